// cpu_baselines.cpp -- the host-CPU sorts BASELINE.md section 3 asks to time beside the library:
// std::sort (1 core, the report's "CPU" column) and __gnu_parallel::sort (all cores).
//   g++ -O3 -fopenmp tools/cpu_baselines.cpp -o build/cpu_baselines ; build/cpu_baselines 24 28
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <parallel/algorithm>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    printf("host threads: %u\n", std::thread::hardware_concurrency());
    for (int a = 1; a < argc; ++a) {
        const size_t n = (size_t)1 << atoi(argv[a]);
        std::vector<int> keys(n), work;
        uint64_t s = 0x9E3779B97F4A7C15ull;
        for (size_t i = 0; i < n; ++i) {                 // splitmix64, seed 1: same recipe as datagen.uniform
            uint64_t z = (s += 0x9E3779B97F4A7C15ull);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            keys[i] = (int)(uint32_t)((z ^ (z >> 31)) >> 32);
        }
        work = keys;
        double t = now();
        std::sort(work.begin(), work.end());
        const double t1 = now() - t;
        work = keys;
        t = now();
        __gnu_parallel::sort(work.begin(), work.end());
        const double tp = now() - t;
        printf("n=2^%s uniform int32: std::sort (1 core) %.3f s = %.1f Mkeys/s | __gnu_parallel::sort (%u threads) %.3f s = %.1f Mkeys/s | sorted %d\n",
               argv[a], t1, n / t1 / 1e6, std::thread::hardware_concurrency(), tp, n / tp / 1e6,
               (int)std::is_sorted(work.begin(), work.end()));
    }
    return 0;
}

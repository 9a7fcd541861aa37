// b200sort_driver.cpp -- checked, parameterised counterpart of the lab's two drivers
// (SRM/main.cpp:17-51 and SRM/performanceTest.cpp:22-53, whose own call is a ToDo at :41).
//
// Same shape as the reference's main: a size sweep, a warm-up sweep first, one timed call per size
// with the MS macro around the host-array operator (so times INCLUDE H2D + D2H, like the report's
// table), CSV rows "Size,Time,Algorithm".  What the reference lacks and this adds: argv, a seeded
// generator, every distribution BASELINE.json names, and a memcmp against std::sort of the same
// input (--check).  The check uses the host's std::sort and is not part of the library.
//
//   b200sort_driver [--min N] [--max N] [--dist NAME] [--seed S] [--check] [--csv FILE]
//   dist: rand100 (SRM/main.cpp:10) | rand1000 (SRM/performanceTest.cpp:35) | uniform | nonneg |
//         and3 | mask16 | skewed | ascending | descending | equal
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lab.h"

static uint64_t splitmix(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void fill(std::vector<int> &v, const std::string &dist, uint64_t seed) {
    uint64_t s = seed * 0x9E3779B97F4A7C15ull;
    const size_t n = v.size();
    for (size_t i = 0; i < n; ++i) {
        const uint32_t r = (uint32_t)(splitmix(s) >> 32);
        if (dist == "rand100") v[i] = rand() % 100;
        else if (dist == "rand1000") v[i] = rand() % 1000;
        else if (dist == "uniform") v[i] = (int)r;
        else if (dist == "nonneg") v[i] = (int)(r >> 1);
        else if (dist == "and3") v[i] = (int)(r & (uint32_t)(splitmix(s) >> 32) & (uint32_t)(splitmix(s) >> 32));
        else if (dist == "mask16") v[i] = (int)(r & 0x0000FFFFu);
        else if (dist == "skewed") v[i] = (splitmix(s) % 10) ? (int)((r & 0x00FFFFFFu) | 0x40000000u) : (int)r;
        else if (dist == "ascending") v[i] = (int)((long long)i - (long long)(n / 2));
        else if (dist == "descending") v[i] = (int)((long long)(n / 2) - 1 - (long long)i);
        else if (dist == "equal") v[i] = 7;
        else { fprintf(stderr, "unknown --dist %s\n", dist.c_str()); exit(2); }
    }
}

int main(int argc, char **argv) {
    size_t min_n = 256, max_n = 65536;          // the reference's sweep, SRM/main.cpp:24,35
    std::string dist = "rand100", csv = "output.txt";
    uint64_t seed = 1;
    bool check = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--min") min_n = strtoull(next(), nullptr, 0);
        else if (a == "--max") max_n = strtoull(next(), nullptr, 0);
        else if (a == "--dist") dist = next();
        else if (a == "--seed") seed = strtoull(next(), nullptr, 0);
        else if (a == "--csv") csv = next();
        else if (a == "--check") check = true;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    srand((unsigned)seed);
    FILE *out = fopen(csv.c_str(), "w");
    if (!out) { perror(csv.c_str()); return 2; }
    fprintf(out, "Size,Time,Algorithm\n");
    int mismatches = 0;
    for (int sweep = 0; sweep < 2; ++sweep) {          // sweep 0 is the ignored warm-up, SRM/main.cpp:23
        for (size_t n = min_n; n <= max_n; n *= 2) {
            std::vector<int> ours(n), trust(n), want;
            fill(ours, dist, seed + n);
            trust = ours;
            if (check) { want = ours; std::sort(want.begin(), want.end()); }
            MS(order_array(ours.data(), (int)n), our_time)
            MS(order_with_trust(trust.data(), (int)n), trust_time)
            if (sweep == 0) continue;
            fprintf(out, "%zu,%f,Our\n%zu,%f,Trust\n", n, our_time, n, trust_time);
            printf("Size = %zu | Our(radix) = %.3f ms (%.1f Mkeys/s) | Trust(merge) = %.3f ms (%.1f Mkeys/s)",
                   n, our_time, n / our_time / 1e3, trust_time, n / trust_time / 1e3);
            if (check) {
                const bool ok1 = memcmp(ours.data(), want.data(), n * sizeof(int)) == 0;
                const bool ok2 = memcmp(trust.data(), want.data(), n * sizeof(int)) == 0;
                printf(" | check %s", (ok1 && ok2) ? "ok" : "MISMATCH");
                mismatches += !(ok1 && ok2);
            }
            printf("\n");
        }
    }
    fclose(out);
    return mismatches ? 1 : 0;
}

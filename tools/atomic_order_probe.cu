// Probe: are same-address shared-memory atomics issued by ONE warp instruction resolved in lane
// order on this GPU?  (Undocumented; the product does not rely on it unless its self-test passes.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(const uint32_t *idx, uint32_t *ret, int steps, int bins) {
    extern __shared__ uint32_t s[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *t = s + warp * bins;
    for (int j = lane; j < bins; j += 32) t[j] = 0;
    __syncwarp();
    const size_t base = ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * steps * 32;
    for (int i = 0; i < steps; ++i) {
        const uint32_t d = idx[base + i * 32 + lane];
        ret[base + i * 32 + lane] = atomicAdd(t + d, 1u);
    }
}
int main() {
    const int blocks = 296, warps = 16, steps = 64, bins = 256;
    const size_t n = (size_t)blocks * warps * steps * 32;
    uint32_t *h = (uint32_t *)malloc(n * 4), *r = (uint32_t *)malloc(n * 4), *d_idx, *d_ret;
    cudaMalloc(&d_idx, n * 4); cudaMalloc(&d_ret, n * 4);
    long long bad_total = 0, groups = 0;
    for (int mode = 0; mode < 6; ++mode) {
        uint64_t x = 88172645463325252ull + mode;
        for (size_t i = 0; i < n; ++i) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            uint32_t v = (uint32_t)(x >> 33);
            switch (mode) {
                case 0: h[i] = v % 256; break;          // uniform digits
                case 1: h[i] = v % 8; break;            // heavy same-address conflicts
                case 2: h[i] = (v % 8) * 32; break;     // same bank, different addresses + duplicates
                case 3: h[i] = 7; break;                // all lanes one address
                case 4: h[i] = (v % 2) ? 5 : (v >> 8) % 256; break;
                case 5: h[i] = ((v % 4) * 32 + (v >> 4) % 2) ; break;
            }
        }
        cudaMemcpy(d_idx, h, n * 4, cudaMemcpyHostToDevice);
        probe<<<blocks, warps * 32, warps * bins * 4>>>(d_idx, d_ret, steps, bins);
        cudaMemcpy(r, d_ret, n * 4, cudaMemcpyDeviceToHost);
        long long bad = 0;
        for (size_t w = 0; w < n / 32; ++w) {           // one warp instruction
            for (int a = 0; a < 32; ++a)
                for (int b = a + 1; b < 32; ++b)
                    if (h[w * 32 + a] == h[w * 32 + b]) { ++groups; if (r[w * 32 + a] >= r[w * 32 + b]) ++bad; }
        }
        printf("mode %d: %lld out-of-lane-order pairs\n", mode, bad);
        bad_total += bad;
    }
    printf("same-address pairs checked %lld, violations %lld => %s\n", groups, bad_total,
           bad_total ? "NOT lane ordered" : "lane ordered in every case tried");
    return 0;
}

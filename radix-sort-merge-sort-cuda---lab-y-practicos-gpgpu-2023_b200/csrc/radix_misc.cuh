// radix_misc.cuh -- the plan's final copy and the lane-order self-test (included by radix.cu).
#pragma once
#include "radix_tile.cuh"

namespace b200sort {

// The plan's final copy (only ever needed with pass skipping): tmp -> out when an in-place sort
// executed an odd number of passes, in -> out when an out-of-place sort executed none.
__global__ void __launch_bounds__(256)
radix_final_copy_kernel(const int32_t *in_buf, int32_t *out_buf, const int32_t *tmp_buf, size_t n,
                        const RadixControl *ctl)
{
    const uint32_t sel = ctl->final_copy;
    if (sel == 0) return;
    const int32_t *src = (sel == kSelIn) ? in_buf : tmp_buf;
    if (src == out_buf) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nvec = n / 4;
    const int4 *s4 = reinterpret_cast<const int4 *>(src);
    int4 *d4 = reinterpret_cast<int4 *>(out_buf);
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out_buf)) & 15) == 0;
    const size_t start = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (aligned) {
        for (size_t i = start; i < nvec; i += stride) d4[i] = s4[i];
        for (size_t i = nvec * 4 + start; i < n; i += stride) out_buf[i] = src[i];
    } else {
        for (size_t i = start; i < n; i += stride) out_buf[i] = src[i];
    }
}

// Self-test behind kRankAdd: that mode is stable only if same-address shared-memory atomics issued
// by one warp instruction are resolved in lane order.  PTX does not promise that; every B200 tried
// does it (tools/atomic_order_probe.cu).  The library checks it once per process on the device it
// runs on, with conflict patterns from none to 32-way, and falls back to ballots if it ever fails.
__global__ void __launch_bounds__(512)
radix_atomic_order_selftest_kernel(uint32_t *violations)
{
    __shared__ uint32_t table[16][kRadixBins];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *t = table[warp];
    uint32_t x = (blockIdx.x * 512u + threadIdx.x) * 2654435761u + 12345u;
    uint32_t bad = 0;
    for (int round = 0; round < 64; ++round) {
        for (int j = lane; j < kRadixBins; j += 32) t[j] = 0;
        __syncwarp();
        x ^= x << 13; x ^= x >> 17; x ^= x << 5;
        const uint32_t bins = 1u << (round & 7);                 // 1, 2, 4 ... 128 distinct digits
        const uint32_t d = ((x >> 8) % bins) * ((round & 8) ? 32u : 1u) % kRadixBins;   // also same-bank sets
        const uint32_t got = atomicAdd(t + d, 1u);
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t want = __popc(peers & lanemask_lt());
        bad += (got != want);
        __syncwarp();
    }
    if (bad) atomicAdd(violations, bad);
}


}  // namespace b200sort

// radix_async.cuh -- TMA (1-D bulk copy) and mbarrier helpers, and the look-back walk over status rows that a bulk
// load has already brought into shared memory (included by radix_pipelined.cuh; used by both persistent pass kernels).
#pragma once
#include "radix_tile.cuh"

namespace b200sort {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- 1-D bulk copies (TMA) ------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// asks the TMA unit to bring `bytes` (a multiple of 16) from global memory at the 16-byte aligned `src` into L2
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void *gsrc, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint4 lds128(const uint32_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void sts128(uint32_t *p, uint4 v) { *reinterpret_cast<uint4 *>(p) = v; }

// Walk back over status rows like walk_back, the nearest `have` of them already sitting in shared memory in
// memory order (win[(have - d) * 256] = the row at distance d, for my digit); eight loads in flight.  A word
// that was not published yet when it was fetched is polled in global memory.
template <int W>
__device__ __forceinline__ uint32_t walk_back_prefetched(const uint32_t *win, uint32_t have, const uint32_t *first,
                                                         uint32_t max_dist) {
    uint32_t acc = 0;
    for (uint32_t base = 0; base < have; base += 8) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = (base + j < have) ? win[(have - 1 - base - j) * kRadixBins] : 0u;
        if (base + 8 <= have) {
            // fast path: eight rows that are all "published, not inclusive" (flag 01) are added whole -- their eight
            // flags add up to 2^33, i.e. to nothing in 32 bits
            const uint32_t all_and = w[0] & w[1] & w[2] & w[3] & w[4] & w[5] & w[6] & w[7];
            const uint32_t all_or  = w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7];
            if ((all_and >> 30) == 1u && (all_or >> 30) == 1u) {
                acc += (w[0] + w[1]) + (w[2] + w[3]) + ((w[4] + w[5]) + (w[6] + w[7]));
                continue;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (base + j < have) {
                uint32_t x = w[j];
                while ((x & ~kValueMask) == 0) x = ld_relaxed_gpu(first - (size_t)(base + j) * kRadixBins);
                acc += x & kValueMask;
                if ((x & ~kValueMask) == kFlagIncl) return acc;
            }
        }
    }
    if (max_dist > have) acc += walk_back<W>(first - (size_t)have * kRadixBins, max_dist - have);
    return acc;
}

}  // namespace b200sort

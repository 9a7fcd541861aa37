// probe_tma_scatter.cu -- how fast can ONE SM's TMA unit scatter a staged tile in small runs?
//
// Persistent CTAs (2 per SM).  A tile of T keys is brought into shared memory by one bulk load
// (cp.async.bulk.shared::cta.global + mbarrier, double-buffered) and leaves as T*4/S bulk stores of S bytes
// each (cp.async.bulk.global.shared::cta), run j of tile t going to region j, slot t -- the address pattern
// of a radix pass's digit runs.  No load/store-pipe work at all: the time is what the TMA unit needs.
//   edges = 1: every run additionally writes a byte-masked 16-byte chunk before and after (.cp_mask), the
//              head / tail of a run that is not 16-byte aligned in the destination.
//   lsu   = 1: the same scatter by ordinary loads and stores, software-pipelined (for comparison).
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int kThreads = 512;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sdst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_masked(void *gdst, uint32_t ssrc, uint32_t mask) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.cp_mask [%0], [%1], 16, %2;"
                 :: "l"(gdst), "r"(ssrc), "h"((uint16_t)mask) : "memory");
}

// region j holds `tiles` slots of `slot_bytes`; run j of tile t goes to slot t of region j
template <int EDGES>
__global__ void __launch_bounds__(kThreads, 2)
tma_scatter_kernel(const uint32_t *__restrict__ in, unsigned char *out, uint32_t tiles, uint32_t T, uint32_t S,
                   uint32_t issuers, uint32_t *ticket)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[2];
    __shared__ uint32_t s_tile[2];
    const uint32_t tid = threadIdx.x;
    const uint32_t tile_bytes = T * 4;
    const uint32_t runs = tile_bytes / S;
    const uint32_t slot_bytes = S + (EDGES ? 32u : 0u);
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        const uint32_t t = atomicAdd(ticket, 1u);
        s_tile[0] = t;
        if (t < tiles) { mbar_expect_tx(&full[0], tile_bytes); bulk_load(smem_u32(smem), in + (size_t)t * T, tile_bytes, &full[0]); }
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};
    int b = 0;
    for (;;) {
        const uint32_t tile = s_tile[b];
        if (tile >= tiles) break;
        // the stores that read buffer b^1 (previous tile) must have read it before it is refilled
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint32_t t = atomicAdd(ticket, 1u);
            s_tile[b ^ 1] = t;
            if (t < tiles) {
                mbar_expect_tx(&full[b ^ 1], tile_bytes);
                bulk_load(smem_u32(smem) + (b ^ 1) * tile_bytes, in + (size_t)t * T, tile_bytes, &full[b ^ 1]);
            }
        }
        mbar_wait(&full[b], phase[b]);
        phase[b] ^= 1;
        const uint32_t sbase = smem_u32(smem) + b * tile_bytes;
        for (uint32_t j = tid; j < runs; j += issuers) {
            if (tid < issuers) {
                unsigned char *dst = out + ((size_t)j * tiles + tile) * slot_bytes;
                if (EDGES) {
                    bulk_store_masked(dst, sbase + j * S, 0xFF00u);
                    bulk_store(dst + 16, sbase + j * S, S);
                    bulk_store_masked(dst + 16 + S, sbase + j * S, 0x00FFu);
                } else {
                    bulk_store(dst, sbase + j * S, S);
                }
            }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        __syncthreads();          // s_tile[b^1] is visible; everybody has issued
        b ^= 1;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// the same scatter by the load/store pipe: tile -> registers (next tile's loads in flight) -> runs
template <int IPT>
__global__ void __launch_bounds__(kThreads, 2)
lsu_scatter_kernel(const uint32_t *__restrict__ in, uint32_t *out, uint32_t tiles, uint32_t S, uint32_t *ticket)
{
    constexpr uint32_t T = kThreads * IPT;
    __shared__ uint32_t s_tile[2];
    const uint32_t tid = threadIdx.x;
    const uint32_t run_words = S / 4;
    if (tid == 0) s_tile[0] = atomicAdd(ticket, 1u);
    __syncthreads();
    uint32_t key[IPT];
    uint32_t tile = s_tile[0];
    if (tile < tiles)
#pragma unroll
        for (int i = 0; i < IPT; ++i) key[i] = in[(size_t)tile * T + i * kThreads + tid];
    int b = 0;
    while (tile < tiles) {
        if (tid == 0) s_tile[b ^ 1] = atomicAdd(ticket, 1u);
        uint32_t cur[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) cur[i] = key[i];
        __syncthreads();
        const uint32_t next = s_tile[b ^ 1];
        if (next < tiles)
#pragma unroll
            for (int i = 0; i < IPT; ++i) key[i] = in[(size_t)next * T + i * kThreads + tid];
        uint32_t j = tid / run_words, o = tid % run_words;
        const uint32_t dj = kThreads / run_words, dof = kThreads % run_words;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            out[((size_t)j * tiles + tile) * run_words + o] = cur[i];
            j += dj; o += dof;
            if (o >= run_words) { o -= run_words; ++j; }
        }
        tile = next;
        b ^= 1;
    }
}

int main()
{
    const uint32_t T = 10240;
    const size_t n = (size_t)1 << 28;
    const uint32_t tiles = (uint32_t)(n / T);
    uint32_t *d_in, *d_ticket;
    unsigned char *d_out;
    const size_t out_bytes = (size_t)tiles * T * 4 * 2;          // room for the edge chunks
    CK(cudaMalloc(&d_in, (size_t)tiles * T * 4));
    CK(cudaMalloc(&d_out, out_bytes));
    CK(cudaMalloc(&d_ticket, 4));
    CK(cudaMemset(d_in, 1, (size_t)tiles * T * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t smem = (size_t)2 * T * 4;
    CK(cudaFuncSetAttribute(tma_scatter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(tma_scatter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t sizes[] = {64, 128, 160, 320, 640, 1280, 2560, 40960};
    for (int edges = 0; edges < 2; ++edges)
        for (uint32_t S : sizes)
            for (uint32_t issuers : {32u, 128u, 512u}) {
                if (edges && S > 1280) continue;
                float best = 1e9f;
                for (int r = 0; r < 4; ++r) {
                    CK(cudaMemset(d_ticket, 0, 4));
                    CK(cudaEventRecord(e0));
                    if (edges) tma_scatter_kernel<1><<<296, kThreads, smem>>>(d_in, d_out, tiles, T, S, issuers, d_ticket);
                    else       tma_scatter_kernel<0><<<296, kThreads, smem>>>(d_in, d_out, tiles, T, S, issuers, d_ticket);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    CK(cudaGetLastError());
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (r > 0 && ms < best) best = ms;
                }
                const double ops = (double)tiles * (T * 4 / S) * (edges ? 3 : 1);
                printf("tma  edges=%d S=%5u issuers=%3u  %.4f ms  %.0f GB/s (8 B/key)  %.1f cycles per op and SM at 1.965 GHz\n", edges, S,
                       issuers, best, 8.0 * tiles * T / best * 1e-6, best * 1e-3 * 1.965e9 / (ops / 148));
                fflush(stdout);
            }
    for (uint32_t S : {64u, 128u, 160u, 320u, 640u, 1280u, 40960u}) {
        float best = 1e9f;
        for (int r = 0; r < 4; ++r) {
            CK(cudaMemset(d_ticket, 0, 4));
            CK(cudaEventRecord(e0));
            lsu_scatter_kernel<20><<<296, kThreads>>>(d_in, (uint32_t *)d_out, tiles, S, d_ticket);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0 && ms < best) best = ms;
        }
        printf("lsu  S=%5u  %.4f ms  %.0f GB/s (8 B/key)\n", S, best, 8.0 * tiles * T / best * 1e-6);
        fflush(stdout);
    }
    return 0;
}

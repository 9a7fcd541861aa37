// probe_r02.cu -- round-2 hardware probes behind the design of the onesweep pass.
//
//   probe A  "writeout": the write-out half of a radix pass in isolation.  Tiles arrive already in digit
//            order (coalesced read, conflict-free staging), and are written to their 256 digit runs
//              mode 0  by the load/store pipe, one key per thread per instruction (what round 1 ships);
//              mode 1  by 1-D TMA bulk stores (cp.async.bulk.global.shared::cta): per digit run one
//                      bulk copy for the 16-byte aligned interior and byte-masked 16-byte copies
//                      (.cp_mask) for the unaligned head and tail, issued by the 256 digit threads;
//              mode 2  bulk copy for the interior, head / tail words by ordinary stores;
//              mode 3  mode 1 with the copies spread over all 512 threads (interior | head + tail).
//            The staging is co-aligned with the destination (staged word index == destination word
//            index mod 4), which is what a bulk copy needs.
//   probe B  "tmem": can a tile's keys be parked in tensor memory (tcgen05.st / tcgen05.ld, 32 lanes x
//            32 columns per warp) underneath shared-memory atomics without slowing them down?
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/probe_r02.cu -o build/probe_r02
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int kThreads = 512;
constexpr int kBins = 256;
constexpr int kRegionLog2 = 21;                 // every digit owns 2^21 destination words

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_masked(void *gdst, uint32_t ssrc, uint32_t mask) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.cp_mask [%0], [%1], 16, %2;"
                 :: "l"(gdst), "r"(ssrc), "h"((uint16_t)mask) : "memory");
}

// One tile: T keys, already grouped by digit.  cnt/gdst rows say how long every run is and where it goes.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
writeout_kernel(const uint32_t *__restrict__ in, uint32_t *out, const uint16_t *__restrict__ cnt,
                const uint32_t *__restrict__ gdst, uint32_t tiles, uint32_t T, uint32_t *ticket)
{
    extern __shared__ __align__(128) uint32_t smem[];
    const uint32_t kStage = T + kBins * 8;                      // words per staging buffer (padding included)
    uint32_t *s_keys  = smem;                                   // [2][kStage]
    uint32_t *s_start = smem + 2 * kStage;                      // [256] first input position of the run
    uint32_t *s_pos   = s_start + kBins;                        // [256] first staged word of the run
    uint32_t *s_g     = s_pos + kBins;                          // [256] destination word index
    uint32_t *s_cnt   = s_g + kBins;                            // [256]
    uint32_t *s_misc  = s_cnt + kBins;                          // [0..7] warp sums A, [8..15] warp sums B, [16] ticket

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int b = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_misc[16] = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_misc[16];
        if (tile >= tiles) break;
        // ---- the digit phase of this probe: run starts in the tile and in the staging buffer -------------
        if (tid < kBins) {
            const uint32_t c = cnt[(size_t)tile * kBins + tid];
            const uint32_t g = gdst[(size_t)tile * kBins + tid];
            const uint32_t a = g & 3u;
            const uint32_t padded = (MODE == 0) ? c : ((a + c + 3u) & ~3u);   // whole 16-byte chunks
            uint32_t x = c, y = padded;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x2 = __shfl_up_sync(0xffffffffu, x, o), y2 = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= (uint32_t)o) { x += x2; y += y2; }
            }
            if (lane == 31) { s_misc[warp] = x; s_misc[8 + warp] = y; }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            uint32_t ax = 0, ay = 0;
            for (uint32_t w = 0; w < warp; ++w) { ax += s_misc[w]; ay += s_misc[8 + w]; }
            s_start[tid] = x - c + ax;
            s_pos[tid] = (MODE == 0) ? (x - c + ax) : (y - padded + ay + a);
            s_g[tid] = g;
            s_cnt[tid] = c;
        }
        // the bulk copies that read staging buffer b two tiles ago must be done with it
        if (MODE != 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        // ---- staging (cheap here: the tile is already in digit order) ---------------------------------------
        uint32_t *sk = s_keys + b * kStage;
        for (uint32_t p = tid; p < T; p += kThreads) {
            const uint32_t k = in[(size_t)tile * T + p];
            const uint32_t d = k >> kRegionLog2;
            sk[s_pos[d] + (p - s_start[d])] = k;
        }
        if (MODE != 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        // ---- write-out --------------------------------------------------------------------------------------
        if (MODE == 0) {
            for (uint32_t p = tid; p < T; p += kThreads) {
                const uint32_t k = sk[p];
                const uint32_t d = k >> kRegionLog2;
                out[s_g[d] + (p - s_start[d])] = k;
            }
        } else {
            const bool lead = (MODE == 3) ? true : (tid < kBins);
            const uint32_t d = tid & (kBins - 1);
            const bool do_edges = (MODE == 1 && tid < kBins) || (MODE == 3 && tid >= kBins);
            const bool do_body  = (MODE == 1 || MODE == 2) ? (tid < kBins) : (tid < kBins);
            if (lead) {
                const uint32_t c = s_cnt[d], g = s_g[d], a = g & 3u;
                const uint32_t first = s_pos[d] - a;                       // 16-byte aligned staged chunk
                uint32_t head = (4u - a) & 3u;                             // words up to the next chunk boundary
                if (head > c) head = c;
                const uint32_t body = (c - head) & ~3u;
                const uint32_t tail = c - head - body;
                const uint32_t sbase = smem_u32(sk);
                if (do_edges && head > 0)
                    bulk_store_masked(out + (g - a), sbase + first * 4, ((1u << (4 * head)) - 1u) << (4 * a));
                if (do_body && body > 0)
                    bulk_store(out + g + head, sbase + (s_pos[d] + head) * 4, body * 4);
                if (do_edges && tail > 0)
                    bulk_store_masked(out + g + head + body, sbase + (s_pos[d] + head + body) * 4, (1u << (4 * tail)) - 1u);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (MODE == 2) {
                // head and tail words by ordinary stores: 6 slots per digit
                for (uint32_t q = tid; q < kBins * 6; q += kThreads) {
                    const uint32_t dd = q / 6, sl = q % 6;
                    const uint32_t c = s_cnt[dd], g = s_g[dd], a = g & 3u;
                    uint32_t head = (4u - a) & 3u;
                    if (head > c) head = c;
                    const uint32_t body = (c - head) & ~3u;
                    const uint32_t tail = c - head - body;
                    if (sl < 3) { if (sl < head) out[g + sl] = sk[s_pos[dd] + sl]; }
                    else if (sl - 3 < tail) out[g + head + body + sl - 3] = sk[s_pos[dd] + head + body + sl - 3];
                }
            }
        }
        b ^= 1;
    }
    if (MODE != 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void check_kernel(const uint32_t *out, const uint16_t *cnt, const uint32_t *gdst, uint32_t tiles,
                             unsigned long long *bad)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)tiles * kBins) return;
    const uint32_t c = cnt[i], g = gdst[i];
    unsigned long long e = 0;
    for (uint32_t j = 0; j < c; ++j) e += (out[g + j] != g + j);
    if (e) atomicAdd(bad, e);
}

static void run_writeout(uint32_t T, int reps)
{
    const size_t n = (size_t)1 << 28;
    const uint32_t tiles = (uint32_t)(n / T);
    const size_t nkeys = (size_t)tiles * T;
    std::vector<uint16_t> h_cnt((size_t)tiles * kBins);
    std::vector<uint32_t> h_g((size_t)tiles * kBins);
    std::vector<uint32_t> h_in(nkeys);
    std::mt19937 rng(12345);
    std::vector<uint32_t> fill(kBins, 0);
    for (uint32_t t = 0; t < tiles; ++t) {
        uint32_t c[kBins] = {0};
        for (uint32_t j = 0; j < T; ++j) c[rng() & 255u]++;          // multinomial, like uniform keys
        size_t p = (size_t)t * T;
        for (int d = 0; d < kBins; ++d) {
            h_cnt[(size_t)t * kBins + d] = (uint16_t)c[d];
            const uint32_t g = ((uint32_t)d << kRegionLog2) + fill[d];
            h_g[(size_t)t * kBins + d] = g;
            for (uint32_t j = 0; j < c[d]; ++j) h_in[p++] = g + j;
            fill[d] += c[d];
            if (fill[d] >= (1u << kRegionLog2)) { fprintf(stderr, "region overflow\n"); exit(3); }
        }
    }
    uint32_t *d_in, *d_out, *d_g, *d_ticket;
    uint16_t *d_cnt;
    unsigned long long *d_bad;
    const size_t out_words = (size_t)kBins << kRegionLog2;
    CK(cudaMalloc(&d_in, nkeys * 4));
    CK(cudaMalloc(&d_out, out_words * 4));
    CK(cudaMalloc(&d_g, h_g.size() * 4));
    CK(cudaMalloc(&d_cnt, h_cnt.size() * 2));
    CK(cudaMalloc(&d_ticket, 4));
    CK(cudaMalloc(&d_bad, 8));
    CK(cudaMemcpy(d_in, h_in.data(), nkeys * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_g, h_g.data(), h_g.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_cnt, h_cnt.data(), h_cnt.size() * 2, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(2 * (T + kBins * 8) + 4 * kBins + 32) * 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 4; ++mode) {
        void (*fn)(const uint32_t *, uint32_t *, const uint16_t *, const uint32_t *, uint32_t, uint32_t, uint32_t *) =
            mode == 0 ? writeout_kernel<0> : mode == 1 ? writeout_kernel<1> : mode == 2 ? writeout_kernel<2> : writeout_kernel<3>;
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float best = 1e9f, sum = 0;
        for (int r = 0; r < reps + 1; ++r) {
            CK(cudaMemset(d_ticket, 0, 4));
            if (r == 0) CK(cudaMemset(d_out, 0xFF, out_words * 4));
            CK(cudaEventRecord(e0));
            fn<<<296, kThreads, smem>>>(d_in, d_out, d_cnt, d_g, tiles, T, d_ticket);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0) { sum += ms; if (ms < best) best = ms; }
        }
        CK(cudaMemset(d_bad, 0, 8));
        check_kernel<<<(unsigned)(((size_t)tiles * kBins + 255) / 256), 256>>>(d_out, d_cnt, d_g, tiles, d_bad);
        unsigned long long bad = 0;
        CK(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
        printf("writeout T=%u mode=%d  avg %.4f ms  best %.4f ms  (%.0f GB/s of 8 B/key)  wrong words: %llu\n", T, mode,
               sum / reps, best, 8.0 * nkeys / (sum / reps) * 1e-6, bad);
        fflush(stdout);
    }
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_g); cudaFree(d_cnt); cudaFree(d_ticket); cudaFree(d_bad);
}

// ---------------------------------------------------------------------------------------------------------
// probe B: tensor memory as a parking place for keys
// ---------------------------------------------------------------------------------------------------------
#define TMEM_ST32(taddr, r)                                                                                   \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16," \
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                           \
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),     \
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),           \
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),         \
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")
#define TMEM_LD32(taddr, r)                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                    \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),          \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),    \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),  \
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])   \
                 : "r"(taddr) : "memory")

// mode bit 0: shared-memory atomics (20 per thread and iteration, random counters, like the rank phase)
// mode bit 1: park 32 words per thread in tensor memory and fetch them back
template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
tmem_kernel(uint32_t *sink, int iters, unsigned long long *cycles)
{
    __shared__ uint32_t s_table[16 * 256];
    __shared__ uint32_t s_taddr;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < 16 * 256; i += kThreads) s_table[i] = 0;
    if (MODE & 2) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(&s_taddr)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (MODE & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // lanes 32*(warp%4).., columns 32*(warp/4)..
    const uint32_t taddr = (MODE & 2) ? (s_taddr + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 32u) : 0u;
    uint32_t *wt = s_table + warp * 256;
    uint32_t x = tid * 2654435761u + blockIdx.x * 40503u + 1u;
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = x + j;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE & 1) {
#pragma unroll
            for (int j = 0; j < 20; ++j) {
                x ^= x << 13; x ^= x >> 17; x ^= x << 5;
                acc += atomicAdd(wt + (x & 255u), 1u);
            }
        }
        if (MODE & 2) {
            TMEM_ST32(taddr, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            TMEM_LD32(taddr, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] += 1u;
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += r[j];
    sink[blockIdx.x * kThreads + tid] = acc;
    if (tid == 0) atomicMax(cycles, (unsigned long long)(t1 - t0));
    __syncthreads();
    if ((MODE & 2) && warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(s_taddr) : "memory");
}

static void run_tmem(int iters)
{
    uint32_t *d_sink;
    unsigned long long *d_cyc;
    CK(cudaMalloc(&d_sink, 296 * kThreads * 4));
    CK(cudaMalloc(&d_cyc, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 1; mode <= 3; ++mode) {
        void (*fn)(uint32_t *, int, unsigned long long *) = mode == 1 ? tmem_kernel<1> : mode == 2 ? tmem_kernel<2> : tmem_kernel<3>;
        for (int r = 0; r < 2; ++r) {
            CK(cudaMemset(d_cyc, 0, 8));
            CK(cudaEventRecord(e0));
            fn<<<296, kThreads>>>(d_sink, iters, d_cyc);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            unsigned long long cyc = 0;
            CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
            if (r == 1)
                printf("tmem mode=%d (%s%s)  %.4f ms  %.1f cycles per iteration and CTA (2 CTAs/SM; an iteration = 640 atomics/warp-row "
                       "and/or 64 KB parked + fetched per CTA)\n", mode, (mode & 1) ? "atomics " : "", (mode & 2) ? "tmem" : "",
                       ms, (double)cyc / iters);
            fflush(stdout);
        }
    }
    // sanity of the round trip: values must come back incremented exactly `iters` times
    cudaFree(d_sink); cudaFree(d_cyc);
}

int main(int argc, char **argv)
{
    const char *what = argc > 1 ? argv[1] : "all";
    if (!strcmp(what, "writeout") || !strcmp(what, "all")) {
        const uint32_t Ts[] = {10240, 8192, 12288, 6144};
        for (uint32_t T : Ts) run_writeout(T, 5);
    }
    if (!strcmp(what, "tmem") || !strcmp(what, "all")) run_tmem(2000);
    return 0;
}

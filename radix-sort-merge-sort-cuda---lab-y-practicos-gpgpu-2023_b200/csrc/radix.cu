// radix.cu -- onesweep LSD radix sort of int32 keys for sm_100a.
//
// Replaces the lab's radix stage (SRM/lab.cu:11-87: exlusiveScan + radix_sort_kernel, one bit per
// iteration on 32-key warp tiles) with a full 4-pass, 8-bit-digit least-significant-digit sort:
//
//   k1  radix_histogram_kernel   ONE read of the keys (128-bit loads) builds all four digit
//                                histograms; its last block turns them into exclusive bases,
//                                decides which passes are skippable and zeroes the first
//                                tile-status buffer.                         4 B/key
//   k2  radix_onesweep_kernel x4 per tile: warp-multisplit ranking (__match_any_sync), block digit
//                                offsets, decoupled look-back over per-tile digit counts (the
//                                chained scan that replaces the lab's separate scan), keys staged
//                                in shared memory in digit order, coalesced scatter.   8 B/key
//
// Signed order: digits are taken from key ^ 0x80000000 (only the top digit changes).
// Stability of every pass is what makes LSD correct: inside a tile the order is (warp, item,
// lane) and keys are loaded warp-striped so that this is memory order.
#include "radix.cuh"

#include <atomic>

namespace b200sort {

// ================================================================================================
// k1: digit histograms
// ================================================================================================
constexpr int kHistThreads = 512;
constexpr int kHistRepl    = 8;                  // replicated counters: lane % 8 picks a replica
constexpr int kHistRow     = kRadixBins + 1;     // 257 words: replica r sits r banks further on
constexpr int kHistUnroll  = 4;                  // 128-bit loads in flight per thread
constexpr int kHistBlocksPerSM = 3;

__device__ __forceinline__ void hist_add(uint32_t *my, int32_t key) {
    const uint32_t k = key_bits(key);
    atomicAdd(my + 0 * kHistRow + (k & 255u), 1u);
    atomicAdd(my + 1 * kHistRow + ((k >> 8) & 255u), 1u);
    atomicAdd(my + 2 * kHistRow + ((k >> 16) & 255u), 1u);
    atomicAdd(my + 3 * kHistRow + (k >> 24), 1u);
}

__global__ void __launch_bounds__(kHistThreads)
radix_histogram_kernel(const int32_t *__restrict__ keys, size_t n, RadixControl *ctl,
                       uint32_t *status_to_zero, size_t status_words, uint32_t skip_enabled,
                       uint32_t in_place)
{
    __shared__ uint32_t sh[kHistRepl * kRadixPasses * kHistRow];
    __shared__ uint32_t s_warp_sums[kRadixBins / 32];
    __shared__ uint32_t s_skip[kRadixPasses];
    __shared__ uint32_t s_is_last;

    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < kHistRepl * kRadixPasses * kHistRow; i += kHistThreads) sh[i] = 0;

    // Zero the tile-status buffer the first pass will use.
    if (status_to_zero != nullptr) {
        uint4 *z = reinterpret_cast<uint4 *>(status_to_zero);
        const size_t nz = status_words / 4;
        for (size_t i = (size_t)blockIdx.x * kHistThreads + tid; i < nz;
             i += (size_t)gridDim.x * kHistThreads)
            z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    uint32_t *my = sh + (tid % kHistRepl) * (kRadixPasses * kHistRow);

    // Scalar head up to 16-byte alignment, 128-bit body, scalar tail.
    size_t head = ((16 - (reinterpret_cast<uintptr_t>(keys) & 15)) & 15) / 4;
    if (head > n) head = n;
    const size_t nvec = (n - head) / 4;
    const size_t tail_start = head + nvec * 4;
    const int4 *v = reinterpret_cast<const int4 *>(keys + head);

    constexpr size_t kChunk = (size_t)kHistThreads * kHistUnroll;
    for (size_t base = (size_t)blockIdx.x * kChunk; base < nvec; base += (size_t)gridDim.x * kChunk) {
        int4 r[kHistUnroll];
        bool ok[kHistUnroll];
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
            const size_t idx = base + (size_t)u * kHistThreads + tid;
            ok[u] = idx < nvec;
            if (ok[u]) r[u] = ld_stream_v4(v + idx);
        }
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
            if (ok[u]) {
                hist_add(my, r[u].x); hist_add(my, r[u].y);
                hist_add(my, r[u].z); hist_add(my, r[u].w);
            }
        }
    }
    if (blockIdx.x == 0) {
        for (size_t i = tid; i < head; i += kHistThreads) hist_add(my, keys[i]);
        for (size_t i = tail_start + tid; i < n; i += kHistThreads) hist_add(my, keys[i]);
    }
    __syncthreads();

    // Fold the replicas and add into the global histogram.
    for (uint32_t i = tid; i < kRadixPasses * kRadixBins; i += kHistThreads) {
        const uint32_t p = i >> kRadixBits, d = i & (kRadixBins - 1);
        uint32_t sum = 0;
#pragma unroll
        for (int r = 0; r < kHistRepl; ++r) sum += sh[(r * kRadixPasses + p) * kHistRow + d];
        if (sum) atomicAdd(&ctl->hist[p][d], sum);
    }

    // The last block to finish turns counts into exclusive bases.
    __threadfence();
    __syncthreads();
    if (tid == 0) s_is_last = (atomicAdd(&ctl->hist_blocks_done, 1u) == gridDim.x - 1) ? 1u : 0u;
    if (tid < kRadixPasses) s_skip[tid] = 0;
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();

    const uint32_t lane = tid & 31, warp = tid >> 5;
    for (int p = 0; p < kRadixPasses; ++p) {
        const uint32_t c = (tid < kRadixBins) ? __ldcg(&ctl->hist[p][tid]) : 0u;
        uint32_t x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o) x += y;
        }
        if (tid < kRadixBins && lane == 31) s_warp_sums[warp] = x;
        __syncthreads();
        if (tid < kRadixBins) {
            uint32_t add = 0;
            for (uint32_t w = 0; w < warp; ++w) add += s_warp_sums[w];
            ctl->base[p][tid] = x - c + add;
            if (skip_enabled && n > 0 && c == (uint32_t)n) s_skip[p] = 1;
        }
        __syncthreads();
    }
    // The buffer plan: executed pass j reads what pass j-1 wrote (the input for j = 0).
    //   in place     : writes alternate tmp, out, tmp, ...; an odd count leaves the result in tmp
    //                  and the final-copy kernel brings it home;
    //   out of place : writes alternate so that the LAST executed pass lands in out; the input
    //                  is never written.  No executed pass at all (all keys equal): copy in -> out.
    if (tid == 0) {
        uint32_t executed = 0;
        for (int p = 0; p < kRadixPasses; ++p) executed += s_skip[p] ? 0u : 1u;
        uint32_t j = 0, cur = kSelIn;
        for (int p = 0; p < kRadixPasses; ++p) {
            ctl->skip[p] = s_skip[p];
            ctl->src_sel[p] = cur;
            uint32_t dst = cur;
            if (!s_skip[p]) {
                if (in_place) dst = (j % 2 == 0) ? kSelTmp : kSelOut;
                else          dst = ((executed - 1 - j) % 2 == 0) ? kSelOut : kSelTmp;
                ++j;
                cur = dst;
            }
            ctl->dst_sel[p] = dst;
        }
        uint32_t final_copy = 0;
        if (in_place) { if (cur == kSelTmp) final_copy = kSelTmp; }
        else          { if (executed == 0) final_copy = kSelIn; }
        ctl->final_copy = final_copy;
    }
}

// ================================================================================================
// k2: one onesweep pass
// ================================================================================================
constexpr uint32_t kFlagLocal = 1u << 30;   // this tile's own digit count
constexpr uint32_t kFlagIncl  = 2u << 30;   // inclusive count over tiles 0..this
constexpr uint32_t kValueMask = (1u << 30) - 1;

template <int WARPS, int IPT>
struct OnesweepShape {
    static constexpr int kThreads = WARPS * 32;
    static constexpr int kTile    = kThreads * IPT;
    static constexpr size_t kSmemBytes =
        (size_t)WARPS * kRadixBins * 4      // per-warp digit counters -> per-warp offsets
        + (size_t)kTile * 4                 // keys staged in digit order
        + (size_t)kRadixBins * 4 * 2        // tile_start, global offset
        + 64;                               // warp sums, tile id
};

template <int WARPS, int IPT, int MIN_BLOCKS>
__global__ void __launch_bounds__(WARPS * 32, MIN_BLOCKS)
radix_onesweep_kernel(const int32_t *in_buf, int32_t *out_buf, int32_t *tmp_buf, size_t n, int pass,
                      RadixControl *ctl, uint32_t *status_cur, uint32_t *status_next,
                      int follow_plan)
{
    using Shape = OnesweepShape<WARPS, IPT>;
    constexpr int kThreads = Shape::kThreads;
    constexpr int kTile    = Shape::kTile;
    static_assert(WARPS >= kRadixBins / 32, "need one thread per digit");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_warp_hist  = reinterpret_cast<uint32_t *>(smem_raw);              // [WARPS][256]
    int32_t  *s_keys       = reinterpret_cast<int32_t *>(s_warp_hist + WARPS * kRadixBins);
    uint32_t *s_tile_start = reinterpret_cast<uint32_t *>(s_keys + kTile);        // [256]
    uint32_t *s_gofs       = s_tile_start + kRadixBins;                           // [256]
    uint32_t *s_misc       = s_gofs + kRadixBins;                                 // [16]

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // follow_plan: buffers and skipping come from the plan the histogram kernel wrote.
    const int32_t *in = in_buf;
    int32_t *out = out_buf;
    if (follow_plan) {
        if (ctl->skip[pass]) {
            // Identity pass.  Still hand the next pass a clean status buffer.
            if (status_next != nullptr && tid < kRadixBins)
                status_next[(size_t)blockIdx.x * kRadixBins + tid] = 0;
            return;
        }
        const uint32_t ss = ctl->src_sel[pass], ds = ctl->dst_sel[pass];
        in = (ss == kSelIn) ? in_buf : (ss == kSelTmp) ? tmp_buf : out_buf;
        out = (ds == kSelTmp) ? tmp_buf : out_buf;
    }

    // Tiles are handed out by ticket so that a tile only ever waits on tiles already running.
    if (tid == 0) s_misc[8] = atomicAdd(&ctl->ticket[pass], 1u);
#pragma unroll
    for (int j = lane; j < kRadixBins; j += 32) s_warp_hist[warp * kRadixBins + j] = 0;
    __syncthreads();
    const uint32_t tile = s_misc[8];
    const size_t tile_base = (size_t)tile * kTile;
    const uint32_t valid = (n - tile_base < (size_t)kTile) ? (uint32_t)(n - tile_base) : (uint32_t)kTile;
    const int shift = pass * kRadixBits;

    // ---- load, warp-striped: item i of lane l is key warp*32*IPT + i*32 + l of the tile ---------
    int32_t key[IPT];
    {
        const uint32_t wofs = warp * (32 * IPT) + lane;
        const int32_t *src = in + tile_base + wofs;
        if (valid == (uint32_t)kTile) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) key[i] = ld_stream(src + i * 32);
        } else {
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                key[i] = (wofs + i * 32 < valid) ? ld_stream(src + i * 32) : 0x7FFFFFFF;  // sorts last
        }
    }

    // ---- rank inside the warp: peers with my digit, in lane order ----------------------------
    uint32_t rank[IPT];
    {
        uint32_t *wh = s_warp_hist + warp * kRadixBins;
        const uint32_t lt = lanemask_lt();
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint32_t d = (key_bits(key[i]) >> shift) & (kRadixBins - 1);
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t leader = __ffs(peers) - 1;
            uint32_t before = 0;
            if (lane == leader) {
                before = wh[d];
                wh[d] = before + __popc(peers);
            }
            before = __shfl_sync(0xffffffffu, before, leader);
            rank[i] = before + __popc(peers & lt);
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- per digit: warp counts -> warp offsets, tile total; publish; scan over digits ---------
    uint32_t total = 0;
    if (tid < kRadixBins) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = s_warp_hist[w * kRadixBins + tid];
            s_warp_hist[w * kRadixBins + tid] = total;
            total += c;
        }
        uint32_t *slot = status_cur + (size_t)tile * kRadixBins + tid;
        st_relaxed_gpu(slot, (tile == 0 ? kFlagIncl : kFlagLocal) | total);
        if (status_next != nullptr) status_next[(size_t)tile * kRadixBins + tid] = 0;
    }
    {
        uint32_t x = total;   // threads >= 256 carry 0
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o) x += y;
        }
        if (tid < kRadixBins && lane == 31) s_misc[warp] = x;
        __syncthreads();
        if (tid < kRadixBins) {
            uint32_t add = 0;
#pragma unroll
            for (int w = 0; w < kRadixBins / 32; ++w) add += (w < (int)warp) ? s_misc[w] : 0u;
            s_tile_start[tid] = x - total + add;
        }
    }
    __syncthreads();

    // ---- stage the keys in shared memory in digit order ---------------------------------------------
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const uint32_t d = (key_bits(key[i]) >> shift) & (kRadixBins - 1);
        s_keys[s_tile_start[d] + s_warp_hist[warp * kRadixBins + d] + rank[i]] = key[i];
    }

    // ---- decoupled look-back: one thread per digit walks the predecessors' status words ------------
    if (tid < kRadixBins) {
        uint32_t prev = 0;
        if (tile > 0) {
            const uint32_t *p = status_cur + (size_t)(tile - 1) * kRadixBins + tid;
            for (;;) {
                const uint32_t s = ld_relaxed_gpu(p);
                const uint32_t f = s & ~kValueMask;
                if (f == 0) continue;               // predecessor has not published yet
                prev += s & kValueMask;
                if (f == kFlagIncl) break;
                p -= kRadixBins;                    // local count only: keep walking back
            }
            st_relaxed_gpu(status_cur + (size_t)tile * kRadixBins + tid,
                           kFlagIncl | ((prev + total) & kValueMask));
        }
        s_gofs[tid] = ctl->base[pass][tid] + prev - s_tile_start[tid];
    }
    __syncthreads();

    // ---- scatter: consecutive threads write consecutive addresses inside each digit run -----------
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t p = tid + j * kThreads;
        if (p < valid) {
            const int32_t k = s_keys[p];
            const uint32_t d = (key_bits(k) >> shift) & (kRadixBins - 1);
            st_stream(out + (size_t)(uint32_t)(s_gofs[d] + p), k);
        }
    }
}

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

// ================================================================================================
// host side
// ================================================================================================
namespace {

using OnesweepFn = void (*)(const int32_t *, int32_t *, int32_t *, size_t, int, RadixControl *,
                            uint32_t *, uint32_t *, int);

struct Variant {
    const char *name;
    int threads;
    int tile;
    size_t smem;
    OnesweepFn fn;
};

#define B200_VARIANT(W, I, B)                                                            \
    { "warps" #W "_ipt" #I "_occ" #B, OnesweepShape<W, I>::kThreads, OnesweepShape<W, I>::kTile, \
      OnesweepShape<W, I>::kSmemBytes, radix_onesweep_kernel<W, I, B> }

const Variant kVariants[] = {
    B200_VARIANT(16, 16, 2),   // 8192-key tiles, 2 CTAs/SM                (default)
    B200_VARIANT(8, 16, 4),    // 4096-key tiles, 4 CTAs/SM
    B200_VARIANT(16, 12, 2),   // 6144
    B200_VARIANT(8, 24, 3),    // 6144, fewer threads
    B200_VARIANT(12, 16, 3),   // 6144, 384 threads
    B200_VARIANT(16, 20, 2),   // 10240
    B200_VARIANT(8, 8, 6),     // 2048 (small-n friendly)
    B200_VARIANT(16, 24, 1),   // 12288, 1 CTA/SM, most registers
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

std::atomic<int> g_variant{0};
std::atomic<int> g_skip_enabled{1};
std::atomic<bool> g_attrs_set[kNumVariants];

int ensure_smem_attr(int v) {
    if (!g_attrs_set[v].load(std::memory_order_acquire)) {
        B200_CUDA_TRY(cudaFuncSetAttribute(reinterpret_cast<const void *>(kVariants[v].fn),
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)kVariants[v].smem));
        g_attrs_set[v].store(true, std::memory_order_release);
    }
    return B200SORT_OK;
}

int hist_grid(size_t n) {
    const size_t chunks = div_up(div_up(n, 4), (size_t)kHistThreads * kHistUnroll);
    size_t g = chunks < (size_t)kNumSMs * kHistBlocksPerSM ? chunks : (size_t)kNumSMs * kHistBlocksPerSM;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace

int radix_num_variants() { return kNumVariants; }
const char *radix_variant_name(int v) { return (v >= 0 && v < kNumVariants) ? kVariants[v].name : nullptr; }
int radix_set_variant(int v) {
    if (v < 0 || v >= kNumVariants) return B200SORT_ERR_INVALID;
    g_variant.store(v);
    return B200SORT_OK;
}
void radix_set_skip(int enabled) { g_skip_enabled.store(enabled ? 1 : 0); }
size_t radix_current_tile() { return (size_t)kVariants[g_variant.load()].tile; }

size_t radix_workspace_bytes(size_t n) {
    const size_t tiles = div_up(n > 0 ? n : 1, kRadixMinTile);
    return kRadixControlBytes + 2 * tiles * kRadixBins * sizeof(uint32_t);
}

int radix_histogram(const int32_t *d_keys, size_t n, uint32_t *d_hist, cudaStream_t s) {
    // Standalone histogram (unit tests, per-kernel timing): d_hist doubles as the control block's
    // histogram area, so a scratch control block is not needed -- the kernel is given a control
    // block that lives in a small static device allocation.
    static thread_local RadixControl *scratch = nullptr;
    if (scratch == nullptr) B200_CUDA_TRY(cudaMalloc(&scratch, kRadixControlBytes));
    B200_CUDA_TRY(cudaMemsetAsync(scratch, 0, kRadixZeroBytes, s));
    if (n > 0) {
        radix_histogram_kernel<<<hist_grid(n), kHistThreads, 0, s>>>(d_keys, n, scratch, nullptr, 0, 0, 0);
        B200_LAUNCH_CHECK();
    }
    B200_CUDA_TRY(cudaMemcpyAsync(d_hist, scratch->hist, sizeof(uint32_t) * kRadixPasses * kRadixBins,
                                  cudaMemcpyDeviceToDevice, s));
    return B200SORT_OK;
}

static int check_ws(void *d_ws, size_t ws_bytes, size_t n) {
    if (d_ws == nullptr || (reinterpret_cast<uintptr_t>(d_ws) & 255) != 0) return B200SORT_ERR_WORKSPACE;
    if (ws_bytes < radix_workspace_bytes(n)) return B200SORT_ERR_WORKSPACE;
    return B200SORT_OK;
}

int radix_single_pass(const int32_t *d_in, int32_t *d_out, size_t n, int pass, void *d_ws,
                      size_t ws_bytes, cudaStream_t s) {
    if (pass < 0 || pass >= kRadixPasses) return B200SORT_ERR_INVALID;
    if (n == 0) return B200SORT_OK;
    B200_TRY(check_ws(d_ws, ws_bytes, n));
    const int v = g_variant.load();
    B200_TRY(ensure_smem_attr(v));
    const Variant &var = kVariants[v];
    auto *ctl = static_cast<RadixControl *>(d_ws);
    auto *status0 = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(d_ws) + kRadixControlBytes);
    const size_t tiles = div_up(n, (size_t)var.tile);
    B200_CUDA_TRY(cudaMemsetAsync(ctl, 0, kRadixZeroBytes, s));
    radix_histogram_kernel<<<hist_grid(n), kHistThreads, 0, s>>>(d_in, n, ctl, status0,
                                                                 tiles * kRadixBins, 0, 0);
    B200_LAUNCH_CHECK();
    var.fn<<<(unsigned)tiles, var.threads, var.smem, s>>>(d_in, d_out, nullptr, n, pass, ctl, status0,
                                                          nullptr, 0);
    B200_LAUNCH_CHECK();
    return B200SORT_OK;
}

namespace {

struct StepTimer {
    cudaStream_t s;
    float *ms;
    cudaEvent_t ev[8];
    int n = 0;
    int begin() {
        if (!ms) return B200SORT_OK;
        for (auto &e : ev) B200_CUDA_TRY(cudaEventCreate(&e));
        return mark();
    }
    int mark() {
        if (!ms) return B200SORT_OK;
        B200_CUDA_TRY(cudaEventRecord(ev[n++], s));
        return B200SORT_OK;
    }
    int finish() {
        if (!ms) return B200SORT_OK;
        B200_CUDA_TRY(cudaEventSynchronize(ev[n - 1]));
        for (int i = 0; i + 1 < n; ++i) B200_CUDA_TRY(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
        for (auto &e : ev) cudaEventDestroy(e);
        return B200SORT_OK;
    }
};

int radix_sort_impl(const int32_t *d_in, int32_t *d_out, int32_t *d_tmp, size_t n, void *d_ws,
                    size_t ws_bytes, cudaStream_t s, float *ms) {
    if (ms) for (int i = 0; i < 6; ++i) ms[i] = 0.f;
    if (n == 0) return B200SORT_OK;
    if (n == 1) {
        if (d_in != d_out) B200_CUDA_TRY(cudaMemcpyAsync(d_out, d_in, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
        return B200SORT_OK;
    }
    B200_TRY(check_ws(d_ws, ws_bytes, n));
    const int v = g_variant.load();
    B200_TRY(ensure_smem_attr(v));
    const Variant &var = kVariants[v];
    auto *ctl = static_cast<RadixControl *>(d_ws);
    const size_t tiles = div_up(n, (size_t)var.tile);
    uint32_t *status[2];
    status[0] = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(d_ws) + kRadixControlBytes);
    status[1] = status[0] + tiles * kRadixBins;
    const int skip = g_skip_enabled.load();
    const uint32_t in_place = (d_in == d_out) ? 1u : 0u;
    StepTimer timer{s, ms};

    B200_CUDA_TRY(cudaMemsetAsync(ctl, 0, kRadixZeroBytes, s));
    B200_TRY(timer.begin());
    radix_histogram_kernel<<<hist_grid(n), kHistThreads, 0, s>>>(d_in, n, ctl, status[0], tiles * kRadixBins,
                                                                 (uint32_t)skip, in_place);
    B200_LAUNCH_CHECK();
    B200_TRY(timer.mark());
    for (int pass = 0; pass < kRadixPasses; ++pass) {
        uint32_t *cur = status[pass & 1];
        uint32_t *next = (pass + 1 < kRadixPasses) ? status[(pass + 1) & 1] : nullptr;
        var.fn<<<(unsigned)tiles, var.threads, var.smem, s>>>(d_in, d_out, d_tmp, n, pass, ctl, cur, next, 1);
        B200_LAUNCH_CHECK();
        B200_TRY(timer.mark());
    }
    if (skip) {
        // Only the plan (on the device) knows whether a final copy is needed; the kernel exits at
        // once when it is not.  With skipping off the pass count is always even / lands in out.
        const size_t blocks = div_up(div_up(n, 4), 256);
        const unsigned grid = (unsigned)(blocks < (size_t)kNumSMs * 8 ? blocks : (size_t)kNumSMs * 8);
        radix_final_copy_kernel<<<grid, 256, 0, s>>>(d_in, d_out, d_tmp, n, ctl);
        B200_LAUNCH_CHECK();
    }
    B200_TRY(timer.mark());
    return timer.finish();
}

}  // namespace

int radix_sort(const int32_t *d_in, int32_t *d_out, int32_t *d_tmp, size_t n, void *d_ws,
               size_t ws_bytes, cudaStream_t s) {
    return radix_sort_impl(d_in, d_out, d_tmp, n, d_ws, ws_bytes, s, nullptr);
}

int radix_sort_timed(const int32_t *d_in, int32_t *d_out, int32_t *d_tmp, size_t n, void *d_ws,
                     size_t ws_bytes, cudaStream_t s, float *ms) {
    return radix_sort_impl(d_in, d_out, d_tmp, n, d_ws, ws_bytes, s, ms);
}

}  // namespace b200sort

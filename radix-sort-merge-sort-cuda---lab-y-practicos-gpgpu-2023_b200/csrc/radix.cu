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

#include <cooperative_groups.h>

#include <atomic>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace b200sort {

// ================================================================================================
// k1: digit histograms
// ================================================================================================
constexpr int kHistThreads = 512;
constexpr int kHistUnroll  = 4;                  // 128-bit loads in flight per thread
constexpr int kHistBlocksPerSM = 3;
// Shared-memory counters are 16-bit and LANE-PRIVATE: counter (place p, digit d, lane l) lives in
// half (l & 1) of word (p*256 + d)*16 + (l >> 1).  The bank is 16*(d & 1) + (l >> 1), so the only
// lanes that can ever collide in one atomic instruction are the two lanes of a pair -- at most
// 2 wavefronts whatever the key distribution (32 random words on 32 banks cost ~3.5, and an
// all-equal input would serialise 32 ways).  16-bit counters overflow after 65535 hits, so the
// block flushes to the global histogram every kHistFlushIters iterations (<= 32768 hits each).
constexpr int kHistSmemWords  = kRadixPasses * kRadixBins * 16;                  // 64 KiB
constexpr size_t kHistSmemBytes = (size_t)kHistSmemWords * 4;
constexpr int kHistFlushIters = 128;   // 128 iters * (4 keys * 4 loads) * 16 warps = 32768 per lane column

__device__ __forceinline__ void hist_add(uint32_t *col, uint32_t one, int32_t key) {
    const uint32_t k = key_bits(key);
    atomicAdd(col + (0 * kRadixBins + (k & 255u)) * 16, one);
    atomicAdd(col + (1 * kRadixBins + ((k >> 8) & 255u)) * 16, one);
    atomicAdd(col + (2 * kRadixBins + ((k >> 16) & 255u)) * 16, one);
    atomicAdd(col + (3 * kRadixBins + (k >> 24)) * 16, one);
}

// Sum the 32 lane columns of every (place, digit), add into the global histogram, clear.
__device__ __forceinline__ void hist_flush(uint32_t *sh, RadixControl *ctl, uint32_t tid) {
    __syncthreads();
    for (uint32_t i = tid; i < kRadixPasses * kRadixBins; i += kHistThreads) {
        uint32_t sum = 0;
#pragma unroll
        for (uint32_t w = 0; w < 16; ++w) {
            const uint32_t idx = i * 16 + ((w + (i >> 1)) & 15);   // rotate: conflict-free across threads
            const uint32_t v = sh[idx];
            sh[idx] = 0;
            sum += (v & 0xffffu) + (v >> 16);
        }
        if (sum) atomicAdd(&ctl->hist[i >> kRadixBits][i & (kRadixBins - 1)], sum);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kHistThreads)
radix_histogram_kernel(const int32_t *__restrict__ keys, size_t n, RadixControl *ctl,
                       uint32_t *status_to_zero, size_t status_words, uint32_t skip_enabled,
                       uint32_t in_place)
{
    extern __shared__ __align__(16) uint32_t sh[];
    __shared__ uint32_t s_warp_sums[kRadixBins / 32];
    __shared__ uint32_t s_skip[kRadixPasses];
    __shared__ uint32_t s_hot[kRadixPasses];
    __shared__ uint32_t s_is_last;

    const uint32_t tid = threadIdx.x;
    {
        uint4 *z = reinterpret_cast<uint4 *>(sh);
        for (uint32_t i = tid; i < kHistSmemWords / 4; i += kHistThreads) z[i] = make_uint4(0, 0, 0, 0);
    }

    // Zero the tile-status buffer the first pass will use.
    if (status_to_zero != nullptr) {
        uint4 *z = reinterpret_cast<uint4 *>(status_to_zero);
        const size_t nz = status_words / 4;
        for (size_t i = (size_t)blockIdx.x * kHistThreads + tid; i < nz;
             i += (size_t)gridDim.x * kHistThreads)
            z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    uint32_t *col = sh + ((tid & 31) >> 1);
    const uint32_t one = 1u << (16 * (tid & 1));

    // Scalar head up to 16-byte alignment, 128-bit body, scalar tail.
    size_t head = ((16 - (reinterpret_cast<uintptr_t>(keys) & 15)) & 15) / 4;
    if (head > n) head = n;
    const size_t nvec = (n - head) / 4;
    const size_t tail_start = head + nvec * 4;
    const int4 *v = reinterpret_cast<const int4 *>(keys + head);

    constexpr size_t kChunk = (size_t)kHistThreads * kHistUnroll;
    int iters = 0;
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
                hist_add(col, one, r[u].x); hist_add(col, one, r[u].y);
                hist_add(col, one, r[u].z); hist_add(col, one, r[u].w);
            }
        }
        if (++iters == kHistFlushIters) { hist_flush(sh, ctl, tid); iters = 0; }
    }
    if (blockIdx.x == 0) {   // < 8 keys in total: cannot overflow anything
        for (size_t i = tid; i < head; i += kHistThreads) hist_add(col, one, keys[i]);
        for (size_t i = tail_start + tid; i < n; i += kHistThreads) hist_add(col, one, keys[i]);
    }
    hist_flush(sh, ctl, tid);

    // The last block to finish turns counts into exclusive bases.
    __threadfence();
    __syncthreads();
    if (tid == 0) s_is_last = (atomicAdd(&ctl->hist_blocks_done, 1u) == gridDim.x - 1) ? 1u : 0u;
    if (tid < kRadixPasses) { s_skip[tid] = 0; s_hot[tid] = 0; }
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
            if ((size_t)c * 8 > n) s_hot[p] = 1;
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
            ctl->hot[p] = s_hot[p];
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

// How a warp finds, for each of its 32 current keys, the lanes holding the same digit:
//   kRankMatch   __match_any_sync (one MATCH instruction; runs on the ADU pipe)
//   kRankBallot  eight __ballot_sync, one per digit bit (VOTE + LOP3, no shared memory)
//   kRankAtomic  atomicOr of the lane bit into a per-warp {peer mask, count} table in shared memory
//   kRankAdd     EXPERIMENT: plain atomicAdd, stable only if the hardware resolves same-address
//                lanes of one instruction in lane order (undocumented)
enum RankMode { kRankMatch = 0, kRankBallot = 1, kRankAtomic = 2, kRankAdd = 3 };

template <int WARPS, int IPT, int MODE>
struct OnesweepShape {
    static constexpr int kThreads = WARPS * 32;
    static constexpr int kTile    = kThreads * IPT;
    static constexpr int kTableWords = (MODE == kRankAtomic) ? 2 : 1;   // words per (warp, digit)
    static constexpr size_t kSmemBytes =
        (size_t)WARPS * kRadixBins * 4 * kTableWords   // per-warp digit counters -> offsets
        + (size_t)kTile * 4                            // keys staged in digit order
        + (size_t)kRadixBins * 4 * 4                   // global offset, tile total, tile start, chain prefix
        + 64;                                          // warp sums, tile id
};

// digit of `key` for the pass with this shift; `flip` is 0x80 for the top digit (signed order)
__device__ __forceinline__ uint32_t digit_of(int32_t key, int shift, uint32_t flip) {
    return ((static_cast<uint32_t>(key) >> shift) & (kRadixBins - 1)) ^ flip;
}

// ---- two-level look-back ---------------------------------------------------------------------------
// Measured with the phase probe (tools/phase_timing.py): with one level the look-back takes 4.7 us of
// a 9 us tile lifetime.  The inclusive front can only advance (window / L2 round trip) = 8 / 0.26 us
// = 31 tiles per microsecond, which is exactly the rate the pass ran at: the chain, not the SMs,
// set the speed.  With two levels tiles are grouped kLookGroup at a time and a tile's prefix is
//   (totals of the earlier GROUPS) + (totals of the earlier tiles of ITS group);
// both are walks over rows whose partial values (a tile's own total, a group's own total) do not
// depend on any other walk, so nobody waits for a long serial chain.
constexpr int kLookGroup = 32;

// Walk back over status rows for one digit: the row at distance d (1 <= d <= max_dist) is
// first - (d-1)*256.  Flags: 0 not published (poll again), kFlagLocal partial (keep walking),
// kFlagIncl inclusive (stop).  Rows beyond max_dist count as inclusive zero.
template <int W>
__device__ __forceinline__ uint32_t walk_back(const uint32_t *first, uint32_t max_dist) {
    uint32_t acc = 0, back = 1;
    for (;;) {
        uint32_t win[W];
#pragma unroll
        for (int j = 0; j < W; ++j)
            win[j] = (back + j <= max_dist) ? ld_relaxed_gpu(first - (size_t)(back + j - 1) * kRadixBins) : kFlagIncl;
        bool done = false;
        uint32_t used = 0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if (!done && used == (uint32_t)j) {
                const uint32_t f = win[j] & ~kValueMask;
                if (f != 0) {
                    acc += win[j] & kValueMask;
                    used = j + 1;
                    done = (f == kFlagIncl);
                }
            }
        }
        if (done) return acc;
        back += used;
    }
}

// CL > 1: the CTAs of a thread-block cluster take CL consecutive tiles and act as ONE link of the
// look-back chain: tile totals are exchanged through distributed shared memory, the last CTA of
// the cluster publishes / looks back for all of them and hands the result to its peers.  The
// chain then has CL times fewer links, which is what bounds the pass once ranking is cheap.
__device__ __forceinline__ void cluster_arrive() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Phase-timing probe (TIMING variants only): lane 0 of warp 0 (group A) and of warp 8 (group B)
// stamp clock64() at the phase boundaries into g_phase_dbg[tile][2][16].
__device__ long long *g_phase_dbg = nullptr;
#define B200_STAMP(slot)                                                                  \
    do {                                                                                  \
        if (TIMING && g_phase_dbg != nullptr && lane == 0 && (warp == 0 || warp == 8))    \
            g_phase_dbg[((size_t)dbg_tile * 2 + (warp >> 3)) * 16 + (slot)] = clock64();  \
    } while (0)

// PF > 0: after issuing its own loads a CTA prefetches into L2 the tile PF tickets ahead (the
// tile some CTA will pick up about one CTA-lifetime later), so that tile's loads hit L2.
// BSF: group B stages its keys before consuming the look-back window instead of after.
// TL: two-level look-back (tile rows + group rows, see walk_back below); implies BSF.
template <int WARPS, int IPT, int MIN_BLOCKS, int MODE, int CL, int PF = 0, int BSF = 0, int TIMING = 0, int TL = 0>
__global__ void __launch_bounds__(WARPS * 32, MIN_BLOCKS)
radix_onesweep_kernel(const int32_t *in_buf, int32_t *out_buf, int32_t *tmp_buf, size_t n, int pass,
                      RadixControl *ctl, uint32_t *status_cur, uint32_t *status_next,
                      int follow_plan)
{
    using Shape = OnesweepShape<WARPS, IPT, MODE>;
    constexpr int kThreads = Shape::kThreads;
    constexpr int kTile    = Shape::kTile;
    constexpr int TW       = Shape::kTableWords;
    static_assert(WARPS >= kRadixBins / 32, "need one thread per digit");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    // [WARPS][256] entries of TW words.  Entry word TW-1 is the running count, later the offset
    // of (warp, digit) inside the staged tile; with kRankAtomic word 0 is the peer mask.
    uint32_t *s_table = reinterpret_cast<uint32_t *>(smem_raw);
    int32_t  *s_keys  = reinterpret_cast<int32_t *>(s_table + WARPS * kRadixBins * TW);
    uint32_t *s_gofs  = reinterpret_cast<uint32_t *>(s_keys + kTile);             // [256]
    uint32_t *s_misc  = s_gofs + kRadixBins;                                      // [16]

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t dbg_tile = blockIdx.x;
    B200_STAMP(0);

    // follow_plan: buffers and skipping come from the plan the histogram kernel wrote.
    const int32_t *in = in_buf;
    int32_t *out = out_buf;
    if (follow_plan) {
        if (ctl->skip[pass]) {
            // Identity pass.  Still hand the next pass a clean status buffer.
            if (status_next != nullptr && tid < kRadixBins && blockIdx.x % CL == 0)
                status_next[(size_t)(blockIdx.x / CL) * kRadixBins + tid] = 0;
            if (TL && status_next != nullptr && tid < kRadixBins && blockIdx.x % kLookGroup == 0)
                status_next[((n + kTile - 1) / kTile + blockIdx.x / kLookGroup) * kRadixBins + tid] = 0;
            return;
        }
        const uint32_t ss = ctl->src_sel[pass], ds = ctl->dst_sel[pass];
        in = (ss == kSelIn) ? in_buf : (ss == kSelTmp) ? tmp_buf : out_buf;
        out = (ds == kSelTmp) ? tmp_buf : out_buf;
    }

    // Tiles are handed out by ticket so that a tile only ever waits on tiles already running.
    uint32_t crank = 0;                                      // my rank inside the cluster
    if (CL > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        crank = cluster.block_rank();
        if (crank == 0 && tid == 0) {
            const uint32_t t = atomicAdd(&ctl->ticket[pass], 1u);
            for (int q = 0; q < CL; ++q) cluster.map_shared_rank(s_misc, q)[8] = t;
        }
    } else {
        if (tid == 0) s_misc[8] = atomicAdd(&ctl->ticket[pass], 1u);
    }
    {
        uint4 *z = reinterpret_cast<uint4 *>(s_table + warp * kRadixBins * TW);
#pragma unroll
        for (int j = lane; j < kRadixBins * TW / 4; j += 32) z[j] = make_uint4(0, 0, 0, 0);
    }
    if (CL > 1) { cluster_arrive(); cluster_wait(); } else __syncthreads();
    const uint32_t link = s_misc[8];                         // my link of the look-back chain
    const uint32_t tile = link * CL + crank;
    B200_STAMP(1);
    const size_t tile_base = (size_t)tile * kTile;
    const uint32_t valid = (tile_base >= n) ? 0u
                         : (n - tile_base < (size_t)kTile) ? (uint32_t)(n - tile_base) : (uint32_t)kTile;
    const int shift = pass * kRadixBits;
    const uint32_t flip = (pass == kRadixPasses - 1) ? 0x80u : 0u;

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
    if (PF > 0) {
        constexpr uint32_t kLines = (uint32_t)kTile * 4 / 128;
        const size_t ahead = ((size_t)tile + PF) * kTile + (size_t)tid * 32;
        if (tid < kLines && ahead + 32 <= n)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(in + ahead));
        if (kLines > (uint32_t)kThreads && tid + kThreads < kLines && ahead + (size_t)kThreads * 32 + 32 <= n)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(in + ahead + (size_t)kThreads * 32));
    }

    if (TIMING) { asm volatile("" :: "r"(key[0]), "r"(key[IPT - 1])); B200_STAMP(2); }   // loads have landed
    // ---- rank inside the warp: earlier keys of this warp with my digit ----------------------------
    // (two 16-bit ranks per register: a warp holds at most 32*IPT < 65536 keys)
    static_assert(IPT % 2 == 0 && 32 * IPT < 65536, "ranks are packed in pairs");
    uint32_t rank2[IPT / 2];
    {
        uint32_t *wt = s_table + warp * kRadixBins * TW;
        const uint32_t lt = lanemask_lt();
        // "hot" = some digit value is frequent: globally (the histogram kernel saw one bin with more
        // than 1/8 of the keys) or in this warp's part of the tile (sorted / clustered input: a
        // quarter of the lanes agree with lane 0 on the first key).  Warp-uniform.
        bool hot = false;
        if (MODE == kRankAdd) {
            const uint32_t d0 = digit_of(key[0], shift, flip);
            const uint32_t agree = __ballot_sync(0xffffffffu, d0 == __shfl_sync(0xffffffffu, d0, 0));
            hot = (follow_plan && ctl->hot[pass] != 0) || __popc(agree) >= 8;
        }
        if (MODE == kRankAdd && !hot) {
            // the common case, kept free of any per-key branch
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const uint32_t r = atomicAdd(wt + digit_of(key[i], shift, flip), 1u);
                rank2[i / 2] = (i & 1) ? (rank2[i / 2] | (r << 16)) : r;
            }
        } else {
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint32_t d = digit_of(key[i], shift, flip);
            if (MODE == kRankAdd) {
                // A digit value is frequent: same-address atomics would serialise.  The lanes that
                // share lane 0's digit are ranked with one ballot and ONE atomic.
                const bool same = (d == __shfl_sync(0xffffffffu, d, 0));
                const uint32_t sm = __ballot_sync(0xffffffffu, same);
                uint32_t r = 0;
                if (!same || lane == 0) r = atomicAdd(wt + d, lane == 0 ? (uint32_t)__popc(sm) : 1u);
                const uint32_t r0 = __shfl_sync(0xffffffffu, r, 0);
                if (same) r = r0 + __popc(sm & lt);
                rank2[i / 2] = (i & 1) ? (rank2[i / 2] | (r << 16)) : r;
            } else if (MODE == kRankAtomic) {
                atomicOr(wt + 2 * d, 1u << lane);
                __syncwarp();
                const uint2 e = *reinterpret_cast<const uint2 *>(wt + 2 * d);   // {peers, count}
                const uint32_t lower = e.x & lt;
                const uint32_t r = e.y + __popc(lower);
                rank2[i / 2] = (i & 1) ? (rank2[i / 2] | (r << 16)) : r;
                __syncwarp();
                if (lower == 0)                                                   // lowest peer
                    *reinterpret_cast<uint2 *>(wt + 2 * d) = make_uint2(0u, e.y + __popc(e.x));
                __syncwarp();
            } else {
                uint32_t peers;
                if (MODE == kRankMatch) {
                    peers = __match_any_sync(0xffffffffu, d);
                } else {
                    peers = 0xffffffffu;
#pragma unroll
                    for (int b = 0; b < kRadixBits; ++b) {
                        const bool bit = (d >> b) & 1u;
                        const uint32_t vote = __ballot_sync(0xffffffffu, bit);
                        peers &= bit ? vote : ~vote;
                    }
                }
                const uint32_t lower = peers & lt;
                uint32_t before = 0;
                if (lower == 0) before = atomicAdd(wt + d, (uint32_t)__popc(peers));   // one lane per digit
                before = __shfl_sync(0xffffffffu, before, __ffs(peers) - 1);
                const uint32_t r = before + __popc(lower);
                rank2[i / 2] = (i & 1) ? (rank2[i / 2] | (r << 16)) : r;
            }
        }
        }
    }
    if (TIMING) { asm volatile("" :: "r"(rank2[0]), "r"(rank2[IPT / 2 - 1])); B200_STAMP(3); }   // ranked
    __syncthreads();
    B200_STAMP(4);

    // ---- per digit, two thread groups working side by side --------------------------------------
    //   group A (threads 0..255, thread = digit): tile totals, exclusive scan over the digits,
    //           warp counts -> positions inside the staged tile;
    //   group B (threads 256..511, thread - 256 = digit; the same threads as A when the CTA has
    //           fewer than 16 warps): publish the tile total, decoupled look-back over the
    //           predecessor tiles with kLookWindow status words in flight per thread, publish the
    //           inclusive count, global offset of the digit.
    // Status words only ever move 0 -> local -> inclusive, so a stale (prefetched) read is safe.
    // Named barriers: 1 = inside group A; 2 = "totals are in shared memory" (A arrives, B waits);
    //                 3 = "positions are final" (A arrives, B waits).
    constexpr bool kSplit = (WARPS >= 16);
    constexpr int kLookWindow = (CL == 1 && IPT <= 16) ? 16 : 8;
    uint32_t *s_total = s_misc + 16;                         // [256]
    uint32_t *s_tstart = s_total + kRadixBins;               // [256]
    uint32_t *s_prev = s_tstart + kRadixBins;                // [256] (clusters: written by the looker)

    const bool in_a = tid < kRadixBins;
    const bool in_b = kSplit ? (tid >= kRadixBins && tid < 2 * kRadixBins) : in_a;
    const uint32_t bd = kSplit ? tid - kRadixBins : tid;     // group B's digit
    const uint32_t *look = status_cur + (size_t)link * kRadixBins + bd;   // my digit in my link's row
    const bool looker = (CL == 1) || (crank == CL - 1);      // the CTA that talks to the chain

    uint32_t win[kLookWindow];
    uint32_t digit_base = 0;                                 // global start of my digit (group B)
    if (in_b) digit_base = ctl->base[pass][bd];              // fetched early: it is off the critical path
    if (kSplit && in_b && looker && !TL) {
        // first window, issued before anything else so that it overlaps group A's work
#pragma unroll
        for (int j = 0; j < kLookWindow; ++j)
            win[j] = (link >= (uint32_t)(j + 1)) ? ld_relaxed_gpu(look - (size_t)(j + 1) * kRadixBins)
                                                 : kFlagIncl;            // before link 0: inclusive 0
    }
    if (in_a) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) total += s_table[(w * kRadixBins + tid) * TW + (TW - 1)];
        s_total[tid] = total;
        if (CL > 1) cluster_arrive();                        // #1: my totals are in shared memory
        if (kSplit) { __threadfence_block(); asm volatile("bar.arrive 2, 512;" ::: "memory"); }
        uint32_t x = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o) x += y;
        }
        if (lane == 31) s_misc[warp] = x;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint32_t add = 0;
#pragma unroll
        for (int w = 0; w < kRadixBins / 32; ++w) add += (w < (int)warp) ? s_misc[w] : 0u;
        const uint32_t tile_start = x - total + add;
        uint32_t run = tile_start;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            uint32_t *e = s_table + (w * kRadixBins + tid) * TW + (TW - 1);
            const uint32_t c = *e;
            *e = run;
            run += c;
        }
        s_tstart[tid] = tile_start;
        if (kSplit) { __threadfence_block(); asm volatile("bar.arrive 3, %0;" :: "n"(WARPS * 32) : "memory"); }
        asm volatile("bar.sync 1, 256;" ::: "memory");      // every (warp, digit) position is final
        if (CL > 1) { cluster_wait(); cluster_arrive(); }    // finish #1; #2: nothing to announce
        B200_STAMP(5);                                       // group A done
    }
    if (CL > 1 && !in_a && !in_b) { cluster_arrive(); cluster_wait(); cluster_arrive(); }
    if (kSplit && !in_a && !in_b) asm volatile("bar.sync 3, %0;" :: "n"(WARPS * 32) : "memory");   // warps 16..: wait for the positions
    if (in_b) {
        if (kSplit) asm volatile("bar.sync 2, 512;" ::: "memory");
        uint32_t total = s_total[bd];                        // my tile; becomes my link's total
        uint32_t before = 0;                                 // same digit in earlier tiles of my link
        if (CL > 1) {
            if (kSplit) cluster_arrive();                    // #1 (group A arrived for itself)
            cluster_wait();                                  // every CTA's totals are readable
            cg::cluster_group cluster = cg::this_cluster();
            uint32_t rest = 0;
#pragma unroll
            for (int q = 0; q < CL; ++q) {
                if (looker ? (q < CL - 1) : (q < (int)crank)) {
                    const uint32_t c = cluster.map_shared_rank(s_total, q)[bd];
                    if (q < (int)crank) before += c;
                    rest += c;
                }
            }
            if (looker) total += rest;
        }
        uint32_t prev = 0;
        if (TL) {
            static_assert(!TL || (CL == 1 && WARPS >= 16), "two-level look-back: split CTAs without clusters");
            const size_t num_tiles = (n + kTile - 1) / kTile;
            const uint32_t group = tile / kLookGroup, r = tile % kLookGroup;
            const bool last_of_group = (r == kLookGroup - 1) || ((size_t)tile + 1 == num_tiles);
            uint32_t *row = status_cur + (size_t)tile * kRadixBins + bd;                   // tile rows ...
            uint32_t *grow = status_cur + (num_tiles + group) * kRadixBins + bd;           // ... then group rows
            st_relaxed_gpu(row, (r == 0 ? kFlagIncl : kFlagLocal) | total);               // inclusive WITHIN the group
            if (status_next != nullptr) {
                status_next[(size_t)tile * kRadixBins + bd] = 0;
                if (last_of_group) status_next[(num_tiles + group) * kRadixBins + bd] = 0;
            }
            // stage my keys now (positions are final once group A says so): that frees their
            // registers for the windows below and overlaps with the predecessors' publishing
            asm volatile("bar.sync 3, %0;" :: "n"(WARPS * 32) : "memory");
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const uint32_t d = digit_of(key[i], shift, flip);
                const uint32_t rk = (i & 1) ? (rank2[i / 2] >> 16) : (rank2[i / 2] & 0xffffu);
                s_keys[s_table[(warp * kRadixBins + d) * TW + (TW - 1)] + rk] = key[i];
            }
            B200_STAMP(10);                                  // staged, walks start
            uint32_t inprev = 0;
            if (r > 0) {
                inprev = walk_back<16>(row - kRadixBins, r);
                st_relaxed_gpu(row, kFlagIncl | (inprev + total));
            }
            B200_STAMP(11);                                  // level 1 done
            const uint32_t gtot = inprev + total;
            if (last_of_group) st_relaxed_gpu(grow, (group == 0 ? kFlagIncl : kFlagLocal) | gtot);
            uint32_t gprev = 0;
            if (group > 0) {
                gprev = walk_back<16>(grow - kRadixBins, group);
                if (last_of_group) st_relaxed_gpu(grow, kFlagIncl | ((gprev + gtot) & kValueMask));
            }
            prev = inprev + gprev;
            B200_STAMP(12);                                  // level 2 done
        }
        if (!TL && looker) {
            st_relaxed_gpu(const_cast<uint32_t *>(look), (link == 0 ? kFlagIncl : kFlagLocal) | total);
            if (status_next != nullptr) status_next[(size_t)link * kRadixBins + bd] = 0;
        }
        if (!TL && BSF && kSplit && CL == 1) {
            // positions are final as soon as group A says so: stage my keys while the prefetched
            // status words are still in flight
            asm volatile("bar.sync 3, %0;" :: "n"(WARPS * 32) : "memory");
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const uint32_t d = digit_of(key[i], shift, flip);
                const uint32_t r = (i & 1) ? (rank2[i / 2] >> 16) : (rank2[i / 2] & 0xffffu);
                s_keys[s_table[(warp * kRadixBins + d) * TW + (TW - 1)] + r] = key[i];
            }
        }
        if (!TL && looker) {
            if (link > 0) {
                uint32_t back = 1;                           // distance of the window's first link
                bool have = kSplit;                          // window already loaded?
                for (;;) {
                    if (!have) {
#pragma unroll
                        for (int j = 0; j < kLookWindow; ++j)
                            win[j] = (link >= back + j) ? ld_relaxed_gpu(look - (size_t)(back + j) * kRadixBins)
                                                        : kFlagIncl;
                    }
                    have = false;
                    bool done = false;
                    uint32_t used = 0;
#pragma unroll
                    for (int j = 0; j < kLookWindow; ++j) {
                        if (!done && used == (uint32_t)j) {
                            const uint32_t f = win[j] & ~kValueMask;
                            if (f != 0) {                    // published: take it
                                prev += win[j] & kValueMask;
                                used = j + 1;
                                done = (f == kFlagIncl);
                            }
                        }
                    }
                    if (done) break;
                    back += used;                            // re-poll from the first unpublished link
                }
                st_relaxed_gpu(const_cast<uint32_t *>(look), kFlagIncl | ((prev + total) & kValueMask));
            }
            if (CL > 1) {                                    // hand the chain prefix to my peers
                cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
                for (int q = 0; q < CL - 1; ++q) cluster.map_shared_rank(s_prev, q)[bd] = prev;
            }
        }
        __syncwarp();                                        // the look-back loop diverges per digit
        if (CL > 1) {
            cluster_arrive();                                // #2: the prefix is in everybody's memory
            cluster_wait();
            if (!looker) prev = s_prev[bd];
        }
        if (kSplit && !((BSF || TL) && CL == 1)) asm volatile("bar.sync 3, %0;" :: "n"(WARPS * 32) : "memory");
        s_gofs[bd] = digit_base + prev + before - s_tstart[bd];
        B200_STAMP(5);                                       // group B done (look-back finished)
    }
    // Positions must be final before anybody stages keys: group A knows (its barrier 1), group B
    // knows (barrier 3); a CTA that is not split simply synchronises.
    static_assert(CL == 1 || WARPS == 16, "clustered shapes are exactly groups A and B");
    if (!kSplit) __syncthreads();

    // ---- stage the keys in shared memory in digit order ---------------------------------------------
    if (!((BSF || TL) && kSplit && CL == 1 && in_b)) {
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint32_t d = digit_of(key[i], shift, flip);
            const uint32_t r = (i & 1) ? (rank2[i / 2] >> 16) : (rank2[i / 2] & 0xffffu);
            s_keys[s_table[(warp * kRadixBins + d) * TW + (TW - 1)] + r] = key[i];
        }
    }
    if (CL > 1 && !in_b) cluster_wait();                     // finish #2 (group B already did)
    B200_STAMP(6);                                           // staged
    __syncthreads();
    B200_STAMP(7);

    // ---- scatter: consecutive threads write consecutive addresses inside each digit run -----------
    if (valid == (uint32_t)kTile) {
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t p = tid + j * kThreads;
            const int32_t k = s_keys[p];
            st_stream(out + (size_t)(uint32_t)(s_gofs[digit_of(k, shift, flip)] + p), k);
        }
    } else {
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t p = tid + j * kThreads;
            if (p < valid) {
                const int32_t k = s_keys[p];
                st_stream(out + (size_t)(uint32_t)(s_gofs[digit_of(k, shift, flip)] + p), k);
            }
        }
    }
    B200_STAMP(8);
    if (TIMING && g_phase_dbg != nullptr && lane == 0 && (warp == 0 || warp == 8))
        g_phase_dbg[((size_t)dbg_tile * 2 + (warp >> 3)) * 16 + 9] = tile;
}

// ================================================================================================
// k2': the same pass as a PERSISTENT, software-pipelined CTA
// ================================================================================================
// One CTA per SM slot loops over tiles (tickets).  14 worker warps load / rank / stage / write the
// keys; 2 chain warps own everything that talks to other tiles (publish the tile's digit counts,
// decoupled look-back with 128-bit status loads, publish the inclusive counts, global offsets).
// The workers never wait for the chain on the tile they are ranking: tile i is written out only
// after tile i+1 has been ranked and staged (double-buffered staging area), and the global loads
// of tile i+1 are in flight while tile i-1 is being written.  So neither the look-back latency
// nor the load latency sits on the workers' critical path.
constexpr int kPPWorkerWarps = 14;
constexpr int kPPWorkers = kPPWorkerWarps * 32;      // 448
constexpr int kPPThreads = 512;
constexpr int kPPChain = kPPThreads - kPPWorkers;    // 64 threads, 4 digits each
constexpr uint32_t kPPPoison = 0xFFFFFFFFu;
constexpr int kPPWindow = 8;                         // status rows in flight per chain thread
enum { kBarW = 1, kBarA = 2, kBarTotals = 3, kBarTstart = 5, kBarGofs = 7 };   // +buffer for the last three

__device__ __forceinline__ void bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_gpu_v4(const uint32_t *p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu_v4(uint32_t *p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Two-level look-back.  Tiles are grouped kPPGroup at a time.  A tile's prefix is
//   (sum of the totals of the earlier GROUPS) + (sum of the totals of the earlier tiles of ITS group).
// Both sums are walks over status rows whose partial values do not depend on any other walk (a
// tile's own total, a group's own total), so no tile waits for a long serial chain: the inclusive
// front only has to advance one GROUP per round trip.  (With one level the front must advance one
// tile per round trip times the window, which is what bounded the pass: ~35 tiles start per
// microsecond and a status round trip through L2 takes ~0.4 us.)
constexpr int kPPGroup = kLookGroup;

// Walk back over status rows: row at distance d (1 <= d <= max_dist) is `first - (d-1)*256`; each
// thread handles four digits with 128-bit loads, W rows in flight.  Flags: 0 not published yet
// (poll again), kFlagLocal partial (keep walking), kFlagIncl inclusive (stop).  Rows beyond
// max_dist count as inclusive zero.  acc[k] += everything taken.
template <int W>
__device__ __forceinline__ void chain_walk(const uint32_t *first, uint32_t max_dist, uint32_t (&acc)[4]) {
    uint32_t need[4] = {1, 1, 1, 1};
    bool done[4] = {false, false, false, false};
    uint32_t back = 1;
    for (;;) {
        uint4 win[W];
#pragma unroll
        for (int j = 0; j < W; ++j)
            win[j] = (back + j <= max_dist) ? ld_relaxed_gpu_v4(first - (size_t)(back + j - 1) * kRadixBins)
                                            : make_uint4(kFlagIncl, kFlagIncl, kFlagIncl, kFlagIncl);
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const uint32_t w4[4] = {win[j].x, win[j].y, win[j].z, win[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!done[k] && need[k] == back + j) {
                    const uint32_t f = w4[k] & ~kValueMask;
                    if (f != 0) {
                        acc[k] += w4[k] & kValueMask;
                        need[k] += 1;
                        done[k] = (f == kFlagIncl);
                    }
                }
            }
        }
        if (done[0] && done[1] && done[2] && done[3]) break;
        uint32_t nb = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (!done[k] && need[k] < nb) nb = need[k];
        back = nb;
    }
}

template <int IPT>
struct PipelinedShape {
    static constexpr int kTile = kPPWorkers * IPT;
    static constexpr size_t kSmemBytes =
        (size_t)kPPWorkerWarps * kRadixBins * 4     // per-warp digit counters -> positions
        + (size_t)2 * kTile * 4                     // two staging buffers
        + (size_t)3 * 2 * kRadixBins * 4            // gofs, total, tstart, double-buffered
        + 128;                                      // warp sums, tickets, tile ids
};

template <int IPT>
__global__ void __launch_bounds__(kPPThreads, 2)
radix_onesweep_pipelined_kernel(const int32_t *in_buf, int32_t *out_buf, int32_t *tmp_buf, size_t n,
                                int pass, RadixControl *ctl, uint32_t *status_cur, uint32_t *status_next,
                                int follow_plan)
{
    constexpr int kTile = PipelinedShape<IPT>::kTile;
    static_assert(IPT % 2 == 0 && 32 * IPT < 65536, "ranks are packed in pairs");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_table  = reinterpret_cast<uint32_t *>(smem_raw);                     // [14][256]
    int32_t  *s_keys   = reinterpret_cast<int32_t *>(s_table + kPPWorkerWarps * kRadixBins);   // [2][kTile]
    uint32_t *s_gofs   = reinterpret_cast<uint32_t *>(s_keys + 2 * kTile);           // [2][256]
    uint32_t *s_total  = s_gofs + 2 * kRadixBins;                                    // [2][256]
    uint32_t *s_tstart = s_total + 2 * kRadixBins;                                   // [2][256]
    uint32_t *s_misc   = s_tstart + 2 * kRadixBins;     // [0..7] warp sums, [8..9] next ticket, [10..11] tile id

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t tiles = (n + kTile - 1) / kTile;

    const int32_t *in = in_buf;
    int32_t *out = out_buf;
    if (follow_plan) {
        if (ctl->skip[pass]) {
            const size_t rows = tiles + (tiles + kPPGroup - 1) / kPPGroup;       // tile rows + group rows
            if (status_next != nullptr)
                for (size_t row = blockIdx.x; row < rows; row += gridDim.x)
                    if (tid < kRadixBins) status_next[row * kRadixBins + tid] = 0;
            return;
        }
        const uint32_t ss = ctl->src_sel[pass], ds = ctl->dst_sel[pass];
        in = (ss == kSelIn) ? in_buf : (ss == kSelTmp) ? tmp_buf : out_buf;
        out = (ds == kSelTmp) ? tmp_buf : out_buf;
    }
    const int shift = pass * kRadixBits;
    const uint32_t flip = (pass == kRadixPasses - 1) ? 0x80u : 0u;

    if (warp >= kPPWorkerWarps) {
        // ======================== chain warps ========================
        const uint32_t c4 = (tid - kPPWorkers) * 4;                 // my four digits
        const uint4 base4 = *reinterpret_cast<const uint4 *>(&ctl->base[pass][c4]);
        int b = 0;
        for (;;) {
            bar_sync(kBarTotals + b, kRadixBins + kPPChain);
            const uint32_t tile = s_misc[10 + b];
            if (tile == kPPPoison) break;
            const uint4 tot = *reinterpret_cast<const uint4 *>(s_total + b * kRadixBins + c4);
            const uint32_t group = tile / kPPGroup, r = tile % kPPGroup;
            const bool last_of_group = (r == kPPGroup - 1) || ((size_t)tile + 1 == tiles);
            uint32_t *row = status_cur + (size_t)tile * kRadixBins + c4;                 // tile rows
            uint32_t *grow = status_cur + (tiles + group) * kRadixBins + c4;             // group rows follow
            const uint32_t flag0 = (r == 0) ? kFlagIncl : kFlagLocal;                    // inclusive WITHIN the group
            st_relaxed_gpu_v4(row, make_uint4(flag0 | tot.x, flag0 | tot.y, flag0 | tot.z, flag0 | tot.w));
            if (status_next != nullptr) {
                *reinterpret_cast<uint4 *>(status_next + (size_t)tile * kRadixBins + c4) = make_uint4(0, 0, 0, 0);
                if (last_of_group)
                    *reinterpret_cast<uint4 *>(status_next + (tiles + group) * kRadixBins + c4) = make_uint4(0, 0, 0, 0);
            }
            // level 1: earlier tiles of my group
            uint32_t prev[4] = {0, 0, 0, 0};
            if (r > 0) {
                chain_walk<kPPWindow>(row - kRadixBins, r, prev);
                st_relaxed_gpu_v4(row, make_uint4(kFlagIncl | (prev[0] + tot.x), kFlagIncl | (prev[1] + tot.y),
                                                  kFlagIncl | (prev[2] + tot.z), kFlagIncl | (prev[3] + tot.w)));
            }
            // level 2: earlier groups (the last tile of a group owns the group's row)
            const uint32_t gflag = (group == 0) ? kFlagIncl : kFlagLocal;
            const uint4 gtot = make_uint4(prev[0] + tot.x, prev[1] + tot.y, prev[2] + tot.z, prev[3] + tot.w);
            if (last_of_group)
                st_relaxed_gpu_v4(grow, make_uint4(gflag | gtot.x, gflag | gtot.y, gflag | gtot.z, gflag | gtot.w));
            if (group > 0) {
                uint32_t gprev[4] = {0, 0, 0, 0};
                chain_walk<kPPWindow>(grow - kRadixBins, group, gprev);
                if (last_of_group)
                    st_relaxed_gpu_v4(grow, make_uint4(kFlagIncl | ((gprev[0] + gtot.x) & kValueMask),
                                                       kFlagIncl | ((gprev[1] + gtot.y) & kValueMask),
                                                       kFlagIncl | ((gprev[2] + gtot.z) & kValueMask),
                                                       kFlagIncl | ((gprev[3] + gtot.w) & kValueMask)));
#pragma unroll
                for (int k = 0; k < 4; ++k) prev[k] += gprev[k];
            }
            __syncwarp();
            bar_sync(kBarTstart + b, kRadixBins + kPPChain);
            const uint4 ts = *reinterpret_cast<const uint4 *>(s_tstart + b * kRadixBins + c4);
            *reinterpret_cast<uint4 *>(s_gofs + b * kRadixBins + c4) =
                make_uint4(base4.x + prev[0] - ts.x, base4.y + prev[1] - ts.y,
                           base4.z + prev[2] - ts.z, base4.w + prev[3] - ts.w);
            __threadfence_block();
            bar_arrive(kBarGofs + b, kPPThreads);
            b ^= 1;
        }
        return;
    }

    // ============================ worker warps ============================
    uint32_t *wt = s_table + warp * kRadixBins;
    {
        uint4 *z = reinterpret_cast<uint4 *>(wt);
#pragma unroll
        for (int j = lane; j < kRadixBins / 4; j += 32) z[j] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) s_misc[8] = atomicAdd(&ctl->ticket[pass], 1u);
    bar_sync(kBarW, kPPWorkers);
    uint32_t tile = s_misc[8];
    uint32_t prev_tile = kPPPoison;
    const uint32_t wofs = warp * (32 * IPT) + lane;

    int32_t key[IPT];
    auto load_tile = [&](uint32_t t) {
        const size_t tile_base = (size_t)t * kTile;
        const uint32_t valid = (n - tile_base < (size_t)kTile) ? (uint32_t)(n - tile_base) : (uint32_t)kTile;
        const int32_t *src = in + tile_base + wofs;
        if (valid == (uint32_t)kTile) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) key[i] = ld_stream(src + i * 32);
        } else {
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                key[i] = (wofs + i * 32 < valid) ? ld_stream(src + i * 32) : 0x7FFFFFFF;
        }
    };
    auto write_tile = [&](uint32_t t, int buf) {
        const size_t tile_base = (size_t)t * kTile;
        const uint32_t valid = (n - tile_base < (size_t)kTile) ? (uint32_t)(n - tile_base) : (uint32_t)kTile;
        const int32_t *sk = s_keys + buf * kTile;
        const uint32_t *go = s_gofs + buf * kRadixBins;
        if (valid == (uint32_t)kTile) {
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const uint32_t p = tid + j * kPPWorkers;
                const int32_t k = sk[p];
                st_stream(out + (size_t)(uint32_t)(go[digit_of(k, shift, flip)] + p), k);
            }
        } else {
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const uint32_t p = tid + j * kPPWorkers;
                if (p < valid) {
                    const int32_t k = sk[p];
                    st_stream(out + (size_t)(uint32_t)(go[digit_of(k, shift, flip)] + p), k);
                }
            }
        }
    };

    if (tile < tiles) load_tile(tile);
    int b = 0;
    uint32_t iter = 0;
    while (tile < tiles) {
        // ---- rank: one shared-memory atomicAdd per key (lane-ordered; see the self-test) ----------
        uint32_t rank2[IPT / 2];
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint32_t r = atomicAdd(wt + digit_of(key[i], shift, flip), 1u);
            rank2[i / 2] = (i & 1) ? (rank2[i / 2] | (r << 16)) : r;
        }
        bar_sync(kBarW, kPPWorkers);
        // next ticket (everybody has read the slot being overwritten: that read precedes this barrier)
        if (tid == 0) s_misc[8 + ((iter + 1) & 1)] = atomicAdd(&ctl->ticket[pass], 1u);

        // ---- threads 0..255, thread = digit: totals -> chain; scan; counts -> positions ----------
        if (tid < kRadixBins) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < kPPWorkerWarps; ++w) total += s_table[w * kRadixBins + tid];
            s_total[b * kRadixBins + tid] = total;
            if (tid == 0) s_misc[10 + b] = tile;
            __threadfence_block();
            bar_arrive(kBarTotals + b, kRadixBins + kPPChain);
            uint32_t x = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= (uint32_t)o) x += y;
            }
            if (lane == 31) s_misc[warp] = x;
            bar_sync(kBarA, kRadixBins);
            uint32_t add = 0;
#pragma unroll
            for (int w = 0; w < kRadixBins / 32; ++w) add += (w < (int)warp) ? s_misc[w] : 0u;
            const uint32_t tile_start = x - total + add;
            uint32_t run = tile_start;
#pragma unroll
            for (int w = 0; w < kPPWorkerWarps; ++w) {
                const uint32_t c = s_table[w * kRadixBins + tid];
                s_table[w * kRadixBins + tid] = run;
                run += c;
            }
            s_tstart[b * kRadixBins + tid] = tile_start;
            __threadfence_block();
            bar_arrive(kBarTstart + b, kRadixBins + kPPChain);
        }
        bar_sync(kBarW, kPPWorkers);                       // positions are final
        const uint32_t next = s_misc[8 + ((iter + 1) & 1)];

        // ---- stage this tile's keys in digit order ----------------------------------------------------
        {
            int32_t *sk = s_keys + b * kTile;
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const uint32_t r = (i & 1) ? (rank2[i / 2] >> 16) : (rank2[i / 2] & 0xffffu);
                sk[wt[digit_of(key[i], shift, flip)] + r] = key[i];
            }
        }
        __syncwarp();
        {
            uint4 *z = reinterpret_cast<uint4 *>(wt);      // my warp's counters, for the next tile
#pragma unroll
            for (int j = lane; j < kRadixBins / 4; j += 32) z[j] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();

        // ---- loads of the next tile go out now and land while the previous tile is written ----
        if (next < tiles) load_tile(next);
        if (prev_tile != kPPPoison) {
            bar_sync(kBarGofs + (b ^ 1), kPPThreads);      // the chain finished tile i-1 long ago
            write_tile(prev_tile, b ^ 1);
        }
        prev_tile = tile;
        tile = next;
        b ^= 1;
        ++iter;
    }
    if (prev_tile != kPPPoison) {
        bar_sync(kBarW, kPPWorkers);                       // the last tile is fully staged
        bar_sync(kBarGofs + (b ^ 1), kPPThreads);
        write_tile(prev_tile, b ^ 1);
    }
    if (tid < kRadixBins) {                                // release the chain warps
        if (tid == 0) s_misc[10 + b] = kPPPoison;
        __threadfence_block();
        bar_arrive(kBarTotals + b, kRadixBins + kPPChain);
    }
}

// ================================================================================================
// k2'': persistent CTA, every warp a worker, DELAYED two-level look-back
// ================================================================================================
// What the phase probe showed (profiles/r01_phase_timing.txt): a tile needs the counts of the tiles
// that started a few hundred nanoseconds before it, and those are often not published yet --
// the look-back does not wait for a long chain but for STRAGGLERS among its ~32 nearest
// predecessors, 4-5 us of a 9 us tile lifetime, with the SM's registers and shared memory held idle.
// Here the CTA does not wait: it publishes tile i's counts, then ranks and stages tile i+1, and only
// then resolves tile i's prefix -- by which time every straggler has long published -- and writes
// tile i out.  The two-level rows make that possible: a tile's own total and a group's own total do
// not depend on anybody's look-back, so delaying one's OWN prefix delays nobody else.  Only the last
// tile of each group sums its group right away (1 tile in 32 waits for stragglers).
//   tile row  : kFlagLocal = the tile's digit counts, kFlagIncl = inclusive within its group
//   group row : kFlagLocal = the group's digit counts, kFlagIncl = inclusive over all groups
template <int IPT>
struct Pipelined2Shape {
    static constexpr int kThreads = 512;
    static constexpr int kTile = kThreads * IPT;
    static constexpr size_t kSmemBytes =
        (size_t)16 * kRadixBins * 4                 // per-warp digit counters -> positions
        + (size_t)2 * kTile * 4                     // two staging buffers
        + (size_t)(2 + 1 + 2) * kRadixBins * 4      // gofs[2], total, tstart[2]
        + 128;
};

template <int IPT, int TIMING = 0>
__global__ void __launch_bounds__(512, 2)
radix_onesweep_pipelined2_kernel(const int32_t *in_buf, int32_t *out_buf, int32_t *tmp_buf, size_t n,
                                 int pass, RadixControl *ctl, uint32_t *status_cur, uint32_t *status_next,
                                 int follow_plan)
{
    constexpr int kThreads = 512, kWarps = 16;
    constexpr int kTile = Pipelined2Shape<IPT>::kTile;
    constexpr int W = 8;                                      // status rows in flight per thread
    static_assert(IPT % 2 == 0 && 32 * IPT < 65536, "ranks are packed in pairs");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_table  = reinterpret_cast<uint32_t *>(smem_raw);                      // [16][256]
    int32_t  *s_keys   = reinterpret_cast<int32_t *>(s_table + kWarps * kRadixBins);  // [2][kTile]
    uint32_t *s_gofs   = reinterpret_cast<uint32_t *>(s_keys + 2 * kTile);            // [2][256]
    uint32_t *s_total  = s_gofs + 2 * kRadixBins;                                     // [256]
    uint32_t *s_tstart = s_total + kRadixBins;                                        // [2][256]
    uint32_t *s_misc   = s_tstart + 2 * kRadixBins;        // [0..7] warp sums, [8..9] tickets

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t tiles = (n + kTile - 1) / kTile;

    const int32_t *in = in_buf;
    int32_t *out = out_buf;
    if (follow_plan) {
        if (ctl->skip[pass]) {
            const size_t rows = tiles + (tiles + kLookGroup - 1) / kLookGroup;
            if (status_next != nullptr)
                for (size_t row = blockIdx.x; row < rows; row += gridDim.x)
                    if (tid < kRadixBins) status_next[row * kRadixBins + tid] = 0;
            return;
        }
        const uint32_t ss = ctl->src_sel[pass], ds = ctl->dst_sel[pass];
        in = (ss == kSelIn) ? in_buf : (ss == kSelTmp) ? tmp_buf : out_buf;
        out = (ds == kSelTmp) ? tmp_buf : out_buf;
    }
    const int shift = pass * kRadixBits;
    const uint32_t flip = (pass == kRadixPasses - 1) ? 0x80u : 0u;
    const uint32_t lt = lanemask_lt();
    const bool in_a = tid < kRadixBins;                       // warps 0..7 : thread = digit
    const bool in_b = !in_a;                                  // warps 8..15: thread - 256 = digit
    const uint32_t bd = tid - kRadixBins;
    uint32_t *wt = s_table + warp * kRadixBins;
    const uint32_t wofs = warp * (32 * IPT) + lane;

    int32_t key[IPT];
    auto load_tile = [&](uint32_t t) {
        const size_t tile_base = (size_t)t * kTile;
        const uint32_t valid = (n - tile_base < (size_t)kTile) ? (uint32_t)(n - tile_base) : (uint32_t)kTile;
        const int32_t *src = in + tile_base + wofs;
        if (valid == (uint32_t)kTile) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) key[i] = ld_stream(src + i * 32);
        } else {
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                key[i] = (wofs + i * 32 < valid) ? ld_stream(src + i * 32) : 0x7FFFFFFF;
        }
    };
    auto write_tile = [&](uint32_t t, int buf) {
        const size_t tile_base = (size_t)t * kTile;
        const uint32_t valid = (n - tile_base < (size_t)kTile) ? (uint32_t)(n - tile_base) : (uint32_t)kTile;
        const int32_t *sk = s_keys + buf * kTile;
        const uint32_t *go = s_gofs + buf * kRadixBins;
        if (valid == (uint32_t)kTile) {
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const uint32_t p = tid + j * kThreads;
                const int32_t k = sk[p];
                st_stream(out + (size_t)(uint32_t)(go[digit_of(k, shift, flip)] + p), k);
            }
        } else {
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const uint32_t p = tid + j * kThreads;
                if (p < valid) {
                    const int32_t k = sk[p];
                    st_stream(out + (size_t)(uint32_t)(go[digit_of(k, shift, flip)] + p), k);
                }
            }
        }
    };

    // group B's memory of the previous tile (the one whose prefix is resolved one iteration late)
    uint32_t digit_base = in_b ? ctl->base[pass][bd] : 0u;
    uint32_t p_total = 0, p_in = 0;                           // its count of my digit; in-group prefix if known
    bool p_in_known = false;
    // The previous tile's look-back, run by group B: fills s_gofs[buf].
    auto resolve_prev = [&](uint32_t pt, int buf) {
        const uint32_t group = pt / kLookGroup, r = pt % kLookGroup;
        const bool last_of_group = (r == kLookGroup - 1) || ((size_t)pt + 1 == tiles);
        uint32_t *row = status_cur + (size_t)pt * kRadixBins + bd;
        uint32_t *grow = status_cur + (tiles + group) * kRadixBins + bd;
        uint32_t inprev = p_in;
        if (!p_in_known) {
            inprev = (r > 0) ? walk_back<W>(row - kRadixBins, r) : 0u;
            if (r > 0) st_relaxed_gpu(row, kFlagIncl | (inprev + p_total));   // shortens later walks
        }
        uint32_t gprev = 0;
        if (group > 0) {
            gprev = walk_back<W>(grow - kRadixBins, group);
            if (last_of_group) st_relaxed_gpu(grow, kFlagIncl | ((gprev + inprev + p_total) & kValueMask));
        }
        s_gofs[buf * kRadixBins + bd] = digit_base + inprev + gprev - s_tstart[buf * kRadixBins + bd];
    };

    {
        uint4 *z = reinterpret_cast<uint4 *>(wt);
#pragma unroll
        for (int j = lane; j < kRadixBins / 4; j += 32) z[j] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) s_misc[8] = atomicAdd(&ctl->ticket[pass], 1u);
    __syncthreads();
    uint32_t tile = s_misc[8];
    uint32_t prev_tile = 0xFFFFFFFFu;
    if (tile < tiles) load_tile(tile);
    int b = 0;
    uint32_t iter = 0;

    while (tile < tiles) {
        const uint32_t dbg_tile = tile;
        if (TIMING) { asm volatile("" :: "r"(key[0]), "r"(key[IPT - 1])); }
        B200_STAMP(0);                                        // this tile's keys are in registers
        // ---- rank: one shared-memory atomicAdd per key (lane-ordered; see the self-test) ----------
        uint32_t rank2[IPT / 2];
        {
            const uint32_t d0 = digit_of(key[0], shift, flip);
            const uint32_t agree = __ballot_sync(0xffffffffu, d0 == __shfl_sync(0xffffffffu, d0, 0));
            const bool hot = (follow_plan && ctl->hot[pass] != 0) || __popc(agree) >= 8;
            if (!hot) {
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    const uint32_t r = atomicAdd(wt + digit_of(key[i], shift, flip), 1u);
                    rank2[i / 2] = (i & 1) ? (rank2[i / 2] | (r << 16)) : r;
                }
            } else {
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    const uint32_t d = digit_of(key[i], shift, flip);
                    const bool same = (d == __shfl_sync(0xffffffffu, d, 0));
                    const uint32_t sm = __ballot_sync(0xffffffffu, same);
                    uint32_t r = 0;
                    if (!same || lane == 0) r = atomicAdd(wt + d, lane == 0 ? (uint32_t)__popc(sm) : 1u);
                    const uint32_t r0 = __shfl_sync(0xffffffffu, r, 0);
                    if (same) r = r0 + __popc(sm & lt);
                    rank2[i / 2] = (i & 1) ? (rank2[i / 2] | (r << 16)) : r;
                }
            }
        }
        if (TIMING) { asm volatile("" :: "r"(rank2[0]), "r"(rank2[IPT / 2 - 1])); }
        B200_STAMP(1);                                        // ranked
        __syncthreads();                                      // SYNC1: counts are final
        B200_STAMP(2);
        if (tid == 0) s_misc[8 + ((iter + 1) & 1)] = atomicAdd(&ctl->ticket[pass], 1u);

        if (in_a) {
            // thread = digit: tile totals -> group B; exclusive scan; warp counts -> positions
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) total += s_table[w * kRadixBins + tid];
            s_total[tid] = total;
            __threadfence_block();
            bar_arrive(2, 512);
            uint32_t x = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= (uint32_t)o) x += y;
            }
            if (lane == 31) s_misc[warp] = x;
            bar_sync(1, kRadixBins);
            uint32_t add = 0;
#pragma unroll
            for (int w = 0; w < kRadixBins / 32; ++w) add += (w < (int)warp) ? s_misc[w] : 0u;
            const uint32_t tile_start = x - total + add;
            uint32_t run = tile_start;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const uint32_t c = s_table[w * kRadixBins + tid];
                s_table[w * kRadixBins + tid] = run;
                run += c;
            }
            s_tstart[b * kRadixBins + tid] = tile_start;
            B200_STAMP(3);                                    // group A done
        } else {
            // publish this tile's counts at once ...
            bar_sync(2, 512);
            const uint32_t total = s_total[bd];
            const uint32_t group = tile / kLookGroup, r = tile % kLookGroup;
            const bool last_of_group = (r == kLookGroup - 1) || ((size_t)tile + 1 == tiles);
            uint32_t *row = status_cur + (size_t)tile * kRadixBins + bd;
            st_relaxed_gpu(row, (r == 0 ? kFlagIncl : kFlagLocal) | total);
            if (status_next != nullptr) {
                status_next[(size_t)tile * kRadixBins + bd] = 0;
                if (last_of_group) status_next[(tiles + group) * kRadixBins + bd] = 0;
            }
            B200_STAMP(10);                                   // published
            // ... resolve the PREVIOUS tile's prefix (everything it needs was published long ago) ...
            if (prev_tile != 0xFFFFFFFFu) resolve_prev(prev_tile, b ^ 1);
            B200_STAMP(11);                                   // previous tile resolved
            // ... and, for the last tile of a group only, sum the group now so that nobody after
            // it has to wait an iteration for the group's total
            p_total = total;
            p_in_known = false;
            if (last_of_group) {
                p_in = (r > 0) ? walk_back<W>(row - kRadixBins, r) : 0u;
                p_in_known = true;
                if (r > 0) st_relaxed_gpu(row, kFlagIncl | (p_in + total));
                uint32_t *grow = status_cur + (tiles + group) * kRadixBins + bd;
                st_relaxed_gpu(grow, (group == 0 ? kFlagIncl : kFlagLocal) | (p_in + total));
            }
            __syncwarp();
            B200_STAMP(3);                                    // group B done
        }
        __syncthreads();                                      // SYNC2: positions final, previous tile's offsets ready
        B200_STAMP(4);
        const uint32_t next = s_misc[8 + ((iter + 1) & 1)];

        // ---- stage this tile's keys in digit order ----------------------------------------------------
        {
            int32_t *sk = s_keys + b * kTile;
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const uint32_t r = (i & 1) ? (rank2[i / 2] >> 16) : (rank2[i / 2] & 0xffffu);
                sk[wt[digit_of(key[i], shift, flip)] + r] = key[i];
            }
        }
        __syncwarp();
        {
            uint4 *z = reinterpret_cast<uint4 *>(wt);          // my warp's counters, for the next tile
#pragma unroll
            for (int j = lane; j < kRadixBins / 4; j += 32) z[j] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        B200_STAMP(5);                                        // staged
        // ---- the next tile's loads go out now and land while the previous tile is written --------
        if (next < tiles) load_tile(next);
        B200_STAMP(6);
        if (prev_tile != 0xFFFFFFFFu) write_tile(prev_tile, b ^ 1);
        B200_STAMP(7);                                        // previous tile written
        if (TIMING && g_phase_dbg != nullptr && lane == 0 && (warp == 0 || warp == 8))
            g_phase_dbg[((size_t)dbg_tile * 2 + (warp >> 3)) * 16 + 9] = tile;
        prev_tile = tile;
        tile = next;
        b ^= 1;
        ++iter;
    }
    // ---- drain: the last tile is staged, its prefix is still to be resolved ----------------------------
    if (prev_tile != 0xFFFFFFFFu) {
        __syncthreads();
        if (in_b) resolve_prev(prev_tile, b ^ 1);
        __syncthreads();
        write_tile(prev_tile, b ^ 1);
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

// ================================================================================================
// host side
// ================================================================================================
namespace {

using OnesweepFn = void (*)(const int32_t *, int32_t *, int32_t *, size_t, int, RadixControl *,
                            uint32_t *, uint32_t *, int);

struct Variant {
    const char *name;
    int mode;
    int cluster;       // CTAs per cluster (1 = none); 0 marks the persistent pipelined kernel
    int two_level;     // status rows: one per tile plus one per group of kLookGroup tiles
    int threads;
    int tile;
    size_t smem;
    OnesweepFn fn;
};

#define B200_VARIANT(W, I, B, M, C)                                                                 \
    { "warps" #W "_ipt" #I "_occ" #B "_" #M "_cl" #C, M, C, 0, OnesweepShape<W, I, M>::kThreads,      \
      OnesweepShape<W, I, M>::kTile, OnesweepShape<W, I, M>::kSmemBytes,                            \
      radix_onesweep_kernel<W, I, B, M, C> }

#define B200_VARIANT_X(W, I, B, M, C, P, S)                                                         \
    { "warps" #W "_ipt" #I "_occ" #B "_" #M "_cl" #C "_pf" #P "_bsf" #S, M, C, 0, OnesweepShape<W, I, M>::kThreads, \
      OnesweepShape<W, I, M>::kTile, OnesweepShape<W, I, M>::kSmemBytes,                            \
      radix_onesweep_kernel<W, I, B, M, C, P, S> }

#define B200_VARIANT_T(W, I, B, M, C, P, S, L)                                                       \
    { "TIMING_warps" #W "_ipt" #I "_" #M "_pf" #P "_tl" #L, M, C, L, OnesweepShape<W, I, M>::kThreads,  \
      OnesweepShape<W, I, M>::kTile, OnesweepShape<W, I, M>::kSmemBytes,                            \
      radix_onesweep_kernel<W, I, B, M, C, P, S, 1, L> }

#define B200_VARIANT_TL(W, I, B, M, P)                                                              \
    { "warps" #W "_ipt" #I "_occ" #B "_" #M "_pf" #P "_twolevel", M, 1, 1, OnesweepShape<W, I, M>::kThreads, \
      OnesweepShape<W, I, M>::kTile, OnesweepShape<W, I, M>::kSmemBytes,                            \
      radix_onesweep_kernel<W, I, B, M, 1, P, 1, 0, 1> }

#define B200_PP_VARIANT(I)                                                                          \
    { "pipelined_14w_ipt" #I "_kRankAdd", kRankAdd, 0, 1, kPPThreads, PipelinedShape<I>::kTile,     \
      PipelinedShape<I>::kSmemBytes, radix_onesweep_pipelined_kernel<I> }

#define B200_PP2_VARIANT(I)                                                                         \
    { "pipelined2_16w_ipt" #I "_kRankAdd_delayed_twolevel", kRankAdd, 0, 1, 512,                    \
      Pipelined2Shape<I>::kTile, Pipelined2Shape<I>::kSmemBytes, radix_onesweep_pipelined2_kernel<I> }

const Variant kVariants[] = {
    B200_PP2_VARIANT(18),                      //  0: DEFAULT (fastest measured): persistent CTAs, 9216-key tiles,
                                               //     delayed two-level look-back
    B200_VARIANT(16, 18, 2, kRankAdd, 1),      //  1: 9216
    B200_VARIANT(16, 16, 2, kRankAdd, 1),      //  2: 8192
    B200_VARIANT(8, 24, 3, kRankAdd, 1),       //  3: 6144, 256 threads
    B200_VARIANT(8, 16, 4, kRankAdd, 1),       //  4: 4096, 4 CTAs/SM
    B200_VARIANT(16, 16, 2, kRankBallot, 1),   //  5: the documented-behaviour fallback
    B200_VARIANT(16, 16, 2, kRankAtomic, 1),   //  6
    B200_VARIANT(16, 16, 2, kRankMatch, 1),    //  7
    B200_VARIANT(8, 24, 3, kRankBallot, 1),    //  8
    B200_VARIANT(8, 8, 6, kRankBallot, 1),     //  9: 2048-key tiles
    B200_VARIANT(12, 16, 3, kRankAdd, 1),      // 10: 6144, 384 threads
    B200_VARIANT(16, 12, 2, kRankAdd, 1),      // 11: 6144, 512 threads
    B200_VARIANT(16, 16, 2, kRankAdd, 2),      // 12: clusters of 2 / 4 / 8 CTAs = one chain link
    B200_VARIANT(16, 16, 2, kRankAdd, 4),      // 13
    B200_VARIANT(16, 16, 2, kRankAdd, 8),      // 14
    B200_VARIANT(16, 18, 2, kRankAdd, 4),      // 15
    B200_VARIANT(16, 20, 2, kRankAdd, 4),      // 16
    B200_VARIANT(16, 20, 2, kRankAdd, 8),      // 17
    B200_VARIANT(16, 16, 2, kRankBallot, 4),   // 18
    B200_PP_VARIANT(20),                       // 19: persistent pipelined, 448 x 20 = 8960-key tiles
    B200_PP_VARIANT(16),                       // 20: 7168
    B200_PP_VARIANT(24),                       // 21: 10752
    B200_PP_VARIANT(18),                       // 22: 8064
    B200_VARIANT(32, 10, 2, kRankAdd, 1),      // 23: 1024 threads x 10 keys, 2 CTAs/SM = full occupancy, 32 regs
    B200_VARIANT(32, 8, 2, kRankAdd, 1),       // 24: 8192
    B200_VARIANT(32, 12, 1, kRankAdd, 1),      // 25: 12288, 1 CTA/SM
    B200_VARIANT(24, 12, 2, kRankAdd, 1),      // 26: 768 threads x 12 = 9216, 2 CTAs/SM (42 regs)
    B200_VARIANT_X(16, 20, 2, kRankAdd, 1, 296, 0),   // 27: default shape + L2 prefetch one CTA-lifetime ahead
    B200_VARIANT_X(16, 20, 2, kRankAdd, 1, 0, 1),     // 28: default shape, group B stages before the look-back
    B200_VARIANT_X(16, 20, 2, kRankAdd, 1, 296, 1),   // 29: both
    B200_VARIANT_X(16, 20, 2, kRankAdd, 1, 592, 1),   // 30: both, two lifetimes ahead
    B200_VARIANT_X(16, 16, 2, kRankAdd, 1, 296, 1),   // 31: 8192-key tiles, both
    B200_VARIANT_T(16, 20, 2, kRankAdd, 1, 296, 0, 0),   // 32: variant 27 with the phase-timing probe
    B200_VARIANT_TL(16, 20, 2, kRankAdd, 296),        // 33: two-level look-back, 10240-key tiles
    B200_VARIANT_TL(16, 20, 2, kRankAdd, 0),          // 34: same without the L2 prefetch
    B200_VARIANT_TL(16, 16, 2, kRankAdd, 296),        // 35: 8192
    B200_VARIANT_TL(16, 18, 2, kRankAdd, 296),        // 36: 9216
    B200_VARIANT_TL(16, 22, 2, kRankAdd, 296),        // 37: 11264
    B200_VARIANT_TL(16, 24, 2, kRankAdd, 296),        // 38: 12288
    B200_VARIANT_T(16, 20, 2, kRankAdd, 1, 296, 1, 1),   // 39: variant 33 with the phase-timing probe
    B200_PP2_VARIANT(20),                             // 40: persistent, delayed two-level look-back, 10240
    B200_PP2_VARIANT(16),                             // 41: 8192
    B200_VARIANT(16, 20, 2, kRankAdd, 1),             // 42: one tile per CTA, 10240-key tiles (round-1 default until PP2)
    B200_PP2_VARIANT(22),                             // 43: 11264
    { "TIMING_pipelined2_ipt18", kRankAdd, 0, 1, 512, Pipelined2Shape<18>::kTile,
      Pipelined2Shape<18>::kSmemBytes, radix_onesweep_pipelined2_kernel<18, 1> },   // 44
};
constexpr int kFallbackVariant = 5;
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

std::atomic<int> g_variant{0};
std::atomic<int> g_atomic_order{-1};      // -1 unknown, 0 the self-test failed, 1 it passed

int atomic_order_ok() {
    int v = g_atomic_order.load(std::memory_order_acquire);
    if (v >= 0) return v;
    const char *env = getenv("B200SORT_RANK_SAFE");
    if (env != nullptr && env[0] == '1') { g_atomic_order.store(0); return 0; }
    uint32_t *d_bad = nullptr, h_bad = 1;
    if (cudaMalloc(&d_bad, sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaMemset(d_bad, 0, sizeof(uint32_t));
    radix_atomic_order_selftest_kernel<<<kNumSMs * 2, 512>>>(d_bad);
    ++g_launch_count;
    if (cudaMemcpy(&h_bad, d_bad, sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); h_bad = 1; }
    cudaFree(d_bad);
    v = (h_bad == 0) ? 1 : 0;
    g_atomic_order.store(v, std::memory_order_release);
    return v;
}

// The variant to launch: the selected one, unless it needs lane-ordered atomics and this device
// failed (or was told to skip) the self-test.
int effective_variant();
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

std::atomic<bool> g_hist_attr_set{false};
int ensure_hist_attr() {
    if (!g_hist_attr_set.load(std::memory_order_acquire)) {
        B200_CUDA_TRY(cudaFuncSetAttribute(reinterpret_cast<const void *>(radix_histogram_kernel),
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHistSmemBytes));
        g_hist_attr_set.store(true, std::memory_order_release);
    }
    return B200SORT_OK;
}

// status rows a pass needs: one per tile, plus (pipelined kernel) one per group of tiles
size_t status_rows(const Variant &var, size_t tiles) {
    return var.two_level ? tiles + div_up(tiles, (size_t)kLookGroup) : tiles;
}

int launch_onesweep(const Variant &var, size_t tiles, cudaStream_t s, const int32_t *in, int32_t *out,
                    int32_t *tmp, size_t n, int pass, RadixControl *ctl, uint32_t *cur, uint32_t *next,
                    int follow_plan) {
    if (var.cluster == 0) {          // persistent: one CTA per resident slot, tiles by ticket
        const unsigned slots = 2u * kNumSMs;
        const unsigned grid = (unsigned)(tiles < slots ? tiles : slots);
        var.fn<<<grid, var.threads, var.smem, s>>>(in, out, tmp, n, pass, ctl, cur, next, follow_plan);
        B200_LAUNCH_CHECK();
        return B200SORT_OK;
    }
    const unsigned grid = (unsigned)(div_up(tiles, (size_t)var.cluster) * var.cluster);
    if (var.cluster == 1) {
        var.fn<<<grid, var.threads, var.smem, s>>>(in, out, tmp, n, pass, ctl, cur, next, follow_plan);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid, 1, 1);
        cfg.blockDim = dim3((unsigned)var.threads, 1, 1);
        cfg.dynamicSmemBytes = var.smem;
        cfg.stream = s;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = (unsigned)var.cluster;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        B200_CUDA_TRY(cudaLaunchKernelEx(&cfg, var.fn, in, out, tmp, n, pass, ctl, cur, next, follow_plan));
    }
    B200_LAUNCH_CHECK();
    return B200SORT_OK;
}

int hist_grid(size_t n) {
    const size_t chunks = div_up(div_up(n, 4), (size_t)kHistThreads * kHistUnroll);
    size_t g = chunks < (size_t)kNumSMs * kHistBlocksPerSM ? chunks : (size_t)kNumSMs * kHistBlocksPerSM;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace

namespace {
int effective_variant() {
    const int v = g_variant.load();
    if (kVariants[v].mode == kRankAdd && !atomic_order_ok()) return kFallbackVariant;
    return v;
}
}  // namespace

int radix_set_phase_debug(long long *d_buf) {
    B200_CUDA_TRY(cudaMemcpyToSymbol(g_phase_dbg, &d_buf, sizeof d_buf));
    return B200SORT_OK;
}
int radix_atomic_order_ok() { return atomic_order_ok(); }
int radix_num_variants() { return kNumVariants; }
const char *radix_variant_name(int v) { return (v >= 0 && v < kNumVariants) ? kVariants[v].name : nullptr; }
int radix_set_variant(int v) {
    if (v < 0 || v >= kNumVariants) return B200SORT_ERR_INVALID;
    g_variant.store(v);
    return B200SORT_OK;
}
void radix_set_skip(int enabled) { g_skip_enabled.store(enabled ? 1 : 0); }
size_t radix_current_tile() { return (size_t)kVariants[g_variant.load()].tile; }
const char *radix_effective_variant_name() { return kVariants[effective_variant()].name; }

size_t radix_workspace_bytes(size_t n) {
    const size_t tiles = div_up(n > 0 ? n : 1, kRadixMinTile);
    const size_t rows = tiles + div_up(tiles, (size_t)kPPGroup) + 1;
    return kRadixControlBytes + 2 * rows * kRadixBins * sizeof(uint32_t);
}

int radix_histogram(const int32_t *d_keys, size_t n, uint32_t *d_hist, cudaStream_t s) {
    // Standalone histogram (unit tests, per-kernel timing): d_hist doubles as the control block's
    // histogram area, so a scratch control block is not needed -- the kernel is given a control
    // block that lives in a small static device allocation.
    static thread_local RadixControl *scratch = nullptr;
    if (scratch == nullptr) B200_CUDA_TRY(cudaMalloc(&scratch, kRadixControlBytes));
    B200_TRY(ensure_hist_attr());
    B200_CUDA_TRY(cudaMemsetAsync(scratch, 0, kRadixZeroBytes, s));
    if (n > 0) {
        radix_histogram_kernel<<<hist_grid(n), kHistThreads, kHistSmemBytes, s>>>(d_keys, n, scratch, nullptr, 0, 0, 0);
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
    const int v = effective_variant();
    B200_TRY(ensure_smem_attr(v));
    B200_TRY(ensure_hist_attr());
    const Variant &var = kVariants[v];
    auto *ctl = static_cast<RadixControl *>(d_ws);
    auto *status0 = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(d_ws) + kRadixControlBytes);
    const size_t tiles = div_up(n, (size_t)var.tile);
    B200_CUDA_TRY(cudaMemsetAsync(ctl, 0, kRadixZeroBytes, s));
    radix_histogram_kernel<<<hist_grid(n), kHistThreads, kHistSmemBytes, s>>>(d_in, n, ctl, status0,
                                                                 status_rows(var, tiles) * kRadixBins, 0, 0);
    B200_LAUNCH_CHECK();
    B200_TRY(launch_onesweep(var, tiles, s, d_in, d_out, nullptr, n, pass, ctl, status0, nullptr, 0));
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
    const int v = effective_variant();
    B200_TRY(ensure_smem_attr(v));
    B200_TRY(ensure_hist_attr());
    const Variant &var = kVariants[v];
    auto *ctl = static_cast<RadixControl *>(d_ws);
    const size_t tiles = div_up(n, (size_t)var.tile);
    const size_t rows = status_rows(var, tiles);
    uint32_t *status[2];
    status[0] = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(d_ws) + kRadixControlBytes);
    status[1] = status[0] + rows * kRadixBins;
    const int skip = g_skip_enabled.load();
    const uint32_t in_place = (d_in == d_out) ? 1u : 0u;
    StepTimer timer{s, ms};

    B200_CUDA_TRY(cudaMemsetAsync(ctl, 0, kRadixZeroBytes, s));
    B200_TRY(timer.begin());
    radix_histogram_kernel<<<hist_grid(n), kHistThreads, kHistSmemBytes, s>>>(d_in, n, ctl, status[0], rows * kRadixBins,
                                                                 (uint32_t)skip, in_place);
    B200_LAUNCH_CHECK();
    B200_TRY(timer.mark());
    for (int pass = 0; pass < kRadixPasses; ++pass) {
        uint32_t *cur = status[pass & 1];
        uint32_t *next = (pass + 1 < kRadixPasses) ? status[(pass + 1) & 1] : nullptr;
        B200_TRY(launch_onesweep(var, tiles, s, d_in, d_out, d_tmp, n, pass, ctl, cur, next, 1));
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

// radix_tma.cuh -- k2: one onesweep pass, Blackwell shape (included by radix.cu).
//
// Replaces the lab's radix stage (SRM/lab.cu:47-87 radix_sort_kernel + :11-41 exlusiveScan) like the other
// pass kernels; this is the one that uses what sm_100 added.
//
// Round 1's pass (radix_pipelined.cuh) is bound by the SM's load/store pipe: ~17.5 shared-memory wavefronts
// per 32 keys, 5.5 of them for writing the staged tile out (staged read, offset lookup, a store whose 32 lanes
// span ~3 lines).  tools/probe_tma_scatter.cu measured that the TMA unit scatters a staged tile in 160-byte
// runs at the full HBM rate without touching that pipe (6.3 TB/s; one bulk copy per ~6 cycles and SM) -- but
// a bulk copy needs 16-byte aligned source AND destination, so the staging has to be laid out AFTER the
// tile's global prefix is known, and the round-1 kernel resolves that prefix one iteration late on purpose
// (decoupled look-back without waiting).  The way out is the one on-chip memory round 1 left idle:
//
//   iteration j of a persistent CTA (16 warps, 2 CTAs per SM, 10240-key tiles by ticket; t = the tile ranked
//   now, p = the tile ranked an iteration ago):
//        prefetch   ONE thread asks the TMA unit for the status rows of p's look-back: the earlier tile rows of
//                   p's group and the nearest group rows are contiguous, so that is two bulk loads
//                   (cp.async.bulk.shared::cta.global + mbarrier) that land underneath the ranking.
//     R  rank t     keys (loaded an iteration ago) -> one shared-memory atomicAdd per key -> keys and ranks are
//                   PARKED IN TENSOR MEMORY (tcgen05.st, 32 lanes x 32 columns per warp, lane-private: exactly
//                   the access pattern of key[IPT]).
//     D  digit work group A (thread = digit): t's digit counts -> publish (two-level status rows, as round 1);
//                   group B (thread = digit): p's prefix summed from the prefetched rows (everything it needs
//                   was published an iteration ago), then p's staging layout: every digit run starts at a
//                   shared-memory word that is congruent mod 4 to its first DESTINATION word.
//     S  stage p    keys and ranks come back from tensor memory (tcgen05.ld) and go to their staged words.
//     W  write p    the next tile's loads go out; then per digit run ONE bulk copy shared -> global for the
//                   16-byte aligned interior (cp.async.bulk.global.shared::cta, thread = digit), the <= 3 + 3
//                   words before and after it by ordinary stores.  (Byte-masked bulk copies for those edges
//                   measured 12.7 cycles each: three times slower than the stores.)
//   The ticket of the next tile is drawn AFTER the digit work (the one phase whose length varies), so that from
//   ticket to publication every tile takes the same stage + write + rank time and tiles are published in very
//   nearly ticket order whatever the skew between CTAs.  (Tickets drawn two tiles ahead measured a full
//   iteration of extra look-back waiting -- a convoy: profiles/r02_onesweep_variants.md.)
//
// Per 32 keys the load/store pipe now sees: global load 1, rank atomic ~3.6, position lookup ~3.3, staging
// store ~3.7, digit work ~1, run edges ~2 (~14.5 in total), and the staging area is single instead of
// double (the bulk copies of tile p are done reading it long before tile p+1 is staged).
#pragma once
#include "radix_pipelined.cuh"

namespace b200sort {

// ---- tensor memory: lane-private parking for registers ---------------------------------------------------
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&r)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int kTmaIpt = 20;
constexpr int kTmaThreads = 512;
constexpr int kTmaTile = kTmaThreads * kTmaIpt;                 // 10240 keys
constexpr int kTmaRows = 8;                                     // two warps per counter row, 16 bits each
constexpr int kTmaStageWords = kTmaTile + kRadixBins * 6 + 64;  // every run padded to whole 16-byte chunks
constexpr int kTmaWin1 = kLookGroup - 1;                        // every earlier tile row of the group
constexpr int kTmaWin2 = 12;                                    // the nearest group rows (2 CTAs/SM: 115712 B each)
constexpr int kTmaTmemCols = 256;                               // 2 tiles x 4 warps per lane quadrant x 32 columns
constexpr size_t kTmaSmemBytes =
    (size_t)kTmaStageWords * 4                   // the staged tile
    + (size_t)2 * kTmaRows * kRadixBins * 4      // digit counters -> positions, this tile's and the previous tile's
    + (size_t)7 * kRadixBins * 4                 // run {start, length}, destination, tile counts x 2, in-group prefix x 2, group prefix
    + (size_t)(kTmaWin1 + kTmaWin2) * kRadixBins * 4   // prefetched status rows of the previous tile's look-back
    + 256;

// parked layout of a thread's 20 keys + 10 rank words (two 16-bit ranks per word) in its 32 columns:
//   columns  0..7  keys 0..7    8..11 their ranks | 12..19 keys 8..15  20..23 ranks | 24..27 keys 16..19  28..29 ranks
__host__ __device__ constexpr int tma_key_col(int i)  { return (i / 8) * 12 + (i % 8); }
__host__ __device__ constexpr int tma_rank_col(int i) { return (i / 8) * 12 + ((i / 8) < 2 ? 8 : 4) + (i % 8) / 2; }

template <int TIMING>
__global__ void __launch_bounds__(kTmaThreads, 2)
radix_onesweep_tma_kernel(const int32_t *in_buf, int32_t *out_buf, int32_t *tmp_buf, size_t n, int pass,
                          RadixControl *ctl, uint32_t *status_cur, uint32_t *status_next, int follow_plan)
{
    constexpr int IPT = kTmaIpt;
    constexpr int kTile = kTmaTile;
    constexpr int kRows = kTmaRows;
    constexpr int W = 8;                                        // status rows in flight per thread
    static_assert(IPT == 20, "the tensor-memory layout is written for 20 keys + 10 rank words per thread");

    extern __shared__ __align__(128) unsigned char smem_tma[];
    int32_t  *s_stage  = reinterpret_cast<int32_t *>(smem_tma);                          // [kTmaStageWords]
    uint32_t *s_table  = reinterpret_cast<uint32_t *>(s_stage + kTmaStageWords);         // [2][kRows][256]
    uint32_t *s_run    = s_table + 2 * kRows * kRadixBins;       // [256] first staged word | keys to write << 16
    uint32_t *s_g      = s_run + kRadixBins;                     // [256] first destination word of the run
    uint32_t *s_ptot   = s_g + kRadixBins;                       // [2][256] the tile's digit counts, as published
    uint32_t *s_pin    = s_ptot + 2 * kRadixBins;                // [2][256] in-group prefix, if the tile summed its group
    uint32_t *s_gprev  = s_pin + 2 * kRadixBins;                 // [256] the previous tile's prefix over the earlier groups
    uint32_t *s_win1   = s_gprev + kRadixBins;                   // [kTmaWin1][256] tile rows before the previous tile
    uint32_t *s_win2   = s_win1 + kTmaWin1 * kRadixBins;         // [kTmaWin2][256] group rows before its group
    uint32_t *s_misc   = s_win2 + kTmaWin2 * kRadixBins;         // [0..7] warp sums, [8] ticket, [10] tmem base, [16..17] mbarrier

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
    const bool in_a = tid < kRadixBins;                          // warps 0..7 : thread = digit
    const uint32_t bd = tid - kRadixBins;                        // warps 8..15: thread - 256 = digit
    const uint32_t sh = (warp & 1) * 16;
    const uint32_t wofs = warp * (32 * IPT) + lane;
    // word offset of `out` inside its 16-byte chunk: word g of the array is word g + gmis of the aligned base
    const uint32_t gmis = (uint32_t)((reinterpret_cast<uintptr_t>(out) >> 2) & 3u);
    int32_t *out_al = out - gmis;
    const uint32_t stage_s = smem_u32(s_stage);
    auto pair_bar = [&]() { bar_sync(3 + (warp >> 1), 64); };    // the two warps that share a counter row

    // ---- tensor memory: 256 columns; this warp owns lanes 32*(warp%4).., columns half*128 + 32*(warp/4).. ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&s_misc[10])), "n"(kTmaTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        uint4 *z = reinterpret_cast<uint4 *>(s_table);
        for (uint32_t i = tid; i < 2 * kRows * kRadixBins / 4; i += kTmaThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    const uint32_t mbar = smem_u32(&s_misc[16]);
    if (tid == 0) {
        s_misc[8] = atomicAdd(&ctl->ticket[pass], 1u);
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_misc[10];
    const uint32_t tmem_warp = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 32u;

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
            for (int i = 0; i < IPT; ++i) key[i] = (wofs + i * 32 < valid) ? ld_stream(src + i * 32) : 0x7FFFFFFF;
        }
    };

    uint32_t tile = s_misc[8];
    uint32_t prev_tile = 0xFFFFFFFFu;
    uint32_t win_parity = 0;
    __syncthreads();                                             // s_misc[8] is rewritten inside the loop
    if (tile < tiles) load_tile(tile);
    const uint32_t digit_base = in_a ? 0u : ctl->base[pass][bd];
    uint32_t iter = 0;

    while (tile < tiles || prev_tile != 0xFFFFFFFFu) {
        const bool have_cur = tile < tiles;
        const bool have_prev = prev_tile != 0xFFFFFFFFu;
        const uint32_t cb = iter & 1;                            // counters / tensor-memory half of `tile`
        uint32_t *tab_cur = s_table + cb * kRows * kRadixBins;
        uint32_t *tab_prev = s_table + (cb ^ 1) * kRows * kRadixBins;
        const uint32_t dbg_tile = have_cur ? tile : prev_tile;
        if (TIMING) { asm volatile("" :: "r"(key[0]), "r"(key[IPT - 1])); }
        B200_STAMP(0);                                           // this tile's keys are in registers
        // ---- the previous tile's look-back rows are fetched into shared memory underneath the ranking: the earlier
        // tile rows of its group and the nearest group rows are contiguous, so that is two bulk loads ------------
        uint32_t have1 = 0, have2 = 0;                           // rows fetched (tile rows, group rows)
        if (have_prev) {
            const uint32_t group = prev_tile / kLookGroup, r = prev_tile % kLookGroup;
            const bool last_of_group = (r == kLookGroup - 1) || ((size_t)prev_tile + 1 == tiles);
            have1 = last_of_group ? 0u : r;
            have2 = group < (uint32_t)kTmaWin2 ? group : (uint32_t)kTmaWin2;
            if (tid == kRadixBins && have1 + have2 > 0) {
                fence_proxy_async_smem();                        // last iteration's reads of the windows are done (SYNC2)
                mbar_expect_tx(mbar, (have1 + have2) * kRadixBins * 4);
                if (have1) bulk_load(smem_u32(s_win1), status_cur + ((size_t)prev_tile - have1) * kRadixBins, have1 * kRadixBins * 4, mbar);
                if (have2) bulk_load(smem_u32(s_win2), status_cur + (tiles + group - have2) * kRadixBins, have2 * kRadixBins * 4, mbar);
            }
        }
        // ---- R: rank `tile`, park its keys and ranks in tensor memory ------------------------------------
        if (have_cur) {
            uint32_t *wt = tab_cur + (warp >> 1) * kRadixBins;
            uint32_t pk[32];
            pk[30] = 0; pk[31] = 0;
            {
                const uint32_t d0 = digit_of(key[0], shift, flip);
                const uint32_t agree = __ballot_sync(0xffffffffu, d0 == __shfl_sync(0xffffffffu, d0, 0));
                // hot digit of the pass (the histogram kernel found one value holding > 1/8 of the keys), else a
                // locally hot one (a quarter of the warp's first keys agree with lane 0's: sorted input)
                const uint32_t hot_word = follow_plan ? ctl->hot[pass] : 0u;
                const bool hot = hot_word != 0 || __popc(agree) >= 8;
                if (!hot) {
#pragma unroll
                    for (int i = 0; i < IPT; ++i) {
                        const uint32_t r = (atomicAdd(wt + digit_of(key[i], shift, flip), 1u << sh) >> sh) & 0xffffu;
                        pk[tma_rank_col(i)] = (i & 1) ? (pk[tma_rank_col(i)] | (r << 16)) : r;
                    }
                } else {
                    // the lanes that hold the hot digit are ranked with one ballot and ONE atomic (by their first
                    // lane); the others take the atomic as usual
#pragma unroll
                    for (int i = 0; i < IPT; ++i) {
                        const uint32_t d = digit_of(key[i], shift, flip);
                        const uint32_t hd = hot_word ? hot_word - 1u : __shfl_sync(0xffffffffu, d, 0);
                        const bool same = (d == hd);
                        const uint32_t sm = __ballot_sync(0xffffffffu, same);
                        const uint32_t leader = (uint32_t)(__ffs(sm) - 1) & 31u;
                        uint32_t r = 0;
                        if (!same || lane == leader)
                            r = (atomicAdd(wt + d, (same ? (uint32_t)__popc(sm) : 1u) << sh) >> sh) & 0xffffu;
                        const uint32_t r0 = __shfl_sync(0xffffffffu, r, leader);
                        if (same) r = r0 + __popc(sm & lt);
                        pk[tma_rank_col(i)] = (i & 1) ? (pk[tma_rank_col(i)] | (r << 16)) : r;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < IPT; ++i) pk[tma_key_col(i)] = (uint32_t)key[i];
            tmem_st32(tmem_warp + cb * 128u, pk);                // completion is awaited before it is read back
        }
        B200_STAMP(1);                                           // ranked and parked
        __syncthreads();                                         // SYNC1: `tile`'s counts are final
        B200_STAMP(2);

        // ---- D: digit work, the two thread groups side by side ---------------------------------------------
        if (in_a) {
            if (have_cur) {
                // thread = digit: `tile`'s count of my digit -> its status row (and, for the last tile of a group,
                // the group's row: that tile sums its group at once so that nobody waits an iteration for it)
                uint32_t total = 0;
#pragma unroll
                for (int w = 0; w < kRows; ++w) {
                    const uint32_t c = tab_cur[w * kRadixBins + tid];
                    total += (c & 0xffffu) + (c >> 16);
                }
                const uint32_t group = tile / kLookGroup, r = tile % kLookGroup;
                const bool last_of_group = (r == kLookGroup - 1) || ((size_t)tile + 1 == tiles);
                uint32_t *row = status_cur + (size_t)tile * kRadixBins + tid;
                st_relaxed_gpu(row, (r == 0 ? kFlagIncl : kFlagLocal) | total);
                if (status_next != nullptr) {
                    status_next[(size_t)tile * kRadixBins + tid] = 0;
                    if (last_of_group) status_next[(tiles + group) * kRadixBins + tid] = 0;
                }
                s_ptot[cb * kRadixBins + tid] = total;
                if (last_of_group) {
                    const uint32_t p_in = (r > 0) ? walk_back<W>(row - kRadixBins, r) : 0u;
                    if (r > 0) st_relaxed_gpu(row, kFlagIncl | (p_in + total));
                    uint32_t *grow = status_cur + (tiles + group) * kRadixBins + tid;
                    st_relaxed_gpu(grow, (group == 0 ? kFlagIncl : kFlagLocal) | (p_in + total));
                    s_pin[cb * kRadixBins + tid] = p_in;
                }
            }
            B200_STAMP(3);                                       // group A done
        } else {
            if (have_prev) {
                // thread = digit: the previous tile's prefix (every row it needs was published an iteration ago)
                const uint32_t pb = cb ^ 1;
                const uint32_t p_total = s_ptot[pb * kRadixBins + bd];
                const uint32_t group = prev_tile / kLookGroup, r = prev_tile % kLookGroup;
                const bool last_tile = (size_t)prev_tile + 1 == tiles;
                const bool last_of_group = (r == kLookGroup - 1) || last_tile;
                uint32_t *row = status_cur + (size_t)prev_tile * kRadixBins + bd;
                uint32_t *grow = status_cur + (tiles + group) * kRadixBins + bd;
                if (have1 + have2 > 0) { mbar_wait(mbar, win_parity); win_parity ^= 1; }   // the fetched rows have landed
                B200_STAMP(11);
                uint32_t inprev;
                if (last_of_group) {
                    inprev = s_pin[pb * kRadixBins + bd];        // summed when the tile was published
                } else {
                    inprev = (r > 0) ? walk_back_prefetched<W>(s_win1 + bd, have1, row - kRadixBins, r) : 0u;
                    if (r > 0) st_relaxed_gpu(row, kFlagIncl | (inprev + p_total));   // shortens later walks
                }
                __syncwarp();                                    // the walk diverges per digit
                B200_STAMP(12);
                uint32_t gprev = 0;
                if (group > 0) {
                    gprev = walk_back_prefetched<W>(s_win2 + bd, have2, grow - kRadixBins, group);
                    if (last_of_group) st_relaxed_gpu(grow, kFlagIncl | ((gprev + inprev + p_total) & kValueMask));
                }
                __syncwarp();                                    // the walks diverge per digit
                B200_STAMP(10);                                  // previous tile resolved
                // ... and its staging layout: run d occupies whole 16-byte chunks, its first key sits at the
                // word that is congruent mod 4 to its first destination word
                const uint32_t g = digit_base + inprev + gprev + gmis;       // destination word (from out_al)
                const uint32_t a = g & 3u;
                const uint32_t padded = (a + p_total + 3u) & ~3u;
                uint32_t x = padded;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= (uint32_t)o) x += y;
                }
                if (lane == 31) s_misc[warp - 8] = x;
                bar_sync(1, kRadixBins);
                uint32_t add = 0;
#pragma unroll
                for (int w = 0; w < kRadixBins / 32; ++w) add += (w < (int)warp - 8) ? s_misc[w] : 0u;
                const uint32_t start = x - padded + add + a;
                uint32_t run = start;
#pragma unroll
                for (int w = 0; w < kRows; ++w) {                // warp 2w's keys first, then warp 2w+1's
                    const uint32_t c = tab_prev[w * kRadixBins + bd];
                    const uint32_t lo = c & 0xffffu;
                    tab_prev[w * kRadixBins + bd] = run | ((run + lo) << 16);
                    run += lo + (c >> 16);
                }
                // slots past n (last tile only) carry INT_MAX: digit 255, ranked behind every real key; they are
                // staged but never written
                uint32_t cw = p_total;
                if (last_tile && bd == kRadixBins - 1) cw -= (uint32_t)(tiles * (size_t)kTile - n);
                s_run[bd] = start | (cw << 16);
                s_g[bd] = g;
            }
            B200_STAMP(3);                                       // group B done
        }
        bulk_wait_read_all();                                    // my bulk copies of the tile before are done READING the staging area
        __syncthreads();                                         // SYNC2: the previous tile's positions are final
        B200_STAMP(4);
        // The next tile's ticket is drawn now, AFTER the one phase whose length varies: from ticket to publication
        // every tile then takes the same stage + write + rank time (see the header).
        if (tid == 0) s_misc[8] = atomicAdd(&ctl->ticket[pass], 1u);

        // ---- S: stage the previous tile: keys and ranks come back from tensor memory ------------------------
        if (have_prev) {
            const uint32_t tp = tmem_warp + (cb ^ 1) * 128u;
            const uint32_t *wt = tab_prev + (warp >> 1) * kRadixBins;
            auto stage8 = [&](const uint32_t (&k)[8], const uint32_t (&rk)[4], int count) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < count) {
                        const uint32_t r = (i & 1) ? (rk[i / 2] >> 16) : (rk[i / 2] & 0xffffu);
                        const uint32_t pos = ((wt[digit_of((int32_t)k[i], shift, flip)] >> sh) & 0xffffu) + r;
                        B200_CHECK_AT(4, pos < (uint32_t)kTmaStageWords);
                        s_stage[pos] = (int32_t)k[i];
                    }
                }
            };
            uint32_t ka[8], ra[4], kb[8], rb[4];
            tmem_wait_st();
            tmem_ld8(tp, ka);       tmem_ld4(tp + 8, ra);
            tmem_ld8(tp + 12, kb);  tmem_ld4(tp + 20, rb);
            tmem_wait_ld();
            stage8(ka, ra, 8);
            {
                uint32_t k4[4], r2[2];
                tmem_ld4(tp + 24, k4);
                tmem_ld2(tp + 28, r2);
                stage8(kb, rb, 8);
                tmem_wait_ld();
                ka[0] = k4[0]; ka[1] = k4[1]; ka[2] = k4[2]; ka[3] = k4[3]; ra[0] = r2[0]; ra[1] = r2[1];
                stage8(ka, ra, 4);
            }
            // the counters are cleared for the tile after next once nobody reads positions from them any more
            pair_bar();
            reinterpret_cast<uint4 *>(tab_prev + (warp >> 1) * kRadixBins)[(warp & 1) * 32 + lane] = make_uint4(0, 0, 0, 0);
            fence_proxy_async_smem();                            // staged keys -> visible to the bulk copies
        }
        B200_STAMP(5);                                           // staged
        __syncthreads();                                         // SYNC3: the staged tile is complete
        B200_STAMP(6);
        // the next tile's loads go out now and land while the previous tile is written
        const uint32_t next = s_misc[8];
        if (next < tiles) load_tile(next);

        // ---- W: write the previous tile: interiors by bulk copy, edges by ordinary stores ------------------
        if (have_prev) {
            if (lane < 16) {                                     // 16 warps x 16 lanes: thread = digit
                const uint32_t d = warp * 16 + lane;
                const uint32_t rw = s_run[d], g = s_g[d];
                const uint32_t start = rw & 0xffffu, c = rw >> 16;
                uint32_t head = (4u - (g & 3u)) & 3u;
                if (head > c) head = c;
                const uint32_t body = (c - head) & ~3u;
                B200_CHECK_AT(5, body == 0 || (((g + head) & 3u) == 0 && ((start + head) & 3u) == 0));          // 16-byte aligned on both sides
                B200_CHECK_AT(6, (size_t)g - gmis + c <= n && start + c <= (uint32_t)kTmaStageWords); // inside the array and the tile
                if (body > 0) bulk_store(out_al + g + head, stage_s + (start + head) * 4u, body * 4u);
                bulk_commit();
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint32_t q = tid + j * kTmaThreads;        // 256 digits x 6 edge slots
                const uint32_t d = q / 6u, sl = q - d * 6u;
                const uint32_t rw = s_run[d], g = s_g[d];
                const uint32_t start = rw & 0xffffu, c = rw >> 16;
                uint32_t head = (4u - (g & 3u)) & 3u;
                if (head > c) head = c;
                const uint32_t body = (c - head) & ~3u;
                const uint32_t tail = c - head - body;
                const uint32_t idx = (sl < 3u) ? sl : head + body + (sl - 3u);
                const bool on = (sl < 3u) ? (sl < head) : (sl - 3u < tail);
                B200_CHECK_AT(7, !on || ((size_t)g - gmis + idx < n && start + idx < (uint32_t)kTmaStageWords));
                if (on) st_stream(out_al + g + idx, s_stage[start + idx]);
            }
        }
        B200_STAMP(7);                                           // previous tile written (bulk copies in flight)
        if (TIMING && g_phase_dbg != nullptr && lane == 0 && (warp == 0 || warp == 8))
            g_phase_dbg[((size_t)dbg_tile * 2 + (warp >> 3)) * 16 + 9] = dbg_tile;
        prev_tile = have_cur ? tile : 0xFFFFFFFFu;
        tile = next;
        ++iter;
    }
    bulk_wait_all();                                             // every bulk copy has landed
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(kTmaTmemCols) : "memory");
}

}  // namespace b200sort

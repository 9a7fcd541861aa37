// radix_tma2.cuh -- k2: one onesweep pass, the Blackwell shape with 16384-key tiles (included by radix.cu).
//
// Replaces the lab's radix stage (SRM/lab.cu:47-87 radix_sort_kernel + :11-41 exlusiveScan) like the other pass
// kernels.  radix_tma.cuh showed that writing the staged tile out by TMA bulk copies takes that work off the SM's
// load/store pipe (59 % busy instead of 81 %) but pays for it in instructions: issuing 256 bulk copies, the run
// edges and the look-back sums cost as much per TILE whatever the tile holds.  This kernel spreads them over 1.6x
// the keys, and drops what made a larger tile impossible there:
//
//   * ranks are not kept at all.  The rank phase only COUNTS (one shared-memory atomicAdd per key whose result is
//     not used); the staging phase takes each key's position as the return value of a SECOND atomicAdd on the same
//     counters, which by then hold the first staged word of every (warp, digit) instead of zero.  That replaces the
//     rank's return value + the position lookup + the add (the same number of shared-memory wavefronts, ~10 fewer
//     instructions per key) and rests on the same lane-order property (and its self-test) as every kRankAdd shape.
//   * so only the KEYS are parked in tensor memory: 32 columns per warp and tile, 2 tiles x 4 warps per lane quadrant
//     = the 256 columns a CTA can have with two CTAs per SM.  A tile is two batches of 512 x 16 keys; a warp owns
//     1024 consecutive keys (batch, item, lane order = memory order).
//
// Iteration j of a persistent CTA (t = the tile counted now, p = the tile counted an iteration ago):
//      prefetch  one thread asks the TMA unit for the status rows of p's look-back (two bulk loads + mbarrier)
//   R  count t   batch 0 (loaded during W of the previous iteration), batch 1 (loaded while batch 0 is counted):
//                digit -> atomicAdd -> tcgen05.st
//   D  digits    group A: t's digit counts -> publish;  group B: p's prefix from the prefetched rows, then p's staging
//                layout (every run starts at a word congruent mod 4 to its first destination word) written into
//                p's counters as start positions
//   S  stage p   tcgen05.ld -> position = atomicAdd -> staging store
//   W  write p   batch 0 of the next tile is requested; per digit run one bulk copy for the 16-byte aligned interior,
//                the <= 3 + 3 edge words by ordinary stores
#pragma once
#include "radix_tma.cuh"

namespace b200sort {

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const int32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

constexpr int kT2Batch = 16;                                    // keys per thread and batch
constexpr int kT2Ipt = 2 * kT2Batch;
constexpr int kT2Threads = 512;
constexpr int kT2Tile = kT2Threads * kT2Ipt;                    // 16384 keys
constexpr int kT2Rows = 8;                                      // two warps per counter row, 16 bits each
constexpr int kT2StageWords = kT2Tile + kRadixBins * 6 + 64;    // every run padded to whole 16-byte chunks
constexpr int kT2Win1 = 11;                                     // nearest earlier tile rows of the group ...
constexpr int kT2Win2 = 6;                                      // ... and nearest group rows that fit beside the tile
constexpr int kT2TmemCols = 256;
constexpr size_t kT2SmemBytes =
    (size_t)kT2StageWords * 4
    + (size_t)2 * kT2Rows * kRadixBins * 4       // digit counters -> positions, this tile's and the previous tile's
    + (size_t)9 * kRadixBins * 4                 // x2: run {start, length}, destination, tile counts, in-group prefix; group prefix
    + (size_t)(kT2Win1 + kT2Win2) * kRadixBins * 4
    + 256;
static_assert(kT2SmemBytes <= 115712, "two CTAs per SM");

template <int TIMING, int DEVN = 0>
__global__ void __launch_bounds__(kT2Threads, 2)
radix_onesweep_tma2_kernel(const int32_t *in_buf, int32_t *out_buf, int32_t *tmp_buf, size_t n, int pass,
                           RadixControl *ctl, uint32_t *status_cur, uint32_t *status_next, int follow_plan)
{
    constexpr int kTile = kT2Tile;
    constexpr int kRows = kT2Rows;
    constexpr int W = 8;                                        // status rows in flight per thread (global walks)
    constexpr uint32_t kNone = 0xFFFFFFFFu;

    extern __shared__ __align__(128) unsigned char smem_tma2[];
    int32_t  *s_stage  = reinterpret_cast<int32_t *>(smem_tma2);                         // [kT2StageWords]
    uint32_t *s_table  = reinterpret_cast<uint32_t *>(s_stage + kT2StageWords);          // [2][kRows][256]
    uint32_t *s_run    = s_table + 2 * kRows * kRadixBins;       // [2][256] first staged word | keys to write << 16
    uint32_t *s_g      = s_run + 2 * kRadixBins;                 // [2][256] first destination word of the run
    uint32_t *s_ptot   = s_g + 2 * kRadixBins;                   // [2][256] the tile's digit counts, as published
    uint32_t *s_pin    = s_ptot + 2 * kRadixBins;                // [2][256] in-group prefix, if the tile summed its group
    uint32_t *s_gprev  = s_pin + 2 * kRadixBins;                 // [256] the previous tile's prefix over the earlier groups
    uint32_t *s_win1   = s_gprev + kRadixBins;                   // [kT2Win1][256] tile rows before the previous tile
    uint32_t *s_win2   = s_win1 + kT2Win1 * kRadixBins;          // [kT2Win2][256] group rows before its group
    uint32_t *s_misc   = s_win2 + kT2Win2 * kRadixBins;          // [0..7] warp sums, [8] ticket, [10] tmem base, [16..17] mbarrier,
                                                                 // [20] key count, [21] tile count
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int32_t *in = in_buf;
    int32_t *out = out_buf;
    if (follow_plan) {
        if (ctl->skip[pass]) {
            const size_t n_now = DEVN ? (size_t)ctl->n_dev : n;
            const size_t tl = (n_now + kTile - 1) / kTile;
            const size_t rows = tl + (tl + kLookGroup - 1) / kLookGroup;
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
    const uint32_t wofs = warp * (32 * kT2Ipt) + lane;           // a warp owns 1024 consecutive keys of the tile
    // word offset of `out` inside its 16-byte chunk: word g of the array is word g + gmis of the aligned base
    const uint32_t gmis = (uint32_t)((reinterpret_cast<uintptr_t>(out) >> 2) & 3u);
    int32_t *out_al = out - gmis;
    const uint32_t stage_s = smem_u32(s_stage);
    auto pair_bar = [&]() { bar_sync(3 + (warp >> 1), 64); };    // the two warps that share a counter row

    // ---- tensor memory: 256 columns; this warp owns lanes 32*(warp%4).., columns (warp/4)*64 + half*32 + batch*16.. --
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&s_misc[10])), "n"(kT2TmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        uint4 *z = reinterpret_cast<uint4 *>(s_table);
        for (uint32_t i = tid; i < 2 * kRows * kRadixBins / 4; i += kT2Threads) z[i] = make_uint4(0, 0, 0, 0);
    }
    const uint32_t mbar = smem_u32(&s_misc[16]);
    if (tid == 0) {
        s_misc[8] = atomicAdd(&ctl->ticket[pass], 1u);
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the key count and the tile count live in shared memory (re-read where needed: registers are scarce)
        const uint32_t n32 = DEVN ? ctl->n_dev : (uint32_t)n;
        s_misc[20] = n32;
        s_misc[21] = (uint32_t)(((size_t)n32 + kTile - 1) / kTile);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    auto n_f = [&]() -> size_t { return DEVN ? (size_t)reinterpret_cast<volatile uint32_t *>(s_misc)[20] : n; };
    auto tiles_f = [&]() -> size_t { return (size_t)reinterpret_cast<volatile uint32_t *>(s_misc)[21]; };
    const uint32_t tmem_base = s_misc[10];
    const uint32_t tmem_warp = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 64u;

    int32_t ka[kT2Batch], kb[kT2Batch];                          // batch 0 / batch 1 of the tile being counted
    auto load_batch = [&](uint32_t t, int batch, int32_t (&k)[kT2Batch]) {
        const size_t tile_base = (size_t)t * kTile;
        const size_t n_now = n_f();
        const uint32_t valid = (n_now - tile_base < (size_t)kTile) ? (uint32_t)(n_now - tile_base) : (uint32_t)kTile;
        const uint32_t o = wofs + batch * (32 * kT2Batch);
        const int32_t *src = in + tile_base + o;
        if (valid == (uint32_t)kTile) {
#pragma unroll
            for (int i = 0; i < kT2Batch; ++i) k[i] = ld_stream(src + i * 32);
        } else {
#pragma unroll
            for (int i = 0; i < kT2Batch; ++i) k[i] = (o + i * 32 < valid) ? ld_stream(src + i * 32) : 0x7FFFFFFF;
        }
    };
    // one shared-memory atomicAdd per key on the warp's half of its counter row.  COUNT: the result is not used (rank
    // phase); otherwise it is the key's staged position, and the key goes there.  A hot digit (the histogram kernel
    // found one value holding > 1/8 of the keys, or a quarter of the warp's first keys agree with lane 0's) is handled
    // with one ballot and ONE atomic per instruction, so skewed / sorted inputs do not serialise on one address.
    auto sweep = [&](const int32_t (&k)[kT2Batch], uint32_t *wt, bool count_only) {
        const uint32_t d0 = digit_of(k[0], shift, flip);
        const uint32_t agree = __ballot_sync(0xffffffffu, d0 == __shfl_sync(0xffffffffu, d0, 0));
        const uint32_t hot_word = follow_plan ? ctl->hot[pass] : 0u;
        const bool hot = hot_word != 0 || __popc(agree) >= 8;
        if (!hot) {
            // all the atomics first, then the stores: a store between two atomics would order them (the compiler
            // cannot know that the staging area and the counters do not alias) and expose every atomic's latency
            uint32_t pos[kT2Batch];
#pragma unroll
            for (int i = 0; i < kT2Batch; ++i) pos[i] = atomicAdd(wt + digit_of(k[i], shift, flip), 1u << sh);
            if (!count_only) {
#pragma unroll
                for (int i = 0; i < kT2Batch; ++i) {
                    const uint32_t q = (pos[i] >> sh) & 0xffffu;
                    B200_CHECK_AT(11, q < (uint32_t)kT2StageWords);
                    s_stage[q] = k[i];
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < kT2Batch; ++i) {
                const uint32_t d = digit_of(k[i], shift, flip);
                const uint32_t hd = hot_word ? hot_word - 1u : __shfl_sync(0xffffffffu, d, 0);
                const bool same = (d == hd);
                const uint32_t sm = __ballot_sync(0xffffffffu, same);
                const uint32_t leader = (uint32_t)(__ffs(sm) - 1) & 31u;
                uint32_t r = 0;
                if (!same || lane == leader)
                    r = (atomicAdd(wt + d, (same ? (uint32_t)__popc(sm) : 1u) << sh) >> sh) & 0xffffu;
                const uint32_t r0 = __shfl_sync(0xffffffffu, r, leader);
                if (same) r = r0 + __popc(sm & lt);
                if (!count_only) {
                    B200_CHECK_AT(11, r < (uint32_t)kT2StageWords);
                    s_stage[r] = k[i];
                }
            }
        }
    };

    uint32_t tile = s_misc[8];
    uint32_t prev_tile = kNone;
    uint32_t win_parity = 0;
    bool edge_tile = false;                                      // the tile staged last iteration still has its run edges to write
    __syncthreads();                                             // s_misc[8] is rewritten inside the loop
    if (tile < tiles_f()) load_batch(tile, 0, ka);
    const uint32_t digit_base = in_a ? 0u : ctl->base[pass][bd];
    uint32_t iter = 0;

    while (tile < tiles_f() || prev_tile != kNone || edge_tile) {
        const bool have_cur = tile < tiles_f();
        const bool have_prev = prev_tile != kNone;
        const uint32_t cb = iter & 1;                            // counters / tensor-memory half of `tile`
        uint32_t *tab_cur = s_table + cb * kRows * kRadixBins;
        uint32_t *tab_prev = s_table + (cb ^ 1) * kRows * kRadixBins;
        const uint32_t dbg_tile = have_cur ? tile : have_prev ? prev_tile : (uint32_t)tiles_f();   // edges-only iteration: a dummy row
        B200_STAMP(0);
        // ---- the previous tile's look-back rows are fetched into shared memory underneath the counting ----------
        uint32_t have1 = 0, have2 = 0;                           // rows fetched (tile rows, group rows)
        if (have_prev) {
            const uint32_t group = prev_tile / kLookGroup, r = prev_tile % kLookGroup;
            const bool last_of_group = (r == kLookGroup - 1) || ((size_t)prev_tile + 1 == tiles_f());
            have1 = last_of_group ? 0u : (r < (uint32_t)kT2Win1 ? r : (uint32_t)kT2Win1);
            have2 = group < (uint32_t)kT2Win2 ? group : (uint32_t)kT2Win2;
            if (tid == kRadixBins && have1 + have2 > 0) {
                fence_proxy_async_smem();                        // last iteration's reads of the windows are done (SYNC2)
                mbar_expect_tx(mbar, (have1 + have2) * kRadixBins * 4);
                if (have1) bulk_load(smem_u32(s_win1), status_cur + ((size_t)prev_tile - have1) * kRadixBins, have1 * kRadixBins * 4, mbar);
                if (have2) bulk_load(smem_u32(s_win2), status_cur + (tiles_f() + group - have2) * kRadixBins, have2 * kRadixBins * 4, mbar);
            }
        }
        // ---- R: count `tile`, park its keys in tensor memory ---------------------------------------------------
        if (have_cur) {
            load_batch(tile, 1, kb);                             // lands while batch 0 is counted
            uint32_t *wt = tab_cur + (warp >> 1) * kRadixBins;
            sweep(ka, wt, true);
            tmem_st16(tmem_warp + cb * 32u, ka);
            sweep(kb, wt, true);
            tmem_st16(tmem_warp + cb * 32u + 16u, kb);
        }
        B200_STAMP(1);                                           // counted and parked
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
                const bool last_of_group = (r == kLookGroup - 1) || ((size_t)tile + 1 == tiles_f());
                uint32_t *row = status_cur + (size_t)tile * kRadixBins + tid;
                st_relaxed_gpu(row, (r == 0 ? kFlagIncl : kFlagLocal) | total);
                if (status_next != nullptr) {
                    status_next[(size_t)tile * kRadixBins + tid] = 0;
                    if (last_of_group) status_next[(tiles_f() + group) * kRadixBins + tid] = 0;
                }
                s_ptot[cb * kRadixBins + tid] = total;
                if (last_of_group) {
                    const uint32_t p_in = (r > 0) ? walk_back<16>(row - kRadixBins, r) : 0u;
                    if (r > 0) st_relaxed_gpu(row, kFlagIncl | (p_in + total));
                    uint32_t *grow = status_cur + (tiles_f() + group) * kRadixBins + tid;
                    st_relaxed_gpu(grow, (group == 0 ? kFlagIncl : kFlagLocal) | (p_in + total));
                    s_pin[cb * kRadixBins + tid] = p_in;
                }
            }
            if (have_prev) {
                // ... the previous tile's prefix over the earlier GROUPS (group B sums its tile rows meanwhile) ...
                const uint32_t pgroup = prev_tile / kLookGroup;
                if (have1 + have2 > 0) { mbar_wait(mbar, win_parity); win_parity ^= 1; }   // the fetched rows have landed
                uint32_t gprev = 0;
                if (pgroup > 0)
                    gprev = walk_back_prefetched<W>(s_win2 + tid, have2, status_cur + (tiles_f() + pgroup - 1) * kRadixBins + tid, pgroup);
                s_gprev[tid] = gprev;
                __syncwarp();
                __threadfence_block();
                bar_arrive(11, kT2Threads);
            }
            if (edge_tile) {
                // ... and the run edges of the tile that was staged an iteration ago (its interiors left by bulk copy
                // right after the staging): the <= 3 + 3 words of every run that do not fill a 16-byte chunk
                const uint32_t *run_e = s_run + cb * kRadixBins, *g_e = s_g + cb * kRadixBins;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const uint32_t q = tid + j * kRadixBins;     // 256 digits x 6 edge slots
                    const uint32_t d = q / 6u, sl = q - d * 6u;
                    const uint32_t rw = run_e[d], g = g_e[d];
                    const uint32_t start = rw & 0xffffu, c = rw >> 16;
                    uint32_t head = (4u - (g & 3u)) & 3u;
                    if (head > c) head = c;
                    const uint32_t body = (c - head) & ~3u;
                    const uint32_t tail = c - head - body;
                    const uint32_t idx = (sl < 3u) ? sl : head + body + (sl - 3u);
                    const bool on = (sl < 3u) ? (sl < head) : (sl - 3u < tail);
                    if (on) st_stream(out_al + g + idx, s_stage[start + idx]);
                }
            }
            B200_STAMP(3);                                       // group A done
        } else {
            if (have_prev) {
                // thread = digit: the previous tile's prefix (every row it needs was published an iteration ago)
                const uint32_t pb = cb ^ 1;
                const uint32_t p_total = s_ptot[pb * kRadixBins + bd];
                const uint32_t group = prev_tile / kLookGroup, r = prev_tile % kLookGroup;
                const bool last_tile = (size_t)prev_tile + 1 == tiles_f();
                const bool last_of_group = (r == kLookGroup - 1) || last_tile;
                uint32_t *row = status_cur + (size_t)prev_tile * kRadixBins + bd;
                uint32_t *grow = status_cur + (tiles_f() + group) * kRadixBins + bd;
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
                bar_sync(11, kT2Threads);                        // group A has walked the group rows
                const uint32_t gprev = s_gprev[bd];
                if (group > 0 && last_of_group) st_relaxed_gpu(grow, kFlagIncl | ((gprev + inprev + p_total) & kValueMask));
                B200_STAMP(10);                                  // previous tile resolved
                // ... and its staging layout: run d occupies whole 16-byte chunks, its first key sits at the word that
                // is congruent mod 4 to its first destination word; the counters become start positions
                const uint32_t g = digit_base + inprev + gprev + gmis;       // destination word (from out_al)
                B200_CHECK_AT(12, (size_t)g - gmis + (last_tile ? 0u : p_total) <= n_f());
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
                // slots past n (last tile only) carry INT_MAX: digit 255, counted behind every real key; they are
                // staged but never written
                uint32_t cw = p_total;
                if (last_tile && bd == kRadixBins - 1) cw -= (uint32_t)(tiles_f() * (size_t)kTile - n_f());
                s_run[pb * kRadixBins + bd] = start | (cw << 16);
                s_g[pb * kRadixBins + bd] = g;
            }
            B200_STAMP(3);                                       // group B done
        }
        bulk_wait_read_all();                                    // my bulk copies of the tile before are done READING the staging area
        __syncthreads();                                         // SYNC2: the previous tile's positions are final
        B200_STAMP(4);
        // The next tile's ticket is drawn now, AFTER the one phase whose length varies: from ticket to publication
        // every tile then takes the same stage + write + count time.
        if (tid == 0) s_misc[8] = atomicAdd(&ctl->ticket[pass], 1u);

        // ---- S: stage the previous tile: keys come back from tensor memory, positions from the second atomic ------
        if (have_prev) {
            const uint32_t tp = tmem_warp + (cb ^ 1) * 32u;
            uint32_t *wt = tab_prev + (warp >> 1) * kRadixBins;
            tmem_wait_st();
#pragma unroll
            for (int batch = 0; batch < 2; ++batch) {
                uint32_t u0[8], u1[8];
                tmem_ld8(tp + batch * 16, u0);
                tmem_ld8(tp + batch * 16 + 8, u1);
                tmem_wait_ld();
                int32_t k[kT2Batch];
#pragma unroll
                for (int i = 0; i < 8; ++i) { k[i] = (int32_t)u0[i]; k[8 + i] = (int32_t)u1[i]; }
                sweep(k, wt, false);
            }
            // the counters are cleared for the tile after next once nobody takes positions from them any more
            pair_bar();
            reinterpret_cast<uint4 *>(tab_prev + (warp >> 1) * kRadixBins)[(warp & 1) * 32 + lane] = make_uint4(0, 0, 0, 0);
            fence_proxy_async_smem();                            // staged keys -> visible to the bulk copies
        }
        B200_STAMP(5);                                           // staged
        __syncthreads();                                         // SYNC3: the staged tile is complete
        B200_STAMP(6);
        // batch 0 of the next tile is requested now and lands while the previous tile is written
        const uint32_t next = s_misc[8];
        if (next < tiles_f()) load_batch(next, 0, ka);

        // ---- W: write the previous tile: interiors by bulk copy, edges by ordinary stores --------------------
        if (have_prev) {
            const uint32_t *run_p = s_run + (cb ^ 1) * kRadixBins, *g_p = s_g + (cb ^ 1) * kRadixBins;
            if (lane < 16) {                                     // 16 warps x 16 lanes: thread = digit
                const uint32_t d = warp * 16 + lane;
                const uint32_t rw = run_p[d], g = g_p[d];
                const uint32_t start = rw & 0xffffu, c = rw >> 16;
                uint32_t head = (4u - (g & 3u)) & 3u;
                if (head > c) head = c;
                const uint32_t body = (c - head) & ~3u;
                B200_CHECK_AT(13, body == 0 || (((g + head) & 3u) == 0 && ((start + head) & 3u) == 0));
                B200_CHECK_AT(14, (size_t)g - gmis + c <= n_f() && start + c <= (uint32_t)kT2StageWords);
                if (body > 0) bulk_store(out_al + g + head, stage_s + (start + head) * 4u, body * 4u);
                bulk_commit();
            }
        }
        edge_tile = have_prev;                                   // its run edges are written by group A during the next digit phase
        B200_STAMP(7);                                           // previous tile written (bulk copies in flight)
        if (TIMING && g_phase_dbg != nullptr && lane == 0 && (warp == 0 || warp == 8))
            g_phase_dbg[((size_t)dbg_tile * 2 + (warp >> 3)) * 16 + 9] = dbg_tile;
        prev_tile = have_cur ? tile : kNone;
        tile = next;
        ++iter;
    }
    bulk_wait_all();                                             // every bulk copy has landed
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(kT2TmemCols) : "memory");
}

}  // namespace b200sort

// EXPERIMENTAL scan kernel on the LEVEL formulation of the block pass (csrc/ta_block.cuh: block_window_minmax,
// BlockLevel<T, N>): both label widths, flag 0x4000 | 0x10000 (TA_PAIR_PATH=level).
//
// Why: in scan_block_kernel every lane of a warp pays as many mask builds as the most crowded block of the warp holds
// labels (warp maximum 3.98 against a lane mean of 1.85 on C3-like tissue), and the 41 % of the blocks whose window is one
// label still build a mask.  Here the label count is a property of a LIST:
//
//   P1   one 8 x 4 x 2 block per thread (uint32: two segments wide, 128 per brick): min / max label of its window.  Equal: closed-form moments, merged per warp.
//        Different: the block goes to list 2 with both labels.
//   P2   list 2, full warps, one block per lane: one fused pass over the window rows builds the masks of both labels;
//        their moments and their pair go through warp merges into the per-brick tables.  A window position covered by
//        neither label names a third one -> list 3.
//   P3   list 3: the masks of the first two labels come from a stash in shared memory where P2 parked them (the first
//        LV_STASH blocks of the list; beyond it: fused masks of all three labels), one mask build for the third label;
//        emitted: the third label's moments and its pairs with the other two.  Lanes whose window is still uncovered
//        add one label at a time (one more mask, its moments, its pairs with every older label) up to MAXL labels --
//        lists 4, 5, ... would hold a handful of blocks per brick, so they are steps of this pass instead of phases of
//        their own (a phase costs a block barrier and leaves most warps idle).
//   PF   blocks still uncovered after MAXL labels: per-voxel path inside the warp that found them (its 32 lanes share the
//        voxels of one block; no list, no barrier), restricted to what the steps could not emit (a label outside the
//        block's known set is involved).
//
// Cost model and expected gain: DESIGN.md section 6.  STATUS: exact on the CPU emulation of the CUDA execution model
// (tests/host/kernel_emu_check.cpp) and, for the block arithmetic, on the host (tests/host/block_level_check.cu); compiles
// for sm_100a; written after the round's GPU budget was spent, so it has NOT run on a GPU and is NOT part of the product
// build (-DTA_WITH_BLOCK_KERNEL).  First run: tools/r02_first_call.sh (TA_PAIR_PATH=level; level_simple: plain atomics
// instead of warp merges; level_pf: with an L2 prefetch of the next brick's tile).  Build knobs: -DTA_LEVEL_MAXL, -DTA_LEVEL_STASH, -DTA_LEVEL_MINB, -DTA_DEFAULT_LEVEL.
#pragma once
#include "ta_scan_block.cuh"

// dynamic counters for the CPU emulation (tests/host/kernel_emu_check.cpp --stats); nothing in a CUDA build
#ifndef TA_STAT
#define TA_STAT(which, n) ((void)0)
#endif

namespace ta {

#ifndef TA_LEVEL_MAXL
#define TA_LEVEL_MAXL 4            // labels per block handled by bit algebra; more -> per-voxel path for the rest
#endif
constexpr int LV_MAXL = TA_LEVEL_MAXL;
#ifndef TA_LEVEL_STASH
#define TA_LEVEL_STASH 96          // list-3 blocks per brick whose two masks P2 parks for P3 (multiple of 32; 0: P3 rebuilds)
#endif
constexpr int LV_STASH = TA_LEVEL_STASH;
static_assert(LV_STASH % 32 == 0 && LV_STASH <= NTHREADS, "whole warp rounds");
#ifndef TA_LEVEL_MINB
#define TA_LEVEL_MINB 3            // CTAs per SM the register budget is cut for: 3 -> 80 registers (some spills), 2 -> 128
#endif

template <typename T> constexpr size_t scan_level_smem_bytes() {
    return scan_block_smem_bytes<T>() +
           256 * 4 +                               // packed row moments of every byte
           NTHREADS * LV_MAXL * 4 +                // known labels per block
           2 * NTHREADS * 2 +                      // list 2 and list 3 (block ids)
           (size_t)LV_STASH * LV_STATE_WORDS * 8;  // masks of the first LV_STASH blocks of list 3, parked by P2 for P3
}

// block id (0 .. NBLK - 1) -> first segment, first row and first plane inside the brick
template <typename T>
__device__ __forceinline__ void level_block_origin(int blk, int& fs, int& m0, int& s0) {
    constexpr int NFB = LvBlk<T>::NFB;
    fs = (blk % NFB) * LvBlk<T>::BSEGS;
    m0 = ((blk / NFB) % (BM / BLK_M)) * BLK_M;
    s0 = (blk / (NFB * (BM / BLK_M))) * BLK_S;
}

// per-voxel path for one voxel of a block whose window holds labels outside `known[0 .. LV_MAXL - 1]`: only
// contributions that involve such a label (everything among the known labels was emitted by the levels)
template <typename T>
__device__ __noinline__ void level_fallback_voxel(const BrickShared<T> sh, const LabelTable lt, const PairTable pt, const T* p,
                                                  uint32_t f, uint32_t m, uint32_t s, const uint32_t* known, u64 gF0, u64 gM0,
                                                  u64 gS0, bool do_mom, bool do_p6, bool do_w18) {
    constexpr int ROWE = ROWV * Vox<T>::SEG, PLANEE = (BM + 2) * ROWE;
    const uint32_t a = p[0];
    bool a_in = false;
#pragma unroll
    for (int i = 0; i < LV_MAXL; ++i) a_in = a_in || (known[i] == a);
    if (do_mom && !a_in) {
        uint32_t v[LT_FIELDS] = {1u, f, m, s, f * f, f * m, f * s, m * m, m * s, s * s, f, m, s, f, m, s};
        label_add<T>(sh, lt, pt.status, a, v, gF0, gM0, gS0);
    }
    if (!(do_p6 || do_w18)) return;
    constexpr int offs[18] = {1, ROWE, PLANEE, -1, -ROWE, -PLANEE, -ROWE - 1, -ROWE + 1, ROWE - 1, ROWE + 1,
                              -PLANEE - 1, -PLANEE + 1, PLANEE - 1, PLANEE + 1,
                              -PLANEE - ROWE, -PLANEE + ROWE, PLANEE - ROWE, PLANEE + ROWE};
#pragma unroll 1
    for (int k = 0; k < 18; ++k) {
        const uint32_t b = p[offs[k]];
        if (b == a) continue;
        if (a_in) {
            bool b_in = false;
#pragma unroll
            for (int i = 0; i < LV_MAXL; ++i) b_in = b_in || (known[i] == b);
            if (b_in) continue;
        }
        if (do_p6 && k < 3) pair_add<T>(sh, pt, a, b, 2 * k + (a < b ? 0 : 1), 1u);
        if (do_w18) {
            bool seen = false;
            for (int q = 0; q < k; ++q) seen |= ((uint32_t)p[offs[q]] == b);
            if (!seen) pair_add<T>(sh, pt, a, b, 6, 1u);
        }
    }
}

// One label row per lane in packed form.  A lane's row holds at most one block (64 voxels, brick-local coordinates
// f < 128, m < 16, s < 8), so the ten sums of a whole WARP fit seven words and the m / s boxes one bit mask:
//   w0 = n [12 bits] | sf << 12 [18]     w1 = sm [15] | sss << 15 [17]     w2 = ss [14] | sms << 14 [18]
//   w3 = sff   w4 = sfm   w5 = sfs   w6 = smm          (warp maxima: n 2048, sf 252 928, sm 30 720, sss 100 352, ss 14 336,
//                                                        sms 215 040 -- each below its field)
//   w7 = f min   w8 = f max   w9 = 1 << m min | 1 << m max | (1 << s min | 1 << s max) << 16
// The steps pack their rows as soon as the moments are known and merge afterwards, so that the masks of a block are not
// live (and spilled) across the merge loops.
constexpr int LV_ROW = 10;
__device__ __forceinline__ void level_pack_row(const uint32_t v[LT_FIELDS], uint32_t w[LV_ROW]) {
    w[0] = v[0] | (v[1] << 12); w[1] = v[2] | (v[9] << 15); w[2] = v[3] | (v[8] << 14);
    w[3] = v[4]; w[4] = v[5]; w[5] = v[6]; w[6] = v[7];
    w[7] = v[10]; w[8] = v[13];
    w[9] = (1u << v[11]) | (1u << v[14]) | (((1u << v[12]) | (1u << v[15])) << 16);
}
__device__ __forceinline__ void level_unpack_row(const uint32_t w[LV_ROW], uint32_t u[LT_FIELDS]) {
    u[0] = w[0] & 0xFFFu; u[1] = w[0] >> 12; u[2] = w[1] & 0x7FFFu; u[9] = w[1] >> 15;
    u[3] = w[2] & 0x3FFFu; u[8] = w[2] >> 14; u[4] = w[3]; u[5] = w[4]; u[6] = w[5]; u[7] = w[6];
    u[10] = w[7]; u[13] = w[8];
    u[11] = (uint32_t)__ffs(w[9] & 0xFFFFu) - 1u; u[14] = 31u - (uint32_t)__clz(w[9] & 0xFFFFu);
    u[12] = (uint32_t)__ffs(w[9] >> 16) - 1u; u[15] = 31u - (uint32_t)__clz(w[9] >> 16);
}

// The label row of a lane goes to the per-brick table.  MERGE: warp merge (all 32 lanes call; has = false: nothing to
// add): uniform loop over the distinct labels of the warp, 10 full-mask redux each, the group leaders add in one SIMT
// pass after the loop.  Otherwise: plain atomics per lane.
template <typename T, bool MERGE>
__device__ __forceinline__ void level_put_label(const BrickShared<T>& sh, const LabelTable& lt, uint32_t* status, bool has,
                                                uint32_t L, const uint32_t w[LV_ROW], u64 gF0, u64 gM0, u64 gS0, int lane) {
    uint32_t u[LT_FIELDS];
    if (!MERGE) {
        if (has) { level_unpack_row(w, u); label_add<T>(sh, lt, status, L, u, gF0, gM0, gS0); }
        return;
    }
    unsigned pending = __ballot_sync(0xffffffffu, has);
    uint32_t tot[LV_ROW];
    bool am_leader = false;
    while (pending) {
        if (lane == 0) TA_STAT(10, 1);
        const int leader = __ffs(pending) - 1;
        const uint32_t Lk = __shfl_sync(0xffffffffu, L, leader);
        const bool mine = has && (L == Lk);
        const bool lead = (lane == leader);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const uint32_t r = __reduce_add_sync(0xffffffffu, mine ? w[i] : 0u);
            if (lead) tot[i] = r;
        }
        uint32_t r = __reduce_min_sync(0xffffffffu, mine ? w[7] : 0xFFFFFFFFu);
        if (lead) tot[7] = r;
        r = __reduce_max_sync(0xffffffffu, mine ? w[8] : 0u);
        if (lead) tot[8] = r;
        r = __reduce_or_sync(0xffffffffu, mine ? w[9] : 0u);
        if (lead) tot[9] = r;
        am_leader = am_leader || lead;
        pending &= ~__ballot_sync(0xffffffffu, mine);
    }
    if (am_leader) {
        level_unpack_row(tot, u);
        label_add<T>(sh, lt, status, L, u, gF0, gM0, gS0);
    }
}
template <typename T, bool MERGE>
__device__ __forceinline__ void level_put_pair(const BrickShared<T>& sh, const PairTable& pt, bool has, uint32_t a, uint32_t b,
                                               const uint32_t inc[PT_WORDS], int lane) {
    if (MERGE) block_merge_pair<T>(sh, pt, has ? Vox<T>::key(a, b) : Vox<T>::PEMPTY, inc, lane);
    else if (has) pair_add_packed<T>(sh, pt, Vox<T>::key(a, b), inc);
}

// moments of slot I of a block as a packed row in brick coordinates; false: nothing to add
template <typename T, int CAP>
__device__ __forceinline__ bool level_slot_row(const BlockLevel<T, CAP>& b, int i, const uint32_t* momtab, bool wanted, uint32_t bF,
                                               uint32_t bM, uint32_t bS, uint32_t w[LV_ROW]) {
    uint32_t v[LT_FIELDS];
    const bool has = wanted && b.label_moments(i, momtab, v);
    if (has) { block_shift_moments(v, bF, bM, bS); level_pack_row(v, w); }
    return has;
}

// what a step adds for slot I of a block: the label's moments and its pairs with every older slot (all lanes call).
// Everything that needs the masks is computed first, the merges follow.
template <typename T, int CAP, int I, bool MERGE>
__device__ __forceinline__ void level_emit_slot(const BrickShared<T>& sh, const ScanParams& P, const LabelTable& lt, const PairTable& pt,
                                                const uint32_t* momtab, const BlockLevel<T, CAP>& b, bool active, uint32_t bF,
                                                uint32_t bM, uint32_t bS, u64 gF0, u64 gM0, u64 gS0, int lane) {
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    uint32_t w[LV_ROW], inc[I > 0 ? I : 1][PT_WORDS];
    bool hasp[I > 0 ? I : 1];
    const bool has = level_slot_row<T, CAP>(b, I, momtab, active && do_mom, bF, bM, bS, w);
#pragma unroll
    for (int j = 0; j < I; ++j) hasp[j] = active && (do_p6 || do_w18) && b.pair_increments(I, j, do_p6, do_w18, inc[j]);
    level_put_label<T, MERGE>(sh, lt, pt.status, has, b.lab[I], w, gF0, gM0, gS0, lane);
    if (do_p6 || do_w18) {
#pragma unroll
        for (int j = 0; j < I; ++j) level_put_pair<T, MERGE>(sh, pt, hasp[j], b.lab[I], b.lab[j], inc[j], lane);
    }
}

// extension steps I .. MAXL - 1 of P3: lanes whose window is still uncovered add one label each; the loop ends for the
// whole warp as soon as no lane needs another step, so the merges stay full-mask
template <typename T, int I, bool MERGE>
__device__ __forceinline__ void level_extend(const BrickShared<T>& sh, const ScanParams& P, const LabelTable& lt, const PairTable& pt,
                                             const uint32_t* momtab, uint32_t* known, BlockLevel<T, LV_MAXL>& b,
                                             bool& more, uint32_t& next, int blk, int fs, int m0, int s0, uint32_t bF, uint32_t bM,
                                             uint32_t bS, u64 gF0, u64 gM0, u64 gS0, int lane) {
    if constexpr (I < LV_MAXL) {
        if (!__ballot_sync(0xffffffffu, more)) return;
        const bool act = more;
        if (lane == 0) TA_STAT(7, 1);
        if (act) TA_STAT(8, 1);
        if (act) {
            known[blk * LV_MAXL + I] = next;
            more = !b.template extend<I>(sh.tile, fs, m0, s0, next, next);
        } else {
            b.template clear_slot<I>();
        }
        level_emit_slot<T, LV_MAXL, I, MERGE>(sh, P, lt, pt, momtab, b, act, bF, bM, bS, gF0, gM0, gS0, lane);
        level_extend<T, I + 1, MERGE>(sh, P, lt, pt, momtab, known, b, more, next, blk, fs, m0, s0, bF, bM, bS, gF0, gM0, gS0, lane);
    }
}

// P2 (N = 2) and P3 (N = 3) over their lists.  Every warp runs the same number of rounds for all its lanes (lanes beyond
// the list end idle with active = false), so the merges are full-mask.
template <typename T, int N, bool MERGE>
__device__ __forceinline__ void level_pass(const BrickShared<T>& sh, const ScanParams& P, const LabelTable& lt, const PairTable& pt,
                                           const uint32_t* momtab, uint32_t* known, unsigned short* lists, u64* stash, int F0, int M0,
                                           int S0, u64 gF0, u64 gM0, u64 gS0, int tid) {
    static_assert(N == 2 || N == 3, "list 2 or list 3");
    constexpr int SEG = Vox<T>::SEG;
    constexpr int CAP = (N == 2) ? 2 : LV_MAXL;
    const int lane = tid & 31;
    const int count = (int)sh.ctr[N - 2];
    const unsigned short* list = lists + (N - 2) * NTHREADS;
    unsigned short* next_list = lists + (N - 1) * NTHREADS;       // list 3 (P2 only)
    for (int base = tid - lane; base < count; base += NTHREADS) {
        const int q = base + lane;
        const bool active = q < count;
        if (lane == 0) TA_STAT(N == 2 ? 4 : 6, 1);
        if (active) TA_STAT(N == 2 ? 3 : 5, 1);
        const int blk = active ? (int)list[q] : 0;
        int fs, m0, s0;
        level_block_origin<T>(blk, fs, m0, s0);
        const int nvf = min(LvBlk<T>::BW, (int)P.nf - (F0 + fs * SEG)), nvm = min(BLK_M, (int)P.nm - (M0 + m0)),
                  nvs = min(BLK_S, (int)P.own_hi - (S0 + s0));
        const uint32_t bF = (uint32_t)(fs * SEG), bM = (uint32_t)m0, bS = (uint32_t)s0;
        BlockLevel<T, CAP> b;
        uint32_t L[N];
#pragma unroll
        for (int i = 0; i < N; ++i) L[i] = known[blk * LV_MAXL + i];
        bool more = false;
        uint32_t next = 0u;
        if (N == 3 && LV_STASH > 0 && base < LV_STASH) {
            // a whole round of parked blocks: slots 0 and 1 from the stash, one mask build for the third label
            if (active) {
                b.load_state2(stash, q, LV_STASH, L[0], L[1], nvf, nvm, nvs);
                more = !b.template extend<(N == 3 ? 2 : 0)>(sh.tile, fs, m0, s0, L[N - 1], next);
            } else {
                b.clear();
            }
        } else if (active) {
            // list 2 carries (min, max) of the window: the borrow-free subtraction form of the compares
            more = !b.template build<N, N == 2>(sh.tile, fs, m0, s0, nvf, nvm, nvs, L, next);
        } else {
            b.clear();
        }
        if (N == 2) {
            if (more) {                                           // a third label: list 3; the first LV_STASH park their masks
                known[blk * LV_MAXL + 2] = next;
                const int pos = (int)atomicAdd(&sh.ctr[1], 1u);
                next_list[pos] = (unsigned short)blk;
                if (pos < LV_STASH) b.store_state2(stash, pos, LV_STASH > 0 ? LV_STASH : 1);
            }
            // both rows and the pair first: the masks are dead before the first merge loop
            const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
            uint32_t w0[LV_ROW], w1[LV_ROW], inc[PT_WORDS];
            const bool has0 = level_slot_row<T, CAP>(b, 0, momtab, active && do_mom, bF, bM, bS, w0);
            const bool has1 = level_slot_row<T, CAP>(b, 1, momtab, active && do_mom, bF, bM, bS, w1);
            const bool hasp = active && (do_p6 || do_w18) && b.pair_increments(1, 0, do_p6, do_w18, inc);
            const uint32_t L0 = b.lab[0], L1 = b.lab[1];
            level_put_label<T, MERGE>(sh, lt, pt.status, has0, L0, w0, gF0, gM0, gS0, lane);
            level_put_label<T, MERGE>(sh, lt, pt.status, has1, L1, w1, gF0, gM0, gS0, lane);
            if (do_p6 || do_w18) level_put_pair<T, MERGE>(sh, pt, hasp, L1, L0, inc, lane);
        } else {
            level_emit_slot<T, CAP, 2, MERGE>(sh, P, lt, pt, momtab, b, active, bF, bM, bS, gF0, gM0, gS0, lane);
            if constexpr (N == 3)
                level_extend<T, 3, MERGE>(sh, P, lt, pt, momtab, known, b, more, next, blk, fs, m0, s0, bF, bM, bS, gF0, gM0, gS0, lane);
            // PF, inside the warp that found them (no list, no block barrier): blocks with labels beyond their known set
            // take the per-voxel path for what the steps could not emit; the 32 lanes share the voxels of one block
            unsigned fm = __ballot_sync(0xffffffffu, more);
            if (fm) __syncwarp();                                  // the known labels of those blocks were written by their lanes
            while (fm) {
                const int src = __ffs(fm) - 1;
                fm &= fm - 1u;
                const int cblk = __shfl_sync(0xffffffffu, blk, src);
                if (lane == 0) TA_STAT(9, 1);
                int cfs, cm0, cs0;
                level_block_origin<T>(cblk, cfs, cm0, cs0);
                constexpr int BW = LvBlk<T>::BW, BV = BW * BLK_M * BLK_S, ROWE = ROWV * SEG;
                const T* tileT = reinterpret_cast<const T*>(sh.tile);
                for (int w = lane; w < BV; w += 32) {
                    const int df = w % BW, dm = (w / BW) % BLK_M, ds = w / (BW * BLK_M);
                    const uint32_t f = (uint32_t)(cfs * SEG + df), m = (uint32_t)(cm0 + dm), sp = (uint32_t)(cs0 + ds);
                    if (F0 + (int)f >= (int)P.nf || M0 + (int)m >= (int)P.nm || S0 + (int)sp >= (int)P.own_hi) continue;
                    const T* p = tileT + (size_t)((sp + 1) * (BM + 2) + (m + 1)) * ROWE + SEG + 1 + f;     // shifted tile
                    level_fallback_voxel<T>(sh, lt, pt, p, f, m, sp, known + cblk * LV_MAXL, gF0, gM0, gS0, P.flags & 1u, P.flags & 2u,
                                            P.flags & 4u);
                }
            }
        }
    }
}

// MERGE = true: warp merges for the table updates; false: plain atomics per block (to bisect against).
template <typename T, bool MERGE>
__global__ void __launch_bounds__(NTHREADS, TA_LEVEL_MINB)
scan_level_kernel(ScanParams P, LabelTable lt, PairTable pt, const __grid_constant__ CUtensorMap tmap) {
    typedef typename Vox<T>::PKey PKey;
    constexpr int SEG = Vox<T>::SEG, ROWE = ROWV * SEG, BF = NFS * SEG;
    static_assert(NTHREADS >= LvBlk<T>::NBLK, "at most one block per thread in P1");
    static_assert(LV_MAXL >= 3 && LV_MAXL <= 8, "labels per block handled by masks");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    BrickShared<T> sh;
    sh.tile = reinterpret_cast<uint4*>(smem_raw);
    sh.lt_key = reinterpret_cast<uint32_t*>(sh.tile + TILE_SEGS);
    sh.lt_val = sh.lt_key + LT_SLOTS;
    sh.pt_val = sh.lt_val + LT_SLOTS * LT_FIELDS;
    sh.pt_key = reinterpret_cast<PKey*>(sh.pt_val + PT_SLOTS * PT_WORDS);
    sh.ctr = reinterpret_cast<unsigned int*>(sh.pt_key + PT_SLOTS);
    uint32_t* momtab = reinterpret_cast<uint32_t*>(sh.ctr + 16);                   // [256]
    uint32_t* known = momtab + 256;                                                // [NTHREADS * MAXL] labels per block
    unsigned short* lists = reinterpret_cast<unsigned short*>(known + NTHREADS * LV_MAXL);   // [2][NTHREADS] block ids: list 2, list 3
    u64* stash = reinterpret_cast<u64*>(lists + 2 * NTHREADS);                     // [LV_STATE_WORDS][LV_STASH], 8-byte aligned
    const T* tileT = reinterpret_cast<const T*>(sh.tile);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const unsigned int total = (unsigned int)P.nbf * P.nbm * P.nbs;
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    const int nf = (int)P.nf, nm = (int)P.nm;

    for (int i = tid; i < LT_SLOTS; i += NTHREADS) sh.lt_key[i] = TA_EMPTY32;
    for (int i = tid; i < LT_SLOTS * LT_FIELDS; i += NTHREADS) {
        const int f = i % LT_FIELDS;
        sh.lt_val[i] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
    }
    for (int i = tid; i < PT_SLOTS; i += NTHREADS) sh.pt_key[i] = Vox<T>::PEMPTY;
    for (int i = tid; i < PT_SLOTS * PT_WORDS; i += NTHREADS) sh.pt_val[i] = 0u;
    for (int i = tid; i < 256; i += NTHREADS) momtab[i] = block_byte_moments_packed((uint32_t)i);

    uint64_t* tma_bar = reinterpret_cast<uint64_t*>(sh.ctr + 12);
    uint32_t tma_parity = 0u;
    const bool use_tma = P.use_tma && ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u) == 0u;
    if (use_tma && tid == 0) {
        mbar_init(tma_bar, 1u);
        TA_PTX("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == 0) sh.ctr[6] = atomicAdd(P.brick_counter, 1u);
    __syncthreads();

    for (unsigned iter = 0;; ++iter) {
        const unsigned int brick = sh.ctr[6 + (iter & 1u)];
        if (brick >= total) break;
        if (tid == 0) {
            const unsigned int nb = atomicAdd(P.brick_counter, 1u);
            sh.ctr[6 + ((iter + 1u) & 1u)] = nb;
            if (use_tma && (P.flags & 0x20000u) && nb < total) {
                // experiment (TA_PAIR_PATH=level_pf): ask the L2 for the NEXT brick's tile now (SASS UTMAPF.L2.3D), so that
                // its box copy finds the data on chip; DRAM is idle 95 % of the time
                const int nbf_ = nb % P.nbf, nbm_ = (nb / P.nbf) % P.nbm, nbs_ = nb / (P.nbf * P.nbm);
                TA_PTX("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                       :: "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(nbf_ * BF - SEG - 1), "r"(nbm_ * BM - 1),
                          "r"((int)P.own_lo + nbs_ * BS - 1) : "memory");
            }
            sh.ctr[0] = sh.ctr[1] = 0u;                                // list 2, list 3
        }
        const int bf = brick % P.nbf, bm = (brick / P.nbf) % P.nbm, bs = brick / (P.nbf * P.nbm);
        const int F0 = bf * BF, M0 = bm * BM, S0 = (int)P.own_lo + bs * BS;
        const u64 gF0 = (u64)F0, gM0 = (u64)M0, gS0 = (u64)((long long)S0 + P.slow_offset);

        // ---- phase A: the tile (ends with a block barrier), then the one-label tile shortcut ----------------------------
        block_stage_tile<T, 1>(sh, P, tmap, tma_bar, tma_parity, use_tma, F0, M0, S0, iter, brick, tid);
        if (block_uniform_tile<T, 1>(sh, P, lt, pt, F0, M0, S0, gF0, gM0, gS0, tid)) continue;

        // ---- P1: one block per thread, window min / max ------------------------------------------------------------------
        {
            int fs, m0, s0;
            level_block_origin<T>(tid % LvBlk<T>::NBLK, fs, m0, s0);
            const int nvf = min(LvBlk<T>::BW, nf - (F0 + fs * SEG)), nvm = min(BLK_M, nm - (M0 + m0)),
                      nvs = min(BLK_S, (int)P.own_hi - (S0 + s0));
            const bool valid = (tid < LvBlk<T>::NBLK && nvf > 0 && nvm > 0 && nvs > 0);
            uint32_t lo = 0u, hi = 0u;
            if (valid) block_window_minmax<T>(sh.tile, s0 * PLANEV + m0 * ROWV + (fs + 1), lo, hi);
            const bool one = valid && lo == hi, many = valid && lo != hi;
            if (tid == 0) TA_STAT(0, 1);
            if (valid) TA_STAT(1, 1);
            if (one) TA_STAT(2, 1);
            if (many) {                                       // list 2: both labels are labels of the window
                known[tid * LV_MAXL + 0] = lo;
                known[tid * LV_MAXL + 1] = hi;
            }
            // warp-aggregated append: one shared atomic per warp
            const unsigned mm = __ballot_sync(0xffffffffu, many);
            unsigned int base = 0u;
            if (lane == 0 && mm) base = atomicAdd(&sh.ctr[0], (unsigned int)__popc(mm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (many) lists[base + (unsigned int)__popc(mm & ((1u << lane) - 1u))] = (unsigned short)tid;
            uint32_t v[LT_FIELDS], w[LV_ROW];
            const bool has = one && do_mom;
            if (has) {
                block_uniform_moments((uint32_t)nvf, (uint32_t)nvm, (uint32_t)nvs, v);
                block_shift_moments(v, (uint32_t)(fs * SEG), (uint32_t)m0, (uint32_t)s0);       // block -> brick coordinates
                level_pack_row(v, w);
            }
            level_put_label<T, MERGE>(sh, lt, pt.status, has, lo, w, gF0, gM0, gS0, lane);
        }
        __syncthreads();

        // ---- P2, P3: the lists (list 3 is complete when P2 has ended) ---------------------------------------------------------
        level_pass<T, 2, MERGE>(sh, P, lt, pt, momtab, known, lists, stash, F0, M0, S0, gF0, gM0, gS0, tid);
        __syncthreads();
        level_pass<T, 3, MERGE>(sh, P, lt, pt, momtab, known, lists, stash, F0, M0, S0, gF0, gM0, gS0, tid);
        __syncthreads();

        // ---- flush the per-brick tables ------------------------------------------------------------------------------------
        block_flush_tables<T>(sh, lt, pt, gF0, gM0, gS0, tid);
        __syncthreads();
    }
}

}  // namespace ta

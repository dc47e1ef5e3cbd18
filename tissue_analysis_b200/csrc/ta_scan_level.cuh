// EXPERIMENTAL scan kernel on the LEVEL formulation of the block pass (csrc/ta_block.cuh: block_window_minmax,
// BlockLevel<T, N>): both label widths, flag 0x4000 | 0x10000 (TA_PAIR_PATH=level).
//
// Why: in scan_block_kernel every lane of a warp pays as many mask builds as the most crowded block of the warp holds
// labels (warp maximum 3.98 against a lane mean of 1.85 on C3-like tissue), and the 41 % of the blocks whose window is one
// label still build a mask.  Here the label count is a property of a LIST:
//
//   P1   one 8 x 4 x 2 block per thread: min / max label of its window.  Equal: closed-form moments, merged per warp.
//        Different: the block goes to list 2 with both labels.
//   PN   (N = 2 .. MAXL) list N is processed by full warps, one block per lane: one fused pass over the window rows builds
//        the masks of the N known labels; the level emits what it adds (N = 2: both labels' moments and their pair;
//        N > 2: the newest label's moments and its pairs with the older ones) through warp merges into the per-brick
//        tables; a window position covered by no known label names label N + 1 -> list N + 1.
//   PF   blocks still uncovered after level MAXL: per-voxel path, all threads share their voxels, restricted to what the
//        levels could not emit (a label outside the block's known set is involved).
//
// Cost model and expected gain: DESIGN.md section 6.  STATUS: exact on the CPU emulation of the CUDA execution model
// (tests/host/kernel_emu_check.cpp) and, for the block arithmetic, on the host (tests/host/block_level_check.cu); compiles
// for sm_100a; written after the round's GPU budget was spent, so it has NOT run on a GPU and is NOT part of the product
// build (-DTA_WITH_BLOCK_KERNEL).  First run: as in ta_scan_block.cuh with TA_PAIR_PATH=level (level_simple: plain
// atomics instead of warp merges).
#pragma once
#include "ta_scan_block.cuh"

namespace ta {

#ifndef TA_LEVEL_MAXL
#define TA_LEVEL_MAXL 4            // labels per block handled by bit algebra; more -> per-voxel path for the rest
#endif
constexpr int LV_MAXL = TA_LEVEL_MAXL;

template <typename T> constexpr size_t scan_level_smem_bytes() {
    return scan_block_smem_bytes<T>() +
           256 * 4 +                               // packed row moments of every byte
           NTHREADS * LV_MAXL * 4 +                // known labels per block
           LV_MAXL * NTHREADS * 2;                 // lists 2 .. MAXL and the fallback list (block ids)
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

// What a level hands to the tables, per lane: up to two label rows and up to MAXL - 1 pair rows.
template <typename T, bool MERGE>
__device__ __forceinline__ void level_emit_label(const BrickShared<T>& sh, const LabelTable& lt, uint32_t* status, bool has,
                                                 uint32_t L, uint32_t v[LT_FIELDS], uint32_t bF, uint32_t bM, uint32_t bS, u64 gF0,
                                                 u64 gM0, u64 gS0, int lane) {
    if (has) block_shift_moments(v, bF, bM, bS);                        // block -> brick coordinates
    if (MERGE) block_merge_label<T>(sh, lt, status, has, L, v, gF0, gM0, gS0, lane);
    else if (has) label_add<T>(sh, lt, status, L, v, gF0, gM0, gS0);
}
template <typename T, bool MERGE>
__device__ __forceinline__ void level_emit_pair(const BrickShared<T>& sh, const PairTable& pt, bool has, uint32_t a, uint32_t b,
                                                const uint32_t inc[PT_WORDS], int lane) {
    if (MERGE) block_merge_pair<T>(sh, pt, has ? Vox<T>::key(a, b) : Vox<T>::PEMPTY, inc, lane);
    else if (has) pair_add_packed<T>(sh, pt, Vox<T>::key(a, b), inc);
}

// Level N over list N.  Every warp runs the same number of rounds for all its lanes (lanes beyond the list end idle
// with has = false), so the merges are full-mask.
template <typename T, int N, bool MERGE>
__device__ __forceinline__ void level_pass(const BrickShared<T>& sh, const ScanParams& P, const LabelTable& lt, const PairTable& pt,
                                           const uint32_t* momtab, uint32_t* known, unsigned short* lists, int F0, int M0, int S0,
                                           u64 gF0, u64 gM0, u64 gS0, int tid) {
    constexpr int SEG = Vox<T>::SEG;
    const int lane = tid & 31;
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    const int count = (int)sh.ctr[N - 2];
    const unsigned short* list = lists + (N - 2) * NTHREADS;
    unsigned short* next_list = lists + (N - 1) * NTHREADS;
    for (int base = tid - lane; base < count; base += NTHREADS) {
        const int q = base + lane;
        const bool active = q < count;
        const int blk = active ? (int)list[q] : 0;
        const int fs = blk % NFS, m0 = ((blk / NFS) % (BM / BLK_M)) * BLK_M, s0 = (blk / (NFS * (BM / BLK_M))) * BLK_S;
        const int nvf = min(SEG, (int)P.nf - (F0 + fs * SEG)), nvm = min(BLK_M, (int)P.nm - (M0 + m0)),
                  nvs = min(BLK_S, (int)P.own_hi - (S0 + s0));
        const uint32_t bF = (uint32_t)(fs * SEG), bM = (uint32_t)m0, bS = (uint32_t)s0;
        BlockLevel<T, N> b;
        uint32_t L[N];
#pragma unroll
        for (int i = 0; i < N; ++i) L[i] = known[blk * LV_MAXL + i];
        bool more = false;
        uint32_t next = 0u;
        if (active) more = !b.build(sh.tile, fs, m0, s0, nvf, nvm, nvs, L, next);
        else b.clear();
        if (more) {
            // list N + 1 (for N == MAXL: the per-voxel list); its label slot exists only below MAXL
            if (N < LV_MAXL) known[blk * LV_MAXL + (N < LV_MAXL ? N : 0)] = next;
            next_list[atomicAdd(&sh.ctr[N - 1], 1u)] = (unsigned short)blk;
        }
        uint32_t v[LT_FIELDS], inc[PT_WORDS];
        if (N == 2) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool has = active && do_mom && b.label_moments(i, momtab, v);
                level_emit_label<T, MERGE>(sh, lt, pt.status, has, b.lab[i], v, bF, bM, bS, gF0, gM0, gS0, lane);
            }
            if (do_p6 || do_w18) {
                const bool has = active && b.pair_increments(0, 1, do_p6, do_w18, inc);
                level_emit_pair<T, MERGE>(sh, pt, has, b.lab[0], b.lab[1], inc, lane);
            }
        } else {
            const bool has = active && do_mom && b.label_moments(N - 1, momtab, v);
            level_emit_label<T, MERGE>(sh, lt, pt.status, has, b.lab[N - 1], v, bF, bM, bS, gF0, gM0, gS0, lane);
            if (do_p6 || do_w18) {
#pragma unroll
                for (int j = 0; j < N - 1; ++j) {
                    const bool hasp = active && b.pair_increments(N - 1, j, do_p6, do_w18, inc);
                    level_emit_pair<T, MERGE>(sh, pt, hasp, b.lab[N - 1], b.lab[j], inc, lane);
                }
            }
        }
    }
}

// levels 2 .. MAXL, one block barrier after each (list N + 1 is complete when level N has ended)
template <typename T, int N, bool MERGE>
__device__ __forceinline__ void level_passes(const BrickShared<T>& sh, const ScanParams& P, const LabelTable& lt, const PairTable& pt,
                                             const uint32_t* momtab, uint32_t* known, unsigned short* lists, int F0, int M0, int S0,
                                             u64 gF0, u64 gM0, u64 gS0, int tid) {
    level_pass<T, N, MERGE>(sh, P, lt, pt, momtab, known, lists, F0, M0, S0, gF0, gM0, gS0, tid);
    __syncthreads();
    if constexpr (N < LV_MAXL) level_passes<T, N + 1, MERGE>(sh, P, lt, pt, momtab, known, lists, F0, M0, S0, gF0, gM0, gS0, tid);
}

// MERGE = true: warp merges for the table updates; false: plain atomics per block (to bisect against).
template <typename T, bool MERGE>
__global__ void __launch_bounds__(NTHREADS, 3)
scan_level_kernel(ScanParams P, LabelTable lt, PairTable pt, const __grid_constant__ CUtensorMap tmap) {
    typedef typename Vox<T>::PKey PKey;
    constexpr int SEG = Vox<T>::SEG, ROWE = ROWV * SEG, BF = NFS * SEG;
    static_assert(NTHREADS == NFS * (BM / BLK_M) * (BS / BLK_S), "one block per thread");
    static_assert(LV_MAXL >= 2 && LV_MAXL <= 6, "sh.ctr[0 .. MAXL - 1] count lists 2 .. MAXL + 1; ctr[6 ..] is taken");

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
    unsigned short* lists = reinterpret_cast<unsigned short*>(known + NTHREADS * LV_MAXL);   // [MAXL][NTHREADS] block ids
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
            sh.ctr[6 + ((iter + 1u) & 1u)] = atomicAdd(P.brick_counter, 1u);
#pragma unroll
            for (int n = 0; n < LV_MAXL; ++n) sh.ctr[n] = 0u;          // lists 2 .. MAXL + 1
        }
        const int bf = brick % P.nbf, bm = (brick / P.nbf) % P.nbm, bs = brick / (P.nbf * P.nbm);
        const int F0 = bf * BF, M0 = bm * BM, S0 = (int)P.own_lo + bs * BS;
        const u64 gF0 = (u64)F0, gM0 = (u64)M0, gS0 = (u64)((long long)S0 + P.slow_offset);

        // ---- phase A: the tile (ends with a block barrier), then the one-label tile shortcut ----------------------------
        block_stage_tile<T>(sh, P, tmap, tma_bar, tma_parity, use_tma, F0, M0, S0, iter, brick, tid);
        if (block_uniform_tile<T>(sh, P, lt, pt, F0, M0, S0, gF0, gM0, gS0, tid)) continue;

        // ---- P1: one block per thread, window min / max ------------------------------------------------------------------
        {
            const int fs = tid % NFS, m0 = ((tid / NFS) % (BM / BLK_M)) * BLK_M, s0 = (tid / (NFS * (BM / BLK_M))) * BLK_S;
            const int nvf = min(SEG, nf - (F0 + fs * SEG)), nvm = min(BLK_M, nm - (M0 + m0)),
                      nvs = min(BLK_S, (int)P.own_hi - (S0 + s0));
            const bool valid = (nvf > 0 && nvm > 0 && nvs > 0);
            uint32_t lo = 0u, hi = 0u;
            if (valid) block_window_minmax<T>(sh.tile, s0 * PLANEV + m0 * ROWV + (fs + 1), lo, hi);
            const bool one = valid && lo == hi, many = valid && lo != hi;
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
            uint32_t v[LT_FIELDS];
            const bool has = one && do_mom;
            if (has) block_uniform_moments((uint32_t)nvf, (uint32_t)nvm, (uint32_t)nvs, v);
            level_emit_label<T, MERGE>(sh, lt, pt.status, has, lo, v, (uint32_t)(fs * SEG), (uint32_t)m0, (uint32_t)s0, gF0, gM0,
                                       gS0, lane);
        }
        __syncthreads();

        // ---- P2 .. PMAXL: the lists ----------------------------------------------------------------------------------------
        level_passes<T, 2, MERGE>(sh, P, lt, pt, momtab, known, lists, F0, M0, S0, gF0, gM0, gS0, tid);

        // ---- PF: blocks with labels beyond their known set: all threads share their voxels -------------------------------
        {
            const int ncrowded = (int)sh.ctr[LV_MAXL - 1];
            const unsigned short* crowded = lists + (LV_MAXL - 1) * NTHREADS;
            constexpr int BV = SEG * BLK_M * BLK_S;
            for (int q = tid; q < ncrowded * BV; q += NTHREADS) {
                const int blk = crowded[q / BV], w = q % BV;
                const int cfs = blk % NFS, cm0 = ((blk / NFS) % (BM / BLK_M)) * BLK_M, cs0 = (blk / (NFS * (BM / BLK_M))) * BLK_S;
                const int df = w % SEG, dm = (w / SEG) % BLK_M, ds = w / (SEG * BLK_M);
                const uint32_t f = (uint32_t)(cfs * SEG + df), m = (uint32_t)(cm0 + dm), sp = (uint32_t)(cs0 + ds);
                if (F0 + (int)f >= nf || M0 + (int)m >= nm || S0 + (int)sp >= (int)P.own_hi) continue;
                const T* p = tileT + (size_t)((sp + 1) * (BM + 2) + (m + 1)) * ROWE + SEG + f;
                level_fallback_voxel<T>(sh, lt, pt, p, f, m, sp, known + blk * LV_MAXL, gF0, gM0, gS0, do_mom, do_p6, do_w18);
            }
        }
        __syncthreads();

        // ---- flush the per-brick tables ------------------------------------------------------------------------------------
        block_flush_tables<T>(sh, lt, pt, gF0, gM0, gS0, tid);
        __syncthreads();
    }
}

}  // namespace ta

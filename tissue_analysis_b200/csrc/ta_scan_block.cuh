// EXPERIMENTAL scan kernel on the block-bitmask formulation (csrc/ta_block.cuh): both label widths, flag 0x4000.
//
// STATUS: the block arithmetic, the block -> brick -> global transforms, the pair slot conventions and the slab
// ownership rules are validated on the CPU (tests/host/block_host_check.cu, block_volume_check.cu).  This kernel wires
// them to the staging, the per-brick tables and the flush of the product kernel: one 8 x 4 x 2 block per thread; the
// results go to the per-brick shared tables through slot-wise warp merges (MERGE = true) or, in the simpler form kept
// to bisect against, with plain atomics per block (MERGE = false).  Both forms fill exact tables on the CPU emulation
// of the CUDA execution model (tests/host/kernel_emu_check.cpp: fibers for threads, rendezvous for the collectives;
// scalar and emulated-TMA staging).  Superseded as the candidate for the next product kernel by the level formulation
// (ta_scan_level.cuh, which shares the staging / uniform-tile / flush helpers and the pair merge below); kept as the
// simpler form to measure against.  It compiles for sm_100a but was written after the round's GPU budget was spent: it has NOT run on a GPU, and it is NOT
// part of the product build (ta_api.cu includes it only under -DTA_WITH_BLOCK_KERNEL; a product library answers flag
// 0x4000 with TA_ERR_BAD_ARG).  Known before the first run: the per-voxel fallback and the plain-atomics updates are
// out of line (inlined, the kernel took minutes to compile); at 80 registers ptxas spills ~130 bytes in the simple form
// and ~460 bytes in the merged form (four label slots of five 64-bit planes plus 16 merged fields) -- 2 CTAs/SM with 128
// registers, or a three-slot limit, are the first things to try.  Plan and cost model: DESIGN.md section 6.
//
// First run (needs a B200):
//   TA_NVCC_EXTRA=-DTA_WITH_BLOCK_KERNEL TA_OUT=$PWD/build/libtissue_b200_block.so bash tissue_analysis_b200/csrc/build.sh
//   TA_LIB_PATH=$PWD/build/libtissue_b200_block.so TA_PAIR_PATH=block python -m pytest tests/test_gpu_parity.py -x -q
//   TA_LIB_PATH=$PWD/build/libtissue_b200_block.so TA_PAIR_PATH=block python tools/profile_scan.py --config C3
// (TA_PAIR_PATH=block sets flag 0x4000 for every pass, so the whole parity suite runs on this kernel;
//  TA_PAIR_PATH=block_simple selects the form without warp merges.)
#pragma once
#include "ta_block.cuh"

namespace ta {

template <typename T> constexpr size_t scan_block_smem_bytes() {
    return (size_t)TILE_SEGS * 16 + LT_SLOTS * 4 + LT_SLOTS * LT_FIELDS * 4 + PT_SLOTS * sizeof(typename Vox<T>::PKey) +
           PT_SLOTS * PT_WORDS * 4 + 64 +
           NTHREADS * 2;                        // + the list of blocks that take the per-voxel path
}

// (The out-of-line functions below take the shared-memory descriptor BY VALUE: a reference would force the kernel's copy
// into local memory, and every tile access of the kernel would become a generic load instead of LDS.)
// per-voxel fallback for one voxel of a block whose window holds more labels than slots: moments of the voxel and its
// pairs by a first-occurrence scan of the 18 neighbours (the rare path of phase D2 of the product kernel)
template <typename T>
__device__ __noinline__ void block_fallback_voxel(const BrickShared<T> sh, const LabelTable lt, const PairTable pt,
                                                  const T* p, uint32_t f, uint32_t m, uint32_t s, u64 gF0, u64 gM0, u64 gS0,
                                                  bool do_mom, bool do_p6, bool do_w18) {
    constexpr int ROWE = ROWV * Vox<T>::SEG, PLANEE = (BM + 2) * ROWE;
    const uint32_t a = p[0];
    if (do_mom) {
        uint32_t v[LT_FIELDS] = {1u, f, m, s, f * f, f * m, f * s, m * m, m * s, s * s, f, m, s, f, m, s};
        label_add<T>(sh, lt, pt.status, a, v, gF0, gM0, gS0);
    }
    if (!(do_p6 || do_w18)) return;
    constexpr int offs[18] = {1, ROWE, PLANEE, -1, -ROWE, -PLANEE, -ROWE - 1, -ROWE + 1, ROWE - 1, ROWE + 1,
                              -PLANEE - 1, -PLANEE + 1, PLANEE - 1, PLANEE + 1,
                              -PLANEE - ROWE, -PLANEE + ROWE, PLANEE - ROWE, PLANEE + ROWE};
    if (do_p6) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t b = p[offs[k]];
            if (b != a) pair_add<T>(sh, pt, a, b, 2 * k + (a < b ? 0 : 1), 1u);
        }
    }
    if (do_w18) {
#pragma unroll 1
        for (int k = 0; k < 18; ++k) {
            const uint32_t b = p[offs[k]];
            if (b == a) continue;
            bool seen = false;
            for (int q = 0; q < k; ++q) seen |= ((uint32_t)p[offs[q]] == b);
            if (!seen) pair_add<T>(sh, pt, a, b, 6, 1u);
        }
    }
}

// Out-of-line table updates: the block code calls them from up to 4 label slots and 12 ordered slot pairs; inlining
// every copy multiplies the compile time of this kernel by five for nothing.
template <typename T>
__device__ __noinline__ void block_emit_label(const BrickShared<T> sh, const LabelTable lt, uint32_t* status,
                                              uint32_t L, const uint32_t* vin, uint32_t bF, uint32_t bM, uint32_t bS, u64 gF0,
                                              u64 gM0, u64 gS0) {
    uint32_t v[LT_FIELDS];
#pragma unroll
    for (int i = 0; i < LT_FIELDS; ++i) v[i] = vin[i];
    block_shift_moments(v, bF, bM, bS);                           // block -> brick coordinates
    label_add<T>(sh, lt, status, L, v, gF0, gM0, gS0);
}
template <typename T>
__device__ __noinline__ void block_emit_pair(const BrickShared<T> sh, const PairTable pt, uint32_t a, uint32_t b,
                                             uint32_t w18, uint32_t ff, uint32_t fm, uint32_t fsl) {
    // seen from label a at the lower-index voxel: slot 2k when a is the smaller label, else 2k + 1
    const bool lo = a < b;
    uint32_t inc[PT_WORDS];                                       // [w18|f0] [f1|f2] [f3|f4] [f5|-]
    inc[0] = w18 | ((lo ? ff : 0u) << 16);
    inc[1] = (lo ? 0u : ff) | ((lo ? fm : 0u) << 16);
    inc[2] = (lo ? 0u : fm) | ((lo ? fsl : 0u) << 16);
    inc[3] = lo ? 0u : fsl;
    if (inc[0] | inc[1] | inc[2] | inc[3]) pair_add_packed<T>(sh, pt, Vox<T>::key(a, b), inc);
}

// Warp merges (all 32 lanes call; lanes without a contribution pass has = false).  One shared-table update per
// distinct label / pair of the warp: uniform loop over the distinct keys, full-mask redux, the group leaders add.  Same
// pattern as the column flush and phase D of the product kernel.
template <typename T>
__device__ __forceinline__ void block_merge_label(const BrickShared<T>& sh, const LabelTable& lt, uint32_t* status,
                                                  bool has, uint32_t L, const uint32_t v[LT_FIELDS], u64 gF0, u64 gM0,
                                                  u64 gS0, int lane) {
    unsigned pending = __ballot_sync(0xffffffffu, has);
    uint32_t tot[LT_FIELDS];
    bool am_leader = false;
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const uint32_t Lk = __shfl_sync(0xffffffffu, L, leader);
        const bool mine = has && (L == Lk);
        const bool lead = (lane == leader);
#pragma unroll
        for (int f = 0; f < 10; ++f) {
            const uint32_t r = __reduce_add_sync(0xffffffffu, mine ? v[f] : 0u);
            if (lead) tot[f] = r;
        }
#pragma unroll
        for (int f = 10; f < 13; ++f) {
            const uint32_t r = __reduce_min_sync(0xffffffffu, mine ? v[f] : 0xFFFFFFFFu);
            if (lead) tot[f] = r;
        }
#pragma unroll
        for (int f = 13; f < 16; ++f) {
            const uint32_t r = __reduce_max_sync(0xffffffffu, mine ? v[f] : 0u);
            if (lead) tot[f] = r;
        }
        am_leader = am_leader || lead;
        pending &= ~__ballot_sync(0xffffffffu, mine);
    }
    if (am_leader) label_add<T>(sh, lt, status, L, tot, gF0, gM0, gS0);
}
template <typename T>
__device__ __forceinline__ void block_merge_pair(const BrickShared<T>& sh, const PairTable& pt, typename Vox<T>::PKey key,
                                                 const uint32_t inc[PT_WORDS], int lane) {
    unsigned pending = __ballot_sync(0xffffffffu, key != Vox<T>::PEMPTY);
    uint32_t tot[PT_WORDS] = {0u, 0u, 0u, 0u};
    bool am_leader = false;
    while (pending) {
#ifdef TA_STAT
        if (lane == 0) TA_STAT(11, 1);
#endif
        const int leader = __ffs(pending) - 1;
        const typename Vox<T>::PKey kk = __shfl_sync(0xffffffffu, key, leader);
        const bool mine = (key == kk);
#pragma unroll
        for (int w = 0; w < PT_WORDS; ++w) {
            const uint32_t r = __reduce_add_sync(0xffffffffu, mine ? inc[w] : 0u);
            if (lane == leader) tot[w] = r;
        }
        am_leader = am_leader || (lane == leader);
        pending &= ~__ballot_sync(0xffffffffu, mine);
    }
    if (am_leader) pair_add_packed<T>(sh, pt, key, tot);
}

// ---- pieces shared by the block kernels (this file and ta_scan_level.cuh) ----------------------------------------------
// phase A: the tile (brick + one-voxel halo), as in the product kernel: one TMA box copy, or the scalar path.
// SHIFT = 0: the product kernel's layout (brick column f at tile element SEG + f).  SHIFT = 1 (level kernel): the whole
// tile one voxel to the right (column f at element SEG + 1 + f), so that the 10-voxel window row of an 8-wide block
// STARTS on a 16-byte vector: columns -1 .. 6 are one LDS.128, columns 7 and 8 the first word of the next vector -- no
// separate edge-lane loads (they fall on 8 banks) and no edge-lane special case in the compares.  A TMA box may start at
// any element; the scalar path just reads shifted columns.
template <typename T, int SHIFT = 0>
__device__ __forceinline__ void block_stage_tile(const BrickShared<T>& sh, const ScanParams& P, const CUtensorMap& tmap, uint64_t* tma_bar,
                                                 uint32_t& tma_parity, bool use_tma, int F0, int M0, int S0, unsigned iter,
                                                 unsigned int brick, int tid) {
    constexpr int SEG = Vox<T>::SEG, ROWE = ROWV * SEG, BF = NFS * SEG;
    const T* vol = reinterpret_cast<const T*>(P.vol);
    const int nf = (int)P.nf, nm = (int)P.nm, ns = (int)P.ns;
    if (use_tma) {
        if (tid == 0) {
            TA_PTX("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(tma_bar, (uint32_t)(TILE_SEGS * 16));
            tma_load_box_3d(sh.tile, &tmap, tma_bar, F0 - SEG - SHIFT, M0 - 1, S0 - 1);
        }
        __syncwarp();
        unsigned spins = 0;
        while (!mbar_try_wait(tma_bar, tma_parity)) {
            if (++spins > (1u << 18)) {
                if (P.diag && atomicAdd(&P.diag[0], 1ull) == 0ull) {
                    P.diag[1] = ((u64)blockIdx.x << 32) | (u64)tid;
                    P.diag[2] = ((u64)iter << 32) | (u64)brick;
                    __threadfence_system();
                }
                __trap();
            }
        }
        tma_parity ^= 1u;
        const bool edge = (F0 == 0) | (F0 + BF + 1 > nf) | (M0 == 0) | (M0 + BM + 1 > nm) | (S0 < 1) | (S0 + BS + 1 > ns);
        if (edge) {
            T* tw = reinterpret_cast<T*>(sh.tile);
            const int xl = (F0 == 0) ? SEG + SHIFT : 0;            // elements before volume column 0
            const int xr = min(ROWE, nf - F0 + SEG + SHIFT);       // first element beyond the last volume column
            for (int r = tid; r < TILE_ROWS; r += NTHREADS) {
                T* row = tw + r * ROWE;
                if (xl) { const T v = row[xl]; for (int x = 0; x < xl; ++x) row[x] = v; }
                if (xr < ROWE) { const T v = row[xr - 1]; for (int x = xr; x < ROWE; ++x) row[x] = v; }
            }
            __syncthreads();
            for (int i = tid; i < TILE_SEGS; i += NTHREADS) {
                const int r = i / ROWV, m = r % (BM + 2) - 1;
                const int mc = min(max(M0 + m, 0), nm - 1) - M0;
                if (mc != m) sh.tile[i] = sh.tile[i + (mc - m) * ROWV];
            }
            __syncthreads();
            for (int i = tid; i < TILE_SEGS; i += NTHREADS) {
                const int s = i / PLANEV - 1;
                const int sc = min(max(S0 + s, 0), ns - 1) - S0;
                if (sc != s) sh.tile[i] = sh.tile[i + (sc - s) * PLANEV];
            }
        }
    } else {
        for (int i = tid; i < TILE_SEGS; i += NTHREADS) {
            const int fsv = i % ROWV - 1;
            const int r = i / ROWV;
            const int m = r % (BM + 2) - 1, s = r / (BM + 2) - 1;
            const int gs = min(max(S0 + s, 0), ns - 1);
            const int gm = min(max(M0 + m, 0), nm - 1);
            const int gf = F0 + fsv * SEG - SHIFT;
            const T* row = vol + ((size_t)gs * nm + gm) * (size_t)nf;
            T tmp[SEG];
#pragma unroll
            for (int j = 0; j < SEG; ++j) tmp[j] = row[min(max(gf + j, 0), nf - 1)];
            uint4 v;
            if (SEG == 8) {
                v.x = (uint32_t)tmp[0] | ((uint32_t)tmp[1] << 16); v.y = (uint32_t)tmp[2] | ((uint32_t)tmp[3] << 16);
                v.z = (uint32_t)tmp[4 % SEG] | ((uint32_t)tmp[5 % SEG] << 16);
                v.w = (uint32_t)tmp[6 % SEG] | ((uint32_t)tmp[7 % SEG] << 16);
            } else {
                v.x = tmp[0]; v.y = tmp[1]; v.z = tmp[2 % SEG]; v.w = tmp[3 % SEG];
            }
            sh.tile[i] = v;
        }
    }
    __syncthreads();
}

// one-label tile: closed-form moments, no pairs.  true: the brick is done (all threads agree).
template <typename T, int SHIFT = 0>
__device__ __forceinline__ bool block_uniform_tile(const BrickShared<T>& sh, const ScanParams& P, const LabelTable& lt, const PairTable& pt,
                                                   int F0, int M0, int S0, u64 gF0, u64 gM0, u64 gS0, int tid) {
    constexpr int SEG = Vox<T>::SEG, BF = NFS * SEG;
    const T* tileT = reinterpret_cast<const T*>(sh.tile);
    const int nf = (int)P.nf, nm = (int)P.nm;
    const bool do_mom = P.flags & 1u;
    const uint32_t ref_label = tileT[SEG + SHIFT];
    // OR of the xors against the reference label over columns -1 .. BF of every tile row: vectors 1 .. 16 (one LOP3 per
    // word), then the two elements they leave out (SHIFT = 0: one on either side; SHIFT = 1: both on the right), one row
    // per thread
    uint32_t diff = 0u;
    {
        constexpr int ROWE = ROWV * SEG;
        const uint32_t pat = (SEG == 8) ? ref_label * 0x00010001u : ref_label;
        for (int r = tid / NFS; r < TILE_ROWS; r += NTHREADS / NFS) {
            const uint4 v = sh.tile[r * ROWV + 1 + (tid % NFS)];
            diff |= (v.x ^ pat) | (v.y ^ pat) | (v.z ^ pat) | (v.w ^ pat);
        }
        constexpr int E0 = SHIFT ? SEG + BF : SEG - 1, E1 = SHIFT ? SEG + BF + 1 : SEG + BF;
        if (tid < TILE_ROWS) diff |= ((uint32_t)tileT[tid * ROWE + E0] ^ ref_label) | ((uint32_t)tileT[tid * ROWE + E1] ^ ref_label);
    }
    const bool all_ref = (diff == 0u);
    if (__syncthreads_and(all_ref)) {
        if (tid == 0 && do_mom) {
            const uint32_t a = (uint32_t)min(BF, nf - F0), b = (uint32_t)min(BM, nm - M0),
                           c = (uint32_t)min(BS, (int)P.own_hi - S0);
            const uint32_t ta = a * (a - 1) / 2, tb = b * (b - 1) / 2, tc = c * (c - 1) / 2;
            const uint32_t qa = (a - 1) * a * (2 * a - 1) / 6, qb = (b - 1) * b * (2 * b - 1) / 6,
                           qc = (c - 1) * c * (2 * c - 1) / 6;
            uint32_t v[LT_FIELDS];
            v[0] = a * b * c; v[1] = b * c * ta; v[2] = a * c * tb; v[3] = a * b * tc;
            v[4] = b * c * qa; v[5] = c * ta * tb; v[6] = b * ta * tc;
            v[7] = a * c * qb; v[8] = a * tb * tc; v[9] = a * b * qc;
            v[10] = 0; v[11] = 0; v[12] = 0; v[13] = a - 1; v[14] = b - 1; v[15] = c - 1;
            label_to_global(lt, pt.status, ref_label, v, gF0, gM0, gS0);
        }
        return true;
    }
    return false;
}

// flush the per-brick tables to the global ones and clear them (as phase F of the product kernel)
template <typename T>
__device__ __forceinline__ void block_flush_tables(const BrickShared<T>& sh, const LabelTable& lt, const PairTable& pt, u64 gF0, u64 gM0,
                                                   u64 gS0, int tid) {
    typedef typename Vox<T>::PKey PKey;
    for (int i = tid; i < LT_SLOTS; i += NTHREADS) {
        const uint32_t L = sh.lt_key[i];
        if (L == TA_EMPTY32) continue;
        uint32_t* d = &sh.lt_val[i * LT_FIELDS];
        label_to_global(lt, pt.status, L, d, gF0, gM0, gS0);
#pragma unroll
        for (int f = 0; f < LT_FIELDS; ++f) d[f] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
        sh.lt_key[i] = TA_EMPTY32;
    }
    for (int i = tid; i < PT_SLOTS; i += NTHREADS) {
        const PKey key = sh.pt_key[i];
        if (key == Vox<T>::PEMPTY) continue;
        uint32_t* d = &sh.pt_val[i * PT_WORDS];
        const int slot = ta_pair_slot(pt, Vox<T>::key64(key));
#pragma unroll
        for (int idx = 0; idx < 7; ++idx) {
            const uint32_t n = (d[idx >> 1] >> ((idx & 1) * 16)) & 0xFFFFu;
            if (n && slot >= 0) atomicAdd(&pt.vals[(size_t)slot * TA_PAIR_STRIDE + (idx == 0 ? 6 : idx - 1)], n);
        }
#pragma unroll
        for (int w = 0; w < PT_WORDS; ++w) d[w] = 0u;
        sh.pt_key[i] = Vox<T>::PEMPTY;
    }
}

// MERGE = true: slot-wise uniform loops with warp merges (flag 0x4000); false: plain atomics per block (0x4000 | 0x8000),
// kept as the simpler form to bisect against.
template <typename T, bool MERGE>
__global__ void __launch_bounds__(NTHREADS, 3)
scan_block_kernel(ScanParams P, LabelTable lt, PairTable pt, const __grid_constant__ CUtensorMap tmap) {
    typedef typename Vox<T>::PKey PKey;
    constexpr int SEG = Vox<T>::SEG, ROWE = ROWV * SEG, BF = NFS * SEG;
    static_assert(NTHREADS == NFS * (BM / BLK_M) * (BS / BLK_S), "one block per thread");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    BrickShared<T> sh;
    sh.tile = reinterpret_cast<uint4*>(smem_raw);
    sh.lt_key = reinterpret_cast<uint32_t*>(sh.tile + TILE_SEGS);
    sh.lt_val = sh.lt_key + LT_SLOTS;
    sh.pt_val = sh.lt_val + LT_SLOTS * LT_FIELDS;
    sh.pt_key = reinterpret_cast<PKey*>(sh.pt_val + PT_SLOTS * PT_WORDS);
    sh.ctr = reinterpret_cast<unsigned int*>(sh.pt_key + PT_SLOTS);
    unsigned short* crowded = reinterpret_cast<unsigned short*>(sh.ctr + 16);      // [NTHREADS] blocks for the per-voxel path
    const T* tileT = reinterpret_cast<const T*>(sh.tile);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const T* vol = reinterpret_cast<const T*>(P.vol);
    const unsigned int total = (unsigned int)P.nbf * P.nbm * P.nbs;
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    const int nf = (int)P.nf, nm = (int)P.nm, ns = (int)P.ns;

    for (int i = tid; i < LT_SLOTS; i += NTHREADS) sh.lt_key[i] = TA_EMPTY32;
    for (int i = tid; i < LT_SLOTS * LT_FIELDS; i += NTHREADS) {
        const int f = i % LT_FIELDS;
        sh.lt_val[i] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
    }
    for (int i = tid; i < PT_SLOTS; i += NTHREADS) sh.pt_key[i] = Vox<T>::PEMPTY;
    for (int i = tid; i < PT_SLOTS * PT_WORDS; i += NTHREADS) sh.pt_val[i] = 0u;

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
        if (tid == 0) { sh.ctr[6 + ((iter + 1u) & 1u)] = atomicAdd(P.brick_counter, 1u); sh.ctr[1] = 0u; }
        const int bf = brick % P.nbf, bm = (brick / P.nbf) % P.nbm, bs = brick / (P.nbf * P.nbm);
        const int F0 = bf * BF, M0 = bm * BM, S0 = (int)P.own_lo + bs * BS;
        const u64 gF0 = (u64)F0, gM0 = (u64)M0, gS0 = (u64)((long long)S0 + P.slow_offset);

        // ---- phase A: the tile (ends with a block barrier) -------------------------------------------------------------
        block_stage_tile<T>(sh, P, tmap, tma_bar, tma_parity, use_tma, F0, M0, S0, iter, brick, tid);

        // ---- one-label tile: closed-form moments, no pairs ------------------------------------------------------------
        if (block_uniform_tile<T>(sh, P, lt, pt, F0, M0, S0, gF0, gM0, gS0, tid)) continue;

        // ---- one block per thread ------------------------------------------------------------------------------------------
        if (MERGE) {
            const int fs = tid % NFS, m0 = ((tid / NFS) % (BM / BLK_M)) * BLK_M, s0 = (tid / (NFS * (BM / BLK_M))) * BLK_S;
            const int nvf = min(SEG, nf - (F0 + fs * SEG)), nvm = min(BLK_M, nm - (M0 + m0)),
                      nvs = min(BLK_S, (int)P.own_hi - (S0 + s0));
            const uint32_t bF = (uint32_t)(fs * SEG), bM = (uint32_t)m0, bS = (uint32_t)s0;
            BlockSlots<T, BLK_MAXLAB> b;
            bool ok = (nvf > 0 && nvm > 0 && nvs > 0);
            if (ok && !b.discover(sh.tile, fs, m0, s0, nvf, nvm, nvs)) {
                crowded[atomicAdd(&sh.ctr[1], 1u)] = (unsigned short)tid;        // more labels than slots: per-voxel path
                ok = false;
            }
            if (!ok) b.clear();                                                    // every slot answers "empty"
            // the same trip counts in every lane: the merges are full-mask
#pragma unroll
            for (int i = 0; i < BLK_MAXLAB; ++i) {
                uint32_t v[LT_FIELDS];
                const bool has = do_mom && b.label_moments(i, v);
                if (has) block_shift_moments(v, bF, bM, bS);                       // block -> brick coordinates
                block_merge_label(sh, lt, pt.status, has, b.lab[i], v, gF0, gM0, gS0, lane);
            }
            if (do_p6 || do_w18) {
#pragma unroll
                for (int i = 0; i < BLK_MAXLAB; ++i) {
#pragma unroll
                    for (int j = 0; j < BLK_MAXLAB; ++j) {
                        if (i == j) continue;
                        uint32_t w18, ff, fm, fsl;
                        const bool has = b.pair_counts(i, j, w18, ff, fm, fsl);
                        if (!do_w18) w18 = 0u;
                        if (!do_p6) ff = fm = fsl = 0u;
                        const uint32_t a = b.lab[i], c = b.lab[j];
                        const bool lo = a < c;                 // seen from a at the lower-index voxel: slot 2k, else 2k + 1
                        uint32_t inc[PT_WORDS];                // [w18|f0] [f1|f2] [f3|f4] [f5|-]
                        inc[0] = w18 | ((lo ? ff : 0u) << 16);
                        inc[1] = (lo ? 0u : ff) | ((lo ? fm : 0u) << 16);
                        inc[2] = (lo ? 0u : fm) | ((lo ? fsl : 0u) << 16);
                        inc[3] = lo ? 0u : fsl;
                        const bool any = has && (inc[0] | inc[1] | inc[2] | inc[3]);
                        block_merge_pair(sh, pt, any ? Vox<T>::key(a, c) : Vox<T>::PEMPTY, inc, lane);
                    }
                }
            }
            __syncthreads();
            // crowded blocks: all threads share their voxels (64 per block)
            const int ncrowded = (int)sh.ctr[1];
            constexpr int BV = SEG * BLK_M * BLK_S;
            for (int q = tid; q < ncrowded * BV; q += NTHREADS) {
                const int blk = crowded[q / BV], w = q % BV;
                const int cfs = blk % NFS, cm0 = ((blk / NFS) % (BM / BLK_M)) * BLK_M, cs0 = (blk / (NFS * (BM / BLK_M))) * BLK_S;
                const int df = w % SEG, dm = (w / SEG) % BLK_M, ds = w / (SEG * BLK_M);
                const uint32_t f = (uint32_t)(cfs * SEG + df), m = (uint32_t)(cm0 + dm), sp = (uint32_t)(cs0 + ds);
                if (F0 + (int)f >= nf || M0 + (int)m >= nm || S0 + (int)sp >= (int)P.own_hi) continue;
                const T* p = tileT + (size_t)((sp + 1) * (BM + 2) + (m + 1)) * ROWE + SEG + f;
                block_fallback_voxel(sh, lt, pt, p, f, m, sp, gF0, gM0, gS0, do_mom, do_p6, do_w18);
            }
        } else {
            const int fs = tid % NFS, m0 = ((tid / NFS) % (BM / BLK_M)) * BLK_M, s0 = (tid / (NFS * (BM / BLK_M))) * BLK_S;
            const int nvf = min(SEG, nf - (F0 + fs * SEG)), nvm = min(BLK_M, nm - (M0 + m0)),
                      nvs = min(BLK_S, (int)P.own_hi - (S0 + s0));
            if (nvf > 0 && nvm > 0 && nvs > 0) {
                const uint32_t bF = (uint32_t)(fs * SEG), bM = (uint32_t)m0, bS = (uint32_t)s0;
                const bool ok = block_features_reg<T, BLK_MAXLAB>(
                    sh.tile, fs, m0, s0, nvf, nvm, nvs,
                    [&](uint32_t L, const uint32_t vin[16]) {
                        if (do_mom) block_emit_label(sh, lt, pt.status, L, vin, bF, bM, bS, gF0, gM0, gS0);
                    },
                    [&](uint32_t a, uint32_t b, uint32_t w18, uint32_t ff, uint32_t fm, uint32_t fsl) {
                        block_emit_pair(sh, pt, a, b, do_w18 ? w18 : 0u, do_p6 ? ff : 0u, do_p6 ? fm : 0u, do_p6 ? fsl : 0u);
                    });
                if (!ok) {
                    // more labels than slots in the window (0.9 % of the blocks of a tissue): voxel by voxel
                    for (int ds = 0; ds < nvs; ++ds)
                        for (int dm = 0; dm < nvm; ++dm)
                            for (int df = 0; df < nvf; ++df) {
                                const uint32_t f = bF + df, m = bM + dm, s = bS + ds;
                                const T* p = tileT + (size_t)((s + 1) * (BM + 2) + (m + 1)) * ROWE + SEG + f;
                                block_fallback_voxel(sh, lt, pt, p, f, m, s, gF0, gM0, gS0, do_mom, do_p6, do_w18);
                            }
                }
            }
        }
        __syncthreads();

        // ---- flush the per-brick tables (as phase F of the product kernel) ---------------------------------------------
        block_flush_tables<T>(sh, lt, pt, gF0, gM0, gS0, tid);
        __syncthreads();
    }
}

}  // namespace ta

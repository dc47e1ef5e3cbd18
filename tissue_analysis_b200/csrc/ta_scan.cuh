// The single streaming pass over the label volume (sm_100a).
//
// Work unit: a brick of BF x BM x BS voxels (BF = 16 segments of 16 bytes) staged in shared memory with a
// one-voxel halo (clamped at the buffer edges, which reproduces the reference's "image border contributes
// nothing" rule: a clamped neighbour equals an in-bounds 6/18-neighbour or the voxel itself).
// Thread (fseg, m) owns one 16-byte segment column and marches along s.
//
// Hierarchy of paths, cheapest first:
//   1. interior segment (the 3x3 rows x (SEG+2) voxels around it hold one label): run-length accumulation of
//      (count, sum s, sum s^2) in registers, closed-form f/m moments on flush; no pair work at all.
//   2. boundary segment: per-lane moments by runs, per-lane 6-face compares and 18-neighbour distinct-label
//      dedup into a per-brick shared-memory pair hash.
// Per-brick shared tables (labels: u32 sums in brick-local coordinates; pairs: 7 u32 counters) are flushed to
// the global dense label table (u64 REDs) and the global open-addressing pair table once per brick.
#pragma once
#include "ta_common.cuh"

namespace ta {

constexpr int NFS = 16;                 // 16-byte segments per brick row
constexpr int BM = 16;                  // brick rows (mid axis)
constexpr int BS = 8;                   // brick planes (slow axis)
constexpr int NTHREADS = NFS * BM;      // one thread per segment column
constexpr int LT_SLOTS = 64;            // per-brick label slots
constexpr int LT_FIELDS = 16;           // n, sf, sm, ss, sff, sfm, sfs, smm, sms, sss, min f/m/s, max f/m/s
constexpr int PT_SLOTS = 256;           // per-brick pair slots
constexpr int TILE_ROWS = (BS + 2) * (BM + 2);
constexpr int TILE_SEGS = TILE_ROWS * (NFS + 2);

template <typename T> struct Vox;
template <> struct Vox<uint16_t> { static constexpr int SEG = 8; };
template <> struct Vox<uint32_t> { static constexpr int SEG = 4; };

constexpr size_t scan_smem_bytes() {
    return (size_t)TILE_SEGS * 16 + (size_t)TILE_ROWS * NFS * 4 + LT_SLOTS * 4 + LT_SLOTS * LT_FIELDS * 4 +
           PT_SLOTS * 8 + PT_SLOTS * TA_PAIR_STRIDE * 4 + 16;
}

__device__ __forceinline__ uint4 ld_stream_128(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

struct BrickShared {
    uint4* tile;          // [TILE_SEGS]
    uint32_t* codes;      // [TILE_ROWS * NFS]
    uint32_t* lt_key;     // [LT_SLOTS]
    uint32_t* lt_val;     // [LT_SLOTS * LT_FIELDS]
    u64* pt_key;          // [PT_SLOTS]
    uint32_t* pt_val;     // [PT_SLOTS * TA_PAIR_STRIDE]
    unsigned int* next;   // [1] brick index broadcast
};

// ---- global flush of one label's brick-local sums -----------------------------------------------------------
__device__ __forceinline__ void label_to_global(const LabelTable& lt, uint32_t* status, uint32_t L,
                                                const uint32_t* v, u64 F0, u64 M0, u64 S0) {
    if (L >= lt.nrows) { atomicExch(&status[1], 1u); return; }
    u64 n = v[0], sf = v[1], sm = v[2], ss = v[3];
    atomicAdd(&lt.count[L], n);
    atomicAdd(&lt.s1[(size_t)L * 3 + 0], n * F0 + sf);
    atomicAdd(&lt.s1[(size_t)L * 3 + 1], n * M0 + sm);
    atomicAdd(&lt.s1[(size_t)L * 3 + 2], n * S0 + ss);
    u64* q = &lt.s2[(size_t)L * 6];
    atomicAdd(&q[0], n * F0 * F0 + 2 * F0 * sf + v[4]);
    atomicAdd(&q[1], n * F0 * M0 + F0 * sm + M0 * sf + v[5]);
    atomicAdd(&q[2], n * F0 * S0 + F0 * ss + S0 * sf + v[6]);
    atomicAdd(&q[3], n * M0 * M0 + 2 * M0 * sm + v[7]);
    atomicAdd(&q[4], n * M0 * S0 + M0 * ss + S0 * sm + v[8]);
    atomicAdd(&q[5], n * S0 * S0 + 2 * S0 * ss + v[9]);
    atomicMin(&lt.bmin[(size_t)L * 3 + 0], (int)(F0 + v[10]));
    atomicMin(&lt.bmin[(size_t)L * 3 + 1], (int)(M0 + v[11]));
    atomicMin(&lt.bmin[(size_t)L * 3 + 2], (int)(S0 + v[12]));
    atomicMax(&lt.bmax[(size_t)L * 3 + 0], (int)(F0 + v[13]));
    atomicMax(&lt.bmax[(size_t)L * 3 + 1], (int)(M0 + v[14]));
    atomicMax(&lt.bmax[(size_t)L * 3 + 2], (int)(S0 + v[15]));
}

// ---- per-brick label accumulation (brick-local coordinates, u32) -----------------------------------------
__device__ __forceinline__ void label_add(const BrickShared& sh, const LabelTable& lt, uint32_t* status,
                                          uint32_t L, const uint32_t* v, u64 F0, u64 M0, u64 S0) {
    uint32_t slot = (L * 0x9E3779B1u) >> 26;   // 6 bits
    int found = -1;
    for (int probe = 0; probe < LT_SLOTS; ++probe) {
        uint32_t k = *((volatile uint32_t*)&sh.lt_key[slot]);
        if (k == L) { found = (int)slot; break; }
        if (k == TA_EMPTY32) {
            uint32_t old = atomicCAS(&sh.lt_key[slot], TA_EMPTY32, L);
            if (old == TA_EMPTY32 || old == L) { found = (int)slot; break; }
        }
        slot = (slot + 1) & (LT_SLOTS - 1);
    }
    if (found < 0) { label_to_global(lt, status, L, v, F0, M0, S0); return; }
    uint32_t* d = &sh.lt_val[found * LT_FIELDS];
#pragma unroll
    for (int i = 0; i < 10; ++i) if (v[i]) atomicAdd(&d[i], v[i]);
#pragma unroll
    for (int i = 10; i < 13; ++i) atomicMin(&d[i], v[i]);
#pragma unroll
    for (int i = 13; i < 16; ++i) atomicMax(&d[i], v[i]);
}

// ---- per-brick pair accumulation -----------------------------------------------------------------------------
__device__ __forceinline__ void pair_add(const BrickShared& sh, const PairTable& pt, uint32_t a, uint32_t b,
                                         int field, uint32_t n) {
    u64 key = ta_pair_key(a, b);
    uint32_t slot = ta_hash64(key) & (PT_SLOTS - 1);
    for (int probe = 0; probe < PT_SLOTS; ++probe) {
        u64 k = *((volatile u64*)&sh.pt_key[slot]);
        if (k == key) { atomicAdd(&sh.pt_val[slot * TA_PAIR_STRIDE + field], n); return; }
        if (k == TA_EMPTY64) {
            u64 old = atomicCAS(&sh.pt_key[slot], TA_EMPTY64, key);
            if (old == TA_EMPTY64 || old == key) {
                atomicAdd(&sh.pt_val[slot * TA_PAIR_STRIDE + field], n);
                return;
            }
        }
        slot = (slot + 1) & (PT_SLOTS - 1);
    }
    ta_pair_add(pt, key, field, n);   // brick table full: straight to the global table
}

__device__ __forceinline__ uint32_t sumsq_upto(uint32_t k) { return k * (k + 1) * (2 * k + 1) / 6; }  // 0..k

template <typename T>
__global__ void __launch_bounds__(NTHREADS, 2)
scan_kernel(ScanParams P, LabelTable lt, PairTable pt) {
    constexpr int SEG = Vox<T>::SEG;
    constexpr int ROWE = (NFS + 2) * SEG;          // elements per tile row
    constexpr int PLANEE = (BM + 2) * ROWE;        // elements per tile plane
    constexpr int BF = NFS * SEG;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    BrickShared sh;
    sh.tile = reinterpret_cast<uint4*>(smem_raw);
    sh.codes = reinterpret_cast<uint32_t*>(sh.tile + TILE_SEGS);
    sh.lt_key = sh.codes + TILE_ROWS * NFS;
    sh.lt_val = sh.lt_key + LT_SLOTS;
    sh.pt_key = reinterpret_cast<u64*>(sh.lt_val + LT_SLOTS * LT_FIELDS);
    sh.pt_val = reinterpret_cast<uint32_t*>(sh.pt_key + PT_SLOTS);
    sh.next = reinterpret_cast<unsigned int*>(sh.pt_val + PT_SLOTS * TA_PAIR_STRIDE);
    const T* tileT = reinterpret_cast<const T*>(sh.tile);

    const int tid = threadIdx.x;
    const T* vol = reinterpret_cast<const T*>(P.vol);
    const unsigned int total = (unsigned int)P.nbf * P.nbm * P.nbs;
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;

    // reset the per-brick tables once; the flush at the end of each brick re-arms them
    for (int i = tid; i < LT_SLOTS; i += NTHREADS) sh.lt_key[i] = TA_EMPTY32;
    for (int i = tid; i < LT_SLOTS * LT_FIELDS; i += NTHREADS) {
        int f = i % LT_FIELDS;
        sh.lt_val[i] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
    }
    for (int i = tid; i < PT_SLOTS; i += NTHREADS) sh.pt_key[i] = TA_EMPTY64;
    for (int i = tid; i < PT_SLOTS * TA_PAIR_STRIDE; i += NTHREADS) sh.pt_val[i] = 0u;

    for (;;) {
        if (tid == 0) *sh.next = atomicAdd(P.brick_counter, 1u);
        __syncthreads();
        const unsigned int brick = *sh.next;
        if (brick >= total) break;
        const int bf = brick % P.nbf, bm = (brick / P.nbf) % P.nbm, bs = brick / (P.nbf * P.nbm);
        const long long F0 = (long long)bf * BF, M0 = (long long)bm * BM, S0 = P.own_lo + (long long)bs * BS;

        // ---- phase A: stage brick + halo (clamped) ------------------------------------------------------------
        for (int i = tid; i < TILE_SEGS; i += NTHREADS) {
            int fs = i % (NFS + 2) - 1;
            int r = i / (NFS + 2);
            int m = r % (BM + 2) - 1, s = r / (BM + 2) - 1;
            long long gs = min(max(S0 + s, 0LL), P.ns - 1);
            long long gm = min(max(M0 + m, 0LL), P.nm - 1);
            long long gf = F0 + (long long)fs * SEG;
            const T* row = vol + (gs * P.nm + gm) * P.nf;
            uint4 v;
            if (P.vec_ok && gf >= 0 && gf + SEG <= P.nf) {
                v = ld_stream_128(row + gf);
            } else {
                T tmp[SEG];
#pragma unroll
                for (int j = 0; j < SEG; ++j) tmp[j] = row[min(max(gf + j, 0LL), P.nf - 1)];
                if (SEG == 8) {
                    v.x = (uint32_t)tmp[0] | ((uint32_t)tmp[1] << 16);
                    v.y = (uint32_t)tmp[2] | ((uint32_t)tmp[3] << 16);
                    v.z = (uint32_t)tmp[4 % SEG] | ((uint32_t)tmp[5 % SEG] << 16);
                    v.w = (uint32_t)tmp[6 % SEG] | ((uint32_t)tmp[7 % SEG] << 16);
                } else {
                    v.x = tmp[0]; v.y = tmp[1]; v.z = tmp[2 % SEG]; v.w = tmp[3 % SEG];
                }
            }
            sh.tile[i] = v;
        }
        __syncthreads();

        // ---- phase B: per row-segment uniformity code (label if the SEG+2 voxels are equal) -------------------
        for (int i = tid; i < TILE_ROWS * NFS; i += NTHREADS) {
            int fs = i % NFS, r = i / NFS;
            const T* rp = tileT + r * ROWE + (fs + 1) * SEG;
            uint4 v = sh.tile[r * (NFS + 2) + fs + 1];
            uint32_t l = rp[0];
            uint32_t pat = (SEG == 8) ? (l | (l << 16)) : l;
            bool uni = (v.x == pat) & (v.y == pat) & (v.z == pat) & (v.w == pat) &
                       ((uint32_t)rp[-1] == l) & ((uint32_t)rp[SEG] == l);
            sh.codes[i] = uni ? l : TA_EMPTY32;
        }
        __syncthreads();

        // ---- phase C: march ---------------------------------------------------------------------------------------
        {
            const int fs = tid % NFS, m = tid / NFS;
            const long long gf0 = F0 + (long long)fs * SEG, gm = M0 + m;
            const bool col_valid = (gf0 < P.nf) && (gm < P.nm);
            const int nvalid = col_valid ? (int)min((long long)SEG, P.nf - gf0) : 0;
            const int smax = (int)min((long long)BS, P.own_hi - S0);
            const uint32_t lf0 = fs * SEG;
            const u64 gF0 = (u64)F0, gM0 = (u64)M0, gS0 = (u64)(S0 + P.slow_offset);

            uint32_t run_label = TA_EMPTY32, run_cnt = 0, run_s = 0, run_ss = 0, run_first = 0;
            auto flush_run = [&](int s_end) {
                if (run_cnt == 0) return;
                uint32_t v[LT_FIELDS];
                const uint32_t n = run_cnt * SEG;
                const uint32_t rowsum = SEG * lf0 + SEG * (SEG - 1) / 2;
                const uint32_t rowsq = SEG * lf0 * lf0 + lf0 * SEG * (SEG - 1) + (SEG - 1) * SEG * (2 * SEG - 1) / 6;
                v[0] = n; v[1] = run_cnt * rowsum; v[2] = n * m; v[3] = SEG * run_s;
                v[4] = run_cnt * rowsq; v[5] = m * v[1]; v[6] = rowsum * run_s;
                v[7] = n * m * m; v[8] = SEG * m * run_s; v[9] = SEG * run_ss;
                v[10] = lf0; v[11] = m; v[12] = run_first;
                v[13] = lf0 + SEG - 1; v[14] = m; v[15] = s_end - 1;
                label_add(sh, lt, pt.status, run_label, v, gF0, gM0, gS0);
                run_cnt = 0; run_s = 0; run_ss = 0;
            };

            auto tcode = [&](int s) {
                int base = ((s + 1) * (BM + 2) + (m + 1)) * NFS + fs;
                uint32_t e0 = sh.codes[base - NFS], e1 = sh.codes[base], e2 = sh.codes[base + NFS];
                return (e0 == e1 && e1 == e2) ? e1 : TA_EMPTY32;
            };

            if (col_valid && smax > 0) {
                uint32_t t_prev = tcode(-1), t_cur = tcode(0);
                for (int s = 0; s < smax; ++s) {
                    const uint32_t t_next = tcode(s + 1);
                    const uint32_t e_c = sh.codes[((s + 1) * (BM + 2) + (m + 1)) * NFS + fs];
                    const bool interior = (t_cur != TA_EMPTY32) && (t_prev == t_cur) && (t_next == t_cur);
                    const T* cp = tileT + (s + 1) * PLANEE + (m + 1) * ROWE + (fs + 1) * SEG;

                    if (do_mom) {
                        if (e_c != TA_EMPTY32 && nvalid == SEG) {
                            if (e_c != run_label) { flush_run(s); run_label = e_c; run_first = s; }
                            run_cnt += 1; run_s += s; run_ss += s * s;
                        } else {
                            flush_run(s);
                            run_label = TA_EMPTY32;
                            int j0 = 0;
                            while (j0 < nvalid) {
                                uint32_t L = cp[j0];
                                int j1 = j0 + 1;
                                while (j1 < nvalid && (uint32_t)cp[j1] == L) ++j1;
                                uint32_t len = j1 - j0;
                                uint32_t sj = (uint32_t)(j0 + j1 - 1) * len / 2;
                                uint32_t sjj = sumsq_upto(j1 - 1) - (j0 > 0 ? sumsq_upto(j0 - 1) : 0u);
                                uint32_t v[LT_FIELDS];
                                v[0] = len; v[1] = len * lf0 + sj; v[2] = len * m; v[3] = len * s;
                                v[4] = len * lf0 * lf0 + 2 * lf0 * sj + sjj; v[5] = m * v[1]; v[6] = s * v[1];
                                v[7] = len * m * m; v[8] = len * m * s; v[9] = len * s * s;
                                v[10] = lf0 + j0; v[11] = m; v[12] = s;
                                v[13] = lf0 + j1 - 1; v[14] = m; v[15] = s;
                                label_add(sh, lt, pt.status, L, v, gF0, gM0, gS0);
                                j0 = j1;
                            }
                        }
                    }

                    if (!interior && (do_p6 || do_w18)) {
                        for (int j = 0; j < nvalid; ++j) {
                            const T* p = cp + j;
                            const uint32_t a = p[0];
                            if (do_p6) {
                                uint32_t w = p[1];
                                if (w != a) pair_add(sh, pt, a, w, a < w ? 0 : 1, 1u);
                                w = p[ROWE];
                                if (w != a) pair_add(sh, pt, a, w, a < w ? 2 : 3, 1u);
                                w = p[PLANEE];
                                if (w != a) pair_add(sh, pt, a, w, a < w ? 4 : 5, 1u);
                            }
                            if (do_w18) {
                                const int offs[18] = {
                                    -1, 1, -ROWE, ROWE, -PLANEE, PLANEE,
                                    -ROWE - 1, -ROWE + 1, ROWE - 1, ROWE + 1,
                                    -PLANEE - 1, -PLANEE + 1, PLANEE - 1, PLANEE + 1,
                                    -PLANEE - ROWE, -PLANEE + ROWE, PLANEE - ROWE, PLANEE + ROWE};
                                uint32_t d0 = a, d1 = a, d2 = a, d3 = a;
                                int nd = 0;
#pragma unroll
                                for (int k = 0; k < 18; ++k) {
                                    uint32_t b = p[offs[k]];
                                    if (b != a && b != d0 && b != d1 && b != d2 && b != d3) {
                                        if (nd == 0) d0 = b; else if (nd == 1) d1 = b;
                                        else if (nd == 2) d2 = b; else if (nd == 3) d3 = b;
                                        ++nd;
                                    }
                                }
                                if (nd <= 4) {
                                    if (nd > 0) pair_add(sh, pt, a, d0, 6, 1u);
                                    if (nd > 1) pair_add(sh, pt, a, d1, 6, 1u);
                                    if (nd > 2) pair_add(sh, pt, a, d2, 6, 1u);
                                    if (nd > 3) pair_add(sh, pt, a, d3, 6, 1u);
                                } else {
                                    // more than four distinct neighbour labels: exact first-occurrence rescan
                                    for (int k = 0; k < 18; ++k) {
                                        uint32_t b = p[offs[k]];
                                        if (b == a) continue;
                                        bool seen = false;
                                        for (int q = 0; q < k; ++q) seen |= ((uint32_t)p[offs[q]] == b);
                                        if (!seen) pair_add(sh, pt, a, b, 6, 1u);
                                    }
                                }
                            }
                        }
                    }
                    t_prev = t_cur; t_cur = t_next;
                }
                if (do_mom) flush_run(smax);
            }
        }
        __syncthreads();

        // ---- flush the per-brick tables -----------------------------------------------------------------------------
        {
            const u64 gF0 = (u64)F0, gM0 = (u64)M0, gS0 = (u64)(S0 + P.slow_offset);
            for (int i = tid; i < LT_SLOTS; i += NTHREADS) {
                uint32_t L = sh.lt_key[i];
                if (L == TA_EMPTY32) continue;
                uint32_t* d = &sh.lt_val[i * LT_FIELDS];
                label_to_global(lt, pt.status, L, d, gF0, gM0, gS0);
#pragma unroll
                for (int f = 0; f < LT_FIELDS; ++f) d[f] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
                sh.lt_key[i] = TA_EMPTY32;
            }
            for (int i = tid; i < PT_SLOTS; i += NTHREADS) {
                u64 key = sh.pt_key[i];
                if (key == TA_EMPTY64) continue;
                uint32_t* d = &sh.pt_val[i * TA_PAIR_STRIDE];
                int slot = ta_pair_slot(pt, key);
#pragma unroll
                for (int f = 0; f < 7; ++f) {
                    if (d[f] && slot >= 0) atomicAdd(&pt.vals[(size_t)slot * TA_PAIR_STRIDE + f], d[f]);
                    d[f] = 0u;
                }
                sh.pt_key[i] = TA_EMPTY64;
            }
        }
        __syncthreads();
    }
}

}  // namespace ta

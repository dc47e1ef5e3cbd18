// The streaming pass, bit-mask formulation (sm_100a).  Same tables as scan_kernel (ta_scan.cuh), different arithmetic:
// every comparison of a voxel with a label is made ONCE, packed two voxels per instruction, and turned into one bit;
// everything after that -- 18-neighbourhood dilation, face tests, moments -- is boolean algebra and popcounts on 32-voxel
// words, warp-uniform, without worklists.
//
// Work unit: a brick of RW x OM x ZB = 32 x 30 x 8 voxels, staged with a one-voxel halo as ONE TMA box of
// (32 + 2 segments of 16 bytes) x 32 rows x 10 planes.  One warp per tile plane, one lane per tile row (lanes 0 and 31 are
// the halo rows, planes 0 and 9 the halo planes).
//
//   P1  plane warp p discovers the labels of its plane tile (34 x 32 voxels): the lane rows are held in registers; the
//       leader of the still uncovered voxels names a label L, every lane compares its row with L (VIADDMNMX.U16x2: packed
//       subtract + min 1, then one IMAD per word gathers the bits) and stores (own, halo) = 32 + 2 bits in shared memory:
//       masks[p][slot][lane].  The labels of a plane sit in lists[p][slot], slot < K.  More than K labels in one plane
//       tile (noise, never tissue): the whole brick takes the per-voxel path G below.
//   --  barrier; the tile is dead now: thread 0 issues the NEXT brick's box copy, it lands under P2.
//   P2  plane warp p (owned planes) for every label b of planes p-1, p, p+1:
//         D_b = dilation of b by the 18-neighbourhood, restricted to row `lane` of plane p -- ORs of the nine row masks
//               around, two shifts for the f direction; Bf / Bm / Bs = b as the +f / +m / +s neighbour;
//         for every label a of plane p that meets D_b somewhere in the warp:
//               wall18 += popc(M_a & D_b), faces += popc(M_a & Bf), popc(M_a & Bm), popc(M_a & Bs)    (a is the lower voxel)
//               two full-mask redux per (a, b); the results wait in one lane each and go to the per-brick pair table in one
//               SIMT pass.
//       Moments: per label of the plane n = popc(M), closed forms for sum f, sum f^2 of a run of bits, 5 redux, bounds from
//       redux.or / ballot, one lane per label updates the per-brick label table.
//   G   (rare) every voxel of the brick against its 18 neighbours, straight from the tile.
//   F   flush the per-brick tables (brick-local u32 sums -> shifted u64 global REDs; pair slots -> global hash).
//   A brick whose tile is one label altogether (background, inside of a big cell) ends after P1 with closed-form moments;
//   consecutive such bricks of one label are merged in shared memory before they touch the global table.
#pragma once
#include "ta_scan.cuh"

namespace ta {
namespace mk {

constexpr int RW = 32;                    // owned voxels per lane row
constexpr int OM = 30;                    // owned rows per brick (lanes 1..30)
constexpr int TM = 32;                    // tile rows = lanes
constexpr int ZB = 8;                     // owned planes per brick
constexpr int TP = ZB + 2;                // tile planes = warps
constexpr int NTHREADS = TP * 32;
constexpr int K = 14;                     // label slots per plane tile

template <typename T> struct Geo {
    static constexpr int HV = 16 / (int)sizeof(T);       // halo elements per side (one 16-byte segment)
    static constexpr int TRE = RW + 2 * HV;              // elements per tile row
    static constexpr int NW = RW * (int)sizeof(T) / 4;   // 32-bit words of the owned part of a row
    static constexpr int ROWB = TRE * (int)sizeof(T);    // bytes per tile row (a multiple of 16)
    static constexpr int ROWV = ROWB / 16;
    static constexpr int TILE_BYTES = TP * TM * ROWB;
};

template <typename T> constexpr size_t smem_bytes() {
    return (size_t)Geo<T>::TILE_BYTES + (size_t)TP * K * 32 * 8 + (size_t)TP * 32 * 4 + LT_SLOTS * 4 + LT_SLOTS * LT_FIELDS * 4 +
           PT_SLOTS * PT_WORDS * 4 + PT_SLOTS * sizeof(typename Vox<T>::PKey) + 32 * 4 + 16 * 8 + 128;
}

// ---- packed compares ---------------------------------------------------------------------------------------------------
#ifdef TA_EMU_TMA
inline uint32_t add_u16x2(uint32_t a, uint32_t b) { return ((a + b) & 0xFFFFu) | (((a >> 16) + (b >> 16)) << 16); }
inline uint32_t min_u16x2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
#else
__device__ __forceinline__ uint32_t add_u16x2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("add.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t min_u16x2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
#endif

// bit j set iff voxel j of the row equals L
template <typename T> struct RowMask;
template <> struct RowMask<uint16_t> {
    static __device__ __forceinline__ uint32_t eq(const uint32_t (&w)[16], uint32_t L) {
        const uint32_t NL = ((0u - L) & 0xFFFFu) * 0x10001u;       // -L in both halves: (v - L) mod 2^16 == 0 <=> v == L
        uint32_t a0 = 0u, a1 = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) a0 += min_u16x2(add_u16x2(w[j], NL), 0x00010001u) << (2 * j);
#pragma unroll
        for (int j = 0; j < 8; ++j) a1 += min_u16x2(add_u16x2(w[8 + j], NL), 0x00010001u) << (2 * j);
        const uint32_t ne = ((a0 & 0x5555u) | ((a0 >> 15) & 0xAAAAu)) | (((a1 & 0x5555u) | ((a1 >> 15) & 0xAAAAu)) << 16);
        return ~ne;
    }
};
template <> struct RowMask<uint32_t> {
    static __device__ __forceinline__ uint32_t eq(const uint32_t (&w)[32], uint32_t L) {
        uint32_t ne = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j) ne += min(w[j] - L, 1u) << j;
        return ~ne;
    }
};

// sum of the bit positions and of their squares
__device__ __forceinline__ void bit_moments(uint32_t M, uint32_t n, uint32_t& sf, uint32_t& sff) {
    const uint32_t lowbit = M & (0u - M);
    if (((M + lowbit) & M) == 0u) {                      // one run of bits [lo, lo + n)
        const uint32_t lo = (uint32_t)__ffs(M) - 1u;     // M != 0 here
        const uint32_t t = n * (n - 1u);                 // 2 * sum_{i<n} i
        sf = n * lo + (t >> 1);
        sff = n * lo * lo + lo * t + (t * (2u * n - 1u)) / 6u;
    } else {
        sf = 0u; sff = 0u;
        while (M) {
            const uint32_t j = (uint32_t)__ffs(M) - 1u;
            M &= M - 1u;
            sf += j; sff += j * j;
        }
    }
}

// brick-local sums of one label -> global u64 sums and bounds (the arithmetic of label_to_global, without the atomics)
__device__ __forceinline__ void local_to_global(const uint32_t* v, u64 F0, u64 M0, u64 S0, u64 g[10], int bmn[3], int bmx[3]) {
    const u64 n = v[0], sf = v[1], sm = v[2], ss = v[3];
    g[0] = n;
    g[1] = n * F0 + sf; g[2] = n * M0 + sm; g[3] = n * S0 + ss;
    g[4] = n * F0 * F0 + 2 * F0 * sf + v[4];
    g[5] = n * F0 * M0 + F0 * sm + M0 * sf + v[5];
    g[6] = n * F0 * S0 + F0 * ss + S0 * sf + v[6];
    g[7] = n * M0 * M0 + 2 * M0 * sm + v[7];
    g[8] = n * M0 * S0 + M0 * ss + S0 * sm + v[8];
    g[9] = n * S0 * S0 + 2 * S0 * ss + v[9];
    bmn[0] = (int)(F0 + v[10]); bmn[1] = (int)(M0 + v[11]); bmn[2] = (int)(S0 + v[12]);
    bmx[0] = (int)(F0 + v[13]); bmx[1] = (int)(M0 + v[14]); bmx[2] = (int)(S0 + v[15]);
}
__device__ __forceinline__ void global_apply(const LabelTable& lt, uint32_t* status, uint32_t L, const u64* g, const int* bmn,
                                             const int* bmx) {
    if (L >= lt.nrows) { atomicExch(&status[1], 1u); return; }
    atomicAdd(&lt.count[L], g[0]);
#pragma unroll
    for (int i = 0; i < 3; ++i) atomicAdd(&lt.s1[(size_t)L * 3 + i], g[1 + i]);
#pragma unroll
    for (int i = 0; i < 6; ++i) atomicAdd(&lt.s2[(size_t)L * 6 + i], g[4 + i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) { atomicMin(&lt.bmin[(size_t)L * 3 + i], bmn[i]); atomicMax(&lt.bmax[(size_t)L * 3 + i], bmx[i]); }
}

// One-label bricks of the same label that follow each other in a CTA are summed here before they reach the global table
// (thousands of background bricks would otherwise hit the same sixteen addresses).  Thread 0 only.
struct Carry {
    u64 g[10];
    int bmn[3], bmx[3];
    uint32_t label, valid;
};

template <typename T>
__global__ void __launch_bounds__(NTHREADS, (sizeof(T) == 2 ? 3 : 2))
mask_kernel(ScanParams P, LabelTable lt, PairTable pt, const __grid_constant__ CUtensorMap tmap) {
    typedef typename Vox<T>::PKey PKey;
    typedef Geo<T> G;
    constexpr int HV = G::HV, TRE = G::TRE, NW = G::NW, ROWV = G::ROWV;
    constexpr int PLANEE = TM * TRE;
    constexpr uint32_t FULL = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    uint4* tilev = reinterpret_cast<uint4*>(smem_raw);
    uint2* masks = reinterpret_cast<uint2*>(smem_raw + G::TILE_BYTES);                  // [TP][K][32]
    uint32_t* lists = reinterpret_cast<uint32_t*>(masks + TP * K * 32);                  // [TP][32]
    BrickShared<T> sh{};
    sh.lt_key = lists + TP * 32;
    sh.lt_val = sh.lt_key + LT_SLOTS;
    sh.pt_val = sh.lt_val + LT_SLOTS * LT_FIELDS;
    sh.pt_key = reinterpret_cast<PKey*>(sh.pt_val + PT_SLOTS * PT_WORDS);
    // ctr: [0..1] brick index ping-pong, [2] overflow flag of the brick, [4..5] mbarrier, [8 + p] label count of plane p
    unsigned int* ctr = reinterpret_cast<unsigned int*>(sh.pt_key + PT_SLOTS);
    Carry* carry = reinterpret_cast<Carry*>(ctr + 32);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int p = tid >> 5;                     // tile plane of this warp
    const T* vol = reinterpret_cast<const T*>(P.vol);
    const unsigned int total = (unsigned int)P.nbf * P.nbm * P.nbs;
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    const bool do_pairs = do_p6 || do_w18;
    const int nf = (int)P.nf, nm = (int)P.nm, ns = (int)P.ns;

    for (int i = tid; i < LT_SLOTS; i += NTHREADS) sh.lt_key[i] = TA_EMPTY32;
    for (int i = tid; i < LT_SLOTS * LT_FIELDS; i += NTHREADS) {
        const int f = i % LT_FIELDS;
        sh.lt_val[i] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
    }
    for (int i = tid; i < PT_SLOTS; i += NTHREADS) sh.pt_key[i] = Vox<T>::PEMPTY;
    for (int i = tid; i < PT_SLOTS * PT_WORDS; i += NTHREADS) sh.pt_val[i] = 0u;

    uint64_t* tma_bar = reinterpret_cast<uint64_t*>(ctr + 4);
    uint32_t tma_parity = 0u;
    const bool use_tma = P.use_tma && ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u) == 0u;
    if (tid == 0) {
        if (use_tma) {
            mbar_init(tma_bar, 1u);
            TA_PTX("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        carry->valid = 0u;
        ctr[2] = 0u;
    }

    auto brick_origin = [&](unsigned int brick, int& F0, int& M0, int& S0) {
        const int bf = brick % P.nbf, bm = (brick / P.nbf) % P.nbm, bs = brick / (P.nbf * P.nbm);
        F0 = bf * RW; M0 = bm * OM; S0 = (int)P.own_lo + bs * ZB;
    };
    auto issue_box = [&](unsigned int brick) {          // thread 0, after a barrier that ended every read of the tile
        int F0, M0, S0;
        brick_origin(brick, F0, M0, S0);
        TA_PTX("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(tma_bar, (uint32_t)G::TILE_BYTES);
        tma_load_box_3d(tile, &tmap, tma_bar, F0 - HV, M0 - 1, S0 - 1);
    };

    __syncthreads();
    if (tid == 0) {
        const unsigned int b0 = atomicAdd(P.brick_counter, 1u);
        ctr[0] = b0;
        if (use_tma && b0 < total) issue_box(b0);
    }
    __syncthreads();

    for (unsigned iter = 0;; ++iter) {
        const unsigned int brick = ctr[iter & 1u];
        if (brick >= total) break;
        if (tid == 0) ctr[(iter + 1u) & 1u] = atomicAdd(P.brick_counter, 1u);
        int F0, M0, S0;
        brick_origin(brick, F0, M0, S0);
        const u64 gF0 = (u64)F0, gM0 = (u64)M0, gS0 = (u64)((long long)S0 + P.slow_offset);
        const int nown = min(ZB, (int)P.own_hi - S0);          // owned planes of this brick: tile planes 1 .. nown
        const int fvalid_n = min(RW, nf - F0);
        const uint32_t fvalid = fvalid_n >= 32 ? FULL : ((1u << fvalid_n) - 1u);

        // ---- stage the tile -------------------------------------------------------------------------------------------
        if (use_tma) {
            __syncwarp();
            unsigned spins = 0;
            u64 t_first = 0ull;
            bool lost = false;
            while (!mbar_try_wait(tma_bar, tma_parity)) {
                if ((++spins & 0xFFFu) == 0u) {
                    const u64 now = ta_globaltimer();
                    if (t_first == 0ull) t_first = now;
                    else if (now - t_first > 5000000000ull) { lost = true; break; }
                }
            }
            if (lost) {
                if (P.diag && atomicAdd(&P.diag[0], 1ull) == 0ull) {
                    P.diag[1] = ((u64)blockIdx.x << 32) | (u64)tid;
                    P.diag[2] = ((u64)iter << 32) | (u64)brick;
                    P.diag[3] = ((u64)tma_parity << 32) | (u64)ctr[(iter + 1u) & 1u];
                    P.diag[4] = *reinterpret_cast<volatile u64*>(tma_bar);
                    P.diag[5] = ((u64)(uint32_t)(F0 - HV) << 32) | ((u64)(uint32_t)(M0 - 1) << 16) | (u64)(uint32_t)(S0 - 1);
                    __threadfence_system();
                }
                return;
            }
            tma_parity ^= 1u;
            // elements outside the buffer arrived as zeros; the tile wants replicated edge voxels
            const bool edge = (F0 == 0) | (F0 + RW + HV > nf) | (M0 == 0) | (M0 + TM - 1 > nm) | (S0 < 1) | (S0 + ZB + 1 > ns);
            if (edge) {
                const int xl = (F0 == 0) ? HV : 0;                        // elements [0, xl) <- element xl
                const int xr = min(TRE, nf - F0 + HV);                    // elements [xr, TRE) <- element xr - 1
                for (int r = tid; r < TP * TM; r += NTHREADS) {
                    T* row = tile + r * TRE;
                    if (xl) { const T v = row[xl]; for (int x = 0; x < xl; ++x) row[x] = v; }
                    if (xr < TRE) { const T v = row[xr - 1]; for (int x = xr; x < TRE; ++x) row[x] = v; }
                }
                __syncthreads();
                for (int i = tid; i < TP * TM * ROWV; i += NTHREADS) {
                    const int r = (i / ROWV) % TM;
                    const int m = M0 - 1 + r;
                    const int mc = min(max(m, 0), nm - 1);
                    if (mc != m) tilev[i] = tilev[i + (mc - m) * ROWV];
                }
                __syncthreads();
                for (int i = tid; i < TP * TM * ROWV; i += NTHREADS) {
                    const int q = i / (TM * ROWV);
                    const int s = S0 - 1 + q;
                    const int sc = min(max(s, 0), ns - 1);
                    if (sc != s) tilev[i] = tilev[i + (sc - s) * (TM * ROWV)];
                }
                __syncthreads();
            }
        } else {
            for (int i = tid; i < TP * TM * TRE; i += NTHREADS) {
                const int e = i % TRE, r = (i / TRE) % TM, q = i / PLANEE;
                const int gf = min(max(F0 - HV + e, 0), nf - 1);
                const int gm = min(max(M0 - 1 + r, 0), nm - 1);
                const int gs = min(max(S0 - 1 + q, 0), ns - 1);
                tile[i] = vol[((size_t)gs * nm + gm) * (size_t)nf + gf];
            }
            __syncthreads();
        }

        // ---- P1: labels and row masks of plane p ----------------------------------------------------------------------
        const uint32_t ref_label = tile[HV];               // plane 0, row 0, first owned column
        bool one_label = true;                             // this plane: not needed, or all ref_label
        if (p <= nown + 1) {
            const T* row = tile + (p * TM + lane) * TRE;
            uint32_t w[NW];
            {
                const uint4* rv = reinterpret_cast<const uint4*>(row + HV);
#pragma unroll
                for (int q = 0; q < NW / 4; ++q) {
                    const uint4 v = rv[q];
                    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                }
            }
            const uint32_t hl = row[HV - 1], hr = row[HV + RW];
            uint32_t cov = 0u, covh = 0u, mylab = TA_EMPTY32;
            int k = 0;
            for (;;) {
                const uint32_t unc = ~cov;
                uint32_t cand = hl;
                if (unc) cand = row[HV + __ffs(unc) - 1];
                else if (covh & 1u) cand = hr;
                const unsigned bal = __ballot_sync(FULL, (unc != 0u) || (covh != 3u));
                if (!bal) break;
                if (k == K) { if (lane == 0) ctr[2] = 1u; break; }
                const uint32_t L = __shfl_sync(FULL, cand, __ffs(bal) - 1);
                const uint32_t M = RowMask<T>::eq(w, L);
                const uint32_t el = (hl == L) ? 1u : 0u, er = (hr == L) ? 1u : 0u;
                cov |= M; covh |= el | (er << 1);
                masks[(p * K + k) * 32 + lane] = make_uint2(M, el | (er << 31));
                if (lane == k) mylab = L;
                ++k;
            }
            lists[p * 32 + lane] = mylab;
            if (lane == 0) ctr[8 + p] = (unsigned)k;
            one_label = (k == 1) && (__shfl_sync(FULL, mylab, 0) == ref_label);
        }
        const bool uniform = __syncthreads_and(one_label) != 0;
        const bool overflow = ctr[2] != 0u;
        const unsigned int next_brick = ctr[(iter + 1u) & 1u];
        bool box_issued = false;
        if (use_tma && !overflow && next_brick < total) {
            if (tid == 0) issue_box(next_brick);
            box_issued = true;
        }

        if (uniform) {
            // the whole tile (brick + halo) is one label: closed-form moments, no pairs
            if (tid == 0 && do_mom) {
                const uint32_t a = (uint32_t)fvalid_n, b = (uint32_t)min(OM, nm - M0), c = (uint32_t)nown;
                const uint32_t ta_ = a * (a - 1) / 2, tb = b * (b - 1) / 2, tc = c * (c - 1) / 2;
                const uint32_t qa = (a - 1) * a * (2 * a - 1) / 6, qb = (b - 1) * b * (2 * b - 1) / 6, qc = (c - 1) * c * (2 * c - 1) / 6;
                uint32_t v[LT_FIELDS];
                v[0] = a * b * c; v[1] = b * c * ta_; v[2] = a * c * tb; v[3] = a * b * tc;
                v[4] = b * c * qa; v[5] = c * ta_ * tb; v[6] = b * ta_ * tc;
                v[7] = a * c * qb; v[8] = a * tb * tc; v[9] = a * b * qc;
                v[10] = 0; v[11] = 0; v[12] = 0; v[13] = a - 1; v[14] = b - 1; v[15] = c - 1;
                u64 g[10];
                int bmn[3], bmx[3];
                local_to_global(v, gF0, gM0, gS0, g, bmn, bmx);
                if (carry->valid && carry->label != ref_label) {
                    global_apply(lt, pt.status, carry->label, carry->g, carry->bmn, carry->bmx);
                    carry->valid = 0u;
                }
                if (!carry->valid) {
                    carry->valid = 1u; carry->label = ref_label;
                    for (int i = 0; i < 10; ++i) carry->g[i] = g[i];
                    for (int i = 0; i < 3; ++i) { carry->bmn[i] = bmn[i]; carry->bmx[i] = bmx[i]; }
                } else {
                    for (int i = 0; i < 10; ++i) carry->g[i] += g[i];
                    for (int i = 0; i < 3; ++i) { carry->bmn[i] = min(carry->bmn[i], bmn[i]); carry->bmx[i] = max(carry->bmx[i], bmx[i]); }
                }
            }
        } else if (!overflow) {
            // ---- P2: owned plane p ---------------------------------------------------------------------------------------
            if (p >= 1 && p <= nown) {
                const bool own_row = (lane >= 1) && (lane <= OM) && (M0 + lane - 1 < nm);
                const uint32_t fm = own_row ? fvalid : 0u;
                const int kc = (int)ctr[8 + p], kd = (int)ctr[8 + p - 1], ku = (int)ctr[8 + p + 1];
                const uint32_t Lc = lists[p * 32 + lane], Ld = lists[(p - 1) * 32 + lane], Lu = lists[(p + 1) * 32 + lane];
                const uint2* mc = masks + (p * K) * 32;
                const uint2* md = masks + ((p - 1) * K) * 32;
                const uint2* mu = masks + ((p + 1) * K) * 32;
                const int lm = max(lane - 1, 0), lp = min(lane + 1, 31);
                const uint32_t sl = (uint32_t)(p - 1);                     // brick-local plane

                if (do_mom) {
                    uint32_t k1 = 0, k2 = 0, k3 = 0, k4 = 0, k5 = 0, kx = 0, ky = 0;
                    const uint32_t ml = (uint32_t)(lane - 1);              // brick-local row (owned lanes only matter)
                    for (int i = 0; i < kc; ++i) {
                        const uint32_t M = mc[i * 32 + lane].x & fm;
                        const uint32_t n = (uint32_t)__popc(M);
                        uint32_t sf = 0u, sff = 0u;
                        if (M) bit_moments(M, n, sf, sff);
                        const uint32_t r1 = __reduce_add_sync(FULL, n | (sf << 10));
                        const uint32_t r2 = __reduce_add_sync(FULL, n * ml);
                        const uint32_t r3 = __reduce_add_sync(FULL, sff);
                        const uint32_t r4 = __reduce_add_sync(FULL, n * ml * ml);
                        const uint32_t r5 = __reduce_add_sync(FULL, sf * ml);
                        const uint32_t rx = __reduce_or_sync(FULL, M);
                        const uint32_t ry = __ballot_sync(FULL, M != 0u);
                        if (lane == i) { k1 = r1; k2 = r2; k3 = r3; k4 = r4; k5 = r5; kx = rx; ky = ry; }
                    }
                    if (lane < kc && kx) {
                        const uint32_t n = k1 & 0x3FFu, sf = k1 >> 10;
                        uint32_t v[LT_FIELDS];
                        v[0] = n; v[1] = sf; v[2] = k2; v[3] = n * sl; v[4] = k3; v[5] = k5; v[6] = sf * sl;
                        v[7] = k4; v[8] = k2 * sl; v[9] = n * sl * sl;
                        v[10] = (uint32_t)__ffs(kx) - 1u; v[11] = (uint32_t)__ffs(ky) - 2u; v[12] = sl;
                        v[13] = 31u - (uint32_t)__clz(kx); v[14] = 30u - (uint32_t)__clz(ky); v[15] = sl;
                        label_add<T>(sh, lt, pt.status, Lc, v, gF0, gM0, gS0);
                    }
                }

                if (do_pairs) {
                    int nres = 0;
                    uint32_t res_i = 0, res_b = 0, res_1 = 0, res_2 = 0;
                    auto flush_results = [&]() {
                        if (lane < nres) {
                            const uint32_t a = lists[p * 32 + res_i];
                            const bool lo = a < res_b;
                            const uint32_t w18 = res_1 & 0xFFFFu, ff = res_1 >> 16, fmm = res_2 & 0xFFFFu, fss = res_2 >> 16;
                            uint32_t inc[PT_WORDS];
                            inc[0] = w18 | (lo ? ff << 16 : 0u);
                            inc[1] = (lo ? 0u : ff) | (lo ? fmm << 16 : 0u);
                            inc[2] = (lo ? 0u : fmm) | (lo ? fss << 16 : 0u);
                            inc[3] = lo ? 0u : fss;
                            pair_add_packed<T>(sh, pt, Vox<T>::key(a, res_b), inc);
                        }
                        nres = 0;
                    };
                    const int nb = kc + kd + ku;
                    for (int jj = 0; jj < nb; ++jj) {
                        const int src = jj < kc ? 0 : (jj < kc + kd ? 1 : 2);
                        const int j = jj - (src == 0 ? 0 : (src == 1 ? kc : kc + kd));
                        const uint32_t b = __shfl_sync(FULL, src == 0 ? Lc : (src == 1 ? Ld : Lu), j);
                        const int ic = __ffs(__ballot_sync(FULL, Lc == b)) - 1;
                        if (src >= 1 && ic >= 0) continue;
                        const int id = __ffs(__ballot_sync(FULL, Ld == b)) - 1;
                        if (src == 2 && id >= 0) continue;
                        const int iu = __ffs(__ballot_sync(FULL, Lu == b)) - 1;
                        uint32_t Mc0 = 0, Mc1 = 0, Mc2 = 0, Hc0 = 0, Hc1 = 0, Hc2 = 0;
                        uint32_t Md0 = 0, Md1 = 0, Md2 = 0, Hd1 = 0, Mu0 = 0, Mu1 = 0, Mu2 = 0, Hu1 = 0;
                        if (ic >= 0) {
                            const uint2* q = mc + ic * 32;
                            const uint2 x0 = q[lm], x1 = q[lane], x2 = q[lp];
                            Mc0 = x0.x; Hc0 = x0.y; Mc1 = x1.x; Hc1 = x1.y; Mc2 = x2.x; Hc2 = x2.y;
                        }
                        if (id >= 0) {
                            const uint2* q = md + id * 32;
                            const uint2 x1 = q[lane];
                            Md0 = q[lm].x; Md1 = x1.x; Hd1 = x1.y; Md2 = q[lp].x;
                        }
                        if (iu >= 0) {
                            const uint2* q = mu + iu * 32;
                            const uint2 x1 = q[lane];
                            Mu0 = q[lm].x; Mu1 = x1.x; Hu1 = x1.y; Mu2 = q[lp].x;
                        }
                        const uint32_t Y = Mc0 | Mc1 | Mc2 | Md0 | Md1 | Md2 | Mu0 | Mu1 | Mu2;
                        const uint32_t Pm = Mc0 | Mc1 | Mc2 | Md1 | Mu1;
                        const uint32_t PH = Hc0 | Hc1 | Hc2 | Hd1 | Hu1;
                        const uint32_t Dn = (Y | (Pm << 1) | (Pm >> 1) | PH) & fm & ~Mc1;    // not-b voxels with a b in their N18
                        if (!__ballot_sync(FULL, Dn != 0u)) continue;
                        const uint32_t Bf = (Mc1 >> 1) | (Hc1 & 0x80000000u);                 // b is the +f neighbour
                        for (int i = 0; i < kc; ++i) {
                            if (i == ic) continue;
                            const uint32_t Ma = mc[i * 32 + lane].x;
                            const uint32_t t = Ma & Dn;
                            if (!__ballot_sync(FULL, t != 0u)) continue;
                            uint32_t c1 = do_w18 ? (uint32_t)__popc(t) : 0u, c2 = 0u;
                            if (do_p6) {
                                const uint32_t Mo = Ma & fm;
                                c1 |= (uint32_t)__popc(Mo & Bf) << 16;
                                c2 = (uint32_t)__popc(Mo & Mc2) | ((uint32_t)__popc(Mo & Mu1) << 16);
                            }
                            const uint32_t r1 = __reduce_add_sync(FULL, c1);
                            const uint32_t r2 = __reduce_add_sync(FULL, c2);
                            if (lane == nres) { res_i = (uint32_t)i; res_b = b; res_1 = r1; res_2 = r2; }
                            if (++nres == 32) flush_results();
                        }
                    }
                    flush_results();
                }
            }
        } else {
            // ---- G: more than K labels in a plane tile: every owned voxel on its own, straight from the tile -------------
            const int nvo = RW * OM * nown;
            for (int i = tid; i < nvo; i += NTHREADS) {
                const int f = i % RW, m = (i / RW) % OM, s = i / (RW * OM);
                if (F0 + f >= nf || M0 + m >= nm) continue;
                const T* q = tile + ((s + 1) * TM + (m + 1)) * TRE + HV + f;
                const uint32_t a = q[0];
                if (do_mom) {
                    const uint32_t uf = f, um = m, us = s;
                    uint32_t v[LT_FIELDS] = {1u, uf, um, us, uf * uf, uf * um, uf * us, um * um, um * us, us * us, uf, um, us, uf, um, us};
                    label_add<T>(sh, lt, pt.status, a, v, gF0, gM0, gS0);
                }
                if (do_p6) {
                    const uint32_t n0 = q[1], n1 = q[TRE], n2 = q[PLANEE];
                    if (n0 != a) pair_add<T>(sh, pt, a, n0, a < n0 ? 0 : 1, 1u);
                    if (n1 != a) pair_add<T>(sh, pt, a, n1, a < n1 ? 2 : 3, 1u);
                    if (n2 != a) pair_add<T>(sh, pt, a, n2, a < n2 ? 4 : 5, 1u);
                }
                if (do_w18) {
#pragma unroll 1
                    for (int k = 0; k < 18; ++k) {
                        const uint32_t b = q[neighbour_offset<TRE, PLANEE>(k)];
                        if (b == a) continue;
                        bool seen = false;
                        for (int kk = 0; kk < k; ++kk) seen |= ((uint32_t)q[neighbour_offset<TRE, PLANEE>(kk)] == b);
                        if (!seen) pair_add<T>(sh, pt, a, b, 6, 1u);
                    }
                }
            }
        }
        __syncthreads();

        // ---- F: flush the per-brick tables ---------------------------------------------------------------------------------
        if (!uniform) {
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
                for (int w2 = 0; w2 < PT_WORDS; ++w2) d[w2] = 0u;
                sh.pt_key[i] = Vox<T>::PEMPTY;
            }
        }
        if (tid == 0) {
            ctr[2] = 0u;
            if (use_tma && !box_issued && next_brick < total) issue_box(next_brick);   // after G: the tile was in use until the barrier
        }
        __syncthreads();
    }
    if (tid == 0 && carry->valid) global_apply(lt, pt.status, carry->label, carry->g, carry->bmn, carry->bmx);
}

}  // namespace mk
}  // namespace ta

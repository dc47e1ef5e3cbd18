// The scan as TWO barrier-free kernels over label-free OCT RECORDS (sm_100a).
//
// Why two kernels (measured, profiles/r02_SUMMARY.md): the single-kernel form of this idea (ta_scan_meta.cuh: TMA tile ->
// records in shared memory -> blocks, phases separated by block barriers) is exact but spends 53 % of its samples waiting
// at barriers -- a few warps run long serial label loops while the rest of the CTA idles -- and runs at 26 % issue
// utilisation.  Here no thread ever waits for another warp:
//
//   rec_build_kernel    pure streaming, one thread per oct (8 voxels + the two f-neighbours = 10 window lanes), coalesced
//                       128-bit loads, neighbour lanes by shuffle.  Writes one record per oct to a global scratch buffer:
//                           a   lo | (0xFFFF - hi) << 16          (uint32 labels: lo and hi in two arrays)
//                           q   notlo [0..9] | m1 [10..19] | m2 [20..29] | code [30..31]
//                               code 0: at most two labels; 1 / 2: three / four labels (mid labels in e, their lane masks
//                               m1, m2); 3: five or more (BAD: the blocks around it take the per-voxel path)
//                       A record does not depend on any label chosen later; the mask of ANY label L over the window is
//                           (L == lo ? ~notlo : 0) | (L == hi ? notlo & ~m1 & ~m2 : 0) | (L == mid1 ? m1 : 0) | (L == mid2 ? m2 : 0)
//   rec_blocks_kernel   one WARP per brick of 128 x 16 x 8 voxels (16 octs x 16 rows x 8 planes; both label widths), bricks
//                       from an atomic counter, nothing but warp-level synchronisation.  Records come through L1 / L2.
//                         U    the 180 x 16 records of the brick's tile all say "one label, the same": closed form, done
//                         P1   8 x 4 x 2 blocks, 32 at a time: min / max over the 24 records of the block's window.  One
//                              label: closed-form moments.  Two labels (every record within {lo, hi} of the window): list 2.
//                              Otherwise list 3.  (Lists are per warp, in shared memory, filled by ballot + popcount.)
//                         P2a  list 2, one block per lane: ONE mask (the other label is its complement), both dilations,
//                              pair counts as popcounts, moments from a 256-entry byte table
//                         P2b  list 3: label after label (the label at the first window position no earlier label covers)
//                              up to MK_MAXL labels by bit algebra; what is left: restricted per-voxel path in the warp
//                       Table updates: merged across the warp (one row per distinct label / pair, redux), the group leaders
//                       add to the global tables.
//
// Algorithmic bytes stay sizeof(label) per voxel; the records cost 1 B / voxel written + read on top, which is
// reported as traffic, not as work.
#pragma once
#include "ta_scan_meta.cuh"

#ifndef TA_SHARED
#define TA_SHARED __shared__
#endif
// the warps of rec_blocks_kernel are never in step with each other, so its code has to stay in the instruction cache:
// the big helpers are real functions unless TA_REC_INLINE is defined
#ifdef TA_REC_INLINE
#define TA_REC_FN __device__ __forceinline__
#else
#define TA_REC_FN __device__ __noinline__
#endif

namespace ta {

struct RecBuf {
    uint32_t* a;      // uint16: lo | (0xFFFF - hi) << 16;  uint32: lo
    uint32_t* b;      // uint32 only: hi
    uint32_t* q;
    uint32_t* e;      // uint16: mid1 | mid2 << 16;  uint32: mid1
    uint32_t* e2;     // uint32 only: mid2
    int noct;         // octs per row: ceil(nf / 8)
    int plane0;       // buffer plane of record plane 0
    int nplanes;      // record planes
};

constexpr int RB_NOCT = 16, RB_BF = 128;         // a brick: 16 octs x BM rows x BS planes
constexpr int RB_NBLK = RB_NOCT * (BM / BLK_M) * (BS / BLK_S);     // 256 blocks of 8 x 4 x 2
constexpr int RB_WARPS = NTHREADS / 32;

// ---- record access ---------------------------------------------------------------------------------------------------------
template <typename T> struct RecIO;
template <> struct RecIO<uint16_t> {
    static __device__ __forceinline__ void store(const RecBuf& R, long long i, uint32_t lo, uint32_t hi) { R.a[i] = lo | ((hi ^ 0xFFFFu) << 16); }
    static __device__ __forceinline__ void store_mid(const RecBuf& R, long long i, uint32_t m1, uint32_t m2) { R.e[i] = m1 | (m2 << 16); }
    static __device__ __forceinline__ void lohi(const RecBuf& R, int i, uint32_t& lo, uint32_t& hi) {
        const uint32_t x = R.a[i];
        lo = x & 0xFFFFu; hi = (x >> 16) ^ 0xFFFFu;
    }
    static __device__ __forceinline__ void mids(const RecBuf& R, int i, uint32_t& m1, uint32_t& m2) {
        const uint32_t x = R.e[i];
        m1 = x & 0xFFFFu; m2 = x >> 16;
    }
    // one word per record, min-accumulated: low half = smallest lo, high half = 0xFFFF - largest hi
    typedef uint32_t Packed;
    static __device__ __forceinline__ Packed load(const RecBuf& R, int i) { return R.a[i]; }
    static __device__ __forceinline__ Packed init() { return 0xFFFFFFFFu; }
    static __device__ __forceinline__ Packed acc(Packed m, Packed x) { return ta_vminu2(m, x); }
    static __device__ __forceinline__ uint32_t lo_of(Packed m) { return m & 0xFFFFu; }
    static __device__ __forceinline__ uint32_t hi_of(Packed m) { return (m >> 16) ^ 0xFFFFu; }
    // nonzero when the record holds a label that is neither wlo nor whi
    static __device__ __forceinline__ uint32_t outside(Packed x, uint32_t wlo, uint32_t whi) {
        return ta_vminu2(x ^ (wlo | ((whi ^ 0xFFFFu) << 16)), x ^ (whi | ((wlo ^ 0xFFFFu) << 16)));
    }
    static __device__ __forceinline__ bool is_uniform(Packed x, uint32_t ref) { return x == (ref | ((ref ^ 0xFFFFu) << 16)); }
};
template <> struct RecIO<uint32_t> {
    static __device__ __forceinline__ void store(const RecBuf& R, long long i, uint32_t lo, uint32_t hi) { R.a[i] = lo; R.b[i] = hi; }
    static __device__ __forceinline__ void store_mid(const RecBuf& R, long long i, uint32_t m1, uint32_t m2) { R.e[i] = m1; R.e2[i] = m2; }
    static __device__ __forceinline__ void lohi(const RecBuf& R, int i, uint32_t& lo, uint32_t& hi) { lo = R.a[i]; hi = R.b[i]; }
    static __device__ __forceinline__ void mids(const RecBuf& R, int i, uint32_t& m1, uint32_t& m2) { m1 = R.e[i]; m2 = R.e2[i]; }
    struct Packed { uint32_t lo, hi; };
    static __device__ __forceinline__ Packed load(const RecBuf& R, int i) { Packed p; p.lo = R.a[i]; p.hi = R.b[i]; return p; }
    static __device__ __forceinline__ Packed init() { Packed p; p.lo = 0xFFFFFFFFu; p.hi = 0u; return p; }
    static __device__ __forceinline__ Packed acc(Packed m, Packed x) { m.lo = x.lo < m.lo ? x.lo : m.lo; m.hi = x.hi > m.hi ? x.hi : m.hi; return m; }
    static __device__ __forceinline__ uint32_t lo_of(Packed m) { return m.lo; }
    static __device__ __forceinline__ uint32_t hi_of(Packed m) { return m.hi; }
    static __device__ __forceinline__ uint32_t outside(Packed x, uint32_t wlo, uint32_t whi) {
        return ((x.lo != wlo && x.lo != whi) || (x.hi != wlo && x.hi != whi)) ? 1u : 0u;
    }
    static __device__ __forceinline__ bool is_uniform(Packed x, uint32_t ref) { return x.lo == ref && x.hi == ref; }
};

// ---- kernel 1: the records -------------------------------------------------------------------------------------------------
// 10-bit mask of the window lanes that EQUAL label L (lane 0 = left neighbour, 1 .. 8 = the oct, 9 = right neighbour)
TA_HD uint32_t rec_eq16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wh, uint32_t L) {
    const uint32_t LL = L * 0x00010001u, one = 0x00010001u;
    return mk_fold16(ta_vminu2(w0 ^ LL, one), ta_vminu2(w1 ^ LL, one), ta_vminu2(w2 ^ LL, one), ta_vminu2(w3 ^ LL, one),
                     ta_vminu2(wh ^ LL, one)) ^ 0x3FFu;
}
TA_HD uint32_t rec_lane16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wh, int x) {
    if (x == 0) return wh & 0xFFFFu;
    if (x == 9) return wh >> 16;
    const int j = x - 1;
    const uint32_t w = (j >> 1) == 0 ? w0 : (j >> 1) == 1 ? w1 : (j >> 1) == 2 ? w2 : w3;
    return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
}

template <typename T>
__global__ void __launch_bounds__(256)
rec_build_kernel(ScanParams P, RecBuf R) {
    constexpr int SEG = Vox<T>::SEG;
    const T* vol = reinterpret_cast<const T*>(P.vol);
    const int nf = (int)P.nf, nm = (int)P.nm, noct = R.noct;
    const int lane = threadIdx.x & 31;
    const bool fast = P.vec_ok && (nf % 8) == 0;
    // one warp per run of 32 octs of one row (no warp straddles two rows): two 32-bit divisions per warp, none per thread
    const unsigned cpr = (unsigned)(noct + 31) / 32u, nrows = (unsigned)R.nplanes * (unsigned)nm;
    const unsigned nchunks = nrows * cpr, nwarps = gridDim.x * (blockDim.x >> 5);
    for (unsigned wi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wi < nchunks; wi += nwarps) {
        const unsigned row = wi / cpr, ch = wi - row * cpr;
        const int o = (int)(ch * 32u) + lane;
        const bool active = o < noct;
        const unsigned pl = row / (unsigned)nm, m = row - pl * (unsigned)nm;
        const long long i = (long long)row * noct + o;
        const T* rp = vol + ((size_t)(R.plane0 + (int)pl) * nm + m) * (size_t)nf;
        MkOct mo;
        uint32_t w[10];                       // uint32 labels: the ten lanes; uint16: w[0..3] the oct, w[4] the neighbour word
        if constexpr (sizeof(T) == 2) {
            uint4 c = make_uint4(0u, 0u, 0u, 0u);
            if (fast) {
                if (active) c = ld_stream_128(rp + 8 * o);
                // neighbour lanes: the octs either side are the lanes either side, except at the ends of the warp / the row
                uint32_t pw = __shfl_sync(0xffffffffu, c.w, (lane + 31) & 31), nx = __shfl_sync(0xffffffffu, c.x, (lane + 1) & 31);
                if (active) {
                    if (o == 0) pw = c.x << 16;                                   // clamped: the voxel itself
                    else if (lane == 0) pw = (uint32_t)rp[8 * o - 1] << 16;
                    if (o == noct - 1) nx = c.w >> 16;
                    else if (lane == 31) nx = (uint32_t)rp[8 * o + 8];
                }
                w[0] = c.x; w[1] = c.y; w[2] = c.z; w[3] = c.w; w[4] = (pw >> 16) | (nx << 16);
            } else {
                uint32_t v[10];
#pragma unroll
                for (int x = 0; x < 10; ++x) v[x] = active ? (uint32_t)rp[max(0, min(8 * o - 1 + x, nf - 1))] : 0u;
                w[0] = v[1] | (v[2] << 16); w[1] = v[3] | (v[4] << 16); w[2] = v[5] | (v[6] << 16); w[3] = v[7] | (v[8] << 16);
                w[4] = v[0] | (v[9] << 16);
            }
            mo = mk_oct16(w[0], w[1], w[2], w[3], w[4]);
        } else {
            if (fast) {
                uint4 c = make_uint4(0u, 0u, 0u, 0u), d = c;
                if (active) { c = ld_stream_128(rp + 8 * o); d = ld_stream_128(rp + 8 * o + 4); }
                w[1] = c.x; w[2] = c.y; w[3] = c.z; w[4] = c.w; w[5] = d.x; w[6] = d.y; w[7] = d.z; w[8] = d.w;
                w[0] = (active && o > 0) ? (uint32_t)rp[8 * o - 1] : c.x;
                w[9] = (active && o < noct - 1) ? (uint32_t)rp[8 * o + 8] : d.w;
            } else {
#pragma unroll
                for (int x = 0; x < 10; ++x) w[x] = active ? (uint32_t)rp[max(0, min(8 * o - 1 + x, nf - 1))] : 0u;
            }
            mo = mk_oct32(w);
        }
        if (!active) continue;
        uint32_t q = mo.notlo;
        uint32_t mid = mo.notlo & mo.nothi;                 // lanes that are neither lo nor hi
        if (mid) {
            uint32_t L1, L2 = 0u, m1, m2 = 0u, code = 1u;
            const int x1 = __ffs(mid) - 1;
            if constexpr (sizeof(T) == 2) { L1 = rec_lane16(w[0], w[1], w[2], w[3], w[4], x1); m1 = rec_eq16(w[0], w[1], w[2], w[3], w[4], L1); }
            else { L1 = 0u; m1 = 0u;
#pragma unroll
                for (int x = 0; x < 10; ++x) if (x == x1) L1 = w[x];
#pragma unroll
                for (int x = 0; x < 10; ++x) m1 |= (w[x] == L1 ? 1u : 0u) << x; }
            mid &= ~m1;
            if (mid) {
                code = 2u;
                const int x2 = __ffs(mid) - 1;
                if constexpr (sizeof(T) == 2) { L2 = rec_lane16(w[0], w[1], w[2], w[3], w[4], x2); m2 = rec_eq16(w[0], w[1], w[2], w[3], w[4], L2); }
                else {
#pragma unroll
                    for (int x = 0; x < 10; ++x) if (x == x2) L2 = w[x];
#pragma unroll
                    for (int x = 0; x < 10; ++x) m2 |= (w[x] == L2 ? 1u : 0u) << x; }
                mid &= ~m2;
                if (mid) code = 3u;
            }
            q |= (m1 << 10) | (m2 << 20) | (code << 30);
            RecIO<T>::store_mid(R, i, L1, L2);
        }
        RecIO<T>::store(R, i, mo.lo, mo.hi);
        R.q[i] = q;
    }
}

TA_REC_FN void rec_label_to_global(const LabelTable lt, uint32_t* status, uint32_t L, const uint32_t* v, u64 F0, u64 M0, u64 S0) {
    label_to_global(lt, status, L, v, F0, M0, S0);
}
// ---- kernel 2 helpers ---------------------------------------------------------------------------------------------------
// Where a block's 24 window records live: index = planeoff[p] + rowoff[r]  (p = 0 .. 3 window planes, r = 0 .. 5 window rows)
struct RecWin { int po[4]; int ro[6]; };

template <typename T>
__device__ __forceinline__ uint32_t rec_mask(const RecBuf& R, int i, uint32_t L) {
    uint32_t lo, hi;
    RecIO<T>::lohi(R, i, lo, hi);
    const uint32_t q = R.q[i], nl = q & 0x3FFu, m1 = (q >> 10) & 0x3FFu, m2 = (q >> 20) & 0x3FFu;
    uint32_t a = (L == lo ? (nl ^ 0x3FFu) : 0u) | (L == hi ? (nl & ~m1 & ~m2) : 0u);
    if (q >> 30) {
        uint32_t l1, l2;
        RecIO<T>::mids(R, i, l1, l2);
        a |= (L == l1 ? m1 : 0u) | (L == l2 ? m2 : 0u);
    }
    return a;
}
struct Planes4 { u64 p[4]; };
template <typename T>
TA_REC_FN Planes4 rec_label_planes(const RecBuf R, const RecWin W, uint32_t L) {
    Planes4 out;
    u64* A = out.p;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        uint32_t h0 = 0u, h1 = 0u;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const uint32_t a = rec_mask<T>(R, W.po[p] + W.ro[r], L);
            if (r < 3) h0 |= a << (10 * r); else h1 |= a << (10 * (r - 3));
        }
        A[p] = (u64)h0 | ((u64)h1 << 30);
    }
    return out;
}
template <typename T>
__device__ __forceinline__ uint32_t rec_label_at(const RecBuf& R, const RecWin& W, int p, int bit) {
    const int r = (bit * 205) >> 11, x = bit - r * 10;
    const int po = p == 0 ? W.po[0] : p == 1 ? W.po[1] : p == 2 ? W.po[2] : W.po[3];
    const int ro = r == 0 ? W.ro[0] : r == 1 ? W.ro[1] : r == 2 ? W.ro[2] : r == 3 ? W.ro[3] : r == 4 ? W.ro[4] : W.ro[5];
    const int i = po + ro;
    uint32_t lo, hi;
    RecIO<T>::lohi(R, i, lo, hi);
    const uint32_t q = R.q[i];
    if (!((q >> x) & 1u)) return lo;
    if (q >> 30) {
        uint32_t l1, l2;
        RecIO<T>::mids(R, i, l1, l2);
        if ((q >> (10 + x)) & 1u) return l1;
        if ((q >> (20 + x)) & 1u) return l2;
    }
    return hi;
}

// moments and box of the voxels c0 (plane 0 of the block) | c1 (plane 1) in block-local coordinates; false: no voxel.
// (The table form of BlockLevel::label_moments, on bare masks.)
struct MomRow { uint32_t w[MK_ROW]; bool has; };
__device__ __forceinline__ bool rec_mask_moments(u64 c0, u64 c1, const uint32_t* tab, uint32_t v[16]) {
    if (!(c0 | c1)) return false;
    uint32_t a0 = 0u, a1 = 0u, a2 = 0u, ap = 0u, apm = 0u, colmask = 0u, rows = 0u;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int r = 0; r < BLK_M; ++r) {
            const uint32_t b = (uint32_t)((p ? c1 : c0) >> (10 * (r + 1) + 1)) & 0xFFu;
            const uint32_t t = tab[b];
            a0 += t; a1 += (uint32_t)r * t; a2 += (uint32_t)(r * r) * t;
            if (p) { ap += t; apm += (uint32_t)r * t; }
            colmask |= b;
            rows |= b ? (1u << (p * BLK_M + r)) : 0u;
        }
    const uint32_t mrows = (rows | (rows >> BLK_M)) & ((1u << BLK_M) - 1u);
    v[0] = a0 & 0x3FFu; v[1] = (a0 >> 10) & 0x7FFu; v[2] = a1 & 0x3FFu; v[3] = ap & 0x3FFu;
    v[4] = a0 >> 21; v[5] = (a1 >> 10) & 0x7FFu; v[6] = (ap >> 10) & 0x7FFu; v[7] = a2 & 0x3FFu;
    v[8] = apm & 0x3FFu; v[9] = v[3];
    v[10] = (uint32_t)__ffs(colmask) - 1u; v[11] = (uint32_t)__ffs(mrows) - 1u;
    v[12] = (rows & ((1u << BLK_M) - 1u)) ? 0u : 1u;
    v[13] = (uint32_t)ta_fls(colmask); v[14] = (uint32_t)ta_fls(mrows); v[15] = (rows >> BLK_M) ? 1u : 0u;
    return true;
}

// the same, moved to brick coordinates and packed for the warp merge
TA_REC_FN MomRow rec_mask_row(u64 c0, u64 c1, const uint32_t* tab, uint32_t bF, uint32_t bM, uint32_t bS) {
    MomRow r;
    uint32_t v[16];
    r.has = rec_mask_moments(c0, c1, tab, v);
    if (r.has) { block_shift_moments(v, bF, bM, bS); mk_pack_row(v, r.w); }
    return r;
}
// per-voxel path, restricted to what the label steps could not emit (global memory, clamped).  The 18 neighbours are
// read once; wall18 counts a neighbour label at its first occurrence.
template <typename T>
__device__ __noinline__ void rec_fallback_voxel(const ScanParams P, const LabelTable lt, const PairTable pt, int f, int m, int s,
                                                uint32_t bf, uint32_t bm, uint32_t bs, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3,
                                                int nk, u64 gF0, u64 gM0, u64 gS0) {
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    const uint32_t a = mk_vox<T>(P, f, m, s);
    const bool a_in = (nk > 0 && k0 == a) || (nk > 1 && k1 == a) || (nk > 2 && k2 == a) || (nk > 3 && k3 == a);
    if (do_mom && !a_in) {
        uint32_t v[16] = {1u, bf, bm, bs, bf * bf, bf * bm, bf * bs, bm * bm, bm * bs, bs * bs, bf, bm, bs, bf, bm, bs};
        rec_label_to_global(lt, pt.status, a, v, gF0, gM0, gS0);
    }
    if (!(do_p6 || do_w18)) return;
    constexpr int df[18] = {1, 0, 0, -1, 0, 0, -1, 1, -1, 1, -1, 1, -1, 1, 0, 0, 0, 0};
    constexpr int dm[18] = {0, 1, 0, 0, -1, 0, -1, -1, 1, 1, 0, 0, 0, 0, -1, 1, -1, 1};
    constexpr int ds[18] = {0, 0, 1, 0, 0, -1, 0, 0, 0, 0, -1, -1, 1, 1, -1, -1, 1, 1};
    uint32_t nb[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) nb[k] = mk_vox<T>(P, f + df[k], m + dm[k], s + ds[k]);
#pragma unroll
    for (int k = 0; k < 18; ++k) {
        const uint32_t b = nb[k];
        bool emit = (b != a);
        if (a_in) emit = emit && !((nk > 0 && k0 == b) || (nk > 1 && k1 == b) || (nk > 2 && k2 == b) || (nk > 3 && k3 == b));
        if (!emit) continue;
        if (do_p6 && k < 3) ta_pair_add(pt, ta_pair_key(a, b), 2 * k + (a < b ? 0 : 1), 1u);
        if (do_w18) {
            bool seen = false;
#pragma unroll
            for (int q = 0; q < k; ++q) seen = seen || (nb[q] == b);
            if (!seen) ta_pair_add(pt, ta_pair_key(a, b), 6, 1u);
        }
    }
}


// ---- table updates: match + redux merge across the warp, then per-WARP tables in shared memory, flushed once per brick ----
// A warp works on one brick at a time, so its tables hold brick-local sums (u32) of at most WL_SLOTS labels / WP_SLOTS
// pairs; what does not fit goes straight to the global tables.  Only group leaders (one lane per distinct key) touch a
// table, and a key has one leader per call: plain read-modify-write, a CAS only to claim a slot.
constexpr int WL_SLOTS = 16, WP_SLOTS = 32;
template <typename T> struct WarpTabs {
    uint32_t* lkey;                        // [WL_SLOTS] label, TA_EMPTY32 when free
    uint32_t* lval;                        // [WL_SLOTS][16] fields of label_to_global
    typename Vox<T>::PKey* pkey;           // [WP_SLOTS]
    uint32_t* pval;                        // [WP_SLOTS][4] packed 16-bit counters [w18|f0] [f1|f2] [f3|f4] [f5|-]
};
template <typename T> struct WarpTabsSize {
    static constexpr size_t value = WL_SLOTS * 4 + WL_SLOTS * 16 * 4 + WP_SLOTS * sizeof(typename Vox<T>::PKey) + WP_SLOTS * 4 * 4;
};
template <typename T>
__device__ __forceinline__ void warp_tabs_clear(const WarpTabs<T>& t, int lane) {
    if (lane < WL_SLOTS) {
        t.lkey[lane] = TA_EMPTY32;
#pragma unroll
        for (int f = 0; f < 16; ++f) t.lval[lane * 16 + f] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
    }
    t.pkey[lane] = Vox<T>::PEMPTY;
#pragma unroll
    for (int w = 0; w < 4; ++w) t.pval[lane * 4 + w] = 0u;
}
template <typename T>
__device__ __noinline__ void warp_tabs_add_label(const WarpTabs<T> t, const LabelTable lt, uint32_t* status, uint32_t L,
                                                 uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5, uint32_t w6,
                                                 uint32_t w7, uint32_t w8, uint32_t w9, u64 gF0, u64 gM0, u64 gS0) {
    const uint32_t w[MK_ROW] = {w0, w1, w2, w3, w4, w5, w6, w7, w8, w9};
    uint32_t u[16];
    mk_unpack_row(w, u);
    uint32_t slot = (L * 0x9E3779B1u) >> 28;
    int found = -1;
    for (int probe = 0; probe < WL_SLOTS; ++probe) {
        const uint32_t k = *((volatile uint32_t*)&t.lkey[slot]);
        if (k == L) { found = (int)slot; break; }
        if (k == TA_EMPTY32) {
            const uint32_t old = atomicCAS(&t.lkey[slot], TA_EMPTY32, L);
            if (old == TA_EMPTY32 || old == L) { found = (int)slot; break; }
        }
        slot = (slot + 1) & (WL_SLOTS - 1);
    }
    if (found < 0) { rec_label_to_global(lt, status, L, u, gF0, gM0, gS0); return; }
    uint32_t* d = t.lval + found * 16;
#pragma unroll
    for (int f = 0; f < 10; ++f) d[f] += u[f];
#pragma unroll
    for (int f = 10; f < 13; ++f) d[f] = min(d[f], u[f]);
#pragma unroll
    for (int f = 13; f < 16; ++f) d[f] = max(d[f], u[f]);
}
template <typename T>
__device__ __noinline__ void warp_tabs_add_pair(const WarpTabs<T> t, const PairTable pt, typename Vox<T>::PKey key, uint32_t i0, uint32_t i1,
                                                uint32_t i2, uint32_t i3) {
    typedef typename Vox<T>::PKey PKey;
    uint32_t slot = Vox<T>::hash(key) & (WP_SLOTS - 1);
    int found = -1;
    for (int probe = 0; probe < WP_SLOTS; ++probe) {
        const PKey k = *((volatile PKey*)&t.pkey[slot]);
        if (k == key) { found = (int)slot; break; }
        if (k == Vox<T>::PEMPTY) {
            const PKey old = atomicCAS(&t.pkey[slot], Vox<T>::PEMPTY, key);
            if (old == Vox<T>::PEMPTY || old == key) { found = (int)slot; break; }
        }
        slot = (slot + 1) & (WP_SLOTS - 1);
    }
    const uint32_t inc[4] = {i0, i1, i2, i3};
    if (found >= 0) {
        uint32_t* d = t.pval + found * 4;
#pragma unroll
        for (int w = 0; w < 4; ++w) d[w] += inc[w];
        return;
    }
    const int g = ta_pair_slot(pt, Vox<T>::key64(key));
    if (g < 0) return;
    uint32_t* v = &pt.vals[(size_t)g * TA_PAIR_STRIDE];
#pragma unroll
    for (int idx = 0; idx < 7; ++idx) {
        const uint32_t n = (inc[idx >> 1] >> ((idx & 1) * 16)) & 0xFFFFu;
        if (n) atomicAdd(&v[idx == 0 ? 6 : idx - 1], n);
    }
}
// brick done: every slot to the global tables, tables empty again (all lanes call)
template <typename T>
__device__ __noinline__ void warp_tabs_flush(const WarpTabs<T> t, const LabelTable lt, const PairTable pt, u64 gF0, u64 gM0, u64 gS0, int lane) {
    __syncwarp();
    if (lane < WL_SLOTS) {
        const uint32_t L = t.lkey[lane];
        if (L != TA_EMPTY32) {
            uint32_t* d = t.lval + lane * 16;
            rec_label_to_global(lt, pt.status, L, d, gF0, gM0, gS0);
            t.lkey[lane] = TA_EMPTY32;
#pragma unroll
            for (int f = 0; f < 16; ++f) d[f] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
        }
    }
    {
        const typename Vox<T>::PKey key = t.pkey[lane];
        if (key != Vox<T>::PEMPTY) {
            uint32_t* d = t.pval + lane * 4;
            const int g = ta_pair_slot(pt, Vox<T>::key64(key));
            if (g >= 0) {
                uint32_t* v = &pt.vals[(size_t)g * TA_PAIR_STRIDE];
#pragma unroll
                for (int idx = 0; idx < 7; ++idx) {
                    const uint32_t n = (d[idx >> 1] >> ((idx & 1) * 16)) & 0xFFFFu;
                    if (n) atomicAdd(&v[idx == 0 ? 6 : idx - 1], n);
                }
            }
            t.pkey[lane] = Vox<T>::PEMPTY;
#pragma unroll
            for (int w = 0; w < 4; ++w) d[w] = 0u;
        }
    }
    __syncwarp();
}
// all 32 lanes call; has = false: nothing to add.  A warp-uniform loop over the distinct keys, full-mask redux (a redux
// over a sub-mask that differs from lane to lane is a loop over the groups in the compiler's own code: measured, 12 % of
// the kernel's instructions), then the group leaders add in one SIMT pass.
template <typename T>
__device__ __forceinline__ void rec_put_label(const WarpTabs<T>& t, const LabelTable& lt, uint32_t* status, bool has, uint32_t L,
                                              const uint32_t w[MK_ROW], u64 gF0, u64 gM0, u64 gS0, int lane) {
    __syncwarp();
    unsigned pending = __ballot_sync(0xffffffffu, has);
    uint32_t tot[MK_ROW];
    bool am_leader = false;
    while (pending) {
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
    if (am_leader)
        warp_tabs_add_label<T>(t, lt, status, L, tot[0], tot[1], tot[2], tot[3], tot[4], tot[5], tot[6], tot[7], tot[8], tot[9], gF0, gM0, gS0);
}
template <typename T>
__device__ __forceinline__ void rec_put_pair(const WarpTabs<T>& t, const PairTable& pt, bool has, uint32_t a, uint32_t b, const uint32_t inc[4],
                                             int lane) {
    typedef typename Vox<T>::PKey PKey;
    __syncwarp();
    const PKey key = has ? Vox<T>::key(a, b) : Vox<T>::PEMPTY;
    unsigned pending = __ballot_sync(0xffffffffu, has);
    uint32_t tot[4] = {0u, 0u, 0u, 0u};
    bool am_leader = false;
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const PKey kk = __shfl_sync(0xffffffffu, key, leader);
        const bool mine = has && (key == kk);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t r = __reduce_add_sync(0xffffffffu, mine ? inc[w] : 0u);
            if (lane == leader) tot[w] = r;
        }
        am_leader = am_leader || (lane == leader);
        pending &= ~__ballot_sync(0xffffffffu, mine);
    }
    if (am_leader) warp_tabs_add_pair<T>(t, pt, key, tot[0], tot[1], tot[2], tot[3]);
}

// list 3: what step I adds for a block: the moments of slot I and its pairs with the older slots
template <typename T, int I>
__device__ __forceinline__ void rec_emit_slot(const WarpTabs<T>& tabs, const ScanParams& P, const LabelTable& lt, const PairTable& pt,
                                              const uint32_t* momtab, const BlockLevel<T, MK_MAXL>& b, bool active, uint32_t bF, uint32_t bM,
                                              uint32_t bS, u64 gF0, u64 gM0, u64 gS0, int lane) {
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    uint32_t inc[I > 0 ? I : 1][4];
    bool hasp[I > 0 ? I : 1];
    MomRow row;
    row.has = false;
    if (active && do_mom) row = rec_mask_row(b.M1[I] & b.cv0, b.M2[I] & b.cv1, momtab, bF, bM, bS);
#pragma unroll
    for (int j = 0; j < I; ++j) hasp[j] = active && (do_p6 || do_w18) && b.pair_increments(I, j, do_p6, do_w18, inc[j]);
    rec_put_label<T>(tabs, lt, pt.status, row.has, b.lab[I], row.w, gF0, gM0, gS0, lane);
    if (do_p6 || do_w18) {
#pragma unroll
        for (int j = 0; j < I; ++j) rec_put_pair<T>(tabs, pt, hasp[j], b.lab[I], b.lab[j], inc[j], lane);
    }
}
template <typename T, int I>
__device__ __forceinline__ void rec_steps(const WarpTabs<T>& tabs, const RecBuf& R, const RecWin& W, const ScanParams& P, const LabelTable& lt, const PairTable& pt,
                                          const uint32_t* momtab, BlockLevel<T, MK_MAXL>& b, bool& more, uint32_t bF, uint32_t bM, uint32_t bS,
                                          u64 gF0, u64 gM0, u64 gS0, int lane) {
    if constexpr (I < MK_MAXL) {
        if (!__ballot_sync(0xffffffffu, more)) return;
        const bool act = more;
        if (lane == 0) TA_STAT(5, 1);
        if (act) TA_STAT(6, 1);
        if (act) {
            constexpr u64 ALL = LvBlk<T>::PLANE_ALL;
            const int p = b.R0 ? 0 : b.R1 ? 1 : b.R2 ? 2 : 3;
            const u64 rp = b.R0 ? b.R0 : b.R1 ? b.R1 : b.R2 ? b.R2 : b.R3;
            const uint32_t L = rec_label_at<T>(R, W, p, ta_ffs64(rp) - 1);
            const Planes4 A4 = rec_label_planes<T>(R, W, L);
            const u64 neq[4] = {~A4.p[0] & ALL, ~A4.p[1] & ALL, ~A4.p[2] & ALL, ~A4.p[3] & ALL};
            b.template set_slot<I>(L, neq);
            more = (b.R0 | b.R1 | b.R2 | b.R3) != 0ull;
        } else {
            b.template clear_slot<I>();
        }
        rec_emit_slot<T, I>(tabs, P, lt, pt, momtab, b, act, bF, bM, bS, gF0, gM0, gS0, lane);
        rec_steps<T, I + 1>(tabs, R, W, P, lt, pt, momtab, b, more, bF, bM, bS, gF0, gM0, gS0, lane);
    }
}


#ifndef TA_FORCE_LIST3
#define TA_FORCE_LIST3 0
#endif
#ifndef TA_REC_MINB
#define TA_REC_MINB 3
#endif

// ---- kernel 2: one warp per brick -------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NTHREADS, TA_REC_MINB)
rec_blocks_kernel(ScanParams P, RecBuf R, LabelTable lt, PairTable pt) {
    TA_SHARED uint32_t momtab[256];
    TA_SHARED unsigned short list2[RB_WARPS][RB_NBLK], list3[RB_WARPS][RB_NBLK];
    TA_SHARED uint32_t l2lo[RB_WARPS][RB_NBLK], l2hi[RB_WARPS][RB_NBLK];
    TA_SHARED unsigned long long tabmem[RB_WARPS][(WarpTabsSize<T>::value + 7) / 8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256; i += NTHREADS) momtab[i] = block_byte_moments_packed((uint32_t)i);
    WarpTabs<T> tabs;
    tabs.pkey = reinterpret_cast<typename Vox<T>::PKey*>(&tabmem[warp][0]);
    tabs.lkey = reinterpret_cast<uint32_t*>(tabs.pkey + WP_SLOTS);
    tabs.lval = tabs.lkey + WL_SLOTS;
    tabs.pval = tabs.lval + WL_SLOTS * 16;
    warp_tabs_clear<T>(tabs, lane);
    __syncthreads();

    const unsigned int total = (unsigned int)P.nbf * P.nbm * P.nbs;
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    const int nf = (int)P.nf, nm = (int)P.nm, ns = (int)P.ns, noct = R.noct;
    constexpr int SB = BS / BLK_S, MB = BM / BLK_M;          // 4 x 4 blocks per oct column

    for (;;) {
        unsigned int brick = 0u;
        if (lane == 0) brick = atomicAdd(P.brick_counter, 1u);
        brick = __shfl_sync(0xffffffffu, brick, 0);
        // The warps of a CTA work on different bricks but run the same PHASE at the same time (block barriers between the
        // phases): warps scattered over 12 000 instructions miss the instruction cache on almost every fetch (measured:
        // 79 stall cycles per issue, 6 % issue utilisation).  A warp without a brick walks through the barriers.
        const bool have = brick < total;
        if (!__syncthreads_or(have ? 1 : 0)) break;
        bool skip = !have;
        if (!have) brick = 0u;
        const int bf = brick % P.nbf, bm = (brick / P.nbf) % P.nbm, bs = brick / (P.nbf * P.nbm);
        const int F0 = bf * RB_BF, M0 = bm * BM, S0 = (int)P.own_lo + bs * BS, og0 = bf * RB_NOCT;
        const u64 gF0 = (u64)F0, gM0 = (u64)M0, gS0 = (u64)((long long)S0 + P.slow_offset);
        const int vo = min(RB_NOCT, noct - og0);                 // octs of this brick inside the volume

        // ---- U: every record of the tile is one label, the same one --------------------------------------------------------
        {
            uint32_t ref, dummy;
            RecIO<T>::lohi(R, ((S0 - R.plane0) * nm + M0) * noct + og0, ref, dummy);
            bool same = true;
            const int o = lane & 15;
            for (int r0 = 0; r0 < MK_ROWS && same && !skip; r0 += 16) {      // 8 rows at a time for the early exit
                bool ok = true;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = r0 + 2 * k + (lane >> 4);
                    if (r < MK_ROWS && o < vo) {
                        const int tp = r / (BM + 2), tr = r - tp * (BM + 2);
                        const int cs = max(0, min(S0 - 1 + tp, ns - 1)), cm = max(0, min(M0 - 1 + tr, nm - 1));
                        ok = ok && RecIO<T>::is_uniform(RecIO<T>::load(R, ((cs - R.plane0) * nm + cm) * noct + og0 + o), ref);
                    }
                }
                same = __ballot_sync(0xffffffffu, !ok) == 0u;
            }
            if (same && !skip) {
                if (lane == 0 && do_mom) {
                    uint32_t v[16];
                    block_uniform_moments((uint32_t)min(RB_BF, nf - F0), (uint32_t)min(BM, nm - M0), (uint32_t)min(BS, (int)P.own_hi - S0), v);
                    rec_label_to_global(lt, pt.status, ref, v, gF0, gM0, gS0);
                }
                skip = true;
            }
        }
        if (lane == 0 && !skip) TA_STAT(0, 1);

        // ---- P1: 32 blocks at a time ---------------------------------------------------------------------------------------
        int n2 = 0, n3 = 0;
        __syncthreads();
        for (int it = 0; it < (skip ? 0 : RB_NBLK / 32); ++it) {
            const int blk = it * 32 + lane;
            const int o = blk % RB_NOCT, sbq = (blk / RB_NOCT) % SB, mbq = blk / (RB_NOCT * SB);
            const int nvf = min(8, nf - (F0 + 8 * o)), nvm = min(BLK_M, nm - (M0 + BLK_M * mbq)),
                      nvs = min(BLK_S, (int)P.own_hi - (S0 + BLK_S * sbq));
            const bool valid = nvf > 0 && nvm > 0 && nvs > 0;
            uint32_t wlo = 0u, whi = 0u;
            bool two = false;
            if (valid) {
                int po[4], ro[6];
#pragma unroll
                for (int p = 0; p < 4; ++p) po[p] = (max(0, min(S0 - 1 + BLK_S * sbq + p, ns - 1)) - R.plane0) * nm * noct + og0 + o;
#pragma unroll
                for (int r = 0; r < 6; ++r) ro[r] = max(0, min(M0 - 1 + BLK_M * mbq + r, nm - 1)) * noct;
                typename RecIO<T>::Packed rec[24], mmx = RecIO<T>::init();
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int r = 0; r < 6; ++r) { rec[p * 6 + r] = RecIO<T>::load(R, po[p] + ro[r]); mmx = RecIO<T>::acc(mmx, rec[p * 6 + r]); }
                wlo = RecIO<T>::lo_of(mmx); whi = RecIO<T>::hi_of(mmx);
                if (wlo != whi) {
                    uint32_t bad = 0u;
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int r = 0; r < 6; ++r) bad |= RecIO<T>::outside(rec[p * 6 + r], wlo, whi) | (R.q[po[p] + ro[r]] >> 30);
                    two = (bad == 0u) && !TA_FORCE_LIST3;
                }
            }
            const bool one = valid && wlo == whi, many = valid && wlo != whi;
            if (valid) TA_STAT(1, 1);
            if (one) TA_STAT(2, 1);
            if (many) TA_STAT(3, 1);
            const unsigned m2 = __ballot_sync(0xffffffffu, many && two), m3 = __ballot_sync(0xffffffffu, many && !two);
            const unsigned below = (1u << lane) - 1u;
            if (many && two) { const int q = n2 + __popc(m2 & below); list2[warp][q] = (unsigned short)blk; l2lo[warp][q] = wlo; l2hi[warp][q] = whi; }
            if (many && !two) list3[warp][n3 + __popc(m3 & below)] = (unsigned short)blk;
            n2 += __popc(m2); n3 += __popc(m3);
            uint32_t w[MK_ROW];
            const bool has = one && do_mom;
            if (__ballot_sync(0xffffffffu, has)) {
                if (has) {
                    uint32_t v[16];
                    block_uniform_moments((uint32_t)nvf, (uint32_t)nvm, (uint32_t)nvs, v);
                    block_shift_moments(v, (uint32_t)(8 * o), (uint32_t)(BLK_M * mbq), (uint32_t)(BLK_S * sbq));
                    mk_pack_row(v, w);
                }
                rec_put_label<T>(tabs, lt, pt.status, has, wlo, w, gF0, gM0, gS0, lane);
            }
        }
        __syncthreads();

        // ---- P2a: two-label blocks: one mask, its complement ----------------------------------------------------------------
        for (int base = 0; base < n2; base += 32) {
            const int qi = base + lane;
            const bool active = qi < n2;
            if (lane == 0) TA_STAT(4, 1);
            const int blk = active ? (int)list2[warp][qi] : 0;
            const uint32_t wlo = active ? l2lo[warp][qi] : 0u, whi = active ? l2hi[warp][qi] : 0u;
            const int o = blk % RB_NOCT, sbq = (blk / RB_NOCT) % SB, mbq = blk / (RB_NOCT * SB);
            const int nvf = min(8, nf - (F0 + 8 * o)), nvm = min(BLK_M, nm - (M0 + BLK_M * mbq)),
                      nvs = min(BLK_S, (int)P.own_hi - (S0 + BLK_S * sbq));
            const uint32_t bF = (uint32_t)(8 * o), bM = (uint32_t)(BLK_M * mbq), bS = (uint32_t)(BLK_S * sbq);
            constexpr u64 ALL = LvBlk<T>::PLANE_ALL;
            u64 A[4] = {0ull, 0ull, 0ull, 0ull};
            if (active) {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int po = (max(0, min(S0 - 1 + BLK_S * sbq + p, ns - 1)) - R.plane0) * nm * noct + og0 + o;
                    uint32_t h0 = 0u, h1 = 0u;
#pragma unroll
                    for (int r = 0; r < 6; ++r) {
                        const int i = po + max(0, min(M0 - 1 + BLK_M * mbq + r, nm - 1)) * noct;
                        uint32_t lo, hi;
                        RecIO<T>::lohi(R, i, lo, hi);
                        const uint32_t a = (lo == wlo) ? ((R.q[i] & 0x3FFu) ^ 0x3FFu) : 0u;     // not wlo: the record is all whi
                        if (r < 3) h0 |= a << (10 * r); else h1 |= a << (10 * (r - 3));
                    }
                    A[p] = (u64)h0 | ((u64)h1 << 30);
                }
            }
            const u64 B[4] = {~A[0] & ALL, ~A[1] & ALL, ~A[2] & ALL, ~A[3] & ALL};
            u64 cv = 0ull;
#pragma unroll
            for (int r = 1; r <= BLK_M; ++r)
                if (r <= nvm) cv |= (u64)(((1u << nvf) - 1u) << 1) << (10 * r);
            const u64 cv0 = (active && nvs >= 1) ? cv : 0ull, cv1 = (active && nvs >= 2) ? cv : 0ull;
            const u64 ca0 = A[1] & cv0, ca1 = A[2] & cv1, cb0 = B[1] & cv0, cb1 = B[2] & cv1;
            uint32_t inc[4] = {0u, 0u, 0u, 0u};
            MomRow ra, rb;
            ra.has = rb.has = false;
            bool hasp = false;
            if (do_mom) {
                ra = rec_mask_row(ca0, ca1, momtab, bF, bM, bS);
                rb = rec_mask_row(cb0, cb1, momtab, bF, bM, bS);
            }
            if (do_p6 || do_w18) {
                uint32_t w18 = 0u, e0 = 0u, e1 = 0u, e2 = 0u, o0 = 0u, o1 = 0u, o2 = 0u;
                if (do_w18) {
                    u64 DA[2], DB[2];
                    block_dilate18_rb<10>(A, DA);
                    block_dilate18_rb<10>(B, DB);
                    w18 = ta_popc64(ca0 & DB[0]) + ta_popc64(ca1 & DB[1]) + ta_popc64(cb0 & DA[0]) + ta_popc64(cb1 & DA[1]);
                }
                if (do_p6) {            // wlo < whi: faces whose lower-index voxel is wlo go to the even slots
                    e0 = ta_popc64(ca0 & (B[1] >> 1)) + ta_popc64(ca1 & (B[2] >> 1));
                    e1 = ta_popc64(ca0 & (B[1] >> 10)) + ta_popc64(ca1 & (B[2] >> 10));
                    e2 = ta_popc64(ca0 & B[2]) + ta_popc64(ca1 & B[3]);
                    o0 = ta_popc64(cb0 & (A[1] >> 1)) + ta_popc64(cb1 & (A[2] >> 1));
                    o1 = ta_popc64(cb0 & (A[1] >> 10)) + ta_popc64(cb1 & (A[2] >> 10));
                    o2 = ta_popc64(cb0 & A[2]) + ta_popc64(cb1 & A[3]);
                }
                inc[0] = w18 | (e0 << 16); inc[1] = o0 | (e1 << 16); inc[2] = o1 | (e2 << 16); inc[3] = o2;
                hasp = active && (inc[0] | inc[1] | inc[2] | inc[3]) != 0u;
            }
            rec_put_label<T>(tabs, lt, pt.status, ra.has, wlo, ra.w, gF0, gM0, gS0, lane);
            rec_put_label<T>(tabs, lt, pt.status, rb.has, whi, rb.w, gF0, gM0, gS0, lane);
            if (do_p6 || do_w18) rec_put_pair<T>(tabs, pt, hasp, wlo, whi, inc, lane);
        }

        __syncthreads();
        // ---- P2b: the other blocks, label after label --------------------------------------------------------------------------
        for (int base = 0; base < n3; base += 32) {
            const int qi = base + lane;
            const bool active = qi < n3;
            if (lane == 0) TA_STAT(4, 1);
            const int blk = active ? (int)list3[warp][qi] : 0;
            const int o = blk % RB_NOCT, sbq = (blk / RB_NOCT) % SB, mbq = blk / (RB_NOCT * SB);
            const int nvf = min(8, nf - (F0 + 8 * o)), nvm = min(BLK_M, nm - (M0 + BLK_M * mbq)),
                      nvs = min(BLK_S, (int)P.own_hi - (S0 + BLK_S * sbq));
            const uint32_t bF = (uint32_t)(8 * o), bM = (uint32_t)(BLK_M * mbq), bS = (uint32_t)(BLK_S * sbq);
            RecWin W;
#pragma unroll
            for (int p = 0; p < 4; ++p) W.po[p] = (max(0, min(S0 - 1 + BLK_S * sbq + p, ns - 1)) - R.plane0) * nm * noct + og0 + o;
#pragma unroll
            for (int r = 0; r < 6; ++r) W.ro[r] = max(0, min(M0 - 1 + BLK_M * mbq + r, nm - 1)) * noct;
            bool badblk = false;
            if (active) {
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int r = 0; r < 6; ++r) badblk = badblk || ((R.q[W.po[p] + W.ro[r]] >> 30) == 3u);
            }
            BlockLevel<T, MK_MAXL> b;
            b.clear();
            b.set_centre(nvf, nvm, nvs);
            constexpr u64 ALL = LvBlk<T>::PLANE_ALL;
            b.R0 = b.R1 = b.R2 = b.R3 = ALL;
            bool more = active && !badblk;
            rec_steps<T, 0>(tabs, R, W, P, lt, pt, momtab, b, more, bF, bM, bS, gF0, gM0, gS0, lane);
            const bool fb = active && (more || badblk);
            if (fb) TA_STAT(7, 1);
            if (active && badblk) TA_STAT(8, 1);
            // voxels the per-voxel path has to look at: those not covered by a known label, and those with such a voxel
            // among their 18 neighbours (everything else was emitted by the steps); a BAD block: all of them
            u64 nd0, nd1;
            {
                const u64 Rm[4] = {b.R0, b.R1, b.R2, b.R3};
                u64 dR[2];
                block_dilate18_rb<10>(Rm, dR);
                nd0 = badblk ? b.cv0 : ((dR[0] | b.R1) & b.cv0);
                nd1 = badblk ? b.cv1 : ((dR[1] | b.R2) & b.cv1);
            }
            unsigned fm = __ballot_sync(0xffffffffu, fb);
            while (fm) {
                const int src = __ffs(fm) - 1;
                fm &= fm - 1u;
                const int cblk = __shfl_sync(0xffffffffu, blk, src);
                const int nk = __shfl_sync(0xffffffffu, badblk ? 0 : MK_MAXL, src);
                const uint32_t k0 = __shfl_sync(0xffffffffu, b.lab[0], src), k1 = __shfl_sync(0xffffffffu, b.lab[1], src),
                               k2 = __shfl_sync(0xffffffffu, b.lab[2], src), k3 = __shfl_sync(0xffffffffu, b.lab[3], src);
                const u64 c0 = __shfl_sync(0xffffffffu, nd0, src), c1 = __shfl_sync(0xffffffffu, nd1, src);
                const int co = cblk % RB_NOCT, csb = (cblk / RB_NOCT) % SB, cmb = cblk / (RB_NOCT * SB);
                for (int w = lane; w < 64; w += 32) {
                    const int df = w & 7, dm = (w >> 3) & 3, dsx = w >> 5;
                    if (!(((dsx ? c1 : c0) >> (10 * (dm + 1) + df + 1)) & 1ull)) continue;
                    const uint32_t f = (uint32_t)(8 * co + df), m = (uint32_t)(BLK_M * cmb + dm), sp = (uint32_t)(BLK_S * csb + dsx);
                    rec_fallback_voxel<T>(P, lt, pt, F0 + (int)f, M0 + (int)m, S0 + (int)sp, f, m, sp, k0, k1, k2, k3, nk, gF0, gM0, gS0);
                }
            }
        }
        __syncthreads();
        warp_tabs_flush<T>(tabs, lt, pt, gF0, gM0, gS0, lane);
        __syncwarp();
    }
}

}  // namespace ta

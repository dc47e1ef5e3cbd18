// The single streaming pass over the label volume (sm_100a).
//
// Work unit: a brick of BF x BM x BS voxels (BF = 16 segments of 16 bytes) staged in shared memory with a
// one-voxel halo, clamped at the buffer edges (a clamped neighbour equals an in-bounds 6/18-neighbour or the
// voxel itself, which reproduces the reference's "the image border contributes nothing" rule).
//
// Phases per brick (irregular work is compacted into worklists first so whole warps stay busy):
//   A  stage brick + halo: 128-bit streaming loads -> shared tile.
//   B  per row-segment uniformity code: the label if the SEG+2 voxels (segment + f-halo) are equal.
//   C1 march: thread (fseg, m) walks s.  Moments go to two register slots per thread (the column of a thread
//      rarely sees more than two labels): n, sum f, sum s, sum ff, sum fs, sum ss, f/s bounds; the m terms are
//      closed forms.  Uniform segments cost a handful of adds, mixed segments are split into runs with SIMD
//      compares.  At the end of the column the slots are merged across the warp (match.any + redux) and one lane
//      per label updates the per-brick shared label table.  Segments whose 3x3 rows are not one label go to the
//      SEGMENT worklist.
//   C2 per listed segment, SIMD on the packed lanes: OR of XORs of the segment with its 18 neighbour vectors
//      (f-shifted ones built with funnel shifts) -> exact "has a different 18-neighbour" flag per voxel; flagged
//      voxels go to the VOXEL worklist.
//   D  per listed voxel: 18 neighbour labels -> first other label + "only one other label" test.  The common case
//      is merged across the warp (match.any on the pair key, redux of the packed 16-bit counters: wall18 and the
//      +f/+m/+s faces) and one lane per pair updates the per-brick shared pair table.  Junction voxels (>= 2 other
//      labels) go to a third worklist and are handled with a register dedup of up to 4 labels.
//   F  flush the per-brick label table (u32 brick-local sums -> u64 global REDs) and pair table.
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
constexpr int PT_WORDS = 4;             // packed 16-bit counters: [w18|f0] [f1|f2] [f3|f4] [f5|-]
constexpr int TILE_ROWS = (BS + 2) * (BM + 2);
constexpr int ROWV = NFS + 2;           // vectors per tile row
constexpr int PLANEV = (BM + 2) * ROWV; // vectors per tile plane
constexpr int TILE_SEGS = TILE_ROWS * ROWV;
constexpr int SEGLIST_CAP = NFS * BM * BS;
constexpr int VOXLIST_CAP = NTHREADS * 8;

template <typename T> struct Vox;
template <> struct Vox<uint16_t> {
    static constexpr int SEG = 8, LOG_SEG = 3;
    typedef unsigned short Code;                    // label 0xFFFF reads as "mixed": slower exact paths, same result
    static constexpr uint32_t MIXED = 0xFFFFu;
    typedef uint32_t PKey;                          // (lo << 16) | hi
    static constexpr PKey PEMPTY = 0xFFFFFFFFu;
    static __device__ __forceinline__ PKey key(uint32_t a, uint32_t b) { return a < b ? (a << 16) | b : (b << 16) | a; }
    static __device__ __forceinline__ u64 key64(PKey k) { return ((u64)(k >> 16) << 32) | (k & 0xFFFFu); }
    static __device__ __forceinline__ uint32_t hash(PKey k) { return (k * 0x9E3779B1u) >> 24; }
};
template <> struct Vox<uint32_t> {
    static constexpr int SEG = 4, LOG_SEG = 2;
    typedef uint32_t Code;
    static constexpr uint32_t MIXED = 0xFFFFFFFFu;
    typedef u64 PKey;
    static constexpr PKey PEMPTY = TA_EMPTY64;
    static __device__ __forceinline__ PKey key(uint32_t a, uint32_t b) { return ta_pair_key(a, b); }
    static __device__ __forceinline__ u64 key64(PKey k) { return k; }
    static __device__ __forceinline__ uint32_t hash(PKey k) {
        return (((uint32_t)(k >> 32) * 0x9E3779B1u) ^ ((uint32_t)k * 0x85EBCA77u)) >> 24;
    }
};

template <typename T> constexpr size_t scan_smem_bytes() {
    return (size_t)TILE_SEGS * 16 + (size_t)TILE_ROWS * NFS * sizeof(typename Vox<T>::Code) + LT_SLOTS * 4 +
           LT_SLOTS * LT_FIELDS * 4 + PT_SLOTS * sizeof(typename Vox<T>::PKey) + PT_SLOTS * PT_WORDS * 4 +
           SEGLIST_CAP * 2 + 2 * VOXLIST_CAP * 2 + 32;
}

__device__ __forceinline__ uint4 ld_stream_128(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <typename T> struct BrickShared {
    uint4* tile;                       // [TILE_SEGS]
    typename Vox<T>::Code* codes;      // [TILE_ROWS * NFS]
    uint32_t* lt_key;                  // [LT_SLOTS]
    uint32_t* lt_val;                  // [LT_SLOTS * LT_FIELDS]
    typename Vox<T>::PKey* pt_key;     // [PT_SLOTS]
    uint32_t* pt_val;                  // [PT_SLOTS * PT_WORDS]
    unsigned short* seglist;           // [SEGLIST_CAP]
    unsigned short* voxlist;           // [VOXLIST_CAP]
    unsigned short* junclist;          // [VOXLIST_CAP]
    unsigned int* ctr;                 // [0] next brick, [1] nseg, [2..3] nvox ping-pong, [4..5] njunc ping-pong
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
template <typename T>
__device__ __forceinline__ void label_add(const BrickShared<T>& sh, const LabelTable& lt, uint32_t* status,
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

// all 32 lanes call; lanes with L == TA_EMPTY32 contribute nothing.  One shared-table update per distinct label.
template <typename T>
__device__ __forceinline__ void label_add_warp(const BrickShared<T>& sh, const LabelTable& lt, uint32_t* status,
                                               uint32_t L, const uint32_t* v, u64 F0, u64 M0, u64 S0, int lane) {
    const unsigned grp = __match_any_sync(0xffffffffu, L);
    uint32_t r[LT_FIELDS];
#pragma unroll
    for (int i = 0; i < 10; ++i) r[i] = __reduce_add_sync(grp, v[i]);
#pragma unroll
    for (int i = 10; i < 13; ++i) r[i] = __reduce_min_sync(grp, v[i]);
#pragma unroll
    for (int i = 13; i < 16; ++i) r[i] = __reduce_max_sync(grp, v[i]);
    if (L != TA_EMPTY32 && lane == __ffs(grp) - 1) label_add(sh, lt, status, L, r, F0, M0, S0);
}

// thread-private moment accumulator for one label over the thread's (fseg, m) column
struct MomSlot {
    uint32_t label, n, sf, ss, sff, sfs, sss, fmin, fmax, smin, smax;
    __device__ __forceinline__ void reset(uint32_t L) {
        label = L; n = sf = ss = sff = sfs = sss = 0u; fmin = 0xFFFFFFFFu; fmax = 0u; smin = 0xFFFFFFFFu; smax = 0u;
    }
    __device__ __forceinline__ void add(uint32_t len, uint32_t sfr, uint32_t sffr, uint32_t s, uint32_t f_first,
                                        uint32_t f_last) {
        n += len; sf += sfr; ss += len * s; sff += sffr; sfs += s * sfr; sss += len * s * s;
        fmin = min(fmin, f_first); fmax = max(fmax, f_last); smin = min(smin, s); smax = max(smax, s);
    }
    __device__ __forceinline__ void fields(uint32_t v[LT_FIELDS], uint32_t m) const {
        v[0] = n; v[1] = sf; v[2] = n * m; v[3] = ss; v[4] = sff; v[5] = m * sf; v[6] = sfs;
        v[7] = n * m * m; v[8] = m * ss; v[9] = sss;
        v[10] = fmin; v[11] = m; v[12] = smin; v[13] = fmax; v[14] = m; v[15] = smax;
    }
};

// ---- per-brick pair accumulation (packed 16-bit counters; a brick has < 65536 voxels) ------------------------
// field 6 = wall18, fields 0..5 = directional faces.  idx = field+1 (wall18 -> 0): word idx>>1, half idx&1.
template <typename T>
__device__ __forceinline__ void pair_add_packed(const BrickShared<T>& sh, const PairTable& pt,
                                                typename Vox<T>::PKey key, const uint32_t inc[PT_WORDS]) {
    typedef typename Vox<T>::PKey PKey;
    uint32_t slot = Vox<T>::hash(key);
    for (int probe = 0; probe < PT_SLOTS; ++probe) {
        PKey k = *((volatile PKey*)&sh.pt_key[slot]);
        bool hit = (k == key);
        if (!hit && k == Vox<T>::PEMPTY) {
            PKey old = atomicCAS(&sh.pt_key[slot], Vox<T>::PEMPTY, key);
            hit = (old == Vox<T>::PEMPTY || old == key);
        }
        if (hit) {
#pragma unroll
            for (int w = 0; w < PT_WORDS; ++w) if (inc[w]) atomicAdd(&sh.pt_val[slot * PT_WORDS + w], inc[w]);
            return;
        }
        slot = (slot + 1) & (PT_SLOTS - 1);
    }
    int g = ta_pair_slot(pt, Vox<T>::key64(key));          // brick table full: straight to the global table
    if (g < 0) return;
    uint32_t* v = &pt.vals[(size_t)g * TA_PAIR_STRIDE];
#pragma unroll
    for (int idx = 0; idx < 7; ++idx) {
        uint32_t n = (inc[idx >> 1] >> ((idx & 1) * 16)) & 0xFFFFu;
        if (n) atomicAdd(&v[idx == 0 ? 6 : idx - 1], n);
    }
}

template <typename T>
__device__ __forceinline__ void pair_add(const BrickShared<T>& sh, const PairTable& pt, uint32_t a, uint32_t b,
                                         int field, uint32_t n) {
    uint32_t inc[PT_WORDS] = {0, 0, 0, 0};
    int idx = field == 6 ? 0 : field + 1;
    inc[idx >> 1] = n << ((idx & 1) * 16);
    pair_add_packed<T>(sh, pt, Vox<T>::key(a, b), inc);
}

// packed increments of one voxel (label a) towards other label d: wall18 + the faces to its +f/+m/+s neighbours
__device__ __forceinline__ void voxel_increments(uint32_t inc[PT_WORDS], bool lo, bool w18, bool ff, bool fm, bool fsl) {
    inc[0] = (w18 ? 1u : 0u) + ((ff && lo) ? (1u << 16) : 0u);
    inc[1] = ((ff && !lo) ? 1u : 0u) + ((fm && lo) ? (1u << 16) : 0u);
    inc[2] = ((fm && !lo) ? 1u : 0u) + ((fsl && lo) ? (1u << 16) : 0u);
    inc[3] = ((fsl && !lo) ? 1u : 0u);
}

__device__ __forceinline__ uint32_t sumsq_upto(uint32_t k) { return k * (k + 1) * (2 * k + 1) / 6; }  // 0..k

// ---- SIMD helpers on one 16-byte segment ------------------------------------------------------------------------
template <typename T> struct Boundary;

template <> struct Boundary<uint16_t> {
    static __device__ __forceinline__ void cross(uint32_t acc[4], const uint4& C, const uint4* tile, int t,
                                                 bool unshifted) {
        const uint4 R = tile[t];
        const uint32_t e0 = (uint32_t)(reinterpret_cast<const unsigned short*>(tile + t)[-1]) << 16;
        const uint32_t e5 = reinterpret_cast<const unsigned short*>(tile + t + 1)[0];
        const uint32_t s0 = __funnelshift_r(e0, R.x, 16), s1 = __funnelshift_r(R.x, R.y, 16),
                       s2 = __funnelshift_r(R.y, R.z, 16), s3 = __funnelshift_r(R.z, R.w, 16),
                       s4 = __funnelshift_r(R.w, e5, 16);
        acc[0] |= (C.x ^ s0) | (C.x ^ s1);
        acc[1] |= (C.y ^ s1) | (C.y ^ s2);
        acc[2] |= (C.z ^ s2) | (C.z ^ s3);
        acc[3] |= (C.w ^ s3) | (C.w ^ s4);
        if (unshifted) { acc[0] |= C.x ^ R.x; acc[1] |= C.y ^ R.y; acc[2] |= C.z ^ R.z; acc[3] |= C.w ^ R.w; }
    }
    static __device__ __forceinline__ void diag(uint32_t acc[4], const uint4& C, const uint4* tile, int t) {
        const uint4 R = tile[t];
        acc[0] |= C.x ^ R.x; acc[1] |= C.y ^ R.y; acc[2] |= C.z ^ R.z; acc[3] |= C.w ^ R.w;
    }
    static __device__ __forceinline__ uint32_t mask(const uint32_t acc[4]) {
        const uint32_t one = 0x00010001u;
        uint32_t tt = __vminu2(acc[0], one) | (__vminu2(acc[1], one) << 2) | (__vminu2(acc[2], one) << 4) |
                      (__vminu2(acc[3], one) << 6);
        return (tt & 0x55u) | ((tt >> 15) & 0xAAu);
    }
    // bit j set iff lane j differs from lane j+1 (j = 0..6)
    static __device__ __forceinline__ uint32_t run_breaks(const uint4& C) {
        uint32_t d[4] = {C.x ^ __funnelshift_r(C.x, C.y, 16), C.y ^ __funnelshift_r(C.y, C.z, 16),
                         C.z ^ __funnelshift_r(C.z, C.w, 16), (C.w ^ (C.w >> 16)) & 0xFFFFu};
        return mask(d) & 0x7Fu;
    }
};

template <> struct Boundary<uint32_t> {
    static __device__ __forceinline__ void cross(uint32_t acc[4], const uint4& C, const uint4* tile, int t,
                                                 bool unshifted) {
        const uint4 R = tile[t];
        const uint32_t e0 = reinterpret_cast<const uint32_t*>(tile + t)[-1];
        const uint32_t e5 = reinterpret_cast<const uint32_t*>(tile + t + 1)[0];
        acc[0] |= (C.x ^ e0) | (C.x ^ R.y);
        acc[1] |= (C.y ^ R.x) | (C.y ^ R.z);
        acc[2] |= (C.z ^ R.y) | (C.z ^ R.w);
        acc[3] |= (C.w ^ R.z) | (C.w ^ e5);
        if (unshifted) { acc[0] |= C.x ^ R.x; acc[1] |= C.y ^ R.y; acc[2] |= C.z ^ R.z; acc[3] |= C.w ^ R.w; }
    }
    static __device__ __forceinline__ void diag(uint32_t acc[4], const uint4& C, const uint4* tile, int t) {
        const uint4 R = tile[t];
        acc[0] |= C.x ^ R.x; acc[1] |= C.y ^ R.y; acc[2] |= C.z ^ R.z; acc[3] |= C.w ^ R.w;
    }
    static __device__ __forceinline__ uint32_t mask(const uint32_t acc[4]) {
        return (acc[0] ? 1u : 0u) | (acc[1] ? 2u : 0u) | (acc[2] ? 4u : 0u) | (acc[3] ? 8u : 0u);
    }
    static __device__ __forceinline__ uint32_t run_breaks(const uint4& C) {
        return (C.x != C.y ? 1u : 0u) | (C.y != C.z ? 2u : 0u) | (C.z != C.w ? 4u : 0u);
    }
};

// k-th offset of the 18-neighbourhood (1 <= |df|+|dm|+|ds| <= 2) in tile elements; rare-path helper
template <int ROWE, int PLANEE>
__device__ __noinline__ int neighbour_offset(int k) {
    int c = 0;
    for (int i = 0; i < 27; ++i) {
        int df = i % 3 - 1, dm = (i / 3) % 3 - 1, ds = i / 9 - 1;
        int l1 = abs(df) + abs(dm) + abs(ds);
        if (l1 >= 1 && l1 <= 2) {
            if (c == k) return ds * PLANEE + dm * ROWE + df;
            ++c;
        }
    }
    return 0;
}

template <typename T>
__global__ void __launch_bounds__(NTHREADS, 2)
scan_kernel(ScanParams P, LabelTable lt, PairTable pt) {
    typedef typename Vox<T>::Code Code;
    typedef typename Vox<T>::PKey PKey;
    constexpr int SEG = Vox<T>::SEG;
    constexpr int LOG_SEG = Vox<T>::LOG_SEG;
    constexpr uint32_t MIXED = Vox<T>::MIXED;
    constexpr int ROWE = ROWV * SEG;               // elements per tile row
    constexpr int PLANEE = (BM + 2) * ROWE;        // elements per tile plane
    constexpr int BF = NFS * SEG;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    BrickShared<T> sh;
    sh.tile = reinterpret_cast<uint4*>(smem_raw);
    sh.lt_key = reinterpret_cast<uint32_t*>(sh.tile + TILE_SEGS);
    sh.lt_val = sh.lt_key + LT_SLOTS;
    sh.pt_val = sh.lt_val + LT_SLOTS * LT_FIELDS;
    sh.pt_key = reinterpret_cast<PKey*>(sh.pt_val + PT_SLOTS * PT_WORDS);
    sh.ctr = reinterpret_cast<unsigned int*>(sh.pt_key + PT_SLOTS);
    sh.codes = reinterpret_cast<Code*>(sh.ctr + 8);
    sh.seglist = reinterpret_cast<unsigned short*>(sh.codes + TILE_ROWS * NFS);
    sh.voxlist = sh.seglist + SEGLIST_CAP;
    sh.junclist = sh.voxlist + VOXLIST_CAP;
    const T* tileT = reinterpret_cast<const T*>(sh.tile);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const T* vol = reinterpret_cast<const T*>(P.vol);
    const unsigned int total = (unsigned int)P.nbf * P.nbm * P.nbs;
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    const bool do_pairs = do_p6 || do_w18;
    const int nf = (int)P.nf, nm = (int)P.nm, ns = (int)P.ns;

    // reset the per-brick tables once; the flush at the end of each brick re-arms them
    for (int i = tid; i < LT_SLOTS; i += NTHREADS) sh.lt_key[i] = TA_EMPTY32;
    for (int i = tid; i < LT_SLOTS * LT_FIELDS; i += NTHREADS) {
        int f = i % LT_FIELDS;
        sh.lt_val[i] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
    }
    for (int i = tid; i < PT_SLOTS; i += NTHREADS) sh.pt_key[i] = Vox<T>::PEMPTY;
    for (int i = tid; i < PT_SLOTS * PT_WORDS; i += NTHREADS) sh.pt_val[i] = 0u;

    for (;;) {
        if (tid == 0) {
            sh.ctr[0] = atomicAdd(P.brick_counter, 1u);
            sh.ctr[1] = 0u; sh.ctr[2] = 0u; sh.ctr[3] = 0u; sh.ctr[4] = 0u; sh.ctr[5] = 0u;
        }
        __syncthreads();
        const unsigned int brick = sh.ctr[0];
        if (brick >= total) break;
        const int bf = brick % P.nbf, bm = (brick / P.nbf) % P.nbm, bs = brick / (P.nbf * P.nbm);
        const int F0 = bf * BF, M0 = bm * BM, S0 = (int)P.own_lo + bs * BS;
        const u64 gF0 = (u64)F0, gM0 = (u64)M0, gS0 = (u64)((long long)S0 + P.slow_offset);

        // ---- phase A: stage brick + halo (clamped) ------------------------------------------------------------
        for (int i = tid; i < TILE_SEGS; i += NTHREADS) {
            const int fs = i % ROWV - 1;
            const int r = i / ROWV;
            const int m = r % (BM + 2) - 1, s = r / (BM + 2) - 1;
            const int gs = min(max(S0 + s, 0), ns - 1);
            const int gm = min(max(M0 + m, 0), nm - 1);
            const int gf = F0 + fs * SEG;
            const T* row = vol + ((size_t)gs * nm + gm) * (size_t)nf;
            uint4 v;
            if (P.vec_ok) {
                // rows are whole segments: out-of-range halo segments replicate the edge voxel
                const int gfc = min(max(gf, 0), nf - SEG);
                v = ld_stream_128(row + gfc);
                if (gf != gfc) {
                    uint32_t e = (gf < 0) ? ((SEG == 8) ? (v.x & 0xFFFFu) : v.x) : ((SEG == 8) ? (v.w >> 16) : v.w);
                    if (SEG == 8) e |= e << 16;
                    v.x = v.y = v.z = v.w = e;
                }
            } else {
                T tmp[SEG];
#pragma unroll
                for (int j = 0; j < SEG; ++j) tmp[j] = row[min(max(gf + j, 0), nf - 1)];
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

        if (P.flags & 0x100u) continue;   // debug: staging only
        // ---- phase B: per row-segment uniformity code (label if the SEG+2 voxels are equal) -------------------
        for (int i = tid; i < TILE_ROWS * NFS; i += NTHREADS) {
            const int fs = i % NFS, r = i / NFS;
            const T* rp = tileT + r * ROWE + (fs + 1) * SEG;
            const uint4 v = sh.tile[r * ROWV + fs + 1];
            const uint32_t l = rp[0];
            const uint32_t pat = (SEG == 8) ? (l | (l << 16)) : l;
            const bool uni = (v.x == pat) & (v.y == pat) & (v.z == pat) & (v.w == pat) &
                             ((uint32_t)rp[-1] == l) & ((uint32_t)rp[SEG] == l);
            sh.codes[i] = (Code)(uni ? l : MIXED);
        }
        __syncthreads();

        if (P.flags & 0x200u) continue;   // debug: staging + codes only
        // ---- phase C1: march (moments, interior test, segment worklist) -------------------------------------------
        {
            const int fs = tid % NFS, m = tid / NFS;
            const int gf0 = F0 + fs * SEG, gm = M0 + m;
            const bool col_valid = (gf0 < nf) && (gm < nm);
            const int nvalid = col_valid ? min(SEG, nf - gf0) : 0;
            const int smax = min(BS, (int)P.own_hi - S0);
            const uint32_t lf0 = fs * SEG;
            const uint32_t rowsum = SEG * lf0 + SEG * (SEG - 1) / 2;
            const uint32_t rowsq = SEG * lf0 * lf0 + lf0 * SEG * (SEG - 1) + (SEG - 1) * SEG * (2 * SEG - 1) / 6;

            MomSlot A, B;
            A.reset(TA_EMPTY32); B.reset(TA_EMPTY32);
            bool mruA = true;
            auto evict = [&](MomSlot& sl) {
                if (sl.n) {
                    uint32_t v[LT_FIELDS];
                    sl.fields(v, (uint32_t)m);
                    label_add<T>(sh, lt, pt.status, sl.label, v, gF0, gM0, gS0);
                }
            };
            auto account = [&](uint32_t L, uint32_t len, uint32_t sfr, uint32_t sffr, uint32_t s, uint32_t f0,
                               uint32_t f1) {
                if (L == A.label) { A.add(len, sfr, sffr, s, f0, f1); mruA = true; }
                else if (L == B.label) { B.add(len, sfr, sffr, s, f0, f1); mruA = false; }
                else if (mruA) { evict(B); B.reset(L); B.add(len, sfr, sffr, s, f0, f1); mruA = false; }
                else { evict(A); A.reset(L); A.add(len, sfr, sffr, s, f0, f1); mruA = true; }
            };
            auto tcode = [&](int s) -> uint32_t {
                const int base = ((s + 1) * (BM + 2) + (m + 1)) * NFS + fs;
                const uint32_t e0 = sh.codes[base - NFS], e1 = sh.codes[base], e2 = sh.codes[base + NFS];
                return (e0 == e1 && e1 == e2) ? e1 : MIXED;
            };

            uint32_t t_prev = tcode(-1), t_cur = tcode(0);
            for (int s = 0; s < BS; ++s) {
                const bool active = col_valid && (s < smax);
                const uint32_t t_next = tcode(s + 1);
                const uint32_t e_c = sh.codes[((s + 1) * (BM + 2) + (m + 1)) * NFS + fs];
                const bool interior = (t_cur != MIXED) && (t_prev == t_cur) && (t_next == t_cur);

                if (do_mom && active) {
                    if (e_c != MIXED && nvalid == SEG) {
                        account(e_c, SEG, rowsum, rowsq, (uint32_t)s, lf0, lf0 + SEG - 1);
                    } else {
                        const int tv = (s + 1) * PLANEV + (m + 1) * ROWV + (fs + 1);
                        const uint4 C = sh.tile[tv];
                        const T* cp = reinterpret_cast<const T*>(sh.tile + tv);
                        uint32_t brk = Boundary<T>::run_breaks(C);
                        int j0 = 0;
                        while (j0 < nvalid) {
                            const uint32_t rest = brk >> j0;
                            int j1 = rest ? j0 + __ffs(rest) : SEG;
                            j1 = min(j1, nvalid);
                            const uint32_t L = cp[j0];
                            const uint32_t len = j1 - j0;
                            const uint32_t sj = (uint32_t)(j0 + j1 - 1) * len / 2;
                            const uint32_t sjj = sumsq_upto(j1 - 1) - (j0 > 0 ? sumsq_upto(j0 - 1) : 0u);
                            account(L, len, len * lf0 + sj, len * lf0 * lf0 + 2 * lf0 * sj + sjj, (uint32_t)s,
                                    lf0 + j0, lf0 + j1 - 1);
                            j0 = j1;
                        }
                    }
                }
                // warp-aggregated append of non-interior segments
                const bool want = do_pairs && active && !interior;
                const unsigned ball = __ballot_sync(0xffffffffu, want);
                if (ball) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(&sh.ctr[1], (unsigned)__popc(ball));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (want) sh.seglist[base + __popc(ball & ((1u << lane) - 1u))] =
                        (unsigned short)((s * BM + m) * NFS + fs);
                }
                t_prev = t_cur; t_cur = t_next;
            }
            if (do_mom) {
                uint32_t v[LT_FIELDS];
                A.fields(v, (uint32_t)m);
                label_add_warp<T>(sh, lt, pt.status, A.n ? A.label : TA_EMPTY32, v, gF0, gM0, gS0, lane);
                B.fields(v, (uint32_t)m);
                label_add_warp<T>(sh, lt, pt.status, B.n ? B.label : TA_EMPTY32, v, gF0, gM0, gS0, lane);
            }
        }
        __syncthreads();

        // ---- phases C2 + D in rounds of NTHREADS listed segments ------------------------------------------------------
        if (do_pairs) {
            const int nseg = (int)sh.ctr[1];
            for (int base = 0, round = 0; base < nseg; base += NTHREADS, ++round) {
                unsigned int* nvox = &sh.ctr[2 + (round & 1)];
                unsigned int* njunc = &sh.ctr[4 + (round & 1)];
                if (tid == 0) { sh.ctr[2 + ((round + 1) & 1)] = 0u; sh.ctr[4 + ((round + 1) & 1)] = 0u; }
                // C2: boundary bits of one listed segment per thread
                const int idx = base + tid;
                uint32_t bits = 0, id = 0;
                if (idx < nseg) {
                    id = sh.seglist[idx];
                    const int fs = id % NFS, m = (id / NFS) % BM, s = id / (NFS * BM);
                    const int t = (s + 1) * PLANEV + (m + 1) * ROWV + (fs + 1);
                    const uint4 C = sh.tile[t];
                    uint32_t acc[4] = {0u, 0u, 0u, 0u};
                    Boundary<T>::cross(acc, C, sh.tile, t, false);
                    Boundary<T>::cross(acc, C, sh.tile, t - ROWV, true);
                    Boundary<T>::cross(acc, C, sh.tile, t + ROWV, true);
                    Boundary<T>::cross(acc, C, sh.tile, t - PLANEV, true);
                    Boundary<T>::cross(acc, C, sh.tile, t + PLANEV, true);
                    Boundary<T>::diag(acc, C, sh.tile, t - PLANEV - ROWV);
                    Boundary<T>::diag(acc, C, sh.tile, t - PLANEV + ROWV);
                    Boundary<T>::diag(acc, C, sh.tile, t + PLANEV - ROWV);
                    Boundary<T>::diag(acc, C, sh.tile, t + PLANEV + ROWV);
                    bits = Boundary<T>::mask(acc);
                    const int left = nf - (F0 + fs * SEG);
                    if (left < SEG) bits &= (1u << left) - 1u;
                }
                {
                    // warp exclusive scan of popcounts, one shared atomic per warp
                    const unsigned cnt = __popc(bits);
                    unsigned inc = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        unsigned y = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += y;
                    }
                    unsigned wtot = __shfl_sync(0xffffffffu, inc, 31);
                    unsigned wbase = 0;
                    if (lane == 31 && wtot) wbase = atomicAdd(nvox, wtot);
                    wbase = __shfl_sync(0xffffffffu, wbase, 31);
                    unsigned pos = wbase + inc - cnt;
                    while (bits) {
                        int j = __ffs(bits) - 1;
                        bits &= bits - 1;
                        sh.voxlist[pos++] = (unsigned short)((id << LOG_SEG) | j);
                    }
                }
                __syncthreads();

                // D: listed voxels, one per thread per iteration; simple voxels are merged across the warp
                const int nv = (int)*nvox;
                for (int ib = 0; ib < nv; ib += NTHREADS) {
                    const int i = ib + tid;
                    PKey key = Vox<T>::PEMPTY;
                    uint32_t inc[PT_WORDS] = {0u, 0u, 0u, 0u};
                    bool junction = false;
                    uint32_t e = 0;
                    if (i < nv) {
                        e = sh.voxlist[i];
                        const uint32_t sid = e >> LOG_SEG;
                        const int j = e & (SEG - 1);
                        const int fs = sid % NFS, m = (sid / NFS) % BM, s = sid / (NFS * BM);
                        const T* p = tileT + (s + 1) * PLANEE + (m + 1) * ROWE + (fs + 1) * SEG + j;
                        const uint32_t a = p[0];
                        constexpr int offs[18] = {
                            1, ROWE, PLANEE, -1, -ROWE, -PLANEE,
                            -ROWE - 1, -ROWE + 1, ROWE - 1, ROWE + 1,
                            -PLANEE - 1, -PLANEE + 1, PLANEE - 1, PLANEE + 1,
                            -PLANEE - ROWE, -PLANEE + ROWE, PLANEE - ROWE, PLANEE + ROWE};
                        uint32_t nb[18];
#pragma unroll
                        for (int k = 0; k < 18; ++k) nb[k] = p[offs[k]];
                        uint32_t d0 = a;
                        bool simple = true;
#pragma unroll
                        for (int k = 0; k < 18; ++k) {
                            const uint32_t b = nb[k];
                            const bool ne = b != a;
                            const bool first = ne && (d0 == a);
                            simple = simple && (!ne || first || b == d0);
                            d0 = first ? b : d0;
                        }
                        if (d0 != a) {
                            if (simple) {
                                key = Vox<T>::key(a, d0);
                                voxel_increments(inc, a < d0, do_w18, do_p6 && nb[0] != a, do_p6 && nb[1] != a,
                                                 do_p6 && nb[2] != a);
                            } else {
                                junction = true;
                            }
                        }
                    }
                    // one shared-table update per distinct pair in the warp
                    {
                        const unsigned grp = __match_any_sync(0xffffffffu, key);
                        uint32_t tot[PT_WORDS];
#pragma unroll
                        for (int w = 0; w < PT_WORDS; ++w) tot[w] = __reduce_add_sync(grp, inc[w]);
                        if (key != Vox<T>::PEMPTY && lane == __ffs(grp) - 1) pair_add_packed<T>(sh, pt, key, tot);
                    }
                    // junction voxels -> third worklist
                    {
                        const unsigned ball = __ballot_sync(0xffffffffu, junction);
                        if (ball) {
                            unsigned jb = 0;
                            if (lane == 0) jb = atomicAdd(njunc, (unsigned)__popc(ball));
                            jb = __shfl_sync(0xffffffffu, jb, 0);
                            if (junction) sh.junclist[jb + __popc(ball & ((1u << lane) - 1u))] = (unsigned short)e;
                        }
                    }
                }
                __syncthreads();

                // D2: junction voxels: distinct other labels in registers (up to 4), one packed add per label
                const int nj = (int)*njunc;
                for (int i = tid; i < nj; i += NTHREADS) {
                    const uint32_t e = sh.junclist[i];
                    const uint32_t sid = e >> LOG_SEG;
                    const int j = e & (SEG - 1);
                    const int fs = sid % NFS, m = (sid / NFS) % BM, s = sid / (NFS * BM);
                    const T* p = tileT + (s + 1) * PLANEE + (m + 1) * ROWE + (fs + 1) * SEG + j;
                    const uint32_t a = p[0];
                    constexpr int offs[18] = {
                        1, ROWE, PLANEE, -1, -ROWE, -PLANEE,
                        -ROWE - 1, -ROWE + 1, ROWE - 1, ROWE + 1,
                        -PLANEE - 1, -PLANEE + 1, PLANEE - 1, PLANEE + 1,
                        -PLANEE - ROWE, -PLANEE + ROWE, PLANEE - ROWE, PLANEE + ROWE};
                    uint32_t nb[18];
#pragma unroll
                    for (int k = 0; k < 18; ++k) nb[k] = p[offs[k]];
                    uint32_t d0 = a, d1 = a, d2 = a, d3 = a;
                    int nd = 0;
#pragma unroll
                    for (int k = 0; k < 18; ++k) {
                        const uint32_t b = nb[k];
                        const bool isnew = (b != a) & (b != d0) & (b != d1) & (b != d2) & (b != d3);
                        d0 = (isnew && nd == 0) ? b : d0;
                        d1 = (isnew && nd == 1) ? b : d1;
                        d2 = (isnew && nd == 2) ? b : d2;
                        d3 = (isnew && nd == 3) ? b : d3;
                        nd += isnew ? 1 : 0;
                    }
                    if (nd <= 4) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t d = q == 0 ? d0 : q == 1 ? d1 : q == 2 ? d2 : d3;
                            if (q < nd) {
                                uint32_t inc[PT_WORDS];
                                voxel_increments(inc, a < d, do_w18, do_p6 && nb[0] == d, do_p6 && nb[1] == d,
                                                 do_p6 && nb[2] == d);
                                pair_add_packed<T>(sh, pt, Vox<T>::key(a, d), inc);
                            }
                        }
                    } else {
                        // more than four distinct other labels (noise-like data): exact first-occurrence rescan
                        if (do_p6) {
                            if (nb[0] != a) pair_add<T>(sh, pt, a, nb[0], a < nb[0] ? 0 : 1, 1u);
                            if (nb[1] != a) pair_add<T>(sh, pt, a, nb[1], a < nb[1] ? 2 : 3, 1u);
                            if (nb[2] != a) pair_add<T>(sh, pt, a, nb[2], a < nb[2] ? 4 : 5, 1u);
                        }
                        if (do_w18) {
#pragma unroll 1
                            for (int k = 0; k < 18; ++k) {
                                const uint32_t b = p[neighbour_offset<ROWE, PLANEE>(k)];
                                if (b == a) continue;
                                bool seen = false;
                                for (int q = 0; q < k; ++q)
                                    seen |= ((uint32_t)p[neighbour_offset<ROWE, PLANEE>(q)] == b);
                                if (!seen) pair_add<T>(sh, pt, a, b, 6, 1u);
                            }
                        }
                    }
                }
                __syncthreads();
            }
        }

        // ---- phase F: flush the per-brick tables ---------------------------------------------------------------------
        if (!(P.flags & 0x400u)) {
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
                PKey key = sh.pt_key[i];
                if (key == Vox<T>::PEMPTY) continue;
                uint32_t* d = &sh.pt_val[i * PT_WORDS];
                int slot = ta_pair_slot(pt, Vox<T>::key64(key));
#pragma unroll
                for (int idx = 0; idx < 7; ++idx) {
                    uint32_t n = (d[idx >> 1] >> ((idx & 1) * 16)) & 0xFFFFu;
                    if (n && slot >= 0) atomicAdd(&pt.vals[(size_t)slot * TA_PAIR_STRIDE + (idx == 0 ? 6 : idx - 1)], n);
                }
#pragma unroll
                for (int w = 0; w < PT_WORDS; ++w) d[w] = 0u;
                sh.pt_key[i] = Vox<T>::PEMPTY;
            }
        }
        __syncthreads();
    }
}

}  // namespace ta

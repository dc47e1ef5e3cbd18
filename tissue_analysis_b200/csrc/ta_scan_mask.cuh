// The streaming pass, bit-mask formulation (sm_100a).  Same tables as scan_kernel (ta_scan.cuh), different arithmetic:
// a row of 32 voxels becomes one 32-bit mask per label it holds; everything after that -- 18-neighbourhood dilation, face
// tests, moments -- is boolean algebra and popcounts on 32-voxel words, warp-uniform, without worklists.
//
// Work unit: a brick of RW x OM x ZB = 32 x 30 x 8 voxels, staged with a one-voxel halo as ONE TMA box of
// (32 + 2 segments of 16 bytes) x 32 rows x 10 planes = 320 tile rows of 32 (+ 2 halo) voxels.  A lane owns a row.
//
//   P1  warp w takes a block of 16 rows x 2 planes of the tile, one row per lane.  The lane finds the run starts of its
//       row (one packed compare of the row with itself shifted by a voxel: SHF + LOP3 + VIMNMX.U16x2 + IMAD per two
//       voxels) and walks its runs: label from the tile, slot = the label's place in the brick's open-addressing list
//       (blab[K], CAS insert, no waiting), bits [j0, j1) OR-ed into masks[plane][slot][row] (all masks are zero when P1
//       starts: the end of a brick clears the slots it used).  hwl / hwr[plane][row] = the slots of the two halo voxels,
//       pres[plane][row] = the slots present in the row.  More than K labels in one brick (noise, small cells): the whole
//       brick takes the per-voxel path G.
//   --  barrier; the tile is dead now: thread 0 issues the NEXT brick's box copy, it lands under P2.
//   P2  warp w < 8 takes a block of 15 owned rows x 2 owned planes.  For every label b present in the rows around the
//       block (redux.or of the row presence words):
//         D_b = dilation of b by the 18-neighbourhood, restricted to the lane's row -- ORs of the nine row masks around,
//               two shifts for the f direction; Bf / Bm / Bs = b as the +f / +m / +s neighbour;
//         the labels a of the block that meet D_b somewhere (8 masks in registers per round, one LOP3 + one predicated OR
//         each, one redux.or) get wall18 += popc(M_a & D_b), faces += popc(M_a & Bf), popc(M_a & Bm), popc(M_a & Bs), two
//         full-mask redux per (a, b); the results wait in one lane each and go to the per-brick pair table in one SIMT pass.
//       Moments: per label of the block n = popc(M), closed forms for sum f, sum f^2 of a run of bits, 6 redux, bounds from
//       redux.or / ballot, one lane per label updates the per-brick label table.
//       Warps 8 and 9 have no block: they flush the tables of the PREVIOUS brick meanwhile (F below; the per-brick tables
//       are ping-pong).  Phase clocks before this: the flush was 17 % of warp 0's time per brick and made it late for the
//       next P1, where the other warps then waited for it.
//   G   (rare) every voxel of the brick against its 18 neighbours, straight from the tile.
//   F   flush a set of per-brick tables (brick-local u32 sums -> shifted u64 global REDs; pair slots -> global hash).
//   A brick whose tile is one label altogether (background, inside of a big cell) ends after P1 with closed-form moments;
//   consecutive such bricks of one label are merged in shared memory before they touch the global table.  It uses no
//   tables, so a pending flush waits for the next brick that has a P2 to hide it behind.
//   Everything rare (G, edge replication of a tile at the volume border, the closed form, the pair-table insert) is a
//   __noinline__ function and the hash probe loops are not unrolled: the hot loop body has to stay near the 32 KB of
//   instruction cache an SM has (ncu: 1.9 `no_instruction` stall cycles per issue with everything inlined, 0.5 now).
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
constexpr int K = 24;                     // label slots per brick (16: C2 and C4 take the per-voxel path often, four times slower)
constexpr int KB = 6;                     // labels of a P2 block held in registers at a time (8: 3 % slower on C3, 5 % on C2; 4: 1 %)
constexpr int NW2 = 8;                    // warps with a P2 block (15 rows x 2 planes each)
constexpr int MPLANE = K * 32 + 16;       // + 16: the two half-warps of a P2 block (planes q, q + 1) use different banks
constexpr int PSTR = 32;

template <typename T> struct Geo {
    static constexpr int HV = 16 / (int)sizeof(T);       // halo elements per side (one 16-byte segment)
    static constexpr int TRE = RW + 2 * HV;              // elements per tile row
    static constexpr int NW = RW * (int)sizeof(T) / 4;   // 32-bit words of the owned part of a row
    static constexpr int ROWB = TRE * (int)sizeof(T);    // bytes per tile row (a multiple of 16)
    static constexpr int ROWV = ROWB / 16;
    static constexpr int TILE_BYTES = TP * TM * ROWB;
};

template <typename T> constexpr size_t table_bytes() {          // one set of per-brick tables: label keys / values, pair values / keys
    return LT_SLOTS * 4 + LT_SLOTS * LT_FIELDS * 4 + PT_SLOTS * PT_WORDS * 4 + PT_SLOTS * sizeof(typename Vox<T>::PKey);
}
template <typename T> constexpr size_t smem_bytes() {
    return (size_t)Geo<T>::TILE_BYTES + (size_t)TP * MPLANE * 4 + (size_t)3 * TP * PSTR * 4 + 2 * table_bytes<T>() +
           96 * 4 + 16 * 8 + 128;
}

// ---- packed compares ---------------------------------------------------------------------------------------------------
#ifdef TA_EMU_TMA
inline uint32_t min_u16x2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
#else
__device__ __forceinline__ uint32_t min_u16x2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
#endif

#ifdef TA_EMU_TMA
inline uint4 mk_ld128(const void* p) { return *reinterpret_cast<const uint4*>(p); }
#else
__device__ __forceinline__ uint4 mk_ld128(const void* p) { return ld_stream_128(p); }
#endif

#ifdef TA_EMU_TMA
inline int mk_tid() { return (int)threadIdx.x; }
#else
__device__ __forceinline__ int mk_tid() {
    int t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    return t;
}
#endif

// bit j set iff voxel j of the row opens a run (differs from voxel j - 1); bit 0 is always set
#ifdef TA_EMU_TMA
inline uint32_t mk_funnel_l16(uint32_t lo, uint32_t hi) { return (hi << 16) | (lo >> 16); }
inline uint32_t mk_umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((u64)a * (u64)b) >> 32); }
#else
__device__ __forceinline__ uint32_t mk_funnel_l16(uint32_t lo, uint32_t hi) { return __funnelshift_l(lo, hi, 16); }
__device__ __forceinline__ uint32_t mk_umulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
#endif
template <typename T> struct RunStarts;
template <> struct RunStarts<uint16_t> {
    // word j holds voxels 2j, 2j + 1; the funnel shift puts their left neighbours 2j - 1, 2j in the same halves
    static __device__ __forceinline__ uint32_t of(const uint32_t (&w)[16]) {
        uint32_t a0 = 0u, a1 = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) a0 += min_u16x2(w[j] ^ mk_funnel_l16(w[j ? j - 1 : 0], w[j]), 0x00010001u) << (2 * j);
#pragma unroll
        for (int j = 0; j < 8; ++j) a1 += min_u16x2(w[8 + j] ^ mk_funnel_l16(w[7 + j], w[8 + j]), 0x00010001u) << (2 * j);
        return ((a0 & 0x5555u) | ((a0 >> 15) & 0xAAAAu)) | (((a1 & 0x5555u) | ((a1 >> 15) & 0xAAAAu)) << 16) | 1u;
    }
    static __device__ __forceinline__ uint32_t first(const uint32_t (&w)[16]) { return w[0] & 0xFFFFu; }
    static __device__ __forceinline__ uint32_t last(const uint32_t (&w)[16]) { return w[15] >> 16; }
};
template <> struct RunStarts<uint32_t> {
    static __device__ __forceinline__ uint32_t of(const uint32_t (&w)[32]) {
        uint32_t st = 1u;
#pragma unroll
        for (int j = 1; j < 32; ++j) if (w[j] != w[j - 1]) st |= 1u << j;
        return st;
    }
    static __device__ __forceinline__ uint32_t first(const uint32_t (&w)[32]) { return w[0]; }
    static __device__ __forceinline__ uint32_t last(const uint32_t (&w)[32]) { return w[31]; }
};

// sum of the bit positions and of their squares
__device__ __forceinline__ void bit_moments(uint32_t M, uint32_t n, uint32_t& sf, uint32_t& sff) {
    const uint32_t lowbit = M & (0u - M);
    if (((M + lowbit) & M) == 0u) {                      // one run of bits [lo, lo + n)
        const uint32_t lo = 32u - (uint32_t)__clz(M) - n; // M != 0 here; one FLO (ffs would be BREV + FLO)
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

// ---- out-of-line pieces: the kernel's hot loop has to fit the 32 KB instruction cache of an SM (ncu: `no_instruction`
// stalls grew from 0.3 to 1.9 cycles per issue when the loop body passed it), so everything rare is a call ----------------

// thread 0: closed-form moments of a brick whose tile is one label; consecutive ones of the same label are merged
// brick-local sums of an a x b x c box of one label at the brick's origin
__device__ __forceinline__ void one_label_sums(uint32_t a, uint32_t b, uint32_t c, uint32_t v[LT_FIELDS]) {
    const uint32_t ta_ = a * (a - 1) / 2, tb = b * (b - 1) / 2, tc = c * (c - 1) / 2;
    const uint32_t qa = (a - 1) * a * (2 * a - 1) / 6, qb = (b - 1) * b * (2 * b - 1) / 6, qc = (c - 1) * c * (2 * c - 1) / 6;
    v[0] = a * b * c; v[1] = b * c * ta_; v[2] = a * c * tb; v[3] = a * b * tc;
    v[4] = b * c * qa; v[5] = c * ta_ * tb; v[6] = b * ta_ * tc;
    v[7] = a * c * qb; v[8] = a * tb * tc; v[9] = a * b * qc;
    v[10] = 0; v[11] = 0; v[12] = 0; v[13] = a - 1; v[14] = b - 1; v[15] = c - 1;
}
__device__ __noinline__ void uniform_brick(Carry* carry, LabelTable lt, uint32_t* status, uint32_t label, uint32_t a, uint32_t b,
                                           uint32_t c, u64 gF0, u64 gM0, u64 gS0) {
    uint32_t v[LT_FIELDS];
    one_label_sums(a, b, c, v);
    u64 g[10];
    int bmn[3], bmx[3];
    local_to_global(v, gF0, gM0, gS0, g, bmn, bmx);
    if (carry->valid && carry->label != label) {
        global_apply(lt, status, carry->label, carry->g, carry->bmn, carry->bmx);
        carry->valid = 0u;
    }
    if (!carry->valid) {
        carry->valid = 1u; carry->label = label;
        for (int i = 0; i < 10; ++i) carry->g[i] = g[i];
        for (int i = 0; i < 3; ++i) { carry->bmn[i] = bmn[i]; carry->bmx[i] = bmx[i]; }
    } else {
        for (int i = 0; i < 10; ++i) carry->g[i] += g[i];
        for (int i = 0; i < 3; ++i) { carry->bmn[i] = min(carry->bmn[i], bmn[i]); carry->bmx[i] = max(carry->bmx[i], bmx[i]); }
    }
}

// every thread of the CTA: elements of a box copy outside the buffer arrived as zeros; the tile wants replicated edge voxels
template <typename T>
__device__ __noinline__ void replicate_edges(T* tile, int tid, int F0, int M0, int S0, int nf, int nm, int ns) {
    typedef Geo<T> G;
    constexpr int HV = G::HV, TRE = G::TRE, ROWV = G::ROWV;
    uint4* tilev = reinterpret_cast<uint4*>(tile);
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

// one pair's packed increments into the per-brick table (both call sites of P2 share this copy)
template <typename T>
__device__ __noinline__ void pair_add_call(BrickShared<T> sh, PairTable pt, typename Vox<T>::PKey key, uint32_t i0, uint32_t i1,
                                           uint32_t i2, uint32_t i3) {
    const uint32_t inc[PT_WORDS] = {i0, i1, i2, i3};
    pair_add_packed<T>(sh, pt, key, inc);
}

// G: more than K labels in the brick (noise; cells under ~20 voxels across: C1): every owned voxel on its own, straight
// from the tile.  One thread per owned row: moments per RUN of a label (closed forms, one table update per run), the 18
// neighbours of a voxel in registers (offsets are compile-time constants), a voxel whose neighbours all carry its own label
// is done after 18 compares.
template <typename T>
__device__ __noinline__ void brick_per_voxel(const T* tile, BrickShared<T> sh, LabelTable lt, PairTable pt, int tid, int nown, int F0,
                                             int M0, int nf, int nm, u64 gF0, u64 gM0, u64 gS0, bool do_mom, bool do_p6, bool do_w18) {
    typedef Geo<T> G;
    constexpr int HV = G::HV, TRE = G::TRE;
    constexpr int PLANEE = TM * TRE;
    const int nrows = OM * nown, nfv = min(RW, nf - F0);
    for (int row = tid; row < nrows; row += NTHREADS) {
        const int m = row % OM, s = row / OM;
        if (M0 + m >= nm) continue;
        const T* rowp = tile + ((s + 1) * TM + (m + 1)) * TRE + HV;
        const uint32_t um = (uint32_t)m, us = (uint32_t)s;
        uint32_t run_label = rowp[0];
        int run_f0 = 0;
        for (int f = 0; f <= nfv; ++f) {
            const uint32_t a = f < nfv ? (uint32_t)rowp[f] : ~run_label;        // f == nfv: closes the last run
            if (a != run_label) {
                if (do_mom) {
                    const uint32_t lo = (uint32_t)run_f0, n = (uint32_t)(f - run_f0), t = n * (n - 1u);
                    const uint32_t sf = n * lo + (t >> 1), sff = n * lo * lo + lo * t + (t * (2u * n - 1u)) / 6u;
                    uint32_t v[LT_FIELDS] = {n, sf, n * um, n * us, sff, sf * um, sf * us, n * um * um, n * um * us, n * us * us,
                                             lo, um, us, (uint32_t)f - 1u, um, us};
                    label_add<T>(sh, lt, pt.status, run_label, v, gF0, gM0, gS0);
                }
                run_label = a; run_f0 = f;
            }
            if (f == nfv) break;
            const T* qv = rowp + f;
            if (do_p6) {
#pragma unroll 1
                for (int d = 0; d < 3; ++d) {
                    const uint32_t n = qv[d == 0 ? 1 : (d == 1 ? TRE : PLANEE)];
                    if (n != a) pair_add<T>(sh, pt, a, n, 2 * d + (a < n ? 0 : 1), 1u);
                }
            }
            if (do_w18) {
                uint32_t nb[18];
                uint32_t differ = 0u;
                {
                    int c = 0;
#pragma unroll
                    for (int ds = -1; ds <= 1; ++ds)
#pragma unroll
                        for (int dm = -1; dm <= 1; ++dm)
#pragma unroll
                            for (int df = -1; df <= 1; ++df) {
                                const int l1 = (ds ? 1 : 0) + (dm ? 1 : 0) + (df ? 1 : 0);
                                if (l1 >= 1 && l1 <= 2) { nb[c] = qv[ds * PLANEE + dm * TRE + df]; differ |= nb[c] ^ a; ++c; }
                            }
                }
                if (differ) {
                    uint32_t fresh = 0u;                               // bit k: neighbour k carries a label not seen at 0 .. k - 1
#pragma unroll
                    for (int k = 0; k < 18; ++k) {
                        bool fr = nb[k] != a;
#pragma unroll
                        for (int kk = 0; kk < k; ++kk) fr = fr && (nb[kk] != nb[k]);
                        fresh |= fr ? (1u << k) : 0u;
                    }
#pragma unroll 1
                    for (; fresh; fresh &= fresh - 1u) pair_add<T>(sh, pt, a, nb[__ffs(fresh) - 1], 6, 1u);
                }
            }
        }
    }
}

template <typename T, int FLAGS = -1>                      // FLAGS >= 0: the pass flags at compile time (the full pass of the product)
__global__ void __launch_bounds__(NTHREADS, (sizeof(T) == 2 ? 3 : 2))
mask_kernel(ScanParams P, LabelTable lt, PairTable pt, const __grid_constant__ CUtensorMap tmap) {
    typedef typename Vox<T>::PKey PKey;
    typedef Geo<T> G;
    constexpr int HV = G::HV, TRE = G::TRE, NW = G::NW;
    constexpr int PLANEE = TM * TRE;
    constexpr uint32_t FULL = 0xffffffffu;
    static_assert(K <= 32 && 5 * KB <= 64, "lane l looks at blab[l]; the slots of a round are packed 5 bits each");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    uint32_t* masks = reinterpret_cast<uint32_t*>(smem_raw + G::TILE_BYTES);            // [TP][MPLANE]: own bits of (slot, row)
    uint32_t* pres = masks + TP * MPLANE;                                                // [TP][32]: slots with a bit in the row
    uint32_t* hwl = pres + TP * PSTR;                                                      // [TP][32]: slots in the left halo voxel of the row
    uint32_t* hwr = hwl + TP * PSTR;                                                       // [TP][32]: slots in the right halo voxel
    constexpr int TBLW = LT_SLOTS + LT_SLOTS * LT_FIELDS + PT_SLOTS * PT_WORDS + PT_SLOTS * (int)sizeof(PKey) / 4;   // words of one set (table_bytes)
    constexpr int NSETS = 2;                                       // ping-pong: see the deferred flush below
    uint32_t* const tables = hwr + TP * PSTR;
    auto table_set = [&](unsigned which) {
        BrickShared<T> x{};
        x.lt_key = tables + which * TBLW;
        x.lt_val = x.lt_key + LT_SLOTS;
        x.pt_val = x.lt_val + LT_SLOTS * LT_FIELDS;
        x.pt_key = reinterpret_cast<PKey*>(x.pt_val + PT_SLOTS * PT_WORDS);
        return x;
    };
    // ctr: [0..1] brick index ping-pong, [2..3] overflow flag ping-pong, [4..5] mbarrier, [6] moment queue, [8..10] / [12..14] brick origin
    // ping-pong, [16 .. 16 + 2 K) the brick's label list, ping-pong
    unsigned int* ctr = reinterpret_cast<unsigned int*>(tables + NSETS * TBLW);
    Carry* carry = reinterpret_cast<Carry*>(ctr + 16 + 2 * K + (2 * K) % 2 + 16);

    // read once and kept: the compiler otherwise re-reads the special register (S2R, a slow-pipe instruction) inside the loops
    const int tid = mk_tid();
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const T* vol = reinterpret_cast<const T*>(P.vol);
    // the queue: every brick of the launch, or the work list the pre-pass left (ta_prepass.cuh)
    const unsigned int total = P.work_list ? *P.work_count : (unsigned int)P.nbf * P.nbm * P.nbs;
    const uint32_t pflags = FLAGS >= 0 ? (uint32_t)FLAGS : P.flags;
    const bool do_mom = pflags & 1u, do_p6 = pflags & 2u, do_w18 = pflags & 4u;
    const bool do_pairs = do_p6 || do_w18;
    const int nf = (int)P.nf, nm = (int)P.nm, ns = (int)P.ns;

    for (unsigned which = 0; which < (unsigned)NSETS; ++which) {
        const BrickShared<T> x = table_set(which);
        for (int i = tid; i < LT_SLOTS; i += NTHREADS) x.lt_key[i] = TA_EMPTY32;
        for (int i = tid; i < LT_SLOTS * LT_FIELDS; i += NTHREADS) {
            const int f = i % LT_FIELDS;
            x.lt_val[i] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
        }
        for (int i = tid; i < PT_SLOTS; i += NTHREADS) x.pt_key[i] = Vox<T>::PEMPTY;
        for (int i = tid; i < PT_SLOTS * PT_WORDS; i += NTHREADS) x.pt_val[i] = 0u;
    }
    // F: one set of per-brick tables -> the global tables (brick-local u32 sums -> shifted u64 global REDs; pair slots ->
    // global hash), by threads first .. first + count - 1 of the CTA; leaves the set empty
    auto flush_tables = [&](const BrickShared<T>& x, u64 oF, u64 oM, u64 oS, int first, int count) {
        for (int i = tid - first; i < LT_SLOTS; i += count) {
            const uint32_t L = x.lt_key[i];
            if (L == TA_EMPTY32) continue;
            uint32_t* d = &x.lt_val[i * LT_FIELDS];
            label_to_global(lt, pt.status, L, d, oF, oM, oS);
#pragma unroll
            for (int f = 0; f < LT_FIELDS; ++f) d[f] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
            x.lt_key[i] = TA_EMPTY32;
        }
        for (int i = tid - first; i < PT_SLOTS; i += count) {
            const PKey key = x.pt_key[i];
            if (key == Vox<T>::PEMPTY) continue;
            uint32_t* d = &x.pt_val[i * PT_WORDS];
            const int slot = ta_pair_slot(pt, Vox<T>::key64(key));
#pragma unroll
            for (int idx = 0; idx < 7; ++idx) {
                const uint32_t n = (d[idx >> 1] >> ((idx & 1) * 16)) & 0xFFFFu;
                if (n && slot >= 0) atomicAdd(&pt.vals[(size_t)slot * TA_PAIR_STRIDE + (idx == 0 ? 6 : idx - 1)], n);
            }
#pragma unroll
            for (int w2 = 0; w2 < PT_WORDS; ++w2) d[w2] = 0u;
            x.pt_key[i] = Vox<T>::PEMPTY;
        }
    };
    bool pending = false;                                  // deferred flush: the other set holds the previous brick's sums
    u64 pF0 = 0, pM0 = 0, pS0 = 0;                         // ... taken at this origin
    unsigned tset = 0u;                                    // the set the next brick that is not one label will use; the pending one is the other
    if (tid < 2 * K) ctr[16 + tid] = TA_EMPTY32;
    for (int i = tid; i < TP * MPLANE; i += NTHREADS) masks[i] = 0u;       // invariant: every mask is zero when a brick's P1 starts

    uint64_t* tma_bar = reinterpret_cast<uint64_t*>(ctr + 4);
    uint32_t tma_parity = 0u;
    const bool use_tma = P.use_tma && ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u) == 0u;
    if (tid == 0) {
        if (use_tma) {
            mbar_init(tma_bar, 1u);
            TA_PTX("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        carry->valid = 0u;
        ctr[2] = 0u; ctr[3] = 0u; ctr[6] = 0u;
    }

    // thread 0: origin of a brick into its ping-pong slot (the divisions are made once per brick, not once per thread)
    auto set_origin = [&](unsigned int brick, unsigned which) {
        const unsigned bf = brick % (unsigned)P.nbf, rest = brick / (unsigned)P.nbf;
        const unsigned bm = rest % (unsigned)P.nbm, bs = rest / (unsigned)P.nbm;
        ctr[8 + 4 * which] = bf * RW; ctr[9 + 4 * which] = bm * OM; ctr[10 + 4 * which] = (unsigned)((int)P.own_lo + (int)bs * ZB);
    };
    auto issue_box = [&](unsigned which) {          // thread 0, after a barrier that ended every read of the tile
        TA_PTX("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(tma_bar, (uint32_t)G::TILE_BYTES);
        tma_load_box_3d(tile, &tmap, tma_bar, (int)ctr[8 + 4 * which] - HV, (int)ctr[9 + 4 * which] - 1, (int)ctr[10 + 4 * which] - 1);
    };

    __syncthreads();
    if (tid == 0) {
        const unsigned int b0 = atomicAdd(P.brick_counter, 1u);
        ctr[0] = b0;
        if (b0 < total) {
            set_origin(P.work_list ? P.work_list[b0] : b0, 0u);
            if (use_tma) issue_box(0u);
        }
        const unsigned int b1 = atomicAdd(P.brick_counter, 1u);
        ctr[1] = b1;
        if (b1 < total) set_origin(P.work_list ? P.work_list[b1] : b1, 1u);
    }
    __syncthreads();

#ifdef TA_WITH_PHASE_TIMING
    // phase clocks of lane 0 of warp 0 ([0..6]) and of the last warp ([8..14]): tile wait, P1, barrier, P2, barrier, flush;
    // [6]: one-label bricks, [14]: cycles the last warp spent in them; [7], [15]: bricks
    u64 tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u64 t_last = (u64)clock64();
    const bool clocked = P.phase_cycles && lane == 0 && (warp == 0 || warp == TP - 1);
#define MK_TICK(k) if (clocked) { const u64 now_ = (u64)clock64(); tk[k] += now_ - t_last; t_last = now_; }
#else
#define MK_TICK(k)
#endif
    for (unsigned iter = 0;; ++iter) {
        const unsigned cur = iter & 1u, nxt = cur ^ 1u;
#ifdef TA_WITH_PHASE_TIMING
        const u64 t_iter0 = (u64)clock64();
#endif
        const unsigned int brick = ctr[cur];
        if (brick >= total) break;
        uint32_t* blab = ctr + 16 + K * cur;
        const int F0 = (int)ctr[8 + 4 * cur], M0 = (int)ctr[9 + 4 * cur], S0 = (int)ctr[10 + 4 * cur];
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
                    P.diag[3] = ((u64)tma_parity << 32) | (u64)ctr[nxt];
                    P.diag[4] = *reinterpret_cast<volatile u64*>(tma_bar);
                    P.diag[5] = ((u64)(uint32_t)(F0 - HV) << 32) | ((u64)(uint32_t)(M0 - 1) << 16) | (u64)(uint32_t)(S0 - 1);
                    __threadfence_system();
                }
                return;
            }
            tma_parity ^= 1u;
            // elements outside the buffer arrived as zeros; the tile wants replicated edge voxels
            const bool edge = (F0 == 0) | (F0 + RW + HV > nf) | (M0 == 0) | (M0 + TM - 1 > nm) | (S0 < 1) | (S0 + ZB + 1 > ns);
            if (edge) replicate_edges<T>(tile, tid, F0, M0, S0, nf, nm, ns);
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

        MK_TICK(0);
#ifdef TA_WITH_PHASE_TIMING
        // measurement aid of the timing build (flag 0x800): staging only -- every tile is fetched and dropped, which times the
        // box copies on their own, one in flight per CTA
        if (FLAGS < 0 && (pflags & 0x800u)) {
            __syncthreads();
            if (tid == 0 && use_tma && ctr[nxt] < total) issue_box(nxt);
            continue;
        }
#endif
        // ---- P1: block of 16 rows x 2 planes: runs of the lane's row -> brick slots, row masks ---------------------------
        const uint32_t ref_label = tile[HV];               // plane 0, row 0, first owned column
        const int q1 = 2 * (warp >> 1) + (lane >> 4);      // tile plane of the lane
        const int r1 = 16 * (warp & 1) + (lane & 15);      // tile row of the lane
        const bool needed = 2 * (warp >> 1) <= nown + 1;   // the block has a plane somebody needs (uniform per warp)
        bool one_label = true;                             // this block: not needed, or all ref_label
        if (needed) {
            const T* row = tile + (q1 * TM + r1) * TRE;
            uint32_t w[NW];
            {
                const uint4* rv = reinterpret_cast<const uint4*>(row + HV);
#pragma unroll
                for (int x = 0; x < NW / 4; ++x) {
                    const uint4 v = rv[x];
                    w[4 * x] = v.x; w[4 * x + 1] = v.y; w[4 * x + 2] = v.z; w[4 * x + 3] = v.w;
                }
            }
            const uint32_t hl = row[HV - 1], hr = row[HV + RW];
            uint32_t* mrow = masks + q1 * MPLANE + r1;
            // slot of a label = its place in the brick's open-addressing list (K entries, CAS insert; every lane that
            // looks for the same label walks the same places, so nobody waits for anybody)
            auto find_slot = [&](uint32_t L) -> int {
                uint32_t h = mk_umulhi(L * 0x9E3779B1u, (uint32_t)K);
#pragma unroll 1
                for (int probe = 0; probe < K; ++probe) {
                    uint32_t e = *((volatile uint32_t*)&blab[h]);
                    if (e == TA_EMPTY32) e = atomicCAS(&blab[h], TA_EMPTY32, L);
                    if (e == L || e == TA_EMPTY32) return (int)h;
                    h = (h + 1u == (uint32_t)K) ? 0u : h + 1u;
                }
                return -1;
            };
            // the runs of the row, last one first (the highest set bit costs one FLO): label from the tile, bits [j0, j1)
            const uint32_t v0 = RunStarts<T>::first(w), v31 = RunStarts<T>::last(w);
            uint32_t rem = RunStarts<T>::of(w);
            one_label = (rem == 1u) && (v0 == ref_label) && (hl == v0) && (hr == v0);
            uint32_t mypres = 0u;
            int j1 = 32, slot_first = 0, slot_last = -1;
            bool ok = true;
            while (rem) {
                const int j0 = 31 - __clz(rem);
                rem ^= 1u << j0;
                const int slot = find_slot((uint32_t)row[HV + j0]);
                if (slot < 0) { ok = false; break; }
                mrow[slot * 32] |= (FULL >> (32 - j1)) & (FULL << j0);
                mypres |= 1u << slot;
                if (slot_last < 0) slot_last = slot;
                slot_first = slot;
                j1 = j0;
            }
            int sl = slot_first, sr = slot_last;
            if (ok && hl != v0) sl = find_slot(hl);
            if (ok && hr != v31) sr = find_slot(hr);
            if (!ok || sl < 0 || sr < 0) { ctr[2 + cur] = 1u; sl = 0; sr = 0; }
            pres[q1 * PSTR + r1] = mypres | (1u << sl) | (1u << sr);
            hwl[q1 * PSTR + r1] = 1u << sl; hwr[q1 * PSTR + r1] = 1u << sr;
        }
        MK_TICK(1);
        const bool uniform = __syncthreads_and(one_label) != 0;
        MK_TICK(2);
        const bool overflow = ctr[2 + cur] != 0u;
        const unsigned int next_brick = ctr[nxt];
        bool box_issued = false;
        if (use_tma && !overflow && next_brick < total) {
            if (tid == 0) issue_box(nxt);
            box_issued = true;
        }
        // The queue index of the brick AFTER next goes into the slot this brick has just stopped needing (every thread read
        // its origin before the barrier).  One lane of the last warp asks for it here, under P2: at the top of the loop the
        // round trip of the global atomic made warp 0 late for P1 and the other warps waited for it at the barrier.
        if (tid == (TP - 1) * 32) {
            const unsigned int nb2 = atomicAdd(P.brick_counter, 1u);
            ctr[cur] = nb2;
            if (nb2 < total) set_origin(P.work_list ? P.work_list[nb2] : nb2, cur);
        }

        // the warps without a P2 block empty the tables of the last brick that used them while the others work on this brick's
        // (a one-label brick uses no tables: nothing to hide the flush behind, so it waits for the next brick that does)
        const BrickShared<T> sh = table_set(tset);
        if (pending && !uniform && warp >= NW2) flush_tables(table_set(tset ^ 1u), pF0, pM0, pS0, NW2 * 32, NTHREADS - NW2 * 32);
        if (uniform) {
            // the whole tile (brick + halo) is one label: closed-form moments, no pairs
            if (tid == 0 && do_mom)
                uniform_brick(carry, lt, pt.status, ref_label, (uint32_t)fvalid_n, (uint32_t)min(OM, nm - M0), (uint32_t)nown, gF0, gM0, gS0);
        } else if (!overflow) {
            // ---- P2: blocks of 15 owned rows x 2 owned planes.  Warp w < 8 does the PAIRS of block w; the MOMENTS of the eight
            // blocks are items of a queue (ctr[6]) that every warp draws from when it has nothing else to do: warps 8 and 9
            // after their flush, the others after their pairs (blocks differ in their number of labels, so the warps do
            // not finish together: phase clocks had the slowest at 1.4 times the mean).
            int blk = warp;
            bool pairs_turn = warp < NW2;
            for (;;) {
                if (!pairs_turn) {
                    if (!do_mom) break;
                    unsigned int it = 0u;
                    if (lane == 0) it = atomicAdd(&ctr[6], 1u);
                    blk = (int)__shfl_sync(FULL, it, 0);
                    if (blk >= NW2) break;
                }
                if (1 + 2 * (blk >> 1) <= nown) {
                    const int wm = blk & 1, s0 = 2 * (blk >> 1);                // s0: brick-local plane of the lower half-warp
                    const int lm = lane & 15, ls = lane >> 4;
                    const bool own_row = (lm < 15) && (s0 + ls < nown) && (M0 + 15 * wm + lm < nm);
                    const int r = own_row ? 1 + 15 * wm + lm : 1;               // idle lanes look at a harmless row
                    const int q = own_row ? 1 + s0 + ls : 1;
                    const uint32_t fm = own_row ? fvalid : 0u;
                    const uint32_t ml = (uint32_t)(15 * wm + lm);               // brick-local row
                    const uint32_t* mbase = masks + q * MPLANE + r;             // + slot * 32; rows +-1, planes +- MPLANE
                    const uint32_t* pbase = pres + q * PSTR + r;
                    const uint32_t PC = __reduce_or_sync(FULL, own_row ? pbase[0] : 0u);  // labels in the block
                    if (pairs_turn) {
                        const uint32_t* hlb = hwl + q * PSTR + r;
                        const uint32_t* hrb = hwr + q * PSTR + r;
                        const uint32_t hrc1 = hrb[0];
                        // halo slots of the five rows with an f-shifted neighbour
                        const uint32_t hlall = hlb[-1] | hlb[0] | hlb[1] | hlb[-PSTR] | hlb[PSTR], hrall = hrb[-1] | hrc1 | hrb[1] | hrb[-PSTR] | hrb[PSTR];
                        uint32_t pn = pbase[-PSTR - 1] | pbase[-PSTR] | pbase[-PSTR + 1] | pbase[-1] | pbase[0] | pbase[1] | pbase[PSTR - 1] | pbase[PSTR] | pbase[PSTR + 1];
                        const uint32_t PU = __reduce_or_sync(FULL, own_row ? pn : 0u);       // labels around the block
                        if (do_pairs) {
                            int nres = 0;
                            uint32_t res_a = 0, res_b = 0, res_1 = 0, res_2 = 0;
                            auto flush_results = [&]() {
                                if (lane < nres) {
                                    const uint32_t a = blab[res_a], b = blab[res_b];
                                    const bool lo = a < b;
                                    const uint32_t w18 = res_1 & 0xFFFFu, ff = res_1 >> 16, fmm = res_2 & 0xFFFFu, fss = res_2 >> 16;
                                    uint32_t inc[PT_WORDS];
                                    inc[0] = w18 | (lo ? ff << 16 : 0u);
                                    inc[1] = (lo ? 0u : ff) | (lo ? fmm << 16 : 0u);
                                    inc[2] = (lo ? 0u : fmm) | (lo ? fss << 16 : 0u);
                                    inc[3] = lo ? 0u : fss;
                                    pair_add_call<T>(sh, pt, Vox<T>::key(a, b), inc[0], inc[1], inc[2], inc[3]);
                                }
                                nres = 0;
                            };
                            // rounds of up to KB labels of the block: their own-row masks in registers, their slots packed 5 bits each
                            for (uint32_t pcr = PC; pcr;) {
                                uint32_t Ma[KB];
                                u64 slots = 0ull;
#pragma unroll
                                for (int j = 0; j < KB; ++j) {
                                    const int i = pcr ? __ffs(pcr) - 1 : 0;
                                    Ma[j] = pcr ? (mbase[i * 32] & fm) : 0u;
                                    slots |= (u64)i << (5 * j);
                                    pcr &= pcr - 1u;
                                }
                                for (uint32_t rest = PU; rest;) {
                                    const int b = 31 - __clz(rest);
                                    rest ^= 1u << b;
                                    const uint32_t* qb = mbase + b * 32;
                                    const uint32_t c0 = qb[-1], c1 = qb[0], c2 = qb[1];
                                    const uint32_t d0 = qb[-MPLANE - 1], d1 = qb[-MPLANE], d2 = qb[-MPLANE + 1];
                                    const uint32_t u0 = qb[MPLANE - 1], u1 = qb[MPLANE], u2 = qb[MPLANE + 1];
                                    const uint32_t Y = c0 | c1 | c2 | d0 | d1 | d2 | u0 | u1 | u2;
                                    const uint32_t Pm = c0 | c1 | c2 | d1 | u1;
                                    const uint32_t PH = ((hlall >> b) & 1u) | (((hrall >> b) & 1u) << 31);   // b in a left / right halo voxel of the five rows
                                    const uint32_t Dn = (Y | (Pm << 1) | (Pm >> 1) | PH) & fm & ~c1;          // not-b voxels with a b in their N18
                                    // labels of the round that meet Dn somewhere in the warp (b itself cannot: Dn excludes its voxels)
                                    uint32_t ts = 0u;
#pragma unroll
                                    for (int j = 0; j < KB; ++j) ts |= (Ma[j] & Dn) ? (1u << j) : 0u;
                                    ts = __reduce_or_sync(FULL, ts);
                                    if (!ts) continue;
                                    const uint32_t Bf = (c1 >> 1) | (((hrc1 >> b) & 1u) << 31);                // b is the +f neighbour
#pragma unroll
                                    for (int j = 0; j < KB; ++j) {
                                        if (!((ts >> j) & 1u)) continue;
                                        const uint32_t Mo = Ma[j];
                                        uint32_t cnt1 = do_w18 ? (uint32_t)__popc(Mo & Dn) : 0u, cnt2 = 0u;
                                        if (do_p6) {
                                            cnt1 |= (uint32_t)__popc(Mo & Bf) << 16;
                                            cnt2 = (uint32_t)__popc(Mo & c2) | ((uint32_t)__popc(Mo & u1) << 16);
                                        }
                                        const uint32_t r1 = __reduce_add_sync(FULL, cnt1);
                                        const uint32_t r2 = __reduce_add_sync(FULL, cnt2);
                                        const bool mine = lane == nres;
                                        res_a = mine ? (uint32_t)((slots >> (5 * j)) & 31ull) : res_a;
                                        res_b = mine ? (uint32_t)b : res_b;
                                        res_1 = mine ? r1 : res_1;
                                        res_2 = mine ? r2 : res_2;
                                        if (++nres == 32) flush_results();
                                    }
                                }
                            }
                            flush_results();
                        }
                    } else {
                        uint32_t k1 = 0, k2 = 0, k3 = 0, k4 = 0, k5 = 0, k6 = 0, kx = 0, ky = 0, ks = 0;
                        const uint32_t up = ls ? 0xFFFFFFFFu : 0u;
                        int idx = 0;
                        for (uint32_t rest = PC; rest;) {                     // the slots are places in a hashed list: not dense
                            const int i = 31 - __clz(rest);
                            rest ^= 1u << i;
                            const uint32_t M = mbase[i * 32] & fm;
                            const uint32_t n = (uint32_t)__popc(M);
                            uint32_t sf = 0u, sff = 0u;
                            if (M) bit_moments(M, n, sf, sff);
                            const uint32_t nml = n * ml;
                            const uint32_t r1 = __reduce_add_sync(FULL, n | (sf << 10));                 // n < 2^10, sum f < 2^15
                            const uint32_t r2 = __reduce_add_sync(FULL, nml | ((sf & up) << 15));        // sum n m < 2^15, upper sum f < 2^14
                            const uint32_t r3 = __reduce_add_sync(FULL, sff);
                            const uint32_t r4 = __reduce_add_sync(FULL, nml * ml);
                            const uint32_t r5 = __reduce_add_sync(FULL, sf * ml);
                            const uint32_t r6 = __reduce_add_sync(FULL, (n & up) | ((nml & up) << 10));  // upper n < 2^9, upper sum n m < 2^14
                            const uint32_t rx = __reduce_or_sync(FULL, M);
                            const uint32_t ry = __ballot_sync(FULL, M != 0u);
                            if (lane == idx) { k1 = r1; k2 = r2; k3 = r3; k4 = r4; k5 = r5; k6 = r6; kx = rx; ky = ry; ks = (uint32_t)i; }
                            ++idx;
                        }
                        if (lane < idx && kx) {
                            const uint32_t n = k1 & 0x3FFu, sf = k1 >> 10, nm_ = k2 & 0x7FFFu, sf1 = k2 >> 15;
                            const uint32_t n1 = k6 & 0x3FFu, nm1 = k6 >> 10;
                            const uint32_t rows = (ky | (ky >> 16)) & 0x7FFFu;
                            const uint32_t u0 = (uint32_t)s0;
                            uint32_t v[LT_FIELDS];
                            v[0] = n; v[1] = sf; v[2] = nm_; v[3] = u0 * n + n1;
                            v[4] = k3; v[5] = k5; v[6] = u0 * sf + sf1;
                            v[7] = k4; v[8] = u0 * nm_ + nm1; v[9] = u0 * u0 * n + (2u * u0 + 1u) * n1;
                            v[10] = (uint32_t)__ffs(kx) - 1u; v[11] = 15u * wm + (uint32_t)__ffs(rows) - 1u; v[12] = (ky & 0xFFFFu) ? u0 : u0 + 1u;
                            v[13] = 31u - (uint32_t)__clz(kx); v[14] = 15u * wm + 31u - (uint32_t)__clz(rows); v[15] = (ky >> 16) ? u0 + 1u : u0;
                            label_add<T>(sh, lt, pt.status, blab[ks], v, gF0, gM0, gS0);
                        }
                    }
                }
                pairs_turn = false;
            }
        } else {
            // ---- G: more than K labels in the brick: every owned voxel on its own, straight from the tile -------------------
            brick_per_voxel<T>(tile, sh, lt, pt, tid, nown, F0, M0, nf, nm, gF0, gM0, gS0, do_mom, do_p6, do_w18);
        }
        MK_TICK(3);
        // (skipping this barrier for a one-label brick costs registers: ptxas spills 188 bytes instead of 84 around it)
        __syncthreads();
        MK_TICK(4);

        // ---- F: flush the per-brick tables ---------------------------------------------------------------------------------
        if (!uniform) { pending = true; tset ^= 1u; pF0 = gF0; pM0 = gM0; pS0 = gS0; }   // left to a later iteration (or to the end of the kernel)
        // No barrier here: the next brick's P1 touches neither these tables nor this brick's label list (the list and the
        // overflow flag are ping-pong), and the barrier after it comes before anything that does.
        {
            // back to all-zero masks: every thread clears the slots its own P1 row may have written (the same thread writes
            // them again in the next P1, so program order is enough)
            const int q = 2 * (warp >> 1) + (lane >> 4), r = 16 * (warp & 1) + (lane & 15);   // as in P1
            if (2 * (warp >> 1) <= nown + 1) {
                uint32_t* mrow = masks + q * MPLANE + r;
                for (uint32_t mine = pres[q * PSTR + r]; mine;) {
                    const int sl = 31 - __clz(mine);
                    mine ^= 1u << sl;
                    mrow[sl * 32] = 0u;
                }
            }
        }
        if (tid < K) blab[tid] = TA_EMPTY32;
        if (tid == 0) {
            ctr[2 + cur] = 0u; ctr[6] = 0u;
            if (use_tma && !box_issued && next_brick < total) issue_box(nxt);   // after G: the tile was in use until the barrier
        }
        MK_TICK(5);
#ifdef TA_WITH_PHASE_TIMING
        if (clocked) { tk[7] += 1; if (uniform) tk[6] += (warp == 0) ? 1ull : (u64)clock64() - t_iter0; }   // [6]: count (first warp), cycles (last warp)
#endif
    }
#ifdef TA_WITH_PHASE_TIMING
    if (clocked) for (int k = 0; k < 8; ++k) atomicAdd(&P.phase_cycles[(warp == 0 ? 0 : 8) + k], tk[k]);
#endif
#undef MK_TICK
    // the last brick's tables (`pending` is CTA-uniform)
    if (pending) {
        __syncthreads();
        flush_tables(table_set(tset ^ 1u), pF0, pM0, pS0, 0, NTHREADS);
    }
    if (tid == 0 && carry->valid) global_apply(lt, pt.status, carry->label, carry->g, carry->bmn, carry->bmx);
}

}  // namespace mk
}  // namespace ta

// The scan kernel on label-free OCT RECORDS (sm_100a) -- candidate product kernel of round 2.
//
// Why: every earlier formulation compared raw voxels against a label once per (label, consumer): the per-voxel pair
// phases of ta_scan.cuh read 18 neighbours per wall voxel, the block / level kernels build a mask per (block, label) from
// 24 raw rows.  Here the raw tile is read ONCE.  Phase M turns every run of 8 voxels plus its two f-neighbours (an "oct
// window", 10 lanes) into a record that does not depend on any label chosen later:
//
//      lo, hi      smallest / largest label of the window
//      notlo       10 bits: lanes that differ from lo   (lo == hi: 0)
//      ovf         the window holds 3 or 4 labels: (labels, lane masks) parked in a small side table (2.9 % of the octs of a
//                  C3-like tissue); 5 or more: BAD, the blocks around it take the per-voxel path
//
// Everything after phase M is scalar work on records.  The mask of ANY label L over a window row is
//      (L == lo ? ~notlo : 0) | (L == hi ? notlo : 0)
// -- four instructions instead of a 16-instruction SIMD compare of five words -- and "is this block one label" is a
// min / max over 24 records.
//
// Phases per brick (128 x 16 x 8 voxels for uint16, 64 x 16 x 8 for uint32; 256 threads):
//   A   the tile (brick + one-voxel halo) by ONE TMA box copy of 18 x 18 x 10 vectors (the box must start on a 16-byte
//       boundary of the row: a box shifted by one voxel faults with "illegal instruction" on the hardware); issued for
//       brick k + 1 as soon as phase M of brick k has ended (the raw tile is not read after phase M: even the per-voxel
//       fallback reads global memory), so the copy overlaps P1 / P2
//   U   one-label tile (background, cell interior): closed-form moments, done
//   M   one record per oct of the tile (2880 / 1440), M2: the side-table entries of the 3- and 4-label octs
//   P1  one 8 x 4 x 2 block per thread: min / max over its 24 records.  One label: closed-form moments.  Otherwise -> list
//   P2  listed blocks, one per lane, full warps: label after label (the label at the first window position no earlier label
//       covers), window masks from the records, 18-dilation, moments from a 256-entry byte table, pair counts as
//       popcounts (the bit algebra of ta_block.cuh); blocks with more than MK_MAXL labels finish on a restricted per-voxel
//       path inside the warp
//   Every table update is merged across the warp first (one row per distinct label / pair, redux) and the group leaders
//   add straight into the GLOBAL tables: no per-brick shared tables, no flush phase.
#pragma once
#include "ta_block.cuh"

#ifndef TA_STAT
#define TA_STAT(which, n) ((void)0)
#endif

namespace ta {

constexpr int MK_ROWV = ROWV;                        // vectors per tile row: one halo vector, the brick row, one halo vector
constexpr int MK_ROWS = (BS + 2) * (BM + 2);         // 180 tile rows
constexpr int MK_VECS = MK_ROWS * MK_ROWV;
constexpr int MK_OVF = 256;                          // side-table entries per brick
constexpr int MK_MAXL = 4;                           // labels per block by bit algebra
constexpr uint32_t MK_QOVF = 0x8000u, MK_QBAD = 0x81FFu;
constexpr int MK_HDR = 128;                          // mbarrier + counters, in front of the tile

template <typename T> struct MkGeo {
    static constexpr int SEG = Vox<T>::SEG;
    static constexpr int ROWE = MK_ROWV * SEG;                 // elements per tile row
    static constexpr int BF = NFS * SEG;                       // brick width in voxels
    static constexpr int NOCT = BF / 8;                        // octs per brick row
    static constexpr int OV = 8 / SEG;                         // vectors per oct
    static constexpr int NBLK = NOCT * (BM / BLK_M) * (BS / BLK_S);
    static constexpr int PSTRIDE = (BM + 2) * NOCT + 8;        // records per tile plane (+8: the two planes of a warp's
                                                               // blocks fall on different banks)
    static constexpr int NREC = (BS + 2) * PSTRIDE;
    static constexpr int NITEMS = MK_ROWS * NOCT;
};

template <typename T> constexpr size_t scan_meta_smem_bytes() {
    return (size_t)MK_HDR + (size_t)MK_VECS * 16 + (size_t)MkGeo<T>::NREC * 4 * (sizeof(T) == 4 ? 2 : 1) + (size_t)MkGeo<T>::NREC * 2 +
           (size_t)MK_OVF * 4 * sizeof(T) + MK_OVF * 4 + MK_OVF * 2 + NTHREADS * 2 + 256 * 4;
}

template <typename T> struct MkShared {
    uint64_t* bar;             // the tile copy's mbarrier
    unsigned int* ctr;         // [0..1] brick index ping-pong, [2] listed blocks, [3] 3+-label octs, [4] tile copy timed out
    uint4* tile;               // [MK_VECS]
    uint32_t* recA;            // uint16: lo | (0xFFFF - hi) << 16;  uint32: lo
    uint32_t* recB;            // uint32 only: hi
    unsigned short* Q;         // notlo (bits 0..9) | MK_QOVF + side-table index | MK_QBAD
    T* ovf_lab;                // [MK_OVF][4]
    uint32_t* ovf_msk;         // [MK_OVF] lane masks of labels 0, 1, 2 at bits 0, 10, 20 (label 3: the rest)
    unsigned short* ovf_rec;   // [MK_OVF] record index of the entry
    unsigned short* list;      // [NTHREADS] blocks that are not one label
    uint32_t* momtab;          // [256] block_byte_moments_packed
};

TA_HD uint32_t mk_swap16(uint32_t x) { return (x >> 16) | (x << 16); }

// ---- records --------------------------------------------------------------------------------------------------------------
template <typename T> struct MkRec;
template <> struct MkRec<uint16_t> {
    static TA_HD void store(const MkShared<uint16_t>& sh, int i, uint32_t lo, uint32_t hi) { sh.recA[i] = lo | ((hi ^ 0xFFFFu) << 16); }
    static TA_HD void load(const MkShared<uint16_t>& sh, int i, uint32_t& lo, uint32_t& hi) {
        const uint32_t x = sh.recA[i];
        lo = x & 0xFFFFu; hi = (x >> 16) ^ 0xFFFFu;
    }
    struct MinMax {
        uint32_t acc = 0xFFFFFFFFu;
        TA_HD void add(const MkShared<uint16_t>& sh, int i) { acc = ta_vminu2(acc, sh.recA[i]); }
        TA_HD uint32_t lo() const { return acc & 0xFFFFu; }
        TA_HD uint32_t hi() const { return (acc >> 16) ^ 0xFFFFu; }
    };
};
template <> struct MkRec<uint32_t> {
    static TA_HD void store(const MkShared<uint32_t>& sh, int i, uint32_t lo, uint32_t hi) { sh.recA[i] = lo; sh.recB[i] = hi; }
    static TA_HD void load(const MkShared<uint32_t>& sh, int i, uint32_t& lo, uint32_t& hi) { lo = sh.recA[i]; hi = sh.recB[i]; }
    struct MinMax {
        uint32_t mn = 0xFFFFFFFFu, mx = 0u;
        TA_HD void add(const MkShared<uint32_t>& sh, int i) {
            const uint32_t a = sh.recA[i], b = sh.recB[i];
            mn = a < mn ? a : mn; mx = b > mx ? b : mx;
        }
        TA_HD uint32_t lo() const { return mn; }
        TA_HD uint32_t hi() const { return mx; }
    };
};

// ---- phase M: one oct window -> lo, hi, lanes != lo, lanes != hi -----------------------------------------------------------
struct MkOct { uint32_t lo, hi, notlo, nothi; };
// uint16: the oct's four words of two lanes (low half = the earlier column) and the halo word: left neighbour in the low
// half, right neighbour in the high half.  Mask bit 0 = left neighbour, bits 1 .. 8 = the oct, bit 9 = right neighbour.
TA_HD uint32_t mk_fold16(uint32_t z0, uint32_t z1, uint32_t z2, uint32_t z3, uint32_t zh) {
    const uint32_t tt = z0 + (z1 << 2) + (z2 << 4) + (z3 << 6);
    return ((((tt & 0x55u) | ((tt >> 15) & 0xAAu)) << 1) | (zh & 1u)) | ((zh >> 7) & 0x200u);
}
TA_HD MkOct mk_oct16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wh) {
    const uint32_t mn = ta_vminu2(ta_vminu2(ta_vminu2(w0, w1), ta_vminu2(w2, w3)), wh);
    const uint32_t mx = ta_vmaxu2(ta_vmaxu2(ta_vmaxu2(w0, w1), ta_vmaxu2(w2, w3)), wh);
    const uint32_t LL = ta_vminu2(mn, mk_swap16(mn)), HH = ta_vmaxu2(mx, mk_swap16(mx)), one = 0x00010001u;
    MkOct r;
    r.lo = LL & 0xFFFFu; r.hi = HH & 0xFFFFu;
    r.notlo = mk_fold16(ta_vminu2(w0 ^ LL, one), ta_vminu2(w1 ^ LL, one), ta_vminu2(w2 ^ LL, one), ta_vminu2(w3 ^ LL, one),
                        ta_vminu2(wh ^ LL, one));
    r.nothi = mk_fold16(ta_vminu2(w0 ^ HH, one), ta_vminu2(w1 ^ HH, one), ta_vminu2(w2 ^ HH, one), ta_vminu2(w3 ^ HH, one),
                        ta_vminu2(wh ^ HH, one));
    return r;
}
TA_HD MkOct mk_oct32(const uint32_t w[10]) {
    uint32_t mn = w[0], mx = w[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) { mn = w[i] < mn ? w[i] : mn; mx = w[i] > mx ? w[i] : mx; }
    MkOct r;
    r.lo = mn; r.hi = mx; r.notlo = 0u; r.nothi = 0u;
#pragma unroll
    for (int i = 0; i < 10; ++i) { r.notlo += ta_minu(w[i] ^ mn, 1u) << i; r.nothi += ta_minu(w[i] ^ mx, 1u) << i; }
    return r;
}

// record index of tile row (plane tp, row tr), oct o
template <typename T> TA_HD int mk_rec_index(int tp, int tr, int o) { return tp * MkGeo<T>::PSTRIDE + tr * MkGeo<T>::NOCT + o; }

// ---- table updates: warp merge, then the group leaders add to the GLOBAL tables ---------------------------------------
// One label row per lane in packed form (a lane's row holds at most one block: 64 voxels, brick-local f < 128, m < 16,
// s < 8, so the ten sums of a whole warp fit seven words and the m / s boxes one bit mask):
//   w0 = n [12] | sf << 12 [18]     w1 = sm [15] | sss << 15 [17]     w2 = ss [14] | sms << 14 [18]
//   w3 = sff   w4 = sfm   w5 = sfs   w6 = smm      w7 = f min   w8 = f max
//   w9 = 1 << m min | 1 << m max | (1 << s min | 1 << s max) << 16
constexpr int MK_ROW = 10;
TA_HD void mk_pack_row(const uint32_t v[16], uint32_t w[MK_ROW]) {
    w[0] = v[0] | (v[1] << 12); w[1] = v[2] | (v[9] << 15); w[2] = v[3] | (v[8] << 14);
    w[3] = v[4]; w[4] = v[5]; w[5] = v[6]; w[6] = v[7];
    w[7] = v[10]; w[8] = v[13];
    w[9] = (1u << v[11]) | (1u << v[14]) | (((1u << v[12]) | (1u << v[15])) << 16);
}
TA_HD void mk_unpack_row(const uint32_t w[MK_ROW], uint32_t u[16]) {
    u[0] = w[0] & 0xFFFu; u[1] = w[0] >> 12; u[2] = w[1] & 0x7FFFu; u[9] = w[1] >> 15;
    u[3] = w[2] & 0x3FFFu; u[8] = w[2] >> 14; u[4] = w[3]; u[5] = w[4]; u[6] = w[5]; u[7] = w[6];
    u[10] = w[7]; u[13] = w[8];
    u[11] = (uint32_t)ta_ffs(w[9] & 0xFFFFu) - 1u; u[14] = (uint32_t)ta_fls(w[9] & 0xFFFFu);
    u[12] = (uint32_t)ta_ffs(w[9] >> 16) - 1u; u[15] = (uint32_t)ta_fls(w[9] >> 16);
}

// all 32 lanes call; has = false: nothing to add
__device__ __forceinline__ void mk_put_label(const LabelTable& lt, uint32_t* status, bool has, uint32_t L, const uint32_t w[MK_ROW],
                                             u64 gF0, u64 gM0, u64 gS0, int lane) {
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
    if (am_leader) {
        uint32_t u[16];
        mk_unpack_row(tot, u);
        label_to_global(lt, status, L, u, gF0, gM0, gS0);
    }
}

// packed pair increments [w18|f0] [f1|f2] [f3|f4] [f5|-] (16-bit counters: a warp adds at most 32 blocks of 64 voxels)
template <typename T>
__device__ __forceinline__ void mk_put_pair(const PairTable& pt, bool has, uint32_t a, uint32_t b, const uint32_t inc[4], int lane) {
    typedef typename Vox<T>::PKey PKey;
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
    if (am_leader) {
        const int g = ta_pair_slot(pt, Vox<T>::key64(key));
        if (g >= 0) {
            uint32_t* v = &pt.vals[(size_t)g * TA_PAIR_STRIDE];
#pragma unroll
            for (int idx = 0; idx < 7; ++idx) {
                const uint32_t n = (tot[idx >> 1] >> ((idx & 1) * 16)) & 0xFFFFu;
                if (n) atomicAdd(&v[idx == 0 ? 6 : idx - 1], n);
            }
        }
    }
}

// ---- masks of one label over a block's window, from the records ------------------------------------------------------------
// rbase: record of window plane 0, row 0.  ovfrows: bit (p * 6 + r) set where the record is a side-table entry.
template <typename T>
__device__ __forceinline__ void mk_label_planes(const MkShared<T>& sh, int rbase, uint32_t L, uint32_t ovfrows, u64 A[4]) {
    constexpr int PS = MkGeo<T>::PSTRIDE, NO = MkGeo<T>::NOCT;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        uint32_t h0 = 0u, h1 = 0u;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const int i = rbase + p * PS + r * NO;
            uint32_t lo, hi;
            MkRec<T>::load(sh, i, lo, hi);
            const uint32_t nl = (uint32_t)sh.Q[i] & 0x3FFu;
            uint32_t a = (L == lo ? (nl ^ 0x3FFu) : 0u) | (L == hi ? nl : 0u);
            a = ((ovfrows >> (p * 6 + r)) & 1u) ? 0u : a;
            if (r < 3) h0 |= a << (10 * r); else h1 |= a << (10 * (r - 3));
        }
        A[p] = (u64)h0 | ((u64)h1 << 30);
    }
    uint32_t rem = ovfrows;
    while (rem) {                                     // side-table rows: a few per block at most
        const int rr = __ffs(rem) - 1;
        rem &= rem - 1u;
        const int p = rr / 6, r = rr - p * 6;
        const int e = (int)(sh.Q[rbase + p * PS + r * NO] & 0xFFu);
        const uint32_t mm = sh.ovf_msk[e];
        const uint32_t m0 = mm & 0x3FFu, m1 = (mm >> 10) & 0x3FFu, m2 = (mm >> 20) & 0x3FFu, m3 = ~(m0 | m1 | m2) & 0x3FFu;
        const T* lb = sh.ovf_lab + e * 4;
        const uint32_t a = (L == (uint32_t)lb[0] ? m0 : 0u) | (L == (uint32_t)lb[1] ? m1 : 0u) | (L == (uint32_t)lb[2] ? m2 : 0u) |
                           (L == (uint32_t)lb[3] ? m3 : 0u);
        const u64 sa = (u64)a << (10 * r);
        if (p == 0) A[0] |= sa; else if (p == 1) A[1] |= sa; else if (p == 2) A[2] |= sa; else A[3] |= sa;
    }
}

// the label at window position (plane p, bit = row * 10 + lane)
template <typename T>
__device__ __forceinline__ uint32_t mk_label_at(const MkShared<T>& sh, int rbase, int p, int bit) {
    const int r = (bit * 205) >> 11, x = bit - r * 10;
    const int i = rbase + p * MkGeo<T>::PSTRIDE + r * MkGeo<T>::NOCT;
    const uint32_t q = sh.Q[i];
    if (!(q & MK_QOVF)) {
        uint32_t lo, hi;
        MkRec<T>::load(sh, i, lo, hi);
        return ((q >> x) & 1u) ? hi : lo;
    }
    const int e = (int)(q & 0xFFu);
    const uint32_t mm = sh.ovf_msk[e];
    const int k = ((mm >> x) & 1u) ? 0 : ((mm >> (10 + x)) & 1u) ? 1 : ((mm >> (20 + x)) & 1u) ? 2 : 3;
    return (uint32_t)sh.ovf_lab[e * 4 + k];
}

// ---- per-voxel path (global memory, clamped), restricted to what the label steps could not emit -------------------------------
template <typename T>
__device__ __forceinline__ uint32_t mk_vox(const ScanParams& P, int f, int m, int s) {
    f = max(0, min(f, (int)P.nf - 1)); m = max(0, min(m, (int)P.nm - 1)); s = max(0, min(s, (int)P.ns - 1));
    return (uint32_t)reinterpret_cast<const T*>(P.vol)[((size_t)s * (size_t)P.nm + (size_t)m) * (size_t)P.nf + (size_t)f];
}
// (f, m, s): buffer coordinates of the voxel; (bf, bm, bs): its brick-local coordinates; known[0 .. nk - 1]: the labels
// whose mutual contributions were emitted by the steps
template <typename T>
__device__ __noinline__ void mk_fallback_voxel(const ScanParams P, const LabelTable lt, const PairTable pt, int f, int m, int s,
                                               uint32_t bf, uint32_t bm, uint32_t bs, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3,
                                               int nk, u64 gF0, u64 gM0, u64 gS0) {
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    const uint32_t known[4] = {k0, k1, k2, k3};
    const uint32_t a = mk_vox<T>(P, f, m, s);
    bool a_in = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) a_in = a_in || (i < nk && known[i] == a);
    if (do_mom && !a_in) {
        uint32_t v[16] = {1u, bf, bm, bs, bf * bf, bf * bm, bf * bs, bm * bm, bm * bs, bs * bs, bf, bm, bs, bf, bm, bs};
        label_to_global(lt, pt.status, a, v, gF0, gM0, gS0);
    }
    if (!(do_p6 || do_w18)) return;
    // 18-neighbourhood, faces +f / +m / +s first
    const signed char df[18] = {1, 0, 0, -1, 0, 0, -1, 1, -1, 1, -1, 1, -1, 1, 0, 0, 0, 0};
    const signed char dm[18] = {0, 1, 0, 0, -1, 0, -1, -1, 1, 1, 0, 0, 0, 0, -1, 1, -1, 1};
    const signed char ds[18] = {0, 0, 1, 0, 0, -1, 0, 0, 0, 0, -1, -1, 1, 1, -1, -1, 1, 1};
#pragma unroll 1
    for (int k = 0; k < 18; ++k) {
        const uint32_t b = mk_vox<T>(P, f + df[k], m + dm[k], s + ds[k]);
        if (b == a) continue;
        if (a_in) {
            bool b_in = false;
#pragma unroll
            for (int i = 0; i < 4; ++i) b_in = b_in || (i < nk && known[i] == b);
            if (b_in) continue;
        }
        if (do_p6 && k < 3) ta_pair_add(pt, ta_pair_key(a, b), 2 * k + (a < b ? 0 : 1), 1u);
        if (do_w18) {
            bool seen = false;
            for (int q = 0; q < k; ++q) seen |= (mk_vox<T>(P, f + df[q], m + dm[q], s + ds[q]) == b);
            if (!seen) ta_pair_add(pt, ta_pair_key(a, b), 6, 1u);
        }
    }
}

// ---- P2: what step I adds for a block: the moments of slot I and its pairs with the older slots --------------------------
template <typename T, int I>
__device__ __forceinline__ void mk_emit_slot(const ScanParams& P, const LabelTable& lt, const PairTable& pt, const uint32_t* momtab,
                                             const BlockLevel<T, MK_MAXL>& b, bool active, uint32_t bF, uint32_t bM, uint32_t bS,
                                             u64 gF0, u64 gM0, u64 gS0, int lane) {
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    uint32_t w[MK_ROW], inc[I > 0 ? I : 1][4];
    bool hasp[I > 0 ? I : 1];
    bool has = false;
    if (active && do_mom) {
        uint32_t v[16];
        has = b.label_moments(I, momtab, v);
        if (has) { block_shift_moments(v, bF, bM, bS); mk_pack_row(v, w); }
    }
#pragma unroll
    for (int j = 0; j < I; ++j) hasp[j] = active && (do_p6 || do_w18) && b.pair_increments(I, j, do_p6, do_w18, inc[j]);
    mk_put_label(lt, pt.status, has, b.lab[I], w, gF0, gM0, gS0, lane);
    if (do_p6 || do_w18) {
#pragma unroll
        for (int j = 0; j < I; ++j) mk_put_pair<T>(pt, hasp[j], b.lab[I], b.lab[j], inc[j], lane);
    }
}

template <typename T, int I>
__device__ __forceinline__ void mk_steps(const MkShared<T>& sh, const ScanParams& P, const LabelTable& lt, const PairTable& pt,
                                         BlockLevel<T, MK_MAXL>& b, bool& more, int rbase, uint32_t ovfrows, uint32_t bF, uint32_t bM,
                                         uint32_t bS, u64 gF0, u64 gM0, u64 gS0, int lane) {
    if constexpr (I < MK_MAXL) {
        if (!__ballot_sync(0xffffffffu, more)) return;
        const bool act = more;
        if (lane == 0) TA_STAT(5, 1);
        if (act) TA_STAT(6, 1);
        if (act) {
            constexpr u64 ALL = LvBlk<T>::PLANE_ALL;
            const int p = b.R0 ? 0 : b.R1 ? 1 : b.R2 ? 2 : 3;
            const u64 rp = b.R0 ? b.R0 : b.R1 ? b.R1 : b.R2 ? b.R2 : b.R3;
            const uint32_t L = mk_label_at<T>(sh, rbase, p, ta_ffs64(rp) - 1);
            u64 A[4];
            mk_label_planes<T>(sh, rbase, L, ovfrows, A);
            const u64 neq[4] = {~A[0] & ALL, ~A[1] & ALL, ~A[2] & ALL, ~A[3] & ALL};
            b.template set_slot<I>(L, neq);
            more = (b.R0 | b.R1 | b.R2 | b.R3) != 0ull;
        } else {
            b.template clear_slot<I>();
        }
        mk_emit_slot<T, I>(P, lt, pt, sh.momtab, b, act, bF, bM, bS, gF0, gM0, gS0, lane);
        mk_steps<T, I + 1>(sh, P, lt, pt, b, more, rbase, ovfrows, bF, bM, bS, gF0, gM0, gS0, lane);
    }
}

__device__ __forceinline__ u64 mk_globaltimer() {
#if defined(TA_EMU_TMA)
    return 0ull;
#else
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#endif
}

#ifndef TA_META_MINB
#define TA_META_MINB 3
#endif
template <typename T>
__global__ void __launch_bounds__(NTHREADS, TA_META_MINB)
scan_meta_kernel(ScanParams P, LabelTable lt, PairTable pt, const __grid_constant__ CUtensorMap tmap) {
    typedef MkGeo<T> G;
    constexpr int SEG = G::SEG, ROWE = G::ROWE, BF = G::BF, NOCT = G::NOCT, OV = G::OV, PS = G::PSTRIDE;
    static_assert(G::NBLK <= NTHREADS, "at most one block per thread in P1");
    static_assert(G::NITEMS % 32 == 0 && NTHREADS % NOCT == 0, "whole warps in phase M");
    static_assert(G::NREC < 65536, "record indices fit 16 bits");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    MkShared<T> sh;
    sh.bar = reinterpret_cast<uint64_t*>(smem_raw);
    sh.ctr = reinterpret_cast<unsigned int*>(smem_raw + 16);
    sh.tile = reinterpret_cast<uint4*>(smem_raw + MK_HDR);
    sh.recA = reinterpret_cast<uint32_t*>(sh.tile + MK_VECS);
    sh.recB = sh.recA + (sizeof(T) == 4 ? G::NREC : 0);
    sh.ovf_msk = sh.recB + G::NREC;
    sh.momtab = sh.ovf_msk + MK_OVF;
    sh.ovf_lab = reinterpret_cast<T*>(sh.momtab + 256);
    sh.Q = reinterpret_cast<unsigned short*>(sh.ovf_lab + MK_OVF * 4);
    sh.ovf_rec = sh.Q + G::NREC;
    sh.list = sh.ovf_rec + MK_OVF;
    T* tileT = reinterpret_cast<T*>(sh.tile);

    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned int total = (unsigned int)P.nbf * P.nbm * P.nbs;
    const bool do_mom = P.flags & 1u;
    const int nf = (int)P.nf, nm = (int)P.nm, ns = (int)P.ns;

    for (int i = tid; i < 256; i += NTHREADS) sh.momtab[i] = block_byte_moments_packed((uint32_t)i);
    const bool use_tma = P.use_tma && ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u) == 0u;
    uint32_t tma_parity = 0u;
    if (tid == 0) {
        if (use_tma) {
            mbar_init(sh.bar, 1u);
            TA_PTX("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        sh.ctr[4] = 0u;
        const unsigned int b0 = atomicAdd(P.brick_counter, 1u);
        sh.ctr[0] = b0;
        if (use_tma && b0 < total) {
            const int bf = b0 % P.nbf, bm = (b0 / P.nbf) % P.nbm, bs = b0 / (P.nbf * P.nbm);
            TA_PTX("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(sh.bar, (uint32_t)(MK_VECS * 16));
            tma_load_box_3d(sh.tile, &tmap, sh.bar, bf * BF - SEG, bm * BM - 1, (int)P.own_lo + bs * BS - 1);
        }
    }
    __syncthreads();

    for (unsigned iter = 0;; ++iter) {
        const unsigned int brick = sh.ctr[iter & 1u];
        if (brick >= total) break;
        unsigned int next_brick = 0xFFFFFFFFu;                 // thread 0 only
        if (tid == 0) {
            next_brick = atomicAdd(P.brick_counter, 1u);
            sh.ctr[(iter + 1u) & 1u] = next_brick;
            sh.ctr[2] = 0u; sh.ctr[3] = 0u;
        }
        const int bf = brick % P.nbf, bm = (brick / P.nbf) % P.nbm, bs = brick / (P.nbf * P.nbm);
        const int F0 = bf * BF, M0 = bm * BM, S0 = (int)P.own_lo + bs * BS;
        const u64 gF0 = (u64)F0, gM0 = (u64)M0, gS0 = (u64)((long long)S0 + P.slow_offset);

        // ---- phase A: the tile ----------------------------------------------------------------------------------------------
        if (use_tma) {
            unsigned spins = 0;
            u64 t0 = 0ull;
            while (!mbar_try_wait(sh.bar, tma_parity)) {
                if ((++spins & 0x3FFu) == 0u) {
                    const u64 now = mk_globaltimer();
                    if (t0 == 0ull) t0 = now;
                    else if (now - t0 > 4000000000ull) {      // 4 s: the copy is lost; report, do not trap
                        sh.ctr[4] = 1u;
                        if (P.diag && atomicAdd(&P.diag[0], 1ull) == 0ull) {
                            P.diag[1] = ((u64)blockIdx.x << 32) | (u64)tid;
                            P.diag[2] = ((u64)iter << 32) | (u64)brick;
                            __threadfence_system();
                        }
                        break;
                    }
                }
            }
            tma_parity ^= 1u;
            // out-of-buffer elements arrive as zeros: bricks on a face of the buffer re-clamp them
            const bool edge = (F0 == 0) | (F0 + BF + 1 > nf) | (M0 == 0) | (M0 + BM + 1 > nm) | (S0 < 1) | (S0 + BS + 1 > ns);
            if (edge) {
                const int xl = (F0 == 0) ? SEG : 0;                  // elements before volume column 0
                const int xr = min(ROWE, nf - F0 + SEG);             // first element beyond the last volume column
                for (int r = tid; r < MK_ROWS; r += NTHREADS) {
                    T* row = tileT + r * ROWE;
                    if (xl) { const T v = row[xl]; for (int x = 0; x < xl; ++x) row[x] = v; }
                    if (xr < ROWE) { const T v = row[xr - 1]; for (int x = xr; x < ROWE; ++x) row[x] = v; }
                }
                __syncthreads();
                for (int i = tid; i < MK_VECS; i += NTHREADS) {
                    const int r = i / MK_ROWV, tp = r / (BM + 2), tr = r - tp * (BM + 2);
                    const int m = M0 - 1 + tr, s = S0 - 1 + tp;
                    const int cm = max(0, min(m, nm - 1)), cs = max(0, min(s, ns - 1));
                    if (cm != m || cs != s) {
                        const int src = (cs - (S0 - 1)) * (BM + 2) + (cm - (M0 - 1));
                        sh.tile[i] = sh.tile[src * MK_ROWV + (i - r * MK_ROWV)];
                    }
                }
                __syncthreads();
            }
        } else {
            const T* vol = reinterpret_cast<const T*>(P.vol);
            for (int e = tid; e < MK_ROWS * ROWE; e += NTHREADS) {
                const int r = e / ROWE, x = e - r * ROWE, tp = r / (BM + 2), tr = r - tp * (BM + 2);
                const int f = max(0, min(F0 - SEG + x, nf - 1)), m = max(0, min(M0 - 1 + tr, nm - 1)), s = max(0, min(S0 - 1 + tp, ns - 1));
                tileT[e] = vol[((size_t)s * nm + m) * (size_t)nf + f];
            }
            __syncthreads();
        }

        // ---- phase U: one-label tile (columns -1 .. BF of every row) -------------------------------------------------------
        {
            const uint32_t ref = (uint32_t)tileT[SEG];
            const uint32_t pat = (SEG == 8) ? ref * 0x00010001u : ref;
            uint32_t diff = 0u;
            for (int i = tid; i < MK_VECS; i += NTHREADS) {
                const uint4 v = sh.tile[i];
                const int j = i % MK_ROWV;
                if (j == 0) diff |= (SEG == 8) ? ((v.w ^ pat) >> 16) : (v.w ^ pat);                   // column -1
                else if (j == MK_ROWV - 1) diff |= (SEG == 8) ? ((v.x ^ pat) & 0xFFFFu) : (v.x ^ pat);   // column BF
                else diff |= (v.x ^ pat) | (v.y ^ pat) | (v.z ^ pat) | (v.w ^ pat);
            }
            const bool uniform = __syncthreads_and(diff == 0u);
            if (sh.ctr[4]) {                                   // a tile copy timed out somewhere in this CTA
                if (tid == 0) atomicExch(&pt.status[2], 1u);
                return;
            }
            if (uniform) {
                if (tid == 0) {
                    if (use_tma && next_brick < total) {
                        const int nbf_ = next_brick % P.nbf, nbm_ = (next_brick / P.nbf) % P.nbm, nbs_ = next_brick / (P.nbf * P.nbm);
                        TA_PTX("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_arrive_expect_tx(sh.bar, (uint32_t)(MK_VECS * 16));
                        tma_load_box_3d(sh.tile, &tmap, sh.bar, nbf_ * BF - SEG, nbm_ * BM - 1, (int)P.own_lo + nbs_ * BS - 1);
                    }
                    if (do_mom) {
                        uint32_t v[16];
                        block_uniform_moments((uint32_t)min(BF, nf - F0), (uint32_t)min(BM, nm - M0), (uint32_t)min(BS, (int)P.own_hi - S0), v);
                        label_to_global(lt, pt.status, ref, v, gF0, gM0, gS0);
                    }
                }
                if (!use_tma) __syncthreads();                 // the scalar staging of the next brick overwrites the tile
                continue;
            }
        }

        if (tid == 0) TA_STAT(0, 1);
        // ---- phase M: one record per oct of the tile -------------------------------------------------------------------------
        for (int base = tid - lane; base < G::NITEMS; base += NTHREADS) {
            const int i = base + lane;
            const int r = i / NOCT, o = i - r * NOCT, tp = r / (BM + 2), tr = r - tp * (BM + 2);
            const uint4* vp = sh.tile + r * MK_ROWV + 1 + o * OV;
            MkOct mo;
            if constexpr (sizeof(T) == 2) {
                const uint4 c = vp[0];
                // the neighbour lanes come from the octs either side: lanes of this warp, except at the ends of the row
                uint32_t pw = __shfl_sync(0xffffffffu, c.w, (lane + 31) & 31), nx = __shfl_sync(0xffffffffu, c.x, (lane + 1) & 31);
                if (o == 0) pw = vp[-1].w;
                if (o == NOCT - 1) nx = vp[1].x;
                mo = mk_oct16(c.x, c.y, c.z, c.w, (pw >> 16) | (nx << 16));
            } else {
                const uint4 c = vp[0], d = vp[1];
                const uint32_t w[10] = {vp[-1].w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w, vp[2].x};
                mo = mk_oct32(w);
            }
            const int ri = mk_rec_index<T>(tp, tr, o);
            MkRec<T>::store(sh, ri, mo.lo, mo.hi);
            const bool ov = (mo.notlo & mo.nothi) != 0u;            // a lane that is neither lo nor hi: 3 or more labels
            uint32_t q = mo.notlo;
            const unsigned bal = __ballot_sync(0xffffffffu, ov);
            if (bal) {
                unsigned int pos0 = 0u;
                if (lane == 0) pos0 = atomicAdd(&sh.ctr[3], (unsigned int)__popc(bal));
                pos0 = __shfl_sync(0xffffffffu, pos0, 0);
                if (ov) {
                    const unsigned int pos = pos0 + (unsigned int)__popc(bal & ((1u << lane) - 1u));
                    if (pos < (unsigned)MK_OVF) { sh.ovf_rec[pos] = (unsigned short)ri; q = MK_QOVF | pos; }
                    else q = MK_QBAD;
                }
            }
            sh.Q[ri] = (unsigned short)q;
        }
        __syncthreads();
        // ---- M2: side-table entries of the octs with 3 or 4 labels (5 and more: BAD) ------------------------------------------
        {
            const int novf = min((int)sh.ctr[3], MK_OVF);
            if (tid == 0) TA_STAT(9, sh.ctr[3]);
            for (int e = tid; e < novf; e += NTHREADS) {
                const int ri = (int)sh.ovf_rec[e];
                const int tp = ri / PS, rem = ri - tp * PS, tr = rem / NOCT, o = rem - tr * NOCT;
                const T* el = tileT + (tp * (BM + 2) + tr) * ROWE + SEG - 1 + 8 * o;
                uint32_t l0 = (uint32_t)el[0], l1 = l0, l2 = l0, l3 = l0, m0 = 1u, m1 = 0u, m2 = 0u, m3 = 0u;
                int nl = 1;
                bool bad = false;
#pragma unroll
                for (int x = 1; x < 10; ++x) {
                    const uint32_t v = (uint32_t)el[x], bit = 1u << x;
                    if (v == l0) m0 |= bit;
                    else if (nl > 1 && v == l1) m1 |= bit;
                    else if (nl > 2 && v == l2) m2 |= bit;
                    else if (nl > 3 && v == l3) m3 |= bit;
                    else if (nl == 1) { l1 = v; m1 = bit; nl = 2; }
                    else if (nl == 2) { l2 = v; m2 = bit; nl = 3; }
                    else if (nl == 3) { l3 = v; m3 = bit; nl = 4; }
                    else bad = true;
                }
                if (bad) {
                    sh.Q[ri] = (unsigned short)MK_QBAD;
                } else {
                    // unused slots repeat label 0 with an empty mask; label 3's mask is "the rest"
                    sh.ovf_lab[e * 4 + 0] = (T)l0; sh.ovf_lab[e * 4 + 1] = (T)(nl > 1 ? l1 : l0);
                    sh.ovf_lab[e * 4 + 2] = (T)(nl > 2 ? l2 : l0); sh.ovf_lab[e * 4 + 3] = (T)(nl > 3 ? l3 : l0);
                    sh.ovf_msk[e] = m0 | (m1 << 10) | (m2 << 20);
                    (void)m3;
                }
            }
        }
        __syncthreads();
        // the raw tile is not read again: the copy of the next brick's tile overlaps P1 / P2
        if (tid == 0 && use_tma && next_brick < total) {
            const int nbf_ = next_brick % P.nbf, nbm_ = (next_brick / P.nbf) % P.nbm, nbs_ = next_brick / (P.nbf * P.nbm);
            TA_PTX("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(sh.bar, (uint32_t)(MK_VECS * 16));
            tma_load_box_3d(sh.tile, &tmap, sh.bar, nbf_ * BF - SEG, nbm_ * BM - 1, (int)P.own_lo + nbs_ * BS - 1);
        }

        // ---- P1: one block per thread: one label -> closed form, else -> list ------------------------------------------------
        {
            const int o = tid % NOCT, sbq = (tid / NOCT) % (BS / BLK_S), mbq = tid / (NOCT * (BS / BLK_S));
            const int nvf = min(8, nf - (F0 + 8 * o)), nvm = min(BLK_M, nm - (M0 + BLK_M * mbq)),
                      nvs = min(BLK_S, (int)P.own_hi - (S0 + BLK_S * sbq));
            const bool valid = tid < G::NBLK && nvf > 0 && nvm > 0 && nvs > 0;
            const int rbase = mk_rec_index<T>(BLK_S * sbq, BLK_M * mbq, o);
            typename MkRec<T>::MinMax mmx;
            if (valid) {
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int r = 0; r < 6; ++r) mmx.add(sh, rbase + p * PS + r * NOCT);
            }
            const uint32_t wlo = mmx.lo(), whi = mmx.hi();
            const bool one = valid && wlo == whi, many = valid && wlo != whi;
            if (valid) TA_STAT(1, 1);
            if (one) TA_STAT(2, 1);
            if (many) TA_STAT(3, 1);
            const unsigned mm = __ballot_sync(0xffffffffu, many);
            unsigned int pos0 = 0u;
            if (lane == 0 && mm) pos0 = atomicAdd(&sh.ctr[2], (unsigned int)__popc(mm));
            pos0 = __shfl_sync(0xffffffffu, pos0, 0);
            if (many) sh.list[pos0 + (unsigned int)__popc(mm & ((1u << lane) - 1u))] = (unsigned short)tid;
            uint32_t w[MK_ROW];
            const bool has = one && do_mom;
            if (has) {
                uint32_t v[16];
                block_uniform_moments((uint32_t)nvf, (uint32_t)nvm, (uint32_t)nvs, v);
                block_shift_moments(v, (uint32_t)(8 * o), (uint32_t)(BLK_M * mbq), (uint32_t)(BLK_S * sbq));
                mk_pack_row(v, w);
            }
            mk_put_label(lt, pt.status, has, wlo, w, gF0, gM0, gS0, lane);
        }
        __syncthreads();

        // ---- P2: listed blocks, one per lane, label after label ---------------------------------------------------------------
        {
            const int count = (int)sh.ctr[2];
            for (int base = tid - lane; base < count; base += NTHREADS) {
                const int qi = base + lane;
                const bool active = qi < count;
                if (lane == 0) TA_STAT(4, 1);
                const int blk = active ? (int)sh.list[qi] : 0;
                const int o = blk % NOCT, sbq = (blk / NOCT) % (BS / BLK_S), mbq = blk / (NOCT * (BS / BLK_S));
                const int nvf = min(8, nf - (F0 + 8 * o)), nvm = min(BLK_M, nm - (M0 + BLK_M * mbq)),
                          nvs = min(BLK_S, (int)P.own_hi - (S0 + BLK_S * sbq));
                const uint32_t bF = (uint32_t)(8 * o), bM = (uint32_t)(BLK_M * mbq), bS = (uint32_t)(BLK_S * sbq);
                const int rbase = mk_rec_index<T>(BLK_S * sbq, BLK_M * mbq, o);
                uint32_t ovfrows = 0u;
                bool badblk = false;
                if (active) {
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int r = 0; r < 6; ++r) {
                            const uint32_t q = sh.Q[rbase + p * PS + r * NOCT];
                            ovfrows |= (q >> 15) << (p * 6 + r);
                            badblk = badblk || (q == MK_QBAD);
                        }
                }
                BlockLevel<T, MK_MAXL> b;
                b.clear();
                b.set_centre(nvf, nvm, nvs);
                constexpr u64 ALL = LvBlk<T>::PLANE_ALL;
                b.R0 = b.R1 = b.R2 = b.R3 = ALL;
                bool more = active && !badblk;
                mk_steps<T, 0>(sh, P, lt, pt, b, more, rbase, ovfrows, bF, bM, bS, gF0, gM0, gS0, lane);
                // blocks the steps did not finish (more than MK_MAXL labels, or a BAD record: nothing emitted): per-voxel
                // path inside this warp, restricted to contributions that involve a label outside the known set
                const bool fb = active && (more || badblk);
                if (fb) TA_STAT(7, 1);
                if (active && badblk) TA_STAT(8, 1);
                unsigned fm = __ballot_sync(0xffffffffu, fb);
                while (fm) {
                    const int src = __ffs(fm) - 1;
                    fm &= fm - 1u;
                    const int cblk = __shfl_sync(0xffffffffu, blk, src);
                    const int nk = __shfl_sync(0xffffffffu, badblk ? 0 : MK_MAXL, src);
                    const uint32_t k0 = __shfl_sync(0xffffffffu, b.lab[0], src), k1 = __shfl_sync(0xffffffffu, b.lab[1], src),
                                   k2 = __shfl_sync(0xffffffffu, b.lab[2], src), k3 = __shfl_sync(0xffffffffu, b.lab[3], src);
                    const int co = cblk % NOCT, csb = (cblk / NOCT) % (BS / BLK_S), cmb = cblk / (NOCT * (BS / BLK_S));
                    for (int w = lane; w < 64; w += 32) {
                        const int df = w & 7, dm = (w >> 3) & 3, ds = w >> 5;
                        const uint32_t f = (uint32_t)(8 * co + df), m = (uint32_t)(BLK_M * cmb + dm), sp = (uint32_t)(BLK_S * csb + ds);
                        if (F0 + (int)f >= nf || M0 + (int)m >= nm || S0 + (int)sp >= (int)P.own_hi) continue;
                        mk_fallback_voxel<T>(P, lt, pt, F0 + (int)f, M0 + (int)m, S0 + (int)sp, f, m, sp, k0, k1, k2, k3, nk, gF0, gM0, gS0);
                    }
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace ta

// Block-bitmask feature extraction: groundwork for the next scan kernel (NOT launched by the product yet).
//
// Motivation (tools/simt_stats.py): the current pair phases work per segment / per voxel and run
// at a third to a half of their per-thread cost estimate because a warp executes the maximum over its lanes of every
// data-dependent loop.  Here the work unit is a small 3-D block, one per thread -- a brick of 128 x 16 x 8 voxels is
// exactly 256 blocks of 8 x 4 x 2 -- and everything after the discovery of the block's labels is straight-line bit
// algebra:
//
//   window   the block plus a one-voxel halo: 10 x 6 x 4 voxels = four planes of 60 bits (bit = row * 10 + x);
//   labels   at most BLK_MAXLAB distinct labels in the window (41 % of the windows of a C3-like tissue hold one label,
//            38 % two, 16 % three, 4 % four, 0.9 % more: those blocks report `overflow` and take the per-voxel path);
//   masks    one window mask per label: a SIMD compare of every window row against the label, collapsed to bits;
//   moments  of a label = row-wise (count, sum x, sum x^2) of its centre bits, weighted by the row / plane index;
//   pairs    wall18(i -> j) = popcount(centre_i & dilate18(mask_j)); faces = popcount(centre_i & shifted mask_j).
//
// No atomics, no hash, no worklist inside the block.  Everything here is host-compilable (TA_HD);
// tests/host/block_host_check.cu runs it over whole tiles against a brute-force count.  Both label widths: a block row is
// one 16-byte segment, i.e. 8 x 4 x 2 voxels for uint16 (window planes of 60 bits) and 4 x 4 x 2 for uint32 (36 bits);
// the level formulation further down uses 8 x 4 x 2 for both (LvBlk).
#pragma once
#include "ta_scan.cuh"

namespace ta {

constexpr int BLK_M = 4, BLK_S = 2;            // block rows and planes; its f extent is one segment (SEG voxels)
constexpr int BLK_MAXLAB = 4;
// per label width: a window row has SEG + 2 positions (10 for uint16, 6 for uint32), a window plane BLK_M + 2 rows
template <typename T> struct Blk {
    static constexpr int SEG = Vox<T>::SEG;
    static constexpr int ROWBITS = SEG + 2;
    static constexpr u64 PLANE_ALL = (1ull << (ROWBITS * (BLK_M + 2))) - 1ull;
    static constexpr uint32_t LANES = (1u << SEG) - 1u;
};

TA_HD int ta_fls(uint32_t x) {                  // index of the highest set bit (x != 0)
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
TA_HD int ta_popc64(u64 x) { return ta_popc((uint32_t)x) + ta_popc((uint32_t)(x >> 32)); }
TA_HD int ta_ffs64(u64 x) {                    // 1-based index of the lowest set bit, 0 if none
    const uint32_t lo = (uint32_t)x;
    if (lo) return ta_ffs(lo);
    const uint32_t hi = (uint32_t)(x >> 32);
    return hi ? 32 + ta_ffs(hi) : 0;
}

// Mask of one window row (vector index t of its segment in the tile): bit 0 = the lane left of the segment, bits
// 1 .. SEG = the segment's lanes, bit SEG + 1 = the lane right of it; set where the voxel equals label L.
template <typename T> TA_HD uint32_t block_row_mask(const uint4* tile, int t, uint32_t L);
template <> TA_HD uint32_t block_row_mask<uint16_t>(const uint4* tile, int t, uint32_t L) {
    const uint4 c = tile[t];
    const unsigned short* e = reinterpret_cast<const unsigned short*>(tile + t);
    const uint32_t pat = L * 0x00010001u, one = 0x00010001u;
    // 1 in every lane that DIFFERS from L, collapsed to one bit per lane (same collapse as the kernel's boundary flags)
    const uint32_t tt = ta_vminu2(c.x ^ pat, one) | (ta_vminu2(c.y ^ pat, one) << 2) | (ta_vminu2(c.z ^ pat, one) << 4) |
                        (ta_vminu2(c.w ^ pat, one) << 6);
    const uint32_t neq = (tt & 0x55u) | ((tt >> 15) & 0xAAu);
    return ((~neq & 0xFFu) << 1) | ((uint32_t)e[-1] == L ? 1u : 0u) | ((uint32_t)e[8] == L ? 0x200u : 0u);
}
template <> TA_HD uint32_t block_row_mask<uint32_t>(const uint4* tile, int t, uint32_t L) {
    const uint4 c = tile[t];
    const uint32_t* e = reinterpret_cast<const uint32_t*>(tile + t);
    return (e[-1] == L ? 1u : 0u) | (c.x == L ? 2u : 0u) | (c.y == L ? 4u : 0u) | (c.z == L ? 8u : 0u) |
           (c.w == L ? 16u : 0u) | (e[4] == L ? 32u : 0u);
}

// Window masks of label L: plane p = tile plane (s0 - 1 + p), rows (m0 - 1 .. m0 + 4).  `t0` = vector index of the
// window's first row (m0 - 1, s0 - 1) at the block's segment.
template <typename T>
TA_HD void block_label_masks(const uint4* tile, int t0, uint32_t L, u64 mask[4]) {
#pragma unroll
    for (int p = 0; p < BLK_S + 2; ++p) {
        u64 m = 0ull;
        // two rows in flight: enough to cover the shared-memory latency, few enough to keep the row data out of the
        // register budget of the label slots (fully unrolled, ptxas hoists all 24 row loads and spills)
#pragma unroll 2
        for (int r = 0; r < BLK_M + 2; ++r)
            m |= (u64)block_row_mask<T>(tile, t0 + p * PLANEV + r * ROWV, L) << (Blk<T>::ROWBITS * r);
        mask[p] = m;
    }
}

// (count, sum x, sum x^2) of the set bits of a byte, x = bit position 0..7, packed as n | sx << 8 | sxx << 16.
TA_HD uint32_t block_byte_moments(uint32_t b) {
    const uint32_t n = ta_popc(b);
    const uint32_t b0 = ta_popc(b & 0xAAu), b1 = ta_popc(b & 0xCCu), b2 = ta_popc(b & 0xF0u);
    const uint32_t sx = b0 + 2u * b1 + 4u * b2;
    // x^2 = (x0 + 2 x1 + 4 x2)^2 = x0 + 4 x1 + 16 x2 + 4 x0 x1 + 8 x0 x2 + 16 x1 x2 for bit values x0, x1, x2
    const uint32_t sxx = b0 + 4u * b1 + 16u * b2 + 4u * ta_popc(b & 0x88u) + 8u * ta_popc(b & 0xA0u) + 16u * ta_popc(b & 0xC0u);
    return n | (sx << 8) | (sxx << 16);
}

// 18-neighbourhood dilation of a label's window masks, for the two centre planes (p = 1, 2).  Bits outside the centre
// (halo columns / rows, bits beyond the plane) are not meaningful: the callers AND with centre masks.
template <int BLK_ROWBITS>
TA_HD void block_dilate18_rb(const u64 mask[4], u64 dil[2]) {
    u64 own[4], cross[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const u64 P = mask[p];
        const u64 hx = (P << 1) | (P >> 1);
        const u64 hy = (P << BLK_ROWBITS) | (P >> BLK_ROWBITS);
        own[p] = hx | hy | (hx << BLK_ROWBITS) | (hx >> BLK_ROWBITS);      // the 8 in-plane neighbours
        cross[p] = P | hx | hy;                                            // the 5 neighbours in an adjacent plane
    }
    dil[0] = own[1] | cross[0] | cross[2];
    dil[1] = own[2] | cross[1] | cross[3];
}

template <typename T> TA_HD void block_dilate18(const u64 mask[4], u64 dil[2]) { block_dilate18_rb<Blk<T>::ROWBITS>(mask, dil); }

// Move a label's sums from coordinates local to a block (or brick) to coordinates shifted by (F, M, S): the algebra of
// the kernel's label_to_global, in 32 bits (a brick: coordinates < 128, at most 16 384 voxels, sums < 2^32).
TA_HD void block_shift_moments(uint32_t v[16], uint32_t F, uint32_t M, uint32_t S) {
    const uint32_t n = v[0], sf = v[1], sm = v[2], ss = v[3];
    v[4] += 2u * F * sf + n * F * F;
    v[5] += F * sm + M * sf + n * F * M;
    v[6] += F * ss + S * sf + n * F * S;
    v[7] += 2u * M * sm + n * M * M;
    v[8] += M * ss + S * sm + n * M * S;
    v[9] += 2u * S * ss + n * S * S;
    v[1] += n * F; v[2] += n * M; v[3] += n * S;
    v[10] += F; v[11] += M; v[12] += S; v[13] += F; v[14] += M; v[15] += S;
}

// One block.  tile: the brick tile of the scan kernel (labels, clamped halo); (fs, m0, s0): the block's segment, first
// row and first plane inside the brick; nvf / nvm / nvs: how many of its 8 x 4 x 2 voxels per axis lie inside the volume
// and the owned plane range (ragged edges).
//   on_label(L, v[16]): n, sf, sm, ss, sff, sfm, sfs, smm, sms, sss, fmin, mmin, smin, fmax, mmax, smax -- block-local
//                       coordinates, the field order of the kernel's per-brick label table;
//   on_pair(La, Lb, w18, ff, fm, fs): seen from the voxels of La: wall18 voxels towards Lb and +f / +m / +s faces whose
//                       lower voxel is La and upper voxel Lb.
// false: more than BLK_MAXLAB labels in the window (nothing was emitted).
template <typename T, typename OnLabel, typename OnPair>
TA_HD bool block_features(const uint4* tile, int fs, int m0, int s0, int nvf, int nvm, int nvs, OnLabel&& on_label,
                          OnPair&& on_pair) {
    constexpr int SEG = Blk<T>::SEG, BLK_ROWBITS = Blk<T>::ROWBITS;
    constexpr u64 BLK_PLANE_ALL = Blk<T>::PLANE_ALL;
    const int t0 = s0 * PLANEV + m0 * ROWV + (fs + 1);          // tile row (m0 - 1, s0 - 1): the tile itself has a halo
    const T* tl = reinterpret_cast<const T*>(tile);
    uint32_t lab[BLK_MAXLAB];
    u64 mask[BLK_MAXLAB][4];
    u64 rest[4] = {BLK_PLANE_ALL, BLK_PLANE_ALL, BLK_PLANE_ALL, BLK_PLANE_ALL};
    int k = 0;
    for (;;) {
        int p = 0;
        while (p < 4 && rest[p] == 0ull) ++p;
        if (p == 4) break;
        if (k == BLK_MAXLAB) return false;
        const int bit = ta_ffs64(rest[p]) - 1, r = bit / BLK_ROWBITS, x = bit % BLK_ROWBITS;
        // window position (x, r, p) = tile element of row t0 + p * PLANEV + r * ROWV, lane x - 1
        const uint32_t L = tl[(size_t)(t0 + p * PLANEV + r * ROWV) * SEG + (x - 1)];
        lab[k] = L;
        block_label_masks<T>(tile, t0, L, mask[k]);
#pragma unroll
        for (int q = 0; q < 4; ++q) rest[q] &= ~mask[k][q];
        ++k;
    }
    // centre voxels inside the volume: x = 1 .. nvf, rows 1 .. nvm, planes 1 .. nvs
    u64 cv = 0ull;
    for (int r = 1; r <= nvm; ++r) cv |= (u64)(((1u << nvf) - 1u) << 1) << (BLK_ROWBITS * r);
    const u64 cvp[2] = {nvs >= 1 ? cv : 0ull, nvs >= 2 ? cv : 0ull};

    u64 cen[BLK_MAXLAB][2], dil[BLK_MAXLAB][2];
    for (int i = 0; i < k; ++i) {
        cen[i][0] = mask[i][1] & cvp[0];
        cen[i][1] = mask[i][2] & cvp[1];
        block_dilate18<T>(mask[i], dil[i]);
    }
    for (int i = 0; i < k; ++i) {
        if (!(cen[i][0] | cen[i][1])) continue;
        // ---- moments of label i over its centre voxels
        uint32_t v[16] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
        uint32_t colmask = 0u;
        for (int p = 0; p < 2; ++p)
            for (int r = 0; r < BLK_M; ++r) {
                const uint32_t b = (uint32_t)(cen[i][p] >> (BLK_ROWBITS * (r + 1) + 1)) & Blk<T>::LANES;
                if (!b) continue;
                const uint32_t t = block_byte_moments(b), n = t & 0xFFu, sx = (t >> 8) & 0xFFu, sxx = t >> 16;
                const uint32_t m = (uint32_t)r, s = (uint32_t)p;
                v[0] += n; v[1] += sx; v[2] += m * n; v[3] += s * n;
                v[4] += sxx; v[5] += m * sx; v[6] += s * sx; v[7] += m * m * n; v[8] += m * s * n; v[9] += s * s * n;
                colmask |= b;
                v[11] = v[11] < m ? v[11] : m; v[14] = v[14] > m ? v[14] : m;
                v[12] = v[12] < s ? v[12] : s; v[15] = v[15] > s ? v[15] : s;
            }
        v[10] = (uint32_t)ta_ffs(colmask) - 1u;
        uint32_t top = 7u;
        while (!((colmask >> top) & 1u)) --top;
        v[13] = top;
        on_label(lab[i], v);
        // ---- pairs seen from label i
        for (int j = 0; j < k; ++j) {
            if (j == i) continue;
            const uint32_t w18 = ta_popc64(cen[i][0] & dil[j][0]) + ta_popc64(cen[i][1] & dil[j][1]);
            const uint32_t ff = ta_popc64(cen[i][0] & (mask[j][1] >> 1)) + ta_popc64(cen[i][1] & (mask[j][2] >> 1));
            const uint32_t fm = ta_popc64(cen[i][0] & (mask[j][1] >> BLK_ROWBITS)) +
                                ta_popc64(cen[i][1] & (mask[j][2] >> BLK_ROWBITS));
            const uint32_t fsl = ta_popc64(cen[i][0] & mask[j][2]) + ta_popc64(cen[i][1] & mask[j][3]);
            if (w18 | ff | fm | fsl) on_pair(lab[i], lab[j], w18, ff, fm, fsl);
        }
    }
    return true;
}

// The same block in register-resident form: every array index is a compile-time constant after unrolling (label slots
// 0 .. MAXLAB - 1, a slot is used iff its index is below the number of labels found), so nothing lives in local memory.
// Per slot: the label, planes 1 .. 3 of its window mask (plane 0 is folded into the dilation at once) and the dilation
// of the two centre planes: five 64-bit words.  A kernel calls discover() once and then label_moments(i) /
// pair_counts(i, j) from loops with the same trip count in every lane (empty slots answer false), so that the warp
// merges around them stay full-mask.
template <typename T, int MAXLAB> struct BlockSlots {
    uint32_t lab[MAXLAB];
    u64 M1[MAXLAB], M2[MAXLAB], M3[MAXLAB], D0[MAXLAB], D1[MAXLAB];
    u64 cv0, cv1;             // centre voxels inside the volume, planes 1 and 2 of the window
    int k;                    // labels found

    TA_HD void clear() {                              // every slot empty
        k = 0; cv0 = cv1 = 0ull;
#pragma unroll
        for (int sl = 0; sl < MAXLAB; ++sl) { lab[sl] = 0u; M1[sl] = M2[sl] = M3[sl] = D0[sl] = D1[sl] = 0ull; }
    }

    // false: more than MAXLAB labels in the window
    TA_HD bool discover(const uint4* tile, int fs, int m0, int s0, int nvf, int nvm, int nvs) {
        constexpr int SEG = Blk<T>::SEG, ROWBITS = Blk<T>::ROWBITS;
        constexpr u64 ALL = Blk<T>::PLANE_ALL;
        const int t0 = s0 * PLANEV + m0 * ROWV + (fs + 1);
        const T* tl = reinterpret_cast<const T*>(tile);
        u64 r0 = ALL, r1 = ALL, r2 = ALL, r3 = ALL;
        k = 0;
#pragma unroll
        for (int sl = 0; sl < MAXLAB; ++sl) {
            lab[sl] = 0u; M1[sl] = M2[sl] = M3[sl] = D0[sl] = D1[sl] = 0ull;
            if (r0 | r1 | r2 | r3) {
                // first window position that no label found so far covers
                const int p = r0 ? 0 : r1 ? 1 : r2 ? 2 : 3;
                const u64 rp = r0 ? r0 : r1 ? r1 : r2 ? r2 : r3;
                const int bit = ta_ffs64(rp) - 1, r = bit / ROWBITS, x = bit % ROWBITS;
                const uint32_t L = tl[(size_t)(t0 + p * PLANEV + r * ROWV) * SEG + (x - 1)];
                u64 m[4];
                block_label_masks<T>(tile, t0, L, m);
                r0 &= ~m[0]; r1 &= ~m[1]; r2 &= ~m[2]; r3 &= ~m[3];
                u64 d[2];
                block_dilate18<T>(m, d);
                lab[sl] = L; M1[sl] = m[1]; M2[sl] = m[2]; M3[sl] = m[3]; D0[sl] = d[0]; D1[sl] = d[1];
                k = sl + 1;
            }
        }
        u64 cv = 0ull;
#pragma unroll
        for (int r = 1; r <= BLK_M; ++r)
            if (r <= nvm) cv |= (u64)(((1u << nvf) - 1u) << 1) << (ROWBITS * r);
        cv0 = nvs >= 1 ? cv : 0ull;
        cv1 = nvs >= 2 ? cv : 0ull;
        return !(r0 | r1 | r2 | r3);
    }

    // moments and box of slot i over its centre voxels, block-local coordinates (field order of the per-brick label
    // table); false: the slot is empty or has no centre voxel
    TA_HD bool label_moments(int i, uint32_t v[16]) const {
        constexpr int ROWBITS = Blk<T>::ROWBITS;
        if (i >= k) return false;
        const u64 c0 = M1[i] & cv0, c1 = M2[i] & cv1;
        if (!(c0 | c1)) return false;
        uint32_t n_ = 0u, sf = 0u, sm = 0u, ss = 0u, sff = 0u, sfm = 0u, sfs = 0u, smm = 0u, sm1 = 0u, colmask = 0u, rows = 0u;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int r = 0; r < BLK_M; ++r) {
                const uint32_t b = (uint32_t)((p ? c1 : c0) >> (ROWBITS * (r + 1) + 1)) & Blk<T>::LANES;
                const uint32_t t = block_byte_moments(b), n = t & 0xFFu, sx = (t >> 8) & 0xFFu, sxx = t >> 16;
                n_ += n; sf += sx; sm += r * n; ss += p * n; sff += sxx; sfm += r * sx; sfs += p * sx; smm += r * r * n;
                sm1 += p * r * n;
                colmask |= b;
                rows |= b ? (1u << (p * BLK_M + r)) : 0u;
            }
        const uint32_t mrows = (rows | (rows >> BLK_M)) & ((1u << BLK_M) - 1u);
        v[0] = n_; v[1] = sf; v[2] = sm; v[3] = ss; v[4] = sff; v[5] = sfm; v[6] = sfs; v[7] = smm;
        v[8] = sm1;                  // sum m * s: s is 0 or 1
        v[9] = ss;                   // sum s * s = sum s
        v[10] = (uint32_t)ta_ffs(colmask) - 1u; v[11] = (uint32_t)ta_ffs(mrows) - 1u; v[12] = (rows & ((1u << BLK_M) - 1u)) ? 0u : 1u;
        v[13] = (uint32_t)ta_fls(colmask); v[14] = (uint32_t)ta_fls(mrows); v[15] = (rows >> BLK_M) ? 1u : 0u;
        return true;
    }

    // seen from the voxels of slot i towards slot j: wall18 voxels and +f / +m / +s faces whose lower voxel is i and upper
    // voxel j; false: nothing (or an empty slot)
    TA_HD bool pair_counts(int i, int j, uint32_t& w18, uint32_t& ff, uint32_t& fm, uint32_t& fsl) const {
        constexpr int ROWBITS = Blk<T>::ROWBITS;
        w18 = ff = fm = fsl = 0u;
        if (i >= k || j >= k || i == j) return false;
        const u64 c0 = M1[i] & cv0, c1 = M2[i] & cv1;
        w18 = ta_popc64(c0 & D0[j]) + ta_popc64(c1 & D1[j]);
        ff = ta_popc64(c0 & (M1[j] >> 1)) + ta_popc64(c1 & (M2[j] >> 1));
        fm = ta_popc64(c0 & (M1[j] >> ROWBITS)) + ta_popc64(c1 & (M2[j] >> ROWBITS));
        fsl = ta_popc64(c0 & M2[j]) + ta_popc64(c1 & M3[j]);
        return (w18 | ff | fm | fsl) != 0u;
    }
};

// ---- level formulation ----------------------------------------------------------------------------------------------
// The discovery loop of BlockSlots costs every lane of a warp as many mask builds as the most crowded block of the warp
// holds labels (measured warp maximum 3.98 against a lane mean of 1.85, tools/simt_stats.py).  The level formulation
// makes the label count a property of a LIST instead of a lane:
//
//   level 1   min / max label over the window (one pass over its 24 rows).  min == max: the window is one label, the
//             block contributes closed-form moments and nothing else.  Otherwise the block goes to list 2 with the two
//             labels (both ARE labels of the window).
//   level N   (N = 2, 3, ...) for the blocks of list N, N labels known: ONE fused pass over the rows builds the N masks
//             (row loads shared between the labels).  Emitted: for N = 2 the moments of both labels and their pair; for
//             N > 2 the moments of the newest label and its pairs with the N - 1 older ones -- exactly what the earlier
//             levels could not know.  Window positions covered by none of the N labels name label N + 1: list N + 1.
//   fallback  blocks still uncovered after the last level take the per-voxel path for everything that involves a label
//             outside their known set S (moments of voxels not in S, pairs with at least one side not in S); all
//             S-internal contributions were emitted by the levels.
//
// Every list is processed by full warps of blocks with the same label count, so the cost follows the mean of the label
// count distribution, not the warp maximum.  tests/host/block_level_check.cu runs this on the CPU against a direct pass.
// Block geometry of the level formulation: 8 x 4 x 2 voxels for BOTH label widths (uint32: two 16-byte segments per block
// row), so that the window planes, the dilation, the moments table and the pair counts are the same code and a uint32
// voxel costs as little of them as a uint16 one.
template <typename T> struct LvBlk {
    static constexpr int BW = 8;                               // block width in voxels
    static constexpr int BSEGS = BW / Vox<T>::SEG;             // segments per block row: 1 (uint16) or 2 (uint32)
    static constexpr int NFB = NFS / BSEGS;                    // blocks per brick row
    static constexpr int NBLK = NFB * (BM / BLK_M) * (BS / BLK_S);   // blocks per brick: 256 or 128
    static constexpr int ROWBITS = BW + 2;
    static constexpr u64 PLANE_ALL = (1ull << (ROWBITS * (BLK_M + 2))) - 1ull;
    static constexpr uint32_t LANES = (1u << BW) - 1u;
};
TA_HD uint32_t ta_vmaxu2(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __vmaxu2(a, b);
#else
    const uint32_t l = (a & 0xFFFFu) > (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
    const uint32_t h = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
    return l | (h << 16);
#endif
}
// keeps the row loads of one half plane from being hoisted above the arithmetic of the previous one (register pressure)
TA_HD void block_sched_fence() {
#ifdef __CUDA_ARCH__
    TA_PTX("" ::: "memory");
#endif
}

// The level formulation reads a SHIFTED tile (block_stage_tile<T, 1>: brick column f at tile element SEG + 1 + f): the
// window row of a block (columns -1 .. 8 relative to the block) starts on a 16-byte vector.  uint16: one vector (columns
// -1 .. 6) + the first word of the next (columns 7, 8); uint32: two vectors + the first two words of the third.
//
// Smallest and largest label of the window whose first row (m0 - 1, s0 - 1) starts at vector t0.
template <typename T> TA_HD void block_window_minmax(const uint4* tile, int t0, uint32_t& lo, uint32_t& hi);
template <> TA_HD void block_window_minmax<uint16_t>(const uint4* tile, int t0, uint32_t& lo, uint32_t& hi) {
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
    for (int p = 0; p < BLK_S + 2; ++p) {
#pragma unroll
        for (int r = 0; r < BLK_M + 2; ++r) {
            const int t = t0 + p * PLANEV + r * ROWV;
            const uint4 c = tile[t];
            const uint32_t w = reinterpret_cast<const uint32_t*>(tile + t + 1)[0];
            mn = ta_vminu2(ta_vminu2(ta_vminu2(mn, c.x), ta_vminu2(c.y, c.z)), ta_vminu2(c.w, w));
            mx = ta_vmaxu2(ta_vmaxu2(ta_vmaxu2(mx, c.x), ta_vmaxu2(c.y, c.z)), ta_vmaxu2(c.w, w));
        }
        block_sched_fence();
    }
    lo = (mn & 0xFFFFu) < (mn >> 16) ? (mn & 0xFFFFu) : (mn >> 16);
    hi = (mx & 0xFFFFu) > (mx >> 16) ? (mx & 0xFFFFu) : (mx >> 16);
}
template <> TA_HD void block_window_minmax<uint32_t>(const uint4* tile, int t0, uint32_t& lo, uint32_t& hi) {
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
    for (int p = 0; p < BLK_S + 2; ++p) {
#pragma unroll
        for (int r = 0; r < BLK_M + 2; ++r) {
            const int t = t0 + p * PLANEV + r * ROWV;
            const uint4 c = tile[t], d = tile[t + 1];
            const uint2 e = reinterpret_cast<const uint2*>(tile + t + 2)[0];
            const uint32_t v[10] = {c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w, e.x, e.y};
#pragma unroll
            for (int i = 0; i < 10; ++i) { mn = mn < v[i] ? mn : v[i]; mx = mx > v[i] ? mx : v[i]; }
        }
        block_sched_fence();
    }
    lo = mn; hi = mx;
}

// NOT-equal bits of one window row (starting at vector t) against N labels at once: out[i] bit x (0 .. 9, column x - 1
// relative to the block) is set where the voxel differs from label i.  The row is loaded once for all labels.
// uint16: five words of two lanes each; xor, VIMNMX against 1 -> one bit per lane at bits 0 and 16; the words are summed
// at 2-bit steps (even columns at bits 0, 2, .. 8, odd columns at 16, 18, .. 24) and folded.
TA_HD uint32_t block_fold10(uint32_t tt) { return (tt & 0x155u) | ((tt >> 15) & 0x2AAu); }
template <typename T, int N> struct BlockRowNeq;
template <int N> struct BlockRowNeq<uint16_t, N> {
    static TA_HD void run(const uint4* tile, int t, const uint32_t* L, uint32_t* out) {
        const uint4 c = tile[t];
        const uint32_t w = reinterpret_cast<const uint32_t*>(tile + t + 1)[0], one = 0x00010001u;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t pat = L[i] * 0x00010001u;
            const uint32_t tt = ta_vminu2(c.x ^ pat, one) + (ta_vminu2(c.y ^ pat, one) << 2) + (ta_vminu2(c.z ^ pat, one) << 4) +
                                (ta_vminu2(c.w ^ pat, one) << 6) + (ta_vminu2(w ^ pat, one) << 8);
            out[i] = block_fold10(tt);
        }
    }
};
// The same for level 2, whose two labels are the SMALLEST and the LARGEST label of the window (block_window_minmax over
// exactly these lanes): every lane of x - A and of B - x is non-negative, so the packed 32-bit subtractions never borrow
// across the 16-bit lanes and are exact per lane.  A subtraction may run on the FMA pipe (IMAD), the xor of the general
// form cannot -- the mask build is bound by the ALU pipe.
struct BlockRowNeqMinMax16 {
    static TA_HD void run(const uint4* tile, int t, const uint32_t* L, uint32_t* out) {
        const uint4 c = tile[t];
        const uint32_t w = reinterpret_cast<const uint32_t*>(tile + t + 1)[0], one = 0x00010001u;
        const uint32_t pa = L[0] * 0x00010001u, pb = L[1] * 0x00010001u;
        out[0] = block_fold10(ta_vminu2(c.x - pa, one) + (ta_vminu2(c.y - pa, one) << 2) + (ta_vminu2(c.z - pa, one) << 4) +
                              (ta_vminu2(c.w - pa, one) << 6) + (ta_vminu2(w - pa, one) << 8));
        out[1] = block_fold10(ta_vminu2(pb - c.x, one) + (ta_vminu2(pb - c.y, one) << 2) + (ta_vminu2(pb - c.z, one) << 4) +
                              (ta_vminu2(pb - c.w, one) << 6) + (ta_vminu2(pb - w, one) << 8));
    }
};
// uint32: ten words of one lane each; min(x ^ L, 1) is the lane's bit, summed at 1-bit steps (one compare-free ALU op and
// one shift-add per lane instead of compare + select + or).  Level 2 (MINMAX: L[0] / L[1] = smallest / largest label of
// the window) subtracts instead of xoring, which may run on the FMA pipe.
TA_HD uint32_t ta_minu(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;                        // spelled in PTX: the compiler turns min(x, 1) back into compare + select
    asm("min.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    return a < b ? a : b;
#endif
}
template <int N> struct BlockRowNeq<uint32_t, N> {
    static TA_HD void run(const uint4* tile, int t, const uint32_t* L, uint32_t* out) {
        const uint4 c = tile[t], d = tile[t + 1];
        const uint2 e = reinterpret_cast<const uint2*>(tile + t + 2)[0];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t l = L[i];
            out[i] = ta_minu(c.x ^ l, 1u) + (ta_minu(c.y ^ l, 1u) << 1) + (ta_minu(c.z ^ l, 1u) << 2) + (ta_minu(c.w ^ l, 1u) << 3) +
                     (ta_minu(d.x ^ l, 1u) << 4) + (ta_minu(d.y ^ l, 1u) << 5) + (ta_minu(d.z ^ l, 1u) << 6) +
                     (ta_minu(d.w ^ l, 1u) << 7) + (ta_minu(e.x ^ l, 1u) << 8) + (ta_minu(e.y ^ l, 1u) << 9);
        }
    }
};
struct BlockRowNeqMinMax32 {
    static TA_HD void run(const uint4* tile, int t, const uint32_t* L, uint32_t* out) {
        const uint4 c = tile[t], d = tile[t + 1];
        const uint2 e = reinterpret_cast<const uint2*>(tile + t + 2)[0];
        const uint32_t a = L[0], b = L[1];
        out[0] = ta_minu(c.x - a, 1u) + (ta_minu(c.y - a, 1u) << 1) + (ta_minu(c.z - a, 1u) << 2) + (ta_minu(c.w - a, 1u) << 3) +
                 (ta_minu(d.x - a, 1u) << 4) + (ta_minu(d.y - a, 1u) << 5) + (ta_minu(d.z - a, 1u) << 6) +
                 (ta_minu(d.w - a, 1u) << 7) + (ta_minu(e.x - a, 1u) << 8) + (ta_minu(e.y - a, 1u) << 9);
        out[1] = ta_minu(b - c.x, 1u) + (ta_minu(b - c.y, 1u) << 1) + (ta_minu(b - c.z, 1u) << 2) + (ta_minu(b - c.w, 1u) << 3) +
                 (ta_minu(b - d.x, 1u) << 4) + (ta_minu(b - d.y, 1u) << 5) + (ta_minu(b - d.z, 1u) << 6) +
                 (ta_minu(b - d.w, 1u) << 7) + (ta_minu(b - e.x, 1u) << 8) + (ta_minu(b - e.y, 1u) << 9);
    }
};

// Packed row moments for the table form of the moments: n [0..9] | sum x [10..20] | sum x^2 [21..31] of the set bits of
// a byte.  The field widths hold every weighted sum over the 8 rows of a block (weights r <= 3, r^2 <= 9).
TA_HD uint32_t block_byte_moments_packed(uint32_t b) {
    const uint32_t t = block_byte_moments(b);
    return (t & 0xFFu) | (((t >> 8) & 0xFFu) << 10) | ((t >> 16) << 21);
}

// Closed-form sums and box of a one-label block with a x b x c voxels inside the volume (block-local coordinates).
TA_HD void block_uniform_moments(uint32_t a, uint32_t b, uint32_t c, uint32_t v[16]) {
    const uint32_t ta_ = a * (a - 1) / 2, tb = b * (b - 1) / 2, tc = c * (c - 1) / 2;
    const uint32_t qa = (a - 1) * a * (2 * a - 1) / 6, qb = (b - 1) * b * (2 * b - 1) / 6, qc = (c - 1) * c * (2 * c - 1) / 6;
    v[0] = a * b * c; v[1] = b * c * ta_; v[2] = a * c * tb; v[3] = a * b * tc;
    v[4] = b * c * qa; v[5] = c * ta_ * tb; v[6] = b * ta_ * tc; v[7] = a * c * qb; v[8] = a * tb * tc; v[9] = a * b * qc;
    v[10] = 0u; v[11] = 0u; v[12] = 0u; v[13] = a - 1u; v[14] = b - 1u; v[15] = c - 1u;
}

// Up to CAP known labels of one block: window masks, dilations, coverage.  build<N0>() fills slots 0 .. N0 - 1 in one
// fused pass over the window rows; extend<I>() adds slot I for one more label (the rare labels 4, 5, ... of a block).
constexpr int LV_STATE_WORDS = 14;
template <typename T, int CAP> struct BlockLevel {
    uint32_t lab[CAP];
    u64 M1[CAP], M2[CAP], M3[CAP], D0[CAP], D1[CAP];
    u64 cv0, cv1;             // centre voxels inside the volume, planes 1 and 2 of the window
    u64 R0, R1, R2, R3;       // window positions covered by none of the labels so far

    TA_HD void clear() {
        cv0 = cv1 = 0ull; R0 = R1 = R2 = R3 = 0ull;
#pragma unroll
        for (int i = 0; i < CAP; ++i) { lab[i] = 0u; M1[i] = M2[i] = M3[i] = D0[i] = D1[i] = 0ull; }
    }
    template <int I> TA_HD void clear_slot() { lab[I] = 0u; M1[I] = M2[I] = M3[I] = D0[I] = D1[I] = 0ull; }

    // NOT-equal planes of N labels, one fused pass: neq[i][p]
    template <int N, bool MINMAX = false>
    static TA_HD void neq_planes(const uint4* tile, int t0, const uint32_t* L, u64 neq[][BLK_S + 2]) {
        constexpr int ROWBITS = LvBlk<T>::ROWBITS, HALF = (BLK_M + 2) / 2;
#pragma unroll
        for (int p = 0; p < BLK_S + 2; ++p) {
            uint32_t half[N][2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int i = 0; i < N; ++i) half[i][h] = 0u;
#pragma unroll
                for (int r = HALF - 1; r >= 0; --r) {                  // descending: acc = (acc << ROWBITS) + row
                    uint32_t row[N];
                    if constexpr (MINMAX && N == 2 && sizeof(T) == 2)
                        BlockRowNeqMinMax16::run(tile, t0 + p * PLANEV + (h * HALF + r) * ROWV, L, row);
                    else if constexpr (MINMAX && N == 2 && sizeof(T) == 4)
                        BlockRowNeqMinMax32::run(tile, t0 + p * PLANEV + (h * HALF + r) * ROWV, L, row);
                    else
                        BlockRowNeq<T, N>::run(tile, t0 + p * PLANEV + (h * HALF + r) * ROWV, L, row);
#pragma unroll
                    for (int i = 0; i < N; ++i) half[i][h] = (half[i][h] << ROWBITS) + row[i];
                }
                block_sched_fence();
            }
#pragma unroll
            for (int i = 0; i < N; ++i) neq[i][p] = (u64)half[i][0] | ((u64)half[i][1] << (ROWBITS * HALF));
        }
    }
    // first uncovered position -> its label
    TA_HD uint32_t first_uncovered(const uint4* tile, int t0) const {
        constexpr int SEG = Vox<T>::SEG, ROWBITS = LvBlk<T>::ROWBITS;
        const int p = R0 ? 0 : R1 ? 1 : R2 ? 2 : 3;
        const u64 rp = R0 ? R0 : R1 ? R1 : R2 ? R2 : R3;
        const int bit = ta_ffs64(rp) - 1, r = bit / ROWBITS, x = bit % ROWBITS;
        return reinterpret_cast<const T*>(tile)[(size_t)(t0 + p * PLANEV + r * ROWV) * SEG + x];      // shifted tile
    }
    template <int I> TA_HD void set_slot(uint32_t L, const u64 neq[BLK_S + 2]) {
        constexpr u64 ALL = LvBlk<T>::PLANE_ALL;
        R0 &= neq[0]; R1 &= neq[1]; R2 &= neq[2]; R3 &= neq[3];
        u64 m[4], d[2];
#pragma unroll
        for (int p = 0; p < 4; ++p) m[p] = ~neq[p] & ALL;
        block_dilate18_rb<LvBlk<T>::ROWBITS>(m, d);
        lab[I] = L; M1[I] = m[1]; M2[I] = m[2]; M3[I] = m[3]; D0[I] = d[0]; D1[I] = d[1];
    }

    // Slots 0 and 1 and the uncovered positions, parked between level 2 and level 3 of a block (field-major: lane `pos`
    // of a warp reads and writes consecutive 64-bit words).  LV_STATE_WORDS words per block.
    TA_HD void store_state2(u64* st, int pos, int cap) const {
        st[0 * cap + pos] = M1[0]; st[1 * cap + pos] = M2[0]; st[2 * cap + pos] = M3[0]; st[3 * cap + pos] = D0[0];
        st[4 * cap + pos] = D1[0]; st[5 * cap + pos] = M1[1]; st[6 * cap + pos] = M2[1]; st[7 * cap + pos] = M3[1];
        st[8 * cap + pos] = D0[1]; st[9 * cap + pos] = D1[1]; st[10 * cap + pos] = R0; st[11 * cap + pos] = R1;
        st[12 * cap + pos] = R2; st[13 * cap + pos] = R3;
    }
    TA_HD void load_state2(const u64* st, int pos, int cap, uint32_t L0, uint32_t L1, int nvf, int nvm, int nvs) {
        static_assert(CAP >= 2, "two slots");
        lab[0] = L0; lab[1] = L1;
        M1[0] = st[0 * cap + pos]; M2[0] = st[1 * cap + pos]; M3[0] = st[2 * cap + pos]; D0[0] = st[3 * cap + pos];
        D1[0] = st[4 * cap + pos]; M1[1] = st[5 * cap + pos]; M2[1] = st[6 * cap + pos]; M3[1] = st[7 * cap + pos];
        D0[1] = st[8 * cap + pos]; D1[1] = st[9 * cap + pos]; R0 = st[10 * cap + pos]; R1 = st[11 * cap + pos];
        R2 = st[12 * cap + pos]; R3 = st[13 * cap + pos];
        set_centre(nvf, nvm, nvs);
    }
    TA_HD void set_centre(int nvf, int nvm, int nvs) {
        constexpr int ROWBITS = LvBlk<T>::ROWBITS;
        u64 cv = 0ull;
#pragma unroll
        for (int r = 1; r <= BLK_M; ++r)
            if (r <= nvm) cv |= (u64)(((1u << nvf) - 1u) << 1) << (ROWBITS * r);
        cv0 = nvs >= 1 ? cv : 0ull;
        cv1 = nvs >= 2 ? cv : 0ull;
    }

    // L[0 .. N0 - 1]: distinct labels -> slots 0 .. N0 - 1.  true: they cover the window; false: `next` = a label of the
    // window that is none of them (the label at the first uncovered position).
    // MINMAX (N0 == 2 only): L[0] / L[1] are the smallest / largest label of the window, as block_window_minmax returns them.
    template <int N0, bool MINMAX = false>
    TA_HD bool build(const uint4* tile, int fs, int m0, int s0, int nvf, int nvm, int nvs, const uint32_t* L, uint32_t& next) {
        static_assert(N0 <= CAP, "more labels than slots");
        constexpr int ROWBITS = LvBlk<T>::ROWBITS;
        constexpr u64 ALL = LvBlk<T>::PLANE_ALL;
        const int t0 = s0 * PLANEV + m0 * ROWV + (fs + 1);
        u64 neq[N0][BLK_S + 2];
        neq_planes<N0, MINMAX>(tile, t0, L, neq);
        R0 = R1 = R2 = R3 = ALL;
        set_slots<N0, 0>(L, neq);
        set_centre(nvf, nvm, nvs);
        if (!(R0 | R1 | R2 | R3)) return true;
        next = first_uncovered(tile, t0);
        return false;
    }
    template <int N0, int I> TA_HD void set_slots(const uint32_t* L, const u64 neq[][BLK_S + 2]) {
        if constexpr (I < N0) {
            set_slot<I>(L[I], neq[I]);
            set_slots<N0, I + 1>(L, neq);
        }
    }

    // one more label (not among slots 0 .. I - 1) -> slot I; same return as build
    template <int I>
    TA_HD bool extend(const uint4* tile, int fs, int m0, int s0, uint32_t Lnew, uint32_t& next) {
        static_assert(I < CAP, "more labels than slots");
        const int t0 = s0 * PLANEV + m0 * ROWV + (fs + 1);
        u64 neq[1][BLK_S + 2];
        neq_planes<1>(tile, t0, &Lnew, neq);
        set_slot<I>(Lnew, neq[0]);
        if (!(R0 | R1 | R2 | R3)) return true;
        next = first_uncovered(tile, t0);
        return false;
    }

    // moments and box of label i over its centre voxels, block-local coordinates; `tab` = block_byte_moments_packed of
    // every byte (256 entries, shared memory in the kernel).  false: no centre voxel.
    TA_HD bool label_moments(int i, const uint32_t* tab, uint32_t v[16]) const {
        constexpr int ROWBITS = LvBlk<T>::ROWBITS;
        const u64 c0 = M1[i] & cv0, c1 = M2[i] & cv1;
        if (!(c0 | c1)) return false;
        uint32_t a0 = 0u, a1 = 0u, a2 = 0u, ap = 0u, apm = 0u, colmask = 0u, rows = 0u;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int r = 0; r < BLK_M; ++r) {
                const uint32_t b = (uint32_t)((p ? c1 : c0) >> (ROWBITS * (r + 1) + 1)) & LvBlk<T>::LANES;
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
        v[10] = (uint32_t)ta_ffs(colmask) - 1u; v[11] = (uint32_t)ta_ffs(mrows) - 1u;
        v[12] = (rows & ((1u << BLK_M) - 1u)) ? 0u : 1u;
        v[13] = (uint32_t)ta_fls(colmask); v[14] = (uint32_t)ta_fls(mrows); v[15] = (rows >> BLK_M) ? 1u : 0u;
        return true;
    }

    // both directions of the unordered pair (i, j) as packed increments of the per-brick pair table
    // ([w18|f0] [f1|f2] [f3|f4] [f5|-]: a face goes to slot 2k when its lower-index voxel carries the smaller label)
    TA_HD bool pair_increments(int i, int j, bool do_p6, bool do_w18, uint32_t inc[4]) const {
        constexpr int ROWBITS = LvBlk<T>::ROWBITS;
        const u64 ci0 = M1[i] & cv0, ci1 = M2[i] & cv1, cj0 = M1[j] & cv0, cj1 = M2[j] & cv1;
        uint32_t w18 = 0u, fi[3] = {0u, 0u, 0u}, fj[3] = {0u, 0u, 0u};
        if (do_w18)
            w18 = ta_popc64(ci0 & D0[j]) + ta_popc64(ci1 & D1[j]) + ta_popc64(cj0 & D0[i]) + ta_popc64(cj1 & D1[i]);
        if (do_p6) {
            fi[0] = ta_popc64(ci0 & (M1[j] >> 1)) + ta_popc64(ci1 & (M2[j] >> 1));
            fi[1] = ta_popc64(ci0 & (M1[j] >> ROWBITS)) + ta_popc64(ci1 & (M2[j] >> ROWBITS));
            fi[2] = ta_popc64(ci0 & M2[j]) + ta_popc64(ci1 & M3[j]);
            fj[0] = ta_popc64(cj0 & (M1[i] >> 1)) + ta_popc64(cj1 & (M2[i] >> 1));
            fj[1] = ta_popc64(cj0 & (M1[i] >> ROWBITS)) + ta_popc64(cj1 & (M2[i] >> ROWBITS));
            fj[2] = ta_popc64(cj0 & M2[i]) + ta_popc64(cj1 & M3[i]);
        }
        const bool ilo = lab[i] < lab[j];            // faces seen from the smaller label go to the even slots
        const uint32_t e0 = ilo ? fi[0] : fj[0], o0 = ilo ? fj[0] : fi[0];
        const uint32_t e1 = ilo ? fi[1] : fj[1], o1 = ilo ? fj[1] : fi[1];
        const uint32_t e2 = ilo ? fi[2] : fj[2], o2 = ilo ? fj[2] : fi[2];
        inc[0] = w18 | (e0 << 16); inc[1] = o0 | (e1 << 16); inc[2] = o1 | (e2 << 16); inc[3] = o2;
        return (inc[0] | inc[1] | inc[2] | inc[3]) != 0u;
    }
};

// block_features on top of BlockSlots: same callbacks and results as the reference form above.
template <typename T, int MAXLAB, typename OnLabel, typename OnPair>
TA_HD bool block_features_reg(const uint4* tile, int fs, int m0, int s0, int nvf, int nvm, int nvs, OnLabel&& on_label,
                              OnPair&& on_pair) {
    BlockSlots<T, MAXLAB> b;
    if (!b.discover(tile, fs, m0, s0, nvf, nvm, nvs)) return false;
#pragma unroll
    for (int i = 0; i < MAXLAB; ++i) {
        uint32_t v[16];
        if (!b.label_moments(i, v)) continue;
        on_label(b.lab[i], v);
#pragma unroll
        for (int j = 0; j < MAXLAB; ++j) {
            uint32_t w18, ff, fm, fsl;
            if (b.pair_counts(i, j, w18, ff, fm, fsl)) on_pair(b.lab[i], b.lab[j], w18, ff, fm, fsl);
        }
    }
    return true;
}

}  // namespace ta

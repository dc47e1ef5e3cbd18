// CPU check of the LEVEL formulation of the block pass (csrc/ta_block.cuh: block_window_minmax, BlockLevel<T, N>; test
// infrastructure, no GPU needed).  A volume of several bricks is tiled as the scan kernel tiles it; every 8 x 4 x 2 block
// (uint32: two 16-byte segments wide) goes through level 1 (window min / max, closed-form moments when they agree), level 2 (fused masks of both
// labels), level 3 (fused masks of three labels) with its extension steps up to MAXL labels (one more mask per step;
// emitted: what the step adds) and, when labels are still uncovered, the per-voxel fallback restricted to contributions
// with a label outside the known set.  The global tables must equal a direct pass.
// Usage: block_level_check <seed> [maxl] ; exit code 0 = every case equal.
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>
#include "../../tissue_analysis_b200/csrc/ta_block.cuh"

using namespace ta;

struct LabelRow { u64 v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long bmin[3] = {1L << 40, 1L << 40, 1L << 40}, bmax[3] = {-1, -1, -1};
    bool operator==(const LabelRow& o) const { return std::equal(v, v + 10, o.v) && std::equal(bmin, bmin + 3, o.bmin) && std::equal(bmax, bmax + 3, o.bmax); } };
typedef std::map<uint32_t, LabelRow> LabelTab;
typedef std::map<std::pair<uint32_t, uint32_t>, std::array<u64, 7>> PairTab;     // faces[6], wall18

struct Vol {
    int nf, nm, ns; std::vector<uint32_t> d;
    uint32_t at(int f, int m, int s) const {
        f = std::min(std::max(f, 0), nf - 1); m = std::min(std::max(m, 0), nm - 1); s = std::min(std::max(s, 0), ns - 1);
        return d[((size_t)s * nm + m) * nf + f];
    }
};

// One voxel of a direct pass.  known / nknown: contributions whose labels ALL lie in known[] are skipped (the levels
// emitted them); nknown = 0 is the plain direct pass.
static void add_voxel(const Vol& V, int f, int m, int s, long slow_offset, LabelTab& lt, PairTab& pt, const uint32_t* known = nullptr,
                      int nknown = 0) {
    auto in_s = [&](uint32_t x) { for (int i = 0; i < nknown; ++i) if (known[i] == x) return true; return false; };
    const uint32_t a = V.at(f, m, s);
    const bool a_in = in_s(a);
    if (!a_in) {
        LabelRow& r = lt[a];
        const u64 F = f, M = m, S = (u64)(s + slow_offset);
        r.v[0] += 1; r.v[1] += F; r.v[2] += M; r.v[3] += S; r.v[4] += F * F; r.v[5] += F * M; r.v[6] += F * S; r.v[7] += M * M;
        r.v[8] += M * S; r.v[9] += S * S;
        const long c[3] = {(long)F, (long)M, (long)S};
        for (int k = 0; k < 3; ++k) { r.bmin[k] = std::min(r.bmin[k], c[k]); r.bmax[k] = std::max(r.bmax[k], c[k]); }
    }
    std::map<uint32_t, int> seen;
    for (int z = -1; z <= 1; ++z) for (int y = -1; y <= 1; ++y) for (int x = -1; x <= 1; ++x) {
        const int l1 = abs(z) + abs(y) + abs(x);
        if (l1 < 1 || l1 > 2) continue;
        const uint32_t b = V.at(f + x, m + y, s + z);
        if (b != a && !(a_in && in_s(b))) seen[b] = 1;
    }
    for (auto& kv : seen) pt[{std::min(a, kv.first), std::max(a, kv.first)}][6] += 1;
    const uint32_t nb[3] = {V.at(f + 1, m, s), V.at(f, m + 1, s), V.at(f, m, s + 1)};
    for (int k = 0; k < 3; ++k)
        if (nb[k] != a && !(a_in && in_s(nb[k]))) pt[{std::min(a, nb[k]), std::max(a, nb[k])}][2 * k + (a < nb[k] ? 0 : 1)] += 1;
}

struct Sink {
    LabelTab* lt; PairTab* pt;
    uint32_t bf, bm, bs;          // block origin inside the brick
    u64 F0, M0, S0;               // brick origin in global coordinates
    void label(uint32_t L, const uint32_t vin[16]) const {
        uint32_t v[16];
        for (int i = 0; i < 16; ++i) v[i] = vin[i];
        block_shift_moments(v, bf, bm, bs);                       // block -> brick, 32 bits as the kernel does
        LabelRow& r = (*lt)[L];
        const u64 n = v[0], sf = v[1], sm = v[2], ss = v[3];      // brick -> global: label_to_global of ta_scan.cuh
        r.v[0] += n; r.v[1] += n * F0 + sf; r.v[2] += n * M0 + sm; r.v[3] += n * S0 + ss;
        r.v[4] += n * F0 * F0 + 2 * F0 * sf + v[4];
        r.v[5] += n * F0 * M0 + F0 * sm + M0 * sf + v[5];
        r.v[6] += n * F0 * S0 + F0 * ss + S0 * sf + v[6];
        r.v[7] += n * M0 * M0 + 2 * M0 * sm + v[7];
        r.v[8] += n * M0 * S0 + M0 * ss + S0 * sm + v[8];
        r.v[9] += n * S0 * S0 + 2 * S0 * ss + v[9];
        const long lo[3] = {(long)(F0 + v[10]), (long)(M0 + v[11]), (long)(S0 + v[12])};
        const long hi[3] = {(long)(F0 + v[13]), (long)(M0 + v[14]), (long)(S0 + v[15])};
        for (int k = 0; k < 3; ++k) { r.bmin[k] = std::min(r.bmin[k], lo[k]); r.bmax[k] = std::max(r.bmax[k], hi[k]); }
    }
    void pair(uint32_t a, uint32_t b, const uint32_t inc[4]) const {      // packed increments: [w18|f0] [f1|f2] [f3|f4] [f5|-]
        auto& r = (*pt)[{std::min(a, b), std::max(a, b)}];
        r[6] += inc[0] & 0xFFFFu; r[0] += inc[0] >> 16; r[1] += inc[1] & 0xFFFFu; r[2] += inc[1] >> 16;
        r[3] += inc[2] & 0xFFFFu; r[4] += inc[2] >> 16; r[5] += inc[3] & 0xFFFFu;
    }
};

// Extension steps I .. MAXL - 1 of the level-3 pass: one more label per step (slot I), its moments and its pairs with
// every older slot.  Returns the number of known labels when the window is covered, -MAXL when it is still not.
template <typename T, int I, int MAXL>
static int run_extend(BlockLevel<T, (MAXL > 3 ? MAXL : 3)>& b, const uint4* tile, int fs, int m0, int s0, uint32_t* L, const uint32_t* tab,
                      const Sink& sink) {
    if constexpr (I < MAXL) {
        uint32_t next = 0u, v[16], inc[4];
        const bool covered = b.template extend<I>(tile, fs, m0, s0, L[I], next);
        if (b.label_moments(I, tab, v)) sink.label(b.lab[I], v);
        for (int j = 0; j < I; ++j) if (b.pair_increments(I, j, true, true, inc)) sink.pair(b.lab[I], b.lab[j], inc);
        if (covered) return I + 1;
        L[I + 1] = next;
        return run_extend<T, I + 1, MAXL>(b, tile, fs, m0, s0, L, tab, sink);
    } else {
        return -MAXL;
    }
}

// Level 2, then (labels left) level 3 with its extension steps, as the kernel runs them.  Returns the number of known
// labels when the window is covered, -(number known) when labels are still uncovered after the last step.
template <typename T, int MAXL>
static int run_levels(const uint4* tile, int fs, int m0, int s0, int nvf, int nvm, int nvs, uint32_t* L, const uint32_t* tab,
                      const Sink& sink) {
    uint32_t next = 0u, v[16], inc[4];
    {
        BlockLevel<T, 2> b;
        const bool covered = b.template build<2, true>(tile, fs, m0, s0, nvf, nvm, nvs, L, next);      // L[0], L[1] = min, max
        for (int i = 0; i < 2; ++i) if (b.label_moments(i, tab, v)) sink.label(b.lab[i], v);
        if (b.pair_increments(0, 1, true, true, inc)) sink.pair(b.lab[0], b.lab[1], inc);
        if (covered) return 2;
        L[2] = next;
    }
    if constexpr (MAXL < 3) return -2;
    else {
        BlockLevel<T, (MAXL > 3 ? MAXL : 3)> b;
        const bool covered = b.template build<3>(tile, fs, m0, s0, nvf, nvm, nvs, L, next);
        if (b.label_moments(2, tab, v)) sink.label(b.lab[2], v);
        for (int j = 0; j < 2; ++j) if (b.pair_increments(2, j, true, true, inc)) sink.pair(b.lab[2], b.lab[j], inc);
        if (covered) return 3;
        L[3] = next;
        return run_extend<T, 3, MAXL>(b, tile, fs, m0, s0, L, tab, sink);
    }
}

template <typename T, int MAXL>
static int run_case(int nf, int nm, int nbuf, int own_lo, int own_hi, long slow_offset, int nlabels, int mode, unsigned seed,
                    long* hist) {
    constexpr int SEG = Vox<T>::SEG, ROWE = ROWV * SEG, BF = NFS * SEG;
    std::mt19937 rng(seed);
    Vol V{nf, nm, nbuf, std::vector<uint32_t>((size_t)nf * nm * nbuf)};
    std::vector<uint32_t> names(nlabels);
    for (auto& n : names) n = (sizeof(T) == 2 ? rng() % 65536u : rng() % 0xFFFFFF00u);      // 0 and 0xFFFF included
    if (mode == 0) {
        for (auto& v : V.d) v = names[rng() % nlabels];
    } else {
        std::vector<int> sx(nlabels), sy(nlabels), sz(nlabels);
        for (int k = 0; k < nlabels; ++k) { sx[k] = rng() % nf; sy[k] = rng() % nm; sz[k] = rng() % nbuf; }
        for (int s = 0; s < nbuf; ++s) for (int m = 0; m < nm; ++m) for (int f = 0; f < nf; ++f) {
            long best = 1L << 60; int bk = 0;
            for (int k = 0; k < nlabels; ++k) {
                long d = (long)(f - sx[k]) * (f - sx[k]) + (long)(m - sy[k]) * (m - sy[k]) * 2 + (long)(s - sz[k]) * (s - sz[k]) * 3;
                if (d < best) { best = d; bk = k; }
            }
            V.d[((size_t)s * nm + m) * nf + f] = names[bk];
        }
    }
    uint32_t tab[256];
    for (uint32_t b = 0; b < 256; ++b) tab[b] = block_byte_moments_packed(b);
    LabelTab gotL, refL; PairTab gotP, refP;
    for (int s = own_lo; s < own_hi; ++s) for (int m = 0; m < nm; ++m) for (int f = 0; f < nf; ++f)
        add_voxel(V, f, m, s, slow_offset, refL, refP);
    std::vector<uint4> tile(TILE_SEGS);
    T* tl = reinterpret_cast<T*>(tile.data());
    for (int S0 = own_lo; S0 < own_hi; S0 += BS) for (int M0 = 0; M0 < nm; M0 += BM) for (int F0 = 0; F0 < nf; F0 += BF) {
        for (int r = 0; r < TILE_ROWS; ++r) {
            const int m = r % (BM + 2) - 1, s = r / (BM + 2) - 1;
            for (int e = 0; e < ROWE; ++e) tl[(size_t)r * ROWE + e] = (T)V.at(F0 + e - SEG - 1, M0 + m, S0 + s);      // the SHIFTED tile
        }
        for (int s0 = 0; s0 < BS && S0 + s0 < own_hi; s0 += BLK_S) for (int m0 = 0; m0 < BM && M0 + m0 < nm; m0 += BLK_M)
            for (int fs = 0; fs < NFS && F0 + fs * SEG < nf; fs += LvBlk<T>::BSEGS) {
                const int nvf = std::min((int)LvBlk<T>::BW, nf - F0 - fs * SEG), nvm = std::min(BLK_M, nm - M0 - m0),
                          nvs = std::min(BLK_S, own_hi - S0 - s0);
                const Sink sink{&gotL, &gotP, (uint32_t)(fs * SEG), (uint32_t)m0, (uint32_t)s0, (u64)F0, (u64)M0, (u64)(S0 + slow_offset)};
                const int t0 = s0 * PLANEV + m0 * ROWV + (fs + 1);
                uint32_t L[MAXL + 1];
                block_window_minmax<T>(tile.data(), t0, L[0], L[1]);
                if (L[0] == L[1]) {
                    uint32_t v[16];
                    block_uniform_moments((uint32_t)nvf, (uint32_t)nvm, (uint32_t)nvs, v);
                    sink.label(L[0], v);
                    ++hist[1];
                    continue;
                }
                const int k = run_levels<T, MAXL>(tile.data(), fs, m0, s0, nvf, nvm, nvs, L, tab, sink);
                if (k > 0) { ++hist[k]; continue; }
                ++hist[0];            // uncovered after the last level: everything that involves a label outside L[0 .. MAXL - 1]
                for (int ds = 0; ds < nvs; ++ds) for (int dm = 0; dm < nvm; ++dm) for (int df = 0; df < nvf; ++df)
                    add_voxel(V, F0 + fs * SEG + df, M0 + m0 + dm, S0 + s0 + ds, slow_offset, gotL, gotP, L, MAXL);
            }
    }
    for (auto it = gotP.begin(); it != gotP.end();) { bool z = true; for (u64 x : it->second) z = z && x == 0; it = z ? gotP.erase(it) : std::next(it); }
    if (gotL != refL || gotP != refP) {
        fprintf(stderr, "MISMATCH T=%d MAXL=%d nf=%d nm=%d nbuf=%d own=[%d,%d) offset=%ld labels=%d mode=%d seed=%u: labels %zu/%zu (equal %d) pairs %zu/%zu (equal %d)\n",
                (int)sizeof(T), MAXL, nf, nm, nbuf, own_lo, own_hi, slow_offset, nlabels, mode, seed, gotL.size(), refL.size(),
                (int)(gotL == refL), gotP.size(), refP.size(), (int)(gotP == refP));
        return 1;
    }
    return 0;
}

int main(int argc, char** argv) {
    std::mt19937 rng(argc > 1 ? (unsigned)atoi(argv[1]) : 1u);
    int bad = 0; long hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < 72; ++c) {
        const int nf = 1 + rng() % 300, nm = 1 + rng() % 40, nbuf = 1 + rng() % 22;
        int lo = 0, hi = nbuf; long off = 0;
        if (c % 3 == 1 && nbuf >= 3) { lo = 1; hi = nbuf - 1; off = 1000 + rng() % 5000; }     // a slab with halo planes
        const int nl = 1 + rng() % (c % 4 == 0 ? 40 : 12), mode = c % 3 == 0 ? 0 : 1;
        const unsigned seed = rng();
        switch (c % 6) {
            case 0: bad += run_case<uint16_t, 4>(nf, nm, nbuf, lo, hi, off, nl, mode, seed, hist); break;
            case 1: bad += run_case<uint32_t, 4>(nf, nm, nbuf, lo, hi, off, nl, mode, seed, hist); break;
            case 2: bad += run_case<uint16_t, 2>(nf, nm, nbuf, lo, hi, off, nl, mode, seed, hist); break;
            case 3: bad += run_case<uint32_t, 3>(nf, nm, nbuf, lo, hi, off, nl, mode, seed, hist); break;
            case 4: bad += run_case<uint16_t, 6>(nf, nm, nbuf, lo, hi, off, nl, mode, seed, hist); break;
            default: bad += run_case<uint32_t, 5>(nf, nm, nbuf, lo, hi, off, nl, mode, seed, hist); break;
        }
    }
    printf("block_level_check: 72 volumes; blocks by labels known at the end: 1: %ld, 2: %ld, 3: %ld, 4: %ld, 5: %ld, fallback: %ld; "
           "%d mismatching volumes\n", hist[1], hist[2], hist[3], hist[4], hist[5], hist[0], bad);
    return bad ? 1 : 0;
}

// CPU check of the whole block-bitmask pass (csrc/ta_block.cuh; test infrastructure, no GPU needed): a volume of several
// bricks is tiled exactly as the scan kernel tiles it (128 x 16 x 8 bricks, clamped one-voxel halo, ragged last bricks,
// an owned plane range inside a taller buffer), every brick is cut into 8 x 4 x 2 blocks, block_features_reg fills a
// global label table (sums moved block -> brick -> global with block_shift_moments) and a global pair table (six
// directional face slots + wall18, the conventions of include/tissue_b200.h); blocks with more labels than slots take a
// per-voxel fallback.  The result must equal a direct pass over the owned voxels.
// Usage: block_volume_check <seed> ; exit code 0 = every case equal.
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

static void add_voxel(const Vol& V, int f, int m, int s, long slow_offset, LabelTab& lt, PairTab& pt) {
    const uint32_t a = V.at(f, m, s);
    LabelRow& r = lt[a];
    const u64 F = f, M = m, S = (u64)(s + slow_offset);
    r.v[0] += 1; r.v[1] += F; r.v[2] += M; r.v[3] += S; r.v[4] += F * F; r.v[5] += F * M; r.v[6] += F * S; r.v[7] += M * M;
    r.v[8] += M * S; r.v[9] += S * S;
    const long c[3] = {(long)F, (long)M, (long)S};
    for (int k = 0; k < 3; ++k) { r.bmin[k] = std::min(r.bmin[k], c[k]); r.bmax[k] = std::max(r.bmax[k], c[k]); }
    std::map<uint32_t, int> seen;
    for (int z = -1; z <= 1; ++z) for (int y = -1; y <= 1; ++y) for (int x = -1; x <= 1; ++x) {
        const int l1 = abs(z) + abs(y) + abs(x);
        if (l1 < 1 || l1 > 2) continue;
        const uint32_t b = V.at(f + x, m + y, s + z);
        if (b != a) seen[b] = 1;
    }
    for (auto& kv : seen) pt[{std::min(a, kv.first), std::max(a, kv.first)}][6] += 1;
    const uint32_t nb[3] = {V.at(f + 1, m, s), V.at(f, m + 1, s), V.at(f, m, s + 1)};
    for (int k = 0; k < 3; ++k)
        if (nb[k] != a) pt[{std::min(a, nb[k]), std::max(a, nb[k])}][2 * k + (a < nb[k] ? 0 : 1)] += 1;
}

struct OnLabel {
    LabelTab* lt; uint32_t bf, bm, bs;          // block origin inside the brick
    u64 F0, M0, S0;                              // brick origin in global coordinates
    __host__ __device__ void operator()(uint32_t L, const uint32_t vin[16]) const {
#ifndef __CUDA_ARCH__
        uint32_t v[16];
        for (int i = 0; i < 16; ++i) v[i] = vin[i];
        block_shift_moments(v, bf, bm, bs);                       // block -> brick, 32 bits as the kernel would
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
#endif
    }
};
struct OnPair {
    PairTab* pt;
    __host__ __device__ void operator()(uint32_t a, uint32_t b, uint32_t w18, uint32_t ff, uint32_t fm, uint32_t fs) const {
#ifndef __CUDA_ARCH__
        auto& r = (*pt)[{std::min(a, b), std::max(a, b)}];
        const int side = a < b ? 0 : 1;            // the lower-index voxel carries the smaller label: slot 2a, else 2a + 1
        r[6] += w18; r[0 + side] += ff; r[2 + side] += fm; r[4 + side] += fs;
#endif
    }
};

template <typename T>
static int run_case(int nf, int nm, int nbuf, int own_lo, int own_hi, long slow_offset, int nlabels, int mode, unsigned seed,
                    long* nover) {
    constexpr int SEG = Vox<T>::SEG, ROWE = ROWV * SEG, BF = NFS * SEG;
    std::mt19937 rng(seed);
    Vol V{nf, nm, nbuf, std::vector<uint32_t>((size_t)nf * nm * nbuf)};
    std::vector<uint32_t> names(nlabels);
    for (auto& n : names) n = 1 + (sizeof(T) == 2 ? rng() % 65000u : rng() % 0xFFFFFF00u);
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
    LabelTab gotL, refL; PairTab gotP, refP;
    for (int s = own_lo; s < own_hi; ++s) for (int m = 0; m < nm; ++m) for (int f = 0; f < nf; ++f)
        add_voxel(V, f, m, s, slow_offset, refL, refP);
    std::vector<uint4> tile(TILE_SEGS);
    T* tl = reinterpret_cast<T*>(tile.data());
    for (int S0 = own_lo; S0 < own_hi; S0 += BS) for (int M0 = 0; M0 < nm; M0 += BM) for (int F0 = 0; F0 < nf; F0 += BF) {
        for (int r = 0; r < TILE_ROWS; ++r) {
            const int m = r % (BM + 2) - 1, s = r / (BM + 2) - 1;
            for (int e = 0; e < ROWE; ++e) tl[(size_t)r * ROWE + e] = (T)V.at(F0 + e - SEG, M0 + m, S0 + s);
        }
        for (int s0 = 0; s0 < BS && S0 + s0 < own_hi; s0 += BLK_S) for (int m0 = 0; m0 < BM && M0 + m0 < nm; m0 += BLK_M)
            for (int fs = 0; fs < NFS && F0 + fs * SEG < nf; ++fs) {
                const int nvf = std::min(SEG, nf - F0 - fs * SEG), nvm = std::min(BLK_M, nm - M0 - m0),
                          nvs = std::min(BLK_S, own_hi - S0 - s0);
                OnLabel ol{&gotL, (uint32_t)(fs * SEG), (uint32_t)m0, (uint32_t)s0, (u64)F0, (u64)M0, (u64)(S0 + slow_offset)};
                if (!block_features_reg<T, BLK_MAXLAB>(tile.data(), fs, m0, s0, nvf, nvm, nvs, ol, OnPair{&gotP})) {
                    ++*nover;       // more labels than slots: the per-voxel path
                    for (int ds = 0; ds < nvs; ++ds) for (int dm = 0; dm < nvm; ++dm) for (int df = 0; df < nvf; ++df)
                        add_voxel(V, F0 + fs * SEG + df, M0 + m0 + dm, S0 + s0 + ds, slow_offset, gotL, gotP);
                }
            }
    }
    for (auto it = gotP.begin(); it != gotP.end();) { bool z = true; for (u64 x : it->second) z = z && x == 0; it = z ? gotP.erase(it) : std::next(it); }
    if (gotL != refL || gotP != refP) {
        fprintf(stderr, "MISMATCH nf=%d nm=%d nbuf=%d own=[%d,%d) offset=%ld labels=%d mode=%d seed=%u: labels %zu/%zu pairs %zu/%zu\n",
                nf, nm, nbuf, own_lo, own_hi, slow_offset, nlabels, mode, seed, gotL.size(), refL.size(), gotP.size(), refP.size());
        return 1;
    }
    return 0;
}

int main(int argc, char** argv) {
    std::mt19937 rng(argc > 1 ? (unsigned)atoi(argv[1]) : 1u);
    int bad = 0; long nover = 0;
    for (int c = 0; c < 60; ++c) {
        const int nf = 1 + rng() % 300, nm = 1 + rng() % 40, nbuf = 1 + rng() % 22;
        int lo = 0, hi = nbuf; long off = 0;
        if (c % 3 == 1 && nbuf >= 3) { lo = 1; hi = nbuf - 1; off = 1000 + rng() % 5000; }     // a slab with halo planes
        const int nl = 1 + rng() % (c % 4 == 0 ? 40 : 12), mode = c % 3 == 0 ? 0 : 1;
        bad += (c & 1) ? run_case<uint32_t>(nf, nm, nbuf, lo, hi, off, nl, mode, rng(), &nover)
                       : run_case<uint16_t>(nf, nm, nbuf, lo, hi, off, nl, mode, rng(), &nover);
    }
    printf("block_volume_check: 60 volumes, %ld blocks on the per-voxel fallback, %d mismatching volumes\n", nover, bad);
    return bad ? 1 : 0;
}

// CPU check of the block-bitmask feature extraction (csrc/ta_block.cuh; test infrastructure, no GPU needed).
// Builds one brick tile as phase A of the scan kernel leaves it (labels, clamped halo), runs ta::block_features -- the
// code a kernel would run, one call per 8 x 4 x 2 block -- and compares, block by block, the per-label moments / boxes and
// the per (own label, other label) wall18 and face counts with a brute-force pass over the block's voxels.
// Usage: block_host_check <seed> ; exit code 0 = every block of every case equal.
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>
#include "../../tissue_analysis_b200/csrc/ta_block.cuh"

using namespace ta;

typedef std::map<uint32_t, std::array<uint32_t, 16>> LabelMap;
typedef std::map<std::pair<uint32_t, uint32_t>, std::array<uint32_t, 4>> PairMap;

struct OnLabel {
    LabelMap* out;
    __host__ __device__ void operator()(uint32_t L, const uint32_t v[16]) const {
#ifndef __CUDA_ARCH__
        std::array<uint32_t, 16> a;
        for (int i = 0; i < 16; ++i) a[i] = v[i];
        (*out)[L] = a;
#endif
    }
};
struct OnPair {
    PairMap* out;
    __host__ __device__ void operator()(uint32_t a, uint32_t b, uint32_t w18, uint32_t ff, uint32_t fm, uint32_t fs) const {
#ifndef __CUDA_ARCH__
        (*out)[{a, b}] = {w18, ff, fm, fs};
#endif
    }
};

template <typename T>
static int run_case(int nf, int nm, int ns, int nlabels, int mode, unsigned seed, long* nblocks, long* noverflow) {
    constexpr int SEG = Vox<T>::SEG, ROWE = ROWV * SEG;
    std::mt19937 rng(seed);
    std::vector<uint32_t> vol((size_t)nf * nm * ns), names(nlabels);
    for (auto& n : names) n = sizeof(T) == 2 ? rng() % 65535u : rng() % 0xFFFFFFF0u;
    if (mode == 0) {
        for (auto& v : vol) v = names[rng() % nlabels];
    } else {
        std::vector<int> sx(nlabels), sy(nlabels), sz(nlabels);
        for (int k = 0; k < nlabels; ++k) { sx[k] = rng() % nf; sy[k] = rng() % nm; sz[k] = rng() % ns; }
        for (int s = 0; s < ns; ++s) for (int m = 0; m < nm; ++m) for (int f = 0; f < nf; ++f) {
            long best = 1L << 60; int bk = 0;
            for (int k = 0; k < nlabels; ++k) {
                long d = (long)(f - sx[k]) * (f - sx[k]) + (long)(m - sy[k]) * (m - sy[k]) * 3 + (long)(s - sz[k]) * (s - sz[k]) * 5;
                if (d < best) { best = d; bk = k; }
            }
            vol[((size_t)s * nm + m) * nf + f] = names[bk];
        }
    }
    auto at = [&](int f, int m, int s) -> uint32_t {
        f = std::min(std::max(f, 0), nf - 1); m = std::min(std::max(m, 0), nm - 1); s = std::min(std::max(s, 0), ns - 1);
        return vol[((size_t)s * nm + m) * nf + f];
    };
    std::vector<uint4> tile(TILE_SEGS);
    T* tl = reinterpret_cast<T*>(tile.data());
    for (int r = 0; r < TILE_ROWS; ++r) {
        const int m = r % (BM + 2) - 1, s = r / (BM + 2) - 1;
        for (int e = 0; e < ROWE; ++e) tl[(size_t)r * ROWE + e] = (T)at(e - SEG, m, s);
    }
    int bad = 0;
    for (int s0 = 0; s0 < std::min(ns, BS); s0 += BLK_S) for (int m0 = 0; m0 < std::min(nm, BM); m0 += BLK_M)
        for (int fs = 0; fs * SEG < nf; ++fs) {
            const int nvf = std::min(SEG, nf - fs * SEG), nvm = std::min(BLK_M, nm - m0), nvs = std::min(BLK_S, ns - s0);
            LabelMap gotL, refL;
            PairMap gotP, refP;
            ++*nblocks;
            if (!block_features<T>(tile.data(), fs, m0, s0, nvf, nvm, nvs, OnLabel{&gotL}, OnPair{&gotP})) {
                LabelMap l2; PairMap p2;
                if (block_features_reg<T, BLK_MAXLAB>(tile.data(), fs, m0, s0, nvf, nvm, nvs, OnLabel{&l2}, OnPair{&p2})) ++bad;
                ++*noverflow;
                continue;
            }
            {
                // the register-resident form must agree with the reference form, also with a smaller slot budget
                LabelMap l2, l3; PairMap p2, p3;
                const bool ok4 = block_features_reg<T, BLK_MAXLAB>(tile.data(), fs, m0, s0, nvf, nvm, nvs, OnLabel{&l2}, OnPair{&p2});
                const bool ok3 = block_features_reg<T, 3>(tile.data(), fs, m0, s0, nvf, nvm, nvs, OnLabel{&l3}, OnPair{&p3});
                if (!ok4 || l2 != gotL || p2 != gotP) ++bad;
                if (ok3 && (l3 != gotL || p3 != gotP)) ++bad;
            }
            for (int ds = 0; ds < nvs; ++ds) for (int dm = 0; dm < nvm; ++dm) for (int df = 0; df < nvf; ++df) {
                const int f = fs * SEG + df, m = m0 + dm, s = s0 + ds;
                const uint32_t a = at(f, m, s);
                auto it = refL.find(a);
                if (it == refL.end()) {
                    std::array<uint32_t, 16> z{};
                    z[10] = z[11] = z[12] = 0xFFFFFFFFu;
                    it = refL.emplace(a, z).first;
                }
                auto& v = it->second;
                const uint32_t uf = df, um = dm, us = ds;
                v[0] += 1; v[1] += uf; v[2] += um; v[3] += us; v[4] += uf * uf; v[5] += uf * um; v[6] += uf * us;
                v[7] += um * um; v[8] += um * us; v[9] += us * us;
                v[10] = std::min(v[10], uf); v[11] = std::min(v[11], um); v[12] = std::min(v[12], us);
                v[13] = std::max(v[13], uf); v[14] = std::max(v[14], um); v[15] = std::max(v[15], us);
                std::map<uint32_t, int> seen;
                for (int z = -1; z <= 1; ++z) for (int y = -1; y <= 1; ++y) for (int x = -1; x <= 1; ++x) {
                    const int l1 = abs(z) + abs(y) + abs(x);
                    if (l1 < 1 || l1 > 2) continue;
                    const uint32_t b = at(f + x, m + y, s + z);
                    if (b != a) seen[b] = 1;
                }
                for (auto& kv : seen) refP[{a, kv.first}][0] += 1;
                uint32_t b;
                b = at(f + 1, m, s); if (b != a) refP[{a, b}][1] += 1;
                b = at(f, m + 1, s); if (b != a) refP[{a, b}][2] += 1;
                b = at(f, m, s + 1); if (b != a) refP[{a, b}][3] += 1;
            }
            if (gotL != refL || gotP != refP) {
                if (!bad) fprintf(stderr, "MISMATCH nf=%d nm=%d ns=%d labels=%d mode=%d seed=%u block (fs=%d, m0=%d, s0=%d): "
                                          "%zu/%zu labels, %zu/%zu pairs\n", nf, nm, ns, nlabels, mode, seed, fs, m0, s0,
                                  gotL.size(), refL.size(), gotP.size(), refP.size());
                ++bad;
            }
        }
    return bad;
}

int main(int argc, char** argv) {
    const unsigned seed0 = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
    std::mt19937 rng(seed0);
    long bad = 0, nblocks = 0, noverflow = 0;
    for (int c = 0; c < 300; ++c) {
        const bool wide = (c & 1) != 0;                         // uint32 labels: 4-voxel segments, 64-voxel rows
        const int maxf = NFS * (wide ? 4 : 8);
        int nf = 1 + rng() % maxf, nm = 1 + rng() % BM, ns = 1 + rng() % BS;
        if (c % 5 == 0) { nf = maxf; nm = BM; ns = BS; }
        const int nl = 1 + rng() % (c % 4 == 0 ? 30 : 9);
        const int mode = (c % 3 == 0) ? 0 : 1;
        bad += wide ? run_case<uint32_t>(nf, nm, ns, nl, mode, rng(), &nblocks, &noverflow)
                    : run_case<uint16_t>(nf, nm, ns, nl, mode, rng(), &nblocks, &noverflow);
    }
    printf("block_host_check: %ld blocks, %ld with more than %d labels in the window (skipped), %ld mismatching blocks\n",
           nblocks, noverflow, BLK_MAXLAB, bad);
    return (bad || nblocks - noverflow < 10000) ? 1 : 0;
}

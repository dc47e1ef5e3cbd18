// CPU check of the one-hot pair arithmetic of csrc/ta_scan.cuh (test infrastructure, no GPU needed).
// Builds one brick tile as phase A of the scan kernel leaves it (clamped halo), computes the uniformity codes of phase
// B, then runs the kernel's own phase R (ta::oh_relabel_segment: quarter id tables, in-place one-hot rewrite, edge
// array) and phase S (ta::oh_segment_pairs) code on the host for every segment of the brick, and compares the per
// (own label, other label) counters with a brute-force count over the voxels (the definitions of SURVEY.md appendix A:
// 18-connected wall voxels with per-voxel de-duplication, +f / +m / +s faces).
// Usage: oh_host_check <seed> ; exit code 0 = every case equal.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <random>
#include <vector>
#include "../../tissue_analysis_b200/csrc/ta_scan.cuh"

using namespace ta;

struct Cnt { uint32_t w18 = 0, f = 0, m = 0, s = 0; bool operator==(const Cnt& o) const { return w18 == o.w18 && f == o.f && m == o.m && s == o.s; } };
typedef std::map<std::pair<uint32_t, uint32_t>, Cnt> PairMap;       // (own label, other label)

template <typename T> struct HostEmit {
    PairMap* out;
    const uint32_t* idk;
    __host__ __device__ void operator()(uint32_t key, uint32_t c0, uint32_t c1) const {
#ifndef __CUDA_ARCH__
        constexpr int NID = OneHot<T>::NID;
        const uint32_t q = key / (NID * NID), i = (key / NID) % NID, j = key % NID;
        Cnt& c = (*out)[{idk[q * NID + i], idk[q * NID + j]}];
        c.w18 += c0 & 0xFFFFu; c.f += c0 >> 16; c.m += c1 & 0xFFFFu; c.s += c1 >> 16;
#endif
    }
};

template <typename T>
static int run_case(int nf, int nm, int ns, int nlabels, int mode, unsigned seed, bool do_w18, bool do_p6) {
    constexpr int SEG = Vox<T>::SEG;
    constexpr int ROWE = ROWV * SEG;
    constexpr int NID = OneHot<T>::NID, NQ = OneHot<T>::NQ;
    std::mt19937 rng(seed);
    std::vector<uint32_t> vol((size_t)nf * nm * ns);
    std::vector<uint32_t> names(nlabels);
    for (auto& n : names) n = (sizeof(T) == 2) ? (rng() % 65535u) : (rng() % 0xFFFFFFF0u);
    if (mode == 0) {                      // noise
        for (auto& v : vol) v = names[rng() % nlabels];
    } else {                              // blobs: nearest of nlabels seeds (walls, junction lines, thin slivers)
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
    // phase A: labels with the clamped halo
    std::vector<uint4> tile(TILE_SEGS);
    T* tl = reinterpret_cast<T*>(tile.data());
    for (int r = 0; r < TILE_ROWS; ++r) {
        const int m = r % (BM + 2) - 1, s = r / (BM + 2) - 1;
        for (int e = 0; e < ROWE; ++e) tl[(size_t)r * ROWE + e] = (T)at(e - SEG, m, s);
    }
    // phase B: uniformity codes
    std::vector<uint32_t> codes((size_t)TILE_ROWS * NFS);
    for (int r = 0; r < TILE_ROWS; ++r) for (int fs = 0; fs < NFS; ++fs) {
        const T* rp = tl + (size_t)r * ROWE + (fs + 1) * SEG;
        bool uni = true;
        for (int j = -1; j <= SEG; ++j) uni = uni && rp[j] == rp[0];
        codes[r * NFS + fs] = uni ? (uint32_t)rp[0] : Vox<T>::MIXED;
    }
    // phase R (kernel code): in any order, here a shuffled one
    std::vector<uint32_t> idk(64, TA_EMPTY32);
    std::vector<T> edg((size_t)TILE_ROWS * (NQ > 1 ? NQ - 1 : 1) * 2, (T)0);
    std::vector<int> order(TILE_ROWS * NFS);
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::shuffle(order.begin(), order.end(), rng);
    uint32_t lastL = TA_EMPTY32, lastQ = 0, lastOH = 0;
    bool ok = true;
    for (int i : order)
        ok = oh_relabel_segment<T>(tile.data(), edg.data(), idk.data(), i / NFS, i % NFS, codes[i], lastL, lastQ, lastOH) && ok;
    if (!ok) return -1;                   // a quarter holds more than NID labels: the kernel takes the per-voxel path
    // phase S (kernel code)
    PairMap got, ref;
    const uint32_t keep0 = (do_w18 ? 0xFFFFu : 0u) | (do_p6 ? 0xFFFF0000u : 0u);
    for (int s = 0; s < std::min(ns, BS); ++s) for (int m = 0; m < std::min(nm, BM); ++m)
        for (int fs = 0; fs * SEG < nf; ++fs)
            oh_segment_pairs<T>(tile.data(), edg.data(), (s + 1) * (BM + 2) + (m + 1), fs, nf - fs * SEG, keep0, do_p6,
                                HostEmit<T>{&got, idk.data()});
    // brute force
    for (int s = 0; s < ns; ++s) for (int m = 0; m < nm; ++m) for (int f = 0; f < nf; ++f) {
        const uint32_t a = at(f, m, s);
        std::map<uint32_t, int> seen;
        for (int ds = -1; ds <= 1; ++ds) for (int dm = -1; dm <= 1; ++dm) for (int df = -1; df <= 1; ++df) {
            const int l1 = abs(ds) + abs(dm) + abs(df);
            if (l1 < 1 || l1 > 2) continue;
            const uint32_t b = at(f + df, m + dm, s + ds);
            if (b != a) seen[b] = 1;
        }
        if (do_w18) for (auto& kv : seen) ref[{a, kv.first}].w18 += 1;
        if (do_p6) {
            uint32_t b;
            b = at(f + 1, m, s); if (b != a) ref[{a, b}].f += 1;
            b = at(f, m + 1, s); if (b != a) ref[{a, b}].m += 1;
            b = at(f, m, s + 1); if (b != a) ref[{a, b}].s += 1;
        }
    }
    for (auto it = got.begin(); it != got.end();) { if (it->second == Cnt()) it = got.erase(it); else ++it; }
    if (got != ref) {
        fprintf(stderr, "MISMATCH T=%d nf=%d nm=%d ns=%d labels=%d mode=%d seed=%u (%zu vs %zu pairs)\n", (int)sizeof(T), nf, nm,
                ns, nlabels, mode, seed, got.size(), ref.size());
        for (auto& kv : ref) {
            const Cnt g = got.count(kv.first) ? got[kv.first] : Cnt();
            if (!(g == kv.second)) {
                fprintf(stderr, "  pair (%u,%u): got %u %u %u %u want %u %u %u %u\n", kv.first.first, kv.first.second, g.w18, g.f,
                        g.m, g.s, kv.second.w18, kv.second.f, kv.second.m, kv.second.s);
                break;
            }
        }
        return 1;
    }
    return 0;
}

int main(int argc, char** argv) {
    const unsigned seed0 = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
    int bad = 0, ran = 0, skipped = 0;
    std::mt19937 rng(seed0);
    for (int c = 0; c < 400; ++c) {
        const bool wide = c & 1;
        const int maxf = wide ? NFS * 4 : NFS * 8;
        int nf = 1 + rng() % maxf, nm = 1 + rng() % BM, ns = 1 + rng() % BS;
        if (c % 7 == 0) { nf = maxf; nm = BM; ns = BS; }
        const int nl = 1 + rng() % (wide ? 24 : (c % 4 == 0 ? 40 : 14));
        const int mode = (c % 3 == 0) ? 0 : 1;
        const bool w18 = (c % 5) != 1, p6 = (c % 5) != 2;
        const int rc = wide ? run_case<uint32_t>(nf, nm, ns, nl, mode, rng(), w18, p6)
                            : run_case<uint16_t>(nf, nm, ns, nl, mode, rng(), w18, p6);
        if (rc < 0) ++skipped; else { ++ran; bad += rc; }
    }
    printf("oh_host_check: %d cases, %d skipped (more labels than ids), %d mismatches\n", ran, skipped, bad);
    return (bad || ran < 300) ? 1 : 0;
}

// Runs the scan kernel ITSELF on the CPU (tests/host/emu/cuda_emu.h: fibers for threads, rendezvous for the warp and block
// collectives) and compares the tables it fills with a direct pass over the voxels: scan_kernel<T, false> for both label
// widths (march, worklists, per-voxel pair phases, flush, slab ownership, ragged bricks) on the scalar staging path and on
// the TMA staging path with the box copy itself emulated.  Test infrastructure: g++ only, no GPU, no nvcc.
// Usage: kernel_emu_check <seed> ; exit code 0 = every case equal.
#include "emu/cuda_emu.h"

#define TA_EMU_TMA 1
static unsigned long long g_user_stat[16];
#define TA_STAT(which, n) (g_user_stat[which] += (unsigned long long)(n))

#include "../../tissue_analysis_b200/csrc/ta_scan.cuh"
#include "../../tissue_analysis_b200/csrc/ta_scan_mask.cuh"
#include "../../tissue_analysis_b200/csrc/ta_prepass.cuh"

namespace ta { alignas(128) unsigned char smem_raw[160 * 1024]; namespace mk { alignas(128) unsigned char smem_raw[160 * 1024]; } }
void ta::ta_emu_yield() { emu::g_progress = true; emu::yield(); }

using namespace ta;

struct LabelRow { u64 v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long bmin[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, bmax[3] = {-1, -1, -1};
    bool operator==(const LabelRow& o) const { return std::equal(v, v + 10, o.v) && std::equal(bmin, bmin + 3, o.bmin) && std::equal(bmax, bmax + 3, o.bmax); } };
typedef std::map<uint32_t, LabelRow> LabelTab;
typedef std::map<std::pair<uint32_t, uint32_t>, std::array<u64, 7>> PairTab;

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

static long g_wm = 2, g_ws = 3;      // metric weights of the blob volumes (mid, slow axis); --stats uses 1, 1 (round cells)
static unsigned long g_skipped = 0, g_pp_bricks = 0;      // bricks the pre-pass took out of the queue / saw
enum Which { PRODUCT, MASK, MASK_PP, NWHICH };      // MASK_PP: the pre-pass (ta_prepass.cuh) in front of the mask kernel
static const char* which_name[] = {"scan_kernel<T,false>", "mk::mask_kernel<T>", "pp:: pre-pass + mk::mask_kernel<T>"};

template <typename T>
static int run_case(Which which, int nf, int nm, int nbuf, int own_lo, int own_hi, long slow_offset, int nlabels, int mode,
                    unsigned seed, int use_tma = 0) {
    std::mt19937 rng(seed);
    Vol V{nf, nm, nbuf, std::vector<uint32_t>((size_t)nf * nm * nbuf)};
    const uint32_t top = sizeof(T) == 2 ? 4000u : 9000u;            // dense label table: keep it small
    std::vector<uint32_t> names(nlabels);
    for (auto& n : names) n = 1 + rng() % top;
    if (sizeof(T) == 2 && nlabels >= 3 && (seed & 1u)) names[1] = 0xFFFFu;      // the uint16 value that doubles as the MIXED code
    if (mode == 0) {
        for (auto& v : V.d) v = names[rng() % nlabels];
    } else if (mode == 3) {          // background with a few balls in it: whole regions of one-label bricks (the pre-pass's case)
        for (auto& v : V.d) v = names[0];
        for (int k = 1; k < nlabels; ++k) {
            const int cx = rng() % nf, cy = rng() % nm, cz = rng() % nbuf, rad = 4 + rng() % 9;
            for (int s = std::max(0, cz - rad); s < std::min(nbuf, cz + rad + 1); ++s)
                for (int m = std::max(0, cy - rad); m < std::min(nm, cy + rad + 1); ++m)
                    for (int f = std::max(0, cx - rad); f < std::min(nf, cx + rad + 1); ++f)
                        if ((f - cx) * (f - cx) + (m - cy) * (m - cy) + (s - cz) * (s - cz) <= rad * rad) V.d[((size_t)s * nm + m) * nf + f] = names[k];
        }
    } else if (mode == 2) {          // one label almost everywhere: every lane of a warp carries a nearly full block of it
        for (auto& v : V.d) v = names[rng() % 97 == 0 ? rng() % nlabels : 0];
    } else {
        std::vector<int> sx(nlabels), sy(nlabels), sz(nlabels);
        for (int k = 0; k < nlabels; ++k) { sx[k] = rng() % nf; sy[k] = rng() % nm; sz[k] = rng() % nbuf; }
        for (int s = 0; s < nbuf; ++s) for (int m = 0; m < nm; ++m) for (int f = 0; f < nf; ++f) {
            long best = 1L << 60; int bk = 0;
            for (int k = 0; k < nlabels; ++k) {
                long d = (long)(f - sx[k]) * (f - sx[k]) + (long)(m - sy[k]) * (m - sy[k]) * g_wm + (long)(s - sz[k]) * (s - sz[k]) * g_ws;
                if (d < best) { best = d; bk = k; }
            }
            V.d[((size_t)s * nm + m) * nf + f] = names[bk];
        }
    }
    LabelTab refL; PairTab refP;
    for (int s = own_lo; s < own_hi; ++s) for (int m = 0; m < nm; ++m) for (int f = 0; f < nf; ++f)
        add_voxel(V, f, m, s, slow_offset, refL, refP);

    // device-side structures in host memory
    std::vector<T> vol(V.d.size());
    for (size_t i = 0; i < vol.size(); ++i) vol[i] = (T)V.d[i];
    const uint32_t nrows = sizeof(T) == 2 ? 65536u : top + 2;
    std::vector<u64> sums((size_t)nrows * 10, 0);
    std::vector<int> boxes((size_t)nrows * 6);
    for (size_t i = 0; i < (size_t)nrows * 3; ++i) { boxes[i] = 0x7FFFFFFF; boxes[(size_t)nrows * 3 + i] = -1; }
    LabelTable lt;
    lt.count = sums.data(); lt.s1 = lt.count + nrows; lt.s2 = lt.s1 + (size_t)nrows * 3;
    lt.bmin = boxes.data(); lt.bmax = lt.bmin + (size_t)nrows * 3; lt.nrows = nrows;
    const uint32_t cap = 1u << 17;                    // room for the noise cases (tens of thousands of pairs)
    std::vector<u64> keys(cap, TA_EMPTY64);
    std::vector<uint32_t> vals((size_t)cap * TA_PAIR_STRIDE, 0u), status(8, 0u);
    PairTable pt;
    pt.keys = keys.data(); pt.vals = vals.data(); pt.cap_mask = cap - 1; pt.status = status.data();
    unsigned int brick_counter = 0;
    ScanParams P{};
    P.vol = vol.data(); P.nf = nf; P.nm = nm; P.ns = nbuf; P.own_lo = own_lo; P.own_hi = own_hi; P.slow_offset = slow_offset;
    const int seg = 16 / (int)sizeof(T);
    P.nbf = (nf + NFS * seg - 1) / (NFS * seg); P.nbm = (nm + BM - 1) / BM; P.nbs = (own_hi - own_lo + BS - 1) / BS;
    if (which != PRODUCT) { P.nbf = (nf + mk::RW - 1) / mk::RW; P.nbm = (nm + mk::OM - 1) / mk::OM; P.nbs = (own_hi - own_lo + mk::ZB - 1) / mk::ZB; }
    P.flags = 7u; P.vec_ok = 0; P.use_tma = use_tma; P.brick_counter = &brick_counter; P.phase_cycles = nullptr; P.diag = nullptr;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    {
        // what ta_api.cu encodes for the kernels: the bound buffer, one box = tile (brick + halo)
        ta::EmuTmap em{vol.data(), nf, nm, nbuf, (int)sizeof(T), ROWV * seg, BM + 2, BS + 2};
        if (which != PRODUCT) { em.box0 = mk::Geo<T>::TRE; em.box1 = mk::TM; em.box2 = mk::TP; }
        static_assert(sizeof(ta::EmuTmap) <= sizeof(CUtensorMap), "the emulated map lives in the bytes of the real one");
        memcpy(&tmap, &em, sizeof em);
    }
    bool ok = true;
    // the pre-pass needs rows of whole 16-byte vectors (the host checks vec_ok); other shapes run the mask kernel alone
    std::vector<uint32_t> core;
    std::vector<unsigned int> work_list;
    unsigned int work_count = 0;
    if (which == MASK_PP && nf % seg == 0) {
        const unsigned total = (unsigned)P.nbf * P.nbm * P.nbs;
        core.assign(total, 0u); work_list.assign(total, 0u);
        pp::PrepassParams Q{};
        Q.vol = vol.data(); Q.nf = nf; Q.nm = nm; Q.ns = nbuf; Q.own_lo = own_lo; Q.own_hi = own_hi; Q.slow_offset = slow_offset;
        Q.nbf = P.nbf; Q.nbm = P.nbm; Q.nbs = P.nbs; Q.core = core.data(); Q.work_list = work_list.data(); Q.work_count = &work_count;
        Q.do_mom = 1u;
        ok = emu::run_block(0, 1, 256, [&]() { pp::classify_cores_kernel<T>(Q); });
        for (unsigned block = 0; block < (total + 255) / 256 && ok; ++block)
            ok = emu::run_block(block, (total + 255) / 256, 256, [&]() { pp::decide_kernel(Q, lt, status.data()); });
        P.work_list = work_list.data(); P.work_count = &work_count;
        g_skipped += total - work_count; g_pp_bricks += total;
    }
    for (unsigned block = 0; block < 2 && ok; ++block) {           // the second block finds the brick counter exhausted
        if (which != PRODUCT) ok = emu::run_block(block, 2, mk::NTHREADS, [&]() { if (P.flags == 7u) mk::mask_kernel<T, 7>(P, lt, pt, tmap); else mk::mask_kernel<T, -1>(P, lt, pt, tmap); });
        else ok = emu::run_block(block, 2, NTHREADS, [&]() {
            scan_kernel<T, false>(P, lt, pt, tmap);
        });
    }
    LabelTab gotL; PairTab gotP;
    for (uint32_t L = 0; L < nrows; ++L) {
        if (!lt.count[L]) continue;
        LabelRow r;
        r.v[0] = lt.count[L];
        for (int k = 0; k < 3; ++k) { r.v[1 + k] = lt.s1[(size_t)L * 3 + k]; r.bmin[k] = lt.bmin[(size_t)L * 3 + k]; r.bmax[k] = lt.bmax[(size_t)L * 3 + k]; }
        for (int k = 0; k < 6; ++k) r.v[4 + k] = lt.s2[(size_t)L * 6 + k];
        gotL[L] = r;
    }
    for (uint32_t i = 0; i < cap; ++i) {
        if (keys[i] == TA_EMPTY64) continue;
        std::array<u64, 7> a;
        u64 any = 0;
        for (int k = 0; k < 7; ++k) { a[k] = vals[(size_t)i * TA_PAIR_STRIDE + k]; any |= a[k]; }
        if (any) gotP[{(uint32_t)(keys[i] >> 32), (uint32_t)keys[i]}] = a;
    }
    if (!ok || status[0] || status[1] || gotL != refL || gotP != refP) {
        fprintf(stderr, "MISMATCH %s T=%d nf=%d nm=%d nbuf=%d own=[%d,%d) labels=%d mode=%d seed=%u: run %s, status %u %u, labels %zu/%zu "
                        "(equal %d), pairs %zu/%zu (equal %d)%s\n", which_name[which], (int)sizeof(T), nf, nm, nbuf, own_lo, own_hi, nlabels,
                mode, seed, ok ? "ok" : "DEADLOCK", status[0], status[1], gotL.size(), refL.size(), (int)(gotL == refL), gotP.size(),
                refP.size(), (int)(gotP == refP), use_tma ? " [emulated TMA staging]" : "");
        if (getenv("EMU_DIFF")) {
            int shown = 0;
            for (auto& kv : refL) {
                auto it = gotL.find(kv.first);
                if (it == gotL.end()) { if (shown++ < 6) fprintf(stderr, "  label %u missing (ref n=%llu)\n", kv.first, kv.second.v[0]); continue; }
                if (!(it->second == kv.second) && shown++ < 6) {
                    fprintf(stderr, "  label %u ref:", kv.first); for (int k = 0; k < 10; ++k) fprintf(stderr, " %llu", kv.second.v[k]);
                    fprintf(stderr, " box %ld %ld %ld - %ld %ld %ld\n             got:", kv.second.bmin[0], kv.second.bmin[1], kv.second.bmin[2], kv.second.bmax[0], kv.second.bmax[1], kv.second.bmax[2]);
                    for (int k = 0; k < 10; ++k) fprintf(stderr, " %llu", it->second.v[k]);
                    fprintf(stderr, " box %ld %ld %ld - %ld %ld %ld\n", it->second.bmin[0], it->second.bmin[1], it->second.bmin[2], it->second.bmax[0], it->second.bmax[1], it->second.bmax[2]);
                }
            }
            shown = 0;
            for (auto& kv : refP) {
                auto it = gotP.find(kv.first);
                if ((it == gotP.end() || it->second != kv.second) && shown++ < 6) {
                    fprintf(stderr, "  pair (%u,%u) ref:", kv.first.first, kv.first.second); for (int k = 0; k < 7; ++k) fprintf(stderr, " %llu", kv.second[k]);
                    fprintf(stderr, "  got:"); if (it != gotP.end()) for (int k = 0; k < 7; ++k) fprintf(stderr, " %llu", it->second[k]); else fprintf(stderr, " none");
                    fprintf(stderr, "\n");
                }
            }
        }
        return 1;
    }
    return 0;
}

// Operation counts per voxel of the six kernels on a tissue-like volume (cells of ~21 500 voxels as in C3): what the
// table updates cost in shared-memory atomics and warp collectives.  Not a time; a GPU decides that.
static void print_stats() {
    const int nf = 256, nm = 128, nbuf = 64;
    g_wm = g_ws = 1;
    const int ncell = (int)((double)nf * nm * nbuf / 21500.0 + 0.5);
    printf("tissue-like volume %d x %d x %d, %d cells; operations per voxel\n", nf, nm, nbuf, ncell);
    printf("%-28s %9s %9s %9s %9s %9s %9s\n", "kernel", "atom.smem", "atom.glob", "redux/w", "ballot/w", "shfl/w", "bar/blk");
    for (int w = 0; w < NWHICH; ++w) {
        emu::g_stats.clear();
        const int bad = run_case<uint16_t>((Which)w, nf, nm, nbuf, 0, nbuf, 0, ncell, 1, 12345u);
        const double nv = (double)nf * nm * nbuf;
        const emu::Stats& s = emu::g_stats;
        printf("%-28s %9.4f %9.4f %9.4f %9.4f %9.4f %9.5f%s\n", which_name[w], s.atom_shared / nv, s.atom_global / nv, s.redux / nv,
               s.ballot / nv, s.shfl / nv, s.syncthreads / nv, bad ? "  (MISMATCH)" : "");
    }
}

int main(int argc, char** argv) {
    emu::g_smem_lo = ta::smem_raw;
    emu::g_smem_hi = ta::smem_raw + sizeof ta::smem_raw;
    if (argc > 1 && !strcmp(argv[1], "--stats")) { print_stats(); return 0; }
    std::mt19937 rng(argc > 1 ? (unsigned)atoi(argv[1]) : 1u);
    const int ncases = argc > 2 ? atoi(argv[2]) : 36;
    const int only = argc > 3 ? atoi(argv[3]) : -1;
    int bad = 0, ran = 0;
    for (int c = 0; c < ncases; ++c) {
        const Which which = only >= 0 ? (Which)only : (Which)(c % NWHICH);
        const bool wide = (c / NWHICH) % 3 == 2;                                   // every third round: uint32 labels
        const int maxf = wide ? 150 : 300;
        int nf = 1 + rng() % maxf;
        const int nm = 1 + rng() % 36, nbuf = 1 + rng() % 19;
        if (which == MASK_PP) nf = (nf + 7) / 8 * 8;                               // the pre-pass wants rows of whole 16-byte vectors
        int lo = 0, hi = nbuf; long off = 0;
        if (c % 5 == 1 && nbuf >= 3) { lo = 1; hi = nbuf - 1; off = 1000 + rng() % 5000; }
        // c % 11 == 3: hundreds of labels in noise -- the per-brick label and pair tables fill up and spill to the global ones
        const int nl = 1 + rng() % (c % 11 == 3 ? 300 : c % 7 == 0 ? 40 : 12), mode = c % 13 == 5 ? 2 : ((c / NWHICH) % 2 == 0 || c % 11 == 3) ? 0 : 1;
        const unsigned seed = rng();
        const int use_tma = (c / NWHICH + c) % 2;          // every kernel alternates between the scalar and the (emulated) TMA staging
        bad += wide ? run_case<uint32_t>(which, nf, nm, nbuf, lo, hi, off, nl, mode, seed, use_tma)
                    : run_case<uint16_t>(which, nf, nm, nbuf, lo, hi, off, nl, mode, seed, use_tma);
        ++ran;
    }
    if (only < 0 || only == MASK_PP) {
        // background with balls, several bricks each way: one launch, a slab in the middle (halo planes owned by nobody), uint32
        bad += run_case<uint16_t>(MASK_PP, 136, 95, 26, 0, 26, 0, 4, 3, rng(), 1);
        bad += run_case<uint16_t>(MASK_PP, 104, 64, 30, 3, 27, 4000, 3, 3, rng(), 0);
        bad += run_case<uint32_t>(MASK_PP, 100, 70, 20, 0, 20, 0, 3, 3, rng(), 1);
        ran += 3;
    }
    printf("pre-pass: %lu of %lu bricks left the queue\n", g_skipped, g_pp_bricks);
    printf("kernel_emu_check: %d kernel runs on the CPU emulation, %d mismatches\n", ran, bad);
    return bad ? 1 : 0;
}

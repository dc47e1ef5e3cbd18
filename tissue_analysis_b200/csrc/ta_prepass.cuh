// One-label bricks, found at memory speed before the scan (sm_100a).
//
// Half of a dome-shaped tissue volume is background: bricks whose tile is one label altogether.  The scan kernel finds
// that out one box copy at a time (9 400 cycles per such brick on C3: 0.8 of its 2.6 ms); nothing hides the latency of the
// copy because such a brick has no work to hide it behind.  These two kernels take the interior of the one-label regions
// out of the scan's queue:
//
//   classify_cores_kernel  one warp per brick, 16-byte loads; a first look at eight rows settles most tissue bricks, then six
//                          loads in flight per lane, out at the first batch that shows a second label:
//                          core[b] = the label of a brick whose OWNED voxels are one label, TA_EMPTY32 otherwise
//   decide_kernel          one thread per brick: a brick is skipped when its core and the cores of its (up to 26) neighbours
//                          are the same label -- then its tile, halo included, is that label, which is exactly the scan
//                          kernel's own one-label case: closed-form moments (summed per warp and per CTA before they reach
//                          the global table: tens of thousands of them are background), no pairs.  Every other brick goes to
//                          the work list the scan kernel draws from, in order.
//   The criterion is sufficient, not necessary: a one-label tile next to a mixed brick stays in the list and the scan
//   kernel treats it as before (measured on C3: 83 % of the one-label bricks leave; testing the halo voxels themselves
//   takes all of them but costs 0.65 ms instead of 0.29).  Bricks at the ends of a slab (halo planes that belong to no brick
//   of the launch) stay too.
#pragma once
#include "ta_scan_mask.cuh"

#ifndef TA_SHARED
#define TA_SHARED __shared__                 /* the CPU emulation of the tests makes these statics */
#endif

namespace ta {
namespace pp {

struct PrepassParams {
    const void* vol;
    int nf, nm, ns;              // bound buffer
    int own_lo, own_hi;          // owned planes of the launch
    long long slow_offset;
    int nbf, nbm, nbs;           // bricks of the launch
    uint32_t* core;              // [nbf * nbm * nbs]: label of a brick whose owned voxels are one label, TA_EMPTY32 otherwise
    unsigned int* work_list;     // [nbf * nbm * nbs]
    unsigned int* work_count;
    uint32_t do_mom;
};

template <typename T>
__global__ void __launch_bounds__(256) classify_cores_kernel(PrepassParams P) {
    constexpr int VE = 16 / (int)sizeof(T);     // voxels per 16-byte vector
    constexpr int VPR = mk::RW / VE;            // vectors per owned row
    constexpr int RPW = 32 / VPR;               // rows per warp-wide load
    constexpr int UNR = 6;
    const int lane = threadIdx.x & 31;
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const T* vol = reinterpret_cast<const T*>(P.vol);
    const unsigned total = (unsigned)P.nbf * P.nbm * P.nbs;
    const int vx = lane % VPR, rsub = lane / VPR;
    for (unsigned b = gw; b < total; b += nw) {
        const unsigned bf = b % (unsigned)P.nbf, rest = b / (unsigned)P.nbf;
        const unsigned bm = rest % (unsigned)P.nbm, bs = rest / (unsigned)P.nbm;
        const int F0 = (int)bf * mk::RW, M0 = (int)bm * mk::OM, S0 = P.own_lo + (int)bs * mk::ZB;
        const int nrm = min(mk::OM, P.nm - M0), nown = min(mk::ZB, P.own_hi - S0);
        const int nrow = nrm * nown;
        const bool fin = F0 + vx * VE < P.nf;                       // rows are whole vectors (the caller checked)
        const T* base = vol + ((size_t)S0 * P.nm + M0) * (size_t)P.nf + F0;
        const uint32_t ref = base[0];
        const uint32_t pat = sizeof(T) == 2 ? ref * 0x10001u : ref;
        auto fetch = [&](int r) {
            uint4 v = make_uint4(pat, pat, pat, pat);
            if (fin && r < nrow) {
                const int s = r / nrm, m = r - s * nrm;
                v = mk::mk_ld128(base + ((size_t)s * P.nm + m) * (size_t)P.nf + vx * VE);
            }
            return v;
        };
        // a look at the first rows settles most tissue bricks for 512 bytes; then six loads in flight per lane
        uint32_t diff;
        {
            const uint4 v = fetch(rsub);
            diff = (v.x ^ pat) | (v.y ^ pat) | (v.z ^ pat) | (v.w ^ pat);
        }
        bool mixed = __any_sync(0xffffffffu, diff != 0u);
        for (int r0 = RPW; r0 < nrow && !mixed; r0 += RPW * UNR) {
            uint4 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) v[u] = fetch(r0 + u * RPW + rsub);
#pragma unroll
            for (int u = 0; u < UNR; ++u) diff |= (v[u].x ^ pat) | (v[u].y ^ pat) | (v[u].z ^ pat) | (v[u].w ^ pat);
            mixed = __any_sync(0xffffffffu, diff != 0u);
        }
        if (lane == 0) P.core[b] = mixed ? TA_EMPTY32 : ref;
    }
}

__global__ void __launch_bounds__(256) decide_kernel(PrepassParams P, LabelTable lt, uint32_t* status) {
    TA_SHARED u64 acc[10];
    TA_SHARED int accmn[3], accmx[3];
    TA_SHARED uint32_t acc_label;
    TA_SHARED unsigned int wcount[8], wbase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned total = (unsigned)P.nbf * P.nbm * P.nbs;
    const unsigned b = blockIdx.x * 256u + (unsigned)tid;
    uint32_t L = TA_EMPTY32;
    bool skip = false;
    if (b < total) {
        const int bf = (int)(b % (unsigned)P.nbf);
        const unsigned rest = b / (unsigned)P.nbf;
        const int bm = (int)(rest % (unsigned)P.nbm), bs = (int)(rest / (unsigned)P.nbm);
        L = P.core[b];
        skip = L != TA_EMPTY32;
        // ends of a slab: the halo plane belongs to no brick of this launch
        if ((bs == 0 && P.own_lo > 0) || (bs == P.nbs - 1 && P.own_hi < P.ns)) skip = false;
        for (int ds = -1; ds <= 1 && skip; ++ds)
            for (int dm = -1; dm <= 1 && skip; ++dm)
                for (int df = -1; df <= 1; ++df) {
                    const int f = bf + df, m = bm + dm, s = bs + ds;
                    if (f < 0 || f >= P.nbf || m < 0 || m >= P.nbm || s < 0 || s >= P.nbs) continue;   // replicated edge voxels
                    if (P.core[((size_t)s * P.nbm + m) * P.nbf + f] != L) { skip = false; break; }
                }
    }
    // ---- the work list keeps the order of the bricks inside a CTA ---------------------------------------------------------
    const bool work = b < total && !skip;
    const unsigned bal = __ballot_sync(0xffffffffu, work);
    if (lane == 0) wcount[warp] = (unsigned)__popc(bal);
    if (tid == 0) acc_label = TA_EMPTY32;
    if (tid < 10) acc[tid] = 0ull;
    if (tid < 3) { accmn[tid] = 0x7FFFFFFF; accmx[tid] = -1; }
    __syncthreads();
    if (tid == 0) {
        unsigned sum = 0;
        for (int w = 0; w < 8; ++w) { const unsigned c = wcount[w]; wcount[w] = sum; sum += c; }
        wbase = sum ? atomicAdd(P.work_count, sum) : 0u;
    }
    if (skip && P.do_mom) atomicCAS(&acc_label, TA_EMPTY32, L);       // the CTA sums one label in shared memory: the first one
    __syncthreads();
    if (work) P.work_list[wbase + wcount[warp] + (unsigned)__popc(bal & ((1u << lane) - 1u))] = b;
    // ---- closed-form moments of the skipped bricks: summed per warp, then per CTA, then one update of the global table -----
    if (!P.do_mom) return;
    u64 g[10];
    int bmn[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, bmx[3] = {-1, -1, -1};
#pragma unroll
    for (int i = 0; i < 10; ++i) g[i] = 0ull;
    const bool mine = skip && L == acc_label;
    if (skip) {
        const int bf = (int)(b % (unsigned)P.nbf);
        const unsigned rest = b / (unsigned)P.nbf;
        const int bm = (int)(rest % (unsigned)P.nbm), bs = (int)(rest / (unsigned)P.nbm);
        const int F0 = bf * mk::RW, M0 = bm * mk::OM, S0 = P.own_lo + bs * mk::ZB;
        uint32_t v[LT_FIELDS];
        mk::one_label_sums((uint32_t)min(mk::RW, P.nf - F0), (uint32_t)min(mk::OM, P.nm - M0), (uint32_t)min(mk::ZB, P.own_hi - S0), v);
        mk::local_to_global(v, (u64)F0, (u64)M0, (u64)((long long)S0 + P.slow_offset), g, bmn, bmx);
        if (!mine) {                                                  // another label than the CTA's (rare): on its own
            mk::global_apply(lt, status, L, g, bmn, bmx);
#pragma unroll
            for (int i = 0; i < 10; ++i) g[i] = 0ull;
#pragma unroll
            for (int i = 0; i < 3; ++i) { bmn[i] = 0x7FFFFFFF; bmx[i] = -1; }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int i = 0; i < 10; ++i) g[i] += __shfl_xor_sync(0xffffffffu, g[i], off);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            bmn[i] = min(bmn[i], __shfl_xor_sync(0xffffffffu, bmn[i], off));
            bmx[i] = max(bmx[i], __shfl_xor_sync(0xffffffffu, bmx[i], off));
        }
    }
    if (lane == 0 && g[0]) {
        for (int i = 0; i < 10; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&acc[i]), (unsigned long long)g[i]);
        for (int i = 0; i < 3; ++i) { atomicMin(&accmn[i], bmn[i]); atomicMax(&accmx[i], bmx[i]); }
    }
    __syncthreads();
    if (tid == 0 && acc_label != TA_EMPTY32) mk::global_apply(lt, status, acc_label, acc, accmn, accmx);
}

}  // namespace pp
}  // namespace ta

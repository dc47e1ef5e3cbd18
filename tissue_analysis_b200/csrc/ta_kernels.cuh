// Auxiliary kernels of libtissue_b200: table init / compaction / merge, batched inertia eigen-solve,
// u32 max-label reduction, synthetic Voronoi generator.
#pragma once
#include "ta_common.cuh"

namespace ta {

constexpr int REC_WORDS = 9;   // lo, hi, faces[6], wall18

__global__ void init_label_table_kernel(LabelTable lt) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t n = lt.nrows;
    if (i < n) lt.count[i] = 0;
    if (i < n * 3) { lt.s1[i] = 0; lt.bmin[i] = 0x7FFFFFFF; lt.bmax[i] = -1; }
    if (i < n * 6) lt.s2[i] = 0;
}

__global__ void max_label_kernel(const uint32_t* __restrict__ v, size_t n, unsigned int* out) {
    unsigned int m = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = max(m, v[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// occupied hash slots -> (key, slot) lists, unordered
__global__ void compact_pairs_kernel(PairTable pt, u64* keys_out, uint32_t* slots_out, unsigned int* n_out) {
    size_t cap = (size_t)pt.cap_mask + 1;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
        u64 k = pt.keys[i];
        if (k != TA_EMPTY64) {
            unsigned int idx = atomicAdd(n_out, 1u);
            keys_out[idx] = k;
            slots_out[idx] = (uint32_t)i;
        }
    }
}

// sorted (key, slot) -> packed records
__global__ void gather_records_kernel(PairTable pt, const u64* keys, const uint32_t* slots, unsigned int n,
                                      uint32_t* rec) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 k = keys[i];
    const uint32_t* v = &pt.vals[(size_t)slots[i] * TA_PAIR_STRIDE];
    uint32_t* r = &rec[(size_t)i * REC_WORDS];
    r[0] = (uint32_t)(k >> 32);
    r[1] = (uint32_t)k;
#pragma unroll
    for (int f = 0; f < 7; ++f) r[2 + f] = v[f];
}

// The same without knowing n on the host (deferred pass): n comes from device memory, the records start at row 1 and row 0
// is a header {rows written, 0, ...}.  More than `cap_rows` records: the overflow flag status[3] and a clipped header.
__global__ void gather_records_deferred_kernel(PairTable pt, const u64* keys, const uint32_t* slots, const unsigned int* n_dev,
                                               unsigned int cap_rows, uint32_t* rec) {
    const unsigned int n_all = *n_dev, n = n_all < cap_rows ? n_all : cap_rows;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rec[0] = n;
#pragma unroll
        for (int f = 1; f < REC_WORDS; ++f) rec[f] = 0u;
        if (n_all > cap_rows) atomicExch(&pt.status[3], 1u);
    }
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u64 k = keys[i];
        const uint32_t* v = &pt.vals[(size_t)slots[i] * TA_PAIR_STRIDE];
        uint32_t* r = &rec[(size_t)(i + 1) * REC_WORDS];
        r[0] = (uint32_t)(k >> 32);
        r[1] = (uint32_t)k;
#pragma unroll
        for (int f = 0; f < 7; ++f) r[2 + f] = v[f];
    }
}

// sum-merge the gathered records of all ranks ([world][1 + cap_rows][REC_WORDS], header row first) into a cleared hash
__global__ void merge_records_deferred_kernel(PairTable pt, const uint32_t* all, unsigned int cap_rows, int world) {
    const size_t total = (size_t)world * cap_rows;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t rank = i / cap_rows, j = i - rank * cap_rows;
        const uint32_t* base = all + rank * (size_t)(cap_rows + 1) * REC_WORDS;
        if (j >= base[0]) continue;
        const uint32_t* r = base + (j + 1) * REC_WORDS;
        const u64 key = ((u64)r[0] << 32) | r[1];
        const int slot = ta_pair_slot(pt, key);
        if (slot < 0) continue;
#pragma unroll
        for (int f = 0; f < 7; ++f)
            if (r[2 + f]) atomicAdd(&pt.vals[(size_t)slot * TA_PAIR_STRIDE + f], r[2 + f]);
    }
}

// fetch layouts, made on the device so that the host side is plain copies into the caller's arrays
__global__ void interleave_boxes_kernel(const int* __restrict__ bmin, const int* __restrict__ bmax, size_t n, int* __restrict__ bbox) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n * 3; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i / 3, a = i - row * 3;
        bbox[row * 6 + a] = bmin[i];
        bbox[row * 6 + 3 + a] = bmax[i];
    }
}
// packed records [n][9] -> lo[n] | hi[n] | faces[n][6] | wall18[n] in one scratch block
__global__ void split_records_kernel(const uint32_t* __restrict__ rec, size_t n, uint32_t* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t* r = rec + i * REC_WORDS;
        out[i] = r[0];
        out[n + i] = r[1];
#pragma unroll
        for (int f = 0; f < 6; ++f) out[2 * n + i * 6 + f] = r[2 + f];
        out[8 * n + i] = r[8];
    }
}

// sum-merge packed records (from all ranks) into a cleared hash
__global__ void merge_records_kernel(PairTable pt, const uint32_t* rec, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* r = &rec[i * REC_WORDS];
    u64 key = ((u64)r[0] << 32) | r[1];
    int slot = ta_pair_slot(pt, key);
    if (slot < 0) return;
#pragma unroll
    for (int f = 0; f < 7; ++f)
        if (r[2 + f]) atomicAdd(&pt.vals[(size_t)slot * TA_PAIR_STRIDE + f], r[2 + f]);
}

// ---- batched 3x3 symmetric eigen-solve (cyclic Jacobi, fp64) ----------------------------------------------
__device__ __forceinline__ void jacobi_rotate(double a[3][3], double v[3][3], int p, int q) {
    double apq = a[p][q];
    if (apq == 0.0) return;
    double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
    double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
    double app = a[p][p], aqq = a[q][q];
    a[p][p] = app - t * apq;
    a[q][q] = aqq + t * apq;
    a[p][q] = a[q][p] = 0.0;
    int r = 3 - p - q;
    double arp = a[r][p], arq = a[r][q];
    a[r][p] = a[p][r] = c * arp - s * arq;
    a[r][q] = a[q][r] = s * arp + c * arq;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double vkp = v[k][p], vkq = v[k][q];
        v[k][p] = c * vkp - s * vkq;
        v[k][q] = s * vkp + c * vkq;
    }
}

// cov6 = (a00,a01,a02,a11,a12,a22) -> evals descending, evecs rows (sign: largest |component| positive)
__device__ __forceinline__ void eig3_sym(const double* cov6, double* evals, double* evecs) {
    double a[3][3] = {{cov6[0], cov6[1], cov6[2]}, {cov6[1], cov6[3], cov6[4]}, {cov6[2], cov6[4], cov6[5]}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double scale = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]) + fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    for (int sweep = 0; sweep < 32; ++sweep) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off <= 1e-300 || off <= scale * 1e-22) break;
        jacobi_rotate(a, v, 0, 1);
        jacobi_rotate(a, v, 0, 2);
        jacobi_rotate(a, v, 1, 2);
    }
    int idx[3] = {0, 1, 2};
    double w[3] = {a[0][0], a[1][1], a[2][2]};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2 - i; ++j)
            if (w[idx[j]] < w[idx[j + 1]]) { int t = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = t; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int c = idx[i];
        evals[i] = w[c];
        double x = v[0][c], y = v[1][c], z = v[2][c];
        double ax = fabs(x), ay = fabs(y), az = fabs(z);
        double big = (ax >= ay && ax >= az) ? x : (ay >= az ? y : z);
        double sg = big < 0.0 ? -1.0 : 1.0;
        evecs[i * 3 + 0] = sg * x; evecs[i * 3 + 1] = sg * y; evecs[i * 3 + 2] = sg * z;
    }
}

__device__ __forceinline__ double i128_to_double(__int128 x) {
    bool neg = x < 0;
    unsigned __int128 u = neg ? (unsigned __int128)(-x) : (unsigned __int128)x;
    double d = (double)(u64)(u >> 64) * 18446744073709551616.0 + (double)(u64)u;
    return neg ? -d : d;
}

// covariance from exact integer sums: C_ab = (n*S_ab - S_a*S_b) / (n * max(3, n))      (SIA:137-150)
__global__ void inertia_from_moments_kernel(LabelTable lt, const uint32_t* labels, size_t n, double* evals,
                                            double* evecs) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t L = labels ? labels[i] : i;
    double cov[6];
    u64 cnt = (L < lt.nrows) ? lt.count[L] : 0;
    if (cnt == 0) {
        for (int k = 0; k < 3; ++k) evals[i * 3 + k] = nan("");
        for (int k = 0; k < 9; ++k) evecs[i * 9 + k] = nan("");
        return;
    }
    const u64* s1 = &lt.s1[L * 3];
    const u64* s2 = &lt.s2[L * 6];
    const int ia[6] = {0, 0, 0, 1, 1, 2}, ib[6] = {0, 1, 2, 1, 2, 2};
    double denom = (double)cnt * (double)(cnt < 3 ? 3 : cnt);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        __int128 num = (__int128)cnt * (__int128)s2[k] - (__int128)s1[ia[k]] * (__int128)s1[ib[k]];
        cov[k] = i128_to_double(num) / denom;
    }
    eig3_sym(cov, &evals[i * 3], &evecs[i * 9]);
}

__global__ void inertia_eig_kernel(const double* cov, size_t n, double* evals, double* evecs) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    eig3_sym(&cov[i * 6], &evals[i * 3], &evecs[i * 9]);
}

// ---- synthetic Voronoi tissue (bench / test utility) ---------------------------------------------------------
// Seeds are binned into a coarse grid (cell lists); every voxel searches growing shells of bins until the best
// squared distance cannot be beaten by an unexplored shell.  Pure integer arithmetic: identical to the numpy
// generator in tissue_analysis_b200/synth.py, ties broken towards the lower seed index.
struct SynthParams {
    long long nf, nm, ns, slow_offset, global_slow;
    int gx, gy, gz;            // bins per axis
    int bin;                   // bin edge in weighted fixed-point units
    int wf, wm, ws;            // axis weights
    int dome;
    long long kf, km, ks;      // dome coefficients: inside iff uf^2*kf + um^2*km + us^2*ks <= 2^40
    uint32_t ncell;
};

template <typename T>
__global__ void synth_voronoi_kernel(T* out, SynthParams P, const int* __restrict__ bin_start,
                                     const int* __restrict__ bin_seed, const int* __restrict__ seeds) {
    size_t total = (size_t)P.nf * P.nm * P.ns;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        long long f = i % P.nf, m = (i / P.nf) % P.nm, s = i / (P.nf * P.nm) + P.slow_offset;
        if (P.dome) {
            long long uf = 2 * f + 1 - P.nf, um = 2 * m + 1 - P.nm, us = 2 * s + 1 - P.global_slow;
            if (uf * uf * P.kf + um * um * P.km + us * us * P.ks > (1LL << 40)) { out[i] = (T)1; continue; }
        }
        // weighted fixed-point position of the voxel centre
        long long pf = (16 * f + 8) * P.wf, pm = (16 * m + 8) * P.wm, ps = (16 * s + 8) * P.ws;
        int cf = (int)(pf / P.bin), cm = (int)(pm / P.bin), cs = (int)(ps / P.bin);
        long long best = 0x7FFFFFFFFFFFFFFFLL;
        int besti = 0x7FFFFFFF;
        int maxr = max(P.gx, max(P.gy, P.gz));
        for (int r = 0; r <= maxr; ++r) {
            // shell r: bins with Chebyshev distance exactly r from (cf, cm, cs)
            for (int dz = -r; dz <= r; ++dz) {
                int z = cs + dz; if (z < 0 || z >= P.gz) continue;
                for (int dy = -r; dy <= r; ++dy) {
                    int y = cm + dy; if (y < 0 || y >= P.gy) continue;
                    bool face = (dz == -r || dz == r || dy == -r || dy == r);
                    for (int dx = -r; dx <= r; dx += (face ? 1 : 2 * r > 0 ? 2 * r : 1)) {
                        int x = cf + dx; if (x < 0 || x >= P.gx) continue;
                        int b = (z * P.gy + y) * P.gx + x;
                        for (int q = bin_start[b]; q < bin_start[b + 1]; ++q) {
                            int si = bin_seed[q];
                            long long df = pf - (long long)seeds[si * 3 + 0] * P.wf;
                            long long dm = pm - (long long)seeds[si * 3 + 1] * P.wm;
                            long long ds = ps - (long long)seeds[si * 3 + 2] * P.ws;
                            long long d = df * df + dm * dm + ds * ds;
                            if (d < best || (d == best && si < besti)) { best = d; besti = si; }
                        }
                    }
                }
            }
            // any seed in shell r+1 or beyond is at least r*bin away along some axis
            long long reach = (long long)r * P.bin;
            if (besti != 0x7FFFFFFF && best < reach * reach) break;
        }
        out[i] = (T)(besti + 2);
    }
}

}  // namespace ta

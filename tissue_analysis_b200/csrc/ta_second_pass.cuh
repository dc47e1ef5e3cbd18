// Optional second passes over the bound volume: wall-voxel coordinates for a list of label pairs
// (reference: wall_voxels_per_cell, spatial_image_analysis.py:835-863) and the first voxel layer image
// (reference: __voxel_first_layer, spatial_image_analysis.py:1024-1038).
#pragma once
#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/tissue_b200.h"
#include "ta_common.cuh"

namespace ta {

struct VolDims {
    long long nf, nm, ns, own_lo, own_hi, slow_offset;
};

template <typename T>
__device__ __forceinline__ uint32_t vox_at(const T* v, const VolDims& D, long long f, long long m, long long s) {
    f = min(max(f, 0LL), D.nf - 1);
    m = min(max(m, 0LL), D.nm - 1);
    s = min(max(s, 0LL), D.ns - 1);
    return v[(s * D.nm + m) * D.nf + f];
}

__device__ __forceinline__ long long find_pair(const u64* keys, long long n, u64 key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < n && keys[lo] == key) ? lo : -1;
}

// One voxel of the wall-voxel passes: its distinct other labels among the 18 neighbours, looked up in the sorted pair keys.
// mode 0: count per pair; mode 1: append composite keys (pair rank * nvox_global + linear index)
template <typename T, int MODE>
__device__ __forceinline__ void wall_voxel_one(const T* __restrict__ vol, const VolDims& D, long long f, long long m, long long s,
                                               const u64* __restrict__ keys, long long nkeys, u64* counts, u64* cursor, u64* records,
                                               u64 lin_span) {
    const uint32_t a = vol[(s * D.nm + m) * D.nf + f];
    uint32_t seen[18];
    int ns = 0;
#pragma unroll 1
    for (int k = 0; k < 27; ++k) {
        int df = k % 3 - 1, dm = (k / 3) % 3 - 1, ds = k / 9 - 1;
        int l1 = abs(df) + abs(dm) + abs(ds);
        if (l1 < 1 || l1 > 2) continue;
        uint32_t b = vox_at(vol, D, f + df, m + dm, s + ds);
        if (b == a) continue;
        bool dup = false;
        for (int q = 0; q < ns; ++q) dup |= (seen[q] == b);
        if (dup) continue;
        seen[ns++] = b;
        long long idx = find_pair(keys, nkeys, ta_pair_key(a, b));
        if (idx < 0) continue;
        if (MODE == 0) {
            atomicAdd(&counts[idx], 1ull);
        } else {
            u64 pos = atomicAdd(cursor, 1ull);
            u64 lin = (u64)((s + D.slow_offset) * D.nm + m) * D.nf + f;
            records[pos] = (u64)idx * lin_span + lin;
        }
    }
}

template <typename T, int MODE>
__global__ void wall_voxels_kernel(const T* __restrict__ vol, VolDims D, const u64* __restrict__ keys,
                                   long long nkeys, u64* counts, u64* cursor, u64* records, u64 lin_span) {
    const long long owned = (D.own_hi - D.own_lo) * D.nm * D.nf;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < owned;
         i += (long long)gridDim.x * blockDim.x) {
        long long f = i % D.nf, m = (i / D.nf) % D.nm, s = i / (D.nf * D.nm) + D.own_lo;
        wall_voxel_one<T, MODE>(vol, D, f, m, s, keys, nkeys, counts, cursor, records, lin_span);
    }
}

// sorted composite keys -> three coordinate rows per pair block
__global__ void decode_wall_voxels_kernel(const u64* __restrict__ recs, u64 n, u64 lin_span, long long nf,
                                          long long nm, const u64* __restrict__ offsets,
                                          const u64* __restrict__ counts, long long* xyz) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 r = recs[i];
    u64 idx = r / lin_span, lin = r % lin_span;
    u64 off = offsets[idx], cnt = counts[idx], j = i - off;
    long long* base = xyz + 3 * off;
    base[j] = (long long)(lin % nf);
    base[cnt + j] = (long long)((lin / nf) % nm);
    base[2 * cnt + j] = (long long)(lin / ((u64)nf * nm));
}

// ---- row windows: the vectorised form of the stencil passes ------------------------------------------------------------------
// A thread takes one 16-byte vector of a row (V voxels) and looks at its neighbourhood as whole rows: the same vector of the
// rows around it, plus one voxel left and right where the stencil shifts in f.  Nine vector loads per V voxels instead of 18
// scalar ones per voxel, no 64-bit division per voxel; neighbouring threads share every line through L1 / L2.  Needs rows of
// whole vectors (nf a multiple of V, 16-byte aligned base); the per-voxel kernels above remain for everything else.
template <typename T> struct Row {
    static constexpr int V = 16 / (int)sizeof(T);
    uint32_t w[V + 2];            // w[0] = the voxel left of the vector (the edge voxel itself at f = 0), w[1 .. V], w[V + 1] right
};
template <typename T>
__device__ __forceinline__ void row_load(const T* __restrict__ rowp, int f0, int nf, bool edges, Row<T>& R) {
    constexpr int V = Row<T>::V;
    const uint4 v = *reinterpret_cast<const uint4*>(rowp + f0);
    if (sizeof(T) == 2) {
        const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { R.w[1 + 2 * k] = x[k] & 0xFFFFu; R.w[2 + 2 * k] = x[k] >> 16; }
    } else {
        R.w[1] = v.x; R.w[2] = v.y; R.w[3] = v.z; R.w[4] = v.w;
    }
    R.w[0] = R.w[1]; R.w[V + 1] = R.w[V];
    if (edges) {
        if (f0 > 0) R.w[0] = rowp[f0 - 1];
        if (f0 + V < nf) R.w[V + 1] = rowp[f0 + V];
    }
}
template <typename T>
__device__ __forceinline__ void row_store(T* __restrict__ rowp, int f0, const uint32_t* r) {
    uint4 v;
    if (sizeof(T) == 2) {
        v.x = r[0] | (r[1] << 16); v.y = r[2] | (r[3] << 16); v.z = r[4] | (r[5] << 16); v.w = r[6] | (r[7] << 16);
    } else {
        v.x = r[0]; v.y = r[1]; v.z = r[2]; v.w = r[3];
    }
    *reinterpret_cast<uint4*>(rowp + f0) = v;
}

// KIND 0 / 2: hollow_out_cells / its 0-1 mask (wrap-around Laplacian, reflect border); KIND 1: 18-connected outer shell;
// KIND 3: voxel_first_layer (cells' voxels with a background 6-neighbour; bg / keep_bg as in voxel_first_layer_kernel)
template <typename T, int KIND>
__global__ void __launch_bounds__(256) stencil_rows_kernel(const T* __restrict__ vol, T* __restrict__ out, VolDims D, uint32_t bg,
                                                           int keep_bg) {
    constexpr int V = Row<T>::V;
    const int nf = (int)D.nf, nm = (int)D.nm, ns = (int)D.ns, nvr = nf / V;
    const unsigned long long total = (unsigned long long)nvr * nm * ns;
    const size_t plane = (size_t)nf * nm;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const int fv = (int)(i % (unsigned)nvr);
        const unsigned long long rest = i / (unsigned)nvr;
        const int m = (int)(rest % (unsigned)nm), s = (int)(rest / (unsigned)nm), f0 = fv * V;
        const T* c_row = vol + (size_t)s * plane + (size_t)m * nf;
        Row<T> c;
        row_load(c_row, f0, nf, true, c);
        uint32_t res[V];
        if (KIND == 1) {
            // out of the volume counts as another label
            bool outside[V + 2];
#pragma unroll
            for (int k = 0; k < V + 2; ++k) outside[k] = false;
            outside[0] = f0 == 0; outside[V + 1] = f0 + V >= nf;
            uint32_t diff[V];
#pragma unroll
            for (int k = 0; k < V; ++k) diff[k] = (c.w[k] ^ c.w[k + 1]) | (c.w[k + 2] ^ c.w[k + 1]) | (outside[k] ? 1u : 0u) | (outside[k + 2] ? 1u : 0u);
#pragma unroll
            for (int dm = -1; dm <= 1; ++dm)
#pragma unroll
                for (int ds = -1; ds <= 1; ++ds) {
                    if (dm == 0 && ds == 0) continue;
                    const int mm = m + dm, ss = s + ds;
                    if (mm < 0 || mm >= nm || ss < 0 || ss >= ns) {
#pragma unroll
                        for (int k = 0; k < V; ++k) diff[k] |= 1u;
                        continue;
                    }
                    Row<T> r;
                    const bool shifted = (dm == 0) != (ds == 0);           // face rows bring their f - 1 / f + 1 voxels too
                    row_load(vol + (size_t)ss * plane + (size_t)mm * nf, f0, nf, shifted, r);
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        diff[k] |= r.w[k + 1] ^ c.w[k + 1];
                        if (shifted) diff[k] |= (outside[k] ? 0u : (r.w[k] ^ c.w[k + 1])) | (outside[k + 2] ? 0u : (r.w[k + 2] ^ c.w[k + 1]));
                    }
                }
#pragma unroll
            for (int k = 0; k < V; ++k) res[k] = diff[k] ? 1u : 0u;
        } else {
            // the four face rows; a row outside the volume is the centre row (reflect: no contribution; first layer: no background)
            Row<T> mlo, mhi, slo, shi;
            row_load(m > 0 ? c_row - nf : c_row, f0, nf, false, mlo);
            row_load(m + 1 < nm ? c_row + nf : c_row, f0, nf, false, mhi);
            row_load(s > 0 ? c_row - plane : c_row, f0, nf, false, slo);
            row_load(s + 1 < ns ? c_row + plane : c_row, f0, nf, false, shi);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const uint32_t a = c.w[k + 1];
                if (KIND == 3) {
                    const bool touch = c.w[k] == bg || c.w[k + 2] == bg || mlo.w[k + 1] == bg || mhi.w[k + 1] == bg || slo.w[k + 1] == bg ||
                                       shi.w[k + 1] == bg;
                    res[k] = a == bg ? (keep_bg ? 1u : 0u) : (touch ? a : 0u);
                } else {
                    // scipy.ndimage.laplace in the voxel type: per axis x[i-1] - 2 x[i] + x[i+1], sums wrap
                    const T sum = (T)((T)(c.w[k] + c.w[k + 2] - 2u * a) + (T)(mlo.w[k + 1] + mhi.w[k + 1] - 2u * a) +
                                      (T)(slo.w[k + 1] + shi.w[k + 1] - 2u * a));
                    res[k] = sum != 0 ? (KIND == 2 ? 1u : a) : 0u;
                }
            }
        }
        row_store(out + (size_t)s * plane + (size_t)m * nf, f0, res);
    }
}

// bit k of the result: voxel k of the vector has a different label among its 18 neighbours (clamped at the volume's faces, as
// vox_at does).  The wall-voxel passes look at single voxels only where this says so.
template <typename T>
__device__ __forceinline__ uint32_t row_wall_bits(const T* __restrict__ vol, const VolDims& D, int f0, int m, long long s) {
    constexpr int V = Row<T>::V;
    const int nf = (int)D.nf, nm = (int)D.nm;
    const size_t plane = (size_t)nf * nm;
    Row<T> c;
    row_load(vol + (size_t)s * plane + (size_t)m * nf, f0, nf, true, c);
    uint32_t diff[V];
#pragma unroll
    for (int k = 0; k < V; ++k) diff[k] = (c.w[k] ^ c.w[k + 1]) | (c.w[k + 2] ^ c.w[k + 1]);
#pragma unroll
    for (int dm = -1; dm <= 1; ++dm)
#pragma unroll
        for (int ds = -1; ds <= 1; ++ds) {
            if (dm == 0 && ds == 0) continue;
            const int mm = min(max(m + dm, 0), nm - 1);
            const long long ss = min(max(s + ds, 0LL), D.ns - 1);
            Row<T> r;
            const bool shifted = (dm == 0) != (ds == 0);
            row_load(vol + (size_t)ss * plane + (size_t)mm * nf, f0, nf, shifted, r);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                diff[k] |= r.w[k + 1] ^ c.w[k + 1];
                if (shifted) diff[k] |= (r.w[k] ^ c.w[k + 1]) | (r.w[k + 2] ^ c.w[k + 1]);
            }
        }
    uint32_t bits = 0u;
#pragma unroll
    for (int k = 0; k < V; ++k) bits |= diff[k] ? (1u << k) : 0u;
    return bits;
}

// the same passes over rows of whole vectors: single voxels are looked at only where row_wall_bits finds another label around
// label_bits: 2^16 bits, bit (label & 0xFFFF) set for every label of a requested pair (exact for uint16 volumes, a filter for
// uint32 ones): a wall voxel of any other label -- nearly all of them when a few pairs are asked for -- is dropped at once
template <typename T, int MODE>
__global__ void __launch_bounds__(256) wall_voxels_rows_kernel(const T* __restrict__ vol, VolDims D, const u64* __restrict__ keys,
                                                               long long nkeys, u64* counts, u64* cursor, u64* records, u64 lin_span,
                                                               const uint32_t* __restrict__ label_bits) {
    constexpr int V = Row<T>::V;
    const int nf = (int)D.nf, nm = (int)D.nm, nvr = nf / V;
    const unsigned long long total = (unsigned long long)nvr * nm * (unsigned long long)(D.own_hi - D.own_lo);
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const int fv = (int)(i % (unsigned)nvr);
        const unsigned long long rest = i / (unsigned)nvr;
        const int m = (int)(rest % (unsigned)nm), f0 = fv * V;
        const long long s = (long long)(rest / (unsigned)nm) + D.own_lo;
        for (uint32_t bits = row_wall_bits(vol, D, f0, m, s); bits; bits &= bits - 1u) {
            const int f = f0 + __ffs(bits) - 1;
            const uint32_t a = vol[((size_t)s * nm + m) * (size_t)nf + f] & 0xFFFFu;
            if ((label_bits[a >> 5] >> (a & 31u)) & 1u)
                wall_voxel_one<T, MODE>(vol, D, f, m, s, keys, nkeys, counts, cursor, records, lin_span);
        }
    }
}

#define TA2_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            *err = std::string(#call) + " -> " + cudaGetErrorString(e_);                   \
            return TA_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

// the row-window kernels want rows that are whole, aligned 16-byte vectors
inline bool rows_are_vectors(const void* vol, int elem, long long nf) {
    return nf > 0 && (nf * elem) % 16 == 0 && ((uintptr_t)vol & 15u) == 0 && nf < 0x7FFFFFF0LL;
}

inline int wall_voxel_coords_impl(const void* vol, int elem, long long nf, long long nm, long long ns,
                                  long long own_lo, long long own_hi, long long slow_offset, const uint32_t* lo,
                                  const uint32_t* hi, uint64_t npairs, uint64_t* counts, int64_t* xyz,
                                  cudaStream_t st, int num_sms, uint64_t* launches, std::string* err) {
    // sorted unique keys; rank -> caller index
    std::vector<size_t> order(npairs);
    std::iota(order.begin(), order.end(), 0);
    std::vector<u64> key(npairs);
    for (size_t i = 0; i < npairs; ++i) key[i] = ta_pair_key(lo[i], hi[i]);
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return key[a] < key[b]; });
    std::vector<u64> skeys;
    std::vector<size_t> rank_of(npairs);
    for (size_t r = 0; r < npairs; ++r) {
        size_t i = order[r];
        if (skeys.empty() || skeys.back() != key[i]) skeys.push_back(key[i]);
        rank_of[i] = skeys.size() - 1;
    }
    const size_t nk = skeys.size();
    VolDims D{nf, nm, ns, own_lo, own_hi, slow_offset};
    TaDevBuf b_keys, b_counts, b_rec0, b_rec1, b_off, b_xyz, b_tmp;        // freed on every way out
    TA2_CUDA(cudaMalloc(&b_keys.p, nk * sizeof(u64)));
    TA2_CUDA(cudaMalloc(&b_counts.p, (nk + 1) * sizeof(u64)));
    u64 *d_keys = b_keys.as<u64>(), *d_counts = b_counts.as<u64>(), *d_cursor = d_counts + nk;
    TA2_CUDA(cudaMemcpyAsync(d_keys, skeys.data(), nk * sizeof(u64), cudaMemcpyHostToDevice, st));
    TA2_CUDA(cudaMemsetAsync(d_counts, 0, (nk + 1) * sizeof(u64), st));
    const int grid = num_sms * 16;
    // the composite key needs the GLOBAL linear span; planes above this slab only enlarge it harmlessly
    const u64 lin_span = (u64)(slow_offset + ns) * (u64)nm * (u64)nf;
    const bool rows = rows_are_vectors(vol, elem, nf);
    TaDevBuf b_bits;
    uint32_t* d_bits = nullptr;
    if (rows) {
        std::vector<uint32_t> bits(2048, 0u);
        for (size_t i = 0; i < npairs; ++i) {
            bits[(lo[i] & 0xFFFFu) >> 5] |= 1u << (lo[i] & 31u);
            bits[(hi[i] & 0xFFFFu) >> 5] |= 1u << (hi[i] & 31u);
        }
        TA2_CUDA(cudaMalloc(&b_bits.p, bits.size() * sizeof(uint32_t)));
        d_bits = b_bits.as<uint32_t>();
        TA2_CUDA(cudaMemcpy(d_bits, bits.data(), bits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (elem == 2) {
        if (rows) wall_voxels_rows_kernel<uint16_t, 0><<<grid, 256, 0, st>>>((const uint16_t*)vol, D, d_keys, (long long)nk, d_counts, d_cursor, nullptr, lin_span, d_bits);
        else wall_voxels_kernel<uint16_t, 0><<<grid, 256, 0, st>>>((const uint16_t*)vol, D, d_keys, (long long)nk, d_counts, d_cursor, nullptr, lin_span);
    } else {
        if (rows) wall_voxels_rows_kernel<uint32_t, 0><<<grid, 256, 0, st>>>((const uint32_t*)vol, D, d_keys, (long long)nk, d_counts, d_cursor, nullptr, lin_span, d_bits);
        else wall_voxels_kernel<uint32_t, 0><<<grid, 256, 0, st>>>((const uint32_t*)vol, D, d_keys, (long long)nk, d_counts, d_cursor, nullptr, lin_span);
    }
    (*launches)++;
    std::vector<u64> kcount(nk);
    TA2_CUDA(cudaMemcpyAsync(kcount.data(), d_counts, nk * sizeof(u64), cudaMemcpyDeviceToHost, st));
    TA2_CUDA(cudaStreamSynchronize(st));
    for (size_t i = 0; i < npairs; ++i) counts[i] = kcount[rank_of[i]];
    if (!xyz) return TA_OK;

    // unique-key blocks in rank order on the device; caller blocks follow the caller's pair order
    std::vector<u64> koff(nk + 1, 0);
    for (size_t r = 0; r < nk; ++r) koff[r + 1] = koff[r] + kcount[r];
    const u64 total = koff[nk];
    int rc = TA_OK;
    if (total > 0) {
        if ((double)nk * (double)lin_span > 9.0e18) { *err = "too many pairs x voxels for one coordinate pass"; rc = TA_ERR_BAD_ARG; }
        if (!rc) {
            TA2_CUDA(cudaMalloc(&b_rec0.p, total * sizeof(u64)));
            TA2_CUDA(cudaMalloc(&b_rec1.p, total * sizeof(u64)));
            TA2_CUDA(cudaMalloc(&b_off.p, (nk + 1) * sizeof(u64)));
            TA2_CUDA(cudaMalloc(&b_xyz.p, total * 3 * sizeof(long long)));
            u64* d_rec[2] = {b_rec0.as<u64>(), b_rec1.as<u64>()};
            u64* d_off = b_off.as<u64>();
            long long* d_xyz = b_xyz.as<long long>();
            TA2_CUDA(cudaMemcpyAsync(d_off, koff.data(), (nk + 1) * sizeof(u64), cudaMemcpyHostToDevice, st));
            if (elem == 2) {
                if (rows) wall_voxels_rows_kernel<uint16_t, 1><<<grid, 256, 0, st>>>((const uint16_t*)vol, D, d_keys, (long long)nk, d_counts, d_cursor, d_rec[0], lin_span, d_bits);
                else wall_voxels_kernel<uint16_t, 1><<<grid, 256, 0, st>>>((const uint16_t*)vol, D, d_keys, (long long)nk, d_counts, d_cursor, d_rec[0], lin_span);
            } else {
                if (rows) wall_voxels_rows_kernel<uint32_t, 1><<<grid, 256, 0, st>>>((const uint32_t*)vol, D, d_keys, (long long)nk, d_counts, d_cursor, d_rec[0], lin_span, d_bits);
                else wall_voxels_kernel<uint32_t, 1><<<grid, 256, 0, st>>>((const uint32_t*)vol, D, d_keys, (long long)nk, d_counts, d_cursor, d_rec[0], lin_span);
            }
            (*launches)++;
            size_t need = 0;
            TA2_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, need, d_rec[0], d_rec[1], (long long)total, 0, 64, st));
            TA2_CUDA(cudaMalloc(&b_tmp.p, need));
            TA2_CUDA(cub::DeviceRadixSort::SortKeys(b_tmp.p, need, d_rec[0], d_rec[1], (long long)total, 0, 64, st));
            decode_wall_voxels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_rec[1], total, lin_span, nf, nm,
                                                                                       d_off, d_counts, d_xyz);
            (*launches)++;
            std::vector<long long> host(total * 3);
            TA2_CUDA(cudaMemcpyAsync(host.data(), d_xyz, total * 3 * sizeof(long long), cudaMemcpyDeviceToHost, st));
            TA2_CUDA(cudaStreamSynchronize(st));
            u64 out_off = 0;
            for (size_t i = 0; i < npairs; ++i) {
                size_t r = rank_of[i];
                std::copy(host.begin() + 3 * koff[r], host.begin() + 3 * koff[r + 1], xyz + 3 * out_off);
                out_off += kcount[r];
            }
        }
    }
    return rc;
}

// out[i] = lut[vol[i]] (labels >= n_lut -> fill).  A thread takes UNR 16-byte vectors of input voxels per round (all loads
// first), looks a label up only where it differs from the voxel before it (runs of one label are the rule in a tissue) and
// stores whole vectors.  SMEM: the table is copied to shared memory first (it fits for uint16 labels: 128 / 256 KB are 65536
// entries of 2 / 4 bytes -- the latter does not, and goes through L1 like every uint32-label table).
template <typename TI, typename TO, bool SMEM>
__global__ void __launch_bounds__(1024) map_labels_kernel(const TI* __restrict__ vol, TO* __restrict__ out, const TO* __restrict__ lut,
                                                          unsigned long long n_lut, TO fill, size_t n) {
    extern __shared__ __align__(16) unsigned char lut_raw[];
    const TO* table = lut;
    if (SMEM) {
        TO* t = reinterpret_cast<TO*>(lut_raw);
        for (unsigned long long i = threadIdx.x; i < n_lut; i += blockDim.x) t[i] = lut[i];
        __syncthreads();
        table = t;
    }
    constexpr int V = 16 / sizeof(TI);
    constexpr int UNR = 4;
    const size_t nvec = n / V, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * UNR) {
        uint4 in[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
            if (i0 + u * stride < nvec) in[u] = *reinterpret_cast<const uint4*>(vol + (i0 + u * stride) * V);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const size_t i = i0 + u * stride;
            if (i >= nvec) break;
            const TI* lab = reinterpret_cast<const TI*>(&in[u]);
            __align__(16) TO res[V];
#pragma unroll
            for (int k = 0; k < V; ++k)
                res[k] = (k > 0 && lab[k] == lab[k - 1]) ? res[k - 1] : (lab[k] < n_lut ? table[lab[k]] : fill);
            TO* o = out + i * V;
            if (sizeof(TO) * V == 16) {
                *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(res);
            } else if (sizeof(TO) * V == 32) {
                reinterpret_cast<uint4*>(o)[0] = reinterpret_cast<const uint4*>(res)[0];
                reinterpret_cast<uint4*>(o)[1] = reinterpret_cast<const uint4*>(res)[1];
            } else {
                *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(res);
            }
        }
    }
    for (size_t i = nvec * V + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = vol[i] < n_lut ? table[vol[i]] : fill;
}

template <typename T>
__global__ void voxel_first_layer_kernel(const T* __restrict__ vol, T* __restrict__ out, VolDims D, uint32_t bg,
                                         int keep_bg) {
    const long long total = D.nf * D.nm * D.ns;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long f = i % D.nf, m = (i / D.nf) % D.nm, s = i / (D.nf * D.nm);
        uint32_t a = vol[i];
        uint32_t r = 0;
        if (a == bg) {
            r = keep_bg ? 1u : 0u;
        } else {
            bool touch = (f > 0 && vol[i - 1] == bg) || (f + 1 < D.nf && vol[i + 1] == bg) ||
                         (m > 0 && vol[i - D.nf] == bg) || (m + 1 < D.nm && vol[i + D.nf] == bg) ||
                         (s > 0 && vol[i - D.nf * D.nm] == bg) || (s + 1 < D.ns && vol[i + D.nf * D.nm] == bg);
            r = touch ? a : 0u;
        }
        out[i] = (T)r;
    }
}

// kind 0: hollow_out_cells (label where the wrap-around Laplacian is non-zero); kind 2: the same as a 0/1 mask;
// kind 1: 18-connected outer shell
template <typename T, int KIND>
__global__ void stencil_image_kernel(const T* __restrict__ vol, T* __restrict__ out, VolDims D) {
    const long long total = D.nf * D.nm * D.ns;
    const long long plane = D.nf * D.nm;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i % D.nf, m = (i / D.nf) % D.nm, s = i / plane;
        const T c = vol[i];
        if (KIND == 0 || KIND == 2) {
            // scipy.ndimage.laplace: per axis x[i-1] - 2 x[i] + x[i+1], 'reflect' border (x[-1] = x[0]), dtype wrap
            T sum = 0;
            sum += (T)((f > 0 ? vol[i - 1] : c) + (f + 1 < D.nf ? vol[i + 1] : c) - (T)2 * c);
            sum += (T)((m > 0 ? vol[i - D.nf] : c) + (m + 1 < D.nm ? vol[i + D.nf] : c) - (T)2 * c);
            sum += (T)((s > 0 ? vol[i - plane] : c) + (s + 1 < D.ns ? vol[i + plane] : c) - (T)2 * c);
            out[i] = sum != 0 ? (KIND == 2 ? (T)1 : c) : (T)0;
        } else {
            bool shell = false;
#pragma unroll 1
            for (int k = 0; k < 27; ++k) {
                const int df = k % 3 - 1, dm = (k / 3) % 3 - 1, ds = k / 9 - 1;
                const int l1 = abs(df) + abs(dm) + abs(ds);
                if (l1 < 1 || l1 > 2) continue;
                const long long ff = f + df, mm = m + dm, ss = s + ds;
                if (ff < 0 || ff >= D.nf || mm < 0 || mm >= D.nm || ss < 0 || ss >= D.ns) { shell = true; break; }
                if (vol[(ss * D.nm + mm) * D.nf + ff] != c) { shell = true; break; }
            }
            out[i] = shell ? (T)1 : (T)0;
        }
    }
}

inline int stencil_image_impl(const void* vol, int elem, long long nf, long long nm, long long ns, int kind,
                              void* out_host, cudaStream_t st, int num_sms, uint64_t* launches, std::string* err) {
    const size_t bytes = (size_t)nf * nm * ns * elem;
    TaDevBuf out_buf;
    TA2_CUDA(cudaMalloc(&out_buf.p, bytes));
    void* d_out = out_buf.p;
    VolDims D{nf, nm, ns, 0, ns, 0};
    const int grid = num_sms * 16;
    if (rows_are_vectors(vol, elem, nf)) {
        if (elem == 2) {
            const uint16_t* v = (const uint16_t*)vol; uint16_t* o = (uint16_t*)d_out;
            if (kind == 0) stencil_rows_kernel<uint16_t, 0><<<grid, 256, 0, st>>>(v, o, D, 0u, 0);
            else if (kind == 1) stencil_rows_kernel<uint16_t, 1><<<grid, 256, 0, st>>>(v, o, D, 0u, 0);
            else stencil_rows_kernel<uint16_t, 2><<<grid, 256, 0, st>>>(v, o, D, 0u, 0);
        } else {
            const uint32_t* v = (const uint32_t*)vol; uint32_t* o = (uint32_t*)d_out;
            if (kind == 0) stencil_rows_kernel<uint32_t, 0><<<grid, 256, 0, st>>>(v, o, D, 0u, 0);
            else if (kind == 1) stencil_rows_kernel<uint32_t, 1><<<grid, 256, 0, st>>>(v, o, D, 0u, 0);
            else stencil_rows_kernel<uint32_t, 2><<<grid, 256, 0, st>>>(v, o, D, 0u, 0);
        }
    } else if (elem == 2) {
        const uint16_t* v = (const uint16_t*)vol; uint16_t* o = (uint16_t*)d_out;
        if (kind == 0) stencil_image_kernel<uint16_t, 0><<<grid, 256, 0, st>>>(v, o, D);
        else if (kind == 1) stencil_image_kernel<uint16_t, 1><<<grid, 256, 0, st>>>(v, o, D);
        else stencil_image_kernel<uint16_t, 2><<<grid, 256, 0, st>>>(v, o, D);
    } else {
        const uint32_t* v = (const uint32_t*)vol; uint32_t* o = (uint32_t*)d_out;
        if (kind == 0) stencil_image_kernel<uint32_t, 0><<<grid, 256, 0, st>>>(v, o, D);
        else if (kind == 1) stencil_image_kernel<uint32_t, 1><<<grid, 256, 0, st>>>(v, o, D);
        else stencil_image_kernel<uint32_t, 2><<<grid, 256, 0, st>>>(v, o, D);
    }
    (*launches)++;
    TA2_CUDA(cudaMemcpyAsync(out_host, d_out, bytes, cudaMemcpyDeviceToHost, st));
    TA2_CUDA(cudaStreamSynchronize(st));
    return TA_OK;
}

inline int voxel_first_layer_impl(const void* vol, int elem, long long nf, long long nm, long long ns,
                                  uint32_t background, int keep_background, void* out_host, cudaStream_t st,
                                  int num_sms, uint64_t* launches, std::string* err) {
    const size_t bytes = (size_t)nf * nm * ns * elem;
    TaDevBuf out_buf;
    TA2_CUDA(cudaMalloc(&out_buf.p, bytes));
    void* d_out = out_buf.p;
    VolDims D{nf, nm, ns, 0, ns, 0};
    if (rows_are_vectors(vol, elem, nf)) {
        if (elem == 2) stencil_rows_kernel<uint16_t, 3><<<num_sms * 16, 256, 0, st>>>((const uint16_t*)vol, (uint16_t*)d_out, D, background, keep_background);
        else stencil_rows_kernel<uint32_t, 3><<<num_sms * 16, 256, 0, st>>>((const uint32_t*)vol, (uint32_t*)d_out, D, background, keep_background);
    } else if (elem == 2)
        voxel_first_layer_kernel<uint16_t><<<num_sms * 16, 256, 0, st>>>((const uint16_t*)vol, (uint16_t*)d_out, D,
                                                                         background, keep_background);
    else
        voxel_first_layer_kernel<uint32_t><<<num_sms * 16, 256, 0, st>>>((const uint32_t*)vol, (uint32_t*)d_out, D,
                                                                         background, keep_background);
    (*launches)++;
    TA2_CUDA(cudaMemcpyAsync(out_host, d_out, bytes, cudaMemcpyDeviceToHost, st));
    TA2_CUDA(cudaStreamSynchronize(st));
    return TA_OK;
}

}  // namespace ta

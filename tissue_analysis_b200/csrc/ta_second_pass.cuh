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

// mode 0: count per pair (+ total); mode 1: append composite keys (pair rank * nvox_global + linear index)
template <typename T, int MODE>
__global__ void wall_voxels_kernel(const T* __restrict__ vol, VolDims D, const u64* __restrict__ keys,
                                   long long nkeys, u64* counts, u64* cursor, u64* records, u64 lin_span) {
    const long long owned = (D.own_hi - D.own_lo) * D.nm * D.nf;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < owned;
         i += (long long)gridDim.x * blockDim.x) {
        long long f = i % D.nf, m = (i / D.nf) % D.nm, s = i / (D.nf * D.nm) + D.own_lo;
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

#define TA2_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            *err = std::string(#call) + " -> " + cudaGetErrorString(e_);                   \
            return TA_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

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
    if (elem == 2)
        wall_voxels_kernel<uint16_t, 0><<<grid, 256, 0, st>>>((const uint16_t*)vol, D, d_keys, (long long)nk, d_counts,
                                                              d_cursor, nullptr, lin_span);
    else
        wall_voxels_kernel<uint32_t, 0><<<grid, 256, 0, st>>>((const uint32_t*)vol, D, d_keys, (long long)nk, d_counts,
                                                              d_cursor, nullptr, lin_span);
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
            if (elem == 2)
                wall_voxels_kernel<uint16_t, 1><<<grid, 256, 0, st>>>((const uint16_t*)vol, D, d_keys, (long long)nk,
                                                                      d_counts, d_cursor, d_rec[0], lin_span);
            else
                wall_voxels_kernel<uint32_t, 1><<<grid, 256, 0, st>>>((const uint32_t*)vol, D, d_keys, (long long)nk,
                                                                      d_counts, d_cursor, d_rec[0], lin_span);
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

// out[i] = lut[vol[i]] (labels >= n_lut -> fill); one 16-byte vector of input voxels per thread iteration
template <typename TI, typename TO>
__global__ void map_labels_kernel(const TI* __restrict__ vol, TO* __restrict__ out, const TO* __restrict__ lut,
                                  unsigned long long n_lut, TO fill, size_t n) {
    constexpr int V = 16 / sizeof(TI);
    const size_t nvec = n / V;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        TI in[V];
        *reinterpret_cast<uint4*>(in) = *reinterpret_cast<const uint4*>(vol + i * V);
        TO res[V];
#pragma unroll
        for (int k = 0; k < V; ++k) res[k] = in[k] < n_lut ? lut[in[k]] : fill;
#pragma unroll
        for (int k = 0; k < V; ++k) out[i * V + k] = res[k];
    }
    for (size_t i = nvec * V + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = vol[i] < n_lut ? lut[vol[i]] : fill;
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
    if (elem == 2) {
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
    if (elem == 2)
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

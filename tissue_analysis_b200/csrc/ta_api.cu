// C ABI of libtissue_b200.so (see include/tissue_b200.h).  Host-side context, buffers, launches.
#include "../../include/tissue_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "ta_common.cuh"
#include "ta_kernels.cuh"
#include "ta_scan.cuh"
#include "ta_scan_mask.cuh"
#include "ta_prepass.cuh"
#include "ta_second_pass.cuh"

struct ta_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;

    const void* vol = nullptr;   // device pointer in use
    void* vol_owned = nullptr;   // our copy when bound from host memory
    size_t vol_owned_bytes = 0;
    int elem = 0;
    long long nf = 0, nm = 0, ns = 0, own_lo = 0, own_hi = 0, slow_offset = 0;

    LabelTable lt{};
    size_t lt_alloc_rows = 0;
    PairTable pt{};
    size_t pt_alloc_cap = 0;
    uint32_t* status = nullptr;       // [8]: status[4] (pair overflow, label range, evictions, spare) + counters[4]
    unsigned int* counters = nullptr; // = status + 4: brick counter, compact count, max label, spare
    uint32_t host_flags[8] = {};      // status + counters as read back at the one synchronisation of build_records
    bool timing_pending = false;

    u64* sort_keys[2] = {nullptr, nullptr};
    uint32_t* sort_vals[2] = {nullptr, nullptr};
    size_t sort_alloc = 0;
    void* cub_temp = nullptr;
    size_t cub_temp_bytes = 0;
    uint32_t* records = nullptr;
    size_t records_alloc = 0;
    uint64_t nrecords = 0;
    bool have_tables = false;
    uint32_t* fetch_scratch = nullptr;   // device scratch of the fetch functions (interleaved boxes, split pair columns)
    size_t fetch_scratch_have = 0;
    bool pending = false;             // deferred pass / merge: records, their count and the status flags are still on the device
    uint64_t deferred_cap = 0;        // rows of the deferred record buffer (without the header row)
    double* d_evals = nullptr;
    double* d_evecs = nullptr;
    size_t eig_alloc_rows = 0;
    u64* phase_cycles = nullptr;
    uint32_t* brick_core = nullptr;      // pre-pass (ta_prepass.cuh): one-label cores, then the scan's work list
    unsigned int* work_list = nullptr;
    size_t prepass_alloc = 0;            // bricks both have room for
    u64* diag_host = nullptr;         // host-mapped [8], survives a kernel trap
    u64* diag_dev = nullptr;

    cudaEvent_t ev[6] = {};
    cudaStream_t copy_stream = nullptr;          // H2D chunks of ta_run_pass_host
    std::vector<cudaEvent_t> chunk_ev;
    float scan_ms = 0, pass_ms = 0, h2d_ms = 0;
    uint64_t launches = 0;
};

static thread_local std::string g_err;

#define TA_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char buf_[512];                                                                   \
            snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call,          \
                     cudaGetErrorString(e_));                                                 \
            if (ctx) ctx->err = buf_; else g_err = buf_;                                      \
            return TA_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

static int fail(ta_ctx* ctx, int code, const char* msg) {
    if (ctx) ctx->err = msg; else g_err = msg;
    return code;
}

static size_t next_pow2(size_t v) { size_t p = 1; while (p < v) p <<= 1; return p; }

template <typename P> static int ensure(ta_ctx* ctx, P** ptr, size_t* have, size_t want_elems) {
    if (*have >= want_elems && *ptr) return TA_OK;
    if (*ptr) TA_CUDA(cudaFree(*ptr));
    *ptr = nullptr; *have = 0;
    TA_CUDA(cudaMalloc((void**)ptr, want_elems * sizeof(P)));
    *have = want_elems;
    return TA_OK;
}

// The scan kernel is the bit-mask kernel (mk::mask_kernel, ta_scan_mask.cuh); the product library carries the kernels it
// launches and nothing else.  A -DTA_WITH_BRICK_KERNEL build also carries round 1's worklist kernel (scan_kernel, ta_scan.cuh:
// the A / B baseline, same tables) and launches it under TA_SCAN_KERNEL=brick; a -DTA_WITH_PHASE_TIMING build the phase clocks.
typedef void (*scan_kernel_fn)(ScanParams, LabelTable, PairTable, const CUtensorMap);
static bool use_mask_kernel() {
#ifdef TA_WITH_BRICK_KERNEL
    static int v = -1;
    if (v < 0) { const char* e = getenv("TA_SCAN_KERNEL"); v = (e && !strcmp(e, "brick")) ? 0 : 1; }
    return v == 1;
#else
    return true;
#endif
}
// full: all three accumulations (TA_PASS_ALL, what the product runs) with the flags folded at compile time; the other
// instantiation reads them from the parameters
static scan_kernel_fn mask_kernel_variant(int elem, bool full) {
    if (full) return elem == 2 ? ta::mk::mask_kernel<uint16_t, 7> : ta::mk::mask_kernel<uint32_t, 7>;
    return elem == 2 ? ta::mk::mask_kernel<uint16_t, -1> : ta::mk::mask_kernel<uint32_t, -1>;
}
#ifdef TA_WITH_BRICK_KERNEL
static scan_kernel_fn scan_kernel_variant(int elem, bool timing) {
#ifdef TA_WITH_PHASE_TIMING
    if (timing) return elem == 2 ? ta::scan_kernel<uint16_t, true> : ta::scan_kernel<uint32_t, true>;
#endif
    (void)timing;
    return elem == 2 ? ta::scan_kernel<uint16_t, false> : ta::scan_kernel<uint32_t, false>;
}
#endif

extern "C" {

const char* ta_version(void) { return "tissue_b200 0.1 (sm_100a)"; }

const char* ta_last_error(ta_ctx* ctx) {
    if (!ctx) return g_err.c_str();
    if (ctx->diag_host && ctx->diag_host[0] && ctx->err.find("tile copy") == std::string::npos) {
        const u64* d = ctx->diag_host;
        char buf[384];
        snprintf(buf, sizeof buf,
                 " [scan kernel: a TMA tile copy did not complete: %llu waits timed out; first: CTA %llu thread %llu, "
                 "iteration %llu brick %llu, parity %llu next brick %llu, mbarrier word 0x%016llx, box origin (%d, %d, %d)]",
                 d[0], d[1] >> 32, d[1] & 0xFFFFFFFFull, d[2] >> 32, d[2] & 0xFFFFFFFFull, d[3] >> 32,
                 d[3] & 0xFFFFFFFFull, d[4], (int)(d[5] >> 32), (int)(int16_t)((d[5] >> 16) & 0xFFFF),
                 (int)(int16_t)(d[5] & 0xFFFF));
        ctx->err += buf;
    }
    return ctx->err.c_str();
}

int ta_ctx_create(ta_ctx** out, int device) {
    ta_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, TA_ERR_BAD_ARG, "ta_ctx_create: out is null");
    int ndev = 0;
    TA_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return fail(nullptr, TA_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    if (device >= 0) TA_CUDA(cudaSetDevice(device));
    else TA_CUDA(cudaGetDevice(&device));
    cudaDeviceProp prop;
    TA_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(nullptr, TA_ERR_CUDA, "device is not sm_100 class (Blackwell) hardware");
    ctx = new ta_ctx();
    struct CtxGuard { ta_ctx* c; ~CtxGuard() { if (c) ta_ctx_destroy(c); } } guard{ctx};    // an early return below frees it
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    TA_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    for (auto& e : ctx->ev) TA_CUDA(cudaEventCreate(&e));
    TA_CUDA(cudaMalloc((void**)&ctx->status, 8 * sizeof(uint32_t)));
    ctx->counters = ctx->status + 4;
#ifdef TA_WITH_BRICK_KERNEL
    for (int e = 0; e < 2; ++e)
        for (int tm = 0; tm < 2; ++tm)
            TA_CUDA(cudaFuncSetAttribute((const void*)scan_kernel_variant(e ? 4 : 2, tm != 0),
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(e ? ta::scan_smem_bytes<uint32_t>() : ta::scan_smem_bytes<uint16_t>())));
#endif
    for (int full = 0; full < 2; ++full) {
        TA_CUDA(cudaFuncSetAttribute((const void*)mask_kernel_variant(2, full != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)ta::mk::smem_bytes<uint16_t>()));
        TA_CUDA(cudaFuncSetAttribute((const void*)mask_kernel_variant(4, full != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)ta::mk::smem_bytes<uint32_t>()));
    }
    guard.c = nullptr;
    *out = ctx;
    return TA_OK;
}

int ta_ctx_destroy(ta_ctx* ctx) {
    if (!ctx) return TA_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->vol_owned);
    cudaFree(ctx->lt.count); cudaFree(ctx->lt.bmin);   // two blocks: sums (count|s1|s2) and boxes (bmin|bmax)
    cudaFree(ctx->pt.keys); cudaFree(ctx->pt.vals);
    cudaFree(ctx->status);
    for (int i = 0; i < 2; ++i) { cudaFree(ctx->sort_keys[i]); cudaFree(ctx->sort_vals[i]); }
    cudaFree(ctx->cub_temp); cudaFree(ctx->records);
    cudaFree(ctx->d_evals); cudaFree(ctx->d_evecs); cudaFree(ctx->phase_cycles); cudaFree(ctx->fetch_scratch);
    cudaFree(ctx->brick_core); cudaFree(ctx->work_list);
    for (auto& e : ctx->ev) cudaEventDestroy(e);
    for (auto& e : ctx->chunk_ev) cudaEventDestroy(e);
    if (ctx->diag_host) cudaFreeHost(ctx->diag_host);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return TA_OK;
}

int ta_set_stream(ta_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return TA_OK;
}

static int check_volume_args(ta_ctx* ctx, const void* data, int elem_bytes, int64_t n_fast, int64_t n_mid,
                             int64_t n_slow) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!data) return fail(ctx, TA_ERR_BAD_ARG, "ta_bind_volume: data is null");
    if (elem_bytes != 2 && elem_bytes != 4) return fail(ctx, TA_ERR_BAD_ARG, "ta_bind_volume: elem_bytes must be 2 or 4");
    if (n_fast <= 0 || n_mid <= 0 || n_slow <= 0 || n_fast > 0x7FFFFFF0LL || n_mid > 0x7FFFFFF0LL ||
        n_slow > 0x7FFFFFF0LL)
        return fail(ctx, TA_ERR_BAD_ARG, "ta_bind_volume: bad shape");
    return TA_OK;
}

static int ensure_owned_volume(ta_ctx* ctx, size_t bytes) {
    if (ctx->vol_owned_bytes < bytes) {
        if (ctx->vol_owned) TA_CUDA(cudaFree(ctx->vol_owned));
        ctx->vol_owned = nullptr; ctx->vol_owned_bytes = 0;
        TA_CUDA(cudaMalloc(&ctx->vol_owned, bytes));
        ctx->vol_owned_bytes = bytes;
    }
    return TA_OK;
}

int ta_bind_volume(ta_ctx* ctx, const void* data, int is_device, int elem_bytes, int64_t n_fast, int64_t n_mid,
                   int64_t n_slow) {
    int rc0 = check_volume_args(ctx, data, elem_bytes, n_fast, n_mid, n_slow);
    if (rc0) return rc0;
    TA_CUDA(cudaSetDevice(ctx->device));
    size_t bytes = (size_t)n_fast * n_mid * n_slow * elem_bytes;
    if (is_device) {
        ctx->vol = data;
    } else {
        rc0 = ensure_owned_volume(ctx, bytes);
        if (rc0) return rc0;
        TA_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
        TA_CUDA(cudaMemcpyAsync(ctx->vol_owned, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
        TA_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
        TA_CUDA(cudaStreamSynchronize(ctx->stream));
        TA_CUDA(cudaEventElapsedTime(&ctx->h2d_ms, ctx->ev[4], ctx->ev[5]));
        ctx->vol = ctx->vol_owned;
    }
    ctx->elem = elem_bytes;
    ctx->nf = n_fast; ctx->nm = n_mid; ctx->ns = n_slow;
    ctx->own_lo = 0; ctx->own_hi = n_slow; ctx->slow_offset = 0;
    ctx->have_tables = false;
    return TA_OK;
}

int ta_set_slab(ta_ctx* ctx, int64_t own_lo, int64_t own_hi, int64_t slow_offset) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->vol) return fail(ctx, TA_ERR_NO_VOLUME, "ta_set_slab: no volume bound");
    if (own_lo < 0 || own_hi > ctx->ns || own_lo > own_hi || slow_offset < 0)
        return fail(ctx, TA_ERR_BAD_ARG, "ta_set_slab: bad plane range");
    ctx->own_lo = own_lo; ctx->own_hi = own_hi; ctx->slow_offset = slow_offset;
    ctx->have_tables = false;
    return TA_OK;
}

static int ensure_sort_buffers(ta_ctx* ctx, size_t cap) {
    if (ctx->sort_alloc < cap) {
        for (int i = 0; i < 2; ++i) {
            if (ctx->sort_keys[i]) TA_CUDA(cudaFree(ctx->sort_keys[i]));
            if (ctx->sort_vals[i]) TA_CUDA(cudaFree(ctx->sort_vals[i]));
            ctx->sort_keys[i] = nullptr; ctx->sort_vals[i] = nullptr;
        }
        ctx->sort_alloc = 0;
        for (int i = 0; i < 2; ++i) {
            TA_CUDA(cudaMalloc((void**)&ctx->sort_keys[i], cap * sizeof(u64)));
            TA_CUDA(cudaMalloc((void**)&ctx->sort_vals[i], cap * sizeof(uint32_t)));
        }
        ctx->sort_alloc = cap;
    }
    return TA_OK;
}

// compaction + sort + gather of the pair hash into ctx->records
static int build_records(ta_ctx* ctx, bool sorted = true) {
    cudaStream_t st = ctx->stream;
    size_t cap = (size_t)ctx->pt.cap_mask + 1;
    int rcs = ensure_sort_buffers(ctx, cap);
    if (rcs) return rcs;
    TA_CUDA(cudaMemsetAsync(&ctx->counters[1], 0, sizeof(unsigned int), st));
    int blocks = (int)std::min<size_t>((cap + 255) / 256, (size_t)ctx->num_sms * 8);
    ta::compact_pairs_kernel<<<blocks, 256, 0, st>>>(ctx->pt, ctx->sort_keys[0], ctx->sort_vals[0], &ctx->counters[1]);
    ctx->launches++;
    // the only host synchronisation of a pass: status flags and the record count in one read-back
    TA_CUDA(cudaMemcpyAsync(ctx->host_flags, ctx->status, sizeof ctx->host_flags, cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaStreamSynchronize(st));
    const unsigned int n = ctx->host_flags[4 + 1];
    ctx->nrecords = n;
    if (n == 0) return TA_OK;
    if (!sorted) {
        int rc0 = ensure(ctx, &ctx->records, &ctx->records_alloc, (size_t)n * ta::REC_WORDS);
        if (rc0) return rc0;
        ta::gather_records_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx->pt, ctx->sort_keys[0], ctx->sort_vals[0], n,
                                                                  ctx->records);
        ctx->launches++;
        TA_CUDA(cudaGetLastError());
        return TA_OK;
    }
    size_t need = 0;
    const int end_bit = (ctx->elem == 2) ? 48 : 64;   // uint16 labels: key bits 16..31 and 48..63 are zero
    TA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, ctx->sort_keys[0], ctx->sort_keys[1], ctx->sort_vals[0],
                                            ctx->sort_vals[1], (int)n, 0, end_bit, st));
    if (need > ctx->cub_temp_bytes) {
        if (ctx->cub_temp) TA_CUDA(cudaFree(ctx->cub_temp));
        ctx->cub_temp = nullptr; ctx->cub_temp_bytes = 0;
        TA_CUDA(cudaMalloc(&ctx->cub_temp, need));
        ctx->cub_temp_bytes = need;
    }
    TA_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_temp, need, ctx->sort_keys[0], ctx->sort_keys[1],
                                            ctx->sort_vals[0], ctx->sort_vals[1], (int)n, 0, end_bit, st));
    int rc = ensure(ctx, &ctx->records, &ctx->records_alloc, (size_t)n * ta::REC_WORDS);
    if (rc) return rc;
    ta::gather_records_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx->pt, ctx->sort_keys[1], ctx->sort_vals[1], n,
                                                              ctx->records);
    ctx->launches++;
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

// Deferred pass: the records go to ctx->records on the device -- header row {count}, then the rows, hash order -- without
// the host ever learning the count.  No synchronisation.
static int pack_records_deferred(ta_ctx* ctx, uint64_t cap_rows) {
    cudaStream_t st = ctx->stream;
    const size_t cap = (size_t)ctx->pt.cap_mask + 1;
    int rc = ensure_sort_buffers(ctx, cap);
    if (rc) return rc;
    rc = ensure(ctx, &ctx->records, &ctx->records_alloc, (size_t)(cap_rows + 1) * ta::REC_WORDS);
    if (rc) return rc;
    TA_CUDA(cudaMemsetAsync(&ctx->counters[1], 0, sizeof(unsigned int), st));
    const int blocks = (int)std::min<size_t>((cap + 255) / 256, (size_t)ctx->num_sms * 8);
    ta::compact_pairs_kernel<<<blocks, 256, 0, st>>>(ctx->pt, ctx->sort_keys[0], ctx->sort_vals[0], &ctx->counters[1]);
    ta::gather_records_deferred_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(ctx->pt, ctx->sort_keys[0], ctx->sort_vals[0], &ctx->counters[1],
                                                                        (unsigned int)cap_rows, ctx->records);
    ctx->launches += 2;
    TA_CUDA(cudaGetLastError());
    ctx->deferred_cap = cap_rows;
    return TA_OK;
}

// What a deferred pass / merge left for later: the one host synchronisation (status flags, record count) and the sorted
// records.  Called by everything that hands results to the host.
static int resolve_pending(ta_ctx* ctx) {
    if (!ctx->pending) return TA_OK;
    ctx->pending = false;
    int rc = build_records(ctx);
    if (rc) { ctx->have_tables = false; return rc; }
    const uint32_t* status = ctx->host_flags;
    if (status[0] || status[3]) {
        ctx->have_tables = false;
        return fail(ctx, TA_ERR_PAIR_OVERFLOW, status[0] ? "pair table overflow (deferred pass): retry with a larger pair_capacity_hint"
                                                         : "more pair records than the deferred record buffer holds: retry with more rows");
    }
    if (status[1]) { ctx->have_tables = false; return fail(ctx, TA_ERR_LABEL_RANGE, "a label exceeds the label table (max_label_hint too small)"); }
    if (ctx->diag_host && ctx->diag_host[0]) {
        ctx->have_tables = false;
        int rcd = fail(ctx, TA_ERR_CUDA, "scan kernel: a CTA gave up waiting for its brick");
        ta_last_error(ctx);
        ctx->diag_host[0] = 0;
        return rcd;
    }
    return TA_OK;
}

static int ensure_pair_table(ta_ctx* ctx, size_t cap) {
    if (ctx->pt_alloc_cap < cap) {
        if (ctx->pt.keys) TA_CUDA(cudaFree(ctx->pt.keys));
        if (ctx->pt.vals) TA_CUDA(cudaFree(ctx->pt.vals));
        ctx->pt.keys = nullptr; ctx->pt.vals = nullptr; ctx->pt_alloc_cap = 0;
        TA_CUDA(cudaMalloc((void**)&ctx->pt.keys, cap * sizeof(u64)));
        TA_CUDA(cudaMalloc((void**)&ctx->pt.vals, cap * TA_PAIR_STRIDE * sizeof(uint32_t)));
        ctx->pt_alloc_cap = cap;
    }
    ctx->pt.cap_mask = (uint32_t)(cap - 1);
    ctx->pt.status = ctx->status;
    TA_CUDA(cudaMemsetAsync(ctx->pt.keys, 0xFF, cap * sizeof(u64), ctx->stream));
    TA_CUDA(cudaMemsetAsync(ctx->pt.vals, 0, cap * TA_PAIR_STRIDE * sizeof(uint32_t), ctx->stream));
    return TA_OK;
}

// Tensor map of the bound buffer for the scan kernel's TMA staging: dims (fast, mid, slow), box = one tile (brick +
// halo; 18 segments x 18 rows x 10 planes), no swizzle, out-of-bounds elements read as zero (the kernel re-clamps
// them).  false: TMA not usable (encoder missing or it rejected the shape) -> the kernel stages with cp.async.
typedef CUresult (*ta_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_tile_map(ta_ctx* ctx, CUtensorMap* map, bool mask_geometry) {
    static ta_encode_tiled_fn encode = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            encode = (ta_encode_tiled_fn)fn;
        else
            cudaGetLastError();
    }
    if (!encode) return false;
    const int seg = 16 / ctx->elem;
    const cuuint64_t dims[3] = {(cuuint64_t)ctx->nf, (cuuint64_t)ctx->nm, (cuuint64_t)ctx->ns};
    const cuuint64_t strides[2] = {(cuuint64_t)ctx->nf * ctx->elem, (cuuint64_t)ctx->nf * ctx->nm * ctx->elem};
    cuuint32_t box[3] = {(cuuint32_t)(ta::ROWV * seg), (cuuint32_t)(ta::BM + 2), (cuuint32_t)(ta::BS + 2)};
    if (mask_geometry) { box[0] = (cuuint32_t)(ta::mk::RW + 2 * seg); box[1] = (cuuint32_t)ta::mk::TM; box[2] = (cuuint32_t)ta::mk::TP; }
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, ctx->elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 3,
                        const_cast<void*>(ctx->vol), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// One launch of the scan kernel over owned planes [own_lo, own_hi) of the bound buffer; tables accumulate.
static int launch_scan(ta_ctx* ctx, ScanParams P, const CUtensorMap& tmap, long long own_lo, long long own_hi) {
    cudaStream_t st = ctx->stream;
    P.own_lo = own_lo; P.own_hi = own_hi;
    const bool mask = use_mask_kernel();
    P.nbs = (int)((own_hi - own_lo + ta::BS - 1) / ta::BS);         // BS == mk::ZB
    const size_t total = (size_t)P.nbf * P.nbm * P.nbs;
    if (total > 0xFFFFFFF0ull) return fail(ctx, TA_ERR_BAD_ARG, "volume too large for one pass");
    if (total == 0) return TA_OK;
    TA_CUDA(cudaMemsetAsync(&ctx->counters[0], 0, sizeof(unsigned int), st));
    if (mask) {
        // The pre-pass takes the interior of one-label regions (background) out of the queue.  TA_PREPASS=0 / 1 forces it off /
        // on; by default it runs from 16384 bricks (126 Mvoxel) up: measured -4 % on C3 and C4, whose domes are half background,
        // +4 % (13 us: two launches and a first look at every brick) on C2, 9216 bricks of tissue only.
        const char* pe = getenv("TA_PREPASS");
        const bool prepass = P.vec_ok && (pe ? atoi(pe) != 0 : total >= 16384);
        if (prepass) {
            if (ctx->prepass_alloc < total) {
                cudaFree(ctx->brick_core); cudaFree(ctx->work_list);
                ctx->brick_core = nullptr; ctx->work_list = nullptr; ctx->prepass_alloc = 0;
                TA_CUDA(cudaMalloc(&ctx->brick_core, total * sizeof(uint32_t)));
                TA_CUDA(cudaMalloc(&ctx->work_list, total * sizeof(unsigned int)));
                ctx->prepass_alloc = total;
            }
            TA_CUDA(cudaMemsetAsync(&ctx->counters[3], 0, sizeof(unsigned int), st));
            ta::pp::PrepassParams Q{};
            Q.vol = P.vol; Q.nf = (int)P.nf; Q.nm = (int)P.nm; Q.ns = (int)P.ns; Q.own_lo = (int)own_lo; Q.own_hi = (int)own_hi;
            Q.slow_offset = P.slow_offset; Q.nbf = P.nbf; Q.nbm = P.nbm; Q.nbs = P.nbs;
            Q.core = ctx->brick_core; Q.work_list = ctx->work_list; Q.work_count = &ctx->counters[3]; Q.do_mom = P.flags & 1u;
            // a warp per brick, many more blocks than are resident: measured on C3 at 8 / 16 / 32 blocks per SM 0.29 / 0.24 / 0.21 ms
            // (128: the scan another 0.6 % shorter on C3, 1 % on C4)
            const int cgrid = (int)std::min<size_t>((total + 7) / 8, (size_t)ctx->num_sms * 128);
            if (ctx->elem == 2) ta::pp::classify_cores_kernel<uint16_t><<<cgrid, 256, 0, st>>>(Q);
            else ta::pp::classify_cores_kernel<uint32_t><<<cgrid, 256, 0, st>>>(Q);
            ta::pp::decide_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Q, ctx->lt, ctx->pt.status);
            ctx->launches += 2;
            TA_CUDA(cudaGetLastError());
            P.work_list = ctx->work_list; P.work_count = &ctx->counters[3];
        }
        const int per_sm = ctx->elem == 2 ? 3 : 2;
        const int grid = (int)std::min<size_t>(total, (size_t)ctx->num_sms * per_sm);
        const size_t smem = ctx->elem == 2 ? ta::mk::smem_bytes<uint16_t>() : ta::mk::smem_bytes<uint32_t>();
        mask_kernel_variant(ctx->elem, (P.flags & 7u) == 7u)<<<grid, ta::mk::NTHREADS, smem, st>>>(P, ctx->lt, ctx->pt, tmap);
        ctx->launches++;
        TA_CUDA(cudaGetLastError());
        return TA_OK;
    }
#ifdef TA_WITH_BRICK_KERNEL
    int grid = (int)std::min<size_t>(total, (size_t)ctx->num_sms * 3);
    const size_t smem = ctx->elem == 2 ? ta::scan_smem_bytes<uint16_t>() : ta::scan_smem_bytes<uint32_t>();
    scan_kernel_variant(ctx->elem, P.phase_cycles != nullptr)<<<grid, ta::NTHREADS, smem, st>>>(P, ctx->lt, ctx->pt, tmap);
    ctx->launches++;
    TA_CUDA(cudaGetLastError());
    return TA_OK;
#else
    return fail(ctx, TA_ERR_BAD_ARG, "round 1's scan kernel is not in this build (-DTA_WITH_BRICK_KERNEL)");
#endif
}

// host_src != nullptr: the bound (context-owned) buffer is filled from host_src in chunks of `chunk_planes` planes on
// a copy stream while the scan kernel already works on the planes that have arrived.
struct PassRanges {           // ta_run_pass_ranges: plane ranges of the owned region, each behind an optional event
    int n = 0;
    const int64_t* lo_hi = nullptr;
    void* const* wait_events = nullptr;
};

static int run_pass_impl(ta_ctx* ctx, uint32_t flags, uint32_t max_label_hint, uint64_t pair_capacity_hint,
                         const void* host_src, long long chunk_planes, const PassRanges* ranges = nullptr) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->vol) return fail(ctx, TA_ERR_NO_VOLUME, "ta_run_pass: no volume bound");
    if ((flags & (TA_PASS_ALL | 0x300u)) == 0) return fail(ctx, TA_ERR_BAD_ARG, "ta_run_pass: empty flags");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->have_tables = false;
    const size_t nvox = (size_t)ctx->nf * ctx->nm * ctx->ns;
    TA_CUDA(cudaEventRecord(ctx->ev[0], st));

    // ---- label table rows ---------------------------------------------------------------------------------------
    size_t nrows;
    if (ctx->elem == 2) {
        nrows = 65536;
    } else if (max_label_hint) {
        nrows = (size_t)max_label_hint + 1;
    } else {
        TA_CUDA(cudaMemsetAsync(&ctx->counters[2], 0, sizeof(unsigned int), st));
        ta::max_label_kernel<<<ctx->num_sms * 8, 256, 0, st>>>((const uint32_t*)ctx->vol, nvox, &ctx->counters[2]);
        ctx->launches++;
        unsigned int mx = 0;
        TA_CUDA(cudaMemcpyAsync(&mx, &ctx->counters[2], sizeof mx, cudaMemcpyDeviceToHost, st));
        TA_CUDA(cudaStreamSynchronize(st));
        if (mx == 0xFFFFFFFFu) return fail(ctx, TA_ERR_LABEL_RANGE, "label 0xFFFFFFFF is reserved");
        nrows = (size_t)mx + 1;
    }
    if (ctx->lt_alloc_rows < nrows) {
        cudaFree(ctx->lt.count); cudaFree(ctx->lt.bmin);
        ctx->lt = LabelTable{}; ctx->lt_alloc_rows = 0;
        // one block for the sums (count | s1 | s2 = 10 u64 per row) and one for the boxes (bmin | bmax): the sharded
        // driver all_reduces each block with a single collective
        TA_CUDA(cudaMalloc((void**)&ctx->lt.count, nrows * 10 * sizeof(u64)));
        ctx->lt.s1 = ctx->lt.count + nrows;
        ctx->lt.s2 = ctx->lt.s1 + nrows * 3;
        TA_CUDA(cudaMalloc((void**)&ctx->lt.bmin, nrows * 6 * sizeof(int)));
        ctx->lt.bmax = ctx->lt.bmin + nrows * 3;
        ctx->lt_alloc_rows = nrows;
    }
    ctx->lt.nrows = (uint32_t)nrows;
    ta::init_label_table_kernel<<<(unsigned)((nrows * 6 + 255) / 256), 256, 0, st>>>(ctx->lt);
    ctx->launches++;

    // ---- pair table -----------------------------------------------------------------------------------------------
    size_t cap = pair_capacity_hint ? next_pow2((size_t)pair_capacity_hint * 3)
                                    : next_pow2(std::min<size_t>(std::max<size_t>(nvox / 1024, 1u << 14), 1u << 24));
    cap = std::max<size_t>(cap, 1024);
    if (cap > (1ull << 31)) return fail(ctx, TA_ERR_BAD_ARG, "pair capacity too large");
    int rc = ensure_pair_table(ctx, cap);
    if (rc) return rc;
    TA_CUDA(cudaMemsetAsync(ctx->status, 0, 4 * sizeof(uint32_t), st));
    TA_CUDA(cudaMemsetAsync(ctx->counters, 0, 4 * sizeof(unsigned int), st));

    // ---- the scan ---------------------------------------------------------------------------------------------------
    ScanParams P{};
    P.vol = ctx->vol;
    P.nf = ctx->nf; P.nm = ctx->nm; P.ns = ctx->ns;
    P.own_lo = ctx->own_lo; P.own_hi = ctx->own_hi; P.slow_offset = ctx->slow_offset;
    const int seg = 16 / ctx->elem;
    const long long BF = (long long)ta::NFS * seg;
    P.nbf = (int)((ctx->nf + BF - 1) / BF);
    P.nbm = (int)((ctx->nm + ta::BM - 1) / ta::BM);
    P.nbs = (int)((ctx->own_hi - ctx->own_lo + ta::BS - 1) / ta::BS);
    if (use_mask_kernel()) {
        P.nbf = (int)((ctx->nf + ta::mk::RW - 1) / ta::mk::RW);
        P.nbm = (int)((ctx->nm + ta::mk::OM - 1) / ta::mk::OM);
    }
    P.flags = flags;
    P.vec_ok = ((ctx->nf % seg) == 0) && (((uintptr_t)ctx->vol & 15) == 0);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    P.use_tma = (P.vec_ok && !getenv("TA_NO_TMA") && make_tile_map(ctx, &tmap, use_mask_kernel())) ? 1 : 0;
    P.brick_counter = &ctx->counters[0];
    if (!ctx->diag_host) {
        if (cudaHostAlloc((void**)&ctx->diag_host, 8 * sizeof(u64), cudaHostAllocMapped) == cudaSuccess) {
            memset(ctx->diag_host, 0, 8 * sizeof(u64));
            if (cudaHostGetDevicePointer((void**)&ctx->diag_dev, ctx->diag_host, 0) != cudaSuccess) ctx->diag_dev = nullptr;
        } else {
            cudaGetLastError();
            ctx->diag_host = nullptr;
        }
    }
    P.diag = ctx->diag_dev;
    P.phase_cycles = nullptr;
#ifdef TA_WITH_PHASE_TIMING
    const bool phase_timing = getenv("TA_PHASE_TIMING") != nullptr;
#else
    const bool phase_timing = false;         // the product library has no kernel with phase clocks
#endif
    if (phase_timing) {
        if (!ctx->phase_cycles) TA_CUDA(cudaMalloc((void**)&ctx->phase_cycles, 16 * sizeof(u64)));
        TA_CUDA(cudaMemsetAsync(ctx->phase_cycles, 0, 16 * sizeof(u64), st));
        P.phase_cycles = ctx->phase_cycles;
    }
    const size_t total = (size_t)P.nbf * P.nbm * P.nbs;
    if (!(ranges && ranges->n == 1 && ranges->wait_events && ranges->wait_events[0])) TA_CUDA(cudaEventRecord(ctx->ev[1], st));
    if (ranges) {
        for (int k = 0; k < ranges->n; ++k) {
            if (ranges->wait_events && ranges->wait_events[k]) {
                TA_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)ranges->wait_events[k], 0));
                // one range behind one event (the halo exchange): the wait is not scan time
                if (ranges->n == 1) TA_CUDA(cudaEventRecord(ctx->ev[1], st));
            }
            rc = launch_scan(ctx, P, tmap, ranges->lo_hi[2 * k], ranges->lo_hi[2 * k + 1]);
            if (rc) return rc;
        }
    } else if (!host_src) {
        rc = launch_scan(ctx, P, tmap, ctx->own_lo, ctx->own_hi);
        if (rc) return rc;
    } else {
        // chunk k = planes [c_k, c_k+1) is copied on the copy stream; the scan of planes [c_k - 1, c_k+1 - 1) (the last
        // chunk runs to the end) follows on the pass stream as soon as the chunk has landed: it needs plane c_k+1 - 1 as
        // its upper halo and nothing beyond.
        const long long ns = ctx->ns, cp = std::max<long long>(chunk_planes, 2);
        const size_t plane_bytes = (size_t)ctx->nf * ctx->nm * ctx->elem;
        const size_t nchunks = (size_t)((ns + cp - 1) / cp);
        if (!ctx->copy_stream) TA_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        while (ctx->chunk_ev.size() < nchunks + 1) {
            cudaEvent_t e;
            TA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->chunk_ev.push_back(e);
        }
        // the copy stream starts after everything queued on the pass stream so far (table resets included)
        TA_CUDA(cudaEventRecord(ctx->chunk_ev[nchunks], st));
        TA_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[nchunks], 0));
        TA_CUDA(cudaEventRecord(ctx->ev[4], ctx->copy_stream));
        for (size_t k = 0; k < nchunks; ++k) {
            const long long c0 = (long long)k * cp, c1 = std::min<long long>(c0 + cp, ns);
            TA_CUDA(cudaMemcpyAsync((char*)ctx->vol_owned + (size_t)c0 * plane_bytes,
                                    (const char*)host_src + (size_t)c0 * plane_bytes, (size_t)(c1 - c0) * plane_bytes,
                                    cudaMemcpyHostToDevice, ctx->copy_stream));
            TA_CUDA(cudaEventRecord(ctx->chunk_ev[k], ctx->copy_stream));
            TA_CUDA(cudaStreamWaitEvent(st, ctx->chunk_ev[k], 0));
            const long long lo = std::max<long long>(c0 - 1, ctx->own_lo);
            const long long hi = std::min<long long>((c1 == ns) ? ns : c1 - 1, ctx->own_hi);
            if (lo < hi) {
                rc = launch_scan(ctx, P, tmap, lo, hi);
                if (rc) return rc;
            }
        }
        TA_CUDA(cudaEventRecord(ctx->ev[5], ctx->copy_stream));
    }
    (void)total;
    TA_CUDA(cudaEventRecord(ctx->ev[2], st));
    if (phase_timing) {
        u64 cyc[16];
        TA_CUDA(cudaMemcpyAsync(cyc, ctx->phase_cycles, sizeof cyc, cudaMemcpyDeviceToHost, st));
        TA_CUDA(cudaStreamSynchronize(st));
        double tot = 0;
        for (int k = 0; k < 8; ++k) tot += (double)cyc[k];
        const char* nm[8] = {"sched", "A stage", "B codes", "C1 march", "C2 flags", "D voxels", "D2 junctions", "F flush"};
        if (use_mask_kernel()) {
            const char* pn[6] = {"tile wait", "P1", "barrier 1", "P2", "barrier 2", "flush"};
            for (int w = 0; w < 2; ++w) {
                double t = 0;
                for (int k = 0; k < 6; ++k) t += (double)cyc[8 * w + k];
                fprintf(stderr, "[ta mask kernel, %s warp]", w ? "last" : "first");
                for (int k = 0; k < 6; ++k) fprintf(stderr, " %s %.1f%%", pn[k], t > 0 ? 100.0 * cyc[8 * w + k] / t : 0.0);
                if (w == 0) fprintf(stderr, " | bricks %llu, one-label %llu, cycles per brick %.0f\n", cyc[7], cyc[6], cyc[7] ? t / cyc[7] : 0.0);
                else fprintf(stderr, " | cycles per one-label brick %.0f, per other brick %.0f\n", cyc[6] ? (double)cyc[14] / cyc[6] : 0.0,
                             cyc[7] > cyc[6] ? (t - (double)cyc[14]) / (cyc[7] - cyc[6]) : 0.0);
            }
        }
        else {
            fprintf(stderr, "[ta phase cycles, thread 0 of each CTA]");
            for (int k = 0; k < 8; ++k) fprintf(stderr, " %s %.1f%%", nm[k], tot > 0 ? 100.0 * cyc[k] / tot : 0.0);
            fprintf(stderr, "\n");
        }
    }

    if (flags & TA_PASS_DEFERRED) {
        // no host synchronisation: records packed on the device, flags checked at the first fetch (resolve_pending)
        const uint64_t cap_rows = pair_capacity_hint ? pair_capacity_hint : std::min<uint64_t>((uint64_t)ctx->pt.cap_mask + 1, 1ull << 20);
        rc = pack_records_deferred(ctx, cap_rows);
        if (rc) return rc;
        TA_CUDA(cudaEventRecord(ctx->ev[3], st));
        ctx->timing_pending = true;
        ctx->pending = true;
        ctx->have_tables = true;
        return TA_OK;
    }
    ctx->pending = false;
    rc = build_records(ctx, !(flags & TA_PASS_UNSORTED));
    if (rc) return rc;
    TA_CUDA(cudaEventRecord(ctx->ev[3], st));
    const uint32_t* status = ctx->host_flags;      // read back inside build_records, after the scan kernel
    ctx->timing_pending = true;                     // events are resolved lazily by ta_last_timing
    if (ctx->diag_host && ctx->diag_host[0]) {      // a CTA gave up waiting for a tile copy and left (ta_scan.cuh, phase A)
        int rcd = fail(ctx, TA_ERR_CUDA, "scan kernel: a CTA gave up waiting for its brick");
        ta_last_error(ctx);                         // appends the diagnostic while it is still there
        ctx->diag_host[0] = 0;
        return rcd;
    }
    if (phase_timing) fprintf(stderr, "[ta] moment-slot evictions: %u (%.3f per segment column)\n", status[2],
                              (double)status[2] / ((double)total * ta::NTHREADS));
    if (status[0]) return fail(ctx, TA_ERR_PAIR_OVERFLOW, "pair table overflow: retry with a larger pair_capacity_hint");
    if (status[1]) return fail(ctx, TA_ERR_LABEL_RANGE, "a label exceeds max_label_hint");
    ctx->have_tables = true;
    return TA_OK;
}

int ta_run_pass(ta_ctx* ctx, uint32_t flags, uint32_t max_label_hint, uint64_t pair_capacity_hint) {
    return run_pass_impl(ctx, flags, max_label_hint, pair_capacity_hint, nullptr, 0);
}

int ta_run_pass_ranges(ta_ctx* ctx, uint32_t flags, uint32_t max_label_hint, uint64_t pair_capacity_hint, int n_ranges,
                       const int64_t* lo_hi, void* const* wait_events) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (n_ranges <= 0 || !lo_hi) return fail(ctx, TA_ERR_BAD_ARG, "ta_run_pass_ranges: no ranges");
    // the ranges must tile the owned planes exactly once (any order)
    std::vector<std::pair<int64_t, int64_t>> r;
    for (int k = 0; k < n_ranges; ++k)
        if (lo_hi[2 * k] < lo_hi[2 * k + 1]) r.push_back({lo_hi[2 * k], lo_hi[2 * k + 1]});
    std::sort(r.begin(), r.end());
    int64_t at = ctx->own_lo;
    for (auto& x : r) {
        if (x.first != at) return fail(ctx, TA_ERR_BAD_ARG, "ta_run_pass_ranges: ranges must tile the owned planes");
        at = x.second;
    }
    if (at != ctx->own_hi) return fail(ctx, TA_ERR_BAD_ARG, "ta_run_pass_ranges: ranges must tile the owned planes");
    PassRanges pr;
    pr.n = n_ranges; pr.lo_hi = lo_hi; pr.wait_events = wait_events;
    return run_pass_impl(ctx, flags, max_label_hint, pair_capacity_hint, nullptr, 0, &pr);
}

int ta_run_pass_host(ta_ctx* ctx, const void* host_data, int elem_bytes, int64_t n_fast, int64_t n_mid, int64_t n_slow,
                     const int64_t* slab, uint32_t flags, uint32_t max_label_hint, uint64_t pair_capacity_hint,
                     int64_t chunk_planes) {
    int rc = check_volume_args(ctx, host_data, elem_bytes, n_fast, n_mid, n_slow);
    if (rc) return rc;
    if (slab && (slab[0] < 0 || slab[1] > n_slow || slab[0] > slab[1] || slab[2] < 0))
        return fail(ctx, TA_ERR_BAD_ARG, "ta_run_pass_host: bad plane range");
    if (elem_bytes == 4 && max_label_hint == 0) {
        // the height of the label table needs the largest label first: copy, then the ordinary pass
        rc = ta_bind_volume(ctx, host_data, 0, elem_bytes, n_fast, n_mid, n_slow);
        if (!rc && slab) rc = ta_set_slab(ctx, slab[0], slab[1], slab[2]);
        return rc ? rc : run_pass_impl(ctx, flags, max_label_hint, pair_capacity_hint, nullptr, 0);
    }
    TA_CUDA(cudaSetDevice(ctx->device));
    rc = ensure_owned_volume(ctx, (size_t)n_fast * n_mid * n_slow * elem_bytes);
    if (rc) return rc;
    ctx->vol = ctx->vol_owned;
    ctx->elem = elem_bytes;
    ctx->nf = n_fast; ctx->nm = n_mid; ctx->ns = n_slow;
    ctx->own_lo = 0; ctx->own_hi = n_slow; ctx->slow_offset = 0;
    if (slab) { ctx->own_lo = slab[0]; ctx->own_hi = slab[1]; ctx->slow_offset = slab[2]; }
    if (chunk_planes <= 0) {
        // about 64 MiB per chunk: long enough to run PCIe at full rate, short enough that the last chunk's scan is a
        // small tail; a multiple of the brick height
        const size_t plane_bytes = (size_t)n_fast * n_mid * elem_bytes;
        chunk_planes = (int64_t)std::max<size_t>(ta::BS, ((64u << 20) / std::max<size_t>(plane_bytes, 1)) / ta::BS * ta::BS);
    }
    rc = run_pass_impl(ctx, flags, max_label_hint, pair_capacity_hint, host_data, chunk_planes);
    if (rc == TA_OK || rc == TA_ERR_PAIR_OVERFLOW || rc == TA_ERR_LABEL_RANGE) {
        // the whole volume is resident now (second passes, retries with ta_run_pass)
        cudaStreamSynchronize(ctx->copy_stream);
        cudaEventElapsedTime(&ctx->h2d_ms, ctx->ev[4], ctx->ev[5]);
    }
    return rc;
}

int ta_label_table_size(ta_ctx* ctx, uint64_t* n) {
    if (!ctx || !n) return fail(ctx, TA_ERR_BAD_ARG, "null argument");
    if (!ctx->have_tables) return fail(ctx, TA_ERR_NO_TABLES, "no tables: call ta_run_pass first");
    { int rcp = resolve_pending(ctx); if (rcp) return rcp; }
    *n = ctx->lt.nrows;
    return TA_OK;
}

int ta_fetch_label_table(ta_ctx* ctx, uint64_t* count, uint64_t* s1, uint64_t* s2, int32_t* bbox) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->have_tables) return fail(ctx, TA_ERR_NO_TABLES, "no tables: call ta_run_pass first");
    { int rcp = resolve_pending(ctx); if (rcp) return rcp; }
    TA_CUDA(cudaSetDevice(ctx->device));
    size_t n = ctx->lt.nrows;
    cudaStream_t st = ctx->stream;
    // the layouts the caller wants are made on the device; the host side is four copies and ONE synchronisation
    if (bbox) {
        int rc = ensure(ctx, &ctx->fetch_scratch, &ctx->fetch_scratch_have, n * 6);
        if (rc) return rc;
        ta::interleave_boxes_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(ctx->lt.bmin, ctx->lt.bmax, n, (int*)ctx->fetch_scratch);
        ctx->launches++;
        TA_CUDA(cudaGetLastError());
    }
    if (count) TA_CUDA(cudaMemcpyAsync(count, ctx->lt.count, n * sizeof(u64), cudaMemcpyDeviceToHost, st));
    if (s1) TA_CUDA(cudaMemcpyAsync(s1, ctx->lt.s1, n * 3 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    if (s2) TA_CUDA(cudaMemcpyAsync(s2, ctx->lt.s2, n * 6 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    if (bbox) TA_CUDA(cudaMemcpyAsync(bbox, ctx->fetch_scratch, n * 6 * sizeof(int), cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaStreamSynchronize(st));
    return TA_OK;
}

int ta_pair_table_size(ta_ctx* ctx, uint64_t* n) {
    if (!ctx || !n) return fail(ctx, TA_ERR_BAD_ARG, "null argument");
    if (!ctx->have_tables) return fail(ctx, TA_ERR_NO_TABLES, "no tables: call ta_run_pass first");
    { int rcp = resolve_pending(ctx); if (rcp) return rcp; }
    *n = ctx->nrecords;
    return TA_OK;
}

int ta_fetch_pair_table(ta_ctx* ctx, uint32_t* lo, uint32_t* hi, uint32_t* faces, uint32_t* wall18) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->have_tables) return fail(ctx, TA_ERR_NO_TABLES, "no tables: call ta_run_pass first");
    { int rcp = resolve_pending(ctx); if (rcp) return rcp; }
    TA_CUDA(cudaSetDevice(ctx->device));
    size_t n = ctx->nrecords;
    if (n == 0) return TA_OK;
    cudaStream_t st = ctx->stream;
    int rc = ensure(ctx, &ctx->fetch_scratch, &ctx->fetch_scratch_have, n * ta::REC_WORDS);
    if (rc) return rc;
    ta::split_records_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(ctx->records, n, ctx->fetch_scratch);
    ctx->launches++;
    TA_CUDA(cudaGetLastError());
    const uint32_t* d = ctx->fetch_scratch;
    if (lo) TA_CUDA(cudaMemcpyAsync(lo, d, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (hi) TA_CUDA(cudaMemcpyAsync(hi, d + n, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (faces) TA_CUDA(cudaMemcpyAsync(faces, d + 2 * n, n * 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (wall18) TA_CUDA(cudaMemcpyAsync(wall18, d + 8 * n, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaStreamSynchronize(st));
    return TA_OK;
}

int ta_label_table_device(ta_ctx* ctx, void** count, void** s1, void** s2, void** bmin, void** bmax, uint64_t* n) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->have_tables) return fail(ctx, TA_ERR_NO_TABLES, "no tables: call ta_run_pass first");
    if (count) *count = ctx->lt.count;
    if (s1) *s1 = ctx->lt.s1;
    if (s2) *s2 = ctx->lt.s2;
    if (bmin) *bmin = ctx->lt.bmin;
    if (bmax) *bmax = ctx->lt.bmax;
    if (n) *n = ctx->lt.nrows;
    return TA_OK;
}

int ta_pair_records_device(ta_ctx* ctx, void** records, uint64_t* n) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->have_tables) return fail(ctx, TA_ERR_NO_TABLES, "no tables: call ta_run_pass first");
    { int rcp = resolve_pending(ctx); if (rcp) return rcp; }
    if (records) *records = ctx->records;
    if (n) *n = ctx->nrecords;
    return TA_OK;
}

int ta_merge_pair_records(ta_ctx* ctx, const void* device_records, uint64_t n) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (n && !device_records) return fail(ctx, TA_ERR_BAD_ARG, "ta_merge_pair_records: null records");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    size_t cap = std::max<size_t>(next_pow2((size_t)n * 2), 1024);
    int rc = ensure_pair_table(ctx, cap);
    if (rc) return rc;
    TA_CUDA(cudaMemsetAsync(ctx->status, 0, 4 * sizeof(uint32_t), st));
    if (n) {
        ta::merge_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->pt, (const uint32_t*)device_records, n);
        ctx->launches++;
    }
    rc = build_records(ctx);
    if (rc) return rc;
    const uint32_t* status = ctx->host_flags;
    if (status[0]) return fail(ctx, TA_ERR_PAIR_OVERFLOW, "pair table overflow in merge");
    ctx->have_tables = true;
    return TA_OK;
}

int ta_pair_records_deferred(ta_ctx* ctx, void** records, uint64_t* cap_rows) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->have_tables || !ctx->pending) return fail(ctx, TA_ERR_NO_TABLES, "no deferred pass: call ta_run_pass with TA_PASS_DEFERRED first");
    if (records) *records = ctx->records;
    if (cap_rows) *cap_rows = ctx->deferred_cap;
    return TA_OK;
}

int ta_merge_pair_records_deferred(ta_ctx* ctx, const void* gathered, uint64_t cap_rows, int world) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!gathered || world <= 0 || cap_rows == 0) return fail(ctx, TA_ERR_BAD_ARG, "ta_merge_pair_records_deferred: bad arguments");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t cap = std::max<size_t>(next_pow2((size_t)cap_rows * (size_t)world * 2), 1024);
    // the status flags of the local pass stay (checked at the first fetch); only the table is emptied
    int rc = ensure_pair_table(ctx, cap);
    if (rc) return rc;
    const size_t total = (size_t)cap_rows * (size_t)world;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->num_sms * 16);
    ta::merge_records_deferred_kernel<<<blocks, 256, 0, st>>>(ctx->pt, (const uint32_t*)gathered, (unsigned int)cap_rows, world);
    ctx->launches++;
    TA_CUDA(cudaGetLastError());
    ctx->pending = true;
    ctx->have_tables = true;
    return TA_OK;
}

int ta_inertia_from_moments(ta_ctx* ctx, const uint32_t* labels, uint64_t n, double* evals, double* evecs) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->have_tables) return fail(ctx, TA_ERR_NO_TABLES, "no tables: call ta_run_pass first");
    if (n == 0) return TA_OK;
    if (!evals || !evecs) return fail(ctx, TA_ERR_BAD_ARG, "null output");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // results go to the context's eigen buffers (kept between calls), the label list to a scoped temporary
    if (ctx->eig_alloc_rows < n) {
        cudaFree(ctx->d_evals); cudaFree(ctx->d_evecs);
        ctx->d_evals = ctx->d_evecs = nullptr; ctx->eig_alloc_rows = 0;
        TA_CUDA(cudaMalloc((void**)&ctx->d_evals, n * 3 * sizeof(double)));
        TA_CUDA(cudaMalloc((void**)&ctx->d_evecs, n * 9 * sizeof(double)));
        ctx->eig_alloc_rows = n;
    }
    double *dw = ctx->d_evals, *dv = ctx->d_evecs;
    TaDevBuf dl;
    if (labels) {
        TA_CUDA(cudaMalloc(&dl.p, n * sizeof(uint32_t)));
        TA_CUDA(cudaMemcpyAsync(dl.p, labels, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
    ta::inertia_from_moments_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ctx->lt, dl.as<uint32_t>(), n, dw, dv);
    ctx->launches++;
    TA_CUDA(cudaMemcpyAsync(evals, dw, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaMemcpyAsync(evecs, dv, n * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaStreamSynchronize(st));
    return TA_OK;
}

int ta_inertia_table(ta_ctx* ctx, double* evals, double* evecs) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->have_tables) return fail(ctx, TA_ERR_NO_TABLES, "no tables: call ta_run_pass first");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    size_t n = ctx->lt.nrows;
    if (ctx->eig_alloc_rows < n) {
        cudaFree(ctx->d_evals); cudaFree(ctx->d_evecs);
        ctx->d_evals = ctx->d_evecs = nullptr; ctx->eig_alloc_rows = 0;
        TA_CUDA(cudaMalloc((void**)&ctx->d_evals, n * 3 * sizeof(double)));
        TA_CUDA(cudaMalloc((void**)&ctx->d_evecs, n * 9 * sizeof(double)));
        ctx->eig_alloc_rows = n;
    }
    ta::inertia_from_moments_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ctx->lt, nullptr, n, ctx->d_evals,
                                                                               ctx->d_evecs);
    ctx->launches++;
    TA_CUDA(cudaGetLastError());
    if (evals) TA_CUDA(cudaMemcpyAsync(evals, ctx->d_evals, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (evecs) TA_CUDA(cudaMemcpyAsync(evecs, ctx->d_evecs, n * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (evals || evecs) TA_CUDA(cudaStreamSynchronize(st));
    return TA_OK;
}

int ta_inertia_eig(ta_ctx* ctx, const double* cov, uint64_t n, double* evals, double* evecs) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (n == 0) return TA_OK;
    if (!cov || !evals || !evecs) return fail(ctx, TA_ERR_BAD_ARG, "null argument");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    TaDevBuf dc, dw, dv;
    TA_CUDA(cudaMalloc(&dc.p, n * 6 * sizeof(double)));
    TA_CUDA(cudaMalloc(&dw.p, n * 3 * sizeof(double)));
    TA_CUDA(cudaMalloc(&dv.p, n * 9 * sizeof(double)));
    TA_CUDA(cudaMemcpyAsync(dc.p, cov, n * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
    ta::inertia_eig_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(dc.as<double>(), n, dw.as<double>(), dv.as<double>());
    ctx->launches++;
    TA_CUDA(cudaMemcpyAsync(evals, dw.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaMemcpyAsync(evecs, dv.p, n * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaStreamSynchronize(st));
    return TA_OK;
}

int ta_wall_voxel_coords(ta_ctx* ctx, const uint32_t* lo, const uint32_t* hi, uint64_t npairs, uint64_t* counts,
                         int64_t* xyz) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->vol) return fail(ctx, TA_ERR_NO_VOLUME, "no volume bound");
    if (npairs == 0) return TA_OK;
    if (!lo || !hi || !counts) return fail(ctx, TA_ERR_BAD_ARG, "null argument");
    TA_CUDA(cudaSetDevice(ctx->device));
    return ta::wall_voxel_coords_impl(ctx->vol, ctx->elem, ctx->nf, ctx->nm, ctx->ns, ctx->own_lo, ctx->own_hi,
                                      ctx->slow_offset, lo, hi, npairs, counts, xyz, ctx->stream, ctx->num_sms,
                                      &ctx->launches, &ctx->err);
}

int ta_voxel_first_layer(ta_ctx* ctx, uint32_t background, int keep_background, void* out_host) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->vol) return fail(ctx, TA_ERR_NO_VOLUME, "no volume bound");
    if (!out_host) return fail(ctx, TA_ERR_BAD_ARG, "null output");
    TA_CUDA(cudaSetDevice(ctx->device));
    return ta::voxel_first_layer_impl(ctx->vol, ctx->elem, ctx->nf, ctx->nm, ctx->ns, background, keep_background,
                                      out_host, ctx->stream, ctx->num_sms, &ctx->launches, &ctx->err);
}

static int stencil_entry(ta_ctx* ctx, int kind, void* out_host) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->vol) return fail(ctx, TA_ERR_NO_VOLUME, "no volume bound");
    if (!out_host) return fail(ctx, TA_ERR_BAD_ARG, "null output");
    TA_CUDA(cudaSetDevice(ctx->device));
    return ta::stencil_image_impl(ctx->vol, ctx->elem, ctx->nf, ctx->nm, ctx->ns, kind, out_host, ctx->stream,
                                  ctx->num_sms, &ctx->launches, &ctx->err);
}

int ta_hollow_out_cells(ta_ctx* ctx, int mask_only, void* out_host) { return stencil_entry(ctx, mask_only ? 2 : 0, out_host); }

int ta_cell_shell18(ta_ctx* ctx, void* out_host) { return stencil_entry(ctx, 1, out_host); }

int ta_map_labels(ta_ctx* ctx, const void* lut_host, int lut_elem_bytes, uint64_t n_lut, uint32_t fill, void* out_host,
                  int in_place) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!ctx->vol) return fail(ctx, TA_ERR_NO_VOLUME, "no volume bound");
    if (!lut_host || n_lut == 0) return fail(ctx, TA_ERR_BAD_ARG, "ta_map_labels: empty lookup table");
    if (lut_elem_bytes != 2 && lut_elem_bytes != 4) return fail(ctx, TA_ERR_BAD_ARG, "ta_map_labels: lut_elem_bytes must be 2 or 4");
    if (in_place && (lut_elem_bytes != ctx->elem || ctx->vol != ctx->vol_owned))
        return fail(ctx, TA_ERR_BAD_ARG, "ta_map_labels: in-place needs the context-owned volume and a table of its dtype");
    if (((uintptr_t)ctx->vol & 15) != 0) return fail(ctx, TA_ERR_BAD_ARG, "ta_map_labels: volume is not 16-byte aligned");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)ctx->nf * ctx->nm * ctx->ns;
    TaDevBuf lut_buf, out_buf;
    TA_CUDA(cudaMalloc(&lut_buf.p, n_lut * lut_elem_bytes));
    TA_CUDA(cudaMalloc(&out_buf.p, n * lut_elem_bytes));
    void *d_lut = lut_buf.p, *d_out = out_buf.p;
    TA_CUDA(cudaMemcpyAsync(d_lut, lut_host, n_lut * lut_elem_bytes, cudaMemcpyHostToDevice, st));
    // the table in shared memory when it fits next to nothing else (one CTA of 1024 threads per SM), else through L1
    {
        const size_t lut_bytes = (size_t)n_lut * lut_elem_bytes;
        const bool smem = lut_bytes <= 160 * 1024;
        const int threads = smem ? 1024 : 256, grid = smem ? ctx->num_sms : ctx->num_sms * 16;
        const size_t dyn = smem ? lut_bytes : 0;
#define TA_MAP_LAUNCH(TI, TO)                                                                                                  \
        do {                                                                                                                       \
            if (smem) {                                                                                                            \
                TA_CUDA(cudaFuncSetAttribute((const void*)ta::map_labels_kernel<TI, TO, true>,                                   \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));                              \
                ta::map_labels_kernel<TI, TO, true><<<grid, threads, dyn, st>>>((const TI*)ctx->vol, (TO*)d_out, (const TO*)d_lut, \
                                                                                n_lut, (TO)fill, n);                               \
            } else {                                                                                                               \
                ta::map_labels_kernel<TI, TO, false><<<grid, threads, 0, st>>>((const TI*)ctx->vol, (TO*)d_out, (const TO*)d_lut,  \
                                                                               n_lut, (TO)fill, n);                                \
            }                                                                                                                      \
        } while (0)
        if (ctx->elem == 2 && lut_elem_bytes == 2) TA_MAP_LAUNCH(uint16_t, uint16_t);
        else if (ctx->elem == 2) TA_MAP_LAUNCH(uint16_t, uint32_t);
        else if (lut_elem_bytes == 2) TA_MAP_LAUNCH(uint32_t, uint16_t);
        else TA_MAP_LAUNCH(uint32_t, uint32_t);
#undef TA_MAP_LAUNCH
    }
    ctx->launches++;
    TA_CUDA(cudaGetLastError());
    if (in_place) {
        TA_CUDA(cudaMemcpyAsync(ctx->vol_owned, d_out, n * lut_elem_bytes, cudaMemcpyDeviceToDevice, st));
        ctx->have_tables = false;
    }
    if (out_host) TA_CUDA(cudaMemcpyAsync(out_host, d_out, n * lut_elem_bytes, cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaStreamSynchronize(st));
    return TA_OK;
}

int ta_last_timing(ta_ctx* ctx, float* scan_ms, float* pass_ms, float* h2d_ms) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (ctx->timing_pending) {
        TA_CUDA(cudaEventSynchronize(ctx->ev[3]));
        TA_CUDA(cudaEventElapsedTime(&ctx->scan_ms, ctx->ev[1], ctx->ev[2]));
        TA_CUDA(cudaEventElapsedTime(&ctx->pass_ms, ctx->ev[0], ctx->ev[3]));
        ctx->timing_pending = false;
    }
    if (scan_ms) *scan_ms = ctx->scan_ms;
    if (pass_ms) *pass_ms = ctx->pass_ms;
    if (h2d_ms) *h2d_ms = ctx->h2d_ms;
    return TA_OK;
}

int ta_launch_count(ta_ctx* ctx, uint64_t* n) {
    if (!ctx || !n) return fail(ctx, TA_ERR_BAD_ARG, "null argument");
    *n = ctx->launches;
    return TA_OK;
}

int ta_synth_voronoi(ta_ctx* ctx, void* device_out, int elem_bytes, int64_t n_fast, int64_t n_mid, int64_t n_slow,
                     int64_t slow_offset, int64_t global_slow, const int32_t* seeds_host, uint32_t ncell,
                     const int32_t* weight, int dome) {
    if (!ctx) return fail(nullptr, TA_ERR_BAD_ARG, "null context");
    if (!device_out || !seeds_host || !weight || ncell == 0) return fail(ctx, TA_ERR_BAD_ARG, "null argument");
    if (elem_bytes != 2 && elem_bytes != 4) return fail(ctx, TA_ERR_BAD_ARG, "elem_bytes must be 2 or 4");
    if (elem_bytes == 2 && ncell + 1 > 65535) return fail(ctx, TA_ERR_BAD_ARG, "too many cells for uint16");
    TA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ta::SynthParams P{};
    P.nf = n_fast; P.nm = n_mid; P.ns = n_slow; P.slow_offset = slow_offset; P.global_slow = global_slow;
    P.wf = weight[0]; P.wm = weight[1]; P.ws = weight[2];
    P.dome = dome; P.ncell = ncell;
    // weighted extents (fixed point 1/16 voxel) and a bin edge giving ~2 seeds per bin
    double ext[3] = {16.0 * n_fast * P.wf, 16.0 * n_mid * P.wm, 16.0 * global_slow * P.ws};
    double edge = cbrt(ext[0] * ext[1] * ext[2] * 2.0 / (double)ncell);
    P.bin = (int)std::max(16.0, std::min(edge, 1.0e8));
    P.gx = (int)(ext[0] / P.bin) + 1; P.gy = (int)(ext[1] / P.bin) + 1; P.gz = (int)(ext[2] / P.bin) + 1;
    {
        auto coef = [](long long dim) {
            long long D = (long long)(0.94 * (double)dim + 0.5);
            if (D < 1) D = 1;
            return (1LL << 40) / (D * D);
        };
        P.kf = coef(n_fast); P.km = coef(n_mid); P.ks = coef(global_slow);
    }
    size_t nb = (size_t)P.gx * P.gy * P.gz;
    std::vector<int> start(nb + 1, 0), order(ncell);
    std::vector<size_t> binof(ncell);
    for (uint32_t i = 0; i < ncell; ++i) {
        long long x = (long long)seeds_host[i * 3 + 0] * P.wf / P.bin, y = (long long)seeds_host[i * 3 + 1] * P.wm / P.bin,
                  z = (long long)seeds_host[i * 3 + 2] * P.ws / P.bin;
        x = std::min<long long>(std::max<long long>(x, 0), P.gx - 1);
        y = std::min<long long>(std::max<long long>(y, 0), P.gy - 1);
        z = std::min<long long>(std::max<long long>(z, 0), P.gz - 1);
        binof[i] = ((size_t)z * P.gy + y) * P.gx + x;
        start[binof[i] + 1]++;
    }
    for (size_t b = 0; b < nb; ++b) start[b + 1] += start[b];
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (uint32_t i = 0; i < ncell; ++i) order[fill[binof[i]]++] = (int)i;
    TaDevBuf b_start, b_order, b_seeds;
    TA_CUDA(cudaMalloc(&b_start.p, (nb + 1) * sizeof(int)));
    TA_CUDA(cudaMalloc(&b_order.p, ncell * sizeof(int)));
    TA_CUDA(cudaMalloc(&b_seeds.p, (size_t)ncell * 3 * sizeof(int)));
    int *d_start = b_start.as<int>(), *d_order = b_order.as<int>(), *d_seeds = b_seeds.as<int>();
    TA_CUDA(cudaMemcpyAsync(d_start, start.data(), (nb + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    TA_CUDA(cudaMemcpyAsync(d_order, order.data(), ncell * sizeof(int), cudaMemcpyHostToDevice, st));
    TA_CUDA(cudaMemcpyAsync(d_seeds, seeds_host, (size_t)ncell * 3 * sizeof(int), cudaMemcpyHostToDevice, st));
    int grid = ctx->num_sms * 16;
    if (elem_bytes == 2)
        ta::synth_voronoi_kernel<uint16_t><<<grid, 256, 0, st>>>((uint16_t*)device_out, P, d_start, d_order, d_seeds);
    else
        ta::synth_voronoi_kernel<uint32_t><<<grid, 256, 0, st>>>((uint32_t*)device_out, P, d_start, d_order, d_seeds);
    ctx->launches++;
    TA_CUDA(cudaGetLastError());
    TA_CUDA(cudaStreamSynchronize(st));
    return TA_OK;
}

}  // extern "C"

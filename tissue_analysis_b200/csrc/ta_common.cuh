// Shared device-side definitions for libtissue_b200 (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda.h>            // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>

#define TA_EMPTY64 0xFFFFFFFFFFFFFFFFull
#define TA_EMPTY32 0xFFFFFFFFu
#define TA_PAIR_STRIDE 8   // u32 per pair row in the hash: faces[6], wall18, pad

typedef unsigned long long u64;

// Host side: a temporary device buffer that is freed when its scope ends, on the error paths too (TA_CUDA returns early).
struct TaDevBuf {
    void* p = nullptr;
    TaDevBuf() {}
    TaDevBuf(const TaDevBuf&) = delete;
    TaDevBuf& operator=(const TaDevBuf&) = delete;
    ~TaDevBuf() { if (p) cudaFree(p); }
    template <typename U> U* as() const { return static_cast<U*>(p); }
};

// inline PTX goes through this macro so that the CPU emulation of the kernels (tests/host/emu) can compile them with g++
#ifndef TA_PTX
#define TA_PTX(...) asm volatile(__VA_ARGS__)
#endif

// Dense per-label table (row index = label value).  Sums are exact u64 integers so the result does not
// depend on accumulation order (bit-reproducible, mergeable across ranks by plain addition).
struct LabelTable {
    u64* count;   // [nrows]
    u64* s1;      // [nrows*3]  sum f, m, s (global indices)
    u64* s2;      // [nrows*6]  sum ff, fm, fs, mm, ms, ss
    int* bmin;    // [nrows*3]
    int* bmax;    // [nrows*3]
    uint32_t nrows;
};

// Global open-addressing table keyed by (lo << 32 | hi).
struct PairTable {
    u64* keys;        // [cap], TA_EMPTY64 when free
    uint32_t* vals;   // [cap * TA_PAIR_STRIDE]
    uint32_t cap_mask;
    uint32_t* status; // [0] pair overflow, [1] label out of range
};

struct ScanParams {
    const void* vol;
    long long nf, nm, ns;        // bound buffer dims (memory axes)
    long long own_lo, own_hi;    // owned planes [own_lo, own_hi) of the buffer
    long long slow_offset;       // global index of buffer plane 0
    int nbf, nbm, nbs;           // bricks per axis over the owned region
    uint32_t flags;
    int vec_ok;                  // rows are 16-byte aligned: 128-bit loads allowed
    int use_tma;                 // the tile (brick + halo) is one TMA box copy (needs vec_ok and a tensor map)
    unsigned int* brick_counter; // dynamic brick scheduler
    const unsigned int* work_list;   // optional: the bricks to scan, in this order (mask kernel; ta_prepass.cuh) ...
    const unsigned int* work_count;  // ... and how many
    u64* phase_cycles;           // optional [16]: per-phase clock64 totals of thread 0 of every CTA (profiling aid)
    u64* diag;                   // [8] host-mapped: what a CTA was doing when it gave up waiting for a tile copy
};

__device__ __forceinline__ uint32_t ta_hash64(u64 k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (uint32_t)k;
}

__host__ __device__ __forceinline__ u64 ta_pair_key(uint32_t a, uint32_t b) {
    return a < b ? ((u64)a << 32) | b : ((u64)b << 32) | a;
}

// Find-or-claim the slot of `key`; -1 (and the overflow flag) when the probe sequence is exhausted.
__device__ __forceinline__ int ta_pair_slot(const PairTable& t, u64 key) {
    uint32_t slot = ta_hash64(key) & t.cap_mask;
    const uint32_t limit = t.cap_mask < 8191u ? t.cap_mask : 8191u;
    for (uint32_t probe = 0; probe <= limit; ++probe) {
        u64 k = *((volatile u64*)&t.keys[slot]);
        if (k == key) return (int)slot;
        if (k == TA_EMPTY64) {
            u64 old = atomicCAS(&t.keys[slot], TA_EMPTY64, key);
            if (old == TA_EMPTY64 || old == key) return (int)slot;
        }
        slot = (slot + 1) & t.cap_mask;
    }
    atomicExch(&t.status[0], 1u);
    return -1;
}

__device__ __forceinline__ void ta_pair_add(const PairTable& t, u64 key, int field, uint32_t n) {
    int slot = ta_pair_slot(t, key);
    if (slot >= 0) atomicAdd(&t.vals[(size_t)slot * TA_PAIR_STRIDE + field], n);
}

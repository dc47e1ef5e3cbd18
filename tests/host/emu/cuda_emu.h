// A small CPU emulation of the CUDA execution model, just enough to run the scan kernels of csrc/ta_scan*.cuh unmodified
// under g++ (test infrastructure; nothing in the product includes this).
//
//   * one thread block at a time; every CUDA thread is a ucontext fiber with its own stack, scheduled round-robin;
//   * __syncthreads / __syncthreads_and: block-wide rendezvous; __ballot_sync, __shfl_sync, __reduce_*_sync, __syncwarp:
//     warp-wide rendezvous (full masks only, which is all the kernels use) -- a fiber that arrives yields until the last
//     participant has arrived and computed everybody's result;
//   * atomics are plain read-modify-writes (the fibers are cooperative, one OS thread);
//   * inline PTX is swallowed (the kernels spell it TA_PTX(...)); the TMA box copy and its mbarrier have plain-code
//     stand-ins under TA_EMU_TMA (ta_scan.cuh), so use_tma = 1 runs too; the cp.async path (vec_ok) does not;
//   * a scheduler round in which nothing progresses is reported as a deadlock (a missing participant of a collective).
//
// Include this INSTEAD of compiling with nvcc, before the kernel headers:  #include "emu/cuda_emu.h"
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <ucontext.h>

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <random>
#include <vector>

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

namespace emu {

struct Dim3 { unsigned x = 0, y = 0, z = 0; };
inline Dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

constexpr int MAX_THREADS = 1024, WARP = 32;
constexpr size_t STACK_BYTES = 256 * 1024;

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = true;
};
struct Rendezvous {                    // one per warp, plus one for the block
    int arrived = 0;
    unsigned gen = 0;
    unsigned long long val[MAX_THREADS];
    unsigned long long res[2][MAX_THREADS];
};

inline Fiber g_fibers[MAX_THREADS];
inline ucontext_t g_sched;
inline int g_cur = 0, g_nthreads = 0;
inline bool g_progress = false;
inline Rendezvous g_block;
inline Rendezvous g_warp[MAX_THREADS / WARP];
inline std::function<void()> g_kernel;

inline void yield() { swapcontext(&g_fibers[g_cur].ctx, &g_sched); }

// Operation counts of the emulated launches (cost indicators, not instruction counts): collectives are counted once per
// warp / block, atomics once per calling thread and by address space.
struct Stats {
    unsigned long long syncthreads = 0, ballot = 0, shfl = 0, redux = 0, syncwarp = 0, atom_shared = 0, atom_global = 0;
    void clear() { *this = Stats(); }
};
inline Stats g_stats;
inline const unsigned char* g_smem_lo = nullptr;
inline const unsigned char* g_smem_hi = nullptr;      // set by the harness: [lo, hi) is the block's shared memory
inline void count_atomic(const void* p) {
    const unsigned char* q = (const unsigned char*)p;
    if (q >= g_smem_lo && q < g_smem_hi) ++g_stats.atom_shared; else ++g_stats.atom_global;
}

// Every participant calls with its value; `compute(vals, results, n)` runs once, in the last arriver.
template <typename F>
inline unsigned long long rendezvous(Rendezvous& r, int slot, int n, unsigned long long v, F compute,
                                     unsigned long long* counter = nullptr) {
    const unsigned my_gen = r.gen;
    r.val[slot] = v;
    g_progress = true;
    if (++r.arrived == n) {
        if (counter) ++*counter;
        compute(r.val, r.res[my_gen & 1u], n);
        r.arrived = 0;
        ++r.gen;
    } else {
        while (r.gen == my_gen) yield();
    }
    return r.res[my_gen & 1u][slot];
}

inline int lane() { return g_cur % WARP; }
inline Rendezvous& my_warp() { return g_warp[g_cur / WARP]; }
inline int warp_size_here() { return std::min(WARP, g_nthreads - (g_cur / WARP) * WARP); }

inline void fiber_entry() {
    g_kernel();
    g_fibers[g_cur].done = true;
    g_progress = true;
    swapcontext(&g_fibers[g_cur].ctx, &g_sched);
}

// Runs `kernel` for every thread of one block.  false: deadlock.
inline bool run_block(unsigned block, unsigned grid, int nthreads, std::function<void()> kernel) {
    g_kernel = kernel;
    g_nthreads = nthreads;
    g_blockIdx.x = block; g_gridDim.x = grid; g_blockDim.x = (unsigned)nthreads;
    g_block = Rendezvous();
    for (auto& w : g_warp) { w.arrived = 0; w.gen = 0; }
    for (int t = 0; t < nthreads; ++t) {
        Fiber& f = g_fibers[t];
        if (!f.stack) f.stack = (char*)malloc(STACK_BYTES);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = &g_sched;
        f.done = false;
        makecontext(&f.ctx, fiber_entry, 0);
    }
    for (;;) {
        bool alive = false;
        g_progress = false;
        for (int t = 0; t < nthreads; ++t) {
            if (g_fibers[t].done) continue;
            alive = true;
            g_cur = t;
            g_threadIdx.x = (unsigned)t;
            swapcontext(&g_sched, &g_fibers[t].ctx);
        }
        if (!alive) return true;
        if (!g_progress) {
            fprintf(stderr, "emu: deadlock in block %u (a collective is missing a participant)\n", block);
            return false;
        }
    }
}

}  // namespace emu

// ---- the CUDA names the kernels use -----------------------------------------------------------------------------------
#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define volatile                                   /* single OS thread: plain accesses (define AFTER every std header) */
#define TA_SHARED static                           /* statically sized __shared__ variables: one block at a time, so a static */
#define TA_PTX(...) ((void)0)                      /* inline PTX: only on paths the emulation does not take */

inline void __syncthreads() {
    emu::rendezvous(emu::g_block, emu::g_cur, emu::g_nthreads, 0ull, [](unsigned long long*, unsigned long long*, int) {},
                    &emu::g_stats.syncthreads);
}
inline int __syncthreads_and(int pred) {
    return (int)emu::rendezvous(emu::g_block, emu::g_cur, emu::g_nthreads, pred ? 1ull : 0ull,
                                [](unsigned long long* v, unsigned long long* r, int n) {
                                    unsigned long long all = 1ull;
                                    for (int i = 0; i < n; ++i) all &= v[i];
                                    for (int i = 0; i < n; ++i) r[i] = all;
                                }, &emu::g_stats.syncthreads);
}
inline int __syncthreads_or(int pred) {
    return (int)emu::rendezvous(emu::g_block, emu::g_cur, emu::g_nthreads, pred ? 1ull : 0ull,
                                [](unsigned long long* v, unsigned long long* r, int n) {
                                    unsigned long long any = 0ull;
                                    for (int i = 0; i < n; ++i) any |= v[i];
                                    for (int i = 0; i < n; ++i) r[i] = any;
                                }, &emu::g_stats.syncthreads);
}
inline void emu_check_mask(unsigned mask) {
    if (mask != 0xffffffffu) { fprintf(stderr, "emu: only full-mask warp collectives are emulated\n"); abort(); }
}
inline void __syncwarp(unsigned mask = 0xffffffffu) {
    emu_check_mask(mask);
    emu::rendezvous(emu::my_warp(), emu::lane(), emu::warp_size_here(), 0ull, [](unsigned long long*, unsigned long long*, int) {},
                    &emu::g_stats.syncwarp);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu_check_mask(mask);
    return (unsigned)emu::rendezvous(emu::my_warp(), emu::lane(), emu::warp_size_here(), pred ? 1ull : 0ull,
                                     [](unsigned long long* v, unsigned long long* r, int n) {
                                         unsigned long long b = 0;
                                         for (int i = 0; i < n; ++i) b |= (v[i] & 1ull) << i;
                                         for (int i = 0; i < n; ++i) r[i] = b;
                                     }, &emu::g_stats.ballot);
}
template <typename V> inline V __shfl_sync(unsigned mask, V var, int src, int width = 32) {
    emu_check_mask(mask);
    (void)width;
    static_assert(sizeof(V) <= 8, "shuffle of at most 64 bits");
    unsigned long long bits = 0;
    memcpy(&bits, &var, sizeof(V));
    // every lane may name a different source: gather all values, then pick
    emu::Rendezvous& w = emu::my_warp();
    const int ln = emu::lane();
    emu::rendezvous(w, ln, emu::warp_size_here(), bits, [](unsigned long long* v, unsigned long long* r, int n) {
        for (int i = 0; i < n; ++i) r[i] = v[i];          // results = a snapshot of all values; picked below per lane
    }, &emu::g_stats.shfl);
    // the snapshot of generation g lives in res[g & 1]; read the source lane's entry of the generation just completed
    const unsigned done_gen = w.gen - 1u;
    unsigned long long out = w.res[done_gen & 1u][src & 31];
    V o;
    memcpy(&o, &out, sizeof(V));
    // nobody may start overwriting this snapshot before all lanes have read it: the NEXT collective uses the other buffer,
    // and the one after that needs all lanes to arrive first, i.e. to have left this function
    return o;
}
template <typename V> inline V __shfl_xor_sync(unsigned mask, V var, int lane_mask, int width = 32) {
    return __shfl_sync(mask, var, emu::lane() ^ lane_mask, width);
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0u; }
template <typename V> inline V __shfl_up_sync(unsigned mask, V var, unsigned delta, int width = 32) {
    const int ln = emu::lane();
    const V got = __shfl_sync(mask, var, ln >= (int)delta ? ln - (int)delta : ln, width);
    return got;
}
// redux with per-lane member masks (the match_any + redux pattern: every lane of the warp calls, each names its own group;
// a lane's result is the reduction over the lanes of ITS mask)
#define EMU_REDUX(name, init, op)                                                                                    \
    inline unsigned name(unsigned mask, unsigned v) {                                                                 \
        if (!((mask >> emu::lane()) & 1u)) { fprintf(stderr, "emu: redux: the caller is not in its member mask\n"); abort(); } \
        return (unsigned)emu::rendezvous(emu::my_warp(), emu::lane(), emu::warp_size_here(),                          \
                                         ((unsigned long long)mask << 32) | (unsigned long long)v,                    \
                                         [](unsigned long long* x, unsigned long long* r, int n) {                    \
                                             for (int me = 0; me < n; ++me) {                                         \
                                                 const unsigned mk = (unsigned)(x[me] >> 32);                         \
                                                 unsigned a = (init);                                                 \
                                                 for (int i = 0; i < n; ++i)                                          \
                                                     if ((mk >> i) & 1u) { const unsigned b = (unsigned)x[i]; a = (op); } \
                                                 r[me] = a;                                                           \
                                             }                                                                        \
                                         }, &emu::g_stats.redux);                                                     \
    }
EMU_REDUX(__reduce_add_sync, 0u, a + b)
EMU_REDUX(__reduce_min_sync, 0xFFFFFFFFu, (a < b ? a : b))
EMU_REDUX(__reduce_max_sync, 0u, (a > b ? a : b))
EMU_REDUX(__reduce_or_sync, 0u, a | b)
EMU_REDUX(__reduce_and_sync, 0xFFFFFFFFu, a & b)

template <typename V> inline unsigned __match_any_sync(unsigned mask, V key) {
    emu_check_mask(mask);
    static_assert(sizeof(V) <= 8, "match on at most 64 bits");
    unsigned long long bits = 0;
    memcpy(&bits, &key, sizeof(V));
    return (unsigned)emu::rendezvous(emu::my_warp(), emu::lane(), emu::warp_size_here(), bits,
                                     [](unsigned long long* x, unsigned long long* r, int n) {
                                         for (int me = 0; me < n; ++me) {
                                             unsigned long long m = 0;
                                             for (int i = 0; i < n; ++i) if (x[i] == x[me]) m |= 1ull << i;
                                             r[me] = m;
                                         }
                                     }, &emu::g_stats.ballot);
}
template <typename V> inline V atomicAdd(V* p, V v) { emu::count_atomic(p); V o = *p; *p = (V)(o + v); return o; }
inline unsigned atomicAdd(unsigned* p, int v) { emu::count_atomic(p); unsigned o = *p; *p = o + (unsigned)v; return o; }
template <typename V> inline V atomicMin(V* p, V v) { emu::count_atomic(p); V o = *p; if (v < o) *p = v; return o; }
template <typename V> inline V atomicMax(V* p, V v) { emu::count_atomic(p); V o = *p; if (v > o) *p = v; return o; }
template <typename V> inline V atomicExch(V* p, V v) { emu::count_atomic(p); V o = *p; *p = v; return o; }
template <typename V> inline V atomicCAS(V* p, V cmp, V v) { emu::count_atomic(p); V o = *p; if (o == cmp) *p = v; return o; }
template <typename V> inline V atomicOr(V* p, V v) { emu::count_atomic(p); V o = *p; *p = o | v; return o; }

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31u));
}
inline unsigned __vminu2(unsigned a, unsigned b) {
    const unsigned l = std::min(a & 0xFFFFu, b & 0xFFFFu), h = std::min(a >> 16, b >> 16);
    return l | (h << 16);
}
inline long long clock64() { return 0; }
inline void __trap() { fprintf(stderr, "emu: __trap() in thread %d of block %u\n", emu::g_cur, emu::g_blockIdx.x); abort(); }
inline void __threadfence_system() {}
inline void __threadfence() {}
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
using std::max;
using std::min;
inline int min(int a, unsigned b) { return a < (int)b ? a : (int)b; }
inline int min(unsigned a, int b) { return (int)a < b ? (int)a : b; }

// The single streaming pass over the label volume (sm_100a).
//
// Work unit: a brick of BF x BM x BS voxels (BF = 16 segments of 16 bytes) staged in shared memory with a
// one-voxel halo, clamped at the buffer edges (a clamped neighbour equals an in-bounds 6/18-neighbour or the
// voxel itself, which reproduces the reference's "the image border contributes nothing" rule).
//
// Phases per brick (irregular work is compacted into worklists first so whole warps stay busy):
//   A  stage brick + halo: one TMA box copy (cp.async.bulk.tensor.3d, completion on an mbarrier) issued by thread 0;
//      bricks on a face of the buffer re-clamp the zero-filled out-of-bounds elements in shared memory.  Rows that are
//      not 16-byte multiples (and TA_NO_TMA=1) take 16-byte cp.async copies / scalar loads instead.
//   B  per row-segment uniformity code: the label if the SEG+2 voxels (segment + f-halo) are equal.  A tile that is
//      one label altogether (background, inside of a large cell) is finished here with closed-form moments.
//   C1 march: thread (fseg, m) walks s.  Moments go to three bit-packed register slots per thread (a column rarely
//      sees more labels; 6 registers per slot: n|sum s|sum s^2, sum f|sum fs, sum f^2, f bounds, s bounds; the m
//      terms are closed forms).  Mixed segments are split into runs with SIMD compares.  At the end of the column
//      the slots are merged across the warp (a warp-uniform loop over the distinct labels with full-mask redux, 9
//      redux per label) and the group leaders update the per-brick shared label table in one SIMT pass.  Segments
//      whose 3x3 rows are not one label go to the SEGMENT worklist.
//   C2 per listed segment, SIMD on the packed lanes: OR of XORs of the segment with its 18 neighbour vectors
//      (f-shifted ones built with funnel shifts), min.u16x2 to turn non-zero lanes into bits -> exact "has a
//      different 18-neighbour" bit per voxel; flagged voxels go to the VOXEL worklist (tile element offsets).
//   D  per listed voxel: 18 neighbours fetched as aligned 32-bit pairs; d0 = a ^ OR(v ^ a) and the "only {a, d0}"
//      test min.u16x2(v ^ a, v ^ d0) == 0, two neighbours per instruction.  Packed 16-bit counters (wall18 and the
//      +f/+m/+s faces) are summed over a thread's chunk of consecutive list entries, merged across the warp, and the
//      group leaders update the per-brick shared pair table.  Junction voxels (>= 2 other labels) go to a third
//      worklist and are handled with a register dedup of up to 4 labels (exact rescan beyond that).
//   F  flush the per-brick label table (u32 brick-local sums -> shifted u64 global REDs) and pair table.
//
#pragma once
#include "ta_common.cuh"

namespace ta {

constexpr int NFS = 16;                 // 16-byte segments per brick row
constexpr int BM = 16;                  // brick rows (mid axis)
constexpr int BS = 8;                   // brick planes (slow axis)
constexpr int NTHREADS = NFS * BM;      // one thread per segment column
constexpr int LT_SLOTS = 32;            // per-brick label slots (power of two)
constexpr int LT_BITS = 5;
constexpr int LT_FIELDS = 16;           // n, sf, sm, ss, sff, sfm, sfs, smm, sms, sss, min f/m/s, max f/m/s
constexpr int PT_SLOTS = 128;           // per-brick pair slots (power of two)
constexpr int PT_BITS = 7;
constexpr int PT_WORDS = 4;             // packed 16-bit counters: [w18|f0] [f1|f2] [f3|f4] [f5|-]
constexpr int TILE_ROWS = (BS + 2) * (BM + 2);
constexpr int ROWV = NFS + 2;           // vectors per tile row
constexpr int PLANEV = (BM + 2) * ROWV; // vectors per tile plane
constexpr int TILE_SEGS = TILE_ROWS * ROWV;
constexpr int SEGLIST_CAP = NFS * BM * BS;
constexpr int VOXLIST_CAP = NTHREADS * 8;

template <typename T> struct Vox;
template <> struct Vox<uint16_t> {
    static constexpr int SEG = 8, LOG_SEG = 3;
    typedef unsigned short Code;                    // label 0xFFFF reads as "mixed": slower exact paths, same result
    static constexpr uint32_t MIXED = 0xFFFFu;
    typedef uint32_t PKey;                          // (lo << 16) | hi
    static constexpr PKey PEMPTY = 0xFFFFFFFFu;
    static __device__ __forceinline__ PKey key(uint32_t a, uint32_t b) { return a < b ? (a << 16) | b : (b << 16) | a; }
    static __device__ __forceinline__ u64 key64(PKey k) { return ((u64)(k >> 16) << 32) | (k & 0xFFFFu); }
    static __device__ __forceinline__ uint32_t hash(PKey k) { return (k * 0x9E3779B1u) >> (32 - PT_BITS); }
};
template <> struct Vox<uint32_t> {
    static constexpr int SEG = 4, LOG_SEG = 2;
    typedef uint32_t Code;
    static constexpr uint32_t MIXED = 0xFFFFFFFFu;
    typedef u64 PKey;
    static constexpr PKey PEMPTY = TA_EMPTY64;
    static __device__ __forceinline__ PKey key(uint32_t a, uint32_t b) { return ta_pair_key(a, b); }
    static __device__ __forceinline__ u64 key64(PKey k) { return k; }
    static __device__ __forceinline__ uint32_t hash(PKey k) {
        return (((uint32_t)(k >> 32) * 0x9E3779B1u) ^ ((uint32_t)k * 0x85EBCA77u)) >> (32 - PT_BITS);
    }
};

template <typename T> constexpr size_t scan_smem_bytes() {
    return (size_t)TILE_SEGS * 16 + (size_t)TILE_ROWS * NFS * sizeof(typename Vox<T>::Code) + LT_SLOTS * 4 +
           LT_SLOTS * LT_FIELDS * 4 + PT_SLOTS * sizeof(typename Vox<T>::PKey) + PT_SLOTS * PT_WORDS * 4 +
           SEGLIST_CAP * 2 + 2 * VOXLIST_CAP * 2 + 64 + 256 + 128;
}

__device__ __forceinline__ uint4 ld_stream_128(const void* p) {
    uint4 r;
    TA_PTX("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    TA_PTX("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    TA_PTX("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ---- TMA staging: one 3-D box copy (brick + halo) per tile, completion on an mbarrier ------------------------------
#ifdef TA_EMU_TMA
// CPU emulation only (tests/host/emu; never defined in a CUDA build): the box copy and its barrier as plain code, so that
// the kernels' TMA paths -- box coordinates, zero fill outside the tensor, the re-clamping of edge tiles -- run under the
// emulation too.  The harness puts an EmuTmap into the bytes of the CUtensorMap.
struct EmuTmap { const void* base; long long n0, n1, n2; int elem, box0, box1, box2; };
inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0ull; }                       // completed phases
inline void mbar_arrive_expect_tx(uint64_t*, uint32_t) {}
void ta_emu_yield();            // the harness: let the other fibers run (a cooperative fiber must not spin)
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    const bool done = ((uint32_t)(*bar) & 1u) != parity;
    if (!done) ta_emu_yield();
    return done;
}
inline u64 ta_globaltimer() { return 0ull; }
inline void tma_load_box_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    EmuTmap m;
    memcpy(&m, map, sizeof m);
    // the hardware faults ("illegal instruction") on a box that does not start on a 16-byte boundary of the row
    if (((long long)c0 * m.elem) % 16 != 0) { fprintf(stderr, "emu: TMA box origin %d is not 16-byte aligned\n", c0); abort(); }
    unsigned char* d = (unsigned char*)smem_dst;
    for (int k = 0; k < m.box2; ++k) for (int j = 0; j < m.box1; ++j) for (int i = 0; i < m.box0; ++i) {
        const long long g0 = c0 + i, g1 = c1 + j, g2 = c2 + k;
        unsigned char* dst = d + (((size_t)k * m.box1 + j) * m.box0 + i) * m.elem;
        if (g0 < 0 || g0 >= m.n0 || g1 < 0 || g1 >= m.n1 || g2 < 0 || g2 >= m.n2) memset(dst, 0, m.elem);
        else memcpy(dst, (const unsigned char*)m.base + ((g2 * m.n1 + g1) * m.n0 + g0) * m.elem, m.elem);
    }
    *bar += 1ull;
}
#else
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    TA_PTX("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    TA_PTX("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    TA_PTX("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return done != 0u;
}
__device__ __forceinline__ u64 ta_globaltimer() {       // nanoseconds
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tma_load_box_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    TA_PTX("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)),
                    "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

#endif

template <typename T> struct BrickShared {
    uint4* tile;                       // [TILE_SEGS]
    typename Vox<T>::Code* codes;      // [TILE_ROWS * NFS]
    uint32_t* lt_key;                  // [LT_SLOTS]
    uint32_t* lt_val;                  // [LT_SLOTS * LT_FIELDS]
    typename Vox<T>::PKey* pt_key;     // [PT_SLOTS]
    uint32_t* pt_val;                  // [PT_SLOTS * PT_WORDS]
    unsigned short* seglist;           // [SEGLIST_CAP]
    unsigned short* voxlist;           // [VOXLIST_CAP]
    unsigned short* junclist;          // [VOXLIST_CAP]
    unsigned int* ctr;                 // [0] next brick, [1] nseg, [2..3] nvox ping-pong, [4..5] njunc ping-pong,
                                       // [6..7] brick index ping-pong, [9] one-hot id table overflow
};

// ---- global flush of one label's brick-local sums -----------------------------------------------------------
__device__ __forceinline__ void label_to_global(const LabelTable& lt, uint32_t* status, uint32_t L,
                                                const uint32_t* v, u64 F0, u64 M0, u64 S0) {
    if (L >= lt.nrows) { atomicExch(&status[1], 1u); return; }
    u64 n = v[0], sf = v[1], sm = v[2], ss = v[3];
    atomicAdd(&lt.count[L], n);
    atomicAdd(&lt.s1[(size_t)L * 3 + 0], n * F0 + sf);
    atomicAdd(&lt.s1[(size_t)L * 3 + 1], n * M0 + sm);
    atomicAdd(&lt.s1[(size_t)L * 3 + 2], n * S0 + ss);
    u64* q = &lt.s2[(size_t)L * 6];
    atomicAdd(&q[0], n * F0 * F0 + 2 * F0 * sf + v[4]);
    atomicAdd(&q[1], n * F0 * M0 + F0 * sm + M0 * sf + v[5]);
    atomicAdd(&q[2], n * F0 * S0 + F0 * ss + S0 * sf + v[6]);
    atomicAdd(&q[3], n * M0 * M0 + 2 * M0 * sm + v[7]);
    atomicAdd(&q[4], n * M0 * S0 + M0 * ss + S0 * sm + v[8]);
    atomicAdd(&q[5], n * S0 * S0 + 2 * S0 * ss + v[9]);
    atomicMin(&lt.bmin[(size_t)L * 3 + 0], (int)(F0 + v[10]));
    atomicMin(&lt.bmin[(size_t)L * 3 + 1], (int)(M0 + v[11]));
    atomicMin(&lt.bmin[(size_t)L * 3 + 2], (int)(S0 + v[12]));
    atomicMax(&lt.bmax[(size_t)L * 3 + 0], (int)(F0 + v[13]));
    atomicMax(&lt.bmax[(size_t)L * 3 + 1], (int)(M0 + v[14]));
    atomicMax(&lt.bmax[(size_t)L * 3 + 2], (int)(S0 + v[15]));
}

// ---- per-brick label accumulation (brick-local coordinates, u32) -----------------------------------------
template <typename T>
__device__ __forceinline__ void label_add(const BrickShared<T>& sh, const LabelTable& lt, uint32_t* status,
                                          uint32_t L, const uint32_t* v, u64 F0, u64 M0, u64 S0) {
    uint32_t slot = (L * 0x9E3779B1u) >> (32 - LT_BITS);
    int found = -1;
#pragma unroll 1
    for (int probe = 0; probe < LT_SLOTS; ++probe) {
        uint32_t k = *((volatile uint32_t*)&sh.lt_key[slot]);
        if (k == L) { found = (int)slot; break; }
        if (k == TA_EMPTY32) {
            uint32_t old = atomicCAS(&sh.lt_key[slot], TA_EMPTY32, L);
            if (old == TA_EMPTY32 || old == L) { found = (int)slot; break; }
        }
        slot = (slot + 1) & (LT_SLOTS - 1);
    }
    if (found < 0) { label_to_global(lt, status, L, v, F0, M0, S0); return; }
    uint32_t* d = &sh.lt_val[found * LT_FIELDS];
#pragma unroll
    for (int i = 0; i < 10; ++i) atomicAdd(&d[i], v[i]);   // branch-free: zero adds are harmless
#pragma unroll
    for (int i = 10; i < 13; ++i) atomicMin(&d[i], v[i]);
#pragma unroll
    for (int i = 13; i < 16; ++i) atomicMax(&d[i], v[i]);
}

// thread-private moment accumulator for one label over the thread's (fseg, m) column
// Six registers, bit-packed (a column is at most 8 lanes x 8 planes, f < 128, s < 8):
//   pS  n [0..6]  | sum s [7..15]   | sum s^2 [16..27]     one IMAD per run: len * (1 | s<<7 | s^2<<16)
//   pF  sum f [0..12] | sum f*s [13..27]                   one IMAD per run: runsum_f * (1 | s<<13)
//   sff sum f^2
//   g   f_min | (255 - f_max) << 16   h   s_min | (255 - s_max) << 16     both updated with one vmin.u16x2
struct MomSlot {
    uint32_t label, pS, pF, sff, g, h;
    __device__ __forceinline__ void reset(uint32_t L) {
        label = L; pS = pF = sff = 0u; g = h = 0xFFFFFFFFu;
    }
    __device__ __forceinline__ void add(uint32_t len, uint32_t sfr, uint32_t sffr, uint32_t kS, uint32_t kF,
                                        uint32_t gf, uint32_t hs) {
        pS += len * kS; pF += sfr * kF; sff += sffr;
        g = __vminu2(g, gf); h = __vminu2(h, hs);
    }
    __device__ __forceinline__ uint32_t n() const { return pS & 0x7Fu; }
    __device__ __forceinline__ uint32_t ss() const { return (pS >> 7) & 0x1FFu; }
    __device__ __forceinline__ uint32_t sss() const { return pS >> 16; }
    __device__ __forceinline__ uint32_t sf() const { return pF & 0x1FFFu; }
    __device__ __forceinline__ uint32_t sfs() const { return pF >> 13; }
    __device__ __forceinline__ uint32_t fmin() const { return g & 0xFFFFu; }
    __device__ __forceinline__ uint32_t fmax() const { return 255u - (g >> 16); }
    __device__ __forceinline__ uint32_t smin() const { return h & 0xFFFFu; }
    __device__ __forceinline__ uint32_t smax() const { return 255u - (h >> 16); }
    __device__ __forceinline__ void fields(uint32_t v[LT_FIELDS], uint32_t m) const {
        const uint32_t n_ = n(), sf_ = sf(), ss_ = ss();
        v[0] = n_; v[1] = sf_; v[2] = n_ * m; v[3] = ss_; v[4] = sff; v[5] = m * sf_; v[6] = sfs();
        v[7] = n_ * m * m; v[8] = m * ss_; v[9] = sss();
        v[10] = fmin(); v[11] = m; v[12] = smin(); v[13] = fmax(); v[14] = m; v[15] = smax();
    }
};

// Column-end merge of one slot across the warp (all 32 lanes call; lanes whose slot is empty pass
// L == TA_EMPTY32).  A warp covers two m rows (lanes 0-15: m0, lanes 16-31: m0 + 1), so the m terms follow from
// the totals and the totals of the upper half: 9 full-mask redux per distinct label instead of 16.
template <typename T>
__device__ __forceinline__ void slot_flush_warp(const BrickShared<T>& sh, const LabelTable& lt, uint32_t* status,
                                                const MomSlot& sl, uint32_t m0, u64 F0, u64 M0, u64 S0, int lane) {
    const uint32_t L = sl.pS ? sl.label : TA_EMPTY32;
    const bool upper = lane >= 16;
    const uint32_t n = sl.n(), sf = sl.sf(), ss = sl.ss();
    const uint32_t w1 = n | (ss << 12);
    const uint32_t w2 = sl.sss() | ((upper ? n : 0u) << 17);
    const uint32_t w3 = sf | ((upper ? ss : 0u) << 18);
    const uint32_t w4 = sl.sfs(), w5 = sl.sff, w6 = upper ? sf : 0u;
    const uint32_t w7 = sl.fmin(), w8 = sl.fmax();
    const uint32_t w9 = sl.pS ? ((1u << sl.smin()) | (1u << sl.smax())) : 0u;
    unsigned pending = __ballot_sync(0xffffffffu, L != TA_EMPTY32);
    uint32_t r1 = 0, r2 = 0, r3 = 0, r4 = 0, r5 = 0, r6 = 0, r7 = 0, r8 = 0, r9 = 0, rb = 0;
    bool am_leader = false;
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const uint32_t Lk = __shfl_sync(0xffffffffu, L, leader);
        const bool mine = (L == Lk);
        const bool lead = (lane == leader);
        const unsigned mb = __ballot_sync(0xffffffffu, mine);
        uint32_t t;
        t = __reduce_add_sync(0xffffffffu, mine ? w1 : 0u); if (lead) r1 = t;
        t = __reduce_add_sync(0xffffffffu, mine ? w2 : 0u); if (lead) r2 = t;
        t = __reduce_add_sync(0xffffffffu, mine ? w3 : 0u); if (lead) r3 = t;
        t = __reduce_add_sync(0xffffffffu, mine ? w4 : 0u); if (lead) r4 = t;
        t = __reduce_add_sync(0xffffffffu, mine ? w5 : 0u); if (lead) r5 = t;
        t = __reduce_add_sync(0xffffffffu, mine ? w6 : 0u); if (lead) r6 = t;
        t = __reduce_min_sync(0xffffffffu, mine ? w7 : 0xFFFFFFFFu); if (lead) r7 = t;
        t = __reduce_max_sync(0xffffffffu, mine ? w8 : 0u); if (lead) r8 = t;
        t = __reduce_or_sync(0xffffffffu, mine ? w9 : 0u); if (lead) r9 = t;
        if (lead) { rb = mb; am_leader = true; }
        pending &= ~mb;
    }
    if (am_leader) {       // every group leader updates the shared table in the same SIMT pass
        const uint32_t nt = r1 & 0xFFFu, sst = r1 >> 12, ssst = r2 & 0x1FFFFu, n1 = r2 >> 17;
        const uint32_t sft = r3 & 0x3FFFFu, ss1 = r3 >> 18, sf1 = r6;
        uint32_t v[LT_FIELDS];
        v[0] = nt; v[1] = sft; v[2] = m0 * nt + n1; v[3] = sst; v[4] = r5; v[5] = m0 * sft + sf1; v[6] = r4;
        v[7] = m0 * m0 * nt + (2 * m0 + 1) * n1; v[8] = m0 * sst + ss1; v[9] = ssst;
        v[10] = r7; v[11] = (rb & 0xFFFFu) ? m0 : m0 + 1; v[12] = __ffs(r9) - 1;
        v[13] = r8; v[14] = (rb >> 16) ? m0 + 1 : m0; v[15] = 31 - __clz(r9);
        label_add(sh, lt, status, L, v, F0, M0, S0);
    }
}

// ---- per-brick pair accumulation (packed 16-bit counters; a brick has < 65536 voxels) ------------------------
// field 6 = wall18, fields 0..5 = directional faces.  idx = field+1 (wall18 -> 0): word idx>>1, half idx&1.
template <typename T>
__device__ __forceinline__ void pair_add_packed(const BrickShared<T>& sh, const PairTable& pt,
                                                typename Vox<T>::PKey key, const uint32_t inc[PT_WORDS]) {
    typedef typename Vox<T>::PKey PKey;
    uint32_t slot = Vox<T>::hash(key);
#pragma unroll 1
    for (int probe = 0; probe < PT_SLOTS; ++probe) {
        PKey k = *((volatile PKey*)&sh.pt_key[slot]);
        bool hit = (k == key);
        if (!hit && k == Vox<T>::PEMPTY) {
            PKey old = atomicCAS(&sh.pt_key[slot], Vox<T>::PEMPTY, key);
            hit = (old == Vox<T>::PEMPTY || old == key);
        }
        if (hit) {
#pragma unroll
            for (int w = 0; w < PT_WORDS; ++w) atomicAdd(&sh.pt_val[slot * PT_WORDS + w], inc[w]);
            return;
        }
        slot = (slot + 1) & (PT_SLOTS - 1);
    }
    int g = ta_pair_slot(pt, Vox<T>::key64(key));          // brick table full: straight to the global table
    if (g < 0) return;
    uint32_t* v = &pt.vals[(size_t)g * TA_PAIR_STRIDE];
#pragma unroll
    for (int idx = 0; idx < 7; ++idx) {
        uint32_t n = (inc[idx >> 1] >> ((idx & 1) * 16)) & 0xFFFFu;
        if (n) atomicAdd(&v[idx == 0 ? 6 : idx - 1], n);
    }
}

template <typename T>
__device__ __forceinline__ void pair_add(const BrickShared<T>& sh, const PairTable& pt, uint32_t a, uint32_t b,
                                         int field, uint32_t n) {
    uint32_t inc[PT_WORDS] = {0, 0, 0, 0};
    int idx = field == 6 ? 0 : field + 1;
    inc[idx >> 1] = n << ((idx & 1) * 16);
    pair_add_packed<T>(sh, pt, Vox<T>::key(a, b), inc);
}

// packed increments of one voxel (label a) towards other label d: wall18 + the faces to its +f/+m/+s neighbours
__device__ __forceinline__ void voxel_increments(uint32_t inc[PT_WORDS], bool lo, bool w18, bool ff, bool fm, bool fsl) {
    inc[0] = (w18 ? 1u : 0u) + ((ff && lo) ? (1u << 16) : 0u);
    inc[1] = ((ff && !lo) ? 1u : 0u) + ((fm && lo) ? (1u << 16) : 0u);
    inc[2] = ((fm && !lo) ? 1u : 0u) + ((fsl && lo) ? (1u << 16) : 0u);
    inc[3] = ((fsl && !lo) ? 1u : 0u);
}

// sum of j^2 for j < k, k = 0..8, from two packed byte tables (no cubic, no division)
__device__ __forceinline__ uint32_t sumsq_below(uint32_t k) {
    const uint32_t lo = 0x05010000u, hi = 0x5B371E0Eu;   // 0,0,1,5 | 14,30,55,91
    return k >= 8 ? 140u : (((k & 4u) ? hi : lo) >> ((k & 3u) * 8)) & 0xFFu;
}

// ---- SIMD helpers on one 16-byte segment ------------------------------------------------------------------------
template <typename T> struct Boundary;

template <> struct Boundary<uint16_t> {
    static __device__ __forceinline__ void cross(uint32_t acc[4], const uint4& C, const uint4* tile, int t,
                                                 bool unshifted) {
        const uint4 R = tile[t];
        const uint32_t e0 = (uint32_t)(reinterpret_cast<const unsigned short*>(tile + t)[-1]) << 16;
        const uint32_t e5 = reinterpret_cast<const unsigned short*>(tile + t + 1)[0];
        const uint32_t s0 = __funnelshift_r(e0, R.x, 16), s1 = __funnelshift_r(R.x, R.y, 16),
                       s2 = __funnelshift_r(R.y, R.z, 16), s3 = __funnelshift_r(R.z, R.w, 16),
                       s4 = __funnelshift_r(R.w, e5, 16);
        acc[0] |= (C.x ^ s0) | (C.x ^ s1);
        acc[1] |= (C.y ^ s1) | (C.y ^ s2);
        acc[2] |= (C.z ^ s2) | (C.z ^ s3);
        acc[3] |= (C.w ^ s3) | (C.w ^ s4);
        if (unshifted) { acc[0] |= C.x ^ R.x; acc[1] |= C.y ^ R.y; acc[2] |= C.z ^ R.z; acc[3] |= C.w ^ R.w; }
    }
    static __device__ __forceinline__ void diag(uint32_t acc[4], const uint4& C, const uint4* tile, int t) {
        const uint4 R = tile[t];
        acc[0] |= C.x ^ R.x; acc[1] |= C.y ^ R.y; acc[2] |= C.z ^ R.z; acc[3] |= C.w ^ R.w;
    }
    static __device__ __forceinline__ uint32_t mask(const uint32_t acc[4]) {
        const uint32_t one = 0x00010001u;
        uint32_t tt = __vminu2(acc[0], one) | (__vminu2(acc[1], one) << 2) | (__vminu2(acc[2], one) << 4) |
                      (__vminu2(acc[3], one) << 6);
        return (tt & 0x55u) | ((tt >> 15) & 0xAAu);
    }
    // bit j set iff lane j differs from lane j+1 (j = 0..6)
    static __device__ __forceinline__ uint32_t run_breaks(const uint4& C) {
        uint32_t d[4] = {C.x ^ __funnelshift_r(C.x, C.y, 16), C.y ^ __funnelshift_r(C.y, C.z, 16),
                         C.z ^ __funnelshift_r(C.z, C.w, 16), (C.w ^ (C.w >> 16)) & 0xFFFFu};
        return mask(d) & 0x7Fu;
    }
};

template <> struct Boundary<uint32_t> {
    static __device__ __forceinline__ void cross(uint32_t acc[4], const uint4& C, const uint4* tile, int t,
                                                 bool unshifted) {
        const uint4 R = tile[t];
        const uint32_t e0 = reinterpret_cast<const uint32_t*>(tile + t)[-1];
        const uint32_t e5 = reinterpret_cast<const uint32_t*>(tile + t + 1)[0];
        acc[0] |= (C.x ^ e0) | (C.x ^ R.y);
        acc[1] |= (C.y ^ R.x) | (C.y ^ R.z);
        acc[2] |= (C.z ^ R.y) | (C.z ^ R.w);
        acc[3] |= (C.w ^ R.z) | (C.w ^ e5);
        if (unshifted) { acc[0] |= C.x ^ R.x; acc[1] |= C.y ^ R.y; acc[2] |= C.z ^ R.z; acc[3] |= C.w ^ R.w; }
    }
    static __device__ __forceinline__ void diag(uint32_t acc[4], const uint4& C, const uint4* tile, int t) {
        const uint4 R = tile[t];
        acc[0] |= C.x ^ R.x; acc[1] |= C.y ^ R.y; acc[2] |= C.z ^ R.z; acc[3] |= C.w ^ R.w;
    }
    static __device__ __forceinline__ uint32_t mask(const uint32_t acc[4]) {
        return (acc[0] ? 1u : 0u) | (acc[1] ? 2u : 0u) | (acc[2] ? 4u : 0u) | (acc[3] ? 8u : 0u);
    }
    static __device__ __forceinline__ uint32_t run_breaks(const uint4& C) {
        return (C.x != C.y ? 1u : 0u) | (C.y != C.z ? 2u : 0u) | (C.z != C.w ? 4u : 0u);
    }
};

// ---- phase D: the 18 neighbours of one voxel -> first other label d0 (== a if none) and "only one other
// label" flag; also returns the +f / +m / +s neighbour labels for the face counters.
template <typename T> struct NeighbourTest;

template <> struct NeighbourTest<uint32_t> {
    template <int ROWE, int PLANEE>
    static __device__ __forceinline__ void run(const uint32_t* p, int, uint32_t& a, uint32_t& d0, bool& simple,
                                               uint32_t& nbf, uint32_t& nbm, uint32_t& nbs) {
        constexpr int offs[18] = {1, ROWE, PLANEE, -1, -ROWE, -PLANEE, -ROWE - 1, -ROWE + 1, ROWE - 1, ROWE + 1,
                                  -PLANEE - 1, -PLANEE + 1, PLANEE - 1, PLANEE + 1,
                                  -PLANEE - ROWE, -PLANEE + ROWE, PLANEE - ROWE, PLANEE + ROWE};
        a = p[0];
        uint32_t nb[18];
#pragma unroll
        for (int k = 0; k < 18; ++k) nb[k] = p[offs[k]];
        uint32_t x = 0;
#pragma unroll
        for (int k = 0; k < 18; ++k) x |= nb[k] ^ a;       // == a ^ d0 when there is a single other label
        d0 = a ^ x;
        uint32_t bad = 0;
#pragma unroll
        for (int k = 0; k < 18; ++k) bad |= min(nb[k] ^ a, nb[k] ^ d0);
        simple = (bad == 0);
        nbf = nb[0]; nbm = nb[1]; nbs = nb[2];
    }
};

// uint16: neighbours are fetched as aligned 32-bit pairs where possible and tested two at a time
// (min(v ^ a, v ^ d0) == 0 per 16-bit lane <=> v is a or d0).
template <> struct NeighbourTest<uint16_t> {
    template <int ROWE, int PLANEE>
    static __device__ __forceinline__ void run(const uint16_t* p, int j, uint32_t& a, uint32_t& d0, bool& simple,
                                               uint32_t& nbf, uint32_t& nbm, uint32_t& nbs) {
        const int odd = j & 1;
        const int far = odd ? 1 : -1;            // the lane next to the aligned pair, on its other side
        const uint16_t* pw = p - odd;            // 4-byte aligned: lanes (j & ~1, (j & ~1) + 1)
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(pw);
        const uint32_t wmm = *reinterpret_cast<const uint32_t*>(pw - ROWE);
        const uint32_t wmp = *reinterpret_cast<const uint32_t*>(pw + ROWE);
        const uint32_t wsm = *reinterpret_cast<const uint32_t*>(pw - PLANEE);
        const uint32_t wsp = *reinterpret_cast<const uint32_t*>(pw + PLANEE);
        const uint32_t e0 = p[far], emm = p[far - ROWE], emp = p[far + ROWE], esm = p[far - PLANEE],
                       esp = p[far + PLANEE];
        const uint32_t g0 = p[-PLANEE - ROWE], g1 = p[-PLANEE + ROWE], g2 = p[PLANEE - ROWE], g3 = p[PLANEE + ROWE];
        a = odd ? (w0 >> 16) : (w0 & 0xFFFFu);
        const uint32_t AA = a * 0x00010001u;
        const uint32_t r5 = e0 | (emm << 16), r6 = emp | (esm << 16), r7 = esp | (g0 << 16), r8 = g1 | (g2 << 16),
                       r9 = g3 | (a << 16);
        uint32_t acc = (w0 ^ AA) | (wmm ^ AA) | (wmp ^ AA) | (wsm ^ AA) | (wsp ^ AA) | (r5 ^ AA) | (r6 ^ AA) |
                       (r7 ^ AA) | (r8 ^ AA) | (r9 ^ AA);
        const uint32_t x = (acc | (acc >> 16)) & 0xFFFFu;
        d0 = a ^ x;
        const uint32_t DD = d0 * 0x00010001u;
        uint32_t bad = __vminu2(w0 ^ AA, w0 ^ DD) | __vminu2(wmm ^ AA, wmm ^ DD) | __vminu2(wmp ^ AA, wmp ^ DD) |
                       __vminu2(wsm ^ AA, wsm ^ DD) | __vminu2(wsp ^ AA, wsp ^ DD) | __vminu2(r5 ^ AA, r5 ^ DD) |
                       __vminu2(r6 ^ AA, r6 ^ DD) | __vminu2(r7 ^ AA, r7 ^ DD) | __vminu2(r8 ^ AA, r8 ^ DD) |
                       __vminu2(r9 ^ AA, r9 ^ DD);
        simple = (bad == 0);
        nbf = odd ? e0 : (w0 >> 16);
        nbm = odd ? (wmp >> 16) : (wmp & 0xFFFFu);
        nbs = odd ? (wsp >> 16) : (wsp & 0xFFFFu);
    }
};

// k-th offset of the 18-neighbourhood (1 <= |df|+|dm|+|ds| <= 2) in tile elements; rare-path helper
template <int ROWE, int PLANEE>
__device__ __noinline__ int neighbour_offset(int k) {
    int c = 0;
    for (int i = 0; i < 27; ++i) {
        int df = i % 3 - 1, dm = (i / 3) % 3 - 1, ds = i / 9 - 1;
        int l1 = abs(df) + abs(dm) + abs(ds);
        if (l1 >= 1 && l1 <= 2) {
            if (c == k) return ds * PLANEE + dm * ROWE + df;
            ++c;
        }
    }
    return 0;
}

// TIMING: compile the per-phase clocks in (profiling aid of a -DTA_WITH_PHASE_TIMING build, TA_PHASE_TIMING=1).  The
// product kernel carries none of it.
template <typename T, bool TIMING>
__global__ void __launch_bounds__(NTHREADS, 3)
scan_kernel(ScanParams P, LabelTable lt, PairTable pt, const __grid_constant__ CUtensorMap tmap) {
    typedef typename Vox<T>::Code Code;
    typedef typename Vox<T>::PKey PKey;
    constexpr int SEG = Vox<T>::SEG;
    constexpr uint32_t MIXED = Vox<T>::MIXED;
    constexpr int ROWE = ROWV * SEG;               // elements per tile row
    constexpr int PLANEE = (BM + 2) * ROWE;        // elements per tile plane
    constexpr int BF = NFS * SEG;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    BrickShared<T> sh;
    sh.tile = reinterpret_cast<uint4*>(smem_raw);
    sh.lt_key = reinterpret_cast<uint32_t*>(sh.tile + TILE_SEGS);
    sh.lt_val = sh.lt_key + LT_SLOTS;
    sh.pt_val = sh.lt_val + LT_SLOTS * LT_FIELDS;
    sh.pt_key = reinterpret_cast<PKey*>(sh.pt_val + PT_SLOTS * PT_WORDS);
    sh.ctr = reinterpret_cast<unsigned int*>(sh.pt_key + PT_SLOTS);
    sh.codes = reinterpret_cast<Code*>(sh.ctr + 16 + 64);
    sh.seglist = reinterpret_cast<unsigned short*>(sh.codes + TILE_ROWS * NFS);
    sh.voxlist = sh.seglist + SEGLIST_CAP;
    sh.junclist = sh.voxlist + VOXLIST_CAP;
    const T* tileT = reinterpret_cast<const T*>(sh.tile);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const T* vol = reinterpret_cast<const T*>(P.vol);
    const unsigned int total = (unsigned int)P.nbf * P.nbm * P.nbs;
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    const bool do_pairs = do_p6 || do_w18;
    const int nf = (int)P.nf, nm = (int)P.nm, ns = (int)P.ns;

    // reset the per-brick tables once; the flush at the end of each brick re-arms them
    for (int i = tid; i < LT_SLOTS; i += NTHREADS) sh.lt_key[i] = TA_EMPTY32;
    for (int i = tid; i < LT_SLOTS * LT_FIELDS; i += NTHREADS) {
        int f = i % LT_FIELDS;
        sh.lt_val[i] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
    }
    for (int i = tid; i < PT_SLOTS; i += NTHREADS) sh.pt_key[i] = Vox<T>::PEMPTY;
    for (int i = tid; i < PT_SLOTS * PT_WORDS; i += NTHREADS) sh.pt_val[i] = 0u;

    // TMA completion barrier (one arrival: the issuing thread's expect_tx); ctr[12..13] is 8-byte aligned
    uint64_t* tma_bar = reinterpret_cast<uint64_t*>(sh.ctr + 12);
    uint32_t tma_parity = 0u;
    const bool use_tma = P.use_tma && ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u) == 0u;
    if (use_tma && tid == 0) {
        mbar_init(tma_bar, 1u);
        TA_PTX("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // Phase clocks (TIMING only): thread 0 keeps per-phase cycle totals and the last time stamp in shared memory -- no
    // register lives across the phases for it -- and adds the totals to P.phase_cycles once, when the CTA is done.
    u64* sh_tick = reinterpret_cast<u64*>(sh.junclist + VOXLIST_CAP);       // [16]: [0..11] totals, [15] last stamp
    if (TIMING && tid == 0) {
        for (int k = 0; k < 15; ++k) sh_tick[k] = 0ull;
        sh_tick[15] = (u64)clock64();
    }
#define TA_TICK(k) if (TIMING && tid == 0) { const u64 now_ = (u64)clock64(); sh_tick[k] += now_ - sh_tick[15]; sh_tick[15] = now_; }

    if (tid == 0) sh.ctr[6] = atomicAdd(P.brick_counter, 1u);
    __syncthreads();
    for (unsigned iter = 0;; ++iter) {
        const unsigned int brick = sh.ctr[6 + (iter & 1u)];
        if (brick >= total) break;
        if (tid == 0) {
            // fetch the next brick index now; it is consumed after the last barrier of this iteration
            sh.ctr[6 + ((iter + 1u) & 1u)] = atomicAdd(P.brick_counter, 1u);
            sh.ctr[1] = 0u; sh.ctr[2] = 0u; sh.ctr[3] = 0u; sh.ctr[4] = 0u; sh.ctr[5] = 0u; sh.ctr[9] = 0u;
        }
        TA_TICK(0);
        const int bf = brick % P.nbf, bm = (brick / P.nbf) % P.nbm, bs = brick / (P.nbf * P.nbm);
        const int F0 = bf * BF, M0 = bm * BM, S0 = (int)P.own_lo + bs * BS;
        const u64 gF0 = (u64)F0, gM0 = (u64)M0, gS0 = (u64)((long long)S0 + P.slow_offset);

        // ---- phase A: stage brick + halo (clamped) ------------------------------------------------------------
        auto stage_tile = [&]() {
        if (use_tma) {
            // the callers' barrier ordered every earlier generic-proxy access of the tile before this point
            if (tid == 0) {
                TA_PTX("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_expect_tx(tma_bar, (uint32_t)(TILE_SEGS * 16));
                tma_load_box_3d(sh.tile, &tmap, tma_bar, F0 - SEG, M0 - 1, S0 - 1);
            }
            // Thread 0 may still be in a divergent tail of the previous iteration (table flush, phase clocks) when its
            // warp mates get here.  They must not start polling before it has issued the copy: a warp whose other lanes
            // spin in try_wait can starve the one lane the barrier is waiting for (seen as a lost copy with
            // TA_PHASE_TIMING=1).  Converge the warp first.
            __syncwarp();
            // A lost copy must not hang the box.  The limit is wall time (5 s of %globaltimer, looked at every 4096
            // polls), not a poll count -- a time-sliced or debugged GPU polls for a long time without anything being wrong
            // -- and the failure is REPORTED, not trapped: every thread of the CTA waits on this barrier, so every thread
            // sees the time-out and leaves the kernel; the host finds the diagnostic (P.diag, host-mapped) after the
            // pass, returns TA_ERR_CUDA with its text and the context stays usable (no sticky error).
            unsigned spins = 0;
            u64 t_first = 0ull;
            bool lost = false;
            while (!mbar_try_wait(tma_bar, tma_parity)) {
                if ((++spins & 0xFFFu) == 0u) {
                    const u64 now = ta_globaltimer();
                    if (t_first == 0ull) t_first = now;
                    else if (now - t_first > 5000000000ull) { lost = true; break; }
                }
            }
            if (lost) {
                if (P.diag && atomicAdd(&P.diag[0], 1ull) == 0ull) {
                    P.diag[1] = ((u64)blockIdx.x << 32) | (u64)tid;
                    P.diag[2] = ((u64)iter << 32) | (u64)brick;
                    P.diag[3] = ((u64)tma_parity << 32) | (u64)sh.ctr[6 + ((iter + 1u) & 1u)];
                    P.diag[4] = *reinterpret_cast<volatile u64*>(tma_bar);
                    P.diag[5] = ((u64)(uint32_t)(F0 - SEG) << 32) | ((u64)(uint32_t)(M0 - 1) << 16) | (u64)(uint32_t)(S0 - 1);
                    __threadfence_system();
                }
                return;
            }
            tma_parity ^= 1u;
            // Elements outside the buffer arrive as zeros; the tile wants them clamped (replicated edge voxels).  Only
            // bricks on a face of the buffer pay for the patch: f lanes, then m rows, then s planes.
            const bool edge = (F0 == 0) | (F0 + BF + 1 > nf) | (M0 == 0) | (M0 + BM + 1 > nm) | (S0 < 1) | (S0 + BS + 1 > ns);
            if (edge) {
                T* tw = reinterpret_cast<T*>(sh.tile);
                const int xl = (F0 == 0) ? SEG : 0;                       // lanes [0, xl) <- lane xl
                const int xr = min(ROWE, nf - F0 + SEG);                  // lanes [xr, ROWE) <- lane xr - 1
                for (int r = tid; r < TILE_ROWS; r += NTHREADS) {
                    T* row = tw + r * ROWE;
                    if (xl) { const T v = row[xl]; for (int x = 0; x < xl; ++x) row[x] = v; }
                    if (xr < ROWE) { const T v = row[xr - 1]; for (int x = xr; x < ROWE; ++x) row[x] = v; }
                }
                __syncthreads();
                for (int i = tid; i < TILE_SEGS; i += NTHREADS) {
                    const int r = i / ROWV, m = r % (BM + 2) - 1;
                    const int mc = min(max(M0 + m, 0), nm - 1) - M0;
                    if (mc != m) sh.tile[i] = sh.tile[i + (mc - m) * ROWV];
                }
                __syncthreads();
                for (int i = tid; i < TILE_SEGS; i += NTHREADS) {
                    const int s = i / PLANEV - 1;
                    const int sc = min(max(S0 + s, 0), ns - 1) - S0;
                    if (sc != s) sh.tile[i] = sh.tile[i + (sc - s) * PLANEV];
                }
            }
            return;
        }
        for (int i = tid; i < TILE_SEGS; i += NTHREADS) {
            const int fs = i % ROWV - 1;
            const int r = i / ROWV;
            const int m = r % (BM + 2) - 1, s = r / (BM + 2) - 1;
            const int gs = min(max(S0 + s, 0), ns - 1);
            const int gm = min(max(M0 + m, 0), nm - 1);
            const int gf = F0 + fs * SEG;
            const T* row = vol + ((size_t)gs * nm + gm) * (size_t)nf;
            uint4 v;
            if (P.vec_ok) {
                // rows are whole segments: in-range segments are asynchronous 16-byte copies (all in flight at
                // once); out-of-range halo segments replicate the edge voxel
                const int gfc = min(max(gf, 0), nf - SEG);
                if (gf == gfc) { cp_async_16(&sh.tile[i], row + gf); continue; }
                v = ld_stream_128(row + gfc);
                uint32_t e = (gf < 0) ? ((SEG == 8) ? (v.x & 0xFFFFu) : v.x) : ((SEG == 8) ? (v.w >> 16) : v.w);
                if (SEG == 8) e |= e << 16;
                v.x = v.y = v.z = v.w = e;
            } else {
                T tmp[SEG];
#pragma unroll
                for (int j = 0; j < SEG; ++j) tmp[j] = row[min(max(gf + j, 0), nf - 1)];
                if (SEG == 8) {
                    v.x = (uint32_t)tmp[0] | ((uint32_t)tmp[1] << 16);
                    v.y = (uint32_t)tmp[2] | ((uint32_t)tmp[3] << 16);
                    v.z = (uint32_t)tmp[4 % SEG] | ((uint32_t)tmp[5 % SEG] << 16);
                    v.w = (uint32_t)tmp[6 % SEG] | ((uint32_t)tmp[7 % SEG] << 16);
                } else {
                    v.x = tmp[0]; v.y = tmp[1]; v.z = tmp[2 % SEG]; v.w = tmp[3 % SEG];
                }
            }
            sh.tile[i] = v;
        }
        cp_async_wait_all();
        };
        stage_tile();
        __syncthreads();
        TA_TICK(1);

        if (P.flags & 0x100u) continue;   // debug: staging only
        // ---- phase B: per row-segment uniformity code (label if the SEG+2 voxels are equal) -------------------
        const uint32_t ref_label = tileT[SEG];            // first in-brick-row element of the tile
        bool all_ref = true;
        for (int i = tid; i < TILE_ROWS * NFS; i += NTHREADS) {
            const int fs = i % NFS, r = i / NFS;
            const T* rp = tileT + r * ROWE + (fs + 1) * SEG;
            const uint4 v = sh.tile[r * ROWV + fs + 1];
            const uint32_t l = rp[0];
            const uint32_t pat = (SEG == 8) ? (l | (l << 16)) : l;
            const bool uni = (v.x == pat) & (v.y == pat) & (v.z == pat) & (v.w == pat) &
                             ((uint32_t)rp[-1] == l) & ((uint32_t)rp[SEG] == l);
            sh.codes[i] = (Code)(uni ? l : MIXED);
            all_ref = all_ref && uni && (l == ref_label);
        }
        // Whole tile (brick + halo) is one label (background, or the inside of a large cell): closed-form moments,
        // no pairs.  MIXED-coded labels (0xFFFF in uint16 volumes) never take this path.
        if (__syncthreads_and(all_ref && ref_label != MIXED) && !(P.flags & 0x300u)) {
            if (tid == 0 && do_mom) {
                const uint32_t a = (uint32_t)min(BF, nf - F0), b = (uint32_t)min(BM, nm - M0),
                               c = (uint32_t)min(BS, (int)P.own_hi - S0);
                const uint32_t ta = a * (a - 1) / 2, tb = b * (b - 1) / 2, tc = c * (c - 1) / 2;
                const uint32_t qa = (a - 1) * a * (2 * a - 1) / 6, qb = (b - 1) * b * (2 * b - 1) / 6,
                               qc = (c - 1) * c * (2 * c - 1) / 6;
                uint32_t v[LT_FIELDS];
                v[0] = a * b * c; v[1] = b * c * ta; v[2] = a * c * tb; v[3] = a * b * tc;
                v[4] = b * c * qa; v[5] = c * ta * tb; v[6] = b * ta * tc;
                v[7] = a * c * qb; v[8] = a * tb * tc; v[9] = a * b * qc;
                v[10] = 0; v[11] = 0; v[12] = 0; v[13] = a - 1; v[14] = b - 1; v[15] = c - 1;
                label_to_global(lt, pt.status, ref_label, v, gF0, gM0, gS0);
            }
            TA_TICK(2);
            continue;
        }

        TA_TICK(2);
        if (P.flags & 0x200u) continue;   // debug: staging + codes only
        // ---- phase C1: march (moments, interior test, segment worklist) -------------------------------------------
        {
            const int fs = tid % NFS, m = tid / NFS;
            const int gf0 = F0 + fs * SEG, gm = M0 + m;
            const bool col_valid = (gf0 < nf) && (gm < nm);
            const int nvalid = col_valid ? min(SEG, nf - gf0) : 0;
            const int smax = min(BS, (int)P.own_hi - S0);
            const uint32_t lf0 = fs * SEG;
            const uint32_t rowsum = SEG * lf0 + SEG * (SEG - 1) / 2;
            const uint32_t rowsq = SEG * lf0 * lf0 + lf0 * SEG * (SEG - 1) + (SEG - 1) * SEG * (2 * SEG - 1) / 6;

            // three slots, most recently used first; a fourth label in one column evicts the oldest (rare)
            MomSlot S0_, S1_, S2_;
            S0_.reset(TA_EMPTY32); S1_.reset(TA_EMPTY32); S2_.reset(TA_EMPTY32);
            // Hits add in place (no slot moves).  `last` / `prev` are the two most recently used slot indices; the
            // victim of a miss is the remaining slot, moved into S2_ first so the eviction code exists once.
            int last = 0, prev = 1;
            auto account = [&](uint32_t L, uint32_t len, uint32_t sfr, uint32_t sffr, uint32_t kS, uint32_t kF,
                               uint32_t gf, uint32_t hs) {
                int k;
                if (L == S0_.label) { S0_.add(len, sfr, sffr, kS, kF, gf, hs); k = 0; }
                else if (L == S1_.label) { S1_.add(len, sfr, sffr, kS, kF, gf, hs); k = 1; }
                else if (L == S2_.label) { S2_.add(len, sfr, sffr, kS, kF, gf, hs); k = 2; }
                else {
                    const int victim = 3 - last - prev;
                    if (victim == 0) { const MomSlot t = S0_; S0_ = S2_; S2_ = t; }
                    else if (victim == 1) { const MomSlot t = S1_; S1_ = S2_; S2_ = t; }
                    if (last == 2) last = victim; else if (prev == 2) prev = victim;   // the old S2_ moved there
                    if (S2_.pS) {
                        if (TIMING) atomicAdd(&pt.status[2], 1u);   // profiling aid: eviction count
                        uint32_t v[LT_FIELDS];
                        S2_.fields(v, (uint32_t)m);
                        label_add<T>(sh, lt, pt.status, S2_.label, v, gF0, gM0, gS0);
                    }
                    S2_.reset(L);
                    S2_.add(len, sfr, sffr, kS, kF, gf, hs);
                    k = 2;
                }
                if (k != last) { prev = last; last = k; }
            };
            auto tcode = [&](int s) -> uint32_t {
                const int base = ((s + 1) * (BM + 2) + (m + 1)) * NFS + fs;
                const uint32_t e0 = sh.codes[base - NFS], e1 = sh.codes[base], e2 = sh.codes[base + NFS];
                return (e0 == e1 && e1 == e2) ? e1 : MIXED;
            };

            uint32_t t_prev = tcode(-1), t_cur = tcode(0);
            for (int s = 0; s < BS; ++s) {
                const bool active = col_valid && (s < smax);
                const uint32_t t_next = tcode(s + 1);
                const uint32_t e_c = sh.codes[((s + 1) * (BM + 2) + (m + 1)) * NFS + fs];
                const bool interior = (t_cur != MIXED) && (t_prev == t_cur) && (t_next == t_cur);

                if (do_mom && active) {
                    // runs of equal labels inside the segment (one run when the code says "uniform")
                    const bool uni = (e_c != MIXED);
                    const int tv = (s + 1) * PLANEV + (m + 1) * ROWV + (fs + 1);
                    const T* cp = reinterpret_cast<const T*>(sh.tile + tv);
                    uint32_t brk = 0u;
                    if (!uni) brk = Boundary<T>::run_breaks(sh.tile[tv]);
                    const uint32_t us = (uint32_t)s;
                    const uint32_t kS = 1u | (us << 7) | ((us * us) << 16), kF = 1u | (us << 13);
                    const uint32_t hs = us | ((255u - us) << 16);
                    int j0 = 0;
                    while (j0 < nvalid) {
                        const uint32_t rest = brk >> j0;
                        const int j1 = min(rest ? j0 + __ffs(rest) : SEG, nvalid);
                        const uint32_t L = uni ? e_c : (uint32_t)cp[j0];
                        const uint32_t len = j1 - j0;
                        uint32_t sfr = rowsum, sffr = rowsq;
                        if (len != SEG) {
                            const uint32_t sj = (uint32_t)(j0 + j1 - 1) * len / 2;
                            const uint32_t sjj = sumsq_below((uint32_t)j1) - sumsq_below((uint32_t)j0);
                            sfr = len * lf0 + sj;
                            sffr = len * lf0 * lf0 + 2 * lf0 * sj + sjj;
                        }
                        account(L, len, sfr, sffr, kS, kF, (lf0 + j0) | ((255u - (lf0 + j1 - 1)) << 16), hs);
                        j0 = j1;
                    }
                }
                // warp-aggregated append of non-interior segments
                const bool want = do_pairs && active && !interior;
                const unsigned ball = __ballot_sync(0xffffffffu, want);
                if (ball) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(&sh.ctr[1], (unsigned)__popc(ball));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (want) sh.seglist[base + __popc(ball & ((1u << lane) - 1u))] =
                        (unsigned short)((s * BM + m) * NFS + fs);
                }
                t_prev = t_cur; t_cur = t_next;
            }
            if (do_mom) {
                const uint32_t m0 = (uint32_t)(m & ~1);
                slot_flush_warp<T>(sh, lt, pt.status, S0_, m0, gF0, gM0, gS0, lane);
                slot_flush_warp<T>(sh, lt, pt.status, S1_, m0, gF0, gM0, gS0, lane);
                slot_flush_warp<T>(sh, lt, pt.status, S2_, m0, gF0, gM0, gS0, lane);
            }
        }
        __syncthreads();
        TA_TICK(3);

        // ---- phases C2 + D in rounds of NTHREADS listed segments ---------------------------------------------------------
        if (do_pairs) {
            const int nseg = (int)sh.ctr[1];
            for (int base = 0, round = 0; base < nseg; base += NTHREADS, ++round) {
                unsigned int* nvox = &sh.ctr[2 + (round & 1)];
                unsigned int* njunc = &sh.ctr[4 + (round & 1)];
                if (tid == 0) { sh.ctr[2 + ((round + 1) & 1)] = 0u; sh.ctr[4 + ((round + 1) & 1)] = 0u; }
                // C2: boundary bits of one listed segment per thread
                const int idx = base + tid;
                uint32_t bits = 0, id = 0, ebase = 0;      // ebase: tile element offset of the segment
                if (idx < nseg) {
                    id = sh.seglist[idx];
                    const int fs = id % NFS, m = (id / NFS) % BM, s = id / (NFS * BM);
                    const int t = (s + 1) * PLANEV + (m + 1) * ROWV + (fs + 1);
                    ebase = (uint32_t)t * SEG;
                    const uint4 C = sh.tile[t];
                    uint32_t acc[4] = {0u, 0u, 0u, 0u};
                    Boundary<T>::cross(acc, C, sh.tile, t, false);
                    Boundary<T>::cross(acc, C, sh.tile, t - ROWV, true);
                    Boundary<T>::cross(acc, C, sh.tile, t + ROWV, true);
                    Boundary<T>::cross(acc, C, sh.tile, t - PLANEV, true);
                    Boundary<T>::cross(acc, C, sh.tile, t + PLANEV, true);
                    Boundary<T>::diag(acc, C, sh.tile, t - PLANEV - ROWV);
                    Boundary<T>::diag(acc, C, sh.tile, t - PLANEV + ROWV);
                    Boundary<T>::diag(acc, C, sh.tile, t + PLANEV - ROWV);
                    Boundary<T>::diag(acc, C, sh.tile, t + PLANEV + ROWV);
                    bits = Boundary<T>::mask(acc);
                    const int left = nf - (F0 + fs * SEG);
                    if (left < SEG) bits &= (1u << left) - 1u;
                }
                {
                    // warp exclusive scan of popcounts, one shared atomic per warp
                    const unsigned cnt = __popc(bits);
                    unsigned inc = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        unsigned y = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += y;
                    }
                    unsigned wtot = __shfl_sync(0xffffffffu, inc, 31);
                    unsigned wbase = 0;
                    if (lane == 31 && wtot) wbase = atomicAdd(nvox, wtot);
                    wbase = __shfl_sync(0xffffffffu, wbase, 31);
                    unsigned pos = wbase + inc - cnt;
                    while (bits) {
                        int j = __ffs(bits) - 1;
                        bits &= bits - 1;
                        sh.voxlist[pos++] = (unsigned short)(ebase + j);   // tile element offset (< 65536)
                    }
                }
                __syncthreads();
                TA_TICK(4);

                // D: listed voxels.  Each thread takes DCH consecutive list entries (they come from one segment, so
                // they mostly share the label pair) and sums their packed counters in registers; one warp merge per
                // chunk then updates the shared pair table.
#ifndef TA_DCH
#define TA_DCH 4          // voxels per thread and sweep in phase D (2 and 8 measured slower on C3)
#endif
                constexpr int DCH = TA_DCH;
                const int nv = (int)*nvox;
                for (int ib = 0; ib < nv; ib += NTHREADS * DCH) {
                    PKey key = Vox<T>::PEMPTY;
                    uint32_t inc[PT_WORDS] = {0u, 0u, 0u, 0u};
#pragma unroll 1
                    for (int c = 0; c < DCH; ++c) {
                        const int i = ib + tid * DCH + c;
                        bool junction = false;
                        uint32_t e = 0;
                        if (i < nv) {
                            e = sh.voxlist[i];
                            const int j = e & (SEG - 1);
                            const T* p = tileT + e;
                            uint32_t a, d0, nbf, nbm, nbs;
                            bool simple;
                            NeighbourTest<T>::template run<ROWE, PLANEE>(p, j, a, d0, simple, nbf, nbm, nbs);
                            if (d0 != a) {
                                if (simple) {
                                    const PKey k2 = Vox<T>::key(a, d0);
                                    uint32_t v[PT_WORDS];
                                    voxel_increments(v, a < d0, do_w18, do_p6 && nbf != a, do_p6 && nbm != a,
                                                     do_p6 && nbs != a);
                                    if (k2 != key && key != Vox<T>::PEMPTY) {   // pair changed inside the chunk (rare)
                                        pair_add_packed<T>(sh, pt, key, inc);
                                        inc[0] = inc[1] = inc[2] = inc[3] = 0u;
                                    }
                                    key = k2;
                                    inc[0] += v[0]; inc[1] += v[1]; inc[2] += v[2]; inc[3] += v[3];
                                } else {
                                    junction = true;
                                }
                            }
                        }
                        // junction voxels -> third worklist
                        const unsigned ball = __ballot_sync(0xffffffffu, junction);
                        if (ball) {
                            unsigned jb = 0;
                            if (lane == 0) jb = atomicAdd(njunc, (unsigned)__popc(ball));
                            jb = __shfl_sync(0xffffffffu, jb, 0);
                            if (junction) sh.junclist[jb + __popc(ball & ((1u << lane) - 1u))] = (unsigned short)e;
                        }
                    }
                    // one shared-table update per distinct pair in the warp (warp-uniform loop, full-mask redux)
                    {
                        unsigned pending = __ballot_sync(0xffffffffu, key != Vox<T>::PEMPTY);
                        uint32_t tot[PT_WORDS] = {0u, 0u, 0u, 0u};
                        bool am_leader = false;
                        while (pending) {
                            const int leader = __ffs(pending) - 1;
                            const PKey kk = __shfl_sync(0xffffffffu, key, leader);
                            const bool mine = (key == kk);
#pragma unroll
                            for (int w = 0; w < PT_WORDS; ++w) {
                                const uint32_t r = __reduce_add_sync(0xffffffffu, mine ? inc[w] : 0u);
                                if (lane == leader) tot[w] = r;
                            }
                            am_leader = am_leader || (lane == leader);
                            pending &= ~__ballot_sync(0xffffffffu, mine);
                        }
                        // all group leaders update the shared pair table in one SIMT pass
                        if (am_leader) pair_add_packed<T>(sh, pt, key, tot);
                    }
                }
                __syncthreads();
                TA_TICK(5);

                // D2: junction voxels: distinct other labels in registers (up to 4), one packed add per label
                const int nj = (int)*njunc;
                for (int i = tid; i < nj; i += NTHREADS) {
                    const uint32_t e = sh.junclist[i];
                    const T* p = tileT + e;
                    const uint32_t a = p[0];
                    constexpr int offs[18] = {
                        1, ROWE, PLANEE, -1, -ROWE, -PLANEE,
                        -ROWE - 1, -ROWE + 1, ROWE - 1, ROWE + 1,
                        -PLANEE - 1, -PLANEE + 1, PLANEE - 1, PLANEE + 1,
                        -PLANEE - ROWE, -PLANEE + ROWE, PLANEE - ROWE, PLANEE + ROWE};
                    uint32_t nb[18];
#pragma unroll
                    for (int k = 0; k < 18; ++k) nb[k] = p[offs[k]];
                    uint32_t d0 = a, d1 = a, d2 = a, d3 = a;
                    int nd = 0;
#pragma unroll
                    for (int k = 0; k < 18; ++k) {
                        const uint32_t b = nb[k];
                        const bool isnew = (b != a) & (b != d0) & (b != d1) & (b != d2) & (b != d3);
                        d0 = (isnew && nd == 0) ? b : d0;
                        d1 = (isnew && nd == 1) ? b : d1;
                        d2 = (isnew && nd == 2) ? b : d2;
                        d3 = (isnew && nd == 3) ? b : d3;
                        nd += isnew ? 1 : 0;
                    }
                    if (nd <= 4) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t d = q == 0 ? d0 : q == 1 ? d1 : q == 2 ? d2 : d3;
                            if (q < nd) {
                                uint32_t inc[PT_WORDS];
                                voxel_increments(inc, a < d, do_w18, do_p6 && nb[0] == d, do_p6 && nb[1] == d,
                                                 do_p6 && nb[2] == d);
                                pair_add_packed<T>(sh, pt, Vox<T>::key(a, d), inc);
                            }
                        }
                    } else {
                        // more than four distinct other labels (noise-like data): exact first-occurrence rescan
                        if (do_p6) {
                            if (nb[0] != a) pair_add<T>(sh, pt, a, nb[0], a < nb[0] ? 0 : 1, 1u);
                            if (nb[1] != a) pair_add<T>(sh, pt, a, nb[1], a < nb[1] ? 2 : 3, 1u);
                            if (nb[2] != a) pair_add<T>(sh, pt, a, nb[2], a < nb[2] ? 4 : 5, 1u);
                        }
                        if (do_w18) {
#pragma unroll 1
                            for (int k = 0; k < 18; ++k) {
                                const uint32_t b = p[neighbour_offset<ROWE, PLANEE>(k)];
                                if (b == a) continue;
                                bool seen = false;
                                for (int q = 0; q < k; ++q)
                                    seen |= ((uint32_t)p[neighbour_offset<ROWE, PLANEE>(q)] == b);
                                if (!seen) pair_add<T>(sh, pt, a, b, 6, 1u);
                            }
                        }
                    }
                }
                __syncthreads();
                TA_TICK(6);
            }
        }

        // ---- phase F: flush the per-brick tables ---------------------------------------------------------------------
        if (!(P.flags & 0x400u)) {
            for (int i = tid; i < LT_SLOTS; i += NTHREADS) {
                uint32_t L = sh.lt_key[i];
                if (L == TA_EMPTY32) continue;
                uint32_t* d = &sh.lt_val[i * LT_FIELDS];
                label_to_global(lt, pt.status, L, d, gF0, gM0, gS0);
#pragma unroll
                for (int f = 0; f < LT_FIELDS; ++f) d[f] = (f >= 10 && f < 13) ? 0xFFFFFFFFu : 0u;
                sh.lt_key[i] = TA_EMPTY32;
            }
            for (int i = tid; i < PT_SLOTS; i += NTHREADS) {
                PKey key = sh.pt_key[i];
                if (key == Vox<T>::PEMPTY) continue;
                uint32_t* d = &sh.pt_val[i * PT_WORDS];
                int slot = ta_pair_slot(pt, Vox<T>::key64(key));
#pragma unroll
                for (int idx = 0; idx < 7; ++idx) {
                    uint32_t n = (d[idx >> 1] >> ((idx & 1) * 16)) & 0xFFFFu;
                    if (n && slot >= 0) atomicAdd(&pt.vals[(size_t)slot * TA_PAIR_STRIDE + (idx == 0 ? 6 : idx - 1)], n);
                }
#pragma unroll
                for (int w = 0; w < PT_WORDS; ++w) d[w] = 0u;
                sh.pt_key[i] = Vox<T>::PEMPTY;
            }
        }
        __syncthreads();
        TA_TICK(7);
    }
#undef TA_TICK
    if (TIMING && tid == 0 && P.phase_cycles)
        for (int k = 0; k < 12; ++k) atomicAdd(&P.phase_cycles[k], sh_tick[k]);
}

}  // namespace ta

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
// One-hot pair path (instantiation OH = true, flag 0x1000; an experiment kept testable, not the default -- it is exact
// but slower than C2 / D / D2 on every measured configuration):
//   R  after the march the tile is rewritten in place: label -> one-hot word (1 << id) of a brick-local id.  Ids come
//      from NID-slot tables (NID = bits of a label word), one per quarter of the row for uint16 so that 16 ids suffice;
//      the two lanes either side of a quarter boundary are kept in both encodings (edge array).
//   S  per listed segment, SIMD on the packed lanes: the OR of the 18 neighbour words is the SET of labels around
//      each voxel (the de-duplication of the wall18 definition is the OR itself); `& ~own` leaves the other labels.
//      Counts per (own id, other id) are sums of one bit column; they go to a direct-indexed table of packed 16-bit
//      counters [wall18 | +f] [+m | +s] after a warp merge.  No per-voxel work, no junction special case.
//   A quarter with more labels than ids (noise-like data): the brick is staged again and takes C2 / D / D2.
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
    uint32_t* idk;                     // [NQ * NID = 64] one-hot id tables: label of id (slot index), TA_EMPTY32 when free
    uint32_t* ohc;                     // [OH_CTR_WORDS] one-hot pair counters (quarter, own id, other id) x 2 words;
                                       // aliases voxlist + junclist (per-voxel path only)
    void* edg;                         // one-hot edge array (see OneHot); aliases pt_val + pt_key (per-voxel path only)
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

// ---- one-hot pair path ------------------------------------------------------------------------------------------------
// The segment arithmetic is host-compilable (TA_HD) so that tests/host/oh_host_check.cu can run exactly this code on the
// CPU against a brute-force count; the kernel is the only product caller.
#ifdef __CUDA_ARCH__
#define TA_HD __device__ __forceinline__
#else
#define TA_HD __host__ __device__ inline
#endif
TA_HD uint32_t ta_funnel_r16(uint32_t lo, uint32_t hi) {          // (hi:lo) >> 16
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, 16);
#else
    return (hi << 16) | (lo >> 16);
#endif
}
TA_HD uint32_t ta_vminu2(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __vminu2(a, b);
#else
    const uint32_t l = (a & 0xFFFFu) < (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
    const uint32_t h = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return l | (h << 16);
#endif
}
TA_HD int ta_popc(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
TA_HD int ta_ffs(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __ffs(x);
#else
    return __builtin_ffs((int)x);
#endif
}

// Brick-local label ids.  A tile row is cut into NQ groups of QS = NFS / NQ segments ("quarters" for uint16); every
// quarter has its own NID-slot open-addressing table (id = slot index, one-hot word = 1 << id), so that NID = the bits
// of a label word is enough for cells down to a few voxels across.  A voxel is encoded with the table of its quarter.
// The two lanes on either side of a quarter boundary are needed by the stencil of the segment across the boundary, in
// THAT segment's encoding: the edge array holds them (edg[(row * (NQ - 1) + (b - 1)) * 2 + side], boundary b = 1 ..
// NQ - 1 between quarters b - 1 and b; side 0: last lane of quarter b - 1 in table b, side 1: first lane of quarter b
// in table b - 1).
template <typename T> struct OneHot;

template <> struct OneHot<uint16_t> {
    static constexpr int NID = 16, LOG_NID = 4, NQ = 4;
    static constexpr uint32_t LANE1 = 0x00010001u;                 // bit 0 of every lane
    static TA_HD uint32_t splat(uint32_t oh) { return oh * 0x00010001u; }
    static TA_HD uint32_t fold(uint32_t x) { return (x | (x >> 16)) & 0xFFFFu; }     // ids in a word
    static TA_HD uint32_t full(uint32_t lanebits) { return lanebits * 0xFFFFu; }     // LANE1 bits -> whole-lane masks
    static TA_HD uint32_t lanesum(uint32_t t) { return (t + (t >> 16)) & 0xFFFFu; }  // sum of the lanes of a word
    // 0xFFFF in every lane of c that equals the lane of pat
    static TA_HD uint32_t eqmask(uint32_t c, uint32_t pat) {
        return (ta_vminu2(c ^ pat, 0x00010001u) ^ 0x00010001u) * 0xFFFFu;
    }
    // whole-lane mask of lanes [j0, j1) restricted to word w
    static TA_HD uint32_t lane_range(int w, int j0, int j1) {
        return ((2 * w >= j0 && 2 * w < j1) ? 0xFFFFu : 0u) | ((2 * w + 1 >= j0 && 2 * w + 1 < j1) ? 0xFFFF0000u : 0u);
    }
    static TA_HD uint32_t run_breaks(const uint4& C) {             // bit j <=> lane j differs from lane j + 1 (j < 7)
        const uint32_t d0 = C.x ^ ta_funnel_r16(C.x, C.y), d1 = C.y ^ ta_funnel_r16(C.y, C.z),
                       d2 = C.z ^ ta_funnel_r16(C.z, C.w), d3 = (C.w ^ (C.w >> 16)) & 0xFFFFu;
        const uint32_t one = 0x00010001u;
        const uint32_t tt = ta_vminu2(d0, one) | (ta_vminu2(d1, one) << 2) | (ta_vminu2(d2, one) << 4) |
                            (ta_vminu2(d3, one) << 6);
        return ((tt & 0x55u) | ((tt >> 15) & 0xAAu)) & 0x7Fu;
    }
};
template <> struct OneHot<uint32_t> {
    static constexpr int NID = 32, LOG_NID = 5, NQ = 1;
    static constexpr uint32_t LANE1 = 1u;
    static TA_HD uint32_t splat(uint32_t oh) { return oh; }
    static TA_HD uint32_t fold(uint32_t x) { return x; }
    static TA_HD uint32_t full(uint32_t lanebits) { return 0u - lanebits; }
    static TA_HD uint32_t lanesum(uint32_t t) { return t; }
    static TA_HD uint32_t eqmask(uint32_t c, uint32_t pat) { return c == pat ? 0xFFFFFFFFu : 0u; }
    static TA_HD uint32_t lane_range(int w, int j0, int j1) { return (w >= j0 && w < j1) ? 0xFFFFFFFFu : 0u; }
    static TA_HD uint32_t run_breaks(const uint4& C) {
        return (C.x != C.y ? 1u : 0u) | (C.y != C.z ? 2u : 0u) | (C.z != C.w ? 4u : 0u);
    }
};
constexpr int OH_CTR_WORDS = 2048;      // NQ * NID * NID * 2 for both label widths

// find-or-insert label L in one quarter's table; the slot index is the id.  -1: table full.
template <typename T>
TA_HD int oh_insert(uint32_t* idk, uint32_t L) {
    constexpr int NID = OneHot<T>::NID;
    uint32_t slot = (L * 0x9E3779B1u) >> (32 - OneHot<T>::LOG_NID);
    for (int probe = 0; probe < NID; ++probe) {
#ifdef __CUDA_ARCH__
        const uint32_t k = *((volatile uint32_t*)&idk[slot]);
        if (k == L) return (int)slot;
        if (k == TA_EMPTY32) {
            const uint32_t old = atomicCAS(&idk[slot], TA_EMPTY32, L);
            if (old == TA_EMPTY32 || old == L) return (int)slot;
        }
#else
        if (idk[slot] == L) return (int)slot;
        if (idk[slot] == TA_EMPTY32) { idk[slot] = L; return (int)slot; }
#endif
        slot = (slot + 1) & (NID - 1);
    }
    return -1;
}

// Phase R for one segment (row r, segment fs) of the tile: labels -> one-hot words of the quarter's ids, in place, plus
// the encodings of its boundary lanes that the neighbouring quarter / the f-halo need.  Only the segment's own lanes
// (and, for the first / last segment of a row, the f-halo lane beside it) are read, so all segments can be rewritten
// concurrently.  `code`: the uniformity code of phase B.  false: some table is full (the tile is then unusable for
// both pair paths and has to be staged again).
template <typename T>
TA_HD bool oh_relabel_segment(uint4* tile, T* edg, uint32_t* idk, int r, int fs, uint32_t code, uint32_t& lastL,
                              uint32_t& lastQ, uint32_t& lastOH) {
    constexpr int SEG = Vox<T>::SEG;
    constexpr int NID = OneHot<T>::NID, NQ = OneHot<T>::NQ, QS = NFS / NQ;
    constexpr int ROWE = ROWV * SEG;
    const uint32_t q = (uint32_t)(fs / QS);
    bool ok = true;
    auto enc = [&](uint32_t qq, uint32_t L) -> uint32_t {          // one-hot of L in quarter qq (one-entry cache)
        if (L == lastL && qq == lastQ) return lastOH;
        const int slot = oh_insert<T>(idk + qq * NID, L);
        if (slot < 0) { ok = false; return 0u; }
        lastL = L; lastQ = qq; lastOH = 1u << slot;
        return lastOH;
    };
    uint4* seg = tile + r * ROWV + fs + 1;
    T* rp = reinterpret_cast<T*>(tile) + r * ROWE + (fs + 1) * SEG;
    uint32_t first, last;                                           // labels of lane 0 and lane SEG - 1
    if (code != Vox<T>::MIXED) {
        const uint32_t w = OneHot<T>::splat(enc(q, code));
        first = last = code;
        if (fs == 0) rp[-1] = (T)lastOH;                            // code != MIXED: the f-halo lanes carry the same label
        if (fs == NFS - 1) rp[SEG] = (T)lastOH;
        *seg = make_uint4(w, w, w, w);
    } else {
        const uint4 v = *seg;
        uint32_t brk = OneHot<T>::run_breaks(v);
        uint32_t out[4] = {0u, 0u, 0u, 0u};
        first = rp[0]; last = rp[SEG - 1];
        const uint32_t hl = (fs == 0) ? (uint32_t)rp[-1] : 0u, hr = (fs == NFS - 1) ? (uint32_t)rp[SEG] : 0u;
        int j0 = 0;
        while (j0 < SEG) {
            const uint32_t rest = brk >> j0;
            const int j1 = rest ? j0 + ta_ffs(rest) : SEG;
            const uint32_t w = OneHot<T>::splat(enc(q, rp[j0]));
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] |= w & OneHot<T>::lane_range(k, j0, j1);
            j0 = j1;
        }
        if (fs == 0) rp[-1] = (T)enc(q, hl);
        if (fs == NFS - 1) rp[SEG] = (T)enc(q, hr);
        *seg = make_uint4(out[0], out[1], out[2], out[3]);
    }
    if (NQ > 1) {
        const int b0 = fs / QS, pos = fs % QS;
        if (pos == 0 && b0 > 0) edg[(r * (NQ - 1) + (b0 - 1)) * 2 + 1] = (T)enc(q - 1, first);
        if (pos == QS - 1 && b0 < NQ - 1) edg[(r * (NQ - 1) + b0) * 2 + 0] = (T)enc(q + 1, last);
    }
    return ok;
}

// One listed segment (row r of the tile, segment fs) of the one-hot tile: C = own words; oth = labels in the
// 18-neighbourhood other than the own one; xf / xm / xs = the +f / +m / +s neighbour's word where it differs from the
// own one (else 0).  The lanes beside the segment come from the tile or, across a quarter boundary, from the edge array.
template <typename T>
TA_HD void oh_stencil(const uint4* tile, const T* edg, int r, int fs, uint32_t C[4], uint32_t oth[4], uint32_t xf[4],
                      uint32_t xm[4], uint32_t xs[4]) {
    constexpr int SEG = Vox<T>::SEG;
    constexpr int NQ = OneHot<T>::NQ, QS = NFS / NQ;
    constexpr int ROWE = ROWV * SEG, PLANEE = (BM + 2) * ROWE;
    constexpr int EROW = (NQ - 1) * 2, EPLANE = (BM + 2) * EROW;
    const int t = r * ROWV + fs + 1;
    const uint4 c = tile[t];
    const uint4 m0 = tile[t - ROWV], m1 = tile[t + ROWV], s0 = tile[t - PLANEV], s1 = tile[t + PLANEV];
    const uint4 d0 = tile[t - PLANEV - ROWV], d1 = tile[t - PLANEV + ROWV], d2 = tile[t + PLANEV - ROWV],
                d3 = tile[t + PLANEV + ROWV];
    const T* e = reinterpret_cast<const T*>(tile + t);
    const T* pl = e - 1;
    const T* pr = e + SEG;
    int lrow = ROWE, lplane = PLANEE, rrow = ROWE, rplane = PLANEE;
    if (NQ > 1) {
        const int b0 = fs / QS, pos = fs % QS;
        if (pos == 0 && b0 > 0) { pl = edg + (r * (NQ - 1) + (b0 - 1)) * 2 + 0; lrow = EROW; lplane = EPLANE; }
        if (pos == QS - 1 && b0 < NQ - 1) { pr = edg + (r * (NQ - 1) + b0) * 2 + 1; rrow = EROW; rplane = EPLANE; }
    }
    const uint32_t cR = pr[0];
    const uint32_t yL = (uint32_t)pl[0] | pl[-lrow] | pl[lrow] | pl[-lplane] | pl[lplane];
    const uint32_t yR = cR | pr[-rrow] | pr[rrow] | pr[-rplane] | pr[rplane];
    C[0] = c.x; C[1] = c.y; C[2] = c.z; C[3] = c.w;
    const uint32_t X[4] = {m0.x | m1.x | s0.x | s1.x, m0.y | m1.y | s0.y | s1.y, m0.z | m1.z | s0.z | s1.z,
                           m0.w | m1.w | s0.w | s1.w};
    const uint32_t D[4] = {d0.x | d1.x | d2.x | d3.x, d0.y | d1.y | d2.y | d3.y, d0.z | d1.z | d2.z | d3.z,
                           d0.w | d1.w | d2.w | d3.w};
    const uint32_t Y[4] = {C[0] | X[0], C[1] | X[1], C[2] | X[2], C[3] | X[3]};   // rows whose f-shifts count
    uint32_t nxt[4];                                                             // centre row shifted by +f
    if (SEG == 8) {
        const uint32_t h0 = ta_funnel_r16(yL << 16, Y[0]), h1 = ta_funnel_r16(Y[0], Y[1]),
                       h2 = ta_funnel_r16(Y[1], Y[2]), h3 = ta_funnel_r16(Y[2], Y[3]),
                       h4 = ta_funnel_r16(Y[3], yR);
        oth[0] = (h0 | h1 | X[0] | D[0]) & ~C[0];
        oth[1] = (h1 | h2 | X[1] | D[1]) & ~C[1];
        oth[2] = (h2 | h3 | X[2] | D[2]) & ~C[2];
        oth[3] = (h3 | h4 | X[3] | D[3]) & ~C[3];
        nxt[0] = ta_funnel_r16(C[0], C[1]); nxt[1] = ta_funnel_r16(C[1], C[2]);
        nxt[2] = ta_funnel_r16(C[2], C[3]); nxt[3] = ta_funnel_r16(C[3], cR);
    } else {
        oth[0] = (yL | Y[1] | X[0] | D[0]) & ~C[0];
        oth[1] = (Y[0] | Y[2] | X[1] | D[1]) & ~C[1];
        oth[2] = (Y[1] | Y[3] | X[2] | D[2]) & ~C[2];
        oth[3] = (Y[2] | yR | X[3] | D[3]) & ~C[3];
        nxt[0] = C[1]; nxt[1] = C[2]; nxt[2] = C[3]; nxt[3] = cR;
    }
    xf[0] = nxt[0] & ~C[0]; xf[1] = nxt[1] & ~C[1]; xf[2] = nxt[2] & ~C[2]; xf[3] = nxt[3] & ~C[3];
    xm[0] = m1.x & ~C[0]; xm[1] = m1.y & ~C[1]; xm[2] = m1.z & ~C[2]; xm[3] = m1.w & ~C[3];
    xs[0] = s1.x & ~C[0]; xs[1] = s1.y & ~C[1]; xs[2] = s1.z & ~C[2]; xs[3] = s1.w & ~C[3];
}

// number of lanes (of the four words) whose bit j is set
template <typename T>
TA_HD uint32_t oh_column_count(const uint32_t o[4], uint32_t j) {
    constexpr uint32_t L1 = OneHot<T>::LANE1;
    return OneHot<T>::lanesum(((o[0] >> j) & L1) + ((o[1] >> j) & L1) + ((o[2] >> j) & L1) + ((o[3] >> j) & L1));
}

// Every (own id i, other id j) contribution of one listed segment, `left` voxels of it inside the volume:
// emit((quarter * NID + i) * NID + j, [wall18 | +f faces], [+m faces | +s faces]) as seen from the voxels of label i.
// The loop runs over the OTHER labels j (one or two per wall segment); the lanes that see j almost always carry one
// own label, so there is no loop over own labels on the common path.
template <typename T, typename Emit>
TA_HD void oh_segment_pairs(const uint4* tile, const T* edg, int r, int fs, int left, uint32_t keep0, bool do_p6,
                            Emit&& emit) {
    constexpr int SEG = Vox<T>::SEG;
    constexpr int NID = OneHot<T>::NID, QS = NFS / OneHot<T>::NQ;
    constexpr uint32_t L1 = OneHot<T>::LANE1;
    uint32_t C[4], oth[4], xf[4], xm[4], xs[4];
    oh_stencil<T>(tile, edg, r, fs, C, oth, xf, xm, xs);
    if (left < SEG) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t vm = OneHot<T>::lane_range(w, 0, left);
            C[w] &= vm; oth[w] &= vm; xf[w] &= vm; xm[w] &= vm; xs[w] &= vm;
        }
    }
    uint32_t J = OneHot<T>::fold(oth[0] | oth[1] | oth[2] | oth[3]);
    const uint32_t kbase = (uint32_t)(fs / QS) * NID;
    while (J) {
        const uint32_t j = ta_ffs(J) - 1;
        J &= J - 1u;
        // lanes that have label j around them, as whole-lane masks
        const uint32_t l0 = OneHot<T>::full((oth[0] >> j) & L1), l1 = OneHot<T>::full((oth[1] >> j) & L1),
                       l2 = OneHot<T>::full((oth[2] >> j) & L1), l3 = OneHot<T>::full((oth[3] >> j) & L1);
        uint32_t own = OneHot<T>::fold((C[0] & l0) | (C[1] & l1) | (C[2] & l2) | (C[3] & l3));
        if ((own & (own - 1u)) == 0u) {
            const uint32_t i = ta_ffs(own) - 1;
            const uint32_t c0 = (oh_column_count<T>(oth, j) | (oh_column_count<T>(xf, j) << 16)) & keep0;
            const uint32_t c1 = do_p6 ? (oh_column_count<T>(xm, j) | (oh_column_count<T>(xs, j) << 16)) : 0u;
            emit((kbase + i) * NID + j, c0, c1);
        } else {
            while (own) {                                           // lanes of several own labels see j (rare)
                const uint32_t i = ta_ffs(own) - 1;
                own &= own - 1u;
                const uint32_t pat = OneHot<T>::splat(1u << i);
                uint32_t o[4], f[4], mm[4], ss[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const uint32_t M = OneHot<T>::eqmask(C[w], pat);
                    o[w] = oth[w] & M; f[w] = xf[w] & M; mm[w] = xm[w] & M; ss[w] = xs[w] & M;
                }
                const uint32_t c0 = (oh_column_count<T>(o, j) | (oh_column_count<T>(f, j) << 16)) & keep0;
                const uint32_t c1 = do_p6 ? (oh_column_count<T>(mm, j) | (oh_column_count<T>(ss, j) << 16)) : 0u;
                emit((kbase + i) * NID + j, c0, c1);
            }
        }
    }
}

// OH: compile the one-hot pair path (phases R + S) in.  The product launches OH = false (per-voxel pair path only, the
// faster one on every measured configuration) unless flag 0x1000 asks for the one-hot path.
// TIMING: compile the per-phase clocks in (profiling aid, TA_PHASE_TIMING=1).  The product kernel carries none of it.
template <typename T, bool OH, bool TIMING>
__global__ void __launch_bounds__(NTHREADS, 3)
scan_kernel(ScanParams P, LabelTable lt, PairTable pt, const __grid_constant__ CUtensorMap tmap) {
    typedef typename Vox<T>::Code Code;
    typedef typename Vox<T>::PKey PKey;
    constexpr int SEG = Vox<T>::SEG;
    constexpr int LOG_SEG = Vox<T>::LOG_SEG;
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
    sh.idk = reinterpret_cast<uint32_t*>(sh.ctr + 16);
    sh.codes = reinterpret_cast<Code*>(sh.idk + 64);
    sh.seglist = reinterpret_cast<unsigned short*>(sh.codes + TILE_ROWS * NFS);
    sh.voxlist = sh.seglist + SEGLIST_CAP;
    sh.junclist = sh.voxlist + VOXLIST_CAP;
    sh.ohc = reinterpret_cast<uint32_t*>(sh.voxlist);
    sh.edg = sh.pt_val;
    constexpr int NID = OneHot<T>::NID, NQ = OneHot<T>::NQ;
    static_assert(NQ * NID * NID * 2 == OH_CTR_WORDS && OH_CTR_WORDS * 4 <= 2 * VOXLIST_CAP * 2,
                  "one-hot counters must fit the two voxel worklists");
    static_assert(NQ * NID <= 64, "id tables");
    static_assert(TILE_ROWS * (NQ - 1) * 2 * sizeof(T) <= PT_SLOTS * PT_WORDS * 4 + PT_SLOTS * sizeof(PKey),
                  "edge array must fit the per-brick pair table");
    const T* tileT = reinterpret_cast<const T*>(sh.tile);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const T* vol = reinterpret_cast<const T*>(P.vol);
    const unsigned int total = (unsigned int)P.nbf * P.nbm * P.nbs;
    const bool do_mom = P.flags & 1u, do_p6 = P.flags & 2u, do_w18 = P.flags & 4u;
    const bool do_pairs = do_p6 || do_w18;
    // Pair path: per-voxel (C2 / D / D2) or, in the OH instantiation, one-hot (R + S).  Both are exact.
    const bool oh_enabled = OH && do_pairs && !(P.flags & 0x800u);
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
        if (tid < 64) sh.idk[tid] = TA_EMPTY32;
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
            unsigned spins = 0;
            while (!mbar_try_wait(tma_bar, tma_parity)) {
                if (++spins > (1u << 18)) {                  // ~1 s: a lost copy must not hang the box; fail loudly
                    if (P.diag && atomicAdd(&P.diag[0], 1ull) == 0ull) {
                        P.diag[1] = ((u64)blockIdx.x << 32) | (u64)tid;
                        P.diag[2] = ((u64)iter << 32) | (u64)brick;
                        P.diag[3] = ((u64)tma_parity << 32) | (u64)sh.ctr[6 + ((iter + 1u) & 1u)];
                        P.diag[4] = *reinterpret_cast<volatile u64*>(tma_bar);
                        P.diag[5] = ((u64)(uint32_t)(F0 - SEG) << 32) | ((u64)(uint32_t)(M0 - 1) << 16) | (u64)(uint32_t)(S0 - 1);
                        __threadfence_system();
                    }
                    __trap();
                }
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

        // ---- one-hot pair path: phases R + S -----------------------------------------------------------------------------
        bool use_oh = oh_enabled && (sh.ctr[1] != 0u);      // uniform over the CTA (ctr[1] is final since the barrier)
        if (use_oh) {
            const int nseg = (int)sh.ctr[1];
            T* edg = reinterpret_cast<T*>(sh.edg);
            for (int i = tid; i < OH_CTR_WORDS; i += NTHREADS) sh.ohc[i] = 0u;
            // R: rewrite the tile in place, label -> one-hot word of its id in the quarter's table
            {
                uint32_t lastL = TA_EMPTY32, lastQ = 0u, lastOH = 0u;
                bool ok = true;
                for (int i = tid; i < TILE_ROWS * NFS; i += NTHREADS)
                    ok = oh_relabel_segment<T>(sh.tile, edg, sh.idk, i / NFS, i % NFS, sh.codes[i], lastL, lastQ, lastOH) && ok;
                if (!ok) sh.ctr[9] = 1u;
            }
            __syncthreads();
            TA_TICK(8);
            if (sh.ctr[9] != 0u) {
                // a quarter holds more than NID labels (noise-like data): the tile is half rewritten, stage it again and
                // take the per-voxel path.  The edge array lies in the per-brick pair table: re-arm that.
                use_oh = false;
                for (int i = tid; i < PT_SLOTS; i += NTHREADS) sh.pt_key[i] = Vox<T>::PEMPTY;
                for (int i = tid; i < PT_SLOTS * PT_WORDS; i += NTHREADS) sh.pt_val[i] = 0u;
                stage_tile();
                __syncthreads();
            }
        }
        if (TIMING && tid == 0 && do_pairs && P.phase_cycles) atomicAdd(&P.phase_cycles[use_oh ? 12 : 13], 1ull);   // bricks per path
        if (use_oh) {
            const int nseg = (int)sh.ctr[1];
            const T* edg = reinterpret_cast<const T*>(sh.edg);
            // S: listed segments.  Per (quarter, own id i, other id j): [wall18 | +f faces] [+m faces | +s faces] as seen
            // from the voxels of label i.
            const uint32_t keep0 = (do_w18 ? 0xFFFFu : 0u) | (do_p6 ? 0xFFFF0000u : 0u);
            for (int base = 0; base < nseg; base += NTHREADS) {
                const int idx = base + tid;
                uint32_t key0 = TA_EMPTY32, a0 = 0u, a1 = 0u;
                if (idx < nseg) {
                    const uint32_t id = sh.seglist[idx];
                    const int fs = id % NFS, m = (id / NFS) % BM, s = id / (NFS * BM);
                    oh_segment_pairs<T>(sh.tile, edg, (s + 1) * (BM + 2) + (m + 1), fs, nf - (F0 + fs * SEG), keep0, do_p6,
                        [&](uint32_t key, uint32_t c0, uint32_t c1) {
                            if (key0 == TA_EMPTY32) { key0 = key; a0 = c0; a1 = c1; }    // first one: through the warp merge
                            else {
                                if (c0) atomicAdd(&sh.ohc[2 * key], c0);
                                if (c1) atomicAdd(&sh.ohc[2 * key + 1], c1);
                            }
                        });
                }
                // one shared-counter update per distinct (own, other) in the warp: uniform loop, full-mask redux
                {
                    unsigned pending = __ballot_sync(0xffffffffu, key0 != TA_EMPTY32);
                    uint32_t t0 = 0u, t1 = 0u;
                    bool am_leader = false;
                    while (pending) {
                        const int leader = __ffs(pending) - 1;
                        const uint32_t kk = __shfl_sync(0xffffffffu, key0, leader);
                        const bool mine = (key0 == kk);
                        const uint32_t r0 = __reduce_add_sync(0xffffffffu, mine ? a0 : 0u);
                        const uint32_t r1 = __reduce_add_sync(0xffffffffu, mine ? a1 : 0u);
                        if (lane == leader) { t0 = r0; t1 = r1; am_leader = true; }
                        pending &= ~__ballot_sync(0xffffffffu, mine);
                    }
                    if (am_leader) {
                        if (t0) atomicAdd(&sh.ohc[2 * key0], t0);
                        if (t1) atomicAdd(&sh.ohc[2 * key0 + 1], t1);
                    }
                }
            }
            __syncthreads();
            TA_TICK(9);
        }

        // ---- phases C2 + D in rounds of NTHREADS listed segments (bricks with more than NID labels) ------------------
        if (do_pairs && !use_oh) {
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
                if (use_oh) break;             // the per-brick pair table held the one-hot edge array: re-armed below
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
            if (use_oh) {
                if (NQ > 1) {
                    for (int i = tid; i < PT_SLOTS; i += NTHREADS) sh.pt_key[i] = Vox<T>::PEMPTY;
                    for (int i = tid; i < PT_SLOTS * PT_WORDS; i += NTHREADS) sh.pt_val[i] = 0u;
                }
                // one-hot counters: entries (q, i, j) and (q, j, i) belong to the same label pair
                for (int e = tid; e < NQ * NID * NID; e += NTHREADS) {
                    const int q = e / (NID * NID), i = (e / NID) % NID, j = e % NID;
                    if (i >= j) continue;
                    const uint32_t* cij = &sh.ohc[2 * ((q * NID + i) * NID + j)];
                    const uint32_t* cji = &sh.ohc[2 * ((q * NID + j) * NID + i)];
                    const uint32_t p0 = cij[0], p1 = cij[1], q0 = cji[0], q1 = cji[1];
                    if (!(p0 | p1 | q0 | q1)) continue;
                    const uint32_t La = sh.idk[q * NID + i], Lb = sh.idk[q * NID + j];
                    const int slot = ta_pair_slot(pt, ta_pair_key(La, Lb));
                    if (slot < 0) continue;
                    uint32_t* v = &pt.vals[(size_t)slot * TA_PAIR_STRIDE];
                    // a face seen from its lower-index voxel: slot 2a when that voxel carries the smaller label
                    const int lo = La < Lb ? 0 : 1;
                    const uint32_t w18 = (p0 & 0xFFFFu) + (q0 & 0xFFFFu);
                    if (w18) atomicAdd(&v[6], w18);
                    if (p0 >> 16) atomicAdd(&v[0 + lo], p0 >> 16);
                    if (q0 >> 16) atomicAdd(&v[1 - lo], q0 >> 16);
                    if (p1 & 0xFFFFu) atomicAdd(&v[2 + lo], p1 & 0xFFFFu);
                    if (q1 & 0xFFFFu) atomicAdd(&v[3 - lo], q1 & 0xFFFFu);
                    if (p1 >> 16) atomicAdd(&v[4 + lo], p1 >> 16);
                    if (q1 >> 16) atomicAdd(&v[5 - lo], q1 >> 16);
                }
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

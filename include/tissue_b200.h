/*
 * tissue_b200.h -- C ABI of libtissue_b200.so: the B200 (sm_100a) per-label voxel scan behind
 * tissue_analysis' SpatialImageAnalysis3D feature extractors.
 *
 * The reference has NO native boundary: its hot path is a Python class whose methods call
 * numpy / scipy.ndimage per label (reference file SIA =
 * src/vplants/tissue_analysis/spatial_image_analysis.py).  Each entry point below names the reference
 * computation it replaces; the Python mirror of the class (tissue_analysis_b200/spatial_image_analysis.py)
 * binds these with ctypes and rebuilds the reference's dict / list return values on the host.
 *
 * Conventions
 *   - every function returns TA_OK (0) or a negative TA_ERR_* code; ta_last_error() gives the text.
 *   - plain pointers and sizes only; "host" pointers are ordinary memory, "device" pointers are CUDA
 *     device memory of the context's GPU.  No pointer returned by the library outlives the context.
 *   - axes are MEMORY axes: fast (contiguous), mid, slow.  The caller maps them to API axes (x,y,z).
 *   - a context is single-threaded: serialise calls on one context.
 *   - there is no CPU fallback: without a CUDA device ta_ctx_create fails with TA_ERR_CUDA.
 */
#ifndef TISSUE_B200_H
#define TISSUE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ta_ctx ta_ctx;

enum {
    TA_OK = 0,
    TA_ERR_CUDA = -1,          /* CUDA runtime error (text in ta_last_error)                    */
    TA_ERR_BAD_ARG = -2,       /* bad dtype / shape / null pointer / call order                 */
    TA_ERR_PAIR_OVERFLOW = -3, /* pair hash table full: call ta_run_pass again with more capacity */
    TA_ERR_LABEL_RANGE = -4,   /* a label exceeds the label table (max_label_hint too small)    */
    TA_ERR_NO_VOLUME = -5,     /* ta_run_pass before ta_bind_volume                             */
    TA_ERR_NO_TABLES = -6      /* fetch before a successful ta_run_pass                         */
};

/* ta_run_pass flags.  MOMENTS: count, sum x/y/z, sum xx..zz, bbox (SIA:1231 nd.sum, SIA:517
 * nd.find_objects, SIA:466 nd.center_of_mass, SIA:1261-1278 centred coordinates).  PAIRS6: six directional
 * face counts per (min,max) pair (SIA:45-60 neighbours, SIA:695-716 + 947-956 one-sided dilations).
 * WALL18: 18-connected wall-voxel count per pair (SIA:796-799, 835-863). */
#define TA_PASS_MOMENTS 1u
#define TA_PASS_PAIRS6 2u
#define TA_PASS_WALL18 4u
#define TA_PASS_ALL 7u
/* Leave the pair records in hash order (no sort by (lo, hi)): for passes whose records only feed
 * ta_merge_pair_records on the ranks of a sharded run, which sorts the merged table anyway. */
#define TA_PASS_UNSORTED 0x2000u
/* No host synchronisation inside ta_run_pass (the sharded driver's steady state): everything is queued on the stream, the
 * pair records are packed on the device behind a header row {count} (ta_pair_records_deferred; pair_capacity_hint = rows of
 * that buffer) and the status flags -- pair / record overflow, label range -- are checked by the first call that hands
 * results to the host (ta_*_size, ta_fetch_*, ta_pair_records_device), which returns the error then. */
#define TA_PASS_DEFERRED 0x4000u
/* Measurement switches, never needed for results: 0x800, in a -DTA_WITH_PHASE_TIMING build only, makes the scan kernel fetch
 * every tile and drop it (times the box copies on their own); 0x100 / 0x200 / 0x400 stop round 1's kernel (TA_SCAN_KERNEL=brick)
 * after staging / after the uniformity codes / before the table flush. */

const char* ta_version(void);

/* One context per (process, GPU).  device < 0 keeps the calling thread's current CUDA device. */
int ta_ctx_create(ta_ctx** out, int device);
int ta_ctx_destroy(ta_ctx* ctx);
const char* ta_last_error(ta_ctx* ctx);

/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the context stream. */
int ta_set_stream(ta_ctx* ctx, void* cuda_stream);

/* Bind the label volume (replaces `self.image`, SIA:224-227).  elem_bytes is 2 (uint16) or 4 (uint32).
 * is_device == 0: `data` is host memory, copied to a context-owned device buffer (H2D on the context stream;
 * pinned memory makes it asynchronous).  is_device != 0: `data` is borrowed, never written. */
int ta_bind_volume(ta_ctx* ctx, const void* data, int is_device, int elem_bytes,
                   int64_t n_fast, int64_t n_mid, int64_t n_slow);

/* z-slab sharding (no reference equivalent: the reference is one process).  The bound buffer holds planes
 * [slow_offset, slow_offset + n_slow) of a taller global volume; this rank OWNS buffer planes
 * [own_lo, own_hi): it accumulates moments and wall18 counts for owned voxels and the faces whose lower voxel
 * it owns.  Planes outside [own_lo, own_hi) are read-only halo.  Default: owns everything, offset 0. */
int ta_set_slab(ta_ctx* ctx, int64_t own_lo, int64_t own_hi, int64_t slow_offset);

/* The single streaming pass.  max_label_hint: largest label value (0 = find it; uint16 always uses 65535).
 * pair_capacity_hint: expected number of distinct touching pairs (0 = default).  Returns
 * TA_ERR_PAIR_OVERFLOW (tables invalid) if the pair table filled up. */
int ta_run_pass(ta_ctx* ctx, uint32_t flags, uint32_t max_label_hint, uint64_t pair_capacity_hint);

/* ta_run_pass for a bound volume whose planes become valid in stages (no reference equivalent; the sharded driver
 * overlaps the NCCL halo exchange with the scan of the interior planes).  Range k = owned planes
 * [lo_hi[2k], lo_hi[2k+1]) is scanned after CUDA event wait_events[k] (cudaEvent_t, NULL = at once).  The ranges must
 * tile the owned planes exactly; empty ranges are skipped.  Results are those of ta_run_pass. */
int ta_run_pass_ranges(ta_ctx* ctx, uint32_t flags, uint32_t max_label_hint, uint64_t pair_capacity_hint, int n_ranges,
                       const int64_t* lo_hi, void* const* wait_events);

/* ta_bind_volume(host) + ta_run_pass with the copy and the scan overlapped: the host volume (pinned memory makes the
 * copy asynchronous) goes to the context-owned device buffer in chunks of `chunk_planes` slow-axis planes (0 = about
 * 64 MiB) on a copy stream, and the scan of each chunk is queued behind its copy, accumulating into the same tables.
 * Same results and error codes as the two calls; afterwards the volume is bound (second passes, ta_run_pass again).
 * This is what the reference's `SpatialImageAnalysis(image)` + first feature request costs end to end (SIA:212-270,
 * 1231): the image lives in host memory.  uint32 volumes with max_label_hint == 0 are copied first (the table height
 * needs the largest label) and then scanned. */
int ta_run_pass_host(ta_ctx* ctx, const void* host_data, int elem_bytes, int64_t n_fast, int64_t n_mid, int64_t n_slow,
                     const int64_t* slab /* NULL, or {own_lo, own_hi, slow_offset} as in ta_set_slab */,
                     uint32_t flags, uint32_t max_label_hint, uint64_t pair_capacity_hint, int64_t chunk_planes);

/* Per-label table, dense by label value: rows 0..n-1.
 *   count[n]     voxels                                  (SIA:1231)
 *   s1[n][3]     sum of fast/mid/slow global indices      (SIA:466)
 *   s2[n][6]     sum of ff, fm, fs, mm, ms, ss products   (SIA:150)
 *   bbox[n][6]   min fast,mid,slow then max fast,mid,slow (inclusive; min>max when absent) (SIA:517) */
int ta_label_table_size(ta_ctx* ctx, uint64_t* n);
int ta_fetch_label_table(ta_ctx* ctx, uint64_t* count, uint64_t* s1, uint64_t* s2, int32_t* bbox);

/* Pair table, sorted by (lo, hi), lo < hi.
 *   faces[n][6]  slot 2a: faces normal to memory axis a whose lower-index voxel has label lo;
 *                slot 2a+1: ... has label hi                     (SIA:947-956)
 *   wall18[n]    voxels of lo or hi with an 18-neighbour of the other label (SIA:860) */
int ta_pair_table_size(ta_ctx* ctx, uint64_t* n);
int ta_fetch_pair_table(ta_ctx* ctx, uint32_t* lo, uint32_t* hi, uint32_t* faces, uint32_t* wall18);

/* Device views for the multi-GPU merge (torch.distributed all_reduce / all_gather run on these buffers).
 * label table: count u64[n], s1 u64[n*3], s2 u64[n*6], bbox i32[n*6] (min rows are all_reduce MIN, max rows
 * MAX: bbox is stored as two blocks: bmin i32[n*3], bmax i32[n*3]).
 * pair records: packed rows of 9 uint32: lo, hi, faces[6], wall18. */
int ta_label_table_device(ta_ctx* ctx, void** count, void** s1, void** s2, void** bmin, void** bmax,
                          uint64_t* n);
int ta_pair_records_device(ta_ctx* ctx, void** records, uint64_t* n);
/* Replace the pair table by the sum-merge of `n` packed device records (gathered from all ranks). */
int ta_merge_pair_records(ta_ctx* ctx, const void* device_records, uint64_t n);
/* The same for a TA_PASS_DEFERRED pass, without host synchronisation.  ta_pair_records_deferred: this rank's record
 * buffer, uint32[1 + cap_rows][9] on the device -- row 0 = {count, 0, ...}, then `count` records in hash order -- ready
 * for a fixed-size all_gather.  ta_merge_pair_records_deferred: `gathered` = the buffers of all `world` ranks back to
 * back; the pair table becomes their sum-merge.  Counts never reach the host; sorting and the status check happen at
 * the first fetch. */
int ta_pair_records_deferred(ta_ctx* ctx, void** records, uint64_t* cap_rows);
int ta_merge_pair_records_deferred(ta_ctx* ctx, const void* gathered, uint64_t cap_rows, int world);

/* Batched inertia axes (replaces compute_covariance_matrix SIA:137-150 + eigen_values_vectors SIA:152-167).
 * For each listed label: covariance = central second moments / max(3, count) from the exact integer sums
 * (128-bit), then a 3x3 Jacobi eigen-solve in fp64.  evals[n][3] descending; evecs[n][9] rows = eigenvectors,
 * in MEMORY axis order.  labels == NULL means rows 0..n-1 of the table. */
int ta_inertia_from_moments(ta_ctx* ctx, const uint32_t* labels, uint64_t n, double* evals, double* evecs);
/* The same for every row of the label table, results kept on the device (and copied out when the pointers are
 * non-null: evals[nrows][3], evecs[nrows][9]); used after the multi-GPU merge and by the benchmark. */
int ta_inertia_table(ta_ctx* ctx, double* evals, double* evecs);
/* Same eigen-solve for caller-provided symmetric matrices cov[n][6] = (a00,a01,a02,a11,a12,a22). */
int ta_inertia_eig(ta_ctx* ctx, const double* cov, uint64_t n, double* evals, double* evecs);

/* Second pass: coordinates of the wall voxels (SIA:799, 860 np.where order) for `npairs` pairs given as
 * (lo, hi).  Call with xyz == NULL to get counts[npairs]; then with xyz sized 3*sum(counts) int64: for pair i
 * the block starting at 3*offset_i holds fast[], mid[], slow[] index rows, each voxel list sorted by
 * (slow, mid, fast)-major = memory order.  The caller re-sorts to API order when axes are permuted. */
int ta_wall_voxel_coords(ta_ctx* ctx, const uint32_t* lo, const uint32_t* hi, uint64_t npairs,
                         uint64_t* counts, int64_t* xyz);

/* out[p] = labels of voxels 6-adjacent to `background`, others 0; keep_background adds 1 on background
 * voxels (SIA:1024-1038).  `out_host` has the volume's shape and dtype. */
int ta_voxel_first_layer(ta_ctx* ctx, uint32_t background, int keep_background, void* out_host);

/* Wall mask of hollow_out_cells (SIA:74-94) / get_all_wall_binary_image (SIA:744-749): out[p] = volume[p] where the
 * discrete Laplacian (scipy.ndimage.laplace: 'reflect' border, wrap-around arithmetic of the label dtype) is non-zero,
 * else 0; with mask_only != 0 the output is 1 / 0 instead (label 0 can sit on a wall).  `out_host` has the volume's
 * shape and dtype. */
int ta_hollow_out_cells(ta_ctx* ctx, int mask_only, void* out_host);

/* Outer voxel shell of every cell at once, cells_voxel_layer (SIA:1399-1448: mask minus its 18-connected erosion inside
 * the cell's bounding box): out[p] = 1 where some 18-neighbour of p is outside the volume or carries another label,
 * else 0.  `out_host` has the volume's shape and dtype. */
int ta_cell_shell18(ta_ctx* ctx, void* out_host);

/* out[p] = lut[volume[p]] (a streaming gather).  Replaces the label -> value image builders
 * (PropertySpatialImage.create_property_image, property_spatial_image.py:207-221; spatial_image_analysis_to_spatial_image,
 * tissue_analysis_oalab/sia_to_spatial_image.py:26-55) and, with in_place != 0, the relabelling image mutators
 * fuse_labels_in_image / remove_labels_from_image (SIA:1114-1165).  `lut_host` has n_lut entries of lut_elem_bytes
 * (2 or 4) bytes; labels >= n_lut map to `fill`.  in_place: the context's own device copy of the volume is rewritten
 * (lut_elem_bytes must equal the volume's; a borrowed device volume is refused) and the tables are invalidated.
 * `out_host`, when not NULL, receives the mapped volume (shape of the volume, lut_elem_bytes per voxel). */
int ta_map_labels(ta_ctx* ctx, const void* lut_host, int lut_elem_bytes, uint64_t n_lut, uint32_t fill,
                  void* out_host, int in_place);

/* Milliseconds (CUDA events on the context stream) of the last ta_run_pass: scan kernel(s) only, and the
 * whole pass including table clear / compaction / sort; and of the last host->device volume copy. */
int ta_last_timing(ta_ctx* ctx, float* scan_ms, float* pass_ms, float* h2d_ms);
/* Number of kernels this library launched since the context was created. */
int ta_launch_count(ta_ctx* ctx, uint64_t* n);

/* Bench / test utility (not a reference function): seeded integer Voronoi tessellation written to device
 * memory.  seeds[ncell][3] are fixed-point positions (1/16 voxel) in memory-axis order; weight[3] are the
 * integer axis weights (voxel size ratios); labels are seed index + 2; dome != 0 writes label 1 outside the
 * inscribed ellipsoid of semi-axes 0.47 * global dims.  Only planes [slow_offset, slow_offset+n_slow) of the
 * global volume (global_slow planes) are written. */
int ta_synth_voronoi(ta_ctx* ctx, void* device_out, int elem_bytes, int64_t n_fast, int64_t n_mid,
                     int64_t n_slow, int64_t slow_offset, int64_t global_slow, const int32_t* seeds_host,
                     uint32_t ncell, const int32_t* weight, int dome);

#ifdef __cplusplus
}
#endif
#endif /* TISSUE_B200_H */

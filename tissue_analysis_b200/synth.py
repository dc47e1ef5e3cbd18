"""Seeded synthetic tissues: integer Voronoi tessellations (bench / test input, not a reference feature).

Definition (pure integer arithmetic, so the numpy and the CUDA generator agree bit for bit):
  * ``ncell`` seeds at fixed-point positions (1/16 voxel) drawn by ``np.random.default_rng(seed)``;
  * voxel (x, y, z) has centre (16x+8, 16y+8, 16z+8); its label is 2 + the index of the seed minimising
    sum_a (w_a * delta_a)^2 (int64), ties to the lower index; ``weights`` w are small integers proportional to
    the voxel size, e.g. (2, 2, 5) for voxelsize (0.2, 0.2, 0.5), so anisotropic voxels give anisotropic cells;
  * ``dome=True`` writes background label 1 outside the ellipsoid of semi-axes 0.47 * shape
    (integer test: sum_a u_a^2 * K_a > 2^40 with u_a = 2 x_a + 1 - dim_a, K_a = 2^40 // round(0.94 dim_a)^2).
"""
import numpy as np

CONFIGS = {
    # BASELINE.md section 3.  shape / voxelsize / weights are in (x, y, z) API order; the image is x-fastest
    # (Fortran order, as openalea's SpatialImage), so z is the slow axis and shards are z-slabs.
    "C1": dict(shape=(128, 128, 128), dtype="uint16", ncell=500, voxelsize=(1.0, 1.0, 1.0), weights=(1, 1, 1),
               dome=False, seed=0),
    "C2": dict(shape=(512, 512, 256), dtype="uint16", ncell=5000, voxelsize=(0.2, 0.2, 0.5), weights=(2, 2, 5),
               dome=False, seed=1),
    "C3": dict(shape=(1024, 1024, 1024), dtype="uint16", ncell=50000, voxelsize=(1.0, 1.0, 1.0), weights=(1, 1, 1),
               dome=True, seed=2),
    "C4": dict(shape=(2048, 2048, 1024), dtype="uint32", ncell=400000, voxelsize=(1.0, 1.0, 1.0), weights=(1, 1, 1),
               dome=True, seed=3),
}


def tissue_image(shape, ncell, seed, weights=(1, 1, 1), dome=False, dtype="uint16", voxelsize=(1.0, 1.0, 1.0),
                 backend="numpy"):
    """SpatialImage of API shape (x, y, z), x fastest in memory.  The generators below work on the C-contiguous
    (z, y, x) array, i.e. their "API order" is the reverse of this image's."""
    from .spatial_image import SpatialImage
    rshape, rw = tuple(shape[::-1]), tuple(weights[::-1])
    if backend == "numpy":
        zyx = voronoi_numpy(rshape, ncell, seed, rw, dome, np.dtype(dtype))
    else:
        zyx = voronoi_device(rshape, ncell, seed, rw, dome, dtype).cpu().numpy()
    return SpatialImage(zyx.transpose(2, 1, 0), voxelsize=voxelsize)


def voronoi_seeds(shape, ncell, seed):
    """int32[ncell, 3] fixed-point seed positions in API axis order."""
    rng = np.random.default_rng(seed)
    hi = np.asarray(shape, np.int64) * 16
    return np.stack([rng.integers(0, hi[a], size=ncell) for a in range(3)], axis=1).astype(np.int32)


def dome_mask(shape, zslice=None):
    """bool array (True = background) for API-ordered ``shape``; ``zslice`` = (lo, hi) restricts axis 0."""
    dims = np.asarray(shape, np.int64)
    D = np.floor(0.94 * dims + 0.5).astype(np.int64)
    D[D < 1] = 1
    K = (1 << 40) // (D * D)
    lo, hi = (0, shape[0]) if zslice is None else zslice
    u0 = 2 * np.arange(lo, hi, dtype=np.int64) + 1 - dims[0]
    u1 = 2 * np.arange(shape[1], dtype=np.int64) + 1 - dims[1]
    u2 = 2 * np.arange(shape[2], dtype=np.int64) + 1 - dims[2]
    q = (u0 * u0 * K[0])[:, None, None] + (u1 * u1 * K[1])[None, :, None] + (u2 * u2 * K[2])[None, None, :]
    return q > (1 << 40)


def voronoi_numpy(shape, ncell, seed, weights=(1, 1, 1), dome=False, dtype=np.uint16, k=8):
    """CPU generator (cKDTree candidates + exact integer re-ranking).  Small volumes only."""
    from scipy.spatial import cKDTree
    seeds = voronoi_seeds(shape, ncell, seed).astype(np.int64)
    w = np.asarray(weights, np.int64)
    tree = cKDTree((seeds * w).astype(np.float64))
    out = np.empty(shape, dtype)
    k = min(k, ncell)
    ys, zs = np.meshgrid(np.arange(shape[1], dtype=np.int64), np.arange(shape[2], dtype=np.int64), indexing="ij")
    for x in range(shape[0]):
        pts = np.stack([np.full(ys.size, 16 * x + 8, np.int64), 16 * ys.ravel() + 8, 16 * zs.ravel() + 8], axis=1) * w
        _, cand = tree.query(pts.astype(np.float64), k=k)
        cand = cand.reshape(len(pts), k)
        d = ((pts[:, None, :] - seeds[cand] * w) ** 2).sum(axis=2)
        best = d.min(axis=1, keepdims=True)
        idx = np.where(d == best, cand, np.iinfo(np.int64).max).min(axis=1)
        out[x] = (idx + 2).reshape(shape[1], shape[2]).astype(dtype)
    if dome:
        out[dome_mask(shape)] = 1
    return out


def voronoi_numpy_box(shape, ncell, seed, lo, hi, weights=(1, 1, 1), dome=False, dtype=np.uint16, k=8):
    """The box [lo, hi) (API axis order) of the volume ``voronoi_numpy(shape, ...)`` would make, without making the rest:
    what a CPU-only process needs of a 1024^3 workload (bench.py's reference arm)."""
    from scipy.spatial import cKDTree
    seeds = voronoi_seeds(shape, ncell, seed).astype(np.int64)
    w = np.asarray(weights, np.int64)
    tree = cKDTree((seeds * w).astype(np.float64))
    ext = tuple(int(h - l) for l, h in zip(lo, hi))
    out = np.empty(ext, dtype)
    k = min(k, ncell)
    ys, zs = np.meshgrid(np.arange(lo[1], hi[1], dtype=np.int64), np.arange(lo[2], hi[2], dtype=np.int64), indexing="ij")
    for i, x in enumerate(range(lo[0], hi[0])):
        pts = np.stack([np.full(ys.size, 16 * x + 8, np.int64), 16 * ys.ravel() + 8, 16 * zs.ravel() + 8], axis=1) * w
        _, cand = tree.query(pts.astype(np.float64), k=k)
        cand = cand.reshape(len(pts), k)
        d = ((pts[:, None, :] - seeds[cand] * w) ** 2).sum(axis=2)
        best = d.min(axis=1, keepdims=True)
        idx = np.where(d == best, cand, np.iinfo(np.int64).max).min(axis=1)
        out[i] = (idx + 2).reshape(ext[1], ext[2]).astype(dtype)
    if dome:
        out[dome_mask(shape, (lo[0], hi[0]))[:, lo[1]:hi[1], lo[2]:hi[2]]] = 1
    return out


def voronoi_device(shape, ncell, seed, weights=(1, 1, 1), dome=False, dtype="uint16", ctx=None, zslice=None,
                   device=None):
    """CUDA generator -> torch tensor (C-contiguous, API order = (slow, mid, fast)) on the current device.
    ``zslice`` = (lo, hi) generates only planes lo..hi-1 of axis 0 of the global volume (z-slab ranks)."""
    import torch
    from . import _native
    own_ctx = ctx is None
    if own_ctx:
        ctx = _native.Context(-1 if device is None else device)
    lo, hi = (0, shape[0]) if zslice is None else zslice
    elem = 2 if str(dtype) in ("uint16", "torch.uint16") else 4
    tdt = torch.uint16 if elem == 2 else torch.uint32
    out = torch.empty((hi - lo, shape[1], shape[2]), dtype=tdt, device="cuda" if device is None else device)
    seeds = voronoi_seeds(shape, ncell, seed)
    seeds_fms = np.ascontiguousarray(seeds[:, ::-1])          # (x,y,z) -> (fast, mid, slow)
    w_fms = np.asarray(weights, np.int32)[::-1].copy()
    torch.cuda.synchronize()
    ctx.synth_voronoi(out.data_ptr(), elem, shape[2], shape[1], hi - lo, lo, shape[0], seeds_fms, w_fms, dome)
    if own_ctx:
        ctx.close()
    return out

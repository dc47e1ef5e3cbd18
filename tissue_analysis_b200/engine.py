"""Host side of the scan: bind a label volume, run the single CUDA pass, hold the result tables.

``ScanTables`` is everything the feature extractors of the reference need
(reference: src/vplants/tissue_analysis/spatial_image_analysis.py, "SIA"):
per-label exact integer moments + bounding boxes (volume SIA:1231, center_of_mass SIA:466,
boundingbox SIA:517, inertia SIA:1261-1278) and per-(min,max)-pair directional face counts and
18-connected wall-voxel counts (neighbors SIA:45-60, cell_wall_area SIA:947-956, wall voxels
SIA:835-863).  Tables are stored in API axis order (x, y, z) whatever the memory order of the image.
"""
import numpy as np

from . import _native

_S2_INDEX = {(0, 0): 0, (0, 1): 1, (0, 2): 2, (1, 1): 3, (1, 2): 4, (2, 2): 5}
_MEM_PAIRS = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]   # ff fm fs mm ms ss


def memory_layout(image):
    """-> (view_smf, ax_of_mem): a C-contiguous (slow, mid, fast) view or copy of ``image`` and, for each
    memory axis k (0 fast, 1 mid, 2 slow), the API axis it corresponds to."""
    img = np.asarray(image)
    if img.ndim != 3:
        raise ValueError("a 3D label image is required, got shape %r" % (img.shape,))
    if img.dtype not in (np.uint16, np.uint32):
        if img.dtype.kind in "iu" and img.size and int(img.min()) >= 0:
            mx = int(img.max())
            img = img.astype(np.uint16 if mx < 65536 else np.uint32)
        else:
            raise ValueError("label image must be an unsigned 16/32-bit integer array, got %s" % img.dtype)
    order = sorted(range(3), key=lambda a: (-abs(img.strides[a]), a))      # slow -> fast API axes
    view = np.transpose(img, order)
    if not view.flags["C_CONTIGUOUS"]:
        order = [0, 1, 2]
        view = np.ascontiguousarray(img)
    ax_of_mem = (order[2], order[1], order[0])
    return view, ax_of_mem


class ScanTables(object):
    """Result of one pass, API axis order."""

    def __init__(self, shape, count, s1, s2, bmin, bmax, pair_lo, pair_hi, faces, wall18):
        self.shape = tuple(int(v) for v in shape)
        self.count = count            # int64[L]
        self.s1 = s1                  # int64[L,3]
        self.s2 = s2                  # int64[L,6]  xx xy xz yy yz zz
        self.bmin = bmin              # int64[L,3]
        self.bmax = bmax              # int64[L,3] inclusive
        self.pair_lo = pair_lo        # int64[P] sorted by (lo, hi)
        self.pair_hi = pair_hi
        self.faces = faces            # int64[P,6]
        self.wall18 = wall18          # int64[P]
        self._key = None

    @property
    def nrows(self):
        return self.count.shape[0]

    def pair_keys(self):
        if self._key is None:
            self._key = (self.pair_lo.astype(np.uint64) << np.uint64(32)) | self.pair_hi.astype(np.uint64)
        return self._key

    def find_pairs(self, a, b):
        """Row index of each (a_i, b_i) pair (any order), -1 when the labels never touch."""
        a = np.asarray(a, np.int64)
        b = np.asarray(b, np.int64)
        lo = np.minimum(a, b).astype(np.uint64)
        hi = np.maximum(a, b).astype(np.uint64)
        key = (lo << np.uint64(32)) | hi
        keys = self.pair_keys()
        if keys.size == 0:
            return np.full(key.shape, -1, np.int64)
        pos = np.searchsorted(keys, key)
        pos_c = np.minimum(pos, keys.size - 1)
        ok = (keys[pos_c] == key) & (a >= 0) & (b >= 0)
        return np.where(ok, pos_c, -1).astype(np.int64)


def tables_from_memory_order(shape_api, ax_of_mem, count, s1, s2, bbox, lo, hi, faces, wall):
    """Permute native (fast, mid, slow) tables into API (x, y, z) order."""
    n = count.shape[0]
    s1_api = np.empty((n, 3), np.int64)
    bmin = np.empty((n, 3), np.int64)
    bmax = np.empty((n, 3), np.int64)
    for k in range(3):
        a = ax_of_mem[k]
        s1_api[:, a] = s1[:, k].astype(np.int64)
        bmin[:, a] = bbox[:, k]
        bmax[:, a] = bbox[:, 3 + k]
    s2_api = np.empty((n, 6), np.int64)
    for j, (k1, k2) in enumerate(_MEM_PAIRS):
        a, b = sorted((ax_of_mem[k1], ax_of_mem[k2]))
        s2_api[:, _S2_INDEX[(a, b)]] = s2[:, j].astype(np.int64)
    faces_api = np.empty((faces.shape[0], 6), np.int64)
    for k in range(3):
        a = ax_of_mem[k]
        faces_api[:, 2 * a] = faces[:, 2 * k]
        faces_api[:, 2 * a + 1] = faces[:, 2 * k + 1]
    return ScanTables(shape_api, count.astype(np.int64), s1_api, s2_api, bmin, bmax,
                      lo.astype(np.int64), hi.astype(np.int64), faces_api, wall.astype(np.int64))


class VolumeScan(object):
    """One bound volume on one GPU: runs the pass and serves the optional second passes."""

    def __init__(self, image, device=-1, flags=_native.PASS_ALL, max_label_hint=0, pair_capacity_hint=0):
        self.view, self.ax_of_mem = memory_layout(image)
        self._image_ref = np.asarray(image)
        self.shape_api = tuple(np.asarray(image).shape)
        self.ctx = _native.Context(device)
        self._bound = False              # the first pass copies and scans in one overlapped call (ta_run_pass_host)
        self.flags = flags
        self.max_label_hint = max_label_hint
        self.pair_capacity_hint = pair_capacity_hint
        self.tables = None

    def _bind(self):
        if not self._bound:
            self.ctx.bind_host(self.view)
            self._bound = True

    def run(self):
        if self._bound:
            self.ctx.run_pass(self.flags, self.max_label_hint, self.pair_capacity_hint)
        else:
            self.ctx.run_pass_host(self.view, self.flags, self.max_label_hint, self.pair_capacity_hint)
            self._bound = True
        count, s1, s2, bbox = self.ctx.label_table()
        lo, hi, faces, wall = self.ctx.pair_table()
        self.tables = tables_from_memory_order(self.shape_api, self.ax_of_mem, count, s1, s2, bbox, lo, hi, faces,
                                               wall)
        return self.tables

    # ---- derived, still on the device -----------------------------------------------------------------
    def inertia(self, labels):
        """-> evals[n,3] descending, evecs[n,3,3] rows, components in API axis order."""
        evals, evecs_mem = self.ctx.inertia_from_moments(np.asarray(labels, np.uint32))
        evecs = np.empty_like(evecs_mem)
        for k in range(3):
            evecs[:, :, self.ax_of_mem[k]] = evecs_mem[:, :, k]
        return evals, evecs

    def wall_voxel_coords(self, lo, hi):
        """-> list of int64[3, n_i] in API axis order, each sorted lexicographically by (x, y, z)."""
        self._bind()
        _, blocks = self.ctx.wall_voxel_coords(lo, hi)
        out = []
        for blk in blocks:
            xyz = np.empty_like(blk)
            for k in range(3):
                xyz[self.ax_of_mem[k]] = blk[k]
            if self.ax_of_mem != (2, 1, 0) and xyz.shape[1] > 1:   # memory order is not API C-order
                order = np.lexsort((xyz[2], xyz[1], xyz[0]))
                xyz = xyz[:, order]
            out.append(xyz)
        return out

    def _to_api(self, out_smf):
        inv = np.argsort([self.ax_of_mem[2], self.ax_of_mem[1], self.ax_of_mem[0]])
        return np.transpose(out_smf, inv)

    def stencil_image(self, kind):
        """'hollow': labels where the Laplacian is non-zero (SIA:74-94); 'shell18': 0/1 outer shell (SIA:1399-1448)."""
        self._bind()
        return self._to_api(self.ctx.stencil_image(kind, self.view.shape, self.view.dtype))

    def map_labels(self, lut, fill=0):
        """API-ordered image ``lut[image]`` (device gather; labels beyond the table map to ``fill``)."""
        self._bind()
        return self._to_api(self.ctx.map_labels(lut, fill, self.view.shape))

    def relabel(self, mapping):
        """Rewrite labels in place (device copy AND the caller's host image) and invalidate the tables.
        ``mapping``: {old label: new label}; other labels are kept."""
        dmax = int(np.iinfo(self.view.dtype).max)
        bad = [v for v in mapping.values() if not 0 <= int(v) <= dmax]
        if bad:
            raise ValueError("relabel: new label %r does not fit the image's %s voxels" % (bad[0], self.view.dtype))
        top = max(int(np.iinfo(self.view.dtype).max) + 1 if self.view.dtype == np.uint16 else 0,
                  max(mapping) + 1 if mapping else 0,
                  self.tables.nrows if self.tables is not None else int(self.view.max()) + 1)
        lut = np.arange(top, dtype=np.int64)
        for old, new in mapping.items():
            lut[old] = new
        self._bind()
        out = self.ctx.map_labels(lut.astype(self.view.dtype), 0, self.view.shape, in_place=True)
        if np.shares_memory(self.view, self._image_ref):
            np.copyto(self.view, out)                      # the view aliases the caller's image memory
        else:
            np.copyto(self.view, out)
            self._image_ref[...] = self._to_api(out)
        self.tables = None

    def voxel_first_layer(self, background, keep_background=True):
        ns, nm, nf = self.view.shape
        self._bind()
        return self._to_api(self.ctx.voxel_first_layer(background, keep_background, (ns, nm, nf), self.view.dtype))


def scan_volume(image, **kw):
    vs = VolumeScan(image, **kw)
    vs.run()
    return vs

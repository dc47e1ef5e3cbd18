"""ctypes binding of libtissue_b200.so (C ABI declared in include/tissue_b200.h).

There is no CPU fallback: if the library is missing, or no B200-class device is present,
every compute entry raises ``NativeError`` loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TA_LIB_PATH: a differently built copy of the same library (kernel tuning experiments); the default is the in-tree build
LIB_PATH = os.environ.get("TA_LIB_PATH") or os.path.join(_HERE, "libtissue_b200.so")

TA_OK = 0
TA_ERR_CUDA = -1
TA_ERR_BAD_ARG = -2
TA_ERR_PAIR_OVERFLOW = -3
TA_ERR_LABEL_RANGE = -4
TA_ERR_NO_VOLUME = -5
TA_ERR_NO_TABLES = -6

PASS_MOMENTS, PASS_PAIRS6, PASS_WALL18, PASS_ALL = 1, 2, 4, 7
PASS_UNSORTED = 0x2000   # pair records stay in hash order (input of merge_pair_records only)
PASS_DEFERRED = 0x4000   # no host synchronisation inside the pass; flags and counts are read at the first fetch

# every symbol include/tissue_b200.h declares (tests/test_cabi_symbols.py checks header <-> library)
EXPORTS = [
    "ta_version", "ta_ctx_create", "ta_ctx_destroy", "ta_last_error", "ta_set_stream", "ta_bind_volume",
    "ta_set_slab", "ta_run_pass", "ta_run_pass_host", "ta_run_pass_ranges", "ta_label_table_size", "ta_fetch_label_table", "ta_pair_table_size",
    "ta_fetch_pair_table", "ta_label_table_device", "ta_pair_records_device", "ta_merge_pair_records",
    "ta_pair_records_deferred", "ta_merge_pair_records_deferred",
    "ta_inertia_from_moments", "ta_inertia_table", "ta_inertia_eig", "ta_wall_voxel_coords", "ta_voxel_first_layer",
    "ta_hollow_out_cells", "ta_cell_shell18", "ta_map_labels", "ta_last_timing", "ta_launch_count", "ta_synth_voronoi",
]


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libtissue_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def load():
    """Load the shared library (built in-tree by ``__graft_entry__.build()`` / csrc/build.sh)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(TA_ERR_CUDA, "%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                       "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i64, u64, u32, ci = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int
    P = C.POINTER
    lib.ta_version.restype = C.c_char_p
    lib.ta_last_error.restype = C.c_char_p
    lib.ta_last_error.argtypes = [vp]
    lib.ta_ctx_create.argtypes = [P(vp), ci]
    lib.ta_ctx_destroy.argtypes = [vp]
    lib.ta_set_stream.argtypes = [vp, vp]
    lib.ta_bind_volume.argtypes = [vp, vp, ci, ci, i64, i64, i64]
    lib.ta_set_slab.argtypes = [vp, i64, i64, i64]
    lib.ta_run_pass.argtypes = [vp, u32, u32, u64]
    lib.ta_run_pass_ranges.argtypes = [vp, u32, u32, u64, ci, P(i64), P(vp)]
    lib.ta_run_pass_host.argtypes = [vp, vp, ci, i64, i64, i64, P(i64), u32, u32, u64, i64]
    lib.ta_label_table_size.argtypes = [vp, P(u64)]
    lib.ta_fetch_label_table.argtypes = [vp, vp, vp, vp, vp]
    lib.ta_pair_table_size.argtypes = [vp, P(u64)]
    lib.ta_fetch_pair_table.argtypes = [vp, vp, vp, vp, vp]
    lib.ta_label_table_device.argtypes = [vp, P(vp), P(vp), P(vp), P(vp), P(vp), P(u64)]
    lib.ta_pair_records_device.argtypes = [vp, P(vp), P(u64)]
    lib.ta_merge_pair_records.argtypes = [vp, vp, u64]
    lib.ta_pair_records_deferred.argtypes = [vp, P(vp), P(u64)]
    lib.ta_merge_pair_records_deferred.argtypes = [vp, vp, u64, ci]
    lib.ta_inertia_from_moments.argtypes = [vp, vp, u64, vp, vp]
    lib.ta_inertia_table.argtypes = [vp, vp, vp]
    lib.ta_inertia_eig.argtypes = [vp, vp, u64, vp, vp]
    lib.ta_wall_voxel_coords.argtypes = [vp, vp, vp, u64, vp, vp]
    lib.ta_voxel_first_layer.argtypes = [vp, u32, ci, vp]
    lib.ta_hollow_out_cells.argtypes = [vp, ci, vp]
    lib.ta_cell_shell18.argtypes = [vp, vp]
    lib.ta_map_labels.argtypes = [vp, vp, ci, u64, u32, vp, ci]
    lib.ta_last_timing.argtypes = [vp, P(C.c_float), P(C.c_float), P(C.c_float)]
    lib.ta_launch_count.argtypes = [vp, P(u64)]
    lib.ta_synth_voronoi.argtypes = [vp, vp, ci, i64, i64, i64, i64, i64, vp, u32, vp, ci]
    for name in EXPORTS:
        if name not in ("ta_version", "ta_last_error"):
            getattr(lib, name).restype = ci
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context(object):
    """Owns one ``ta_ctx`` (device buffers, stream, tables) on one GPU."""

    def __init__(self, device=-1):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.ta_ctx_create(C.byref(h), int(device))
        if rc != TA_OK:
            raise NativeError(rc, self.lib.ta_last_error(None).decode())
        self.h = h
        self._keepalive = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.ta_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != TA_OK:
            raise NativeError(rc, self.lib.ta_last_error(self.h).decode())

    # ---- volume -------------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._check(self.lib.ta_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def bind_host(self, arr):
        """arr: C-contiguous numpy array (slow, mid, fast) of uint16/uint32."""
        assert arr.flags["C_CONTIGUOUS"] and arr.ndim == 3
        ns, nm, nf = arr.shape
        self._check(self.lib.ta_bind_volume(self.h, _ptr(arr), 0, arr.dtype.itemsize, nf, nm, ns))
        self._keepalive = None

    def bind_device(self, dev_ptr, elem_bytes, nf, nm, ns, keepalive=None):
        self._check(self.lib.ta_bind_volume(self.h, C.c_void_p(dev_ptr), 1, elem_bytes, nf, nm, ns))
        self._keepalive = keepalive

    def set_slab(self, own_lo, own_hi, slow_offset):
        self._check(self.lib.ta_set_slab(self.h, own_lo, own_hi, slow_offset))

    # ---- pass + tables ---------------------------------------------------------------------------------
    def run_pass(self, flags=PASS_ALL, max_label_hint=0, pair_capacity_hint=0, max_retries=4):
        cap = int(pair_capacity_hint)
        for attempt in range(max_retries + 1):
            rc = self.lib.ta_run_pass(self.h, flags, int(max_label_hint), cap)
            if rc == TA_ERR_PAIR_OVERFLOW and attempt < max_retries:
                cap = max(cap * 4, 1 << 20) if cap else 1 << 22   # never drop pairs: grow and redo the pass
                continue
            self._check(rc)
            return

    def run_pass_ranges(self, ranges, events, flags=PASS_ALL, max_label_hint=0, pair_capacity_hint=0, max_retries=4):
        """ranges: [(lo, hi), ...] tiling the owned planes; events: cudaEvent_t handles (ints) or None per range."""
        n = len(ranges)
        flat = (C.c_int64 * (2 * n))(*[int(v) for r in ranges for v in r])
        evs = (C.c_void_p * n)(*[C.c_void_p(int(e)) if e else None for e in events])
        cap = int(pair_capacity_hint)
        for attempt in range(max_retries + 1):
            rc = self.lib.ta_run_pass_ranges(self.h, flags, int(max_label_hint), cap, n, flat, evs)
            if rc == TA_ERR_PAIR_OVERFLOW and attempt < max_retries:
                cap = max(cap * 4, 1 << 20) if cap else 1 << 22
                continue
            self._check(rc)
            return

    def run_pass_host(self, arr, flags=PASS_ALL, max_label_hint=0, pair_capacity_hint=0, chunk_planes=0,
                      max_retries=4, slab=None):
        """bind_host(arr) [+ set_slab(*slab)] + run_pass() with the H2D copy and the scan overlapped."""
        assert arr.flags["C_CONTIGUOUS"] and arr.ndim == 3
        ns, nm, nf = arr.shape
        self._keepalive = None
        slab_c = (C.c_int64 * 3)(*[int(v) for v in slab]) if slab is not None else None
        rc = self.lib.ta_run_pass_host(self.h, _ptr(arr), arr.dtype.itemsize, nf, nm, ns, slab_c, flags,
                                       int(max_label_hint), int(pair_capacity_hint), int(chunk_planes))
        if rc == TA_ERR_PAIR_OVERFLOW and max_retries > 0:
            # the volume is resident by now: grow the table and redo the pass on the device copy
            cap = int(pair_capacity_hint)
            return self.run_pass(flags, max_label_hint, max(cap * 4, 1 << 20) if cap else 1 << 22, max_retries - 1)
        self._check(rc)

    def label_table(self):
        n = C.c_uint64()
        self._check(self.lib.ta_label_table_size(self.h, C.byref(n)))
        n = n.value
        count = np.empty(n, np.uint64)
        s1 = np.empty((n, 3), np.uint64)
        s2 = np.empty((n, 6), np.uint64)
        bbox = np.empty((n, 6), np.int32)
        self._check(self.lib.ta_fetch_label_table(self.h, _ptr(count), _ptr(s1), _ptr(s2), _ptr(bbox)))
        return count, s1, s2, bbox

    def pair_table(self):
        n = C.c_uint64()
        self._check(self.lib.ta_pair_table_size(self.h, C.byref(n)))
        n = n.value
        lo = np.empty(n, np.uint32)
        hi = np.empty(n, np.uint32)
        faces = np.empty((n, 6), np.uint32)
        wall = np.empty(n, np.uint32)
        self._check(self.lib.ta_fetch_pair_table(self.h, _ptr(lo), _ptr(hi), _ptr(faces), _ptr(wall)))
        return lo, hi, faces, wall

    def label_table_device(self):
        p = [C.c_void_p() for _ in range(5)]
        n = C.c_uint64()
        self._check(self.lib.ta_label_table_device(self.h, *[C.byref(x) for x in p], C.byref(n)))
        return [x.value for x in p], n.value

    def pair_records_device(self):
        p = C.c_void_p()
        n = C.c_uint64()
        self._check(self.lib.ta_pair_records_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def merge_pair_records(self, dev_ptr, n):
        self._check(self.lib.ta_merge_pair_records(self.h, C.c_void_p(dev_ptr or 0), n))

    def pair_records_deferred(self):
        """(device pointer, cap_rows) of the record buffer of a PASS_DEFERRED pass: uint32[1 + cap_rows][9]."""
        p = C.c_void_p()
        n = C.c_uint64()
        self._check(self.lib.ta_pair_records_deferred(self.h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def merge_pair_records_deferred(self, dev_ptr, cap_rows, world):
        self._check(self.lib.ta_merge_pair_records_deferred(self.h, C.c_void_p(dev_ptr), C.c_uint64(cap_rows), int(world)))

    # ---- derived ---------------------------------------------------------------------------------------------
    def inertia_from_moments(self, labels=None, n=None):
        if labels is not None:
            labels = np.ascontiguousarray(labels, np.uint32)
            n = labels.size
        evals = np.empty((n, 3), np.float64)
        evecs = np.empty((n, 3, 3), np.float64)
        self._check(self.lib.ta_inertia_from_moments(self.h, _ptr(labels), n, _ptr(evals), _ptr(evecs)))
        return evals, evecs

    def inertia_table(self, fetch=True, nrows=None):
        """Eigen-solve every table row on the device; ``fetch=False`` leaves the result resident."""
        if not fetch:
            self._check(self.lib.ta_inertia_table(self.h, None, None))
            return None
        evals = np.empty((nrows, 3), np.float64)
        evecs = np.empty((nrows, 3, 3), np.float64)
        self._check(self.lib.ta_inertia_table(self.h, _ptr(evals), _ptr(evecs)))
        return evals, evecs

    def inertia_eig(self, cov6):
        cov6 = np.ascontiguousarray(cov6, np.float64).reshape(-1, 6)
        n = cov6.shape[0]
        evals = np.empty((n, 3), np.float64)
        evecs = np.empty((n, 3, 3), np.float64)
        self._check(self.lib.ta_inertia_eig(self.h, _ptr(cov6), n, _ptr(evals), _ptr(evecs)))
        return evals, evecs

    def wall_voxel_coords(self, lo, hi):
        """-> counts[npairs], list of int64[3, n_i] arrays in memory-axis order (fast, mid, slow)."""
        lo = np.ascontiguousarray(lo, np.uint32)
        hi = np.ascontiguousarray(hi, np.uint32)
        n = lo.size
        counts = np.zeros(n, np.uint64)
        if n == 0:
            return counts, []
        self._check(self.lib.ta_wall_voxel_coords(self.h, _ptr(lo), _ptr(hi), n, _ptr(counts), None))
        total = int(counts.sum())
        xyz = np.empty(3 * max(total, 1), np.int64)
        self._check(self.lib.ta_wall_voxel_coords(self.h, _ptr(lo), _ptr(hi), n, _ptr(counts), _ptr(xyz)))
        out, off = [], 0
        for c in counts.tolist():
            out.append(xyz[3 * off:3 * (off + c)].reshape(3, c))
            off += c
        return counts, out

    def voxel_first_layer(self, background, keep_background, shape_smf, dtype):
        out = np.empty(shape_smf, dtype)
        self._check(self.lib.ta_voxel_first_layer(self.h, int(background), int(bool(keep_background)), _ptr(out)))
        return out

    def stencil_image(self, kind, shape_smf, dtype):
        """kind 'hollow' (labels where the Laplacian is non-zero), 'laplace' (the same as a 0/1 mask) or 'shell18'
        (0/1 outer shell of every cell)."""
        out = np.empty(shape_smf, dtype)
        if kind == "shell18":
            self._check(self.lib.ta_cell_shell18(self.h, _ptr(out)))
        else:
            self._check(self.lib.ta_hollow_out_cells(self.h, 1 if kind == "laplace" else 0, _ptr(out)))
        return out

    def map_labels(self, lut, fill, shape_smf, in_place=False, fetch=True):
        """out[p] = lut[volume[p]] on the device; lut: 1-D uint16/uint32 array indexed by label."""
        lut = np.ascontiguousarray(lut)
        assert lut.dtype in (np.uint16, np.uint32) and lut.ndim == 1
        out = np.empty(shape_smf, lut.dtype) if fetch else None
        self._check(self.lib.ta_map_labels(self.h, _ptr(lut), lut.dtype.itemsize, lut.size, int(fill), _ptr(out),
                                           int(bool(in_place))))
        return out

    def last_timing(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._check(self.lib.ta_last_timing(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(scan_ms=a.value, pass_ms=b.value, h2d_ms=c.value)

    def launch_count(self):
        n = C.c_uint64()
        self._check(self.lib.ta_launch_count(self.h, C.byref(n)))
        return n.value

    def synth_voronoi(self, dev_ptr, elem_bytes, nf, nm, ns, slow_offset, global_slow, seeds_fms, weight_fms, dome):
        seeds = np.ascontiguousarray(seeds_fms, np.int32)
        w = np.ascontiguousarray(weight_fms, np.int32)
        self._check(self.lib.ta_synth_voronoi(self.h, C.c_void_p(dev_ptr), elem_bytes, nf, nm, ns, slow_offset,
                                              global_slow, _ptr(seeds), seeds.shape[0], _ptr(w), int(bool(dome))))

"""ctypes loader of the C oracle (oracle/c/onepass.c -> oracle/_build/libta_oracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libta_oracle.so")


def _lib():
    if not os.path.exists(_SO):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    lib = C.CDLL(_SO)
    lib.ta_oracle_onepass.restype = C.c_int64
    lib.ta_oracle_onepass.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_uint32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_uint32))]
    lib.ta_oracle_free.argtypes = [C.c_void_p]
    return lib


def onepass(vol_smf, nrows=None):
    """vol_smf: C-contiguous (slow, mid, fast) uint16/uint32 array -> memory-order tables like the C ABI's."""
    vol = np.ascontiguousarray(vol_smf)
    ns, nm, nf = vol.shape
    n = int(vol.max()) + 1 if nrows is None else int(nrows)
    count = np.empty(n, np.uint64)
    s1 = np.empty((n, 3), np.uint64)
    s2 = np.empty((n, 6), np.uint64)
    bbox = np.empty((n, 6), np.int32)
    out = C.POINTER(C.c_uint32)()
    lib = _lib()
    k = lib.ta_oracle_onepass(vol.ctypes.data, vol.dtype.itemsize, nf, nm, ns, n, count.ctypes.data, s1.ctypes.data,
                              s2.ctypes.data, bbox.ctypes.data, C.byref(out))
    if k < 0:
        raise RuntimeError("C oracle failed (allocation or label out of range)")
    rec = np.ctypeslib.as_array(out, shape=(max(k, 1), 9))[:k].copy()
    lib.ta_oracle_free(out)
    return dict(count=count, s1=s1, s2=s2, bbox=bbox, lo=rec[:, 0], hi=rec[:, 1], faces=rec[:, 2:8], wall18=rec[:, 8])

"""The C-ABI library loads and exports exactly the entry points include/tissue_b200.h declares (no compute)."""
import ctypes
import os
import re

from tissue_analysis_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "tissue_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ta_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_functions() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_native.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), name
    lib.ta_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.ta_version()


def test_no_silent_cpu_fallback():
    """Without a CUDA device the context must refuse to exist (the product never routes to the oracle)."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(_native.NativeError):
        _native.Context()
    import numpy as np
    from tissue_analysis_b200 import SpatialImageAnalysis3D
    sia = SpatialImageAnalysis3D(np.ones((4, 4, 4), np.uint16), background=1)
    with pytest.raises(_native.NativeError):
        sia.volume()

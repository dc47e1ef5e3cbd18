"""Golden fixtures (tests/golden/, written by tests/golden/make_golden.py).

reference_docstring_vectors.json holds the only results the reference itself pins (its docstring examples, transcribed with
their line numbers); tables_*.npz freeze the CPU oracle's tables for three small seeded tissues.  CPU: both oracles and
the host mirror reproduce them.  `-m gpu`: the CUDA tables and the class API reproduce them, without the oracle in the loop.
"""
import glob
import json
import os
import warnings

import numpy as np
import pytest

from tissue_analysis_b200 import SpatialImage, SpatialImageAnalysis3D

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "tables_*.npz")))
warnings.filterwarnings("ignore", category=UserWarning)


def load(path):
    z = np.load(path)
    return z, SpatialImage(z["image"], voxelsize=tuple(z["voxelsize"]))


def assert_tables_match_fixture(t, z):
    n = z["count"].shape[0]
    assert not np.asarray(t.count[n:]).any()
    assert np.array_equal(t.count[:n], z["count"])
    assert np.array_equal(t.s1[:n], z["s1"]) and np.array_equal(t.s2[:n], z["s2"])
    present = z["count"] > 0
    assert np.array_equal(np.asarray(t.bmin[:n])[present], z["bmin"][present])
    assert np.array_equal(np.asarray(t.bmax[:n])[present], z["bmax"][present])
    assert np.array_equal(t.pair_lo, z["pair_lo"]) and np.array_equal(t.pair_hi, z["pair_hi"])
    assert np.array_equal(t.faces, z["faces"]) and np.array_equal(t.wall18, z["wall18"])


def check_docstring_vectors(make):
    """`make(image)` -> an analysis object with the reference's method surface."""
    d = json.load(open(os.path.join(GOLDEN, "reference_docstring_vectors.json")))
    toy = np.array(d["image_4x6"], dtype=np.uint16).reshape(4, 6, 1)
    sia = make(toy)
    assert sorted(sia.labels()) == d["labels"]["value"]
    com = sia.center_of_mass()
    assert {str(k): list(v) for k, v in com.items()} == d["center_of_mass"]["value"]
    bb = sia.boundingbox()
    for k, box in d["boundingbox"]["value"].items():
        assert bb[int(k)] == tuple(slice(a, b) for a, b in box)
    nb = sia.neighbors(verbose=False) if "verbose" in sia.neighbors.__code__.co_varnames else sia.neighbors()
    assert {str(k): sorted(int(x) for x in v) for k, v in nb.items()} == d["neighbors"]["value"]
    assert {"%d,%d" % k: v for k, v in sia.cell_wall_area(7, [2, 5]).items()} == d["cell_wall_area_7"]["value"]
    assert {"%d,%d" % k: float(v) for k, v in sia.wall_areas().items()} == d["wall_areas"]["value"]
    assert {str(k): float(v) for k, v in sia.volume().items()} == d["volume"]["value"]


# ---------------------------------------------------------------------------------------------------- CPU
def test_fixture_files_exist():
    assert len(CASES) == 3 and os.path.exists(os.path.join(GOLDEN, "reference_docstring_vectors.json"))


def test_loop_oracle_reproduces_the_reference_docstrings():
    from oracle.sia_loops import LoopOracle
    check_docstring_vectors(lambda im: LoopOracle(im))


def test_host_mirror_reproduces_the_reference_docstrings():
    from tests.helpers import OracleBackend
    check_docstring_vectors(lambda im: SpatialImageAnalysis3D(im, _backend=OracleBackend(im)))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_oracles_reproduce_the_fixtures(path):
    """numpy one-pass oracle and C oracle == the frozen tables (scipy / numpy drift shows up here)."""
    from tests.helpers import oracle_tables
    z, img = load(path)
    assert_tables_match_fixture(oracle_tables(np.asarray(img)), z)
    from oracle import c_onepass
    n = z["count"].shape[0]
    c = c_onepass.onepass(np.ascontiguousarray(np.asarray(img).transpose(2, 1, 0)), nrows=n)   # memory axes f, m, s = x, y, z
    assert np.array_equal(c["count"], z["count"]) and np.array_equal(c["s1"], z["s1"]) and np.array_equal(c["s2"], z["s2"])
    assert np.array_equal(c["lo"], z["pair_lo"]) and np.array_equal(c["hi"], z["pair_hi"])
    assert np.array_equal(c["faces"], z["faces"]) and np.array_equal(c["wall18"], z["wall18"])


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_cuda_tables_reproduce_the_fixtures(path):
    z, img = load(path)
    assert_tables_match_fixture(SpatialImageAnalysis3D(img, background=1)._tables(), z)


@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference_docstrings():
    check_docstring_vectors(lambda im: SpatialImageAnalysis3D(im))

"""Host mirror (tissue_analysis_b200.SpatialImageAnalysis3D) fed by oracle tables == the loop oracle.

Proves two things on the CPU: (1) the one-pass table definitions of oracle/sia_onepass.py reproduce the
per-label loops of oracle/sia_loops.py (the line-for-line restatement of the reference), and (2) the product's
host-side logic (tables -> the reference's dicts / lists / floats) is exact.  The CUDA scan itself is compared
with the same oracles in the `-m gpu` tests.
"""
import warnings

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle.sia_loops import LoopOracle, LIST, NPLIST
from tests.helpers import TOY, OracleBackend, compare_api
from tissue_analysis_b200 import SpatialImage, SpatialImageAnalysis3D
from tissue_analysis_b200.synth import tissue_image

warnings.filterwarnings("ignore", category=UserWarning)
warnings.filterwarnings("ignore", category=RuntimeWarning)


def both(img, **kw):
    prod = SpatialImageAnalysis3D(img, _backend=OracleBackend(img), **kw)
    orc = LoopOracle(np.asarray(img), voxelsize=getattr(img, "voxelsize", None), **kw)
    return prod, orc


def test_docstring_image_through_product_class():
    prod, orc = both(TOY.copy())
    compare_api(prod, orc, eig=False)
    assert list(prod.center_of_mass(7)) == [0.75, 2.75, 0.0]
    assert prod.boundingbox(7) == (slice(0, 3), slice(2, 4), slice(0, 1))
    assert prod.cell_wall_area(7, [2, 5]) == {(2, 7): 1.0, (5, 7): 2.0}
    assert prod.volume(7) == {7: 4.0}


def test_voronoi_dome_isotropic():
    img = tissue_image((40, 36, 30), 40, seed=11, dome=True)
    prod, orc = both(img, background=1)
    compare_api(prod, orc)


def test_voronoi_anisotropic_ignored_labels():
    img = tissue_image((36, 30, 24), 30, seed=5, weights=(2, 2, 5), dome=True, voxelsize=(0.2, 0.2, 0.5))
    prod, orc = both(img, background=1, ignoredlabels=[0, 3, 4])
    compare_api(prod, orc)


def test_c_order_image_and_label_zero():
    rng = np.random.default_rng(3)
    img = tissue_image((24, 20, 28), 18, seed=9, dome=True)
    arr = np.ascontiguousarray(np.asarray(img))
    arr[rng.random(arr.shape) < 0.02] = 0          # sprinkle "removed cell" voxels
    im = SpatialImage(arr, voxelsize=(0.5, 1.0, 2.0))
    prod, orc = both(im, background=1, ignoredlabels=0)
    compare_api(prod, orc)


@settings(max_examples=12, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(2, 9), st.tuples(st.integers(3, 9), st.integers(3, 9), st.integers(3, 9)))
def test_random_noise_volumes(seed, nlab, shape):
    """Worst case for the pair logic: every voxel is a wall voxel with many distinct neighbours."""
    rng = np.random.default_rng(seed)
    arr = rng.integers(1, nlab + 1, size=shape).astype(np.uint16)
    arr[0, 0, 0] = 1
    prod, orc = both(SpatialImage(arr, voxelsize=(1.0, 0.7, 1.3)), background=1)
    compare_api(prod, orc, eig=False)


def test_single_label_and_return_types():
    arr = np.full((5, 6, 7), 4, np.uint16)
    prod, orc = both(arr)
    assert prod.labels() == orc.labels() == [4]
    assert prod.volume() == orc.volume()
    assert np.array_equal(prod.center_of_mass(), orc.center_of_mass())
    assert prod.neighbors() == orc.neighbors()
    assert prod.wall_areas() == orc.wall_areas() == {}
    img = tissue_image((20, 18, 16), 9, seed=2, dome=True)
    for rt in (LIST, NPLIST):
        prod, orc = both(img, background=1, return_type=rt)
        vp, vo = prod.volume(), orc.volume()
        assert list(vp) == list(vo)
        assert sorted(prod.labels()) == sorted(orc.labels())


def test_requests_for_absent_labels():
    img = tissue_image((20, 18, 16), 9, seed=2, dome=True)
    prod, orc = both(img, background=1)
    assert prod.boundingbox(4000) is None and orc.boundingbox(4000) is None
    assert prod.neighbors(4000) == orc.neighbors(4000) == []
    assert prod.volume([2, 4000]) == orc.volume([2, 4000])
    assert np.isnan(np.asarray(prod.center_of_mass(4000), float)).all()
    with pytest.raises(ValueError):
        SpatialImageAnalysis3D(img, background=1.5, _backend=OracleBackend(img))
    with pytest.raises(ValueError):
        prod.label_request(3.2)


def test_wall_areas_fast_path_equals_the_per_label_loop():
    """wall_areas answers the usual input with one table lookup; the floats must be those of the reference's per-label
    loop (SIA:986-992), for real and voxel units, anisotropic voxels, subsets, lists with repeated labels (loop path)
    and neighbours that do not touch."""
    from tissue_analysis_b200 import SpatialImage
    img = SpatialImage(np.asarray(tissue_image((30, 26, 22), 25, seed=9, dome=True)), voxelsize=(0.21, 0.37, 0.53))
    sia = SpatialImageAnalysis3D(img, background=1, _backend=OracleBackend(img))

    def loop(neighbors, real):
        areas = {}
        for label_id, lneighbors in neighbors.items():
            neigh = [n for n in lneighbors if n > label_id]
            if neigh:
                for key, val in sia.cell_wall_area(label_id, neigh, real=real).items():
                    areas[key] = areas.get(key, 0.0) + val
        return areas

    full = sia.neighbors(verbose=False)
    some = dict(list(full.items())[3:9])
    far = {2: [3, 4, 5, 26, 27], 5: [9, 2]}                                   # mostly not touching
    twice = {k: v + v[:1] for k, v in some.items()}                            # repeated labels: the loop path
    for nb in (full, some, far, twice):
        for real in (True, False):
            got, want = sia.wall_areas(nb, real=real), loop(nb, real)
            assert list(got.keys()) == list(want.keys())
            assert all(got[k] == want[k] and type(got[k]) is type(want[k]) for k in want)

"""The plain-C one-pass oracle (oracle/c/onepass.c) equals the numpy one-pass oracle on random and Voronoi volumes."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_onepass, sia_onepass
from tissue_analysis_b200.synth import voronoi_numpy


def _same(vol_zyx):
    c = c_onepass.onepass(vol_zyx)
    img = vol_zyx.transpose(2, 1, 0)            # API (x, y, z) == memory (fast, mid, slow)
    lt, pt = sia_onepass.label_table(img), sia_onepass.pair_table(img)
    n = lt["count"].size
    assert np.array_equal(c["count"][:n].astype(np.int64), lt["count"])
    assert np.array_equal(c["s1"][:n].astype(np.int64), lt["s1"])
    assert np.array_equal(c["s2"][:n].astype(np.int64), lt["s2"])
    p = lt["count"] > 0
    assert np.array_equal(c["bbox"][:n][p][:, :3], lt["bmin"][p]) and np.array_equal(c["bbox"][:n][p][:, 3:], lt["bmax"][p])
    assert np.array_equal(c["lo"], pt["lo"]) and np.array_equal(c["hi"], pt["hi"])
    assert np.array_equal(c["faces"].astype(np.int64), pt["faces"])
    assert np.array_equal(c["wall18"].astype(np.int64), pt["wall18"])


def test_voronoi_dome():
    _same(voronoi_numpy((30, 28, 44), 50, 3, dome=True))
    _same(voronoi_numpy((12, 40, 20), 30, 4, weights=(5, 2, 2), dtype=np.uint32))


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 10 ** 6), st.tuples(st.integers(1, 7), st.integers(1, 7), st.integers(1, 9)))
def test_noise(seed, shape):
    rng = np.random.default_rng(seed)
    _same(rng.integers(0, 6, size=shape).astype(np.uint16))

"""Pin the loop oracle to every known-answer the reference itself states.

The reference has no tests (test/__init__.py:1-11); its only pinned results are the
docstring examples on one 4x6 toy image (SIA = spatial_image_analysis.py):
labels SIA:343-353, center_of_mass SIA:437-450, boundingbox SIA:498-511,
neighbors SIA:561-574, cell_wall_area SIA:924-927, wall_areas SIA:978-982,
volume SIA:1219-1226.  The examples were written when 2D input was reshaped to
(4, 6, 1); that is the shape used here.
"""
import numpy as np
import pytest

from oracle.sia_loops import LoopOracle, LIST

TOY = np.array([[1, 2, 7, 7, 1, 1],
                [1, 6, 5, 7, 3, 3],
                [2, 2, 1, 7, 3, 3],
                [1, 1, 1, 4, 1, 1]], dtype=np.uint16).reshape(4, 6, 1)


@pytest.fixture()
def sia():
    with pytest.warns(UserWarning):
        return LoopOracle(TOY.copy())


def test_labels(sia):
    assert sorted(sia.labels()) == [1, 2, 3, 4, 5, 6, 7]
    assert sia.nb_labels() == 7


def test_center_of_mass(sia):
    assert list(sia.center_of_mass(7)) == [0.75, 2.75, 0.0]
    two = sia.center_of_mass([7, 2])
    assert list(two[7]) == [0.75, 2.75, 0.0]
    assert list(two[2]) == [1.3333333333333333, 0.66666666666666663, 0.0]
    allc = sia.center_of_mass()
    expect = {1: [1.8, 2.2999999999999998, 0.0], 2: [1.3333333333333333, 0.66666666666666663, 0.0],
              3: [1.5, 4.5, 0.0], 4: [3.0, 3.0, 0.0], 5: [1.0, 2.0, 0.0], 6: [1.0, 1.0, 0.0],
              7: [0.75, 2.75, 0.0]}
    assert {k: list(v) for k, v in allc.items()} == expect


def test_boundingbox(sia):
    assert sia.boundingbox(7) == (slice(0, 3), slice(2, 4), slice(0, 1))
    two = sia.boundingbox([7, 2])
    assert two[7] == (slice(0, 3), slice(2, 4), slice(0, 1))
    assert two[2] == (slice(0, 3), slice(0, 2), slice(0, 1))
    expect = [(slice(0, 4), slice(0, 6), slice(0, 1)), (slice(0, 3), slice(0, 2), slice(0, 1)),
              (slice(1, 3), slice(4, 6), slice(0, 1)), (slice(3, 4), slice(3, 4), slice(0, 1)),
              (slice(1, 2), slice(2, 3), slice(0, 1)), (slice(1, 2), slice(1, 2), slice(0, 1)),
              (slice(0, 3), slice(2, 4), slice(0, 1))]
    allb = sia.boundingbox()
    assert [allb[l] for l in range(1, 8)] == expect


def test_neighbors(sia):
    assert sorted(sia.neighbors(7)) == [1, 2, 3, 4, 5]
    two = sia.neighbors([7, 2])
    assert {k: sorted(v) for k, v in two.items()} == {7: [1, 2, 3, 4, 5], 2: [1, 6, 7]}
    expect = {1: [2, 3, 4, 5, 6, 7], 2: [1, 6, 7], 3: [1, 7], 4: [1, 7], 5: [1, 6, 7], 6: [1, 2, 5],
              7: [1, 2, 3, 4, 5]}
    assert {k: sorted(v) for k, v in sia.neighbors().items()} == expect


def test_cell_wall_area(sia):
    assert sia.cell_wall_area(7, 2) == 1.0
    assert sia.cell_wall_area(7, [2, 5]) == {(2, 7): 1.0, (5, 7): 2.0}


def test_wall_areas(sia):
    assert sia.wall_areas({1: [2, 3], 2: [6]}) == {(1, 2): 5.0, (1, 3): 4.0, (2, 6): 2.0}
    expect = {(1, 2): 5.0, (1, 3): 4.0, (1, 4): 2.0, (1, 5): 1.0, (1, 6): 1.0, (1, 7): 2.0, (2, 6): 2.0,
              (2, 7): 1.0, (3, 7): 2, (4, 7): 1, (5, 6): 1.0, (5, 7): 2.0}
    assert sia.wall_areas() == expect


def test_volume(sia):
    assert sia.volume(7) == {7: 4.0}  # today's code returns a 1-entry dict (SIA:1240-1241)
    v = sia.volume([7, 2])
    assert v == {7: 4.0, 2: 3.0}
    allv = sia.volume()
    assert [allv[l] for l in range(1, 8)] == [10.0, 3.0, 4.0, 1.0, 1.0, 1.0, 4.0]
    with pytest.warns(UserWarning):
        aslist = LoopOracle(TOY.copy(), return_type=LIST)
    got = aslist.volume()
    assert sorted(got) == sorted([10.0, 3.0, 4.0, 1.0, 1.0, 1.0, 4.0])

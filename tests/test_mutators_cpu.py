"""Image mutators (SIA:1114-1176) and label -> value images (property_spatial_image.py:207-221) on oracle-fed tables."""
import warnings

import numpy as np

from oracle.sia_loops import LoopOracle
from tests.helpers import OracleBackend, compare_api
from tissue_analysis_b200 import SpatialImage, SpatialImageAnalysis3D
from tissue_analysis_b200.property_spatial_image import create_property_image
from tissue_analysis_b200.synth import tissue_image

warnings.filterwarnings("ignore")


def reference_mutation(arr, fuse=None, remove=None, erase=0):
    """What SIA:1114-1165 do to the image, as whole-array numpy."""
    out = arr.copy()
    if fuse:
        for l in fuse:
            out[arr == l] = min(fuse)
    if remove:
        for l in remove:
            out[arr == l] = erase
    return out


def test_fuse_and_remove_then_features_match_a_fresh_analysis():
    img = tissue_image((36, 30, 28), 30, seed=31, dome=True)
    arr = np.ascontiguousarray(np.asarray(img))
    work = SpatialImage(arr.copy(), voxelsize=img.voxelsize)
    prod = SpatialImageAnalysis3D(work, background=1, _backend=OracleBackend(work))
    labels = sorted(prod.labels())
    fuse, remove = labels[2:5], labels[7:9]
    prod.fuse_labels_in_image(list(fuse), verbose=False)
    prod.remove_labels_from_image(list(remove), verbose=False)
    expect = reference_mutation(arr, fuse=fuse, remove=remove)
    assert np.array_equal(np.asarray(work), expect)                         # the caller's image was edited in place
    assert 0 in prod.ignoredlabels() and fuse[1] not in prod.labels() and remove[0] not in prod.labels()
    fresh = LoopOracle(expect, voxelsize=img.voxelsize, background=1, ignoredlabels=[0])
    compare_api(prod, fresh, check_wall_voxels=False)


def test_property_image_is_a_lookup():
    img = tissue_image((30, 26, 22), 20, seed=32, dome=True)
    prod = SpatialImageAnalysis3D(img, background=1, _backend=OracleBackend(img))
    vol = prod.volume(real=False)
    some = dict((l, v) for l, v in vol.items() if l % 2 == 0)
    out = create_property_image(prod, some, dtype=np.uint16)
    arr = np.asarray(img)
    expect = np.full(arr.shape, 1, np.uint16)                                # missing labels and background -> background
    for l, v in some.items():
        expect[arr == l] = np.float64(v).astype(np.uint16)
    assert np.array_equal(np.asarray(out), expect) and out.voxelsize == img.voxelsize

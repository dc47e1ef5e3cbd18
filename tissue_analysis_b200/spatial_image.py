"""Minimal stand-in for ``openalea.image.SpatialImage``.

The reference wraps every input in ``openalea.image.serial.basics.SpatialImage``
(reference: src/vplants/tissue_analysis/spatial_image_analysis.py:27, 224-227), an
``ndarray`` subclass that carries ``voxelsize`` (and an ``info`` dict).  That
package is not vendored with the reference, so this module provides the two
attributes the hot path reads (``voxelsize``: SIA:246, ``info``: SIA:262-266).

Any memory order is accepted; the scan shards along the slowest memory axis.
"""
import numpy as np


class SpatialImage(np.ndarray):
    def __new__(cls, input_array, voxelsize=None, info=None, dtype=None, **kwargs):
        arr = np.asarray(input_array, dtype=dtype)
        obj = arr.view(cls)
        if voxelsize is None:
            voxelsize = getattr(input_array, "voxelsize", None)
        if voxelsize is None:
            voxelsize = (1.0,) * arr.ndim
        obj.voxelsize = tuple(float(v) for v in voxelsize)
        obj.info = dict(info) if info else dict(getattr(input_array, "info", {}) or {})
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.voxelsize = getattr(obj, "voxelsize", (1.0,) * self.ndim)
        self.info = getattr(obj, "info", {})

    # ``resolution`` is the historical alias used by older openalea code
    @property
    def resolution(self):
        return self.voxelsize

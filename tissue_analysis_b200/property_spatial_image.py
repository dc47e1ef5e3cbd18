"""Label -> value images from a lookup table.

Mirrors the image builders of the reference that turn a per-label property into an image of the same shape:
``PropertySpatialImage.create_property_image`` (/root/reference/src/vplants/tissue_analysis/property_spatial_image.py:207-221)
and the ``'volume'`` branch of ``spatial_image_analysis_to_spatial_image``
(tissue_analysis_oalab/sia_to_spatial_image.py:26-55).  Both evaluate ``property_dict.values(image)`` with openalea's
``array_dict``, i.e. a gather ``out[p] = table[image[p]]``; here that gather is one streaming CUDA kernel over the
volume already resident on the device (``ta_map_labels``).  The rest of ``PropertySpatialImage`` (graph-backed property
computation, VTK meshing) is outside this path.
"""
import numpy as np

from .spatial_image import SpatialImage


def create_property_image(analysis, property_dict, background=None, dtype=np.uint16):
    """property_spatial_image.py:207-221: labels missing from ``property_dict`` and the background map to the
    background value; the value image is cast to ``dtype`` (numpy ``astype`` per value)."""
    bg = analysis.background() if background is None else background
    t = analysis._tables()
    dt = np.dtype(dtype)
    if dt not in (np.dtype(np.uint16), np.dtype(np.uint32)):
        raise ValueError("property images are uint16 or uint32")
    values = np.full(t.nrows, bg, dtype=np.float64)
    for l, v in property_dict.items():
        if 0 <= int(l) < t.nrows:
            values[int(l)] = v
    if bg is not None and 0 <= bg < t.nrows:
        values[bg] = bg
    lut = values.astype(dt)
    out = analysis._scan().map_labels(lut, fill=int(np.float64(bg).astype(dt)))
    return SpatialImage(out, voxelsize=analysis.image.voxelsize)


def spatial_image_analysis_to_spatial_image(analysis, property_name=None):
    """sia_to_spatial_image.py:26-55 without the label-removal copy: the 'volume' image, background kept."""
    if property_name != 'volume':
        return analysis.image
    labels = analysis.labels()
    vol = analysis.volume(labels)
    t = analysis._tables()
    lut = np.zeros(t.nrows, np.float64)
    for l, v in vol.items():
        lut[l] = v
    bg = analysis.background()
    out = analysis._scan().map_labels(lut.astype(np.uint16), fill=0)
    img = SpatialImage(out, voxelsize=analysis.image.voxelsize)
    if bg is not None:
        img[np.asarray(analysis.image) == bg] = bg
    return img

"""Time series of label volumes (BASELINE config C5): independent frames, one frame per GPU at a time.

The reference analyses each time point with a fresh ``SpatialImageAnalysis`` object (there is no temporal coupling in
the feature extractors: spatial_image_analysis.py:1663-1680); frames are therefore independent units.  Under
``torchrun`` rank r takes frames r, r + world, ...; no data-path collective is involved ("replicas only").  One
context (device buffers, tables) is reused across the frames of a rank.
"""
import os

import numpy as np

from . import _native
from .engine import memory_layout, tables_from_memory_order


def frames_of_rank(n_frames, rank=None, world=None):
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    return list(range(rank, n_frames, world))


def analyze_frames(frames, device=-1, flags=_native.PASS_ALL, rank=None, world=None):
    """frames: sequence of 3D uint16/uint32 arrays (or callables returning one).  Returns {frame index: ScanTables}
    for the frames this rank owns."""
    ctx = _native.Context(device)
    out = {}
    try:
        for k in frames_of_rank(len(frames), rank, world):
            img = frames[k]() if callable(frames[k]) else frames[k]
            view, ax = memory_layout(img)
            ctx.run_pass_host(view, flags)
            count, s1, s2, bbox = ctx.label_table()
            lo, hi, faces, wall = ctx.pair_table()
            out[k] = tables_from_memory_order(np.asarray(img).shape, ax, count, s1, s2, bbox, lo, hi, faces, wall)
    finally:
        ctx.close()
    return out

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


def analyze_frames(frames, device=-1, flags=_native.PASS_ALL, rank=None, world=None, pipeline=2):
    """frames: sequence of 3D uint16/uint32 arrays (or callables returning one).  Returns {frame index: ScanTables}
    for the frames this rank owns.

    ``pipeline`` contexts (device buffers, tables, streams) work side by side, one thread each: while one frame's tables
    travel back and become ScanTables, the next frame's upload + scan (``ta_run_pass_host``: chunked H2D with the scan of
    each chunk queued behind its copy) is already running.  The upload dominates a frame (2 GiB over PCIe against 6 ms of
    scan at C5's size), so the pipeline hides everything but it.  ctypes releases the GIL inside the library calls."""
    import threading
    mine = frames_of_rank(len(frames), rank, world)
    out, errors = {}, []
    nctx = max(1, min(int(pipeline), len(mine)))

    def work(slot):
        ctx = _native.Context(device)
        try:
            for k in mine[slot::nctx]:
                img = frames[k]() if callable(frames[k]) else frames[k]
                view, ax = memory_layout(img)
                ctx.run_pass_host(view, flags)
                count, s1, s2, bbox = ctx.label_table()
                lo, hi, faces, wall = ctx.pair_table()
                out[k] = tables_from_memory_order(np.asarray(img).shape, ax, count, s1, s2, bbox, lo, hi, faces, wall)
        except Exception as e:          # noqa: BLE001 -- re-raised in the caller's thread
            errors.append(e)
        finally:
            ctx.close()

    if nctx == 1:
        work(0)
    else:
        threads = [threading.Thread(target=work, args=(i,)) for i in range(nctx)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    return out

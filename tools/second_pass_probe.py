"""The second-pass kernels (ta_second_pass.cuh) once each on a synthetic configuration resident on the device -- the command
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` is pointed at for their GB/s
(tools/r02_second_pass.sh).  The wrappers copy their result image to the host, so wall time says nothing about the kernels."""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tissue_analysis_b200 import _native
from tissue_analysis_b200.synth import CONFIGS, voronoi_device

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
a = ap.parse_args()
cfg = dict(CONFIGS[a.config])
X, Y, Z = cfg["shape"]
vol = voronoi_device((Z, Y, X), cfg["ncell"], cfg["seed"], cfg["weights"][::-1], cfg["dome"], cfg["dtype"])
dt = np.uint16 if vol.element_size() == 2 else np.uint32
ctx = _native.Context()
ctx.bind_device(vol.data_ptr(), vol.element_size(), X, Y, Z, keepalive=vol)
ctx.run_pass(_native.PASS_ALL, cfg["ncell"] + 1 if vol.element_size() == 4 else 0)
shape = (Z, Y, X)
lut = np.arange(max(65536, cfg["ncell"] + 2), dtype=dt)
ctx.map_labels(lut, 0, shape)                       # LUT gather: read + write per voxel
ctx.voxel_first_layer(1, True, shape, dt)           # first layer of voxels against the background
ctx.stencil_image("hollow", shape, dt)              # wrap-around Laplacian mask
ctx.stencil_image("shell18", shape, dt)             # 18-connected outer shell of every cell
lo, hi, faces, wall = ctx.pair_table()
k = min(64, lo.size)
ctx.wall_voxel_coords(lo[:k], hi[:k])               # coordinates of the wall voxels of the first pairs
print("done: %d voxels" % (X * Y * Z))

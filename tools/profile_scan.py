"""Small driver for ncu: a few full passes over one synthetic tissue resident on the device."""
import argparse
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tissue_analysis_b200 import _native
from tissue_analysis_b200.synth import CONFIGS, voronoi_device

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--passes", type=int, default=3)
ap.add_argument("--flags", type=int, default=7)
ap.add_argument("--ncell", type=int, default=0, help="override the number of seeds (cell size study)")
a = ap.parse_args()
cfg = dict(CONFIGS[a.config])
if a.ncell:
    cfg["ncell"] = a.ncell
X, Y, Z = cfg["shape"]
vol = voronoi_device((Z, Y, X), cfg["ncell"], cfg["seed"], cfg["weights"][::-1], cfg["dome"], cfg["dtype"])
ctx = _native.Context()
ctx.bind_device(vol.data_ptr(), vol.element_size(), X, Y, Z, keepalive=vol)
for i in range(a.passes):
    ctx.run_pass(a.flags, cfg["ncell"] + 1 if vol.element_size() == 4 else 0)
    t = ctx.last_timing()
    print("pass %d: scan %.3f ms, pass %.3f ms -> %.1f Gvoxel/s" % (i, t["scan_ms"], t["pass_ms"], X * Y * Z / t["scan_ms"] / 1e6))
lo, hi, faces, wall = ctx.pair_table()
print("pairs %d, sum wall18 / voxels = %.3f, sum faces / voxels = %.3f" % (lo.size, wall.sum() / (X * Y * Z), faces.sum() / (X * Y * Z)))

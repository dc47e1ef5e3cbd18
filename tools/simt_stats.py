"""Data statistics behind the scan kernel's phase costs, computed on the CPU for a C3-like tissue (same cell density).

The kernel is instruction-bound and its irregular phases run warp-wide: what a warp executes is the MAXIMUM over its 32
lanes of every per-lane loop count.  This tool measures, on real synthetic data, the per-lane quantities the phases loop
over and their per-warp maxima for the kernel's thread mappings, so that a candidate formulation can be costed before any
GPU time is spent on it (the one-hot pair path looked 1.6x faster per thread and lost because of exactly this).

    python tools/simt_stats.py [--shape 128 256 256] [--cell 21500]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tissue_analysis_b200.synth import voronoi_numpy  # noqa: E402

SEG, NFS, BM, BS = 8, 16, 16, 8


def shifted_neq(v, axis, step=1):
    """bool array: v differs from its +step neighbour along axis (edge: False)."""
    out = np.zeros(v.shape, bool)
    a = [slice(None)] * 3
    b = [slice(None)] * 3
    a[axis] = slice(0, -step)
    b[axis] = slice(step, None)
    out[tuple(a)] = v[tuple(a)] != v[tuple(b)]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=3, default=[128, 256, 256], help="z y x (x fastest)")
    ap.add_argument("--cell", type=float, default=21500.0, help="voxels per cell (C3: 1024^3 / 50000)")
    ap.add_argument("--seed", type=int, default=2)
    a = ap.parse_args()
    Z, Y, X = a.shape
    ncell = max(2, int(round(Z * Y * X / a.cell)))
    v = voronoi_numpy((Z, Y, X), ncell, a.seed).astype(np.int64)
    print("volume %dx%dx%d (z,y,x), %d cells, %.0f voxels per cell (cube side %.1f)" % (Z, Y, X, ncell, v.size / ncell,
                                                                                   (v.size / ncell) ** (1 / 3)))
    # ---- wall voxels (18-neighbourhood) and junction voxels (>= 2 other labels around) ---------------------------
    pad = np.pad(v, 1, mode="edge")
    others = []                                                  # per voxel: sorted tuple of other labels via set ops
    diff_any = np.zeros(v.shape, bool)
    first_other = np.zeros(v.shape, np.int64)
    multi = np.zeros(v.shape, bool)
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                l1 = abs(dz) + abs(dy) + abs(dx)
                if l1 < 1 or l1 > 2:
                    continue
                nb = pad[1 + dz:1 + dz + Z, 1 + dy:1 + dy + Y, 1 + dx:1 + dx + X]
                d = nb != v
                new_first = d & ~diff_any
                first_other[new_first] = nb[new_first]
                multi |= d & diff_any & (nb != first_other)
                diff_any |= d
    wall = diff_any
    print("wall voxels (18-conn): %.1f %% of voxels; junction voxels (>= 2 other labels): %.1f %% of wall voxels" % (
        100 * wall.mean(), 100 * multi.sum() / max(wall.sum(), 1)))

    # ---- segments (8 voxels along x) ----------------------------------------------------------------------------
    nsx = X // SEG
    seg = v[:, :, :nsx * SEG].reshape(Z, Y, nsx, SEG)
    brk = seg[..., 1:] != seg[..., :-1]
    runs = 1 + brk.sum(-1)                                        # runs per segment
    padx = np.pad(v, ((0, 0), (0, 0), (1, 1)), mode="edge")
    left = padx[:, :, 0:nsx * SEG:SEG]
    right = padx[:, :, SEG + 1:nsx * SEG + 2:SEG]
    uni = (runs == 1) & (left == seg[..., 0]) & (right == seg[..., -1])   # the kernel's uniformity code
    print("segments: %.1f %% uniform (code != MIXED); runs per mixed segment: mean %.2f, 90th pct %d, max %d" % (
        100 * uni.mean(), runs[~uni].mean(), np.percentile(runs[~uni], 90), runs.max()))
    code = np.where(uni, seg[..., 0], -1)
    # interior segment: the 3x3 codes around it are one non-MIXED label
    cpad = np.pad(code, ((1, 1), (1, 1), (0, 0)), mode="edge")
    interior = code >= 0
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            interior &= cpad[1 + dz:1 + dz + Z, 1 + dy:1 + dy + Y] == code
    listed = ~interior
    wseg = wall[:, :, :nsx * SEG].reshape(Z, Y, nsx, SEG).sum(-1)
    jseg = multi[:, :, :nsx * SEG].reshape(Z, Y, nsx, SEG).sum(-1)
    print("listed segments (not interior by the 3x3 codes): %.1f %%; wall voxels per listed segment: mean %.2f; listed "
          "segments without a wall voxel: %.1f %%; with a junction voxel: %.1f %%" % (
              100 * listed.mean(), wseg[listed].mean(), 100 * (wseg[listed] == 0).mean(), 100 * (jseg[listed] > 0).mean()))

    # ---- distinct labels in the 3x3x(8+2) window of a listed segment, and distinct (own, other) pairs -------------
    zz, yy, ss = np.nonzero(listed)
    rng = np.random.default_rng(0)
    pick = rng.choice(len(zz), size=min(20000, len(zz)), replace=False)
    nlab, npair, nown = [], [], []
    for i in pick:
        z, y, s = zz[i], yy[i], ss[i]
        win = pad[z:z + 3, y:y + 3, s * SEG:s * SEG + SEG + 2]
        nlab.append(len(np.unique(win)))
        own = seg[z, y, s]
        w = wall[z, y, s * SEG:(s + 1) * SEG]
        keys = set()
        for j in range(SEG):
            if not w[j]:
                continue
            nbh = pad[z:z + 3, y:y + 3, s * SEG + j:s * SEG + j + 3]
            for o in np.unique(nbh):
                if o != own[j]:
                    keys.add((min(own[j], o), max(own[j], o)))
        npair.append(len(keys))
        nown.append(len(np.unique(own)))
    nlab, npair, nown = np.array(nlab), np.array(npair), np.array(nown)
    print("window of a listed segment (3x3x10): labels mean %.2f; <= 2 labels %.1f %%, 3 labels %.1f %%, >= 4 labels %.1f %%" % (
        nlab.mean(), 100 * (nlab <= 2).mean(), 100 * (nlab == 3).mean(), 100 * (nlab >= 4).mean()))
    print("distinct label pairs contributed by a listed segment: mean %.2f; 0: %.1f %%, 1: %.1f %%, 2: %.1f %%, >= 3: %.1f %%; "
          "own labels in the segment: mean %.2f" % (npair.mean(), 100 * (npair == 0).mean(), 100 * (npair == 1).mean(),
                                                   100 * (npair == 2).mean(), 100 * (npair >= 3).mean(), nown.mean()))

    # ---- per-warp maxima for the kernel's mappings ---------------------------------------------------------------
    # march (C1): a warp = 2 rows x 16 segments at one plane; what it executes is the max over its 32 lanes
    rows2 = runs[:, :(Y // 2) * 2, :(nsx // NFS) * NFS].reshape(Z, Y // 2, 2, nsx // NFS, NFS)
    warp_max = rows2.max(axis=(2, 4))
    warp_mean = rows2.mean(axis=(2, 4))
    print("march, runs per segment: lane mean %.2f, warp max mean %.2f  -> SIMT efficiency of the run loop %.0f %%" % (
        warp_mean.mean(), warp_max.mean(), 100 * warp_mean.mean() / warp_max.mean()))
    # worklists: 32 consecutive listed segments
    def warp_eff(per_item, name):
        n = (len(per_item) // 32) * 32
        w = per_item[:n].reshape(-1, 32)
        print("%s: lane mean %.2f, warp max mean %.2f -> SIMT efficiency %.0f %%" % (
            name, w.mean(), w.max(1).mean(), 100 * w.mean() / max(w.max(1).mean(), 1e-9)))
    order = np.lexsort((ss, yy, zz))
    warp_eff(wseg[listed][order].astype(float), "wall voxels per listed segment (list order)")
    warp_eff(runs[listed][order].astype(float), "runs per listed segment (list order)")
    warp_eff(npair[np.argsort(pick)].astype(float), "label pairs per listed segment (sampled, list order)")

    # ---- labels per tile (brick + halo) and per row quarter ---------------------------------------------------
    per_tile, per_quarter = [], []
    for z0 in range(0, Z - BS + 1, BS):
        for y0 in range(0, Y - BM + 1, BM):
            for x0 in range(0, X - NFS * SEG + 1, NFS * SEG):
                t = pad[z0:z0 + BS + 2, y0:y0 + BM + 2, x0:x0 + NFS * SEG + 2]
                per_tile.append(len(np.unique(t)))
                for q in range(4):
                    per_quarter.append(len(np.unique(t[:, :, q * 32:q * 32 + 34])))
    per_tile, per_quarter = np.array(per_tile), np.array(per_quarter)
    print("labels per tile (130x18x10): mean %.1f, > 16: %.1f %%;  per row quarter (34x18x10): mean %.1f, > 16: %.2f %%" % (
        per_tile.mean(), 100 * (per_tile > 16).mean(), per_quarter.mean(), 100 * (per_quarter > 16).mean()))

    # ---- labels in the window of a small 3-D block (candidate work unit: one thread per block) -------------------
    for (bx, by, bz) in ((8, 4, 2), (8, 2, 4), (4, 4, 4), (8, 4, 4), (8, 8, 2)):
        ks = []
        for z0 in range(0, Z - bz + 1, bz):
            zs = slice(z0, z0 + bz + 2)
            for y0 in range(0, Y - by + 1, by):
                ys = slice(y0, y0 + by + 2)
                row = pad[zs, ys, :]
                for x0 in range(0, X - bx + 1, bx):
                    w = row[:, :, x0:x0 + bx + 2]
                    a = w.flat[0]
                    if (w == a).all():
                        ks.append(1)
                    else:
                        ks.append(len(np.unique(w)))
        ks = np.array(ks)
        n = (len(ks) // 32) * 32
        wk = ks[:n].reshape(-1, 32)                      # 32 consecutive blocks along x = one warp
        print("block %dx%dx%d (window %dx%dx%d): labels in the window 1: %.1f %%, 2: %.1f %%, 3: %.1f %%, 4: %.1f %%, > 4: %.2f %%; "
              "mean %.2f, warp max mean %.2f" % (bx, by, bz, bx + 2, by + 2, bz + 2, 100 * (ks == 1).mean(), 100 * (ks == 2).mean(),
                                                100 * (ks == 3).mean(), 100 * (ks == 4).mean(), 100 * (ks > 4).mean(), ks.mean(),
                                                wk.max(1).mean()))


if __name__ == "__main__":
    main()

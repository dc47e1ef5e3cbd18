"""Bit-level Python model of the 'window of runs' pair phase (design check before writing CUDA).

For every row and every 32-voxel piece: the 9 neighbour rows are described by runs inside a 34-position window;
per label a few 64-bit masks are accumulated (centre row, dilated neighbours, +m row, +s row); wall18 and the
directional faces come out of popcounts.  The result must equal oracle/sia_onepass.pair_table exactly.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import sia_onepass

PIECE = 32


def popc(x):
    return bin(x).count("1")


def pairs_by_windows(img_xyz):
    """img axes (x, y, z) = (fast, mid, slow)."""
    X, Y, Z = img_xyz.shape
    pad = np.pad(img_xyz.astype(np.int64), 1, mode="edge")        # clamped halo
    out = {}
    npiece = (X + PIECE - 1) // PIECE
    roles = []
    for dm in (-1, 0, 1):
        for ds in (-1, 0, 1):
            roles.append((dm, ds))
    for z in range(Z):
        for y in range(Y):
            for q in range(npiece):
                x0 = q * PIECE
                nvalid = min(PIECE, X - x0)
                valid = ((1 << nvalid) - 1) << 1                 # window positions 1..nvalid
                labels, cen, dil, mp, sp = [], {}, {}, {}, {}
                for dm, ds in roles:
                    # window positions p = 0..33 <-> x = x0-1+p ; padded index = x+1 (clamped beyond the volume too)
                    xs = np.clip(np.arange(x0 - 1, x0 + 33), -1, X) + 1
                    win = pad[xs, y + 1 + dm, z + 1 + ds]
                    p = 0
                    while p < 34:
                        e = p
                        while e + 1 < 34 and win[e + 1] == win[p]:
                            e += 1
                        L = int(win[p])
                        I = ((1 << (e - p + 1)) - 1) << p
                        for d in (cen, dil, mp, sp):
                            d.setdefault(L, 0)
                        if dm == 0 and ds == 0:
                            cen[L] |= I
                            dil[L] |= ((I << 1) | (I >> 1))
                        elif dm == 0 or ds == 0:
                            dil[L] |= I | (I << 1) | (I >> 1)
                            if dm == 1:
                                mp[L] |= I
                            if ds == 1:
                                sp[L] |= I
                        else:
                            dil[L] |= I
                        p = e + 1
                labs = sorted(cen)
                for a in labs:
                    ca = cen[a] & valid
                    if not ca:
                        continue
                    for b in labs:
                        if a == b:
                            continue
                        w18 = popc(ca & dil[b])
                        ff = popc(ca & (cen[b] >> 1))            # voxel p is a, voxel p+1 is b
                        fm = popc(ca & mp[b])
                        fs = popc(ca & sp[b])
                        if not (w18 or ff or fm or fs):
                            continue
                        key = (min(a, b), max(a, b))
                        rec = out.setdefault(key, [0] * 7)
                        rec[6] += w18
                        lo = a < b                                # lower-index voxel carries the smaller label
                        rec[0 if lo else 1] += ff
                        rec[2 if lo else 3] += fm
                        rec[4 if lo else 5] += fs
    keys = sorted(out)
    return (np.array([k[0] for k in keys]), np.array([k[1] for k in keys]),
            np.array([out[k][:6] for k in keys]).reshape(-1, 6), np.array([out[k][6] for k in keys]))


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    from tissue_analysis_b200.synth import voronoi_numpy
    cases = [rng.integers(0, 5, size=(s)).astype(np.uint16) for s in [(7, 3, 2), (33, 4, 3), (70, 5, 4), (1, 1, 1), (40, 2, 6)]]
    cases.append(voronoi_numpy((9, 12, 75), 14, 1, dome=True).transpose(2, 1, 0))
    cases.append(voronoi_numpy((7, 10, 40), 6, 2).transpose(2, 1, 0))
    for img in cases:
        lo, hi, faces, wall = pairs_by_windows(img)
        ref = sia_onepass.pair_table(img)
        ok = (np.array_equal(lo, ref["lo"]) and np.array_equal(hi, ref["hi"]) and np.array_equal(faces, ref["faces"])
              and np.array_equal(wall, ref["wall18"]))
        print(img.shape, "pairs", len(lo), "OK" if ok else "MISMATCH")
        assert ok

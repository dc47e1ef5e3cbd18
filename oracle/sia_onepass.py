"""CPU ORACLE (test infrastructure, never shipped, never on the product path).

Vectorised one-pass numpy formulation of the per-label / per-pair *tables* the CUDA scan
emits.  Each definition below is the whole-volume equivalent of a per-label loop of the
reference (SIA = /root/reference/src/vplants/tissue_analysis/spatial_image_analysis.py):

  count / sums / bbox   <- nd.sum SIA:1231, nd.center_of_mass SIA:466, nd.find_objects SIA:517,
                           centred moments SIA:123-150, 1261-1278
  faces[6] per pair     <- six one-sided dilations SIA:695-716, 947-956
  wall18 per pair       <- 18-connected dilations of both masks SIA:796-799, 835-863

tests/test_oracle_equivalence.py proves these tables reproduce oracle/sia_loops.py (the
line-for-line restatement) on random small volumes, so they can stand in for it at sizes
where the per-label loops take hours.  Parity pinning is inherited from sia_loops.py
(docstring known-answers only; otherwise unpinned).

All axes here are API axes (x, y, z) = array axes (0, 1, 2).
"""
import numpy as np

N18 = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)
       if 1 <= abs(dx) + abs(dy) + abs(dz) <= 2]


def label_table(img, nlabels=None, slab=None):
    """Exact integer per-label moments.

    Returns dict of int64 arrays indexed by label: count[L], s1[L,3], s2[L,6] (xx,xy,xz,yy,yz,zz),
    bmin[L,3], bmax[L,3] (bmin > bmax where the label is absent).  ``slab`` = (z0, z1) restricts the
    owned voxels to planes z0 <= z < z1 of the *last* axis (multi-rank emulation).
    """
    img = np.asarray(img)
    L = int(img.max()) + 1 if nlabels is None else int(nlabels)
    X, Y, Z = img.shape
    z0, z1 = (0, Z) if slab is None else slab
    count = np.zeros(L, np.int64)
    s1 = np.zeros((L, 3), np.int64)
    s2 = np.zeros((L, 6), np.int64)
    big = np.iinfo(np.int64).max
    bmin = np.full((L, 3), big, np.int64)
    bmax = np.full((L, 3), -1, np.int64)
    yy, zz = np.meshgrid(np.arange(Y, dtype=np.int64), np.arange(z0, z1, dtype=np.int64), indexing="ij")
    yy = yy.ravel()
    zz = zz.ravel()
    for x in range(X):  # plane by plane keeps every float64 bincount exact and memory small
        lab = img[x, :, z0:z1].ravel().astype(np.int64)
        n = np.bincount(lab, minlength=L)
        count += n
        sy = np.bincount(lab, weights=yy, minlength=L).astype(np.int64)
        sz = np.bincount(lab, weights=zz, minlength=L).astype(np.int64)
        s1[:, 0] += n * x
        s1[:, 1] += sy
        s1[:, 2] += sz
        s2[:, 0] += n * x * x
        s2[:, 1] += sy * x
        s2[:, 2] += sz * x
        s2[:, 3] += np.bincount(lab, weights=yy * yy, minlength=L).astype(np.int64)
        s2[:, 4] += np.bincount(lab, weights=yy * zz, minlength=L).astype(np.int64)
        s2[:, 5] += np.bincount(lab, weights=zz * zz, minlength=L).astype(np.int64)
        present = n > 0
        bmin[present, 0] = np.minimum(bmin[present, 0], x)
        bmax[present, 0] = np.maximum(bmax[present, 0], x)
        for ax, coord in ((1, yy), (2, zz)):
            lo = np.full(L, big, np.int64)
            hi = np.full(L, -1, np.int64)
            np.minimum.at(lo, lab, coord)
            np.maximum.at(hi, lab, coord)
            bmin[:, ax] = np.minimum(bmin[:, ax], lo)
            bmax[:, ax] = np.maximum(bmax[:, ax], hi)
    return dict(count=count, s1=s1, s2=s2, bmin=bmin, bmax=bmax)


def _pair_key(a, b):
    lo = np.minimum(a, b).astype(np.uint64)
    hi = np.maximum(a, b).astype(np.uint64)
    return (lo << np.uint64(32)) | hi


def face_table(img, slab=None):
    """{key: int64[6]} directional face counts.  A face between p and p+e_a with u=img[p],
    w=img[p+e_a], u != w goes to slot 2a if u < w else 2a+1 (from the smaller label's side: slot 2a is
    the reference's kernel 2a, +axis; 2a+1 is kernel 2a+1, -axis).  With ``slab`` a face belongs to the
    rank owning its lower voxel p."""
    img = np.asarray(img)
    Z = img.shape[2]
    z0, z1 = (0, Z) if slab is None else slab
    keys, slots = [], []
    for a in range(3):
        lo_sl = [slice(None)] * 3
        hi_sl = [slice(None)] * 3
        if a == 2:
            top = min(z1, Z - 1)
            lo_sl[2] = slice(z0, top)
            hi_sl[2] = slice(z0 + 1, top + 1)
        else:
            lo_sl[a] = slice(0, -1)
            hi_sl[a] = slice(1, None)
            lo_sl[2] = hi_sl[2] = slice(z0, z1)
        u = img[tuple(lo_sl)]
        w = img[tuple(hi_sl)]
        m = u != w
        uu = u[m].astype(np.int64)
        ww = w[m].astype(np.int64)
        keys.append(_pair_key(uu, ww))
        slots.append(np.where(uu < ww, 2 * a, 2 * a + 1))
    keys = np.concatenate(keys)
    slots = np.concatenate(slots)
    uniq, inv = np.unique(keys, return_inverse=True)
    faces = np.zeros((len(uniq), 6), np.int64)
    np.add.at(faces, (inv, slots), 1)
    return uniq, faces


def wall18_table(img, slab=None):
    """{key: count}: number of voxels p with img[p] in {a, b} having an 18-neighbour of the other
    label (each voxel counted once per distinct other label).  With ``slab`` a voxel belongs to the
    rank owning it."""
    img = np.asarray(img)
    X, Y, Z = img.shape
    z0, z1 = (0, Z) if slab is None else slab
    pad = np.pad(img.astype(np.int64), 1, mode="edge")
    centre = pad[1:-1, 1:-1, 1 + z0:1 + z1]
    lin = np.arange(centre.size, dtype=np.int64).reshape(centre.shape)
    vox, other = [], []
    for dx, dy, dz in N18:
        nb = pad[1 + dx:1 + dx + X, 1 + dy:1 + dy + Y, 1 + dz + z0:1 + dz + z1]
        m = nb != centre
        vox.append(lin[m])
        other.append(nb[m])
    vox = np.concatenate(vox)
    other = np.concatenate(other)
    # one entry per (voxel, distinct other label)
    code = np.unique(vox * np.int64(2 ** 32) + other)
    v = code >> 32
    b = code & (2 ** 32 - 1)
    a = centre.ravel()[v]
    uniq, cnt = np.unique(_pair_key(a, b), return_counts=True)
    return uniq, cnt.astype(np.int64)


def pair_table(img, slab=None):
    """Sorted pair table: lo[P], hi[P], faces[P,6], wall18[P] over the union of 6- and 18-connected
    contacts (18-only contacts have all-zero faces)."""
    fk, faces = face_table(img, slab)
    wk, wcnt = wall18_table(img, slab)
    keys = np.union1d(fk, wk)
    f = np.zeros((len(keys), 6), np.int64)
    w = np.zeros(len(keys), np.int64)
    f[np.searchsorted(keys, fk)] = faces
    w[np.searchsorted(keys, wk)] = wcnt
    lo = (keys >> np.uint64(32)).astype(np.int64)
    hi = (keys & np.uint64(2 ** 32 - 1)).astype(np.int64)
    return dict(lo=lo, hi=hi, faces=f, wall18=w)


def merge_label_tables(parts):
    out = dict(count=sum(p["count"] for p in parts), s1=sum(p["s1"] for p in parts),
               s2=sum(p["s2"] for p in parts))
    out["bmin"] = np.minimum.reduce([p["bmin"] for p in parts])
    out["bmax"] = np.maximum.reduce([p["bmax"] for p in parts])
    return out


def merge_pair_tables(parts):
    keys = np.concatenate([(p["lo"].astype(np.uint64) << np.uint64(32)) | p["hi"].astype(np.uint64)
                           for p in parts])
    faces = np.concatenate([p["faces"] for p in parts])
    wall = np.concatenate([p["wall18"] for p in parts])
    uniq, inv = np.unique(keys, return_inverse=True)
    f = np.zeros((len(uniq), 6), np.int64)
    w = np.zeros(len(uniq), np.int64)
    np.add.at(f, inv, faces)
    np.add.at(w, inv, wall)
    return dict(lo=(uniq >> np.uint64(32)).astype(np.int64),
                hi=(uniq & np.uint64(2 ** 32 - 1)).astype(np.int64), faces=f, wall18=w)

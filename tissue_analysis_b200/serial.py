"""``imread`` / ``imsave`` for INRIMAGE-4 label volumes (``.inr``, ``.inr.gz``): the on-disk format either side of the
hot path.

The reference reads and writes its tissues through ``openalea.image.serial.basics.imread / imsave``
(src/vplants/tissue_analysis/spatial_image_analysis.py:27, 1668-1671; temporal_graph_from_image.py:22), a third-party
package that is not vendored with it.  This module restates the published INRIMAGE-4 container that package uses for
segmented stacks: a text header of ``KEY=value`` lines between ``#INRIMAGE-4#{`` and ``##}``, padded with newlines to
a multiple of 256 bytes, followed by the raw voxels with x fastest, then y, then z (and the ``VDIM`` components of a
voxel innermost).  No fixture of the reference pins it (the reference ships no image): parity unpinned, round trips
and the byte layout are what the tests check.  TIFF stacks are not read.

An uncompressed file is memory-mapped copy-on-write, so a 16 GiB stack is not read twice on its way to
``ta_run_pass_host``; the returned ``SpatialImage`` is x-fastest (Fortran order), the layout the scan uses as is.
"""
import gzip
import os

import numpy as np

from .spatial_image import SpatialImage

_LITTLE = ("decm", "alpha", "pc")
_BIG = ("sun", "sgi")
_RESERVED = ("XDIM", "YDIM", "ZDIM", "VDIM", "TYPE", "PIXSIZE", "SCALE", "CPU", "VX", "VY", "VZ")


def _open(path, mode):
    return gzip.open(path, mode) if path.endswith(".gz") else open(path, mode)


def _read_header(f):
    """-> (dict of str, header length in bytes)."""
    head = f.read(256)
    if not head.startswith(b"#INRIMAGE-4#{"):
        raise IOError("not an INRIMAGE-4 file (bad magic)")
    while b"##}" not in head:
        more = f.read(256)
        if not more:
            raise IOError("INRIMAGE-4 header is not terminated by '##}'")
        head += more
    prop = {}
    for line in head.decode("latin-1").split("\n"):
        line = line.strip()
        if not line or line.startswith("#") or "=" not in line:
            continue
        key, val = line.split("=", 1)
        prop[key.strip()] = val.strip()
    return prop, len(head)


def _dtype_of(prop):
    kind = prop.get("TYPE", "unsigned fixed")
    bits = int(prop.get("PIXSIZE", "8 bits").split()[0])
    if kind == "unsigned fixed":
        base = {8: "u1", 16: "u2", 32: "u4", 64: "u8"}[bits]
    elif kind == "signed fixed":
        base = {8: "i1", 16: "i2", 32: "i4", 64: "i8"}[bits]
    elif kind == "float":
        base = {32: "f4", 64: "f8"}[bits]
    else:
        raise IOError("unsupported INRIMAGE TYPE %r" % kind)
    cpu = prop.get("CPU", "decm")
    if cpu in _BIG:
        return np.dtype(">" + base)
    if cpu in _LITTLE:
        return np.dtype("<" + base)
    raise IOError("unsupported INRIMAGE CPU %r" % cpu)


def imread(filename):
    """Read an INRIMAGE-4 stack -> ``SpatialImage`` of shape (x, y, z), x fastest in memory, ``voxelsize`` from
    VX / VY / VZ; the remaining header entries go to ``info``.  Vector images (VDIM > 1) get shape (x, y, z, v)."""
    if not isinstance(filename, str):
        raise TypeError("imread needs a file name")
    lower = filename.lower()
    if not (lower.endswith(".inr") or lower.endswith(".inr.gz")):
        raise NotImplementedError("only INRIMAGE-4 stacks (.inr, .inr.gz) are read; got %r" % os.path.basename(filename))
    with _open(filename, "rb") as f:
        prop, hlen = _read_header(f)
        dims = [int(prop[k]) for k in ("XDIM", "YDIM", "ZDIM")]
        vdim = int(prop.get("VDIM", 1))
        dt = _dtype_of(prop)
        count = dims[0] * dims[1] * dims[2] * vdim
        if filename.endswith(".gz"):
            raw = f.read(count * dt.itemsize)
            if len(raw) != count * dt.itemsize:
                raise IOError("INRIMAGE data is truncated")
            flat = np.frombuffer(raw, dtype=dt).copy()
        else:
            if os.path.getsize(filename) < hlen + count * dt.itemsize:
                raise IOError("INRIMAGE data is truncated")
            flat = np.memmap(filename, dtype=dt, mode="c", offset=hlen, shape=(count,))
    if not dt.isnative:
        flat = flat.astype(dt.newbyteorder("="))
    if vdim == 1:
        arr = flat.reshape(dims, order="F")
    else:
        arr = np.moveaxis(flat.reshape([vdim] + dims, order="F"), 0, -1)
    vox = tuple(float(prop.get(k, 1.0)) for k in ("VX", "VY", "VZ"))
    info = {k: v for k, v in prop.items() if k not in _RESERVED}
    return SpatialImage(arr, voxelsize=vox, info=info)


def imsave(filename, img):
    """Write a 3-D (or (x, y, z, v)) array as INRIMAGE-4; ``voxelsize`` and ``info`` of a ``SpatialImage`` are kept."""
    lower = filename.lower()
    if not (lower.endswith(".inr") or lower.endswith(".inr.gz")):
        raise NotImplementedError("only INRIMAGE-4 stacks (.inr, .inr.gz) are written")
    arr = np.asarray(img)
    if arr.ndim not in (3, 4):
        raise ValueError("INRIMAGE stacks are 3-D (or 3-D with a vector per voxel)")
    vdim = 1 if arr.ndim == 3 else arr.shape[3]
    dt = arr.dtype
    if dt.kind == "u":
        kind = "unsigned fixed"
    elif dt.kind == "i":
        kind = "signed fixed"
    elif dt.kind == "f" and dt.itemsize in (4, 8):
        kind = "float"
    else:
        raise ValueError("dtype %s cannot be stored in an INRIMAGE" % dt)
    vox = tuple(getattr(img, "voxelsize", (1.0, 1.0, 1.0)))[:3]
    lines = ["#INRIMAGE-4#{", "XDIM=%d" % arr.shape[0], "YDIM=%d" % arr.shape[1], "ZDIM=%d" % arr.shape[2],
             "VDIM=%d" % vdim, "TYPE=%s" % kind, "PIXSIZE=%d bits" % (8 * dt.itemsize), "SCALE=2**0", "CPU=decm",
             "VX=%s" % repr(float(vox[0])), "VY=%s" % repr(float(vox[1])), "VZ=%s" % repr(float(vox[2]))]
    for k, v in (getattr(img, "info", None) or {}).items():
        if k not in _RESERVED and "\n" not in str(k) and "\n" not in str(v):
            lines.append("%s=%s" % (k, v))
    head = ("\n".join(lines) + "\n").encode("latin-1")
    pad = (-(len(head) + 4)) % 256
    head += b"\n" * pad + b"##}\n"
    assert len(head) % 256 == 0
    little = arr.astype(dt.newbyteorder("<"), copy=False)
    data = little if vdim == 1 else np.moveaxis(little, -1, 0)
    with _open(filename, "wb") as f:
        f.write(head)
        f.write(np.asfortranarray(data).tobytes(order="F"))

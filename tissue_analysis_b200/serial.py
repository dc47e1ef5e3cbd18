"""``imread`` / ``imsave`` for INRIMAGE-4 label volumes (``.inr``, ``.inr.gz``): the on-disk format either side of the
hot path.

The reference reads and writes its tissues through ``openalea.image.serial.basics.imread / imsave``
(src/vplants/tissue_analysis/spatial_image_analysis.py:27, 1668-1671; temporal_graph_from_image.py:22), a third-party
package that is not vendored with it.  This module restates the published INRIMAGE-4 container that package uses for
segmented stacks: a text header of ``KEY=value`` lines between ``#INRIMAGE-4#{`` and ``##}``, padded with newlines to
a multiple of 256 bytes, followed by the raw voxels with x fastest, then y, then z (and the ``VDIM`` components of a
voxel innermost).  No fixture of the reference pins it (the reference ships no image): parity unpinned, round trips
and the byte layout are what the tests check.

TIFF stacks (``.tif``, ``.tiff``: the factory's other input, SIA:1668-1671) are read by a baseline reader restated from
the TIFF 6.0 specification: classic (42) files of either byte order, one grey sample of 8 / 16 / 32 unsigned or signed
bits or a 32 / 64-bit float per pixel, uncompressed strips, one page per z plane -- or the contiguous ImageJ stack (one
IFD, ``images=N`` in the description).  Voxel size from X / YResolution (+ ResolutionUnit) and ImageJ's ``spacing=``.
Compressed, tiled, palette and RGB files raise ``NotImplementedError``.

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
    if lower.endswith(".tif") or lower.endswith(".tiff"):
        return _imread_tiff(filename)
    if not (lower.endswith(".inr") or lower.endswith(".inr.gz")):
        raise NotImplementedError("only INRIMAGE-4 (.inr, .inr.gz) and TIFF (.tif, .tiff) stacks are read; got %r"
                                  % os.path.basename(filename))
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
    if lower.endswith(".tif") or lower.endswith(".tiff"):
        return _imsave_tiff(filename, img)
    if not (lower.endswith(".inr") or lower.endswith(".inr.gz")):
        raise NotImplementedError("only INRIMAGE-4 (.inr, .inr.gz) and TIFF (.tif, .tiff) stacks are written")
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


# ---- TIFF 6.0, baseline grey stacks -----------------------------------------------------------------------------------------
_TIFF_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d", 16: "Q"}


def _tiff_ifds(buf):
    """Every image file directory of a classic TIFF -> list of {tag: tuple of values}; byte order character."""
    import struct
    mark = bytes(buf[:2])
    if mark == b"II":
        bo = "<"
    elif mark == b"MM":
        bo = ">"
    else:
        raise IOError("not a TIFF file (bad byte-order mark)")
    magic, off = struct.unpack(bo + "HI", bytes(buf[2:8]))
    if magic != 42:
        raise NotImplementedError("only classic TIFF (42) is read, not BigTIFF")
    ifds, seen = [], set()
    while off and off not in seen:
        seen.add(off)
        (n,) = struct.unpack(bo + "H", bytes(buf[off:off + 2]))
        tags = {}
        for i in range(n):
            e = off + 2 + 12 * i
            tag, typ, cnt = struct.unpack(bo + "HHI", bytes(buf[e:e + 8]))
            if typ not in _TIFF_TYPES:
                continue
            fmt = _TIFF_TYPES[typ]
            size = struct.calcsize(bo + fmt) * cnt
            pos = e + 8 if size <= 4 else struct.unpack(bo + "I", bytes(buf[e + 8:e + 12]))[0]
            if typ == 2:
                tags[tag] = (bytes(buf[pos:pos + cnt]).split(b"\0")[0].decode("latin-1"),)
            else:
                vals = struct.unpack(bo + fmt * cnt, bytes(buf[pos:pos + size]))
                tags[tag] = tuple(vals[k] / vals[k + 1] if vals[k + 1] else 0.0 for k in range(0, len(vals), 2)) if typ in (5, 10) else vals
        ifds.append(tags)
        (off,) = struct.unpack(bo + "I", bytes(buf[off + 2 + 12 * n:off + 6 + 12 * n]))
    if not ifds:
        raise IOError("TIFF file without an image directory")
    return ifds, bo


def _imread_tiff(filename):
    buf = np.memmap(filename, dtype=np.uint8, mode="r")
    ifds, bo = _tiff_ifds(buf)
    t0 = ifds[0]
    width, height = int(t0[256][0]), int(t0[257][0])
    bits, fmt = int(t0.get(258, (1,))[0]), int(t0.get(339, (1,))[0])
    if int(t0.get(259, (1,))[0]) != 1:
        raise NotImplementedError("compressed TIFF (Compression = %d) is not read" % t0[259][0])
    if int(t0.get(277, (1,))[0]) != 1 or int(t0.get(262, (1,))[0]) > 1 or 322 in t0:
        raise NotImplementedError("only grey, strip-organised TIFF stacks are read (no RGB, palette or tiles)")
    try:
        dt = np.dtype(bo + {1: "u", 2: "i", 3: "f"}[fmt] + str(bits // 8))
    except (KeyError, TypeError):
        raise NotImplementedError("TIFF samples of %d bits, format %d" % (bits, fmt))
    plane = width * height * dt.itemsize
    desc = t0.get(270, ("",))[0]
    ij = dict(kv.split("=", 1) for kv in desc.split("\n") if "=" in kv) if desc.startswith("ImageJ") else {}
    nz = len(ifds)
    if nz == 1 and int(ij.get("images", 1)) > 1:                     # ImageJ: the whole stack behind the first strip
        nz = int(ij["images"])
        start = int(t0[273][0])
        if start + nz * plane > buf.size:
            raise IOError("TIFF data is truncated")
        vol = np.frombuffer(buf, dtype=dt, count=nz * width * height, offset=start).reshape(nz, height, width)
    else:
        vol = np.empty((nz, height, width), dt)
        for z, t in enumerate(ifds):
            if int(t[256][0]) != width or int(t[257][0]) != height or int(t.get(258, (1,))[0]) != bits:
                raise NotImplementedError("TIFF pages of different shape or depth")
            offs, counts = t[273], t.get(279, (plane,))
            raw = np.concatenate([buf[o:o + c] for o, c in zip(offs, counts)]) if len(offs) > 1 else buf[offs[0]:offs[0] + plane]
            if raw.size < plane:
                raise IOError("TIFF data is truncated")
            vol[z] = np.frombuffer(raw[:plane].tobytes(), dtype=dt).reshape(height, width)
    if not dt.isnative:
        vol = vol.astype(dt.newbyteorder("="))
    unit = {1: 1.0, 2: 25400.0, 3: 10000.0}.get(int(t0.get(296, (2,))[0]), 1.0) if "unit" not in ij else 1.0
    vx = unit / t0[282][0] if t0.get(282, (0,))[0] else 1.0
    vy = unit / t0[283][0] if t0.get(283, (0,))[0] else 1.0
    vz = float(ij.get("spacing", 1.0))
    # (z, y, x) C order == (x, y, z) with x fastest: the layout the scan uses as is
    return SpatialImage(np.ascontiguousarray(vol).transpose(2, 1, 0), voxelsize=(vx, vy, vz), info={"tiff_description": desc} if desc else {})


def _imsave_tiff(filename, img):
    """One uncompressed little-endian strip per z plane, ImageJ-style description (spacing, unit) for the voxel size."""
    import struct
    arr = np.asarray(img)
    if arr.ndim != 3 or arr.dtype.kind not in "uif" or arr.dtype.itemsize not in (1, 2, 4, 8):
        raise ValueError("a TIFF stack is a 3-D array of 8 / 16 / 32-bit integers or 32 / 64-bit floats")
    nx, ny, nz = arr.shape
    dt = arr.dtype.newbyteorder("<")
    vox = tuple(float(v) for v in tuple(getattr(img, "voxelsize", (1.0, 1.0, 1.0)))[:3])
    desc = ("ImageJ=1.53\nimages=%d\nslices=%d\nunit=micron\nspacing=%r\n" % (nz, nz, vox[2])).encode("latin-1") + b"\0"
    plane = nx * ny * dt.itemsize
    fmt = {"u": 1, "i": 2, "f": 3}[arr.dtype.kind]
    with open(filename, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, 8))
        pos = 8
        for z in range(nz):
            d = desc if z == 0 else b""
            entries = [(256, 4, 1, nx), (257, 4, 1, ny), (258, 3, 1, 8 * dt.itemsize), (259, 3, 1, 1), (262, 3, 1, 1)]
            extra_off = pos + 2 + 12 * (13 if z == 0 else 12) + 4
            if z == 0:
                entries.append((270, 2, len(d), extra_off))
            res_off = extra_off + len(d)
            data_off = res_off + 16
            entries += [(273, 4, 1, data_off), (277, 3, 1, 1), (278, 4, 1, ny), (279, 4, 1, plane), (282, 5, 1, res_off),
                        (283, 5, 1, res_off + 8), (339, 3, 1, fmt)]
            entries.sort()
            nxt = data_off + plane if z + 1 < nz else 0
            out = struct.pack("<H", len(entries))
            for tag, typ, cnt, val in entries:
                out += struct.pack("<HHI", tag, typ, cnt) + (struct.pack("<HH", val, 0) if typ == 3 and cnt == 1 else struct.pack("<I", val))
            out += struct.pack("<I", nxt) + d
            out += struct.pack("<II", 1000000, max(1, int(round(1000000 * vox[0])))) + struct.pack("<II", 1000000, max(1, int(round(1000000 * vox[1]))))
            assert pos + len(out) == data_off
            f.write(out)
            f.write(np.ascontiguousarray(arr[:, :, z].T.astype(dt, copy=False)).tobytes())
            pos = data_off + plane


"""INRIMAGE-4 reader / writer (tissue_analysis_b200/serial.py): byte layout, round trips, the factory's file input
(SIA:1668-1671).  The reference ships no image file, so these pin the published container layout itself."""
import gzip
import os

import numpy as np
import pytest

from tissue_analysis_b200 import SpatialImage, SpatialImageAnalysis, imread, imsave
from tissue_analysis_b200.synth import tissue_image


@pytest.mark.parametrize("ext", [".inr", ".inr.gz"])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.int16, np.float32])
def test_round_trip(tmp_path, ext, dtype):
    rng = np.random.default_rng(3)
    arr = (rng.random((7, 5, 3)) * 200).astype(dtype)
    img = SpatialImage(arr, voxelsize=(0.2, 0.25, 0.5), info={"TX": "1.5"})
    path = str(tmp_path / ("stack" + ext))
    imsave(path, img)
    back = imread(path)
    assert back.dtype == np.dtype(dtype) and back.shape == (7, 5, 3)
    assert np.array_equal(np.asarray(back), arr)
    assert back.voxelsize == (0.2, 0.25, 0.5)
    assert back.info.get("TX") == "1.5"
    assert back.flags["F_CONTIGUOUS"]                     # x fastest: the layout the scan takes without a copy


def test_byte_layout(tmp_path):
    """Header: magic, KEY=value lines, '##}' terminator, padded to a multiple of 256 bytes; data: x fastest,
    little endian."""
    arr = np.arange(2 * 3 * 4, dtype=np.uint16).reshape(2, 3, 4)          # (x, y, z)
    path = str(tmp_path / "layout.inr")
    imsave(path, SpatialImage(arr, voxelsize=(1.0, 2.0, 3.0)))
    raw = open(path, "rb").read()
    assert raw.startswith(b"#INRIMAGE-4#{\n")
    hlen = raw.index(b"##}\n") + 4
    assert hlen % 256 == 0 and len(raw) == hlen + arr.size * 2
    head = raw[:hlen].decode()
    for line in ("XDIM=2", "YDIM=3", "ZDIM=4", "VDIM=1", "TYPE=unsigned fixed", "PIXSIZE=16 bits", "CPU=decm",
                 "VX=1.0", "VY=2.0", "VZ=3.0"):
        assert ("\n" + line + "\n") in head
    data = np.frombuffer(raw[hlen:], "<u2")
    assert data[1] == arr[1, 0, 0] and data[2] == arr[0, 1, 0] and data[6] == arr[0, 0, 1]


def test_big_endian_and_long_header(tmp_path):
    arr = np.arange(24, dtype=np.uint16).reshape(2, 3, 4)
    head = "#INRIMAGE-4#{\nXDIM=2\nYDIM=3\nZDIM=4\nVDIM=1\nTYPE=unsigned fixed\nPIXSIZE=16 bits\nCPU=sun\n"
    head += "VX=0.5\nVY=0.5\nVZ=2\n" + "".join("#comment line %03d\n" % i for i in range(20))
    head += "\n" * ((-(len(head) + 4)) % 256) + "##}\n"
    assert len(head) == 512
    path = str(tmp_path / "be.inr.gz")
    with gzip.open(path, "wb") as f:
        f.write(head.encode())
        f.write(arr.astype(">u2").tobytes(order="F"))
    back = imread(path)
    assert back.dtype == np.uint16 and np.array_equal(np.asarray(back), arr) and back.voxelsize == (0.5, 0.5, 2.0)


def test_errors(tmp_path):
    p = str(tmp_path / "x.inr")
    open(p, "wb").write(b"not an inrimage" + b"\n" * 300)
    with pytest.raises(IOError):
        imread(p)
    with pytest.raises(NotImplementedError):
        imread(str(tmp_path / "stack.png"))
    imsave(p, np.zeros((4, 4, 4), np.uint16))
    with open(p, "r+b") as f:
        f.truncate(os.path.getsize(p) - 10)
    with pytest.raises(IOError):
        imread(p)


def test_factory_reads_a_file_name(tmp_path):
    """SIA:1668-1671: a string is a file name.  The image a mutator edits in place is the copy-on-write mapping, never
    the file."""
    img = tissue_image((20, 16, 12), 8, seed=2, dome=True, voxelsize=(0.3, 0.3, 1.0))
    path = str(tmp_path / "tissue.inr")
    imsave(path, img)
    sia = SpatialImageAnalysis(path, background=1)
    assert sia.image.shape == (20, 16, 12) and tuple(sia.image.voxelsize) == (0.3, 0.3, 1.0)
    assert np.array_equal(np.asarray(sia.image), np.asarray(img))
    before = open(path, "rb").read()
    np.asarray(sia.image)[0, 0, 0] = 777
    assert open(path, "rb").read() == before


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.int16, np.float32])
def test_tiff_round_trip_and_pillow_agrees(tmp_path, dtype):
    """TIFF stacks (the factory's other input, SIA:1668-1671): our writer -> our reader, our writer -> Pillow (an independent
    TIFF implementation, when it is installed), Pillow's multi-page writer -> our reader."""
    rng = np.random.default_rng(3)
    arr = rng.integers(0, 200, size=(9, 6, 4)).astype(dtype)
    path = str(tmp_path / "stack.tif")
    imsave(path, SpatialImage(arr, voxelsize=(0.25, 0.5, 2.0)))
    back = imread(path)
    assert back.dtype == arr.dtype and np.array_equal(np.asarray(back), arr)
    assert np.asarray(back).flags["F_CONTIGUOUS"]                     # x fastest: what the scan uses as is
    assert np.allclose(back.voxelsize, (0.25, 0.5, 2.0))
    Image = pytest.importorskip("PIL.Image")
    if dtype in (np.uint8, np.uint16, np.float32):                    # the sample types Pillow maps to a mode
        im = Image.open(path)
        assert im.n_frames == 4
        for z in range(4):
            im.seek(z)
            assert np.array_equal(np.array(im), arr[:, :, z].T)
        frames = [Image.fromarray(arr[:, :, z].T.copy()) for z in range(4)]
        other = str(tmp_path / "pillow.tif")
        frames[0].save(other, save_all=True, append_images=frames[1:])
        assert np.array_equal(np.asarray(imread(other)), arr)


def test_tiff_big_endian_imagej_stack_and_refusals(tmp_path):
    """A big-endian file, ImageJ's contiguous stack behind one directory, and what is refused instead of misread."""
    import struct
    arr = np.arange(5 * 4 * 3, dtype=np.uint16).reshape(5, 4, 3) * 7       # (x, y, z)
    nx, ny, nz = arr.shape
    desc = b"ImageJ=1.53\nimages=3\nslices=3\nunit=micron\nspacing=1.5\n\0"

    def one_ifd(bo, compression=1, samples=1):
        ents = [(256, 4, 1, nx), (257, 4, 1, ny), (258, 3, 1, 16), (259, 3, 1, compression), (262, 3, 1, 1),
                (270, 2, len(desc), 0), (273, 4, 1, 0), (277, 3, 1, samples), (278, 4, 1, ny), (279, 4, 1, nx * ny * 2)]
        doff = 8 + 2 + 12 * len(ents) + 4
        out = (b"MM" if bo == ">" else b"II") + struct.pack(bo + "HI", 42, 8) + struct.pack(bo + "H", len(ents))
        for tag, typ, cnt, val in ents:
            val = doff if tag == 270 else doff + len(desc) if tag == 273 else val
            out += struct.pack(bo + "HHI", tag, typ, cnt) + (struct.pack(bo + "HH", val, 0) if typ == 3 else struct.pack(bo + "I", val))
        out += struct.pack(bo + "I", 0) + desc
        return out + np.ascontiguousarray(arr.transpose(2, 1, 0)).astype(bo + "u2").tobytes()

    p = str(tmp_path / "ij.tif")
    open(p, "wb").write(one_ifd(">"))
    img = imread(p)
    assert np.array_equal(np.asarray(img), arr) and img.voxelsize[2] == 1.5
    open(p, "wb").write(one_ifd("<", compression=5))
    with pytest.raises(NotImplementedError):
        imread(p)
    open(p, "wb").write(one_ifd("<", samples=3))
    with pytest.raises(NotImplementedError):
        imread(p)
    open(p, "wb").write(one_ifd("<")[:-20])
    with pytest.raises(IOError):
        imread(p)


def test_factory_reads_a_tiff(tmp_path):
    img = tissue_image((20, 16, 12), 8, seed=2, dome=True, voxelsize=(0.5, 0.5, 2.0))
    path = str(tmp_path / "tissue.tif")
    imsave(path, img)
    sia = SpatialImageAnalysis(path, background=1)
    assert sia.image.shape == (20, 16, 12) and np.allclose(sia.image.voxelsize, (0.5, 0.5, 2.0))
    assert np.array_equal(np.asarray(sia.image), np.asarray(img))


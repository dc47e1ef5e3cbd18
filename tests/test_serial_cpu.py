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
        imread(str(tmp_path / "stack.tif"))
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

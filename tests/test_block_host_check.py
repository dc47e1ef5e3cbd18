"""CPU check of the block-bitmask feature extraction (csrc/ta_block.cuh): groundwork for the next scan kernel, not yet
launched by the product.  tests/host/block_host_check.cu runs the per-block code on the CPU over whole brick tiles
(noise and blob volumes, ragged bricks) and compares every block's label moments / boxes and pair counts with a
brute-force pass over its voxels."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_block_features_on_the_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "block_host_check")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", exe,
                    os.path.join(HERE, "host", "block_host_check.cu")], check=True, capture_output=True, timeout=600)
    for seed in (1, 2):
        r = subprocess.run([exe, str(seed)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert " 0 mismatching blocks" in r.stdout


def test_block_pass_over_whole_volumes_on_the_host(tmp_path):
    """Volumes of several bricks (ragged, slabs with halo planes and a global plane offset): block features + the
    block -> brick -> global transforms + the pair slot conventions + the per-voxel fallback == a direct pass."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "block_volume_check")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", exe,
                    os.path.join(HERE, "host", "block_volume_check.cu")], check=True, capture_output=True, timeout=600)
    r = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatching volumes" in r.stdout


def test_level_formulation_over_whole_volumes_on_the_host(tmp_path):
    """The level formulation (window min / max, fused masks of the known labels, what each level adds, restricted
    per-voxel fallback) for both label widths and 2 .. 5 levels == a direct pass."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "block_level_check")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", exe,
                    os.path.join(HERE, "host", "block_level_check.cu")], check=True, capture_output=True, timeout=600)
    for seed in (1, 2):
        r = subprocess.run([exe, str(seed)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert " 0 mismatching volumes" in r.stdout

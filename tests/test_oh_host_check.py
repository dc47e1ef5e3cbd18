"""CPU check of the scan kernel's one-hot pair arithmetic (`ta::oh_segment_pairs`, csrc/ta_scan.cuh).

The segment arithmetic of the kernel is host-compilable; tests/host/oh_host_check.cu runs it on the CPU over whole
brick tiles (uint16 and uint32, noise and blob volumes, ragged rows, every flag combination) against a brute-force
count of 18-connected wall voxels and +f/+m/+s faces.  No GPU, no oracle import: this pins the bit tricks before the
GPU parity tests pin the whole kernel.
"""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_one_hot_segment_arithmetic_on_the_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "oh_host_check")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", exe,
                    os.path.join(HERE, "host", "oh_host_check.cu")], check=True, capture_output=True, timeout=600)
    for seed in (1, 2, 3):
        r = subprocess.run([exe, str(seed)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "0 mismatches" in r.stdout

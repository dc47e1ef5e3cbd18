"""The scan kernel itself, run on the CPU (tests/host/emu/cuda_emu.h: every CUDA thread is a fiber, warp and block
collectives are rendezvous points, atomics are plain) against a direct pass over the voxels.

Covers the source of the product kernel mk::mask_kernel<T, FLAGS> (run walk, hashed label slots, pair and moment phases,
moment queue, ping-pong tables with the deferred flush, per-voxel path, slab ownership, ragged bricks), alone and behind
the pre-pass of ta_prepass.cuh (classify_cores_kernel, decide_kernel, the work list; background volumes with whole
regions of one-label bricks, a slab whose halo planes belong to no brick), and of round 1's scan_kernel<T, false>, for
uint16 and uint32, on the scalar staging path and on the TMA staging path with the box copy itself emulated (zero fill
outside the buffer, re-clamping of edge tiles; a box origin that is not 16-byte aligned aborts, as the hardware faults on
it).  g++ only.
The GPU parity tests stay the authority for the compiled kernel; this one finds logic errors without a GPU.
"""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_scan_kernels_on_the_cpu_emulation(tmp_path):
    gxx = shutil.which("g++")
    inc = "/usr/local/cuda/include"
    if not gxx or not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    exe = str(tmp_path / "kernel_emu_check")
    subprocess.run([gxx, "-std=c++17", "-O1", "-w", "-I" + inc, "-o", exe, os.path.join(HERE, "host", "kernel_emu_check.cpp")],
                   check=True, capture_output=True, timeout=900)
    for seed in (1, 2):
        r = subprocess.run([exe, str(seed), "36"], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        assert " 0 mismatches" in r.stdout

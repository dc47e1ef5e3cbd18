#!/usr/bin/env python
"""Benchmark of the full feature pass (BASELINE.json metric: Gvoxels/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3] [--cpu-sample EDGE]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full feature pass (moments + bbox, 6-face counts, 18-connected wall-voxel counts, pair-table
compaction + sort, inertia eigen-solve for every label) over one synthetic Voronoi tissue.
  N = 1   workload C3: 1024^3 uint16, 50 000 seeds, dome, background 1 (north_star's target configuration).
  N > 1   the same volume z-slab sharded over N ranks (halo exchange + all_reduce + all_gather inside the
          timed step): total work fixed -> "scaling": "strong".
`value`   device-resident volume, tables left on the device (CUDA events on the launching stream).
`e2e`     host (pinned) volume -> C ABI: H2D copy + pass + D2H of both tables, every step.
`--impl reference`  the CPU restatement of the reference's own per-label loops (oracle/sia_loops.py; the
          reference is Python 2 + openalea and cannot run here) on bounded crops of the same volume, one
          process per host core.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tissue_analysis_b200.synth import CONFIGS  # noqa: E402

METRIC = "full_feature_pass_throughput"
UNIT = "Gvoxel/s"


def ncu_traffic_bytes(config, world):
    """DRAM bytes per scan-kernel launch from the committed `ncu --set full` capture (profiles/), when one exists
    for this workload; None otherwise (never measured live: a number taken under a profiler is not a bench value)."""
    if config != "C3" or world != 1:
        return None
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_scan_c3_metrics.json")))["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _crop_origin(shape_zyx, edge, k):
    """k-th crop of the volume: boxes along the dome surface so background / L1 stages do real work."""
    nz, ny, nx = shape_zyx
    rng = np.random.default_rng(1234 + k)
    z0 = int(rng.integers(0, max(nz - edge, 0) + 1))
    y0 = int(rng.integers(0, max(ny - edge, 0) + 1))
    x0 = 0 if k % 2 == 0 else int(rng.integers(0, max(nx - edge, 0) + 1))
    return z0, y0, x0


def _reference_pass(crop_xyz, voxelsize):
    """The reference's feature pass, per-label loops and all (oracle/sia_loops.py), as graph_from_image drives it
    (temporal_graph_from_image.py:109-212): labels, neighbors, boundingbox, volume, center_of_mass, background
    neighbours / L1, stack margins, inertia_axis, wall_areas, wall voxel counts per pair."""
    import io
    import contextlib
    import warnings
    from oracle.sia_loops import LoopOracle
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        o = LoopOracle(crop_xyz, background=1, voxelsize=voxelsize)
        labels = o.labels()
        nb = o.neighbors()
        o.boundingbox()
        o.volume()
        o.center_of_mass()
        if 1 in o._neighbors and len(nb[1]):
            o.cell_first_layer()
        o.labels_at_stack_margins()
        o.inertia_axis()
        o.wall_areas()
        o.wall_voxels_per_cells_pairs(verbose=False)
    return len(labels)


def _reference_worker(args):
    vol_path, shape_zyx, dtype, edge, k, voxelsize = args
    vol = np.load(vol_path, mmap_mode="r")
    z0, y0, x0 = _crop_origin(shape_zyx, edge, k)
    crop = np.ascontiguousarray(vol[z0:z0 + edge, y0:y0 + edge, x0:x0 + edge]).transpose(2, 1, 0)
    t0 = time.perf_counter()
    nl = _reference_pass(crop, voxelsize)
    return crop.size, time.perf_counter() - t0, nl


def run_reference_sample(vol_zyx, voxelsize, edge, nproc, steps=1, warmup=0):
    """-> (Gvoxel/s, seconds per step, description).  Each step: `nproc` processes, one crop each."""
    import multiprocessing as mp
    import tempfile
    tmp = tempfile.NamedTemporaryFile(suffix=".npy", delete=False)
    tmp.close()
    np.save(tmp.name, vol_zyx)
    ctx = mp.get_context("fork")
    times, vox = [], 0
    try:
        with ctx.Pool(nproc) as pool:
            for it in range(warmup + steps):
                jobs = [(tmp.name, vol_zyx.shape, str(vol_zyx.dtype), edge, it * nproc + k, voxelsize)
                        for k in range(nproc)]
                t0 = time.perf_counter()
                res = pool.map(_reference_worker, jobs)
                dt = time.perf_counter() - t0
                if it >= warmup:
                    times.append(dt)
                    vox = sum(r[0] for r in res)
    finally:
        os.unlink(tmp.name)
    sec = float(np.mean(times))
    desc = ("%d crops of %d^3 voxels of the workload volume per step (one per process, %d processes), full feature "
            "pass by oracle/sia_loops.py (py3 restatement of the reference's scipy.ndimage per-label loops)"
            % (nproc, edge, nproc))
    return vox / sec / 1e9, sec, desc


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="edge of the CPU baseline crops (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--equal-planes", action="store_true", help="N > 1: slabs of equal height instead of equal work")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.config]
    X, Y, Z = cfg["shape"]
    shape_zyx = (Z, Y, X)
    nvox = X * Y * Z
    elem = 2 if cfg["dtype"] == "uint16" else 4
    ncores = os.cpu_count() or 1
    workload = "%s: %dx%dx%d %s Voronoi tissue, %d seeds, %s, seed %d" % (
        args.config, X, Y, Z, cfg["dtype"], cfg["ncell"], "dome + background 1" if cfg["dome"] else "no background",
        cfg["seed"])

    import torch
    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample: the whole run (steps + warmup) should end within a few minutes at ~0.3 Mvoxel/s/process
        edge = args.cpu_sample or (128 if args.steps + args.warmup <= 16 else 96 if args.steps + args.warmup <= 40 else 64)
        if torch.cuda.is_available():
            from tissue_analysis_b200.synth import voronoi_device
            torch.cuda.set_device(0)
            vol = voronoi_device(shape_zyx, cfg["ncell"], cfg["seed"], cfg["weights"][::-1], cfg["dome"],
                                 cfg["dtype"]).cpu().numpy()
        else:   # no GPU: a reduced stand-in volume generated on the CPU (same generator definition)
            from tissue_analysis_b200.synth import voronoi_numpy
            shape_zyx = tuple(min(s, 160) for s in shape_zyx)
            vol = voronoi_numpy(shape_zyx, max(cfg["ncell"] * int(np.prod(shape_zyx)) // nvox, 8), cfg["seed"],
                                cfg["weights"][::-1], cfg["dome"], np.dtype(cfg["dtype"]))
        gv, sec, desc = run_reference_sample(vol, cfg["voxelsize"], edge, ncores, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": gv, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u16" if elem == 2 else "u32",
                "data": "synthetic", "config": {"workload": workload, "sample": desc},
                "cpu_baseline": {"value": gv, "unit": UNIT, "cores": ncores, "kind": "port", "sample": desc},
                "e2e": {"value": gv, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------------------------------------- our arm
    assert torch.cuda.is_available(), "bench.py --impl ours needs a B200 (there is no CPU fallback)"
    import torch.distributed as dist
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.distributed import SlabScan
    from tissue_analysis_b200.synth import voronoi_device
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert args.gpus == world or world == 1, "launch N>1 with torch.distributed.run"

    bounds = None
    if world > 1 and cfg["dome"] and not args.equal_planes:
        # Partition step (outside the timed region, as a loader would do it once per volume): plane boundaries of equal
        # estimated work instead of equal height -- a dome leaves the end slabs mostly background.
        from tissue_analysis_b200.distributed import partition_planes, partition_planes_weighted, plane_work_weights
        b0 = partition_planes(shape_zyx[0], world)
        part = voronoi_device(shape_zyx, cfg["ncell"], cfg["seed"], cfg["weights"][::-1], cfg["dome"], cfg["dtype"],
                              zslice=(b0[rank], b0[rank + 1]))
        w = plane_work_weights(part, 1)
        del part
        cap = max(b0[r + 1] - b0[r] for r in range(world))
        mine = torch.zeros(cap, dtype=torch.float64, device="cuda")
        mine[:w.numel()] = w
        allw = torch.empty(world * cap, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allw, mine)
        allw = allw.cpu().numpy()
        weights = np.concatenate([allw[r * cap:r * cap + (b0[r + 1] - b0[r])] for r in range(world)])
        bounds = partition_planes_weighted(weights, world)
    scan = SlabScan(shape_zyx, torch.uint16 if elem == 2 else torch.uint32, rank=rank, world=world, bounds=bounds)
    gen = voronoi_device(shape_zyx, cfg["ncell"], cfg["seed"], cfg["weights"][::-1], cfg["dome"], cfg["dtype"],
                         zslice=(scan.g_lo, scan.g_hi))
    scan.owned().copy_(gen)
    del gen
    torch.cuda.synchronize()
    hint_labels = cfg["ncell"] + 1 if elem == 4 else 0

    def step():
        scan.run(flags=_native.PASS_ALL, max_label_hint=hint_labels, inertia=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    scan.stage_ms.clear()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = scan.ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms = []
    ev0.record()
    for _ in range(args.steps):
        step()
        scan_ms.append(scan.ctx.last_timing()["scan_ms"])
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = scan.ctx.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, float(np.mean(scan_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_rank = torch.zeros(world, dtype=torch.float64, device="cuda")
    per_rank[rank] = float(np.mean(scan_ms))
    if world > 1:
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    scan_ms_per_rank = [round(float(v), 4) for v in per_rank.tolist()]
    ms, scan_ms_avg = float(t[0]), float(t[1])
    value = nvox / (ms * 1e-3) / 1e9

    peak, peak_src = measured_peak_gbs()
    own_vox = (scan.g_hi - scan.g_lo) * Y * X
    achieved = own_vox * elem / (scan_ms_avg * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_bytes(args.config, world), "kernel": "ta::scan_kernel", "kernel_ms": scan_ms_avg,
                "kernel_ms_per_rank": scan_ms_per_rank, "algorithmic_bytes_per_voxel": elem, "peak_source": peak_src}

    # ---- e2e: host volume through the C ABI (H2D + pass + D2H of the tables), rank-local slab --------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty(tuple(scan.buf.shape), dtype=scan.buf.dtype).pin_memory()
        host.copy_(scan.buf)
        harr = host.numpy()
        ctx2 = _native.Context(local_rank)
        ns_b, nm_b, nf_b = harr.shape

        def e2e_step():
            ctx2.run_pass_host(harr, _native.PASS_ALL, hint_labels,
                               slab=(scan.own_lo, scan.own_hi, scan.g_lo - scan.own_lo))
            lt = ctx2.label_table()
            pt = ctx2.pair_table()
            return lt, pt

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(n_e2e):
            lt, pt = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        d2h = sum(a.nbytes for a in lt) + sum(a.nbytes for a in pt)
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": nvox / float(tt[0]) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(harr.nbytes),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(tt[0]) * 1e3,
               "note": "pinned host volume -> ta_run_pass_host (chunked H2D overlapped with the scan) + "
                       "ta_fetch_*_table (D2H); per-rank slab, no cross-rank merge"}
        ctx2.close()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        edge = args.cpu_sample or 128
        vol = scan.buf.cpu().numpy()
        gv, sec, desc = run_reference_sample(vol, cfg["voxelsize"], edge, ncores, steps=1, warmup=0)
        cpu_baseline = {"value": gv, "unit": UNIT, "cores": ncores, "kind": "port", "sample": desc,
                        "seconds": sec}

    if rank == 0 and scan.stage_ms:
        sys.stderr.write("[stage ms per step] " + ", ".join("%s %.3f" % (k, v / args.steps)
                                                            for k, v in scan.stage_ms.items()) + "\n")
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u16" if elem == 2 else "u32",
                "data": "synthetic",
                "config": {"workload": workload, "sharding": "z-slabs x%d%s" % (world, (", plane boundaries of equal estimated work %s" % list(bounds)) if bounds else ""),
                           "l2": "input (%.1f GiB) larger than L2, no flush needed" % (nvox * elem / 2 ** 30),
                           "step": "halo exchange + scan + table compaction/sort + cross-rank merge + inertia eig"},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Benchmark of the full feature pass (BASELINE.json metric: Gvoxels/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3] [--cpu-sample EDGE]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full feature pass (moments + bbox, 6-face counts, 18-connected wall-voxel counts, pair-table
compaction, inertia eigen-solve for every label) over one synthetic Voronoi tissue.
  N = 1   workload C3: 1024^3 uint16, 50 000 seeds, dome, background 1 (north_star's target configuration).
  N > 1   the same volume z-slab sharded over N ranks (halo exchange + all_reduce + all_gather + merge inside the
          timed step): total work fixed -> "scaling": "strong".  Outside the timed region rank 0 scans the whole volume
          alone and compares a digest of the merged tables with it ("parity"); a mismatch ends the run with exit code 3.
`value`   device-resident volume, tables left on the device (CUDA events on the launching stream).
`e2e`     host (pinned) volume -> tables in host memory, every step: N = 1 through ta_run_pass_host (chunked H2D
          overlapped with the scan) + ta_fetch_*; N > 1 every rank uploads its slab, the sharded step runs including the
          cross-rank merge, rank 0 fetches the merged tables.
`api_e2e` (N = 1) what a user of the drop-in class waits for: pageable numpy volume -> SpatialImageAnalysis3D ->
          graph_arrays / the reference's dict-returning methods, wall clock.
`--impl reference`  the reference's OWN SpatialImageAnalysis3D (oracle/_ref/vplants_ref, the mechanical py3 rewrite that
          oracle/make_ref.py makes of /root/reference; oracle/sia_loops.py if that module was not built) on bounded
          crops of the same workload, one process per host core.  The crops are generated on the CPU: this arm loads
          neither the product library nor CUDA.
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
    """DRAM bytes per scan (its three launches: the two pre-pass kernels and the mask kernel) from the committed
    `ncu --set full` capture (profiles/), when one exists
    for this workload; None otherwise (never measured live: a number taken under a profiler is not a bench value)."""
    if config != "C3" or world != 1:
        return None
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_scan_c3_metrics.json")))["dram_bytes_per_launch"])
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


def _reference_factory():
    """-> (make(crop_xyz, voxelsize) -> analysis object, kind).  The reference itself when oracle/_ref holds it."""
    try:
        from oracle import make_ref, ref_stubs
        ref = make_ref.load()
    except Exception:
        ref = None
    if ref is not None:
        def make(crop, voxelsize):
            return ref.SpatialImageAnalysis3D(ref_stubs.SpatialImage(crop, voxelsize=voxelsize), background=1)
        return make, "reference"
    from oracle.sia_loops import LoopOracle

    def make(crop, voxelsize):
        return LoopOracle(crop, background=1, voxelsize=voxelsize)
    return make, "port"


def _reference_pass(crop_xyz, voxelsize):
    """The reference's feature pass, per-label loops and all, as graph_from_image drives it
    (temporal_graph_from_image.py:109-212): labels, neighbors, boundingbox, volume, center_of_mass, background
    neighbours / L1, stack margins, inertia_axis, wall_areas, wall voxels per pair."""
    import io
    import contextlib
    import warnings
    make, _ = _reference_factory()
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        o = make(crop_xyz, voxelsize)
        labels = o.labels()
        if len(labels) == 0:              # a crop of pure background: the reference's inertia_axis indexes an empty list
            return 0
        nb = o.neighbors()
        o.boundingbox()
        o.volume()
        o.center_of_mass()
        if 1 in nb and len(nb[1]):
            o.cell_first_layer()
        o.labels_at_stack_margins()
        o.inertia_axis()
        o.wall_areas()
        o.wall_voxels_per_cells_pairs(verbose=False)
    return len(labels)


def _reference_worker(args):
    cfg_name, vol_path, edge, k = args
    cfg = CONFIGS[cfg_name]
    X, Y, Z = cfg["shape"]
    shape_zyx = (Z, Y, X)
    z0, y0, x0 = _crop_origin(shape_zyx, edge, k)
    e = [min(edge, s) for s in shape_zyx]
    if vol_path:
        vol = np.load(vol_path, mmap_mode="r")
        crop = np.ascontiguousarray(vol[z0:z0 + e[0], y0:y0 + e[1], x0:x0 + e[2]])
    else:
        # CPU generator, this box only (bit-identical to the device generator: tests/test_gpu_parity.py)
        from tissue_analysis_b200.synth import voronoi_numpy_box
        crop = voronoi_numpy_box(shape_zyx, cfg["ncell"], cfg["seed"], (z0, y0, x0), (z0 + e[0], y0 + e[1], x0 + e[2]),
                                 cfg["weights"][::-1], cfg["dome"], np.dtype(cfg["dtype"]))
    crop = crop.transpose(2, 1, 0)
    t0 = time.perf_counter()
    nl = _reference_pass(crop, cfg["voxelsize"])
    return crop.size, time.perf_counter() - t0, nl


def run_reference_sample(cfg_name, edge, nproc, steps=1, warmup=0, vol_zyx=None):
    """-> (Gvoxel/s, seconds per step, description, kind).  Each step: `nproc` processes, one crop each; the timed part of
    a worker is the feature pass alone (crop generation / loading is outside it)."""
    import multiprocessing as mp
    import tempfile
    path = None
    if vol_zyx is not None:
        tmp = tempfile.NamedTemporaryFile(suffix=".npy", delete=False)
        tmp.close()
        np.save(tmp.name, vol_zyx)
        path = tmp.name
    kind = _reference_factory()[1]
    ctx = mp.get_context("fork")
    rates = []
    try:
        with ctx.Pool(nproc) as pool:
            for it in range(warmup + steps):
                jobs = [(cfg_name, path, edge, it * nproc + k) for k in range(nproc)]
                res = pool.map(_reference_worker, jobs)
                if it >= warmup:
                    # the processes run side by side: the step takes as long as its slowest pass
                    rates.append((sum(r[0] for r in res), max(r[1] for r in res)))
    finally:
        if path:
            os.unlink(path)
    vox = float(np.mean([r[0] for r in rates]))
    sec = float(np.mean([r[1] for r in rates]))
    desc = ("%d crops of %d^3 voxels of the workload volume per step (one per process, %d processes), full feature pass by %s"
            % (nproc, edge, nproc, "the reference's own SpatialImageAnalysis3D (oracle/_ref, py3 rewrite by oracle/make_ref.py)"
               if kind == "reference" else "oracle/sia_loops.py (py3 restatement of the reference's per-label loops)"))
    return vox / sec / 1e9, sec, desc, kind


# ------------------------------------------------------------------------------------------------ main
def tables_digest(count, s1, s2, bbox, lo, hi, faces, wall):
    """Order-independent fingerprint of a pair of result tables (the pair rows are sorted by (lo, hi) already)."""
    import hashlib
    h = hashlib.sha256()
    for a in (count, s1, s2, bbox, lo, hi, faces, wall):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


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
    ap.add_argument("--no-extra", action="store_true", help="skip api_e2e and the C2 / C4 scan times (N = 1)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the merged == single-GPU check")
    ap.add_argument("--equal-planes", action="store_true", help="N > 1: slabs of equal height instead of equal work")
    ap.add_argument("--overlap", action="store_true", help="N > 1: scan the interior planes during the halo exchange")
    ap.add_argument("--sync-merge", action="store_true", help="N > 1: the synchronous merge in every step")
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

    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample: the whole run (steps + warmup) should end within a few minutes at ~0.3 Mvoxel/s/process.  No CUDA,
        # no product library in this process: the crops come from the CPU generator.
        edge = args.cpu_sample or (128 if args.steps + args.warmup <= 16 else 96 if args.steps + args.warmup <= 40 else 64)
        gv, sec, desc, kind = run_reference_sample(args.config, edge, ncores, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": gv, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u16" if elem == 2 else "u32",
                "data": "synthetic", "config": {"workload": workload, "sample": desc},
                "cpu_baseline": {"value": gv, "unit": UNIT, "cores": ncores, "kind": kind, "sample": desc},
                "e2e": {"value": gv, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------------------------------------- our arm
    import torch
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
        # estimated work instead of equal height -- a dome leaves the end slabs mostly background.  Multiples of 8 planes:
        # the scan works in bricks of 8.
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
        bounds = partition_planes_weighted(weights, world, align=8)
    scan = SlabScan(shape_zyx, torch.uint16 if elem == 2 else torch.uint32, rank=rank, world=world, bounds=bounds)
    gen = voronoi_device(shape_zyx, cfg["ncell"], cfg["seed"], cfg["weights"][::-1], cfg["dome"], cfg["dtype"],
                         zslice=(scan.g_lo, scan.g_hi))
    scan.owned().copy_(gen)
    del gen
    torch.cuda.synchronize()
    hint_labels = cfg["ncell"] + 1 if elem == 4 else 0

    def step():
        scan.run(flags=_native.PASS_ALL, max_label_hint=hint_labels, inertia=True, overlap=args.overlap,
                 deferred=not args.sync_merge)

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
    ev0.record(scan.stream)                    # the stream the steps are queued on
    for _ in range(args.steps):
        step()
    ev1.record(scan.stream)
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = scan.ctx.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    # the scan kernel's own time: CUDA events of the library around its launches, read back after each of a few extra
    # steps (reading them inside the timed loop would put a host synchronisation into every step)
    scan_ms = []
    for _ in range(3):
        step()
        scan_ms.append(scan.ctx.last_timing()["scan_ms"])
    barrier()
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
                "traffic": ncu_traffic_bytes(args.config, world), "kernel": "ta::mk::mask_kernel (with its pre-pass ta::pp::classify_cores_kernel + decide_kernel: the scan's launches)",
                "kernel_ms": scan_ms_avg,
                "kernel_ms_per_rank": scan_ms_per_rank, "algorithmic_bytes_per_voxel": elem, "peak_source": peak_src}

    # ---- parity of the sharded result: merged tables == one GPU scanning the whole volume (outside the timed region) ------
    parity = None
    if world > 1 and not args.no_parity:
        step()
        if rank == 0:
            merged = tables_digest(*scan.ctx.label_table(), *scan.ctx.pair_table())
            n_pairs = int(scan.ctx.pair_table()[0].size)
            whole = voronoi_device(shape_zyx, cfg["ncell"], cfg["seed"], cfg["weights"][::-1], cfg["dome"], cfg["dtype"])
            c1 = _native.Context(local_rank)
            c1.bind_device(whole.data_ptr(), elem, X, Y, Z, keepalive=whole)
            c1.run_pass(_native.PASS_ALL, hint_labels)
            single = tables_digest(*c1.label_table(), *c1.pair_table())
            c1.close()
            del whole
            parity = {"merged_equals_single": merged == single, "pairs": n_pairs, "sha256": merged[:16],
                      "check": "rank 0 scans the whole volume alone; sha256 over the label table and the sorted pair table"}
        barrier()

    # ---- e2e: host volume -> tables in host memory ---------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty(tuple(scan.owned().shape), dtype=scan.buf.dtype).pin_memory()
        host.copy_(scan.owned())
        n_e2e = max(2, min(args.steps, 5))
        if world == 1:
            harr = host.numpy()
            ctx2 = _native.Context(local_rank)

            def e2e_step():
                ctx2.run_pass_host(harr, _native.PASS_ALL, hint_labels)
                return ctx2.label_table(), ctx2.pair_table()
            note = "pinned host volume -> ta_run_pass_host (chunked H2D overlapped with the scan) + ta_fetch_*_table (D2H)"
        else:
            def e2e_step():
                with torch.cuda.stream(scan.stream):
                    scan.owned().copy_(host, non_blocking=True)
                step()
                if rank == 0:
                    return scan.ctx.label_table(), scan.ctx.pair_table()
                return None, None
            note = ("every rank: pinned host slab -> device (H2D), sharded step with halo exchange and cross-rank merge; "
                    "rank 0: ta_fetch_*_table of the merged tables (D2H)")
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            lt, pt = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        d2h = (sum(a.nbytes for a in lt) + sum(a.nbytes for a in pt)) if lt is not None else 0
        tt = torch.tensor([dt, float(host.numel() * host.element_size()), float(d2h)], dtype=torch.float64, device="cuda")
        if world > 1:
            mx = tt[:1].clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            tt[0] = mx[0]
        e2e = {"value": nvox / float(tt[0]) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(tt[1]),
               "d2h_bytes_per_step": int(tt[2]), "ms_per_step": float(tt[0]) * 1e3, "note": note}
        if world == 1:
            ctx2.close()

    # ---- N = 1 extras: the drop-in class end to end, the other single-GPU configurations ------------------------------------
    api_e2e, extra = None, None
    if world == 1 and not args.no_extra:
        import warnings
        from tissue_analysis_b200 import SpatialImage, SpatialImageAnalysis3D
        from tissue_analysis_b200.temporal_graph_from_image import graph_arrays
        vol_np = scan.buf.cpu().numpy()                  # pageable host memory, (z, y, x)
        img = SpatialImage(vol_np.transpose(2, 1, 0), voxelsize=cfg["voxelsize"])     # x fastest, as openalea's images
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t_runs = []
            for _ in range(2):       # the first run also pays the context's allocations (pinned staging, tables)
                t0 = time.perf_counter()
                an = SpatialImageAnalysis3D(img, background=1, device=local_rank)
                ga = graph_arrays(an)
                t_runs.append(time.perf_counter() - t0)
            t_arrays = min(t_runs)
            t0 = time.perf_counter()
            nl = an.nb_labels()
            an.volume(); an.boundingbox(); an.center_of_mass(); an.neighbors(); an.wall_areas()
            an.cell_first_layer(); an.labels_at_stack_margins(); an.inertia_axis()
            t_dicts = time.perf_counter() - t0
        api_e2e = {"graph_arrays_s": t_arrays, "graph_arrays_s_runs": t_runs, "value": nvox / t_arrays / 1e9, "unit": UNIT,
                   "dict_methods_s": t_dicts, "labels": int(nl),
                   "note": "pageable numpy volume -> SpatialImageAnalysis3D (H2D + scan + D2H) -> graph_arrays (CSR + property "
                           "arrays); dict_methods_s: then volume, boundingbox, center_of_mass, neighbors, wall_areas, "
                           "cell_first_layer, labels_at_stack_margins, inertia_axis as the reference's dicts (tables cached)"}
        del an, ga, img, vol_np
        extra = {}
        for name in ("C2", "C4"):
            if name == args.config:
                continue
            c = CONFIGS[name]
            cx, cy, cz = c["shape"]
            cel = 2 if c["dtype"] == "uint16" else 4
            free = torch.cuda.mem_get_info()[0]
            if cx * cy * cz * cel * 1.3 > free:
                extra[name] = "skipped: not enough free memory"
                continue
            v = voronoi_device((cz, cy, cx), c["ncell"], c["seed"], c["weights"][::-1], c["dome"], c["dtype"])
            cx2 = _native.Context(local_rank)
            cx2.bind_device(v.data_ptr(), cel, cx, cy, cz, keepalive=v)
            tms = []
            for _ in range(3):
                cx2.run_pass(_native.PASS_ALL, c["ncell"] + 1 if cel == 4 else 0)
                tms.append(cx2.last_timing()["scan_ms"])
            extra[name] = {"scan_ms": float(np.min(tms)), "Gvoxel_per_s": cx * cy * cz / float(np.min(tms)) / 1e6,
                           "roofline_frac": cx * cy * cz * cel / (float(np.min(tms)) * 1e-3) / 1e9 / peak}
            cx2.close()
            del v

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        edge = args.cpu_sample or 128
        gv, sec, desc, kind = run_reference_sample(args.config, edge, ncores, steps=1, warmup=0,
                                                   vol_zyx=scan.buf.cpu().numpy())
        cpu_baseline = {"value": gv, "unit": UNIT, "cores": ncores, "kind": kind, "sample": desc, "seconds": sec}

    if rank == 0 and scan.stage_ms:
        sys.stderr.write("[stage ms per step] " + ", ".join("%s %.3f" % (k, v / args.steps)
                                                            for k, v in scan.stage_ms.items()) + "\n")
    rc = 0
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u16" if elem == 2 else "u32",
                "data": "synthetic",
                "config": {"workload": workload, "sharding": "z-slabs x%d%s" % (world, (", plane boundaries of equal estimated work %s" % list(bounds)) if bounds else ""),
                           "l2": "input (%.1f GiB) larger than L2, no flush needed" % (nvox * elem / 2 ** 30),
                           "step": "halo exchange + scan + record packing + cross-rank merge + inertia eig" if world > 1
                                   else "scan + table compaction / sort + inertia eig",
                           "merge": None if world == 1 else ("synchronous" if args.sync_merge else "deferred: no host synchronisation inside a step")},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks}
        if parity is not None:
            line["parity"] = parity
            if not parity["merged_equals_single"]:
                rc = 3
        if api_e2e is not None:
            line["api_e2e"] = api_e2e
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    if world > 1:
        flag = torch.tensor([rc], dtype=torch.int32, device="cuda")
        dist.broadcast(flag, 0)
        rc = int(flag.item())
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())

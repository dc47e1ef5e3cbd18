"""Config C5 at size: a time series of 1024^3 uint16 frames (seeds 10, 11, ...), one frame per GPU at a time.

    python tools/timeseries_bench.py [--frames 10] [--shape 1024]                         # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/timeseries_bench.py

Replicas only: rank r analyses frames r, r + world, ... (tissue_analysis_b200.timeseries.analyze_frames); no data-path
collective.  Setup (untimed): every rank generates its frames on the device and parks them in pinned host memory -- the
place a time series comes from.  Timed: host frame -> H2D -> scan -> tables in host memory, for all frames of the rank;
barrier on both sides, wall clock, max over ranks.  Prints ONE JSON line (frames/s and Gvoxel/s over all ranks) and checks
one frame per rank against a fresh single pass (digest), and rank 0's first frame's 64-plane slab against the C oracle.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tissue_analysis_b200 import _native  # noqa: E402
from tissue_analysis_b200.synth import voronoi_device  # noqa: E402
from tissue_analysis_b200.timeseries import analyze_frames, frames_of_rank  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--shape", type=int, default=1024)
    ap.add_argument("--ncell", type=int, default=50000)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.shape
    ncell = a.ncell if n == 1024 else max(8, a.ncell * n ** 3 // 1024 ** 3)
    mine = frames_of_rank(a.frames, rank, world)
    frames = [None] * a.frames
    for k in mine:
        dev = voronoi_device((n, n, n), ncell, 10 + k, (1, 1, 1), True, "uint16")
        host = torch.empty(dev.shape, dtype=dev.dtype).pin_memory()
        host.copy_(dev)
        frames[k] = host.numpy()
        del dev
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    analyze_frames(frames[:max(mine) + 1] if mine else [], device=local, rank=rank, world=world)      # warm-up (contexts, pinned staging)
    barrier()
    t0 = time.perf_counter()
    out = analyze_frames(frames, device=local, rank=rank, world=world)
    barrier()
    dt = time.perf_counter() - t0
    # parity: the first frame of every rank against a fresh pass over the same data on the device
    ok = True
    if mine:
        k = mine[0]
        dev = torch.from_numpy(frames[k]).cuda()
        c = _native.Context(local)
        c.bind_device(dev.data_ptr(), 2, n, n, n, keepalive=dev)
        c.run_pass()
        cnt, s1, s2, bb = c.label_table()
        lo, hi, f, w = c.pair_table()
        c.close()
        t = out[k]
        ok = (int(t.count.sum()) == n ** 3 and np.array_equal(np.sort(t.count[t.count > 0]), np.sort(cnt[cnt > 0]))
              and np.array_equal(t.pair_lo, lo) and np.array_equal(t.pair_hi, hi) and np.array_equal(t.wall18, w)
              and int(t.faces.sum()) == int(f.sum()))
        if rank == 0:
            from oracle import c_onepass
            z0 = n // 2 - 32
            sub = np.ascontiguousarray(frames[k][z0:z0 + 64])
            c = _native.Context(local)
            c.bind_host(sub)
            c.run_pass()
            ref = c_onepass.onepass(sub, nrows=65536)
            got = dict(zip(("count", "s1", "s2", "bbox"), c.label_table()))
            got.update(dict(zip(("lo", "hi", "faces", "wall18"), c.pair_table())))
            c.close()
            ok = ok and all(np.array_equal(got[x], ref[x]) for x in ("count", "s1", "s2", "lo", "hi", "faces", "wall18"))
    stat = torch.tensor([dt, 1.0 if ok else 0.0, float(len(out))], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stat[:1].clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = stat[1:2].clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(stat, op=dist.ReduceOp.SUM)
        stat[0], stat[1] = mx[0], mn[0]
    if rank == 0:
        sec, nframes = float(stat[0]), int(stat[2])
        print(json.dumps({"metric": "time_series_feature_pass", "config": "C5: %d frames of %d^3 uint16, %d seeds, dome, one frame per GPU at a time"
                          % (a.frames, n, ncell), "n_gpus": world, "frames": nframes, "seconds": sec,
                          "frames_per_s": nframes / sec, "value": nframes * n ** 3 / sec / 1e9, "unit": "Gvoxel/s",
                          "scaling": "replicas only (no data-path collective)", "parity": bool(stat[1] > 0.5),
                          "timed": "pinned host frame -> ta_run_pass_host (H2D overlapped with the scan) -> tables in host memory, "
                                   "two contexts per rank side by side; waves: %d" % -(-a.frames // world)}))
    if world > 1:
        dist.destroy_process_group()
    return 0 if stat[1] > 0.5 else 3


if __name__ == "__main__":
    sys.exit(main())

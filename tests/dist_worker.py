"""torchrun worker for tests/test_gpu_distributed.py: z-slab sharded scan over NCCL == single-GPU scan."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tissue_analysis_b200 import _native  # noqa: E402
from tissue_analysis_b200.distributed import SlabScan  # noqa: E402
from tissue_analysis_b200.synth import voronoi_device  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for shape, ncell, dt, seed in (((77, 96, 160), 300, "uint16", 5), ((64, 40, 72), 120, "uint32", 6)):
        tdt = torch.uint16 if dt == "uint16" else torch.uint32
        # unequal slabs (one of them a single plane when world >= 4) for the first volume, equal planes for the second
        bounds = None
        if seed == 5:
            bounds = {2: [0, 20, 77], 4: [0, 10, 30, 31, 77]}.get(world)
        scan = SlabScan(shape, tdt, rank=rank, world=world, bounds=bounds)
        scan.owned().copy_(voronoi_device(shape, ncell, seed, (1, 1, 1), True, dt, zslice=(scan.g_lo, scan.g_hi)))
        torch.cuda.synchronize()
        scan.run(inertia=True)
        merged = scan.tables()
        first = (scan.ctx.label_table(), scan.ctx.pair_table())
        # the following steps run deferred (no host synchronisation; record counts stay on the device) -- with the scan of
        # the interior planes overlapped with the halo exchange, and without: the same tables as the first, synchronous step
        for overlap in (True, False):
            assert scan._rec_cap > 0
            scan.run(inertia=True, overlap=overlap, deferred=True)
            again = (scan.ctx.label_table(), scan.ctx.pair_table())
            ok = ok and all(np.array_equal(a, b) for a, b in zip(first[0] + first[1], again[0] + again[1]))
        # single-GPU truth on every rank
        whole = voronoi_device(shape, ncell, seed, (1, 1, 1), True, dt)
        ctx = _native.Context(local)
        ctx.bind_device(whole.data_ptr(), whole.element_size(), shape[2], shape[1], shape[0], keepalive=whole)
        ctx.run_pass()
        c, s1, s2, bb = ctx.label_table()
        lo, hi, f, w = ctx.pair_table()
        c2, s12, s22, bb2 = scan.ctx.label_table()
        lo2, hi2, f2, w2 = scan.ctx.pair_table()
        n = min(c.size, c2.size)
        present = c[:n] > 0
        same = (np.array_equal(c[:n], c2[:n]) and np.array_equal(s1[:n], s12[:n]) and np.array_equal(s2[:n], s22[:n])
                and np.array_equal(bb[:n][present], bb2[:n][present]) and np.array_equal(lo, lo2)
                and np.array_equal(hi, hi2) and np.array_equal(f, f2) and np.array_equal(w, w2))
        ok = ok and same and merged.count.sum() == int(np.prod(shape))
        ctx.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_OK" if int(flag) == 1 else "DIST_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()

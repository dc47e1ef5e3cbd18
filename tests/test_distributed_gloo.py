"""world_size-2 gloo run (CPU) of the multi-GPU host logic: plane partition, halo exchange, label-table
all_reduce, pair-record all_gather.  Each rank's local tables come from the numpy oracle restricted to its slab
with the ownership rules of ta_set_slab; the merged result must equal the unsplit tables."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sia_onepass
from tissue_analysis_b200 import distributed as D
from tissue_analysis_b200.synth import voronoi_numpy


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, vol_zyx, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ns = vol_zyx.shape[0]
        b = D.partition_planes(ns, world)
        g_lo, g_hi = b[rank], b[rank + 1]
        has_lo, has_hi = rank > 0, rank < world - 1
        own_lo = 1 if has_lo else 0
        own_hi = own_lo + g_hi - g_lo
        buf = torch.zeros((own_hi + (1 if has_hi else 0),) + vol_zyx.shape[1:], dtype=torch.uint16)
        buf[own_lo:own_hi] = torch.from_numpy(vol_zyx[g_lo:g_hi].copy())
        D.exchange_halo_planes(buf, own_lo, own_hi, rank, world)
        lo_g = g_lo - own_lo
        assert np.array_equal(buf.numpy(), vol_zyx[lo_g:lo_g + buf.shape[0]])        # halos are the neighbours' planes
        # local oracle tables on (x=fast,...) API order: oracle slab axis is the LAST axis -> transpose
        img = buf.numpy().transpose(2, 1, 0)
        L = 4096
        lt = sia_onepass.label_table(img, nlabels=L, slab=(own_lo, own_hi))
        pt = sia_onepass.pair_table(img, slab=(own_lo, own_hi))
        lt["s1"][:, 2] += lt["count"] * lo_g                                        # local -> global slow index
        # (second moments involving z are not needed for this host-logic test)
        present = lt["count"] > 0
        lt["bmin"][present, 2] += lo_g
        lt["bmax"][present, 2] += lo_g
        count = torch.from_numpy(lt["count"])
        s1 = torch.from_numpy(lt["s1"].ravel().copy())
        s2 = torch.from_numpy(lt["s2"].ravel().copy())
        bmin = torch.from_numpy(np.minimum(lt["bmin"], 2 ** 31 - 1).astype(np.int32).ravel())
        bmax = torch.from_numpy(lt["bmax"].astype(np.int32).ravel())
        D.allreduce_label_tables(count, s1, s2, bmin, bmax)
        rec = np.concatenate([pt["lo"][:, None], pt["hi"][:, None], pt["faces"], pt["wall18"][:, None]], axis=1)
        allrec = D.allgather_pair_records(torch.from_numpy(rec.astype(np.int32)), world).numpy().astype(np.int64)
        merged = sia_onepass.merge_pair_tables([dict(lo=allrec[:, 0], hi=allrec[:, 1], faces=allrec[:, 2:8],
                                                     wall18=allrec[:, 8])])
        if rank == 0:
            out.put(dict(count=count.numpy(), s1=s1.numpy().reshape(-1, 3), bmin=bmin.numpy().reshape(-1, 3),
                         bmax=bmax.numpy().reshape(-1, 3), pairs=merged))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_slab_merge_equals_unsplit():
    vol = voronoi_numpy((23, 20, 26), 40, seed=8, dome=True)       # (z, y, x)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, vol, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    img = vol.transpose(2, 1, 0)
    lt = sia_onepass.label_table(img, nlabels=4096)
    pt = sia_onepass.pair_table(img)
    assert np.array_equal(got["count"], lt["count"])
    assert np.array_equal(got["s1"], lt["s1"])
    present = lt["count"] > 0
    assert np.array_equal(got["bmin"][present], lt["bmin"][present])
    assert np.array_equal(got["bmax"][present], lt["bmax"][present])
    for k in ("lo", "hi", "faces", "wall18"):
        assert np.array_equal(got["pairs"][k], pt[k]), k


def test_partition_covers_all_planes():
    for ns in (1, 7, 64, 1024, 1000):
        for w in (1, 2, 3, 4, 8):
            b = D.partition_planes(ns, w)
            assert b[0] == 0 and b[-1] == ns and all(b[i] <= b[i + 1] for i in range(w))


def test_weighted_partition_balances_a_dome():
    """Plane boundaries of equal estimated work: every rank gets at least one plane, the boundaries are monotone, and
    the heaviest slab of a dome-like weight profile is much lighter than with equal heights."""
    z = np.arange(256)
    w = np.maximum(0.0, 1.0 - ((z - 140.0) / 100.0) ** 2) * 1000 + 12.0      # tissue in the middle, background elsewhere
    for world in (2, 3, 4, 8):
        b = D.partition_planes_weighted(w, world)
        assert b[0] == 0 and b[-1] == 256 and all(b[i] < b[i + 1] for i in range(world))
        heavy = max(w[b[r]:b[r + 1]].sum() for r in range(world))
        e = D.partition_planes(256, world)
        heavy_equal = max(w[e[r]:e[r + 1]].sum() for r in range(world))
        assert heavy <= heavy_equal
        assert heavy <= w.sum() / world * 1.08
    assert D.partition_planes_weighted(np.zeros(10), 4) == D.partition_planes(10, 4)
    assert D.partition_planes_weighted([5, 0, 0, 0], 4) == [0, 1, 2, 3, 4]

"""z-slab sharding of one label volume over the GPUs of a node (one process per GPU, torch.distributed).

The reference is a single Python process (no parallelism of any kind: SURVEY.md section 2); this is the multi-GPU
form of the same pass.  Every output of the scan is a commutative reduction, so a rank needs only its planes
plus one halo plane on each side:

  0. partition         contiguous plane ranges, equal planes or equal estimated work (``partition_planes_weighted``)
  1. halo exchange     boundary planes to both neighbours (NCCL send/recv over NVLink; 2 MiB per face at C3), on a
                       side stream, overlapped with the scan of the interior planes (``ta_run_pass_ranges``)
  2. local pass        ownership rules of ``ta_set_slab``: a face belongs to the rank owning its lower voxel, an
                       18-connected wall voxel to the rank owning the voxel
  3. label table       all_reduce SUM of the exact u64 sums, MIN / MAX of the bounding boxes
  4. pair table        all_gather of the packed records and a device hash sum-merge on every rank.  First step: sizes
                       first (one host synchronisation), ``ta_merge_pair_records``.  Steady state (``run`` again on a
                       volume of the same kind): NO host synchronisation anywhere in the step -- the pass is queued with
                       ``TA_PASS_DEFERRED``, every rank contributes a fixed-size record buffer whose header row carries
                       its count, ``ta_merge_pair_records_deferred`` reads the counts on the device, and sorting, the
                       record count and the overflow flags wait for the first fetch.

The collective helpers below act on plain tensors, so the world_size-2 ``gloo`` tests drive them on the CPU with
oracle tables; on GPUs they act directly on the library's device buffers (zero copy).
"""
import numpy as np
import torch
import torch.distributed as dist

REC_WORDS = 9


def partition_planes(n_slow, world):
    """Contiguous, balanced plane ranges: rank r owns [b[r], b[r+1])."""
    return [(n_slow * r) // world for r in range(world + 1)]


def partition_planes_weighted(weights, world, align=1):
    """Contiguous plane ranges of (nearly) equal total weight: rank r owns [b[r], b[r+1]).  ``weights``: one
    non-negative work estimate per plane (see ``plane_work_weights``).  Every rank gets at least one plane.
    ``align``: boundaries are multiples of it (the scan works in bricks of 8 planes: a slab whose height is a multiple
    of 8 has no ragged last brick layer)."""
    if align > 1 and len(weights) % align == 0 and len(weights) // align >= world:
        w = np.asarray(weights, dtype=np.float64).reshape(-1, align).sum(axis=1)
        return [align * v for v in partition_planes_weighted(w, world)]
    w = np.asarray(weights, dtype=np.float64)
    n = w.size
    assert n >= world
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    if total <= 0:
        return partition_planes(n, world)
    b = [0]
    for r in range(1, world):
        cut = int(np.searchsorted(cum, total * r / world, side="left"))
        # the plane boundary nearest to the ideal cumulative weight
        if cut > 0 and abs(cum[cut - 1] - total * r / world) <= abs(cum[min(cut, n)] - total * r / world):
            cut -= 1
        cut = min(max(cut, b[-1] + 1), n - (world - r))
        b.append(cut)
    b.append(n)
    return b


BACKGROUND_PLANE_WEIGHT = 0.12     # a voxel of a one-label brick costs about this much of a tissue voxel


def plane_work_weights(slab, background):
    """Work estimate per plane of a (planes, mid, fast) label tensor: tissue voxels + a small share for background."""
    tissue = (slab != background).sum(dim=(1, 2)).to(torch.float64)
    per_plane = float(slab.shape[1] * slab.shape[2])
    return tissue + BACKGROUND_PLANE_WEIGHT * (per_plane - tissue)


def exchange_halo_planes(buf, own_lo, own_hi, rank, world):
    """``buf`` holds [halo?][owned planes][halo?] along dim 0.  Sends the first / last owned plane to the lower /
    upper neighbour and receives their boundary planes into the halo slots."""
    if world == 1:
        return
    ops, keep = [], []
    raw = buf.view(torch.uint8)
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, raw[own_lo], rank - 1))
        ops.append(dist.P2POp(dist.irecv, raw[own_lo - 1], rank - 1))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, raw[own_hi - 1], rank + 1))
        ops.append(dist.P2POp(dist.irecv, raw[own_hi], rank + 1))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    del keep


def allreduce_label_tables(count, s1, s2, bmin, bmax):
    """In place: sums are exact integers (two's complement SUM is u64 addition), boxes are MIN / MAX."""
    dist.all_reduce(count, op=dist.ReduceOp.SUM)
    dist.all_reduce(s1, op=dist.ReduceOp.SUM)
    dist.all_reduce(s2, op=dist.ReduceOp.SUM)
    dist.all_reduce(bmin, op=dist.ReduceOp.MIN)
    dist.all_reduce(bmax, op=dist.ReduceOp.MAX)


def allgather_pair_records(records, world, sizes_out=None):
    """records: int32[n, 9] (this rank's packed pair rows) -> int32[sum n_r, 9] on every rank.  ``sizes_out``: a list that
    receives the ranks' row counts."""
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=records.device)
    sizes = torch.zeros(world, dtype=torch.int64, device=records.device)
    dist.all_gather_into_tensor(sizes, n)
    sizes = sizes.tolist()                       # one host synchronisation for all ranks' sizes
    if sizes_out is not None:
        sizes_out[:] = sizes
    cap = max(max(sizes), 1)
    padded = torch.zeros((cap, REC_WORDS), dtype=records.dtype, device=records.device)
    padded[:records.shape[0]] = records
    gathered = torch.empty((world * cap, REC_WORDS), dtype=records.dtype, device=records.device)
    dist.all_gather_into_tensor(gathered, padded)             # rank r's rows land at [r * cap, (r + 1) * cap)
    return torch.cat([gathered[r * cap:r * cap + s] for r, s in enumerate(sizes)], dim=0)


class _DeviceView(object):
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3)


def device_tensor(ptr, shape, typestr):
    """Zero-copy torch view of a library-owned device buffer."""
    return torch.as_tensor(_DeviceView(ptr, shape, typestr), device="cuda")


class SlabScan(object):
    """One rank of the z-slab sharded scan.  ``global_shape`` is (slow, mid, fast)."""

    def __init__(self, global_shape, dtype=torch.uint16, rank=None, world=None, device=None, bounds=None):
        """``bounds``: world + 1 plane boundaries (e.g. from ``partition_planes_weighted``); default: equal planes."""
        from . import _native
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.global_shape = tuple(int(v) for v in global_shape)
        ns, nm, nf = self.global_shape
        b = partition_planes(ns, self.world) if bounds is None else [int(v) for v in bounds]
        assert len(b) == self.world + 1 and b[0] == 0 and b[-1] == ns and all(b[i] < b[i + 1] for i in range(self.world))
        self.bounds = b
        self.g_lo, self.g_hi = b[self.rank], b[self.rank + 1]
        self.has_lo, self.has_hi = self.rank > 0, self.rank < self.world - 1
        self.own_lo = 1 if self.has_lo else 0
        self.own_hi = self.own_lo + (self.g_hi - self.g_lo)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.buf = torch.zeros((self.own_hi + (1 if self.has_hi else 0), nm, nf), dtype=dtype, device=self.device)
        self.elem = self.buf.element_size()
        self.ctx = _native.Context(self.device.index)
        # One stream for the library's kernels AND the collectives (torch orders a collective against the stream that is
        # current when it is issued).  torch's default stream has handle 0, which ta_set_stream reads as "the context's own
        # stream": the deferred step, which never synchronises with the host, would then race with the collectives.
        self.stream = torch.cuda.Stream(self.device)
        self.ctx.set_stream(self.stream.cuda_stream)
        self._native = _native
        self.stage_ms, self._t0 = {}, 0.0
        self._halo_stream = self._halo_event = None
        self._rec_cap = 0          # rows of the deferred record buffer: set by the first (synchronous) step
        self._sizes_box = []
        self._gathered = None

    def owned(self):
        """View of the owned planes (fill it by upload or by the device generator)."""
        return self.buf[self.own_lo:self.own_hi]

    def _tick(self, name):
        """Stage timing for profiling (TA_DIST_TIMING=1): host clock after a device synchronisation."""
        import os
        import time
        if not os.environ.get("TA_DIST_TIMING"):
            return
        torch.cuda.synchronize(self.device)
        now = time.perf_counter()
        if name is not None:
            self.stage_ms[name] = self.stage_ms.get(name, 0.0) + (now - self._t0) * 1e3
        self._t0 = now

    def run(self, flags=7, max_label_hint=0, pair_capacity_hint=0, inertia=False, overlap=False, deferred=True):
        """One sharded step.  ``overlap``: scan the interior planes while the halo planes are still in flight (three
        launches) instead of one launch behind the exchange.  ``deferred``: after a first synchronous step has sized the
        record buffer, run without any host synchronisation (see the module docstring).  Everything is queued on
        ``self.stream``."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self._run(flags, max_label_hint, pair_capacity_hint, inertia, overlap, deferred)

    def _run(self, flags, max_label_hint, pair_capacity_hint, inertia, overlap, deferred):
        ns, nm, nf = self.buf.shape
        self._tick(None)
        use_deferred = bool(deferred and self.world > 1 and self._rec_cap)
        if use_deferred:
            flags |= self._native.PASS_DEFERRED
            pair_capacity_hint = self._rec_cap
        if self.elem == 4 and not max_label_hint and self.world > 1:
            # every rank must size its dense label table identically before the all_reduce
            mx = self.owned().view(torch.int32).max().to(torch.int64).reshape(1)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            max_label_hint = int(mx.item())
            if max_label_hint < 0:
                raise ValueError("labels >= 2**31 are not supported in the sharded path")
        self.ctx.bind_device(self.buf.data_ptr(), self.elem, nf, nm, ns, keepalive=self.buf)
        self.ctx.set_slab(self.own_lo, self.own_hi, self.g_lo - self.own_lo)
        if self.world == 1:
            self.ctx.run_pass(flags, max_label_hint, pair_capacity_hint)
        else:
            # The halo exchange runs on a side stream while the pass stream already scans the interior planes (they need
            # no halo); the first and the last owned plane follow behind the exchange's event.  No host synchronisation.
            main = torch.cuda.current_stream(self.device)
            if self._halo_stream is None:
                self._halo_stream = torch.cuda.Stream(self.device)
                self._halo_event = torch.cuda.Event()
            self._halo_stream.wait_stream(main)
            with torch.cuda.stream(self._halo_stream):
                exchange_halo_planes(self.buf, self.own_lo, self.own_hi, self.rank, self.world)
                self._halo_event.record(self._halo_stream)
            lo, hi = self.own_lo, self.own_hi
            ev = self._halo_event.cuda_event
            # the local records only feed the merge, which sorts: skip the local sort
            if overlap:
                first = (lo, min(lo + 1, hi)) if self.has_lo else (lo, lo)
                last = (max(hi - 1, first[1]), hi) if self.has_hi else (hi, hi)
                self.ctx.run_pass_ranges([(first[1], last[0]), first, last], [None, ev, ev],
                                         flags | self._native.PASS_UNSORTED, max_label_hint, pair_capacity_hint)
            else:
                # one launch behind the exchange: a boundary plane scanned on its own stages a whole ten-plane tile per
                # brick for one plane of work, and two extra launches cost more than 2 MiB over NVLink
                self.ctx.run_pass_ranges([(lo, hi)], [ev], flags | self._native.PASS_UNSORTED, max_label_hint,
                                         pair_capacity_hint)
        self._tick("pass")
        if self.world > 1:
            if use_deferred:
                self.merge_deferred()
            else:
                self.merge()
        if inertia:
            self.ctx.inertia_table(fetch=False)
        self._tick("inertia")

    def merge(self):
        (p_count, p_s1, p_s2, p_bmin, p_bmax), n = self.ctx.label_table_device()
        if p_s1 == p_count + 8 * n and p_s2 == p_s1 + 24 * n:
            # the library keeps count | s1 | s2 in one block: one SUM collective for all ten u64 columns
            dist.all_reduce(device_tensor(p_count, (n * 10,), "<i8"), op=dist.ReduceOp.SUM)
            if p_bmax == p_bmin + 12 * n:
                # bmin | bmax are one block too: MAX(x) = -MIN(-x), one collective for both halves of the boxes
                box = device_tensor(p_bmin, (n * 6,), "<i4")
                box[n * 3:].neg_()
                dist.all_reduce(box, op=dist.ReduceOp.MIN)
                box[n * 3:].neg_()
            else:
                dist.all_reduce(device_tensor(p_bmin, (n * 3,), "<i4"), op=dist.ReduceOp.MIN)
                dist.all_reduce(device_tensor(p_bmax, (n * 3,), "<i4"), op=dist.ReduceOp.MAX)
        else:
            allreduce_label_tables(device_tensor(p_count, (n,), "<i8"), device_tensor(p_s1, (n * 3,), "<i8"),
                                   device_tensor(p_s2, (n * 6,), "<i8"), device_tensor(p_bmin, (n * 3,), "<i4"),
                                   device_tensor(p_bmax, (n * 3,), "<i4"))
        self._tick("allreduce")
        p_rec, n_rec = self.ctx.pair_records_device()
        if n_rec:
            mine = device_tensor(p_rec, (n_rec, REC_WORDS), "<i4")
        else:
            mine = torch.zeros((0, REC_WORDS), dtype=torch.int32, device=self.device)
        allrec = allgather_pair_records(mine, self.world, sizes_out=self._sizes_box).contiguous()
        # size the deferred record buffer of the following steps: the largest rank, with head room
        self._rec_cap = max(1024, int(1.5 * max(self._sizes_box)) + 64)
        self._tick("allgather")
        self.ctx.merge_pair_records(allrec.data_ptr(), allrec.shape[0])     # same stream; synchronises internally
        self._tick("pair merge")

    def _allreduce_labels(self):
        (p_count, p_s1, p_s2, p_bmin, p_bmax), n = self.ctx.label_table_device()
        assert p_s1 == p_count + 8 * n and p_s2 == p_s1 + 24 * n and p_bmax == p_bmin + 12 * n
        dist.all_reduce(device_tensor(p_count, (n * 10,), "<i8"), op=dist.ReduceOp.SUM)
        box = device_tensor(p_bmin, (n * 6,), "<i4")
        box[n * 3:].neg_()                       # MAX(x) = -MIN(-x): one collective for both halves of the boxes
        dist.all_reduce(box, op=dist.ReduceOp.MIN)
        box[n * 3:].neg_()

    def merge_deferred(self):
        """The merge of a PASS_DEFERRED step: two all_reduces, one fixed-size all_gather, one merge kernel; the host
        never waits."""
        self._allreduce_labels()
        self._tick("allreduce")
        p_rec, cap = self.ctx.pair_records_deferred()
        rows = (cap + 1) * REC_WORDS
        mine = device_tensor(p_rec, (rows,), "<i4")
        if self._gathered is None or self._gathered.numel() != self.world * rows:
            self._gathered = torch.empty(self.world * rows, dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(self._gathered, mine)
        self._tick("allgather")
        self.ctx.merge_pair_records_deferred(self._gathered.data_ptr(), cap, self.world)
        self._tick("pair merge")

    def tables(self, ax_of_mem=(2, 1, 0)):
        """Merged tables as ScanTables (API shape = global (slow, mid, fast) unless ``ax_of_mem`` says otherwise)."""
        from .engine import tables_from_memory_order
        count, s1, s2, bbox = self.ctx.label_table()
        lo, hi, faces, wall = self.ctx.pair_table()
        ns, nm, nf = self.global_shape
        mem_dims = (nf, nm, ns)
        shape_api = [0, 0, 0]
        for k in range(3):
            shape_api[ax_of_mem[k]] = mem_dims[k]
        return tables_from_memory_order(tuple(shape_api), ax_of_mem, count, s1, s2, bbox, lo, hi, faces, wall)

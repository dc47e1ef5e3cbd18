"""`-m gpu` parity: the CUDA scan (through the C ABI and the Python mirror) against the CPU oracles.

Integer outputs bit-exact; centre of mass and wall areas bit-exact floats; inertia eigen-data rtol 1e-6
(eigenvectors up to sign, skipped inside degenerate eigenspaces) -- the tolerances north_star states.
"""
import warnings

import numpy as np
import pytest

from oracle import sia_onepass
from oracle.sia_loops import LoopOracle
from tests.helpers import TOY, compare_api, oracle_tables
from tissue_analysis_b200 import SpatialImage, SpatialImageAnalysis3D
from tissue_analysis_b200.synth import tissue_image

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore", category=UserWarning)
warnings.filterwarnings("ignore", category=RuntimeWarning)


def assert_tables_equal(t, o):
    n = min(t.nrows, o.nrows)
    assert not t.count[n:].any() and not o.count[n:].any()
    assert np.array_equal(t.count[:n], o.count[:n])
    assert np.array_equal(t.s1[:n], o.s1[:n])
    assert np.array_equal(t.s2[:n], o.s2[:n])
    present = o.count[:n] > 0
    assert np.array_equal(t.bmin[:n][present], o.bmin[:n][present])
    assert np.array_equal(t.bmax[:n][present], o.bmax[:n][present])
    assert np.array_equal(t.pair_lo, o.pair_lo) and np.array_equal(t.pair_hi, o.pair_hi)
    assert np.array_equal(t.faces, o.faces)
    assert np.array_equal(t.wall18, o.wall18)


def both(img, **kw):
    prod = SpatialImageAnalysis3D(img, **kw)
    orc = LoopOracle(np.asarray(img), voxelsize=getattr(img, "voxelsize", None), **kw)
    return prod, orc


def test_docstring_image():
    prod, orc = both(TOY.copy())
    compare_api(prod, orc, eig=False)
    assert prod.cell_wall_area(7, [2, 5]) == {(2, 7): 1.0, (5, 7): 2.0}


def test_voronoi_dome_full_api():
    img = tissue_image((40, 36, 30), 40, seed=11, dome=True)
    prod, orc = both(img, background=1)
    compare_api(prod, orc)


def test_anisotropic_c_order_label_zero():
    rng = np.random.default_rng(3)
    arr = np.ascontiguousarray(np.asarray(tissue_image((36, 30, 24), 30, seed=5, weights=(2, 2, 5), dome=True)))
    arr[rng.random(arr.shape) < 0.02] = 0
    im = SpatialImage(arr, voxelsize=(0.2, 0.2, 0.5))
    prod, orc = both(im, background=1, ignoredlabels=0)
    compare_api(prod, orc)


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 7, 3), (9, 1, 5), (17, 16, 8), (129, 17, 9), (8, 33, 19),
                                   (131, 5, 3), (64, 16, 8), (256, 32, 16)])
@pytest.mark.parametrize("dtype", [np.uint16, np.uint32])
def test_random_noise_tables(shape, dtype):
    """Ragged shapes (not multiples of the brick / vector width) and many-label junctions."""
    rng = np.random.default_rng(sum(shape))
    arr = rng.integers(0, 23, size=shape[::-1]).astype(dtype)      # (z, y, x) C-order == x fastest
    img = SpatialImage(arr.transpose(2, 1, 0))
    prod = SpatialImageAnalysis3D(img, background=1)
    assert_tables_equal(prod._tables(), oracle_tables(np.asarray(img)))


def test_blocky_volume_tables_and_large_u32_labels():
    rng = np.random.default_rng(7)
    small = rng.integers(2, 400, size=(9, 11, 40))
    arr = np.kron(small, np.ones((6, 5, 7), np.int64)).astype(np.uint32)
    arr[arr == 17] = 3_000_000        # sparse huge label in a uint32 volume
    prod = SpatialImageAnalysis3D(SpatialImage(arr), background=2)
    assert_tables_equal(prod._tables(), oracle_tables(arr))


def test_c1_config_tables_and_features():
    """BASELINE config C1: 128^3 uint16, 500 cells (tables exact; API against the loop oracle)."""
    img = tissue_image((128, 128, 128), 500, seed=0)
    prod, orc = both(img)
    assert_tables_equal(prod._tables(), oracle_tables(np.asarray(img)))
    compare_api(prod, orc, check_wall_voxels=False, real_modes=(True,))


def test_slab_split_equals_whole():
    """Two z-slab 'ranks' on one GPU (ta_set_slab) + host merge == unsplit tables."""
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.engine import memory_layout, tables_from_memory_order
    img = tissue_image((48, 40, 37), 60, seed=4, dome=True)
    view, ax = memory_layout(img)
    ns = view.shape[0]
    cut = 19
    parts = []
    for lo, hi in ((0, cut), (cut, ns)):
        b0, b1 = max(lo - 1, 0), min(hi + 1, ns)
        ctx = _native.Context()
        ctx.bind_host(np.ascontiguousarray(view[b0:b1]))
        ctx.set_slab(lo - b0, hi - b0, b0)
        ctx.run_pass()
        parts.append((ctx.label_table(), ctx.pair_table()))
        ctx.close()
    (c0, s10, s20, bb0), (lo0, hi0, f0, w0) = parts[0]
    (c1, s11, s21, bb1), (lo1, hi1, f1, w1) = parts[1]
    bbox = np.concatenate([np.minimum(bb0[:, :3], bb1[:, :3]), np.maximum(bb0[:, 3:], bb1[:, 3:])], axis=1)
    merged = sia_onepass.merge_pair_tables([dict(lo=lo0, hi=hi0, faces=f0.astype(np.int64), wall18=w0.astype(np.int64)),
                                            dict(lo=lo1, hi=hi1, faces=f1.astype(np.int64), wall18=w1.astype(np.int64))])
    t = tables_from_memory_order(img.shape, ax, c0 + c1, s10 + s11, s20 + s21, bbox, merged["lo"], merged["hi"],
                                 merged["faces"], merged["wall18"])
    assert_tables_equal(t, oracle_tables(np.asarray(img)))


def test_native_inertia_eig_against_lapack():
    from tissue_analysis_b200 import _native
    rng = np.random.default_rng(0)
    a = rng.normal(size=(500, 3, 3))
    cov = a @ a.transpose(0, 2, 1)
    cov[:20] = np.diag([3.0, 1.0, 0.0])            # already diagonal, zero eigenvalue
    cov6 = np.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], axis=1)
    ctx = _native.Context()
    evals, evecs = ctx.inertia_eig(cov6)
    ctx.close()
    w = np.linalg.eigvalsh(cov)[:, ::-1]
    np.testing.assert_allclose(evals, w, rtol=1e-10, atol=1e-12)
    recon = np.einsum("nij,ni,nik->njk", evecs, evals, evecs)
    np.testing.assert_allclose(recon, cov, rtol=1e-9, atol=1e-10)


def test_errors_are_loud():
    from tissue_analysis_b200 import _native
    ctx = _native.Context()
    with pytest.raises(_native.NativeError):
        ctx.run_pass()                               # no volume bound
    arr = np.zeros((4, 4, 4), np.uint16)
    ctx.bind_host(arr)
    with pytest.raises(_native.NativeError):
        ctx.label_table()                            # no pass yet
    ctx.close()


def test_device_generator_equals_numpy_generator():
    from tissue_analysis_b200.synth import voronoi_device, voronoi_numpy
    for shape, ncell, w, dome, dt in (((40, 48, 56), 90, (1, 1, 1), True, "uint16"),
                                      ((33, 20, 70), 300, (5, 2, 2), False, "uint32"),
                                      ((16, 16, 16), 3, (1, 1, 1), True, "uint16")):
        dev = voronoi_device(shape, ncell, 7, w, dome, dt).cpu().numpy()
        ref = voronoi_numpy(shape, ncell, 7, w, dome, np.dtype(dt), k=16)
        assert np.array_equal(dev, ref), (shape, ncell)
    # slab generation == the same planes of the whole volume
    whole = voronoi_device((40, 48, 56), 90, 7, (1, 1, 1), True, "uint16").cpu().numpy()
    part = voronoi_device((40, 48, 56), 90, 7, (1, 1, 1), True, "uint16", zslice=(13, 29)).cpu().numpy()
    assert np.array_equal(part, whole[13:29])


def test_time_series_frames_reuse_one_context():
    """Config C5 in miniature: independent frames through one context; each frame equals its own oracle tables."""
    from tissue_analysis_b200.timeseries import analyze_frames, frames_of_rank
    frames = [tissue_image((40, 32, 24 + 3 * k), 30 + 5 * k, seed=10 + k, dome=True) for k in range(3)]
    got = analyze_frames(frames)
    assert sorted(got) == [0, 1, 2]
    for k, f in enumerate(frames):
        assert_tables_equal(got[k], oracle_tables(np.asarray(f)))
    assert frames_of_rank(10, rank=3, world=8) == [3] and frames_of_rank(10, rank=1, world=8) == [1, 9]


def test_c2_config_tables_against_c_oracle_and_sampled_api():
    """BASELINE config C2: 512x512x256 uint16, 5000 seeds, voxelsize (0.2, 0.2, 0.5): tables bit-exact against the C
    oracle, API features (incl. wall areas and inertia axes) against the loop oracle on a sample of labels."""
    import torch
    from oracle import c_onepass
    from tissue_analysis_b200.synth import CONFIGS, voronoi_device
    cfg = CONFIGS["C2"]
    X, Y, Z = cfg["shape"]
    zyx = voronoi_device((Z, Y, X), cfg["ncell"], cfg["seed"], cfg["weights"][::-1], cfg["dome"], cfg["dtype"]).cpu().numpy()
    img = SpatialImage(zyx.transpose(2, 1, 0), voxelsize=cfg["voxelsize"])          # (x, y, z), x fastest
    prod = SpatialImageAnalysis3D(img, background=1)
    t = prod._tables()
    ref = c_onepass.onepass(zyx, nrows=65536)                                          # memory order == (x, y, z) here
    assert np.array_equal(t.count, ref["count"].astype(np.int64))
    assert np.array_equal(t.s1, ref["s1"].astype(np.int64)) and np.array_equal(t.s2, ref["s2"].astype(np.int64))
    assert np.array_equal(t.pair_lo, ref["lo"]) and np.array_equal(t.pair_hi, ref["hi"])
    assert np.array_equal(t.faces, ref["faces"].astype(np.int64)) and np.array_equal(t.wall18, ref["wall18"].astype(np.int64))
    # sampled API parity against the per-label loops of the reference restatement
    orc = LoopOracle(np.asarray(img), voxelsize=cfg["voxelsize"], background=1)
    rng = np.random.default_rng(0)
    labels = sorted(rng.choice(prod.labels(), size=12, replace=False).tolist())
    assert prod.volume(list(labels)) == orc.volume(list(labels))
    cp, co = prod.center_of_mass(list(labels)), orc.center_of_mass(list(labels))
    for l in labels:
        assert np.array_equal(np.asarray(cp[l]), np.asarray(co[l]))
        assert prod.boundingbox(l) == orc.boundingbox(l)
        nb = sorted(map(int, orc.neighbors(l)))
        assert sorted(prod.neighbors(l)) == nb
        wa_o = orc.cell_wall_area(l, nb)
        assert prod.cell_wall_area(l, nb) == dict(((int(a), int(b)), v) for (a, b), v in wa_o.items())
    (vp, wp), (vo, wo) = prod.inertia_axis(list(labels)), orc.inertia_axis(list(labels))
    from tests.helpers import assert_eig_close
    for l in labels:
        assert_eig_close(vp[l], wp[l], vo[l], np.real(wo[l]))


def test_pair_table_overflow_is_detected_and_retried():
    """A pair table that is too small must never drop pairs: the C ABI reports TA_ERR_PAIR_OVERFLOW and the wrapper
    redoes the pass with a larger table."""
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.engine import tables_from_memory_order
    rng = np.random.default_rng(5)
    arr = rng.integers(0, 300, size=(24, 40, 64)).astype(np.uint16)          # ~45 000 distinct touching pairs
    ctx = _native.Context()
    ctx.bind_host(arr)
    rc = ctx.lib.ta_run_pass(ctx.h, _native.PASS_ALL, 0, 300)                 # capacity for ~300 pairs only
    assert rc == _native.TA_ERR_PAIR_OVERFLOW
    with pytest.raises(_native.NativeError):
        ctx.pair_table()                                                      # tables are invalid after an overflow
    ctx.run_pass(pair_capacity_hint=300)                                      # wrapper: grow and redo
    count, s1, s2, bbox = ctx.label_table()
    lo, hi, faces, wall = ctx.pair_table()
    ctx.close()
    t = tables_from_memory_order(arr.shape, (2, 1, 0), count, s1, s2, bbox, lo, hi, faces, wall)
    assert lo.size > 20000
    assert_tables_equal(t, oracle_tables(arr))


def test_label_beyond_hint_is_an_error():
    from tissue_analysis_b200 import _native
    arr = np.full((8, 8, 16), 7, np.uint32)
    arr[3, 3, 3] = 100000
    ctx = _native.Context()
    ctx.bind_host(arr)
    assert ctx.lib.ta_run_pass(ctx.h, _native.PASS_ALL, 50, 0) == _native.TA_ERR_LABEL_RANGE
    ctx.run_pass()                                                            # hint 0: the library finds the maximum
    assert ctx.label_table()[0][100000] == 1
    ctx.close()


def test_graph_from_image_on_the_gpu():
    """the graph packing over the CUDA tables == the restated reference graph builder."""
    from oracle.graph_loops import graph_from_image_oracle
    from tests.test_graph_cpu import PROPS, compare_graph
    from tissue_analysis_b200.temporal_graph_from_image import graph_arrays, graph_from_image
    img = tissue_image((72, 64, 48), 80, seed=41, dome=True, voxelsize=(0.4, 0.4, 1.0), weights=(2, 2, 5))
    g = graph_from_image(img, spatio_temporal_properties=PROPS, ignore_cells_at_stack_margins=False)
    o = graph_from_image_oracle(np.asarray(img), properties=PROPS, voxelsize=img.voxelsize,
                                ignore_cells_at_stack_margins=False)
    compare_graph(g, o)
    arr = graph_arrays(SpatialImageAnalysis3D(img, ignoredlabels=0, background=1), ignore_cells_at_stack_margins=False)
    assert sorted(g.vertices()) == arr.labels.tolist() and g.nb_edges() == arr.edge_lo.size


def test_mutators_and_property_image_on_the_gpu():
    """LUT gather and in-place relabel kernels (ta_map_labels)."""
    from tests.test_mutators_cpu import reference_mutation
    from tissue_analysis_b200.property_spatial_image import create_property_image
    for order in ("C", "F"):
        img = tissue_image((52, 44, 36), 40, seed=43, dome=True)
        arr = np.array(np.asarray(img), order=order)
        work = SpatialImage(arr.copy(order=order), voxelsize=img.voxelsize)
        prod = SpatialImageAnalysis3D(work, background=1)
        labels = sorted(prod.labels())
        vol = prod.volume(real=False)
        some = dict((l, v) for l, v in vol.items() if l % 3 == 0)
        out = create_property_image(prod, some, dtype=np.uint16)
        expect = np.full(arr.shape, 1, np.uint16)
        for l, v in some.items():
            expect[arr == l] = np.float64(v).astype(np.uint16)
        assert np.array_equal(np.asarray(out), expect)
        fuse, remove = labels[1:4], labels[6:9]
        prod.fuse_labels_in_image(list(fuse), verbose=False)
        prod.remove_labels_from_image(list(remove), verbose=False)
        mutated = reference_mutation(arr, fuse=fuse, remove=remove)
        assert np.array_equal(np.asarray(work), mutated)
        fresh = LoopOracle(mutated, voxelsize=img.voxelsize, background=1, ignoredlabels=[0])
        compare_api(prod, fresh, check_wall_voxels=False, real_modes=(True,))
        with pytest.raises(ValueError):                          # a label the voxel type cannot hold must not wrap silently
            prod.remove_labels_from_image([sorted(prod.labels())[0]], erase_value=70000, verbose=False)


@pytest.mark.parametrize("shape", [(5, 3, 2), (17, 16, 8), (129, 17, 9), (131, 5, 3), (260, 40, 21)])
@pytest.mark.parametrize("dtype,nlab", [(np.uint16, 9), (np.uint16, 16), (np.uint16, 23), (np.uint32, 30), (np.uint32, 40)])
def test_noise_with_tens_of_labels(shape, dtype, nlab):
    """Noise volumes: junctions of many labels in every neighbourhood, ragged rows, sparse label values."""
    rng = np.random.default_rng(sum(shape) + nlab)
    names = rng.choice(np.arange(2, 60000), size=nlab, replace=False)
    arr = names[rng.integers(0, nlab, size=shape[::-1])].astype(dtype)
    img = SpatialImage(arr.transpose(2, 1, 0))
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.engine import memory_layout, tables_from_memory_order
    view, ax = memory_layout(img)
    ctx = _native.Context()
    ctx.bind_host(np.ascontiguousarray(view))
    ctx.run_pass(_native.PASS_ALL)
    count, s1, s2, bbox = ctx.label_table()
    lo, hi, faces, wall = ctx.pair_table()
    ctx.close()
    t = tables_from_memory_order(img.shape, ax, count, s1, s2, bbox, lo, hi, faces, wall)
    assert_tables_equal(t, oracle_tables(np.asarray(img)))


@pytest.mark.parametrize("dtype", [np.uint16, np.uint32])
def test_deferred_pass_and_merge_equal_the_synchronous_pass(dtype):
    """TA_PASS_DEFERRED (the sharded driver's steady state: no host synchronisation, record counts stay on the device):
    the packed record buffer, merged back as if it had been gathered from `world` ranks, gives the tables of the plain
    pass -- once (world = 1) and doubled (the same buffer twice: every pair counter exactly twice as large)."""
    import torch
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.engine import memory_layout
    img = tissue_image((150, 70, 33), 260, seed=21, dome=True, dtype=dtype)
    view = np.ascontiguousarray(memory_layout(img)[0])
    hint = 400 if dtype == np.uint32 else 0
    ctx = _native.Context()
    stream = torch.cuda.Stream()                # as the sharded driver does: one (non-default) stream shared with torch
    ctx.set_stream(stream.cuda_stream)
    torch.cuda.set_stream(stream)
    ctx.bind_host(view)
    ctx.run_pass(_native.PASS_ALL, hint)
    want = ctx.label_table() + ctx.pair_table()
    npairs = want[4].size
    for world in (1, 2):
        ctx.run_pass(_native.PASS_ALL | _native.PASS_UNSORTED | _native.PASS_DEFERRED, hint, npairs + 100)
        ptr, cap = ctx.pair_records_deferred()
        assert cap == npairs + 100
        rows = (cap + 1) * 9
        from tissue_analysis_b200.distributed import device_tensor
        mine = device_tensor(ptr, (rows,), "<i4")
        assert int(mine[0]) == npairs                      # the header row carries the count
        gathered = torch.cat([mine] * world).contiguous()
        ctx.merge_pair_records_deferred(gathered.data_ptr(), cap, world)
        got = ctx.label_table() + ctx.pair_table()
        for a, b in zip(got[:4], want[:4]):
            assert np.array_equal(a, b)
        assert np.array_equal(got[4], want[4]) and np.array_equal(got[5], want[5])
        assert np.array_equal(got[6], world * want[6]) and np.array_equal(got[7], world * want[7])
    # too few rows for the records: the error arrives with the first fetch
    ctx.run_pass(_native.PASS_ALL | _native.PASS_DEFERRED, hint, max(npairs // 2, 1))
    with pytest.raises(_native.NativeError):
        ctx.pair_table()
    ctx.close()
    torch.cuda.set_stream(torch.cuda.default_stream())


@pytest.mark.parametrize("chunk", [0, 2, 3, 8, 1000])
def test_overlapped_host_pass_equals_bind_then_pass(chunk):
    """ta_run_pass_host (chunked H2D, the scan of each chunk queued behind its copy) == ta_bind_volume +
    ta_run_pass, for any chunk height, whole volume and slab ownership."""
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.engine import memory_layout
    img = tissue_image((70, 41, 29), 90, seed=8, dome=True)
    view = np.ascontiguousarray(memory_layout(img)[0])
    for slab in (None, (4, 21, 100)):
        ref = _native.Context()
        ref.bind_host(view)
        if slab:
            ref.set_slab(*slab)
        ref.run_pass()
        want = ref.label_table() + ref.pair_table()
        ref.close()
        ctx = _native.Context()
        ctx.run_pass_host(view, chunk_planes=chunk, slab=slab)
        got = ctx.label_table() + ctx.pair_table()
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        ctx.run_pass()                       # the volume stays bound: an ordinary pass gives the same tables again
        for a, b in zip(ctx.label_table() + ctx.pair_table(), want):
            assert np.array_equal(a, b)
        ctx.close()


def test_pass_in_plane_ranges_equals_one_pass():
    """ta_run_pass_ranges (the sharded driver's overlap of halo exchange and interior scan): any tiling of the owned
    planes, in any order, with empty and single-plane ranges, gives the tables of ta_run_pass."""
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.engine import memory_layout
    img = tissue_image((60, 37, 41), 80, seed=12, dome=True)
    view = np.ascontiguousarray(memory_layout(img)[0])
    ctx = _native.Context()
    ctx.bind_host(view)
    ctx.set_slab(2, 39, 7)
    ctx.run_pass()
    want = ctx.label_table() + ctx.pair_table()
    for ranges in ([(3, 38), (2, 3), (38, 39)], [(20, 39), (2, 2), (2, 20)], [(2, 39)], [(2, 3), (3, 4), (4, 39)]):
        ctx.run_pass_ranges(ranges, [None] * len(ranges))
        for a, b in zip(ctx.label_table() + ctx.pair_table(), want):
            assert np.array_equal(a, b)
    with pytest.raises(_native.NativeError):
        ctx.run_pass_ranges([(2, 10), (11, 39)], [None, None])          # a gap
    ctx.close()


def test_kernel_variants_give_the_same_tables(monkeypatch, capfd):
    """TMA staging and cp.async staging (TA_NO_TMA=1) fill identical tables.  (TA_PHASE_TIMING only has an effect in a
    -DTA_WITH_PHASE_TIMING build; the product library ignores it.)"""
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.engine import memory_layout
    img = tissue_image((160, 48, 27), 120, seed=31, dome=True)
    view = np.ascontiguousarray(memory_layout(img)[0])

    def tables(extra_flags=0):
        ctx = _native.Context()
        ctx.bind_host(view)
        ctx.run_pass(_native.PASS_ALL | extra_flags)
        out = ctx.label_table() + ctx.pair_table()
        ctx.close()
        return out

    want = tables()
    for env in ({"TA_PHASE_TIMING": "1"}, {"TA_NO_TMA": "1"}):
        for flags in (0,):
            with monkeypatch.context() as m:
                for k, v in env.items():
                    m.setenv(k, v)
                got = tables(flags)
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (env, flags)
    capfd.readouterr()        # the phase shares go to stderr


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["uint16", "uint32"])
def test_prepass_leaves_the_tables_unchanged(monkeypatch, dtype):
    """The one-label pre-pass (ta_prepass.cuh: bricks whose neighbourhood is one label leave the scan's queue and get their
    closed-form moments from decide_kernel) against the scan on its own, on a dome with thick background around it, ragged
    bricks on every side, as one launch and as three slabs; and against the CPU oracle."""
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.engine import memory_layout
    from oracle import c_onepass
    img = tissue_image((200, 130, 77), 60, seed=5, dome=True, dtype=dtype)
    view = np.ascontiguousarray(memory_layout(img)[0])

    def tables(prepass, ranges=None):
        with monkeypatch.context() as m:
            m.setenv("TA_PREPASS", prepass)
            ctx = _native.Context()
            ctx.bind_host(view)
            if ranges is None:
                ctx.run_pass(_native.PASS_ALL, int(view.max()) if dtype == "uint32" else 0)
            else:
                ctx.run_pass_ranges(ranges, [None] * len(ranges), _native.PASS_ALL, int(view.max()) if dtype == "uint32" else 0)
            out = ctx.label_table() + ctx.pair_table()
            ctx.close()
        return out

    want = tables("0")
    got = tables("1")
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    ns = view.shape[0]
    got3 = tables("1", [(0, 24), (24, 56), (56, ns)])
    for a, b in zip(got3, want):
        assert np.array_equal(a, b)
    nrows = got[0].size
    ref = c_onepass.onepass(view, nrows=nrows)
    count, s1, s2, bbox, lo, hi, faces, wall18 = got
    for a, k in ((count, "count"), (s1, "s1"), (s2, "s2"), (lo, "lo"), (hi, "hi"), (faces, "faces"), (wall18, "wall18")):
        assert np.array_equal(a, ref[k]), k
    present = ref["count"] > 0
    assert np.array_equal(bbox[present], ref["bbox"][present])


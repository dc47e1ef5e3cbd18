"""Parity at BASELINE.json's full sizes -- C3: 1024^3 uint16, 50 000 seeds, dome; C4: 2048 x 2048 x 1024 uint32 (16 GiB),
400 000 seeds, labels beyond 65 535 -- through size-independent properties, plus a bit-exact comparison with the C oracle
on a slab of the same tissue."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tables(ctx):
    c, s1, s2, bb = ctx.label_table()
    lo, hi, f, w = ctx.pair_table()
    return dict(count=c, s1=s1, s2=s2, bbox=bb, lo=lo, hi=hi, faces=f, wall18=w)


@pytest.mark.timeout(900)
def test_c3_full_size_properties_and_slab_against_c_oracle():
    import torch
    from oracle import c_onepass
    from oracle.sia_onepass import merge_pair_tables
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.synth import CONFIGS, voronoi_device
    cfg = CONFIGS["C3"]
    X, Y, Z = cfg["shape"]
    free, _ = torch.cuda.mem_get_info()
    if free < 6 * 2 ** 30:
        pytest.skip("not enough device memory")
    vol = voronoi_device((Z, Y, X), cfg["ncell"], cfg["seed"], (1, 1, 1), True, "uint16")
    ctx = _native.Context()
    ctx.bind_device(vol.data_ptr(), 2, X, Y, Z, keepalive=vol)
    ctx.run_pass()
    whole = _tables(ctx)
    nvox = X * Y * Z

    # --- conservation laws -------------------------------------------------------------------------------------
    assert int(whole["count"].sum()) == nvox
    for k, (n, other) in enumerate(((X, Y * Z), (Y, X * Z), (Z, X * Y))):
        assert int(whole["s1"][:, k].sum()) == other * n * (n - 1) // 2                       # sum of a coordinate
    sq = lambda n: (n - 1) * n * (2 * n - 1) // 6
    assert int(whole["s2"][:, 0].sum()) == Y * Z * sq(X) and int(whole["s2"][:, 5].sum()) == X * Y * sq(Z)
    assert int(whole["s2"][:, 1].sum()) == Z * (X * (X - 1) // 2) * (Y * (Y - 1) // 2)
    present = whole["count"] > 0
    assert present.sum() > 20000 and present[1]
    assert (whole["bbox"][present][:, :3] <= whole["bbox"][present][:, 3:]).all()
    # faces per axis == number of unequal adjacent voxel pairs (independent torch reduction)
    v = vol.view(torch.int16)
    diff_f = int((v[:, :, 1:] != v[:, :, :-1]).sum())
    diff_m = int((v[:, 1:, :] != v[:, :-1, :]).sum())
    diff_s = int((v[1:] != v[:-1]).sum())
    f = whole["faces"].astype(np.int64)
    assert (int(f[:, 0:2].sum()), int(f[:, 2:4].sum()), int(f[:, 4:6].sum())) == (diff_f, diff_m, diff_s)
    assert (whole["lo"] < whole["hi"]).all() and (whole["wall18"][f.sum(axis=1) > 0] > 0).all()

    # --- three slabs on the same buffer (halo planes shared) merge to the unsplit tables, bit for bit ---------------
    parts, cuts = [], [0, 300, 701, Z]
    for lo_p, hi_p in zip(cuts[:-1], cuts[1:]):
        b0, b1 = max(lo_p - 1, 0), min(hi_p + 1, Z)
        c2 = _native.Context()
        c2.bind_device(vol.data_ptr() + b0 * X * Y * 2, 2, X, Y, b1 - b0, keepalive=vol)
        c2.set_slab(lo_p - b0, hi_p - b0, b0)
        c2.run_pass()
        parts.append(_tables(c2))
        c2.close()
    assert np.array_equal(sum(p["count"] for p in parts), whole["count"])
    assert np.array_equal(sum(p["s1"] for p in parts), whole["s1"])
    assert np.array_equal(sum(p["s2"] for p in parts), whole["s2"])
    bmin = np.minimum.reduce([p["bbox"][:, :3] for p in parts])
    bmax = np.maximum.reduce([p["bbox"][:, 3:] for p in parts])
    assert np.array_equal(bmin[present], whole["bbox"][present][:, :3])
    assert np.array_equal(bmax[present], whole["bbox"][present][:, 3:])
    merged = merge_pair_tables([dict(lo=p["lo"], hi=p["hi"], faces=p["faces"].astype(np.int64),
                                     wall18=p["wall18"].astype(np.int64)) for p in parts])
    assert np.array_equal(merged["lo"], whole["lo"]) and np.array_equal(merged["hi"], whole["hi"])
    assert np.array_equal(merged["faces"], f) and np.array_equal(merged["wall18"], whole["wall18"].astype(np.int64))

    # --- a 128-plane slab as a stand-alone volume == the C oracle, bit for bit -------------------------------------------
    z0, z1 = 448, 576
    sub = vol[z0:z1].contiguous()
    c3 = _native.Context()
    c3.bind_device(sub.data_ptr(), 2, X, Y, z1 - z0, keepalive=sub)
    c3.run_pass()
    got = _tables(c3)
    c3.close()
    ref = c_onepass.onepass(sub.cpu().numpy(), nrows=65536)
    for k in ("count", "s1", "s2", "lo", "hi", "faces", "wall18"):
        assert np.array_equal(got[k], ref[k]), k
    p2 = ref["count"] > 0
    assert np.array_equal(got["bbox"][p2], ref["bbox"][p2])
    ctx.close()


@pytest.mark.timeout(2400)
def test_c4_full_size_uint32_properties_slabs_and_c_oracle():
    """Config C4 at size: 16 GiB of uint32 labels, ~190 000 cells present (labels far beyond 65 535: the dense label table
    has 400 002 rows), ~1.5 M pairs.  Conservation laws, face totals against an independent torch reduction, three slabs
    == whole, a 64-plane slab == the C oracle bit for bit."""
    import torch
    from oracle import c_onepass
    from oracle.sia_onepass import merge_pair_tables
    from tissue_analysis_b200 import _native
    from tissue_analysis_b200.synth import CONFIGS, voronoi_device
    cfg = CONFIGS["C4"]
    X, Y, Z = cfg["shape"]
    free, _ = torch.cuda.mem_get_info()
    if free < 28 * 2 ** 30:
        pytest.skip("not enough device memory for the 16 GiB volume")
    hint = cfg["ncell"] + 1
    vol = voronoi_device((Z, Y, X), cfg["ncell"], cfg["seed"], (1, 1, 1), True, "uint32")
    ctx = _native.Context()
    ctx.bind_device(vol.data_ptr(), 4, X, Y, Z, keepalive=vol)
    ctx.run_pass(_native.PASS_ALL, hint)
    whole = _tables(ctx)
    nvox = X * Y * Z
    assert whole["count"].size == hint + 1
    assert int(whole["count"].sum()) == nvox
    for k, (n, other) in enumerate(((X, Y * Z), (Y, X * Z), (Z, X * Y))):
        assert int(whole["s1"][:, k].astype(object).sum()) == other * n * (n - 1) // 2
    sq = lambda n: (n - 1) * n * (2 * n - 1) // 6
    assert int(whole["s2"][:, 0].astype(object).sum()) == Y * Z * sq(X)
    assert int(whole["s2"][:, 5].astype(object).sum()) == X * Y * sq(Z)
    present = whole["count"] > 0
    assert present.sum() > 150000 and present[1] and present[70000:].sum() > 100000       # labels a uint16 cannot hold
    assert (whole["bbox"][present][:, :3] <= whole["bbox"][present][:, 3:]).all()
    v = vol.view(torch.int32)
    diff_f = diff_m = diff_s = 0
    for z0 in range(0, Z, 64):                                         # in chunks: the comparison masks are 4 GiB otherwise
        c = v[z0:z0 + 64]
        diff_f += int((c[:, :, 1:] != c[:, :, :-1]).sum())
        diff_m += int((c[:, 1:, :] != c[:, :-1, :]).sum())
        hi = min(z0 + 65, Z)
        diff_s += int((v[z0 + 1:hi] != v[z0:hi - 1]).sum())
    f = whole["faces"].astype(np.int64)
    assert (int(f[:, 0:2].sum()), int(f[:, 2:4].sum()), int(f[:, 4:6].sum())) == (diff_f, diff_m, diff_s)
    assert whole["lo"].size > 1000000 and (whole["lo"] < whole["hi"]).all() and int(whole["hi"].max()) > 65535

    parts, cuts = [], [0, 217, 696, Z]
    for lo_p, hi_p in zip(cuts[:-1], cuts[1:]):
        b0, b1 = max(lo_p - 1, 0), min(hi_p + 1, Z)
        c2 = _native.Context()
        c2.bind_device(vol.data_ptr() + b0 * X * Y * 4, 4, X, Y, b1 - b0, keepalive=vol)
        c2.set_slab(lo_p - b0, hi_p - b0, b0)
        c2.run_pass(_native.PASS_ALL, hint)
        parts.append(_tables(c2))
        c2.close()
    assert np.array_equal(sum(p["count"] for p in parts), whole["count"])
    assert np.array_equal(sum(p["s1"] for p in parts), whole["s1"])
    assert np.array_equal(sum(p["s2"] for p in parts), whole["s2"])
    merged = merge_pair_tables([dict(lo=p["lo"], hi=p["hi"], faces=p["faces"].astype(np.int64),
                                     wall18=p["wall18"].astype(np.int64)) for p in parts])
    assert np.array_equal(merged["lo"], whole["lo"]) and np.array_equal(merged["hi"], whole["hi"])
    assert np.array_equal(merged["faces"], f) and np.array_equal(merged["wall18"], whole["wall18"].astype(np.int64))

    z0, z1 = 480, 544
    sub = vol[z0:z1].contiguous()
    c3 = _native.Context()
    c3.bind_device(sub.data_ptr(), 4, X, Y, z1 - z0, keepalive=sub)
    c3.run_pass(_native.PASS_ALL, hint)
    got = _tables(c3)
    c3.close()
    ref = c_onepass.onepass(sub.cpu().numpy(), nrows=hint + 1)
    for k in ("count", "s1", "s2", "lo", "hi", "faces", "wall18"):
        assert np.array_equal(got[k], ref[k]), k
    p2 = ref["count"] > 0
    assert np.array_equal(got["bbox"][p2], ref["bbox"][p2])
    ctx.close()

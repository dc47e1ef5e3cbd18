"""Test-only helpers: an oracle-fed backend for the host mirror (CPU tests) and whole-API comparison."""
import numpy as np
import scipy.ndimage as nd

from oracle import sia_onepass
from tissue_analysis_b200.engine import ScanTables

TOY = np.array([[1, 2, 7, 7, 1, 1],
                [1, 6, 5, 7, 3, 3],
                [2, 2, 1, 7, 3, 3],
                [1, 1, 1, 4, 1, 1]], dtype=np.uint16).reshape(4, 6, 1)


def oracle_tables(img):
    img = np.asarray(img)
    lt = sia_onepass.label_table(img)
    pt = sia_onepass.pair_table(img)
    return ScanTables(img.shape, lt["count"], lt["s1"], lt["s2"], lt["bmin"], lt["bmax"], pt["lo"], pt["hi"],
                      pt["faces"], pt["wall18"])


class OracleBackend(object):
    """Stands in for engine.VolumeScan in `-m "not gpu"` tests: same interface, tables from the numpy oracle.
    It exists only under tests/; the product never falls back to it."""

    def __init__(self, image):
        self.img = np.asarray(image)
        self.tables = oracle_tables(self.img)

    def run(self):
        return self.tables

    def inertia(self, labels):
        t = self.tables
        evals = np.full((len(labels), 3), np.nan)
        evecs = np.full((len(labels), 3, 3), np.nan)
        pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
        for i, l in enumerate(labels):
            n = int(t.count[l]) if 0 <= l < t.nrows else 0
            if n == 0:
                continue
            cov = np.zeros((3, 3))
            for k, (a, b) in enumerate(pairs):
                num = n * int(t.s2[l, k]) - int(t.s1[l, a]) * int(t.s1[l, b])   # exact python ints
                cov[a, b] = cov[b, a] = num / (n * max(3, n))
            w, v = np.linalg.eigh(cov)
            evals[i] = w[::-1]
            evecs[i] = v[:, ::-1].T
        return evals, evecs

    def wall_voxel_coords(self, lo, hi):
        s18 = nd.generate_binary_structure(3, 2)
        out = []
        for a, b in zip(lo, hi):
            ma, mb = self.img == a, self.img == b
            hit = (nd.binary_dilation(ma, structure=s18) & mb) | (nd.binary_dilation(mb, structure=s18) & ma)
            out.append(np.array(np.where(hit)).astype(np.int64).reshape(3, -1))
        return out

    def stencil_image(self, kind):
        if kind == "hollow":
            return self.img * (nd.laplace(self.img) != 0)
        if kind == "laplace":
            return (nd.laplace(self.img) != 0).astype(self.img.dtype)
        pad = np.pad(self.img.astype(np.int64), 1, mode="constant", constant_values=-1)
        X, Y, Z = self.img.shape
        shell = np.zeros(self.img.shape, bool)
        for dx, dy, dz in sia_onepass.N18:
            shell |= pad[1 + dx:1 + dx + X, 1 + dy:1 + dy + Y, 1 + dz:1 + dz + Z] != self.img
        return shell.astype(self.img.dtype)

    def map_labels(self, lut, fill=0):
        lut = np.asarray(lut)
        idx = np.minimum(self.img.astype(np.int64), lut.size - 1)
        return np.where(self.img < lut.size, lut[idx], np.asarray(fill, lut.dtype))

    def relabel(self, mapping):
        src = self.img.copy()
        for old, new in mapping.items():
            self.img[src == old] = new          # in place: self.img aliases the caller's image
        self.tables = oracle_tables(self.img)

    def voxel_first_layer(self, background, keep_background=True):
        mask = self.img == background
        layer = nd.binary_dilation(mask) & ~mask
        out = self.img * layer
        return out + mask if keep_background else out


def _sorted_lists(d):
    return dict((int(k), sorted(int(x) for x in v)) for k, v in d.items())


def assert_eig_close(vec_p, val_p, vec_o, val_o, rtol=1e-6):
    val_p, val_o = np.asarray(val_p, float), np.asarray(val_o, float)
    scale = max(np.abs(val_o).max(), 1e-300)
    np.testing.assert_allclose(val_p, val_o, rtol=rtol, atol=1e-9 * scale + 1e-12)
    for i in range(3):
        gap = min(abs(val_o[i] - val_o[j]) for j in range(3) if j != i)
        if gap <= 1e-6 * scale:
            continue   # direction is arbitrary inside a (near-)degenerate eigenspace
        a, b = np.asarray(vec_p[i], float), np.asarray(vec_o[i], float)
        assert min(np.abs(a - b).max(), np.abs(a + b).max()) <= 1e-6 / max(gap / scale, 1e-6) * 1e-2 + 1e-7, (a, b)


def compare_api(prod, orc, check_wall_voxels=True, rtol=1e-6, eig=True, real_modes=(True, False)):
    """Every in-scope feature extractor of `prod` (product class) against `orc` (LoopOracle): integers and
    containers exactly, floats bit-exact where the reference arithmetic is reproduced, eigen-data to rtol."""
    assert sorted(prod.labels()) == sorted(orc.labels())
    assert prod.nb_labels() == orc.nb_labels()
    labels = sorted(orc.labels())
    bg = orc.background()
    for real in real_modes:
        vp, vo = prod.volume(real=real), orc.volume(real=real)
        assert set(vp) == set(vo)
        for l in labels:
            assert vp[l] == vo[l], (l, vp[l], vo[l])
    # bounding boxes (dict over labels + background)
    bp, bo = prod.boundingbox(), orc.boundingbox()
    assert bp == bo
    bpr, bor = prod.boundingbox(real=True), orc.boundingbox(real=True)
    assert bpr == bor
    for l in labels[:5]:
        assert prod.boundingbox(l) == orc.boundingbox(l)
    # centre of mass: bit-exact
    for real in real_modes:
        cp, co = prod.center_of_mass(real=real), orc.center_of_mass(real=real)
        if len(labels) == 1:
            cp, co = {labels[0]: cp}, {labels[0]: co}
        assert set(cp) == set(co)
        for l in labels:
            a, b = np.asarray(cp[l], float), np.asarray(co[l], float)
            assert np.array_equal(a, b, equal_nan=True), (l, a, b)
    # neighbours
    np_, no = prod.neighbors(), orc.neighbors()
    assert _sorted_lists(np_) == _sorted_lists(no)
    assert prod.neighbors_number() == orc.neighbors_number()
    for l in labels[:5]:
        assert sorted(map(int, prod.neighbors(l))) == sorted(map(int, orc.neighbors(l)))
    sub = labels[:7]
    assert _sorted_lists(prod.neighbors(list(sub))) == _sorted_lists(orc.neighbors(list(sub)))
    # wall areas: bit-exact floats
    for real in real_modes:
        wp, wo = prod.wall_areas(real=real), orc.wall_areas(real=real)
        assert set(wp) == set((int(a), int(b)) for a, b in wo)
        for (a, b), v in wo.items():
            assert wp[(int(a), int(b))] == v, ((a, b), wp[(int(a), int(b))], v)
    for l in labels[:4]:
        nb = sorted(map(int, orc.neighbors(l)))
        if nb:
            assert prod.cell_wall_area(l, nb) == dict(((int(a), int(b)), v) for (a, b), v in orc.cell_wall_area(l, nb).items())
            assert prod.cell_wall_area(l, nb[0], real=False) == orc.cell_wall_area(l, nb[0], real=False)
    # min contact area filter
    thr = 3.0
    fp, fo = prod.neighbors(list(sub), min_contact_area=thr, verbose=False), orc.neighbors(list(sub), min_contact_area=thr, verbose=False)
    assert _sorted_lists(fp) == _sorted_lists(fo)
    # layers
    if bg is not None:
        for kw in (dict(filter_by_area=False), dict(), dict(minimal_external_area=2, real_area=False)):
            assert sorted(prod.cell_first_layer(**kw)) == sorted(orc.cell_first_layer(**kw)), kw
        assert sorted(prod.cell_second_layer()) == sorted(map(int, orc.cell_second_layer()))
        assert np.array_equal(np.asarray(prod.voxel_first_layer()), orc.voxel_first_layer())
    # stack margins
    for d in (5, 1, 0, 3):
        assert sorted(map(int, prod.labels_at_stack_margins(d))) == sorted(map(int, orc.labels_at_stack_margins(d)))
    # inertia
    if eig:
        for real in real_modes:
            (vecp, valp), (veco, valo) = prod.inertia_axis(real=real), orc.inertia_axis(real=real)
            if len(labels) == 1:
                vecp, valp, veco, valo = {labels[0]: vecp}, {labels[0]: valp}, {labels[0]: veco}, {labels[0]: valo}
            for l in labels:
                if l == 0:
                    continue
                # `real` rescales each eigenvalue by |v * voxelsize| (sign-free), so compare after the fact
                assert_eig_close(vecp[l], valp[l], veco[l], np.real(valo[l]), rtol=rtol)
    # remaining voxel stencils
    assert np.array_equal(prod.get_all_wall_binary_image(), orc.get_all_wall_binary_image(), equal_nan=True)
    assert [list(map(int, v)) for v in prod.cells_walls_coords()] == [list(map(int, v)) for v in orc.cells_walls_coords()]
    some = labels[:6]
    if len(some) >= 2:
        assert prod.region_boundingbox(list(some)) == orc.region_boundingbox(list(some))
        for kw in (dict(), dict(region_boundingbox=True)):
            lp, lo_ = prod.cells_voxel_layer(list(some), **kw), orc.cells_voxel_layer(list(some), **kw)
            assert set(lp) == set(lo_)
            for l in some:
                assert np.array_equal(lp[l], lo_[l]), (l, kw)
        assert np.array_equal(prod.cells_voxel_layer(list(some), single_frame=True),
                              orc.cells_voxel_layer(list(some), single_frame=True))
        assert np.array_equal(prod.cells_voxel_layer(some[0]), orc.cells_voxel_layer(some[0]))
    # wall voxels
    if check_wall_voxels:
        wvp = prod.wall_voxels_per_cells_pairs(verbose=False)
        wvo = orc.wall_voxels_per_cells_pairs(verbose=False)
        assert set(wvp) == set((int(a), int(b)) for a, b in wvo)
        for (a, b), xyz in wvo.items():
            assert np.array_equal(wvp[(int(a), int(b))], xyz), (a, b)
        t = prod._tables()
        rows = t.find_pairs([k[0] for k in wvp], [k[1] for k in wvp])
        assert (rows >= 0).all()
        for k, r in zip(wvp, rows):
            assert t.wall18[r] == wvp[k].shape[1]
        if bg is not None:
            # SIA:1062-1074: the label list comes from the first voxel layer (np.unique: an ndarray, so the reference's
            # ``labels + [background]`` is an element-wise sum)
            for kw in (dict(), dict(ignore_background=True)):
                ep = prod.wall_voxels_per_cells_pairs(only_epidermis=True, verbose=False, **kw)
                eo = orc.wall_voxels_per_cells_pairs(only_epidermis=True, verbose=False, **kw)
                assert set(ep) == set((int(a), int(b)) for a, b in eo), kw
                for (a, b), xyz in eo.items():
                    assert np.array_equal(ep[(int(a), int(b))], xyz), (a, b)
    k6 = prod.neighbor_kernels()
    assert len(k6) == 6 and all(np.array_equal(x, y) for x, y in zip(k6, orc.neighbor_kernels()))

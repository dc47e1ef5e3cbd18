"""Drop-in mirror of the reference's ``SpatialImageAnalysis`` feature extractors, served from ONE CUDA pass.

Reference surface being mirrored (same names, arguments, defaults, containers, quirks):
``/root/reference/src/vplants/tissue_analysis/spatial_image_analysis.py`` ("SIA") -- factory
``SpatialImageAnalysis`` SIA:1663-1680, ``AbstractSpatialImageAnalysis`` SIA:206-1176,
``SpatialImageAnalysis3D`` SIA:1179-1448.  Where the reference loops over labels calling scipy.ndimage on
bounding-box crops, this class reads the per-label / per-pair tables that ``engine.VolumeScan`` filled in a
single streaming pass of hand-written sm_100a kernels; each method cites the reference lines it reproduces.

Nothing here computes voxel data on the CPU: without the CUDA library the first feature request raises
``_native.NativeError``.
"""
import copy
import warnings

import numpy as np

from .spatial_image import SpatialImage

NPLIST, LIST, DICT = range(3)  # SIA:204


def _as_slices(lo, hi):
    return tuple(slice(int(a), int(b) + 1) for a, b in zip(lo, hi))


def real_indices(slices, resolutions):
    """SIA:63-71."""
    return [(s.start * r, s.stop * r) for s, r in zip(slices, resolutions)]


def hollow_out_cells(image, background, remove_background=True, verbose=True, device=-1, _backend=None):
    """SIA:74-94: keep the labels where the discrete Laplacian of the label image is non-zero (cell walls), zero
    elsewhere; optionally zero the background too.  One stencil kernel (``ta_hollow_out_cells``)."""
    if verbose:
        print('Hollowing out cells... ', end='')
    if _backend is None:
        from .engine import VolumeScan
        _backend = VolumeScan(image, device=device)
    m = _backend.stencil_image("hollow")
    if remove_background:
        m = m * (m != background)
    if verbose:
        print('Done !!')
    return SpatialImage(m, voxelsize=getattr(image, "voxelsize", None))


class AbstractSpatialImageAnalysis(object):
    def __init__(self, image, ignoredlabels=[], return_type=DICT, background=None, *, device=-1, _backend=None):
        # SIA:212-270
        self.image = image if isinstance(image, SpatialImage) else SpatialImage(image)
        if isinstance(ignoredlabels, int):
            ignoredlabels = [ignoredlabels]
        self._ignoredlabels = set(ignoredlabels)
        if background is not None:
            if not isinstance(background, int):
                raise ValueError("The label you provided as background is not an integer !")
            self._ignoredlabels.update([background])
        else:
            warnings.warn("No value defining the background, some functionalities won't work !")
        try:
            self._voxelsize = image.voxelsize
        except AttributeError:
            self._voxelsize = np.ones(len(np.shape(image)))
        self._background = background
        self._labels = None
        self._bbox = None
        self._kernels = None
        self._neighbors = None
        self._cell_layer1 = None
        self._center_of_mass = {}
        try:
            import os
            self.filepath, self.filename = os.path.split(image.info["Filename"])
        except Exception:
            self.filepath, self.filename = None, None
        try:
            self.info = dict((k, v) for k, v in image.info.items() if k != "Filename")
        except Exception:
            pass
        self.return_type = return_type
        # --- scan state (no reference equivalent) ---
        self._device = device
        self._backend = _backend
        self._adj = None
        self._com_all = None
        self._max_label_of = None
        self._max_label_value = 0

    # ------------------------------------------------------------------------------------------- the scan
    def _scan(self):
        """The bound volume + its tables; the single CUDA pass runs on first use."""
        if self._backend is None:
            from .engine import VolumeScan
            self._backend = VolumeScan(self.image, device=self._device)
        if self._backend.tables is None:
            self._backend.run()
            bg = self._background
            t = self._backend.tables
            if bg is not None and not (0 <= bg < t.nrows and t.count[bg] > 0):
                print(" WARNING!!! The background you provided has not been detected in the image !")  # SIA:238-239
        return self._backend

    def _tables(self):
        return self._scan().tables

    def _drop_caches(self):
        """Every derived cache goes (the reference only resets ``_labels`` after its image mutators and then serves
        stale boxes / neighbours; here the next feature request re-runs the pass on the edited volume)."""
        self._labels = self._bbox = self._neighbors = self._cell_layer1 = self._adj = self._com_all = None
        self._center_of_mass = {}
        if hasattr(self, "_voxel_layer1"):
            self._voxel_layer1 = None

    def invalidate(self):
        """Call after editing ``self.image`` by hand: the volume is bound and scanned again on the next request."""
        self._backend = None
        self._drop_caches()

    # ------------------------------------------------------------------------------------------- image mutators
    def fuse_labels_in_image(self, labels, verbose=True):
        """SIA:1114-1134: every label of ``labels`` becomes the smallest one (one in-place device relabel)."""
        assert isinstance(labels, list) and len(labels) >= 2
        assert self.background() not in labels
        min_lab = min(labels)
        labels.remove(min_lab)
        if verbose:
            print("Fusing the following {} labels: {} to value '{}'.".format(len(labels), labels, min_lab))
        t = self._tables()
        present = [l for l in labels if 0 <= l < t.nrows and t.count[l] > 0]
        for l in labels:
            if l not in present:
                print("No boundingbox found for cell id #{}, skipping...".format(l))
        if present:
            self._scan().relabel(dict((l, min_lab) for l in present))
            self._drop_caches()
        if verbose:
            print("Done!")
        return None

    def remove_labels_from_image(self, labels, erase_value=0, verbose=True):
        """SIA:1136-1165."""
        if isinstance(labels, int):
            labels = [labels]
        labels = list(labels)
        if self.background() in labels:
            labels.remove(self.background())
        if verbose:
            print("Removing", len(labels), "cell-labels.")
        t = self._tables()
        present = [l for l in labels if 0 <= l < t.nrows and t.count[l] > 0]
        for l in labels:
            if l not in present:
                print("No boundingbox found for cell id #{}, skipping...".format(l))
        if present:
            self._scan().relabel(dict((l, erase_value) for l in present))
            self._drop_caches()
        self._ignoredlabels.update([erase_value])
        for label in labels:
            self._ignoredlabels.discard(label)
        if verbose:
            print('Done !!')

    def remove_stack_margin_labels_from_image(self, erase_value=0, voxel_distance_from_margin=5, verbose=True):
        """SIA:1168-1176."""
        if verbose:
            print("Deleting cells at the margins of the stack from 'self.image'...")
        self.remove_labels_from_image(self.labels_at_stack_margins(voxel_distance_from_margin), erase_value, verbose)

    def is3D(self):
        return False

    def background(self):
        return self._background

    def ignoredlabels(self):
        return self._ignoredlabels

    def add2ignoredlabels(self, list2add, verbose=False):
        # SIA:279-289
        if isinstance(list2add, int):
            list2add = [list2add]
        if verbose:
            print('Adding labels', list2add, 'to the list of labels to ignore...')
        self._ignoredlabels.update(list2add)
        self._labels = self.__labels()

    def consideronlylabels(self, list2consider, verbose=False):
        # SIA:291-306
        if isinstance(list2consider, int):
            list2consider = [list2consider]
        present = np.nonzero(self._tables().count)[0].tolist()
        toignore = sorted(set(present) - set(list2consider))
        if verbose:
            print('Adding labels', toignore, 'to the list of labels to ignore...')
        self._ignoredlabels.update(toignore)
        self._labels = self.__labels()

    def convert_return(self, values, labels=None, overide_return_type=None):
        # SIA:309-334
        rt = self.return_type if overide_return_type is None else overide_return_type
        if labels is not None and isinstance(labels, int):
            return values
        if rt == NPLIST:
            return values
        if rt == LIST:
            return values if isinstance(values, list) else values.tolist()
        return dict(zip(labels, values))

    # ------------------------------------------------------------------------------------------- labels
    def labels(self):
        # SIA:337-356
        if self._labels is None:
            self._labels = self.__labels()
        return self._labels

    def __labels(self):
        # SIA:358-364: set(np.unique(image)) - ignored
        present = np.nonzero(self._tables().count)[0].tolist()
        ign = self._ignoredlabels
        return [l for l in present if l not in ign]

    def nb_labels(self):
        return len(self.labels())

    def label_request(self, labels):
        # SIA:387-414
        if isinstance(labels, int):
            if labels not in self.labels():
                print("The following id was not found within the image labels: {}".format(labels))
            return [labels]
        if isinstance(labels, list):
            return list(set(labels) & set(self.labels()))
        if labels is None:
            return self.labels()
        if isinstance(labels, str):
            key = labels.lower()
            if key == 'all':
                return self.labels()
            if key == 'l1':
                return self.cell_first_layer()
            if key == 'l2':
                return self.cell_second_layer()
            return labels
        raise ValueError("This is not usable as `labels`: {}".format(labels))

    def _rows(self, labels):
        """(index array clipped into the table, validity mask) for arbitrary label values."""
        t = self._tables()
        lab = np.asarray(labels, dtype=np.int64).reshape(-1)
        ok = (lab >= 0) & (lab < t.nrows)
        idx = np.where(ok, lab, 0)
        ok &= t.count[idx] > 0
        return idx, ok

    # ------------------------------------------------------------------------------------------- center_of_mass
    def _com_voxel(self, labels):
        """SIA:464-467 from exact sums: fl(fl((S_a - n*start_a) / n) + start_a); nan for absent labels and for
        label 0 (nd.center_of_mass weights the crop by the label values themselves)."""
        t = self._tables()
        idx, ok = self._rows(labels)
        n = t.count[idx]
        start = t.bmin[idx]
        with np.errstate(divide="ignore", invalid="ignore"):
            c = (t.s1[idx] - n[:, None] * start).astype(np.float64) / n[:, None].astype(np.float64) + start
        c[~ok | (idx == 0)] = np.nan
        return c

    def center_of_mass(self, labels=None, real=True, verbose=False):
        # SIA:417-480
        labels = self.label_request(labels)
        missing = [l for l in labels if l not in self._center_of_mass]
        if missing:
            c = self._com_voxel(missing)
            for l, row in zip(missing, c):
                self._center_of_mass[l] = [row[0], row[1], row[2]]
        center = dict((l, self._center_of_mass[l]) for l in labels)
        if real:
            center = dict((l, np.multiply(center[l], self._voxelsize)) for l in labels)
        if len(labels) == 1:
            return center[labels[0]]
        return center

    # ------------------------------------------------------------------------------------------- boundingbox
    def _max_label(self):
        t = self._tables()
        if self._max_label_of is not t:                   # one scan of the count column per set of tables
            present = np.nonzero(t.count)[0]
            self._max_label_value = int(present[-1]) if present.size else 0
            self._max_label_of = t
        return self._max_label_value

    def _bbox_entry(self, i):
        """``nd.find_objects(image)[i-1]`` (SIA:517, 526, 533) including its list-indexing behaviour."""
        t = self._tables()
        n = self._max_label()
        k = i - 1
        if k < 0:
            k += n
        if k < 0 or k >= n:
            raise IndexError("list index out of range")
        lab = k + 1
        if t.count[lab] == 0:
            return None
        return _as_slices(t.bmin[lab], t.bmax[lab])

    def _bbox_entries(self, labels):
        """``[_bbox_entry(i) for i in labels]`` without a Python call per label: the table columns go to lists once and the
        slices are made in one comprehension (50 000 labels: tens of milliseconds instead of a third of a second)."""
        t = self._tables()
        n = self._max_label()
        lab = np.asarray(labels, dtype=np.int64).reshape(-1)
        k = lab - 1
        k = np.where(k < 0, k + n, k)
        if lab.size and (k.min() < 0 or k.max() >= n):
            raise IndexError("list index out of range")
        rows = k + 1
        present = (t.count[rows] > 0).tolist()
        lo = t.bmin[rows].tolist()
        hi = (t.bmax[rows] + 1).tolist()
        return [(slice(a[0], b[0]), slice(a[1], b[1]), slice(a[2], b[2])) if p else None
                for a, b, p in zip(lo, hi, present)]

    def boundingbox(self, labels=None, real=False):
        # SIA:483-535
        t = self._tables()
        if isinstance(labels, (int, np.integer)) and labels == 0:       # SIA:513 ``if labels == 0`` (numpy scalars too)
            if t.count[0] == 0:
                raise IndexError("list index out of range")
            return _as_slices(t.bmin[0], t.bmax[0])
        if labels is None:
            labels = copy.copy(self.labels())
            if self.background() is not None:
                labels.append(self.background())
        if isinstance(labels, list):
            bboxes = self._bbox_entries(labels)
            if real:
                return self.convert_return([real_indices(b, self._voxelsize) for b in bboxes], labels)
            return self.convert_return(bboxes, labels)
        try:
            if real:
                return real_indices(self._bbox_entry(labels), self._voxelsize)
            return self._bbox_entry(labels)
        except Exception:
            return None

    # ------------------------------------------------------------------------------------------- neighbors
    def _adjacency(self):
        """CSR of the 6-connected label adjacency: pairs with at least one shared voxel face (SIA:45-60)."""
        if self._adj is None:
            t = self._tables()
            touching = t.faces.sum(axis=1) > 0
            lo, hi = t.pair_lo[touching], t.pair_hi[touching]
            src = np.concatenate([lo, hi])
            dst = np.concatenate([hi, lo])
            order = np.lexsort((dst, src))
            src, dst = src[order], dst[order]
            indptr = np.searchsorted(src, np.arange(t.nrows + 1))
            self._adj = (indptr, dst)
        return self._adj

    def _ring_labels_of(self, label):
        indptr, dst = self._adjacency()
        if not (0 <= label < len(indptr) - 1):
            return []
        return dst[indptr[label]:indptr[label + 1]].tolist()

    def neighbors(self, labels=None, min_contact_area=None, real_area=True, verbose=True):
        # SIA:538-587
        if (min_contact_area is not None) and verbose:
            if real_area:
                print(u"Neighbors will be filtered according to a min contact area of %.2f μm²" % min_contact_area)
            else:
                print("Neighbors will be filtered according to a min contact area of %d voxels" % min_contact_area)
        if labels is None:
            return self._all_neighbors(min_contact_area, real_area)
        elif not isinstance(labels, list):
            return self._neighbors_with_mask(labels, min_contact_area, real_area)
        else:
            return self._neighbors_from_list_with_mask(labels, min_contact_area, real_area)

    def _neighbors_with_mask(self, label, min_contact_area=None, real_area=True):
        # SIA:589-607
        if self._neighbors is not None and label in self._neighbors:
            result = self._neighbors[label]
        else:
            result = self._ring_labels_of(label)
        if min_contact_area is None:
            return result
        return self._neighbors_filtering_by_contact_area(label, result, min_contact_area, real_area)

    def _neighbors_from_list_with_mask(self, labels, min_contact_area=None, real_area=True):
        # SIA:609-630
        if self._neighbors is not None and all(i in self._neighbors for i in labels):
            edges = dict((i, self._neighbors[i]) for i in labels)
        else:
            edges = dict((i, self._ring_labels_of(i)) for i in labels)
        if min_contact_area is None:
            return edges
        return self._filter_with_area(edges, min_contact_area, real_area)

    def _all_neighbors(self, min_contact_area=None, real_area=True):
        # SIA:632-660: keys are labels() plus the background
        if self._neighbors is None:
            keys = self.boundingbox()
            if self.return_type in (NPLIST, LIST):
                keys = range(1, len(keys) + 1)      # SIA:642-645
            indptr, dst = self._adjacency()
            ip, dl, top = indptr.tolist(), dst.tolist(), len(indptr) - 1      # two conversions, then list slices
            self._neighbors = dict((l, dl[ip[l]:ip[l + 1]] if 0 <= l < top else []) for l in keys)
        if min_contact_area is None:
            return self._neighbors
        return self._filter_with_area(self._neighbors, min_contact_area, real_area)

    def _filter_with_area(self, neighborhood_dictionary, min_contact_area, real_area):
        # SIA:662-675
        return dict((label, self._neighbors_filtering_by_contact_area(label, nei, min_contact_area, real_area))
                    for label, nei in neighborhood_dictionary.items())

    def _neighbors_filtering_by_contact_area(self, label, neighbors, min_contact_area, real_area):
        # SIA:677-693
        areas = self.cell_wall_area(label, neighbors, real_area)
        nei = copy.copy(neighbors)
        for i, j in areas.keys():
            if areas[(i, j)] < min_contact_area:
                nei.remove(i if j == label else j)
        return nei

    def neighbor_kernels(self):
        """SIA:695-732: the six one-sided structuring elements the reference dilates with (a = 0: +x, 1: -x, 2: +y,
        3: -y, 4: +z, 5: -z).  Nothing here dilates with them -- the scan counts the faces per direction -- the accessor is
        part of the public surface."""
        if self._kernels is None:
            kernels = []
            for axis in range(3):
                for drop in (0, 2):
                    k = np.zeros((3, 3, 3), np.bool_)
                    idx = [1, 1, 1]
                    idx[axis] = slice(None)
                    k[tuple(idx)] = True
                    idx[axis] = drop
                    k[tuple(idx)] = False
                    kernels.append(k)
            self._kernels = tuple(kernels)
        return self._kernels

    def neighbors_number(self, labels=None, min_contact_area=None, real_area=True, verbose=True):
        # SIA:734-742
        nei = self.neighbors(labels, min_contact_area, real_area, verbose)
        if isinstance(nei, dict):
            return dict((k, len(v)) for k, v in nei.items())
        return len(nei)

    def get_all_wall_binary_image(self):
        """SIA:744-749: ``lp / lp`` of the Laplacian: 1.0 on walls, nan elsewhere (0 / 0)."""
        walls = np.asarray(self._scan().stencil_image("laplace")) != 0
        return np.where(walls, 1.0, np.nan)

    def cells_walls_coords(self):
        """SIA:883-905.  The reference passes the bound method ``self.background`` (not its value) to
        hollow_out_cells, so the background is never removed; the same happens here."""
        m = np.asarray(self._scan().stencil_image("hollow"))
        x, y, z = np.where(m != 0)
        return list(x), list(y), list(z)

    def get_voxel_face_surface(self):
        # SIA:751-756
        a = self._voxelsize
        if len(a) == 3:
            return np.array([a[1] * a[2], a[2] * a[0], a[0] * a[1]])
        return np.array([a[0], a[1]])

    # ------------------------------------------------------------------------------------------- wall areas
    def _directional_counts(self, label_id, neighbors):
        """int64[n,6]: for neighbour n_i, faces in the reference's kernel order a = 0..5 (+x,-x,+y,-y,+z,-z as
        seen from ``label_id``, SIA:695-716).  The table stores slots from the smaller label's side."""
        t = self._tables()
        nb = np.asarray(neighbors, dtype=np.int64).reshape(-1)
        rows = t.find_pairs(np.full(nb.shape, label_id, np.int64), nb)
        c = np.where(rows[:, None] >= 0, t.faces[np.maximum(rows, 0)], 0)
        flip = nb < label_id                      # label_id is the larger label: swap +/- slots
        c[flip] = c[flip][:, [1, 0, 3, 2, 5, 4]]
        return c

    def cell_wall_area(self, label_id, neighbors, real=True):
        # SIA:908-959
        resolution = self.get_voxel_face_surface()
        unique_neighbor = not isinstance(neighbors, list)
        if unique_neighbor:
            neighbors = [neighbors]
        counts = self._directional_counts(label_id, neighbors)
        wall = {}
        keys = [(min(label_id, n), max(label_id, n)) for n in neighbors]
        if len(set(keys)) == len(keys):
            total = np.zeros(len(keys))
            for a in range(6):                     # same left fold as SIA:947-956
                total = total + (counts[:, a] * resolution[a // 2] if real else counts[:, a])
            wall = dict(zip(keys, total.tolist()))
        else:
            for a in range(6):
                for k, key in enumerate(keys):
                    nb_pix = int(counts[k, a])
                    area = float(nb_pix * resolution[a // 2]) if real else nb_pix
                    wall[key] = wall.get(key, 0.0) + area
        if unique_neighbor:
            return next(iter(wall.values()))
        return wall

    def wall_areas(self, neighbors=None, real=True):
        # SIA:962-993
        if neighbors is None:
            neighbors = self.neighbors()
        fast = self._wall_areas_vectorised(neighbors, real)
        if fast is not None:
            return fast
        areas = {}
        for label_id, lneighbors in neighbors.items():
            neigh = [n for n in lneighbors if n > label_id]
            if len(neigh) > 0:
                lareas = self.cell_wall_area(label_id, neigh, real=real)
                for key in lareas:
                    areas[key] = areas.get(key, 0.0) + lareas[key]
        return areas

    def _wall_areas_vectorised(self, neighbors, real):
        """``wall_areas`` for the usual input (int labels, no label twice in a list) in one table lookup: the same
        keys in the same order and, element by element, the same left fold over the six directions as the per-label
        loop (SIA:947-956, 986-992), so the floats are bit-identical.  None: take the loop."""
        src, dst = [], []
        for label_id, lneighbors in neighbors.items():
            if not isinstance(lneighbors, list):
                return None
            neigh = [n for n in lneighbors if n > label_id]
            if len(set(neigh)) != len(neigh):
                return None
            src.extend([label_id] * len(neigh))
            dst.extend(neigh)
        if not src:
            return {}
        try:
            a = np.asarray(src, dtype=np.int64)
            b = np.asarray(dst, dtype=np.int64)
        except (TypeError, ValueError, OverflowError):
            return None
        if a.shape != b.shape or a.ndim != 1:
            return None
        t = self._tables()
        rows = t.find_pairs(a, b)
        counts = np.where(rows[:, None] >= 0, t.faces[np.maximum(rows, 0)], 0)     # label_id < n: slots as stored
        resolution = self.get_voxel_face_surface()
        total = np.zeros(len(src))
        for k in range(6):
            total = total + (counts[:, k] * resolution[k // 2] if real else counts[:, k])
        return dict(zip(zip(src, dst), total.tolist()))

    # ------------------------------------------------------------------------------------------- layers
    def cell_first_layer(self, filter_by_area=True, minimal_external_area=10, real_area=True):
        # SIA:996-1010
        if self._cell_layer1 is None:
            self._cell_layer1 = list(map(int, self.neighbors(self.background())))
        cell_layer1 = self._cell_layer1
        if filter_by_area:
            bg = self.background()
            labels_area = self.cell_wall_area(bg, self._cell_layer1, real_area)
            cell_layer1 = [l for l in self._cell_layer1
                           if ((bg, l) in labels_area) and (labels_area[(bg, l)] > minimal_external_area)]
        return list(set(cell_layer1) - self._ignoredlabels)

    def cell_second_layer(self, filter_by_area=True, minimal_L1_area=10, real_area=True):
        # SIA:1012-1022
        L1_neighbors = self.neighbors(self.cell_first_layer(), minimal_L1_area, real_area, True)
        l2 = set([])
        for nei in L1_neighbors.values():
            l2.update(nei)
        self._cell_layer2 = list(l2 - set(self._cell_layer1) - self._ignoredlabels)
        return self._cell_layer2

    # ------------------------------------------------------------------------------------------- wall voxels
    def wall_voxels_between_two_cells(self, label_1, label_2, bbox=None, verbose=False):
        """SIA:759-804: int[3, N] coordinates of the voxels of either label that have an 18-neighbour of the
        other label, in np.where order.  ``bbox`` only restricted the search region in the reference."""
        return self._scan().wall_voxel_coords([min(label_1, label_2)], [max(label_1, label_2)])[0]

    def wall_voxels_per_cell(self, label_1, bbox=None, neighbors=None, neighbors2ignore=[], verbose=False):
        # SIA:807-880
        if neighbors is None:
            neighbors = self.neighbors(label_1)
        if isinstance(neighbors, int):
            neighbors = [neighbors]
        if isinstance(neighbors, dict) and len(neighbors) != 1:
            neighbors = copy.copy(neighbors[label_1])
        neighbors = [n for n in neighbors if n not in neighbors2ignore]
        lo = [min(label_1, n) for n in neighbors]
        hi = [max(label_1, n) for n in neighbors]
        blocks = self._scan().wall_voxel_coords(lo, hi)
        coord, not_found = {}, []
        for n, l, h, xyz in zip(neighbors, lo, hi, blocks):
            if xyz.shape[1] > 0:
                coord[(l, h)] = xyz
            else:
                not_found.append(n)
        if not_found:
            print("Some walls have not been found comparing to the `neighbors` list of {}: {}".format(label_1, not_found))
        return coord

    def wall_voxels_per_cells_pairs(self, labels=None, neighborhood=None, only_epidermis=False,
                                    ignore_background=False, min_contact_area=None, real_area=True, verbose=True):
        # SIA:1049-1111 -- the per-label loop only decides WHICH pairs are extracted; the voxels of all of them come
        # from one device pass.
        # only_epidermis (SIA:1062-1065, 1073-1074): the first-voxel-layer image only ever supplies the label list,
        # np.unique of it (an ndarray: 0, the background mark and the labels of the layer); the voxels still come from
        # the image itself.
        compute_neighborhood = neighborhood is None
        if isinstance(labels, list) and isinstance(neighborhood, dict):
            labels = [label for label in labels if label in neighborhood]
        if labels is None and not only_epidermis:
            labels = self.labels()
        elif labels is None and only_epidermis:
            labels = np.unique(np.asarray(self.voxel_first_layer(True)))
        elif isinstance(labels, list):
            labels.sort()
            if not isinstance(neighborhood, dict):
                compute_neighborhood = True
        elif isinstance(labels, int):
            labels = [labels]
        else:
            raise ValueError("Couldn't find any labels.")
        if isinstance(labels, np.ndarray):
            # ``labels + [background]`` (SIA:1101) on an ndarray is numpy's element-wise sum, not a concatenation
            allowed = set(labels.tolist()) if ignore_background else set((labels + [self.background()]).tolist())
            labels = [int(l) for l in labels]
        else:
            allowed = set(labels) if ignore_background else set(labels) | set([self.background()])
        wanted, seen = [], set()
        for label in labels:
            if compute_neighborhood:
                neighbors = self.neighbors(label, min_contact_area, real_area, verbose=False)
            elif isinstance(neighborhood, dict):
                neighbors = neighborhood[label]
            else:
                neighbors = neighborhood
            for n in neighbors:
                key = (min(label, n), max(label, n))
                if n in allowed and key not in seen:
                    seen.add(key)
                    wanted.append(key)
        blocks = self._scan().wall_voxel_coords([k[0] for k in wanted], [k[1] for k in wanted])
        return dict((k, xyz) for k, xyz in zip(wanted, blocks) if xyz.shape[1] > 0)


class SpatialImageAnalysis3D(AbstractSpatialImageAnalysis):
    """SIA:1179-1448."""

    def __init__(self, image, ignoredlabels=[], return_type=DICT, background=None, **kw):
        AbstractSpatialImageAnalysis.__init__(self, image, ignoredlabels, return_type, background, **kw)
        self._voxel_layer1 = None
        self.principal_curvatures = {}
        self.principal_curvatures_normal = {}
        self.principal_curvatures_directions = {}
        self.principal_curvatures_origin = {}
        self.curvatures_tensor = {}
        self.external_wall_geometric_median = {}
        self.epidermis_wall_median_voxel = {}

    def is3D(self):
        return True

    def volume(self, labels=None, real=True):
        # SIA:1197-1243 (wide label index instead of np.int16, which overflows above 32767)
        labels = self.label_request(labels)
        t = self._tables()
        idx, ok = self._rows(labels)
        volume = np.where(ok, t.count[idx], 0).astype(np.float64)
        if real is True:
            volume = np.multiply(volume, (self._voxelsize[0] * self._voxelsize[1] * self._voxelsize[2]))
        return self.convert_return(volume, labels)

    def inertia_axis(self, labels=None, real=True, verbose=False):
        # SIA:1246-1292: covariance (1/max(3,N)) sum (p-c)(p-c)^T from exact integer moments, batched Jacobi on
        # the device; eigenvalues descending, eigenvectors by rows.
        labels = self.label_request(labels)
        evals, evecs = self._scan().inertia(labels)
        if real:
            for i in range(3):
                evals[:, i] *= np.linalg.norm(np.multiply(evecs[:, i, :], self._voxelsize), axis=1)
        by_rows = [[v[0], v[1], v[2]] for v in evecs]
        vals = [w for w in evals]
        if len(labels) == 1:
            return by_rows[0], vals[0]
        return self.convert_return(by_rows, labels), self.convert_return(vals, labels)

    def reduced_inertia_axis(self, labels=None, real=True, verbose=False):
        # SIA:1295-1341 performs the same computation as inertia_axis
        return self.inertia_axis(labels, real, verbose)

    def labels_at_stack_margins(self, voxel_distance_from_margin=5):
        # SIA:1344-1358 == bounding box test per label: start < d or stop > dim - d
        t = self._tables()
        d = voxel_distance_from_margin
        present = t.count > 0
        if d == 0:
            hit = present.copy()              # image[-0:] is the whole image
        else:
            dims = np.asarray(t.shape, np.int64)
            hit = present & ((t.bmin < d).any(axis=1) | ((t.bmax + 1) > (dims - d)).any(axis=1))
        return list(set(np.nonzero(hit)[0].tolist()) - set([self._background]))

    def region_boundingbox(self, labels):
        """SIA:1361-1396."""
        if isinstance(labels, list) and len(labels) == 1:
            return self.boundingbox(labels[0])
        if isinstance(labels, int):
            return self.boundingbox(labels)
        dict_slices = self.boundingbox(labels)
        not_found = [c for c in labels if c not in dict_slices]
        if len(not_found) != 0:
            warnings.warn('You have asked for unknown cells labels: ' + " ".join([str(k) for k in not_found]))
        x_start, y_start, z_start, x_stop, y_stop, z_stop = np.inf, np.inf, np.inf, 0, 0, 0
        for c in labels:
            x, y, z = dict_slices[c]
            x_start, y_start, z_start = min(x.start, x_start), min(y.start, y_start), min(z.start, z_start)
            x_stop, y_stop, z_stop = max(x.stop, x_stop), max(y.stop, y_stop), max(z.stop, z_stop)
        return (slice(x_start, x_stop), slice(y_start, y_stop), slice(z_start, z_stop))

    def cells_voxel_layer(self, labels, region_boundingbox=False, single_frame=False):
        """SIA:1399-1448: first voxel layer of each cell = mask minus its 18-connected erosion inside the crop.  A
        cell voxel is eroded away iff one of its 18 neighbours is outside the crop or not the cell, which for the
        bounding box of the cell (or any box containing it) is: outside the image or another label -> one stencil
        kernel for all cells (``ta_cell_shell18``), cropped per cell on the host."""
        if isinstance(labels, int):
            labels = [labels]
        if single_frame:
            region_boundingbox = True
        bbox = None
        if not isinstance(region_boundingbox, bool):
            if sum([isinstance(s, slice) for s in region_boundingbox]) == 3:
                bbox = region_boundingbox
                region_boundingbox = True
            else:
                print("TypeError: Wong type for 'region_boundingbox', should either be bool or la tuple of slices")
                return None
        elif region_boundingbox:
            bbox = self.region_boundingbox(labels)
        else:
            bboxes = self.boundingbox(labels, real=False)
        shell = np.asarray(self._scan().stencil_image("shell18")) != 0
        img = np.asarray(self.image)
        vox_layer = np.zeros_like(img[bbox], dtype=int) if single_frame else {}
        for clabel in labels:
            box = bbox if region_boundingbox else bboxes[clabel]
            layer = np.array((img[box] == clabel) & shell[box], dtype=int)
            if single_frame:
                vox_layer += layer
            else:
                vox_layer[clabel] = layer
        if len(labels) == 1:
            return vox_layer[clabel]
        return vox_layer

    def voxel_first_layer(self, keep_background=True):
        # SIA:1024-1046
        if self._voxel_layer1 is None:
            print("Extracting the first layer of voxels...")
            self._voxel_layer1 = self._scan().voxel_first_layer(self.background(), keep_background)
        return self._voxel_layer1


def SpatialImageAnalysis(image, *args, **kwd):
    """SIA:1663-1680: a file name is read first (SIA:1668-1671), then dispatch on dimensionality.  The reference
    routes 2D input (and shape[2] == 1) to a
    ``SpatialImageAnalysis2D`` class that is not defined anywhere in it; that path raises here as well."""
    if isinstance(image, str):
        from .serial import imread          # SIA:1668-1671 (openalea's imread; INRIMAGE-4 stacks here)
        image = imread(image)
    assert len(image.shape) in [2, 3]
    if len(image.shape) == 2 or image.shape[2] == 1:
        raise NotImplementedError("SpatialImageAnalysis2D is referenced but never defined by the reference (SIA:1677)")
    return SpatialImageAnalysis3D(image, *args, **kwd)

"""CPU ORACLE (test infrastructure, never shipped, never on the product path).

Python-3 restatement of the reference's per-label numpy/scipy.ndimage loops for the
``SpatialImageAnalysis3D`` feature extractors.  Reference file (py2, not importable
here): /root/reference/src/vplants/tissue_analysis/spatial_image_analysis.py ("SIA").
Every method cites the SIA lines it follows; the loop structure (one bounding-box crop
per label, the same scipy.ndimage calls, the same float operation order) is kept so the
results are the reference's results under numpy 2.3 / scipy 1.18.

Parity pinning: the only values the reference itself pins are its docstring examples
on a 4x6 toy image (SIA:343-353, 437-450, 498-511, 561-574, 924-927, 978-982,
1219-1226); tests/test_oracle_docstring_kat.py checks this file against every one of
them.  Beyond that toy image parity is unpinned at the scipy/numpy boundary (the
reference has no tests, test/__init__.py:1-11).

Documented deviations (the reference line cannot run under numpy>=2 / py3):
  * SIA:1231 ``np.int16(labels)`` overflows for labels > 32767 -> a wide index is used.
  * SIA:1033 boolean ``dil - mask`` -> ``dil & ~mask`` (same truth table, dil >= mask).
  * SIA:861  ``x != []`` on an ndarray -> ``len(x) > 0`` (what old numpy evaluated to).
  * SIA:825  ``bbox(label_1)`` on a dict (TypeError) -> ``bbox[label_1]``.
  * SIA:1104-1106 mutates the cached neighbour list while iterating -> iterate a copy.
"""
import copy
import warnings

import numpy as np
import scipy.ndimage as nd

NPLIST, LIST, DICT = range(3)  # SIA:204


# ---------------------------------------------------------------- helpers SIA:35-71
def grow_box(box, amount=1):
    """SIA:35-42 ``dilation`` / ``dilation_by``: widen a tuple of slices."""
    return tuple(slice(max(0, s.start - amount), s.stop + amount) for s in box)


def ring_values(sub, label):
    """SIA:45-52 ``wall``: values on the 6-connected outer ring of ``label`` in ``sub``."""
    inside = sub == label
    ring = nd.binary_dilation(inside) & ~inside
    return sub[ring]


def touching_labels(sub, label):
    """SIA:55-60 ``contact_surface``."""
    return set(np.unique(ring_values(sub, label)))


def box_to_real(box, vs):
    """SIA:63-71 ``real_indices``."""
    return [(s.start * r, s.stop * r) for s, r in zip(box, vs)]


def covariance_of(points):
    """SIA:137-150: 1/max(shape) * P.P^T (divides by max(3, N))."""
    p = np.asarray(points)
    if p.shape[0] > 3:
        p = p.T
    return 1.0 / max(p.shape) * np.dot(p, p.T)


def sorted_eig(cov):
    """SIA:152-167: general eig, descending order, eigenvectors by rows."""
    w, v = np.linalg.eig(cov)
    order = w.argsort()[::-1]
    return w[order], np.array(v[:, order]).T


def one_sided_kernels():
    """SIA:695-716: the six 3x3x3 one-sided structuring elements, in reference order."""
    out = []
    for axis in range(3):
        for drop in (0, 2):
            k = np.zeros((3, 3, 3), bool)
            idx = [1, 1, 1]
            idx[axis] = slice(None)
            k[tuple(idx)] = True
            idx[axis] = drop
            k[tuple(idx)] = False
            out.append(k)
    return tuple(out)


def hollow_out_cells(image, background, remove_background=True):
    """SIA:74-94."""
    image = np.asarray(image)
    m = image * (nd.laplace(image) != 0)
    if remove_background:
        m = m * (m != background)
    return m


class LoopOracle(object):
    """Restates AbstractSpatialImageAnalysis + SpatialImageAnalysis3D (SIA:206-1448)."""

    def __init__(self, image, ignoredlabels=[], return_type=DICT, background=None, voxelsize=None):
        # SIA:212-270
        vs = voxelsize if voxelsize is not None else getattr(image, "voxelsize", None)
        self.image = np.asarray(image)
        if isinstance(ignoredlabels, int):
            ignoredlabels = [ignoredlabels]
        self._ignoredlabels = set(ignoredlabels)
        if background is not None:
            if not isinstance(background, int):
                raise ValueError("The label you provided as background is not an integer !")
            if background not in self.image:
                print(" WARNING!!! The background you provided has not been detected in the image !")
            self._ignoredlabels.update([background])
        else:
            warnings.warn("No value defining the background, some functionalities won't work !")
        self._voxelsize = np.ones(self.image.ndim) if vs is None else tuple(vs)
        self._background = background
        self._labels = None
        self._bbox = None
        self._kernels = None
        self._neighbors = None
        self._cell_layer1 = None
        self._center_of_mass = {}
        self._voxel_layer1 = None
        self.return_type = return_type

    def is3D(self):
        return True

    def background(self):
        return self._background

    def ignoredlabels(self):
        return self._ignoredlabels

    def add2ignoredlabels(self, list2add, verbose=False):
        # SIA:279-289
        if isinstance(list2add, int):
            list2add = [list2add]
        self._ignoredlabels.update(list2add)
        self._labels = self._compute_labels()

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

    # ------------------------------------------------------------ labels SIA:337-414
    def _compute_labels(self):
        return list(map(int, set(np.unique(self.image)) - self._ignoredlabels))

    def labels(self):
        if self._labels is None:
            self._labels = self._compute_labels()
        return self._labels

    def nb_labels(self):
        return len(self.labels())

    def label_request(self, labels):
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
            if key == "all":
                return self.labels()
            if key == "l1":
                return self.cell_first_layer()
            if key == "l2":
                return self.cell_second_layer()
            return labels
        raise ValueError("This is not usable as `labels`: {}".format(labels))

    # ---------------------------------------------------- center_of_mass SIA:417-480
    def center_of_mass(self, labels=None, real=True, verbose=False):
        labels = self.label_request(labels)
        center = {}
        for l in labels:
            if l in self._center_of_mass:
                center[l] = self._center_of_mass[l]
                continue
            try:
                box = self.boundingbox(l, real=False)
                sub = self.image[box]
                com = np.array(nd.center_of_mass(sub, sub, index=l))
                com = [com[i] + s.start for i, s in enumerate(box)]
            except Exception:
                sub = self.image
                com = np.array(nd.center_of_mass(sub, sub, index=l))
            self._center_of_mass[l] = com
            center[l] = com
        if real:
            center = dict((l, np.multiply(center[l], self._voxelsize)) for l in labels)
        if len(labels) == 1:
            return center[labels[0]]
        return center

    # ------------------------------------------------------- boundingbox SIA:483-535
    def boundingbox(self, labels=None, real=False):
        if isinstance(labels, (int, np.integer)) and labels == 0:    # SIA:513 ``if labels == 0`` (numpy scalars too)
            return nd.find_objects((self.image == 0).astype(np.uint8))[0]    # scipy >= 1.12 rejects a boolean input
        if self._bbox is None:
            self._bbox = nd.find_objects(self.image)
        if labels is None:
            labels = copy.copy(self.labels())
            if self.background() is not None:
                labels.append(self.background())
        if isinstance(labels, list):
            boxes = [self._bbox[i - 1] for i in labels]
            if real:
                return self.convert_return([box_to_real(b, self._voxelsize) for b in boxes], labels)
            return self.convert_return(boxes, labels)
        try:
            if real:
                return box_to_real(self._bbox[labels - 1], self._voxelsize)
            return self._bbox[labels - 1]
        except Exception:
            return None

    # --------------------------------------------------------- neighbors SIA:538-693
    def _ring_labels_of(self, label):
        try:
            sub = self.image[grow_box(self.boundingbox(label))]
        except Exception:
            sub = self.image
        return list(touching_labels(sub, label))

    def neighbors(self, labels=None, min_contact_area=None, real_area=True, verbose=True):
        if labels is None:
            return self._all_neighbors(min_contact_area, real_area)
        if not isinstance(labels, list):
            return self._neighbors_of_one(labels, min_contact_area, real_area)
        return self._neighbors_of_list(labels, min_contact_area, real_area)

    def _neighbors_of_one(self, label, min_contact_area=None, real_area=True):
        # SIA:589-607
        if self._neighbors is not None and label in self._neighbors:
            found = self._neighbors[label]
            if min_contact_area is None:
                return found
            return self._drop_small_contacts(label, found, min_contact_area, real_area)
        found = self._ring_labels_of(label)
        if min_contact_area is not None:
            found = self._drop_small_contacts(label, found, min_contact_area, real_area)
        return found

    def _neighbors_of_list(self, labels, min_contact_area=None, real_area=True):
        # SIA:609-630
        if self._neighbors is not None and all(i in self._neighbors for i in labels):
            found = dict((i, self._neighbors[i]) for i in labels)
            if min_contact_area is None:
                return found
            return self._filter_with_area(found, min_contact_area, real_area)
        edges = {}
        for label in labels:
            nei = self._ring_labels_of(label)
            if min_contact_area is not None:
                nei = self._drop_small_contacts(label, nei, min_contact_area, real_area)
            edges[label] = nei
        return edges

    def _all_neighbors(self, min_contact_area=None, real_area=True):
        # SIA:632-660
        if self._neighbors is not None:
            if min_contact_area is None:
                return self._neighbors
            return self._filter_with_area(self._neighbors, min_contact_area, real_area)
        boxes = self.boundingbox()
        if self.return_type in (NPLIST, LIST):
            boxes = dict((i + 1, b) for i, b in enumerate(boxes))
        edges = {}
        for label_id, box in boxes.items():
            try:
                sub = self.image[grow_box(box)]
            except Exception:
                sub = self.image
            edges[label_id] = list(touching_labels(sub, label_id))
        self._neighbors = edges
        if min_contact_area is None:
            return edges
        return self._filter_with_area(edges, min_contact_area, real_area)

    def _filter_with_area(self, neighborhood, min_contact_area, real_area):
        # SIA:662-675
        return dict((label, self._drop_small_contacts(label, nei, min_contact_area, real_area))
                    for label, nei in neighborhood.items())

    def _drop_small_contacts(self, label, neighbors, min_contact_area, real_area):
        # SIA:677-693
        areas = self.cell_wall_area(label, neighbors, real_area)
        kept = copy.copy(neighbors)
        for i, j in areas.keys():
            if areas[(i, j)] < min_contact_area:
                kept.remove(i if j == label else j)
        return kept

    def neighbor_kernels(self):
        if self._kernels is None:
            self._kernels = one_sided_kernels()
        return self._kernels

    def neighbors_number(self, labels=None, min_contact_area=None, real_area=True, verbose=True):
        # SIA:734-742
        nei = self.neighbors(labels, min_contact_area, real_area, verbose)
        if isinstance(nei, dict):
            return dict((k, len(v)) for k, v in nei.items())
        return len(nei)

    def get_voxel_face_surface(self):
        # SIA:751-756
        a = self._voxelsize
        return np.array([a[1] * a[2], a[2] * a[0], a[0] * a[1]])

    # ------------------------------------------------------- wall voxels SIA:759-880
    def wall_voxels_per_cell(self, label_1, bbox=None, neighbors=None, neighbors2ignore=[], verbose=False):
        if isinstance(bbox, dict):
            box = bbox[label_1]
        elif isinstance(bbox, (tuple, list)) and isinstance(bbox[0], slice):
            box = bbox
        else:
            box = self.boundingbox(label_1)
        big = grow_box(grow_box(box))
        sub = self.image[big]
        m1 = sub == label_1
        s18 = nd.generate_binary_structure(3, 2)
        d1 = nd.binary_dilation(m1, structure=s18)
        if neighbors is None:
            neighbors = self.neighbors(label_1)
        if isinstance(neighbors, int):
            neighbors = [neighbors]
        if isinstance(neighbors, dict) and len(neighbors) != 1:
            neighbors = copy.copy(neighbors[label_1])
        neighbors = list(neighbors)
        for nei in neighbors2ignore:
            if nei in neighbors:
                neighbors.remove(nei)
        coord = {}
        for label_2 in neighbors:
            m2 = sub == label_2
            d2 = nd.binary_dilation(m2, structure=s18)
            x, y, z = np.where((d1 & m2) | (d2 & m1))
            if len(x) > 0:
                coord[min(label_1, label_2), max(label_1, label_2)] = np.array(
                    (x + big[0].start, y + big[1].start, z + big[2].start))
        return coord

    def wall_voxels_between_two_cells(self, label_1, label_2, bbox=None, verbose=False):
        """SIA:759-804 with a tuple ``bbox`` (the only branch whose names are all defined);
        ``bbox=None`` falls back to the whole image as SIA:790-791 does."""
        if isinstance(bbox, dict):
            box = bbox[label_1] if label_1 in bbox else bbox[label_2]
        elif isinstance(bbox, tuple) and len(bbox) == 3:
            box = bbox
        else:
            box = None
        if box is not None:
            big = grow_box(box)
            sub = self.image[big]
            off = [s.start for s in big]
        else:
            sub = self.image
            off = [0, 0, 0]
        m1, m2 = sub == label_1, sub == label_2
        s18 = nd.generate_binary_structure(3, 2)
        d1 = nd.binary_dilation(m1, structure=s18)
        d2 = nd.binary_dilation(m2, structure=s18)
        x, y, z = np.where((d1 & m2) | (d2 & m1))
        return np.array((x + off[0], y + off[1], z + off[2]))

    # -------------------------------------------------------- wall areas SIA:908-993
    def cell_wall_area(self, label_id, neighbors, real=True):
        res = self.get_voxel_face_surface()
        try:
            sub = self.image[grow_box(self.boundingbox(label_id))]
        except Exception:
            sub = self.image
        inside = sub == label_id
        single = not isinstance(neighbors, list)
        if single:
            neighbors = [neighbors]
        wall = {}
        for a, kern in enumerate(self.neighbor_kernels()):
            dil = nd.binary_dilation(inside, structure=kern)
            frontier = sub[dil & ~inside]
            for n in neighbors:
                nb_pix = len(frontier[frontier == n])
                area = float(nb_pix * res[a // 2]) if real else nb_pix
                key = (min(label_id, n), max(label_id, n))
                wall[key] = wall.get(key, 0.0) + area
        if single:
            return next(iter(wall.values()))
        return wall

    def wall_areas(self, neighbors=None, real=True):
        if neighbors is None:
            neighbors = self.neighbors()
        areas = {}
        for label_id, lneighbors in neighbors.items():
            upper = [n for n in lneighbors if n > label_id]
            if len(upper) > 0:
                part = self.cell_wall_area(label_id, upper, real=real)
                for key in part:
                    areas[key] = areas.get(key, 0.0) + part[key]
        return areas

    # ------------------------------------------------------------ layers SIA:996-1046
    def cell_first_layer(self, filter_by_area=True, minimal_external_area=10, real_area=True):
        if self._cell_layer1 is None:
            self._cell_layer1 = list(map(int, self.neighbors(self.background())))
        layer = self._cell_layer1
        if filter_by_area:
            bg = self.background()
            area = self.cell_wall_area(bg, self._cell_layer1, real_area)
            layer = [l for l in self._cell_layer1
                     if (bg, l) in area and area[(bg, l)] > minimal_external_area]
        return list(set(layer) - self._ignoredlabels)

    def cell_second_layer(self, filter_by_area=True, minimal_L1_area=10, real_area=True):
        around = self.neighbors(self.cell_first_layer(), minimal_L1_area, real_area, True)
        l2 = set()
        for nei in around.values():
            l2.update(nei)
        self._cell_layer2 = list(l2 - set(self._cell_layer1) - self._ignoredlabels)
        return self._cell_layer2

    def voxel_first_layer(self, keep_background=True):
        if self._voxel_layer1 is None:
            bgmask = self.image == self.background()
            dil = nd.binary_dilation(bgmask, structure=nd.generate_binary_structure(3, 1))
            layer = dil & ~bgmask
            out = self.image * layer
            if keep_background:
                out = out + bgmask
            self._voxel_layer1 = out
        return self._voxel_layer1

    # ------------------------------------------ wall voxels per pair SIA:1049-1111
    def wall_voxels_per_cells_pairs(self, labels=None, neighborhood=None, only_epidermis=False,
                                    ignore_background=False, min_contact_area=None, real_area=True,
                                    verbose=True):
        # SIA:1062-1065: the first-voxel-layer image only ever supplies the label list (SIA:1073-1074); the voxels still
        # come from self.image.  That list is an ndarray (np.unique), so ``labels + [background]`` below is numpy's
        # element-wise sum, as in the reference.
        image = self.voxel_first_layer(True) if only_epidermis else self.image
        compute_neighborhood = neighborhood is None
        if isinstance(labels, list) and isinstance(neighborhood, dict):
            labels = [l for l in labels if l in neighborhood]
        if labels is None and not only_epidermis:
            labels = self.labels()
        elif labels is None and only_epidermis:
            labels = np.unique(image)
        elif isinstance(labels, list):
            labels.sort()
            if not isinstance(neighborhood, dict):
                compute_neighborhood = True
        elif isinstance(labels, int):
            labels = [labels]
        else:
            raise ValueError("Couldn't find any labels.")
        found = {}
        for label in labels:
            if compute_neighborhood:
                neighbors = list(self.neighbors(label, min_contact_area, real_area))
            elif isinstance(neighborhood, dict):
                neighbors = copy.copy(neighborhood[label])
            else:
                neighbors = list(neighborhood)
            allowed = labels if ignore_background else labels + [self.background()]
            skip = [n for n in neighbors if n not in allowed]
            neighbors = [n for n in neighbors
                         if (min(label, n), max(label, n)) not in found]
            if neighbors != []:
                found.update(self.wall_voxels_per_cell(label, self.boundingbox(label), neighbors, skip,
                                                       verbose=False))
        return found

    # ---------------------------------------------------------- volume SIA:1197-1243
    def volume(self, labels=None, real=True):
        labels = self.label_request(labels)
        vol = nd.sum(np.ones_like(self.image), self.image, index=np.asarray(labels, dtype=np.int64))
        if real is True:
            vol = np.multiply(vol, (self._voxelsize[0] * self._voxelsize[1] * self._voxelsize[2]))
        return self.convert_return(vol, labels)

    # ---------------------------------------------------- inertia_axis SIA:1246-1292
    def inertia_axis(self, labels=None, real=True, verbose=False):
        labels = self.label_request(labels)
        vecs, vals = [], []
        for label in labels:
            box = self.boundingbox(label, real=False)
            center = copy.copy(self.center_of_mass(label, real=False))
            if box is not None:
                center = [center[i] - s.start for i, s in enumerate(box)]
                inside = self.image[box] == label
            else:
                inside = self.image == label
            xyz = inside.nonzero()
            pts = np.array([xyz[0] - center[0], xyz[1] - center[1], xyz[2] - center[2]])  # SIA:123-135
            w, v = sorted_eig(covariance_of(pts))
            if real:
                for i in range(3):
                    w[i] *= np.linalg.norm(np.multiply(v[i], self._voxelsize))
            vecs.append(v)
            vals.append(w)
        by_rows = [[v[i] for i in range(len(v))] for v in vecs]  # SIA:191-201
        if len(labels) == 1:
            return by_rows[0], vals[0]
        return self.convert_return(by_rows, labels), self.convert_return(vals, labels)

    reduced_inertia_axis = inertia_axis  # SIA:1295-1341 is the same computation

    # ------------------------------------------------- remaining voxel stencils
    def get_all_wall_binary_image(self):
        # SIA:744-749
        lp = nd.laplace(self.image)
        with np.errstate(divide="ignore", invalid="ignore"):
            return lp / lp

    def cells_walls_coords(self):
        # SIA:883-905; `self.background` (the bound method) is what the reference passes -> nothing is removed
        m = hollow_out_cells(self.image, None, remove_background=False)
        x, y, z = np.where(m != 0)
        return list(x), list(y), list(z)

    def region_boundingbox(self, labels):
        # SIA:1361-1396
        if isinstance(labels, list) and len(labels) == 1:
            return self.boundingbox(labels[0])
        if isinstance(labels, int):
            return self.boundingbox(labels)
        boxes = self.boundingbox(labels)
        lo = [min(boxes[c][a].start for c in labels) for a in range(3)]
        hi = [max(boxes[c][a].stop for c in labels) for a in range(3)]
        return tuple(slice(a, b) for a, b in zip(lo, hi))

    def cells_voxel_layer(self, labels, region_boundingbox=False, single_frame=False):
        # SIA:1399-1448
        if isinstance(labels, int):
            labels = [labels]
        if single_frame:
            region_boundingbox = True
        if region_boundingbox:
            bbox = self.region_boundingbox(labels)
        else:
            bboxes = self.boundingbox(labels, real=False)
        s18 = nd.generate_binary_structure(3, 2)
        out = np.zeros_like(self.image[bbox], dtype=int) if single_frame else {}
        for c in labels:
            sub = self.image[bbox] if region_boundingbox else self.image[bboxes[c]]
            mask = sub == c
            layer = np.array(mask & ~nd.binary_erosion(mask, structure=s18), dtype=int)
            if single_frame:
                out += layer
            else:
                out[c] = layer
        if len(labels) == 1:
            return out[c]
        return out

    # ------------------------------------------ labels_at_stack_margins SIA:1344-1358
    def labels_at_stack_margins(self, voxel_distance_from_margin=5):
        d = voxel_distance_from_margin
        im = self.image
        seen = []
        for sl in (np.s_[:d, :, :], np.s_[-d:, :, :], np.s_[:, :d, :], np.s_[:, -d:, :],
                   np.s_[:, :, :d], np.s_[:, :, -d:]):
            seen.extend(np.unique(im[sl]))
        return list(set(seen) - set([self._background]))

"""graph_from_image on oracle-fed tables == the restated reference graph builder; array form consistent."""
import warnings

import numpy as np

from oracle.graph_loops import graph_from_image_oracle
from tests.helpers import OracleBackend, assert_eig_close
from tissue_analysis_b200 import SpatialImageAnalysis3D
from tissue_analysis_b200.synth import tissue_image
from tissue_analysis_b200.temporal_graph_from_image import graph_arrays, graph_from_image

warnings.filterwarnings("ignore")
PROPS = ['boundingbox', 'volume', 'barycenter', 'L1', 'border', 'inertia_axis', 'wall_surface', 'epidermis_surface']


def compare_graph(g, o):
    assert sorted(g.vertices()) == sorted(o["vertices"])
    got_edges = dict(((g.source(e), g.target(e)), e) for e in g.edges())
    assert set(got_edges) == set(o["edges"])
    vp, ep = o["vertex_properties"], o["edge_properties"]
    for name in ("label", "boundingbox", "volume", "L1", "border", "unlabelled_wall_surface", "epidermis_surface"):
        if name in vp:
            assert dict(g.vertex_property(name)) == dict(vp[name]), name
    for l, c in vp["barycenter"].items():
        assert np.array_equal(np.asarray(g.vertex_property("barycenter")[l]), np.asarray(c))
    for l in vp["inertia_values"]:
        assert_eig_close(g.vertex_property("inertia_axis")[l], g.vertex_property("inertia_values")[l],
                         vp["inertia_axis"][l], np.real(vp["inertia_values"][l]))
    ws = dict(((g.source(e), g.target(e)), v) for e, v in g.edge_property("wall_surface").items())
    inv = dict((eid, k) for k, eid in o["edges"].items())
    assert ws == dict((inv[e], v) for e, v in ep["wall_surface"].items())


def test_graph_default_and_filtered():
    img = tissue_image((44, 40, 36), 45, seed=21, dome=True, voxelsize=(0.4, 0.4, 1.0), weights=(2, 2, 5))
    for kw in (dict(), dict(ignore_cells_at_stack_margins=False, min_contact_area=2.0),
               dict(labels=[5, 9, 12, 17, 20, 23, 31], ignore_cells_at_stack_margins=False)):
        okw = dict(kw)
        if "labels" in okw:
            okw["labels"] = list(okw["labels"])
        g = graph_from_image(img, spatio_temporal_properties=PROPS, _backend=OracleBackend(img),
                             **dict(kw, labels=list(kw["labels"]) if "labels" in kw else None))
        o = graph_from_image_oracle(np.asarray(img), properties=PROPS, voxelsize=img.voxelsize, **okw)
        compare_graph(g, o)


def test_graph_arrays_match_dict_graph():
    img = tissue_image((44, 40, 36), 45, seed=22, dome=True, voxelsize=(0.4, 0.4, 1.0), weights=(2, 2, 5))
    g = graph_from_image(img, spatio_temporal_properties=PROPS, ignore_cells_at_stack_margins=False,
                         _backend=OracleBackend(img))
    sia2 = SpatialImageAnalysis3D(img, ignoredlabels=0, background=1, _backend=OracleBackend(img))
    arr = graph_arrays(sia2, ignore_cells_at_stack_margins=False)
    assert sorted(g.vertices()) == arr.labels.tolist()
    ws = dict(((g.source(e), g.target(e)), v) for e, v in g.edge_property("wall_surface").items())
    assert ws == dict(((int(a), int(b)), v) for a, b, v in zip(arr.edge_lo, arr.edge_hi, arr.wall_surface))
    for i, l in enumerate(arr.labels.tolist()):
        assert g.vertex_property("volume")[l] == arr.volume[i]
        assert np.array_equal(np.asarray(g.vertex_property("barycenter")[l]), arr.barycenter[i])
        assert g.vertex_property("L1")[l] == bool(arr.L1[i]) and g.vertex_property("border")[l] == bool(arr.border[i])
        nb = sorted(arr.labels[arr.indices[arr.indptr[i]:arr.indptr[i + 1]]].tolist())
        assert nb == sorted(g.neighbors(l))

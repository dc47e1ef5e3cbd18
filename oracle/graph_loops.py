"""CPU ORACLE (test infrastructure, never shipped, never on the product path).

Python-3 restatement of the reference's graph builder ``_graph_from_image``
(/root/reference/src/vplants/tissue_analysis/temporal_graph_from_image.py:77-244, ``generate_graph_topology``
:30-60) driving the loop oracle (oracle/sia_loops.py).  ``openalea.container.PropertyGraph`` is absent, so the result is
returned as plain dictionaries: vertices (ids = labels, :45), edges {(source, target): edge id in insertion order},
vertex / edge properties keyed by those ids.  Parity unpinned: the reference has no test or example output for this
function; the property values are those of sia_loops.py, which is pinned to the reference's docstring vectors.
'wall_surface' / 'epidermis_surface' follow the evidently intended meaning of the undefined helpers at :186, :205, :207.
"""
from oracle.sia_loops import DICT, LoopOracle


def graph_from_image_oracle(image, labels=None, background=1, properties=None, property_as_real=True,
                            ignore_cells_at_stack_margins=True, min_contact_area=None, voxelsize=None):
    if properties is None:
        properties = ['L1', 'L2', 'barycenter', 'border', 'boundingbox', 'epidermis_area', 'inertia_axis', 'volume',
                      'wall_area']
    analysis = LoopOracle(image, ignoredlabels=0, return_type=DICT, background=1, voxelsize=voxelsize)        # :109
    if ignore_cells_at_stack_margins:
        analysis.add2ignoredlabels([int(x) for x in analysis.labels_at_stack_margins()])                      # :114
    if labels is None:
        labels = list(analysis.labels())                                                                       # :118
        if background in labels:
            del labels[labels.index(background)]
    else:
        if isinstance(labels, int):
            labels = [labels]
        if background in labels:
            labels.remove(background)
        analysis.add2ignoredlabels(set(analysis.labels()) - set(labels))                                      # :127
    neighborhood = analysis.neighbors(labels, min_contact_area=min_contact_area, verbose=False)               # :129
    labelset = set(labels)
    vertices = list(labels)                                                                                    # :45
    edges, next_eid = {}, 0
    for source, targets in neighborhood.items():                                                              # :51-55
        if source in labelset:
            for target in targets:
                if source < target and target in labelset:
                    edges[(int(source), int(target))] = next_eid
                    next_eid += 1
    vprop, eprop = {'label': dict((l, l) for l in labels)}, {}
    if 'boundingbox' in properties:
        vprop['boundingbox'] = dict(analysis.boundingbox(labels, real=property_as_real))                     # :142
    if 'volume' in properties:
        vprop['volume'] = dict(analysis.volume(labels, real=property_as_real))                                # :147
    barycenters = None
    if 'barycenter' in properties:
        barycenters = analysis.center_of_mass(labels, real=property_as_real)                                  # :153
        vprop['barycenter'] = dict(barycenters)
    bg_nb = set(int(x) for x in analysis.neighbors(background)) & labelset                                    # :157-158
    if 'L1' in properties:
        vprop['L1'] = dict((l, l in bg_nb) for l in labels)                                                   # :161
    if 'border' in properties:
        border = set(int(x) for x in analysis.labels_at_stack_margins()) - set([background])                  # :166-169
        vprop['border'] = dict((l, l in border) for l in labels)
    if 'inertia_axis' in properties:
        axes, vals = analysis.inertia_axis(labels, barycenters)                                               # :174
        vprop['inertia_axis'], vprop['inertia_values'] = dict(axes), dict(vals)
    if 'wall_surface' in properties:
        filtered, unlabelled = {}, {}
        for source, targets in neighborhood.items():                                                          # :181-184
            if source in labelset:
                filtered[source] = [t for t in targets if source < t and t in labelset]
                unlabelled[source] = [t for t in targets if t not in labelset and t != background]
        ws = analysis.wall_areas(filtered, real=property_as_real)                                             # :185
        eprop['wall_surface'] = dict((edges[(int(a), int(b))], v) for (a, b), v in ws.items())
        vprop['unlabelled_wall_surface'] = dict(
            (s, sum(analysis.wall_areas({s: unlabelled[s]}, real=property_as_real).values())) for s in unlabelled)
    if 'epidermis_surface' in properties:
        nb = sorted(bg_nb)
        surf = analysis.cell_wall_area(background, nb, real=property_as_real) if nb else {}
        vprop['epidermis_surface'] = dict(((int(a) if b == background else int(b)), v) for (a, b), v in surf.items())
    return dict(vertices=vertices, edges=edges, vertex_properties=vprop, edge_properties=eprop)

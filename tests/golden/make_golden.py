"""Writes the golden fixtures of this directory.  Run from the repository root: python tests/golden/make_golden.py

* reference_docstring_vectors.json -- the ONLY results the reference itself pins: the docstring examples of
  src/vplants/tissue_analysis/spatial_image_analysis.py on its 4 x 6 toy image (the reference cannot be imported here:
  Python 2 syntax, openalea absent -- the values are transcribed from the cited lines, not computed).
* tables_*.npz -- per-label and per-pair tables of small seeded synthetic tissues, computed by the CPU oracle
  (oracle/sia_onepass.py, proven equal to the loop restatement oracle/sia_loops.py by tests/test_host_mirror_cpu.py) with
  the scipy / numpy of this image.  They freeze today's oracle output: a later scipy / numpy that changes a result shows
  up as a fixture mismatch, and the `-m gpu` run compares the CUDA tables with them without importing the oracle.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

DOCSTRING = {
    "source": "src/vplants/tissue_analysis/spatial_image_analysis.py (docstring examples)",
    "image_4x6": [[1, 2, 7, 7, 1, 1], [1, 6, 5, 7, 3, 3], [2, 2, 1, 7, 3, 3], [1, 1, 1, 4, 1, 1]],
    "labels": {"lines": "343-353", "value": [1, 2, 3, 4, 5, 6, 7]},
    "center_of_mass": {"lines": "437-450", "value": {
        "1": [1.8, 2.2999999999999998, 0.0], "2": [1.3333333333333333, 0.66666666666666663, 0.0], "3": [1.5, 4.5, 0.0],
        "4": [3.0, 3.0, 0.0], "5": [1.0, 2.0, 0.0], "6": [1.0, 1.0, 0.0], "7": [0.75, 2.75, 0.0]}},
    "boundingbox": {"lines": "498-511", "value": {
        "1": [[0, 4], [0, 6], [0, 1]], "2": [[0, 3], [0, 2], [0, 1]], "3": [[1, 3], [4, 6], [0, 1]],
        "4": [[3, 4], [3, 4], [0, 1]], "5": [[1, 2], [2, 3], [0, 1]], "6": [[1, 2], [1, 2], [0, 1]],
        "7": [[0, 3], [2, 4], [0, 1]]}},
    "neighbors": {"lines": "561-574", "value": {
        "1": [2, 3, 4, 5, 6, 7], "2": [1, 6, 7], "3": [1, 7], "4": [1, 7], "5": [1, 6, 7], "6": [1, 2, 5],
        "7": [1, 2, 3, 4, 5]}},
    "cell_wall_area_7": {"lines": "924-927", "value": {"2,7": 1.0, "5,7": 2.0}},
    "wall_areas": {"lines": "978-982", "value": {
        "1,2": 5.0, "1,3": 4.0, "1,4": 2.0, "1,5": 1.0, "1,6": 1.0, "1,7": 2.0, "2,6": 2.0, "2,7": 1.0, "3,7": 2.0,
        "4,7": 1.0, "5,6": 1.0, "5,7": 2.0}},
    "volume": {"lines": "1219-1226", "value": {"1": 10.0, "2": 3.0, "3": 4.0, "4": 1.0, "5": 1.0, "6": 1.0, "7": 4.0}},
}

CASES = {
    # name: (shape xyz, cells, seed, dome, dtype, voxelsize)
    "dome_u16": ((40, 36, 30), 40, 11, True, "uint16", (1.0, 1.0, 1.0)),
    "aniso_u32": ((33, 20, 17), 25, 5, True, "uint32", (0.2, 0.2, 0.5)),
    "ragged_u16": ((131, 9, 5), 12, 3, False, "uint16", (1.0, 1.0, 1.0)),
}


def main():
    from oracle import sia_onepass
    from tissue_analysis_b200.synth import tissue_image
    with open(os.path.join(HERE, "reference_docstring_vectors.json"), "w") as f:
        json.dump(DOCSTRING, f, indent=1, sort_keys=True)
    for name, (shape, ncell, seed, dome, dtype, vox) in CASES.items():
        img = np.asarray(tissue_image(shape, ncell, seed, dome=dome, dtype=dtype, voxelsize=vox))
        lt = sia_onepass.label_table(img)
        pt = sia_onepass.pair_table(img)
        n = int(np.nonzero(lt["count"])[0].max()) + 1
        np.savez_compressed(os.path.join(HERE, "tables_%s.npz" % name), image=img, voxelsize=np.array(vox),
                            count=lt["count"][:n], s1=lt["s1"][:n], s2=lt["s2"][:n], bmin=lt["bmin"][:n],
                            bmax=lt["bmax"][:n], pair_lo=pt["lo"], pair_hi=pt["hi"], faces=pt["faces"], wall18=pt["wall18"])
        print(name, img.shape, img.dtype, "labels", int((lt["count"] > 0).sum()), "pairs", len(pt["lo"]))


if __name__ == "__main__":
    main()

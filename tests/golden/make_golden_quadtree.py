"""Generates tests/golden/quadtree_small.npz from the UNMODIFIED reference (2-D TreeMesh): the cases of
quadtree_cases.py on four QuadTrees built with refine_ball / insert_cells / refine_box / refine_points / refine."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.refload import load_reference  # noqa: E402
import _pack  # noqa: E402
import quadtree_cases as qc  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    out = {}
    for tag, h, origin, kind in qc.specs():
        mesh = qc.build(ref, h, origin, kind)
        pts = qc.points_for(mesh, h, origin)
        out[f"{tag}|points"] = pts
        for k, v in qc.evaluate(mesh, pts).items():
            out[f"{tag}|{k}"] = v
    # diagonal balancing changes the tree: pin the cell list of one mesh built with it
    tag, h, origin, kind = qc.specs()[1]
    mesh = qc.build(ref, h, origin, kind, diagonal_balance=True)
    out["diag|cell_centers"], out["diag|levels"] = mesh.cell_centers, np.asarray(mesh._cell_levels_by_indexes())
    path = os.path.join(HERE, "quadtree_small.npz")
    _pack.save(path, out)
    print(len(out), "entries ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

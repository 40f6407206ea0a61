"""Generates tests/golden/boundary_small.npz from the UNMODIFIED reference: weak-form boundary pieces, boundary
integrals / projections, surface and line inner products, and the small geometry conveniences, on 1-D / 2-D / 3-D
TensorMesh and two OcTrees.  The cases live in boundary_cases.py, shared with tests/test_boundary_ops.py."""
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
import boundary_cases as bc  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    out = {}
    for tag, h, origin in bc.tensor_specs():
        for k, v in bc.evaluate(ref.TensorMesh(h, origin=origin)).items():
            out[f"{tag}|{k}"] = v
    for tag, h, origin in bc.tree_specs():
        mesh = bc.build_tree(ref.TreeMesh, h, origin)
        out[f"{tag}|n_cells"] = np.array(mesh.n_cells)
        for k, v in bc.evaluate(mesh, tree=True).items():
            out[f"{tag}|{k}"] = v
    path = os.path.join(HERE, "boundary_small.npz")
    _pack.save(path, out)
    print(len(out), "entries ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

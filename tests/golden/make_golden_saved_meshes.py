"""Writes tests/golden/saved_*.json with the UNMODIFIED reference's ``mesh.save()``: a TensorMesh, an OcTree and a QuadTree
(the meshes of boundary_cases.py / quadtree_cases.py), for the cross-library load test in tests/test_persistence.py."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.refload import load_reference  # noqa: E402
import boundary_cases as bc  # noqa: E402
import quadtree_cases as qc  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    tag, h, origin = bc.tensor_specs()[0]
    ref.TensorMesh(h, origin=origin).save(os.path.join(HERE, "saved_tensor.json"))
    tag, h, origin = bc.tree_specs()[1]
    bc.build_tree(ref.TreeMesh, h, origin).save(os.path.join(HERE, "saved_octree.json"))
    tag, h, origin, kind = qc.specs()[2]
    qc.build(ref, h, origin, kind).save(os.path.join(HERE, "saved_quadtree.json"))
    for f in ("saved_tensor.json", "saved_octree.json", "saved_quadtree.json"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()

"""Generates tests/golden/tree_interp_small.npz from the UNMODIFIED reference: TreeMesh.get_interpolation_matrix for
nodes, x/y/z edges, x/y/z faces and cell centres on the two OcTrees of boundary_cases.py, at random points (some
outside the mesh) mixed with points placed exactly on cell centres, nodes, faces and edges -- the ties that drive the
reference's order-dependent axis ordering (tree_ext.pyx:5961-5969, 6240-6262).  The points are stored in the order
they were given to the reference; that order is part of the result."""
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.refload import load_reference  # noqa: E402
import boundary_cases as bc  # noqa: E402

LOCS = ["N", "Ex", "Ey", "Ez", "Fx", "Fy", "Fz", "CC"]


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    rng = np.random.default_rng(2024)
    out = {}
    for tag, h, origin in bc.tree_specs():
        mesh = bc.build_tree(ref.TreeMesh, h, origin)
        ext = np.array([x.sum() for x in h])
        pts = rng.random((500, 3)) * ext * 1.2 - 0.1 * ext + origin
        adv = np.r_[mesh.cell_centers[rng.integers(0, mesh.n_cells, 60)], mesh.nodes[rng.integers(0, mesh.n_nodes, 60)],
                    mesh.faces_x[rng.integers(0, mesh.n_faces_x, 40)], mesh.edges_y[rng.integers(0, mesh.n_edges_y, 40)]]
        pts = np.r_[pts, adv][rng.permutation(700)]
        out[f"{tag}|points"] = pts
        out[f"{tag}|inside"] = mesh.is_inside(pts)
        for lt in LOCS:
            A = sp.csr_matrix(mesh.get_interpolation_matrix(pts, lt)).copy()
            A.sum_duplicates()
            A.sort_indices()
            out[f"{tag}|{lt}|indptr"] = A.indptr.astype(np.int32)
            out[f"{tag}|{lt}|indices"] = A.indices.astype(np.int32)
            out[f"{tag}|{lt}|data"] = A.data
            out[f"{tag}|{lt}|shape"] = np.array(A.shape)
    path = os.path.join(HERE, "tree_interp_small.npz")
    np.savez_compressed(path, **out)
    print(len(out), "entries ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

"""Generates tests/golden/tree_n2_small.npz (OcTree node->edge, node->face, edge->face averages) from the UNMODIFIED
reference, on the two trees of tree_small.npz (rebuilt from their stored refinement points)."""
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refload import load_reference  # noqa: E402

TREE_N2_OPS = ["average_node_to_edge", "average_node_to_face", "average_edge_to_face"]
# cell-centred side (tree_ext.pyx:3471-3853, 5080-5512; tree_mesh.py:846-982)
TREE_CC_OPS = ["face_x_divergence", "face_y_divergence", "face_z_divergence", "stencil_cell_gradient",
               "stencil_cell_gradient_x", "stencil_cell_gradient_y", "stencil_cell_gradient_z", "average_cell_to_face",
               "average_cell_to_face_x", "average_cell_to_face_y", "average_cell_to_face_z", "average_cell_vector_to_face"]
TREE_CC_COMPOSITES = ["cell_gradient", "cell_gradient_x", "cell_gradient_y", "cell_gradient_z"]
TREE_CC_METHODS = ["average_cell_to_total_face_x", "average_cell_to_total_face_y", "average_cell_to_total_face_z"]


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tree_small.npz"))
    out = {}
    for tag in ("a", "b"):
        h = [g[f"{tag}_h{a}"] for a in "xyz"]
        mesh = ref.TreeMesh(h, origin=g[f"{tag}_origin"], diagonal_balance=bool(g[f"{tag}_db"]))
        mesh.refine_points(g[f"{tag}_pts"], -1, [1, 1], finalize=True)
        idx, lev = mesh.__getstate__()
        assert np.array_equal(idx, g[f"{tag}_state_indexes"]) and np.array_equal(lev, g[f"{tag}_state_levels"])
        mats = [(name, getattr(mesh, name)) for name in TREE_N2_OPS + TREE_CC_OPS + TREE_CC_COMPOSITES]
        mats += [(name, getattr(mesh, name)()) for name in TREE_CC_METHODS]
        for k, ind in enumerate(mesh.face_boundary_indices):
            out[f"{tag}_face_boundary_{k}"] = np.asarray(ind, dtype=np.int64)
        # node interpolation and point location (tree_ext.pyx:6116-6179, tree.cpp:758-766): random points, some outside
        # the mesh, some exactly on nodes / cell centres
        rng = np.random.default_rng(123 + len(out))
        ext = np.array([x.sum() for x in h])
        ipts = g[f"{tag}_origin"] + ext * (rng.random((120, 3)) * 1.2 - 0.1)
        ipts[:20] = mesh.nodes[rng.integers(0, mesh.nN, 20)]
        ipts[20:40] = mesh.cell_centers[rng.integers(0, mesh.nC, 20)]
        out[f"{tag}_interp_pts"] = ipts
        out[f"{tag}_point2index"] = np.asarray(mesh.point2index(ipts), dtype=np.int64)
        mats += [("interp_N", mesh.get_interpolation_matrix(ipts, "N")),
                 ("interp_N_zeros_outside", mesh.get_interpolation_matrix(ipts, "N", zeros_outside=True))]
        # full-tensor model derivatives (operators/inner_products.py:459-565 with the OcTree corner projections)
        mats += [("dMf_tensor", mesh.get_face_inner_product_deriv(g[f"{tag}_sigma6"])(g[f"{tag}_vF"])),
                 ("dMe_tensor", mesh.get_edge_inner_product_deriv(g[f"{tag}_sigma6"])(g[f"{tag}_vE"]))]
        for name, mat in mats:
            m = sp.csr_matrix(mat).copy()
            m.sum_duplicates()
            m.sort_indices()
            out[f"{tag}_{name}__indptr"] = m.indptr.astype(np.int64)
            out[f"{tag}_{name}__indices"] = m.indices.astype(np.int64)
            out[f"{tag}_{name}__data"] = m.data
            out[f"{tag}_{name}__shape"] = np.array(m.shape, dtype=np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tree_n2_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

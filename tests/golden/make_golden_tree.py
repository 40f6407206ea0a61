"""Generates tests/golden/tree_small.npz from the UNMODIFIED reference (run in the build container)."""
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refload import load_reference  # noqa: E402

TREE_OPS = ["face_divergence", "edge_curl", "nodal_gradient", "average_face_to_cell", "average_face_to_cell_vector",
            "average_edge_to_cell", "average_edge_to_cell_vector", "average_node_to_cell", "average_face_x_to_cell",
            "average_face_y_to_cell", "average_face_z_to_cell", "average_edge_x_to_cell", "average_edge_y_to_cell",
            "average_edge_z_to_cell"]
COUNTS = ["nC", "nN", "nhN", "ntN", "nEx", "nEy", "nEz", "nhEx", "nhEy", "nhEz", "nFx", "nFy", "nFz", "nhFx", "nhFy", "nhFz"]


def canon(m):
    m = sp.csr_matrix(m).copy()
    m.sum_duplicates()
    m.sort_indices()
    return m


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    rng = np.random.default_rng(77)
    out = {}
    for tag, shape, db in (("a", (16, 16, 16), True), ("b", (16, 8, 32), False)):
        h = [rng.random(s) + 0.5 for s in shape]
        origin = np.array([-1.0, 2.0, 0.5])
        ext = np.array([x.sum() for x in h])
        pts = origin + (rng.random((6, 3)) * 0.7 + 0.15) * ext
        mesh = ref.TreeMesh(h, origin=origin, diagonal_balance=db)
        mesh.refine_points(pts, -1, [1, 1], finalize=True)
        idx, lev = mesh.__getstate__()
        for a, nm in enumerate("xyz"):
            out[f"{tag}_h{nm}"] = h[a]
        out[f"{tag}_origin"], out[f"{tag}_pts"], out[f"{tag}_db"] = origin, pts, np.array(int(db))
        out[f"{tag}_state_indexes"], out[f"{tag}_state_levels"] = idx, lev
        out[f"{tag}_counts"] = np.array([getattr(mesh, c) for c in COUNTS], dtype=np.int64)
        out[f"{tag}_cell_volumes"], out[f"{tag}_face_areas"], out[f"{tag}_edge_lengths"] = (
            mesh.cell_volumes, mesh.face_areas, mesh.edge_lengths)
        out[f"{tag}_cell_centers"], out[f"{tag}_nodes"] = mesh.cell_centers, mesh.nodes

        def put(name, m):
            m = canon(m)
            out[f"{tag}_{name}__indptr"] = m.indptr.astype(np.int64)
            out[f"{tag}_{name}__indices"] = m.indices.astype(np.int64)
            out[f"{tag}_{name}__data"] = m.data
            out[f"{tag}_{name}__shape"] = np.array(m.shape, dtype=np.int64)

        for name in TREE_OPS:
            put(name, getattr(mesh, name))
        sig = rng.random(mesh.nC) + 0.5
        sig3 = rng.random((mesh.nC, 3)) + 0.5
        out[f"{tag}_sigma"], out[f"{tag}_sigma3"] = sig, sig3
        put("Mf_iso", mesh.get_face_inner_product(sig))
        put("Me_iso", mesh.get_edge_inner_product(sig))
        put("Me_iso_inv", mesh.get_edge_inner_product(sig, invert_model=True, invert_matrix=True))
        put("Mf_aniso", mesh.get_face_inner_product(sig3))
        put("Me_aniso", mesh.get_edge_inner_product(sig3))
        sig6 = np.c_[rng.random((mesh.nC, 3)) + 2.0, rng.random((mesh.nC, 3)) * 0.5 - 0.25]
        out[f"{tag}_sigma6"] = sig6
        put("Mf_tensor", mesh.get_face_inner_product(sig6))
        put("Me_tensor_inv", mesh.get_edge_inner_product(sig6, invert_model=True))
        vF, vE = rng.standard_normal(mesh.nF), rng.standard_normal(mesh.nE)
        out[f"{tag}_vF"], out[f"{tag}_vE"] = vF, vE
        put("dMf_iso", mesh.get_face_inner_product_deriv(sig)(vF))
        put("dMe_iso_inv", mesh.get_edge_inner_product_deriv(sig, invert_model=True, invert_matrix=True)(vE))
        put("dMe_aniso", mesh.get_edge_inner_product_deriv(sig3)(vE))
        put("dMf_aniso_invmat", mesh.get_face_inner_product_deriv(sig3, invert_matrix=True)(vF))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tree_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

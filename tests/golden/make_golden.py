"""Generates tests/golden/tensor_small.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref built by
oracle/build_ref.sh):   python tests/golden/make_golden.py
The fixture pins the numpy oracle and the CUDA path on machines without the reference.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refload import load_reference  # noqa: E402


def canon(m):
    m = sp.csr_matrix(m).copy()
    m.sum_duplicates()
    m.sort_indices()
    return m


def main():
    ref = load_reference()
    assert ref is not None, "reference not available"
    rng = np.random.default_rng(2024)
    h = [rng.random(4) + 0.5, rng.random(6) + 0.5, rng.random(5) + 0.5]
    origin = np.array([-1.0, 0.25, 3.0])
    mesh = ref.TensorMesh(h, origin=origin)
    out = {"hx": h[0], "hy": h[1], "hz": h[2], "origin": origin}

    def put(name, m):
        m = canon(m)
        out[name + "__indptr"] = m.indptr.astype(np.int64)
        out[name + "__indices"] = m.indices.astype(np.int64)
        out[name + "__data"] = m.data
        out[name + "__shape"] = np.array(m.shape, dtype=np.int64)

    for name in ["face_divergence", "edge_curl", "nodal_gradient", "average_face_to_cell",
                 "average_face_to_cell_vector", "average_edge_to_cell", "average_edge_to_cell_vector",
                 "average_node_to_cell", "average_face_x_to_cell", "average_face_y_to_cell",
                 "average_face_z_to_cell", "average_edge_x_to_cell", "average_edge_y_to_cell",
                 "average_edge_z_to_cell"]:
        put(name, getattr(mesh, name))

    nC = mesh.nC
    models = {
        "iso": rng.random(nC) + 0.5,
        "aniso": rng.random((nC, 3)) + 0.5,
        "tensor": np.c_[rng.random((nC, 3)) + 2.0, rng.random((nC, 3)) * 0.5 - 0.25],
    }
    for k, v in models.items():
        out["model_" + k] = v
    for proj, get in (("F", mesh.get_face_inner_product), ("E", mesh.get_edge_inner_product)):
        for mk, m in models.items():
            for im in (False, True):
                for ix in (False, True):
                    if mk == "tensor" and ix:
                        continue
                    put(f"mass_{proj}_{mk}_{int(im)}{int(ix)}", get(m, invert_model=im, invert_matrix=ix))
    v = {"F": rng.standard_normal(mesh.nF), "E": rng.standard_normal(mesh.nE)}
    out["deriv_v_F"], out["deriv_v_E"] = v["F"], v["E"]
    for proj, get in (("F", mesh.get_face_inner_product_deriv), ("E", mesh.get_edge_inner_product_deriv)):
        for mk, m in models.items():
            for im in (False, True):
                for ix in (False, True):
                    if mk == "tensor" and (ix or im):
                        continue
                    put(f"deriv_{proj}_{mk}_{int(im)}{int(ix)}", get(m, invert_model=im, invert_matrix=ix)(v[proj]))

    # composites
    G, Cc = mesh.nodal_gradient, mesh.edge_curl
    put("dc_nodal", G.T @ mesh.get_edge_inner_product(models["iso"]) @ G)
    mui = rng.random(nC) + 0.5
    out["model_mui"] = mui
    put("maxwell", Cc.T @ mesh.get_face_inner_product(mui) @ Cc + mesh.get_edge_inner_product(models["iso"]))

    # interpolation, incl. adversarial points (on nodes, centres, boundary, outer half cells)
    lo = np.array([mesh.nodes_x[0], mesh.nodes_y[0], mesh.nodes_z[0]])
    hi = np.array([mesh.nodes_x[-1], mesh.nodes_y[-1], mesh.nodes_z[-1]])
    pts = lo + rng.random((64, 3)) * (hi - lo)
    k = 0
    for a, (nodes, cc) in enumerate(((mesh.nodes_x, mesh.cell_centers_x), (mesh.nodes_y, mesh.cell_centers_y),
                                     (mesh.nodes_z, mesh.cell_centers_z))):
        for val in (nodes[0], nodes[-1], nodes[len(nodes) // 2], cc[0], cc[-1],
                    0.5 * (nodes[0] + cc[0]), 0.5 * (nodes[-1] + cc[-1])):
            pts[k, a] = val
            k += 1
    pts[k], pts[k + 1] = lo, hi
    out["interp_pts"] = pts
    for loc in ["CC", "N", "Fx", "Fy", "Fz", "Ex", "Ey", "Ez", "CCVx", "CCVy", "CCVz"]:
        put("interp_" + loc, mesh.get_interpolation_matrix(pts, loc))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tensor_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()

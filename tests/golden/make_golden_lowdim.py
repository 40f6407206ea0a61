"""Generates tests/golden/tensor_lowdim_small.npz (1-D and 2-D TensorMesh operators) from the UNMODIFIED reference."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refload import load_reference  # noqa: E402

OPS_ANY = ["face_divergence", "face_x_divergence", "nodal_gradient", "average_face_to_cell", "average_face_to_cell_vector",
           "average_edge_to_cell", "average_edge_to_cell_vector", "average_face_x_to_cell", "average_edge_x_to_cell",
           "average_node_to_cell", "average_cell_to_face", "average_cell_vector_to_face", "average_node_to_edge",
           "average_node_to_face", "average_cell_to_edge", "average_edge_to_face", "stencil_cell_gradient_x", "cell_gradient_x"]
OPS_2D = ["face_y_divergence", "edge_curl", "average_face_y_to_cell", "average_edge_y_to_cell", "stencil_cell_gradient_y",
          "cell_gradient_y"]
BCS = {"neumann": "neumann", "dirichlet": "dirichlet", "mixed": [["dirichlet", "neumann"], "dirichlet"]}


def bc_for(dim, tag):
    bc = BCS[tag]
    return bc if isinstance(bc, str) else list(bc[:dim])
LOCS = {1: ["CC", "N", "Fx", "Ex", "CCVx"], 2: ["CC", "N", "Fx", "Fy", "Ex", "Ey", "CCVx", "CCVy"]}
MASS_CASES = [(mk, im, ix) for mk in ("iso", "aniso") for im in (False, True) for ix in (False, True)]


def ops_for(dim):
    return OPS_ANY + (OPS_2D if dim == 2 else [])


def main():
    ref = load_reference()
    assert ref is not None
    rng = np.random.default_rng(41)
    out = {}
    for tag, h, origin in (("d2", [rng.random(5) + .5, rng.random(7) + .5], np.array([0.3, -1.0])),
                           ("d1", [rng.random(9) + .5], np.array([2.0]))):
        mesh = ref.TensorMesh(h, origin=origin)
        d = mesh.dim
        for a in range(d):
            out[f"{tag}_h{a}"] = h[a]
        out[f"{tag}_origin"] = origin

        def put(name, m):
            m = sp.csr_matrix(m).copy()
            m.sum_duplicates()
            m.sort_indices()
            out[f"{tag}_{name}__indptr"], out[f"{tag}_{name}__indices"] = m.indptr.astype(np.int64), m.indices.astype(np.int64)
            out[f"{tag}_{name}__data"], out[f"{tag}_{name}__shape"] = m.data, np.array(m.shape, dtype=np.int64)

        for name in ops_for(d):
            put(name, getattr(mesh, name))
        for arr in ("cell_volumes", "face_areas", "edge_lengths"):
            out[f"{tag}_{arr}"] = getattr(mesh, arr)
        import copy
        for bct in BCS:
            mesh.set_cell_gradient_BC(copy.deepcopy(bc_for(d, bct)))
            put("stencil_cell_gradient_" + bct, mesh.stencil_cell_gradient)
            put("cell_gradient_" + bct, mesh.cell_gradient)
        models = {"iso": rng.random(mesh.nC) + .5, "aniso": rng.random((mesh.nC, d)) + .5}
        for k, v in models.items():
            out[f"{tag}_model_{k}"] = v
        for proj, get in (("F", mesh.get_face_inner_product), ("E", mesh.get_edge_inner_product)):
            for mk, im, ix in MASS_CASES:
                out[f"{tag}_mass_{proj}_{mk}_{int(im)}{int(ix)}"] = get(models[mk], invert_model=im, invert_matrix=ix).diagonal()
        v = {"F": rng.standard_normal(mesh.nF), "E": rng.standard_normal(mesh.nE)}
        out[f"{tag}_vF"], out[f"{tag}_vE"] = v["F"], v["E"]
        for proj, get in (("F", mesh.get_face_inner_product_deriv), ("E", mesh.get_edge_inner_product_deriv)):
            for mk, im, ix in MASS_CASES:
                put(f"deriv_{proj}_{mk}_{int(im)}{int(ix)}", get(models[mk], invert_model=im, invert_matrix=ix)(v[proj]))
        ext = np.array([x.sum() for x in h])
        pts = origin + ext * rng.random((40, d))
        pts[0], pts[1] = origin, origin + ext
        pts_out = origin + ext * (rng.random((20, d)) * 1.4 - 0.2)
        out[f"{tag}_pts"], out[f"{tag}_pts_out"] = pts, pts_out
        for loc in LOCS[d]:
            put("interp_" + loc, mesh.get_interpolation_matrix(pts, loc))
        put("interp_zo_CC", mesh.get_interpolation_matrix(pts_out, "CC", zeros_outside=True))
        if d == 2:      # full-tensor model (nC, 3) = [xx, yy, xy]; own generator so the entries above keep their values
            rng_t = np.random.default_rng(97)
            sig = rng_t.random((mesh.nC, 3)) + np.array([2.0, 2.0, 0.0])
            out[f"{tag}_model_tensor"] = sig
            for proj, get, getd in (("F", mesh.get_face_inner_product, mesh.get_face_inner_product_deriv),
                                    ("E", mesh.get_edge_inner_product, mesh.get_edge_inner_product_deriv)):
                for im in (False, True):
                    put(f"mass_{proj}_tensor_{int(im)}0", get(sig, invert_model=im))
                put(f"deriv_{proj}_tensor_00", getd(sig)(v[proj]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tensor_lowdim_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

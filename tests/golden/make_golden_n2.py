"""Generates tests/golden/tensor_n2_small.npz (cell-centred operators) from the UNMODIFIED
reference.  Run in the build container:   python tests/golden/make_golden_n2.py"""
import copy
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refload import load_reference  # noqa: E402

N2_PLAIN = ["average_cell_to_face", "average_cell_vector_to_face", "average_cell_to_edge", "average_edge_to_face",
            "average_node_to_edge", "average_node_to_face", "face_x_divergence", "face_y_divergence",
            "face_z_divergence", "stencil_cell_gradient_x", "stencil_cell_gradient_y", "stencil_cell_gradient_z",
            "cell_gradient_x", "cell_gradient_y", "cell_gradient_z"]
BCS = {"neumann": "neumann", "dirichlet": "dirichlet",
       "mixed": ["dirichlet", ["neumann", "dirichlet"], ["dirichlet", "neumann"]]}


def main():
    ref = load_reference()
    assert ref is not None, "reference not available"
    rng = np.random.default_rng(77)
    h = [rng.random(5) + 0.5, rng.random(4) + 0.5, rng.random(6) + 0.5]   # every axis >= 4 cells (see DESIGN: BSR artefacts)
    origin = np.array([2.0, -1.0, 0.5])
    mesh = ref.TensorMesh(h, origin=origin)
    out = {"hx": h[0], "hy": h[1], "hz": h[2], "origin": origin}

    def put(name, m):
        m = sp.csr_matrix(m).copy()
        m.sum_duplicates()
        m.sort_indices()
        out[name + "__indptr"] = m.indptr.astype(np.int64)
        out[name + "__indices"] = m.indices.astype(np.int64)
        out[name + "__data"] = m.data
        out[name + "__shape"] = np.array(m.shape, dtype=np.int64)

    for name in N2_PLAIN:
        put(name, getattr(mesh, name))
    for tag, bc in BCS.items():
        mesh.set_cell_gradient_BC(copy.deepcopy(bc))
        put("stencil_cell_gradient_" + tag, mesh.stencil_cell_gradient)
        put("cell_gradient_" + tag, mesh.cell_gradient)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tensor_n2_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()

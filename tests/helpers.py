"""Shared helpers for the parity tests."""
import os

import numpy as np
import scipy.sparse as sp

from oracle import tensor_ops as T

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-12  # north_star tolerance for float64 operator values and matvec results


def load_golden(name="tensor_small.npz"):
    return np.load(os.path.join(GOLDEN, name))


def golden_matrix(g, name):
    shape = tuple(int(s) for s in g[name + "__shape"])
    return sp.csr_matrix((g[name + "__data"], g[name + "__indices"], g[name + "__indptr"]), shape=shape)


def assert_same_matrix(a, b, tol=RTOL):
    """Pattern bit-exact after canonicalisation; values within tol relative to max|b|."""
    a, b = T.canonical(a), T.canonical(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    np.testing.assert_array_equal(a.indptr, b.indptr)
    np.testing.assert_array_equal(a.indices, b.indices)
    scale = max(np.abs(b.data).max(), 1e-300) if b.nnz else 1.0
    err = np.abs(a.data - b.data).max(initial=0.0) / scale
    assert err <= tol, f"relative value error {err:.3e} > {tol:.1e}"


def assert_close_vec(a, b, tol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    err = np.abs(a - b).max(initial=0.0) / scale
    assert err <= tol, f"relative error {err:.3e} > {tol:.1e}"


ORACLE_OPS = {
    "face_divergence": T.face_divergence,
    "edge_curl": T.edge_curl,
    "nodal_gradient": T.nodal_gradient,
    "average_face_to_cell": T.average_face_to_cell,
    "average_face_to_cell_vector": T.average_face_to_cell_vector,
    "average_edge_to_cell": T.average_edge_to_cell,
    "average_edge_to_cell_vector": T.average_edge_to_cell_vector,
    "average_node_to_cell": T.average_node_to_cell,
    "average_face_x_to_cell": lambda g: T.average_component_to_cell(g, "Fx"),
    "average_face_y_to_cell": lambda g: T.average_component_to_cell(g, "Fy"),
    "average_face_z_to_cell": lambda g: T.average_component_to_cell(g, "Fz"),
    "average_edge_x_to_cell": lambda g: T.average_component_to_cell(g, "Ex"),
    "average_edge_y_to_cell": lambda g: T.average_component_to_cell(g, "Ey"),
    "average_edge_z_to_cell": lambda g: T.average_component_to_cell(g, "Ez"),
}

ORACLE_N2_OPS = {
    "average_cell_to_face": T.average_cell_to_face,
    "average_cell_vector_to_face": T.average_cell_vector_to_face,
    "average_cell_to_edge": T.average_cell_to_edge,
    "average_edge_to_face": T.average_edge_to_face,
    "average_node_to_edge": T.average_node_to_edge,
    "average_node_to_face": T.average_node_to_face,
    "face_x_divergence": lambda g: T.face_component_divergence(g, 0),
    "face_y_divergence": lambda g: T.face_component_divergence(g, 1),
    "face_z_divergence": lambda g: T.face_component_divergence(g, 2),
    "stencil_cell_gradient_x": lambda g: T.stencil_cell_gradient(g, axis=0),
    "stencil_cell_gradient_y": lambda g: T.stencil_cell_gradient(g, axis=1),
    "stencil_cell_gradient_z": lambda g: T.stencil_cell_gradient(g, axis=2),
    "cell_gradient_x": lambda g: T.cell_gradient(g, axis=0),
    "cell_gradient_y": lambda g: T.cell_gradient(g, axis=1),
    "cell_gradient_z": lambda g: T.cell_gradient(g, axis=2),
}
CELL_GRADIENT_BCS = {"neumann": "neumann", "dirichlet": "dirichlet",
                     "mixed": ["dirichlet", ["neumann", "dirichlet"], ["dirichlet", "neumann"]]}

MESHES = {
    "uniform": lambda rng: ([5, 4, 3], None),
    "nonuniform": lambda rng: ([rng.random(4) + 0.5, rng.random(6) + 0.5, rng.random(5) + 0.5], np.array([-1.0, 0.25, 3.0])),
    "thin": lambda rng: ([rng.random(7) + 0.5, rng.random(1) + 0.5, rng.random(2) + 0.5], None),
    "wide": lambda rng: ([rng.random(70) + 0.5, rng.random(9) + 0.5, rng.random(37) + 0.5], None),
    # strongly graded (padding cells growing by 1.4x): defeats the interpolation kernel's index guess
    "graded": lambda rng: ([np.r_[1.4 ** np.arange(12, 0, -1), np.ones(9), 1.4 ** np.arange(1, 13)],
                            np.r_[np.ones(10), 1.3 ** np.arange(1, 15)], np.r_[2.0 ** np.arange(8, 0, -1), np.ones(8)]],
                           np.array([-300.0, 0.0, -500.0])),
}


def make_pair(name, seed=4):
    """(discretize_b200.TensorMesh, oracle Grid) on the same widths."""
    import discretize_b200 as dz

    rng = np.random.default_rng(seed)
    h, origin = MESHES[name](rng)
    return dz.TensorMesh(h, origin=origin), T.Grid(h, origin)


def adversarial_points(g, rng, n=300):
    lo = np.array([x[0] for x in g.nodes]); hi = np.array([x[-1] for x in g.nodes])
    pts = lo + rng.random((n, 3)) * (hi - lo)
    k = 0
    for a in range(3):
        for val in (g.nodes[a][0], g.nodes[a][-1], g.nodes[a][len(g.nodes[a]) // 2],
                    g.centers[a][0], g.centers[a][-1],
                    0.5 * (g.nodes[a][0] + g.centers[a][0]), 0.5 * (g.nodes[a][-1] + g.centers[a][-1])):
            pts[k, a] = val
            k += 1
    pts[k] = lo; pts[k + 1] = hi
    return pts

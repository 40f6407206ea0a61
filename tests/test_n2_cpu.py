"""CPU: the cell-centred operators of the oracle against the reference-made fixture and, where
/root/reference exists, against the reference itself."""
import copy

import numpy as np
import pytest

from oracle import tensor_ops as T
from oracle.refload import load_reference
from tests.helpers import CELL_GRADIENT_BCS, ORACLE_N2_OPS, assert_same_matrix, golden_matrix, load_golden


@pytest.fixture(scope="module")
def gold():
    return load_golden("tensor_n2_small.npz")


def _grid(g):
    return T.Grid([g["hx"], g["hy"], g["hz"]], g["origin"])


@pytest.mark.parametrize("name", sorted(ORACLE_N2_OPS))
def test_plain_ops_match_fixture(gold, name):
    assert_same_matrix(ORACLE_N2_OPS[name](_grid(gold)), golden_matrix(gold, name), 1e-14)


@pytest.mark.parametrize("tag", sorted(CELL_GRADIENT_BCS))
def test_cell_gradient_bcs_match_fixture(gold, tag):
    g = _grid(gold)
    bc = CELL_GRADIENT_BCS[tag]
    assert_same_matrix(T.stencil_cell_gradient(g, bc), golden_matrix(gold, "stencil_cell_gradient_" + tag), 1e-14)
    assert_same_matrix(T.cell_gradient(g, bc), golden_matrix(gold, "cell_gradient_" + tag), 1e-14)


def test_against_reference_on_other_meshes():
    ref = load_reference()
    if ref is None:
        pytest.skip("reference not available on this machine")
    rng = np.random.default_rng(3)
    for h in ([5, 4, 6], [rng.random(7) + 0.5, rng.random(4) + 0.5, rng.random(5) + 0.5], [4, 9, 4]):
        m, g = ref.TensorMesh(h), T.Grid(h)
        for name, fn in ORACLE_N2_OPS.items():
            assert_same_matrix(fn(g), getattr(m, name), 1e-14)
        for bc in CELL_GRADIENT_BCS.values():
            m.set_cell_gradient_BC(copy.deepcopy(bc))
            assert_same_matrix(T.stencil_cell_gradient(g, bc), m.stencil_cell_gradient, 1e-14)
            assert_same_matrix(T.cell_gradient(g, bc), m.cell_gradient, 1e-14)


def test_known_properties():
    """constants are averaged to constants; cell_gradient of a constant vanishes under Neumann; the Dirichlet
    ghost-point rows are +-2/h (reference docstrings, differential_operators.py:41-82)."""
    g = T.Grid([np.array([1.0, 2.0, 0.5, 1.5]), 4, 5])
    for fn in (T.average_cell_to_face, T.average_cell_to_edge, T.average_node_to_edge, T.average_node_to_face,
               T.average_edge_to_face):
        A = fn(g)
        np.testing.assert_allclose(A @ np.ones(A.shape[1]), 1.0, rtol=1e-14)
    G = T.cell_gradient(g, "neumann")
    np.testing.assert_allclose(G @ np.ones(g.nC), 0.0, atol=1e-13)
    Gd = T.cell_gradient(g, "dirichlet")
    first_x_face = Gd[0].toarray().ravel()
    assert first_x_face[0] == pytest.approx(2.0 / 1.0)


def test_boundary_projections_and_integrals():
    """project_*_to_boundary_*, boundary_faces / nodes / edges / outward normals, boundary_face_scalar_integral
    (base_tensor_mesh.py:300-310, 423-560; differential_operators.py:2183-2197, 3384-3430): against the reference where
    it is available, and by their defining properties everywhere."""
    import discretize_b200 as dz

    ref = load_reference()
    rng = np.random.default_rng(2)
    for h, o in (([rng.random(4) + .5, rng.random(3) + .5, rng.random(5) + .5], np.array([1., 2., 3.])),
                 ([rng.random(5) + .5, rng.random(4) + .5], None), ([6], None), ([1, 1, 1], None)):
        m = dz.TensorMesh(h, origin=o)
        n = m.shape_cells
        Pf = m.project_face_to_boundary_face.tocsr()
        n_bf = 2 * sum(int(np.prod([n[b] for b in range(m.dim) if b != a])) for a in range(m.dim))
        assert Pf.shape == (n_bf, m.n_faces) and np.array_equal((Pf @ Pf.T).toarray(), np.eye(n_bf))
        Pn = m.project_node_to_boundary_node.tocsr()
        interior = int(np.prod([max(k - 1, 0) for k in n]))
        assert Pn.shape[0] == m.n_nodes - interior
        normals = m.boundary_face_outward_normals
        assert np.asarray(normals).shape[0] == n_bf
        A = m.boundary_face_scalar_integral.tocsr()
        if m.dim > 1:   # each boundary face carries +-(its area): the sign of its outward normal
            flux = Pf @ (A @ np.ones(n_bf))
            np.testing.assert_allclose(flux, (Pf @ m.face_areas) * np.asarray(normals).sum(axis=1))
        if ref is None:
            continue
        mr = ref.TensorMesh(h, origin=o)
        for name in ("project_face_to_boundary_face", "project_edge_to_boundary_edge", "project_node_to_boundary_node",
                     "boundary_face_scalar_integral"):
            a, b = getattr(m, name), getattr(mr, name)
            if b is None:
                assert a is None
            else:
                assert_same_matrix(a.tocsr(), b, 1e-14)
        for name in ("boundary_nodes", "boundary_faces", "boundary_edges", "boundary_face_outward_normals"):
            a, b = getattr(m, name), getattr(mr, name)
            if b is None:
                assert a is None
            else:
                np.testing.assert_array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))

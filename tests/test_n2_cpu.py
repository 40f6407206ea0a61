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

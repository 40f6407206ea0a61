"""GPU parity tests for the cell-centred operators: CUDA path through the python mirror -> C ABI
against the numpy oracle on several meshes and against the reference-made fixture.  Patterns bit-exact (explicit
zeros of the Neumann stencil rows included), values and matvecs <= 1e-12 relative."""
import copy

import numpy as np
import pytest

from oracle import tensor_ops as T
from tests.helpers import (CELL_GRADIENT_BCS, MESHES, ORACLE_N2_OPS, assert_close_vec, assert_same_matrix, golden_matrix,
                           load_golden, make_pair)

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device (no CPU fallback exists)")


def _check(op, ref, seed=1):
    assert op.shape == ref.shape
    assert_same_matrix(op.tocsr(), ref)
    assert_same_matrix(op.T.tocsr(), ref.T.tocsr())
    rng = np.random.default_rng(seed)
    x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
    assert_close_vec(op @ x, ref @ x)
    assert_close_vec(op.T @ y, ref.T @ y)


@pytest.mark.parametrize("mesh_name", list(MESHES))
@pytest.mark.parametrize("name", sorted(ORACLE_N2_OPS))
def test_n2_matrix_and_apply(mesh_name, name):
    mesh, grid = make_pair(mesh_name)
    _check(getattr(mesh, name), ORACLE_N2_OPS[name](grid))


@pytest.mark.parametrize("mesh_name", ["nonuniform", "thin", "wide"])
@pytest.mark.parametrize("tag", sorted(CELL_GRADIENT_BCS))
def test_cell_gradient_boundary_conditions(mesh_name, tag):
    mesh, grid = make_pair(mesh_name)
    bc = CELL_GRADIENT_BCS[tag]
    mesh.set_cell_gradient_BC(copy.deepcopy(bc))
    _check(mesh.stencil_cell_gradient, T.stencil_cell_gradient(grid, bc))
    _check(mesh.cell_gradient, T.cell_gradient(grid, bc))


@pytest.mark.parametrize("march", [0, -1])
def test_cell_to_face_family_marching_and_generic(march):
    """average_cell_to_face and (stencil_)cell_gradient have a plane-marching kernel and the generic row kernel."""
    import ctypes

    from discretize_b200._lib import LIB_PATH

    lib = ctypes.CDLL(LIB_PATH)
    lib.dzb_tune_march(march)
    try:
        for mesh_name in ("wide", "thin"):
            mesh, grid = make_pair(mesh_name)
            rng = np.random.default_rng(5)
            x = rng.standard_normal(mesh.nC)
            assert_close_vec(mesh.average_cell_to_face @ x, T.average_cell_to_face(grid) @ x)
            for bc in CELL_GRADIENT_BCS.values():
                mesh.set_cell_gradient_BC(copy.deepcopy(bc))
                assert_close_vec(mesh.cell_gradient @ x, T.cell_gradient(grid, bc) @ x)
                assert_close_vec(mesh.stencil_cell_gradient @ x, T.stencil_cell_gradient(grid, bc) @ x)
    finally:
        lib.dzb_tune_march(-1)


def test_n2_vs_reference_fixture():
    import discretize_b200 as dz

    g = load_golden("tensor_n2_small.npz")
    mesh = dz.TensorMesh([g["hx"], g["hy"], g["hz"]], origin=g["origin"])
    for name in ORACLE_N2_OPS:
        assert_same_matrix(getattr(mesh, name).tocsr(), golden_matrix(g, name))
    for tag, bc in CELL_GRADIENT_BCS.items():
        mesh.set_cell_gradient_BC(copy.deepcopy(bc))
        assert_same_matrix(mesh.stencil_cell_gradient.tocsr(), golden_matrix(g, "stencil_cell_gradient_" + tag))
        assert_same_matrix(mesh.cell_gradient.tocsr(), golden_matrix(g, "cell_gradient_" + tag))


def test_cell_centred_dc_operator_and_errors():
    """the cell-centred DC system -D Mf(1/sigma)^-1 D^T (reference examples/plot_dc_resistivity.py:23-30) composes
    from the new pieces; set_cell_gradient_BC keeps the reference's argument checks."""
    import discretize_b200 as dz

    mesh, grid = make_pair("nonuniform")
    sig = np.random.default_rng(2).random(mesh.nC) + 0.5
    D = mesh.face_divergence
    A = D @ mesh.get_face_inner_product(sig, invert_model=True, invert_matrix=True) @ D.T
    Dr = T.face_divergence(grid)
    Ar = Dr @ T.inner_product(grid, "F", sig, invert_model=True, invert_matrix=True) @ Dr.T
    x = np.random.default_rng(3).standard_normal(mesh.nC)
    assert_close_vec(A @ x, Ar @ x)
    with pytest.raises(ValueError):
        mesh.set_cell_gradient_BC(["neumann", "neumann"])
    with pytest.raises(ValueError):
        mesh.set_cell_gradient_BC("robin")
    with pytest.raises(TypeError):
        mesh.set_cell_gradient_BC(3)
    with pytest.raises(NotImplementedError):
        dz.TensorMesh([4, 4, 4]).nodal_laplacian

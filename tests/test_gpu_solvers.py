"""GPU tests of the solver-side pieces: analytic Jacobi diagonals of the fused operators, the
entrywise-squared stencils behind them, the device-resident CG vector kernels and the PCG driver."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse.linalg as sla

from oracle import tensor_ops as T
from tests.helpers import assert_close_vec, assert_same_matrix, make_pair

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device (no CPU fallback exists)")


def _fused(mesh, grid, seed=3):
    rng = np.random.default_rng(seed)
    sig, mui = np.exp(rng.standard_normal(mesh.nC)), np.exp(rng.standard_normal(mesh.nC))
    G, Cc = mesh.nodal_gradient, mesh.edge_curl
    Adc = G.T @ mesh.get_edge_inner_product(sig) @ G
    Amx = Cc.T @ mesh.get_face_inner_product(mui) @ Cc + mesh.get_edge_inner_product(sig)
    Gr, Cr = T.nodal_gradient(grid), T.edge_curl(grid)
    Rdc = (Gr.T @ T.inner_product(grid, "E", sig) @ Gr).tocsr()
    Rmx = (Cr.T @ T.inner_product(grid, "F", mui) @ Cr + T.inner_product(grid, "E", sig)).tocsr()
    return Adc, Amx, Rdc, Rmx


@pytest.mark.parametrize("mesh_name", ["nonuniform", "wide", "thin"])
def test_squared_stencils_and_fused_diagonals(mesh_name):
    from discretize_b200 import _lib
    from discretize_b200.tensor_ops import StencilOperator

    mesh, grid = make_pair(mesh_name)
    for op_id, ref in ((_lib.OP_GRAD_SQ, T.nodal_gradient(grid)), (_lib.OP_CURL_SQ, T.edge_curl(grid))):
        sq = ref.multiply(ref).tocsr()
        op = StencilOperator(mesh, op_id, name="sq")
        assert_same_matrix(op.tocsr(), sq)
        assert_same_matrix(op.T.tocsr(), sq.T.tocsr())
        x = np.random.default_rng(1).standard_normal(sq.shape[0])
        assert_close_vec(op.T @ x, sq.T @ x)
    Adc, Amx, Rdc, Rmx = _fused(mesh, grid)
    assert type(Adc).__name__ == "DCNodalOperator" and type(Amx).__name__ == "MaxwellOperator"
    assert_close_vec(Adc.diagonal(), Rdc.diagonal())
    assert_close_vec(Amx.diagonal(), Rmx.diagonal())


def test_dot_kernel_is_deterministic_and_exact_enough():
    from discretize_b200._lib import load
    from discretize_b200.operators import ptr, stream_ptr

    lib = load()
    n = 3_000_017
    gen = torch.Generator(device="cuda"); gen.manual_seed(4)
    a = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    ws = torch.zeros(int(lib.dzb_solver_workspace_bytes()), dtype=torch.uint8, device="cuda")
    scal = torch.zeros(6, dtype=torch.float64, device="cuda")
    vals = []
    for _ in range(3):
        assert lib.dzb_dot(n, ptr(a), ptr(b), ptr(ws), ptr(scal), 1, stream_ptr()) == 0
        vals.append(float(scal[1]))
    assert vals[0] == vals[1] == vals[2]                       # block-ordered reduction: bitwise repeatable
    ref = float(torch.dot(a, b))
    assert abs(vals[0] - ref) <= 1e-12 * float(a.norm() * b.norm())


@pytest.mark.parametrize("jacobi", [False, True])
def test_pcg_on_fused_maxwell_and_dc(jacobi):
    from discretize_b200.solvers import pcg

    mesh, grid = make_pair("wide")          # 70 x 9 x 37, strongly varying sigma / mu
    Adc, Amx, Rdc, Rmx = _fused(mesh, grid)
    rng = np.random.default_rng(9)
    # Maxwell (SPD): against a direct solve of the oracle's matrix
    b = rng.standard_normal(mesh.nE)
    dinv = 1.0 / Amx.diagonal_device() if jacobi else None
    x, info = pcg(Amx, b, inv_diag=dinv, rtol=1e-11, maxiter=5000, check_every=25)
    assert info.converged, info
    xr = sla.spsolve(Rmx.tocsc(), b)
    assert_close_vec(x.cpu().numpy(), xr, 1e-8)
    assert np.linalg.norm(Rmx @ x.cpu().numpy() - b) <= 1e-10 * np.linalg.norm(b)
    # DC (singular: constants in the null space) with a right-hand side in the range
    xt = rng.standard_normal(mesh.nN)
    bdc = Rdc @ xt
    dinv = 1.0 / Adc.diagonal_device() if jacobi else None
    x, info = pcg(Adc, bdc, inv_diag=dinv, rtol=1e-10, maxiter=8000, check_every=25)
    assert info.converged, info
    assert np.linalg.norm(Rdc @ x.cpu().numpy() - bdc) <= 1e-9 * np.linalg.norm(bdc)
    if jacobi:
        _, plain = pcg(Adc, bdc, rtol=1e-10, maxiter=8000, check_every=25)
        assert info.iterations <= plain.iterations          # the preconditioner must not hurt on a graded mesh


def test_pcg_cell_centred_dc_on_octree():
    """The cell-centred DC system of SimPEG on an OcTree, V D Mf(rho)^-1 D^T V (+ eps V to pin the constant), composed
    lazily from the CUDA OcTree operators and solved by the device-resident PCG; checked against spsolve of the host
    matrices (which are the reference's, tests/test_tree_cpu.py)."""
    import scipy.sparse as sp

    import discretize_b200 as dz
    from discretize_b200.operators import DiagonalOperator
    from discretize_b200.solvers import pcg

    m = dz.TreeMesh([16, 16, 16], diagonal_balance=True)
    rng = np.random.default_rng(21)
    m.refine_points(rng.random((5, 3)) * 0.6 + 0.2, -1, [1, 1], finalize=True)
    rho = np.exp(rng.standard_normal(m.nC))
    V = m.cell_volumes
    D = m.face_divergence
    MfI = m.get_face_inner_product(rho, invert_matrix=True)
    Vop = DiagonalOperator(V)
    A = Vop @ D @ MfI @ D.T @ Vop + DiagonalOperator(1e-3 * V)
    Dh = m._host_matrix("face_divergence")
    Ah = (sp.diags(V) @ Dh @ sp.diags(1.0 / MfI.diag.cpu().numpy()) @ Dh.T @ sp.diags(V) + sp.diags(1e-3 * V)).tocsc()
    b = rng.standard_normal(m.nC)
    x, info = pcg(A, b, inv_diag=1.0 / Ah.diagonal(), rtol=1e-11, maxiter=4000, check_every=20)
    assert info.converged, info
    assert_close_vec(x.cpu().numpy(), sla.spsolve(Ah, b), 1e-8)


def test_pcg_survives_convergence_inside_a_blind_window():
    """A diagonal system with the exact Jacobi preconditioner converges in ONE iteration; the remaining iterations of the
    check_every window then see r.z = p.Ap = 0 and must be no-ops instead of 0/0 (NaN solution)."""
    from discretize_b200.solvers import pcg

    mesh, grid = make_pair("wide")
    rng = np.random.default_rng(8)
    sig = np.exp(rng.standard_normal(mesh.nC))
    Me = mesh.get_edge_inner_product(sig)
    d = T.inner_product(grid, "E", sig).diagonal()
    b = rng.standard_normal(mesh.nE)
    for x0 in (None, b / d):                      # cold start and an exact initial guess
        x, info = pcg(Me, b, x0=x0, inv_diag=1.0 / d, rtol=1e-12, maxiter=40, check_every=10)
        xs = x.cpu().numpy()
        assert np.isfinite(xs).all() and info.converged
        assert_close_vec(xs, b / d)


@pytest.mark.parametrize("shape", [(33, 20, 17), (70, 41, 64), (600, 9, 130)])
def test_dc_apply_with_fused_dot(shape):
    """K8 with the p.Ap of a CG iteration accumulated by its consumers: same y as the plain kernel (bitwise) and the same
    dot product as a separate reduction (to rounding), deterministic from run to run."""
    import discretize_b200 as dz

    rng = np.random.default_rng(8)
    mesh = dz.TensorMesh([rng.random(n) + 0.5 for n in shape])
    sigma = np.exp(rng.standard_normal(mesh.nC))
    G = mesh.nodal_gradient
    A = G.T @ mesh.get_edge_inner_product(sigma) @ G
    x = torch.as_tensor(rng.standard_normal(mesh.nN), device="cuda")
    y0 = A.apply_device(x)
    y1 = torch.full_like(y0, float("nan"))
    dots = torch.zeros(3, dtype=torch.float64, device="cuda")
    assert A.apply_with_dot(x, y1, dots[1:2])
    assert torch.equal(y0, y1)
    want = float(torch.dot(x, y0))
    assert abs(float(dots[1]) - want) <= 1e-12 * float(torch.linalg.vector_norm(x) * torch.linalg.vector_norm(y0))
    assert float(dots[0]) == 0.0 and float(dots[2]) == 0.0
    again = torch.zeros(1, dtype=torch.float64, device="cuda")
    assert A.apply_with_dot(x, y1, again)
    assert float(again[0]) == float(dots[1])

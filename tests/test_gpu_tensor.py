"""GPU parity tests: the CUDA path (through the python mirror of the reference API, which
calls the C ABI) against the numpy oracle and the committed reference fixtures.

Tolerances (north_star): sparsity patterns / interpolation indices bit-exact; operator
values and matvec results <= 1e-12 relative (float64).
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from oracle import tensor_ops as T
from tests.helpers import (MESHES, ORACLE_OPS, RTOL, adversarial_points, assert_close_vec, assert_same_matrix,
                           golden_matrix, load_golden, make_pair)

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device (no CPU fallback exists)")


# ------------------------------------------------------------------------------------------
# elementary operators
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mesh_name", list(MESHES))
@pytest.mark.parametrize("name", list(ORACLE_OPS))
def test_elementary_matrix_and_apply(mesh_name, name):
    mesh, grid = make_pair(mesh_name)
    ref = ORACLE_OPS[name](grid)
    op = getattr(mesh, name)
    assert op.shape == ref.shape
    assert_same_matrix(op.tocsr(), ref)
    assert_same_matrix(op.T.tocsr(), ref.T.tocsr())
    rng = np.random.default_rng(1)
    x = rng.standard_normal(ref.shape[1])
    y = rng.standard_normal(ref.shape[0])
    assert_close_vec(op @ x, ref @ x)
    assert_close_vec(op.T @ y, ref.T @ y)


@pytest.mark.parametrize("march", [0, -1])
@pytest.mark.parametrize("shape", [(70, 9, 70), (33, 40, 5), (1, 1, 1), (2, 67, 33)])
def test_core_operators_marching_and_generic_kernels(shape, march):
    """D, G, C and their transposes have two CUDA implementations (plane-marching kernels, generic row kernels);
    both must match the oracle, across several z-chunks, x-blocks and degenerate extents."""
    import ctypes

    import discretize_b200 as dz

    rng = np.random.default_rng(12)
    h = [rng.random(n) + 0.5 for n in shape]
    mesh, grid = dz.TensorMesh(h), T.Grid(h)
    from discretize_b200._lib import LIB_PATH

    lib = ctypes.CDLL(LIB_PATH)
    lib.dzb_tune_march(march)
    try:
        for name in ("face_divergence", "nodal_gradient", "edge_curl", "average_face_to_cell", "average_edge_to_cell",
                     "average_node_to_cell"):
            ref = ORACLE_OPS[name](grid)
            op = getattr(mesh, name)
            x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
            assert_close_vec(op @ x, ref @ x)
            assert_close_vec(op.T @ y, ref.T @ y)
    finally:
        lib.dzb_tune_march(-1)


@pytest.mark.parametrize("name", list(ORACLE_OPS))
def test_elementary_vs_reference_fixture(name):
    g = load_golden()
    import discretize_b200 as dz

    mesh = dz.TensorMesh([g["hx"], g["hy"], g["hz"]], origin=g["origin"])
    assert_same_matrix(getattr(mesh, name).tocsr(), golden_matrix(g, name))


def test_config1_64cubed():
    """BASELINE config 1: TensorMesh 64^3, D @ u_f, C @ e, Mf(sigma) @ u_f -- full element-wise parity."""
    import discretize_b200 as dz

    mesh, grid = dz.TensorMesh([64, 64, 64]), T.Grid([64, 64, 64])
    u = np.random.default_rng(1).standard_normal(mesh.nF)
    e = np.random.default_rng(2).standard_normal(mesh.nE)
    sigma = np.random.default_rng(3).random(mesh.nC) + 0.5
    D, C = T.face_divergence(grid), T.edge_curl(grid)
    Mf = T.inner_product(grid, "F", sigma)
    assert_same_matrix(mesh.face_divergence.tocsr(), D)
    assert_same_matrix(mesh.edge_curl.tocsr(), C)
    assert_same_matrix(mesh.get_face_inner_product(sigma).tocsr(), Mf)
    assert_close_vec(mesh.face_divergence @ u, D @ u)
    assert_close_vec(mesh.edge_curl @ e, C @ e)
    assert_close_vec(mesh.get_face_inner_product(sigma) @ u, Mf @ u)
    # non-uniform variant
    rng = np.random.default_rng(4)
    h = [rng.random(64) + 0.5 for _ in range(3)]
    mesh, grid = dz.TensorMesh(h), T.Grid(h)
    assert_close_vec(mesh.face_divergence @ u, T.face_divergence(grid) @ u)
    assert_close_vec(mesh.edge_curl @ e, T.edge_curl(grid) @ e)
    assert_close_vec(mesh.get_face_inner_product(sigma) @ u, T.inner_product(grid, "F", sigma) @ u)


def test_mimetic_identities_device():
    mesh, _ = make_pair("nonuniform")
    rng = np.random.default_rng(0)
    e = rng.standard_normal(mesh.nE)
    phi = rng.standard_normal(mesh.nN)
    dc = mesh.face_divergence @ (mesh.edge_curl @ e)
    cg = mesh.edge_curl @ (mesh.nodal_gradient @ phi)
    assert np.abs(dc).max() < 1e-12 * np.abs(e).max() * 100
    assert np.abs(cg).max() < 1e-12 * np.abs(phi).max() * 100


# ------------------------------------------------------------------------------------------
# mass matrices
# ------------------------------------------------------------------------------------------
def _models(grid, rng):
    return {
        "none": None,
        "scalar": 2.5,
        "iso": rng.random(grid.nC) + 0.5,
        "aniso": rng.random((grid.nC, 3)) + 0.5,
        "tensor": np.c_[rng.random((grid.nC, 3)) + 2.0, rng.random((grid.nC, 3)) * 0.5 - 0.25],
    }


@pytest.mark.parametrize("mesh_name", ["uniform", "nonuniform", "thin"])
@pytest.mark.parametrize("proj", ["F", "E"])
@pytest.mark.parametrize("im", [False, True])
@pytest.mark.parametrize("ix", [False, True])
def test_mass(mesh_name, proj, im, ix):
    mesh, grid = make_pair(mesh_name)
    rng = np.random.default_rng(3)
    get = mesh.get_face_inner_product if proj == "F" else mesh.get_edge_inner_product
    n = grid.nF if proj == "F" else grid.nE
    x = rng.standard_normal(n)
    for mk, m in _models(grid, rng).items():
        if mk == "tensor" and ix:
            with pytest.raises(Exception, match="Solver needed"):
                get(m, invert_model=im, invert_matrix=ix)
            continue
        ref = T.inner_product(grid, proj, m, invert_model=im, invert_matrix=ix)
        op = get(m, invert_model=im, invert_matrix=ix)
        assert_same_matrix(op.tocsr(), ref)
        assert_close_vec(op @ x, ref @ x)
        assert_close_vec(op.T @ x, ref @ x)


@pytest.mark.parametrize("proj", ["F", "E"])
@pytest.mark.parametrize("mk", ["iso", "aniso", "tensor"])
def test_mass_vs_reference_fixture(proj, mk):
    g = load_golden()
    import discretize_b200 as dz

    mesh = dz.TensorMesh([g["hx"], g["hy"], g["hz"]], origin=g["origin"])
    get = mesh.get_face_inner_product if proj == "F" else mesh.get_edge_inner_product
    for im in (False, True):
        for ix in (False, True):
            if mk == "tensor" and ix:
                continue
            assert_same_matrix(get(g["model_" + mk], invert_model=im, invert_matrix=ix).tocsr(),
                               golden_matrix(g, f"mass_{proj}_{mk}_{int(im)}{int(ix)}"))


def test_mass_errors():
    mesh, grid = make_pair("uniform")
    with pytest.raises(TypeError, match="invProp"):
        mesh.get_face_inner_product(invProp=True)
    with pytest.raises(TypeError, match="doFast"):
        mesh.get_edge_inner_product(doFast=True)
    with pytest.raises(Exception, match="Unexpected shape of tensor"):
        mesh.get_face_inner_product(np.ones(grid.nC + 1))


@pytest.mark.parametrize("mesh_name", ["uniform", "nonuniform"])
@pytest.mark.parametrize("proj", ["F", "E"])
@pytest.mark.parametrize("im", [False, True])
@pytest.mark.parametrize("ix", [False, True])
def test_mass_deriv(mesh_name, proj, im, ix):
    mesh, grid = make_pair(mesh_name)
    rng = np.random.default_rng(5)
    n = grid.nF if proj == "F" else grid.nE
    v = rng.standard_normal(n)
    w = rng.standard_normal(n)
    get = mesh.get_face_inner_product_deriv if proj == "F" else mesh.get_edge_inner_product_deriv
    for mk, m in _models(grid, rng).items():
        if mk == "none":
            continue
        if mk == "tensor" and (im or ix):
            with pytest.raises(NotImplementedError):
                get(m, invert_model=im, invert_matrix=ix)
            continue
        if mk == "scalar" and im and not ix:
            continue  # undefined in the reference (raises there)
        ref = T.inner_product_deriv(grid, proj, m, v, invert_model=im, invert_matrix=ix)
        op = get(m, invert_model=im, invert_matrix=ix)(v)
        assert op.shape == ref.shape, (mk, op.shape, ref.shape)
        dm = rng.standard_normal(ref.shape[1])
        assert_close_vec(op @ dm, ref @ dm)
        assert_close_vec(op.T @ w, ref.T @ w)
        assert_same_matrix(op.tocsr(), ref)


@pytest.mark.parametrize("proj", ["F", "E"])
def test_mass_deriv_generic_path(proj):
    mesh, grid = make_pair("nonuniform")
    rng = np.random.default_rng(6)
    n = grid.nF if proj == "F" else grid.nE
    v, w = rng.standard_normal(n), rng.standard_normal(n)
    get = mesh.get_face_inner_product_deriv if proj == "F" else mesh.get_edge_inner_product_deriv
    for mk, m in _models(grid, rng).items():
        if mk == "none":
            continue
        ref = T.inner_product_deriv(grid, proj, m, v, do_fast=False)
        op = get(m, do_fast=False)(v)
        dm = rng.standard_normal(ref.shape[1])
        assert_close_vec(op @ dm, ref @ dm)
        assert_close_vec(op.T @ w, ref.T @ w)


# ------------------------------------------------------------------------------------------
# fused composites
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mesh_name", list(MESHES))
def test_dc_nodal_fused(mesh_name):
    from discretize_b200.tensor_ops import DCNodalOperator

    mesh, grid = make_pair(mesh_name)
    rng = np.random.default_rng(6)
    sigma = np.exp(rng.standard_normal(grid.nC))
    phi = rng.standard_normal(grid.nN)
    G = T.nodal_gradient(grid)
    Me = T.inner_product(grid, "E", sigma)
    ref = (G.T @ Me @ G).tocsr()
    Gd = mesh.nodal_gradient
    A = Gd.T @ mesh.get_edge_inner_product(sigma) @ Gd
    assert isinstance(A, DCNodalOperator)  # the algebra fused it
    assert_close_vec(A @ phi, ref @ phi)
    assert_close_vec(A.T @ phi, ref @ phi)
    # unfused chain of elementary kernels gives the same thing
    chain = Gd.T @ (mesh.get_edge_inner_product(sigma) @ (Gd @ phi))
    assert_close_vec(A @ phi, chain)
    assert_same_matrix(A.tocsr(), ref)


def test_dc_nodal_fixture_and_tuning():
    import ctypes

    import discretize_b200 as dz
    from discretize_b200._lib import load

    g = load_golden()
    mesh = dz.TensorMesh([g["hx"], g["hy"], g["hz"]], origin=g["origin"])
    ref = golden_matrix(g, "dc_nodal")
    phi = np.random.default_rng(7).standard_normal(mesh.nN)
    Gd = mesh.nodal_gradient
    A = Gd.T @ mesh.get_edge_inner_product(g["model_iso"]) @ Gd
    lib = load()
    lib.dzb_tune_dc.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.dzb_tune_dc.restype = None
    try:
        for ry in (2, 4, 8):
            for kchunk in (1, 3, 32):
                lib.dzb_tune_dc(ry, kchunk)
                assert_close_vec(A @ phi, ref @ phi)
    finally:
        lib.dzb_tune_dc(4, 32)


def test_dc_nodal_streamed_host_path():
    """Pinned host tensor in -> pinned host tensor out: the z-chunk streamed apply equals the device apply."""
    import discretize_b200 as dz

    rng = np.random.default_rng(14)
    mesh = dz.TensorMesh([rng.random(40) + 0.5, rng.random(30) + 0.5, rng.random(21) + 0.5])
    sigma = np.exp(rng.standard_normal(mesh.nC))
    Gd = mesh.nodal_gradient
    A = Gd.T @ mesh.get_edge_inner_product(sigma) @ Gd
    phi = rng.standard_normal(mesh.nN)
    want = A @ phi
    xh = torch.from_numpy(phi).pin_memory()
    yh = torch.empty(mesh.nN, dtype=torch.float64).pin_memory()
    got = A.matvec(xh, out=yh)
    assert got is yh and yh.device.type == "cpu"
    np.testing.assert_array_equal(yh.numpy(), want)
    np.testing.assert_array_equal(A.matvec(xh).numpy(), want)


@pytest.mark.parametrize("shape", [(600, 5, 7), (601, 4, 9), (1100, 3, 5), (271, 6, 130)])
def test_fused_kernels_multiple_x_blocks_and_long_z(shape):
    """Meshes wider than one block of the ring kernels (K8: 540 node columns, K9: 270) with even and odd nx,
    and a z-extent longer than one z-chunk: fused == unfused chain of elementary kernels."""
    import discretize_b200 as dz

    rng = np.random.default_rng(15)
    mesh = dz.TensorMesh([rng.random(n) + 0.5 for n in shape])
    sigma = np.exp(rng.standard_normal(mesh.nC))
    mui = np.exp(rng.standard_normal(mesh.nC))
    phi, e = rng.standard_normal(mesh.nN), rng.standard_normal(mesh.nE)
    Gd, Cd = mesh.nodal_gradient, mesh.edge_curl
    Me, Mf = mesh.get_edge_inner_product(sigma), mesh.get_face_inner_product(mui)
    A = Gd.T @ Me @ Gd
    B = Cd.T @ Mf @ Cd + Me
    assert_close_vec(A @ phi, Gd.T @ (Me @ (Gd @ phi)))
    assert_close_vec(B @ e, Cd.T @ (Mf @ (Cd @ e)) + Me @ e)


def test_dc_nodal_128_against_chain_and_cg():
    """Size-independent checks at a larger size: fused == unfused chain; SPD => CG converges."""
    import discretize_b200 as dz

    rng = np.random.default_rng(8)
    h = [rng.random(128) + 0.5, rng.random(96) + 0.5, rng.random(100) + 0.5]
    mesh = dz.TensorMesh(h)
    sigma = np.exp(rng.standard_normal(mesh.nC))
    phi = rng.standard_normal(mesh.nN)
    Gd = mesh.nodal_gradient
    Me = mesh.get_edge_inner_product(sigma)
    A = Gd.T @ Me @ Gd
    fused = A @ phi
    chain = Gd.T @ (Me @ (Gd @ phi))
    assert_close_vec(fused, chain)
    # symmetry <x, A y> == <A x, y>
    psi = rng.standard_normal(mesh.nN)
    assert abs(psi @ fused - (A @ psi) @ phi) <= 1e-10 * abs(psi @ fused)
    # constants are in the null space
    assert np.abs(A @ np.ones(mesh.nN)).max() < 1e-9 * np.abs(fused).max()


@pytest.mark.parametrize("mesh_name", list(MESHES))
def test_maxwell_fused(mesh_name):
    from discretize_b200.tensor_ops import MaxwellOperator

    mesh, grid = make_pair(mesh_name)
    rng = np.random.default_rng(9)
    sigma = np.exp(rng.standard_normal(grid.nC))
    mui = np.exp(rng.standard_normal(grid.nC))
    e = rng.standard_normal(grid.nE)
    C = T.edge_curl(grid)
    ref = (C.T @ T.inner_product(grid, "F", mui) @ C + T.inner_product(grid, "E", sigma)).tocsr()
    Cd = mesh.edge_curl
    A = Cd.T @ mesh.get_face_inner_product(mui) @ Cd + mesh.get_edge_inner_product(sigma)
    assert isinstance(A, MaxwellOperator)
    assert_close_vec(A @ e, ref @ e)
    chain = Cd.T @ (mesh.get_face_inner_product(mui) @ (Cd @ e)) + mesh.get_edge_inner_product(sigma) @ e
    assert_close_vec(A @ e, chain)
    assert_same_matrix(A.tocsr(), ref)


def _guarded(values, lead, dev):
    """A CUDA view that starts `lead` elements into a buffer whose every other entry is NaN: a kernel that reads even one
    element before or after its arrays (or multiplies such an element by zero) turns the result into NaN."""
    buf = torch.full((values.size + lead + 3,), float("nan"), dtype=torch.float64, device=dev)
    view = buf[lead:lead + values.size]
    view.copy_(torch.as_tensor(values))
    return view


@pytest.mark.parametrize("lead", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(33, 18, 7), (40, 24, 36), (5, 4, 3), (64, 31, 12)])
def test_fused_kernels_never_read_outside_their_arrays(shape, lead):
    """K8 and K9 stage rows with 16-byte copies although the reference's layouts only give 8-byte alignment: with the
    arrays placed at even and odd element offsets between NaNs the results must not change (the ring
    kernels used to over-read 8 bytes and to multiply the stray value by a zero measure)."""
    import discretize_b200 as dz

    rng = np.random.default_rng(31)
    h = [rng.random(n) + 0.5 for n in shape]
    mesh, grid = dz.TensorMesh(h), T.Grid(h)
    dev = torch.device("cuda")
    sigma, mui = np.exp(rng.standard_normal(grid.nC)), np.exp(rng.standard_normal(grid.nC))
    phi, e = rng.standard_normal(grid.nN), rng.standard_normal(grid.nE)
    sig_d, mui_d = _guarded(sigma, lead, dev), _guarded(mui, lead + 1, dev)
    G, C = T.nodal_gradient(grid), T.edge_curl(grid)
    Gd, Cd = mesh.nodal_gradient, mesh.edge_curl
    A = Gd.T @ mesh.get_edge_inner_product(sig_d) @ Gd
    y = A.apply_device(_guarded(phi, lead, dev), out=_guarded(np.zeros(grid.nN), lead, dev))
    assert_close_vec(y.cpu().numpy(), G.T @ (T.inner_product(grid, "E", sigma) @ (G @ phi)))
    B = Cd.T @ mesh.get_face_inner_product(mui_d) @ Cd + mesh.get_edge_inner_product(sig_d)
    z = B.apply_device(_guarded(e, lead, dev), out=_guarded(np.zeros(grid.nE), lead, dev))
    assert_close_vec(z.cpu().numpy(), C.T @ (T.inner_product(grid, "F", mui) @ (C @ e)) + T.inner_product(grid, "E", sigma) @ e)


def test_maxwell_tensor_map_path_and_two_ranges():
    """Even-sized meshes (all BASELINE configs, and their z-slabs whatever the slab height) take the tensor-map TMA path;
    a two-range launch writes exactly the planes of its two ranges."""
    import ctypes

    import discretize_b200 as dz
    from discretize_b200._lib import load

    lib = load()
    lib.dzb_tm_maxwell_last_path.restype = ctypes.c_int
    rng = np.random.default_rng(32)
    for shape, want in (((64, 32, 17), 3), ((64, 32, 16), 3), ((33, 17, 9), 3), ((129, 66, 20), 2)):
        h = [rng.random(n) + 0.5 for n in shape]
        mesh, grid = dz.TensorMesh(h), T.Grid(h)
        sigma, mui = np.exp(rng.standard_normal(grid.nC)), np.exp(rng.standard_normal(grid.nC))
        e = rng.standard_normal(grid.nE)
        Cd, C = mesh.edge_curl, T.edge_curl(grid)
        A = Cd.T @ mesh.get_face_inner_product(mui) @ Cd + mesh.get_edge_inner_product(sigma)
        ref = C.T @ (T.inner_product(grid, "F", mui) @ (C @ e)) + T.inner_product(grid, "E", sigma) @ e
        assert_close_vec(A @ e, ref)
        assert lib.dzb_tm_maxwell_last_path() == want, (shape, lib.dzb_tm_maxwell_last_path())
        # planes [0, 1) and [nz - 1, nz + 1) in one launch; everything else stays NaN
        ed = torch.as_tensor(e).cuda()
        out = torch.full((grid.nE,), float("nan"), dtype=torch.float64, device="cuda")
        nz = shape[2]
        A.apply_planes2(ed, out, 0, 1, nz - 1, nz + 1)
        got = out.cpu().numpy()
        nx, ny = shape[0], shape[1]
        px, py, pz = nx * (ny + 1), (nx + 1) * ny, (nx + 1) * (ny + 1)
        offy, offz = px * (nz + 1), px * (nz + 1) + py * (nz + 1)
        written = np.zeros(grid.nE, dtype=bool)
        for k in (0, nz - 1, nz):
            written[k * px:(k + 1) * px] = True
            written[offy + k * py:offy + (k + 1) * py] = True
            if k < nz:
                written[offz + k * pz:offz + (k + 1) * pz] = True
        assert np.isnan(got[~written]).all() and not np.isnan(got[written]).any()
        assert_close_vec(got[written], ref[written])


# ------------------------------------------------------------------------------------------
# interpolation
# ------------------------------------------------------------------------------------------
LOCS = ["CC", "N", "Fx", "Fy", "Fz", "Ex", "Ey", "Ez", "CCVx", "CCVy", "CCVz"]


@pytest.mark.parametrize("mesh_name", ["uniform", "nonuniform", "thin", "graded"])
@pytest.mark.parametrize("loc", LOCS)
def test_interpolation(mesh_name, loc):
    from discretize_b200.interp import make_interpolation_operator

    mesh, grid = make_pair(mesh_name)
    rng = np.random.default_rng(10)
    pts = adversarial_points(grid, rng)
    ref = T.interpolation_matrix(grid, pts, loc)
    Q = mesh.get_interpolation_matrix(pts, loc)
    assert Q.shape == ref.shape
    assert_same_matrix(Q.tocsr(), ref, 0.0)  # indices AND weights bit-exact
    v = rng.standard_normal(ref.shape[1])
    y = rng.standard_normal(ref.shape[0])
    assert_close_vec(Q @ v, ref @ v)
    assert_close_vec(Q.T @ y, ref.T @ y)
    Qf = make_interpolation_operator(mesh, pts, loc, store=False)  # fused locate+apply
    assert_close_vec(Qf @ v, ref @ v)


@pytest.mark.parametrize("loc", LOCS)
def test_interpolation_vs_reference_fixture(loc):
    import discretize_b200 as dz

    g = load_golden()
    mesh = dz.TensorMesh([g["hx"], g["hy"], g["hz"]], origin=g["origin"])
    assert_same_matrix(mesh.get_interpolation_matrix(g["interp_pts"], loc).tocsr(), golden_matrix(g, "interp_" + loc), 0.0)


@pytest.mark.parametrize("loc", ["CC", "N", "Ex"])
def test_interpolation_outside(loc):
    mesh, grid = make_pair("nonuniform")
    rng = np.random.default_rng(11)
    pts = adversarial_points(grid, rng, 60)
    pts[5] += 100.0
    pts[17, 1] -= 50.0
    with pytest.raises(ValueError, match="Points outside of mesh"):
        mesh.get_interpolation_matrix(pts, loc)
    with pytest.raises(TypeError, match="locType"):
        mesh.get_interpolation_matrix(pts, locType=loc)
    ref = T.interpolation_matrix(grid, pts, loc, zeros_outside=True)
    Q = mesh.get_interpolation_matrix(pts, loc, zeros_outside=True)
    v = rng.standard_normal(ref.shape[1])
    out = Q @ v
    assert out[5] == 0.0 and out[17] == 0.0
    assert_close_vec(out, ref @ v)
    assert_same_matrix(Q.tocsr(), ref, 0.0)


def test_interpolation_large_roundtrip():
    """Size-independent property at 2e6 points on 64^3: a trilinear field is reproduced exactly
    (to rounding) at interior points, for nodes and for edge/face grids inside their hulls."""
    import discretize_b200 as dz

    mesh = dz.TensorMesh([64, 64, 64])
    rng = np.random.default_rng(12)
    pts = rng.random((2_000_000, 3)) * 0.9 + 0.05
    f = lambda X: 1.0 + 2.0 * X[:, 0] - 0.5 * X[:, 1] + 3.0 * X[:, 2]
    for loc, grid_pts in (("N", mesh.nodes), ("Ex", mesh.edges_x), ("Fz", mesh.faces_z), ("CC", mesh.cell_centers)):
        Q = mesh.get_interpolation_matrix(pts, loc)
        ncols = Q.shape[1]
        v = np.zeros(ncols)
        off = {"N": 0, "CC": 0, "Ex": 0, "Fz": mesh.nFx + mesh.nFy}[loc]
        v[off:off + grid_pts.shape[0]] = f(grid_pts)
        assert np.abs(Q @ v - f(pts)).max() < 1e-12


# ------------------------------------------------------------------------------------------
# LinearOperator protocol / algebra
# ------------------------------------------------------------------------------------------
def test_linear_operator_protocol():
    mesh, grid = make_pair("nonuniform")
    D = mesh.face_divergence
    Dref = T.face_divergence(grid)
    assert isinstance(D, sla.LinearOperator)
    assert D.dtype == np.float64 and D.shape == Dref.shape
    rng = np.random.default_rng(13)
    x = rng.standard_normal(D.shape[1])
    X = rng.standard_normal((D.shape[1], 3))
    y = rng.standard_normal(D.shape[0])
    for got in (D @ x, D * x, D.dot(x), D.matvec(x), D(x)):
        assert isinstance(got, np.ndarray)
        assert_close_vec(got, Dref @ x)
    assert_close_vec(D.matmat(X), Dref @ X)
    assert_close_vec(D @ X, Dref @ X)
    assert_close_vec(D.rmatvec(y), Dref.T @ y)
    assert_close_vec(D.T @ y, Dref.T @ y)
    assert_close_vec(D.H @ y, Dref.T @ y)
    assert_close_vec(y @ D, Dref.T @ y)
    assert (D @ x.reshape(-1, 1)).shape == (D.shape[0], 1)
    # torch in -> torch out, stays on the device
    xd = torch.from_numpy(x).cuda()
    yd = D @ xd
    assert isinstance(yd, torch.Tensor) and yd.is_cuda
    assert_close_vec(yd.cpu().numpy(), Dref @ x)
    # complex vectors (frequency-domain fields)
    z = x + 1j * rng.standard_normal(D.shape[1])
    assert_close_vec(D @ z, Dref @ z)
    # algebra
    A = 2.0 * D - D / 4.0
    assert_close_vec(A @ x, 1.75 * (Dref @ x))
    assert_close_vec((-D).T @ y, -(Dref.T @ y))
    S = D.T @ D
    assert_close_vec(S @ x, Dref.T @ (Dref @ x))
    assert_same_matrix((D.T @ D).tocsr(), (Dref.T @ Dref).tocsr())
    with pytest.raises(ValueError):
        D @ np.ones(3)
    with pytest.raises(ValueError):
        D @ mesh.edge_curl.T
    # scipy solvers accept it
    M = mesh.get_edge_inner_product(rng.random(grid.nC) + 0.5)
    b = rng.standard_normal(M.shape[0])
    sol, info = sla.cg(M, b, rtol=1e-12, maxiter=5000)
    assert info == 0
    assert_close_vec(M @ sol, b, 1e-9)


def test_aliases_and_counts():
    mesh, grid = make_pair("nonuniform")
    assert (mesh.nC, mesh.nN, mesh.nF, mesh.nE) == (grid.nC, grid.nN, grid.nF, grid.nE)
    assert tuple(mesh.vnF) == grid.vnF and tuple(mesh.vnE) == grid.vnE
    assert mesh.aveF2CC is mesh.average_face_to_cell
    assert mesh.aveE2CCV.shape == (3 * grid.nC, grid.nE)
    np.testing.assert_array_equal(mesh.cell_volumes, grid.cell_volumes())
    np.testing.assert_array_equal(mesh.face_areas, grid.face_areas())
    np.testing.assert_array_equal(mesh.edge_lengths, grid.edge_lengths())


@pytest.mark.parametrize("mesh_name", list(MESHES))
def test_dc_cell_fused(mesh_name):
    """D Mf(rho)^-1 D^T is recognised and replaced by ONE kernel (cell-centred DC); Dirichlet (the plain product) and
    Neumann (boundary faces dropped) variants against the oracle product."""
    from discretize_b200.tensor_ops import DCCellOperator

    mesh, grid = make_pair(mesh_name)
    rng = np.random.default_rng(17)
    rho = np.exp(rng.standard_normal(mesh.nC))
    D = mesh.face_divergence
    A = D @ mesh.get_face_inner_product(rho, invert_matrix=True) @ D.T
    assert isinstance(A, DCCellOperator)
    Dr = T.face_divergence(grid)
    MfI = T.inner_product(grid, "F", rho, invert_matrix=True)
    ref = (Dr @ MfI @ Dr.T).tocsr()
    x = rng.standard_normal(mesh.nC)
    assert_close_vec(A @ x, ref @ x)
    assert_close_vec(A.T @ x, ref @ x)
    assert_same_matrix(A.tocsr(), ref)
    # Neumann: the same with the boundary-face fluxes removed
    interior = np.diff(Dr.tocsc().indptr) == 2
    refn = (Dr @ sp.diags(MfI.diagonal() * interior) @ Dr.T).tocsr()
    An = DCCellOperator(mesh, A.rho, neumann=True)
    assert_close_vec(An @ x, refn @ x)
    assert_same_matrix(An.tocsr(), refn)
    # the on-the-fly variant (face masses generated from rho inside the kernel)
    assert_close_vec(DCCellOperator(mesh, A.rho, cache_inverse_mass=False) @ x, ref @ x)
    assert_close_vec(DCCellOperator(mesh, A.rho, neumann=True, cache_inverse_mass=False) @ x, refn @ x)
    # several z-chunks and x-blocks
    if mesh_name == "wide":
        h = [rng.random(n) + 0.5 for n in (70, 9, 70)]
        import discretize_b200 as dz

        m2, g2 = dz.TensorMesh(h), T.Grid(h)
        rho2 = np.exp(rng.standard_normal(m2.nC))
        A2 = m2.face_divergence @ m2.get_face_inner_product(rho2, invert_matrix=True) @ m2.face_divergence.T
        D2 = T.face_divergence(g2)
        x2 = rng.standard_normal(m2.nC)
        assert_close_vec(A2 @ x2, D2 @ (T.inner_product(g2, "F", rho2, invert_matrix=True) @ (D2.T @ x2)))


@pytest.mark.parametrize("mesh_name", ["graded", "wide"])
@pytest.mark.parametrize("loc", ["CC", "N", "Ez", "Fy"])
def test_interpolation_compact_rows_expand_to_the_full_tables(mesh_name, loc):
    """K10c keeps 28 bytes per point (first column + three weights, sign = clamped axis); expanded, they must be the very
    tables the (npts, 8) locate kernel writes: same columns, same weights, bit for bit -- including points in the
    half-cell clamp zones, on the boundary, and NaN / infinite coordinates."""
    import torch

    from discretize_b200._lib import check, load
    from discretize_b200.interp import InterpolationOperator
    from discretize_b200.operators import ptr, stream_ptr, to_device

    mesh, grid = make_pair(mesh_name)
    rng = np.random.default_rng(43)
    lo = np.array([t[0] for t in grid.nodes])
    hi = np.array([t[-1] for t in grid.nodes])
    pts = lo + rng.random((4000, 3)) * (hi - lo)
    pts[:50] = lo
    pts[50:100] = hi
    pts[100:140, 0] = grid.nodes[0][rng.integers(0, grid.nodes[0].size, 40)]      # exactly on grid lines
    pts[140:150, 1] = [np.nan, np.inf, -np.inf, 1e300, -1e300] * 2
    Q = InterpolationOperator(mesh, to_device(pts), loc, False, None)
    cols, w = Q.tables()
    full_c = torch.empty_like(cols)
    full_w = torch.empty_like(w)
    check(load().dzb_interp_locate(*Q._axes_args(), ptr(Q.pts), pts.shape[0], Q.col_offset, ptr(full_c), 0, ptr(full_w),
                                   stream_ptr()), "dzb_interp_locate")
    assert torch.equal(cols, full_c)
    assert torch.equal(w.view(torch.int64), full_w.view(torch.int64))
    assert (Q.wx < 0).any() and (Q.wy < 0).any()                # clamped axes are present in the sample
    # apply / transpose on those rows against the stored-table kernels
    v = to_device(rng.standard_normal(Q.shape[1]))
    x = to_device(rng.standard_normal(Q.shape[0]))
    good = torch.isfinite(to_device(pts)).all(dim=1)
    y_full = torch.empty(Q.shape[0], dtype=torch.float64, device=v.device)
    check(load().dzb_interp_apply(ptr(full_c), 0, ptr(full_w), None, pts.shape[0], ptr(v), ptr(y_full), stream_ptr()),
          "dzb_interp_apply")
    y = Q.apply_device(v)
    scale = float(y_full[good].abs().max())
    assert float((y - y_full)[good].abs().max()) <= RTOL * scale
    z_full = torch.zeros(Q.shape[1], dtype=torch.float64, device=v.device)
    check(load().dzb_interp_apply_t(ptr(full_c), 0, ptr(full_w), None, pts.shape[0], ptr(x), ptr(z_full), stream_ptr()),
          "dzb_interp_apply_t")
    z = Q.T.apply_device(x)
    assert float((z - z_full).abs().max()) <= RTOL * float(z_full.abs().max())


@pytest.mark.parametrize("loc", ["CC", "N", "Ey", "Fx", "CCVz"])
@pytest.mark.parametrize("zeros_outside", [False, True])
def test_interpolation_bucketed_tables(loc, zeros_outside):
    """K11 sorts the table rows of large operators by grid index before the first apply; the operator, its transpose
    and its matrix must not change (here the sort is forced on a small operator)."""
    mesh, grid = make_pair("graded")
    rng = np.random.default_rng(41)
    lo = np.array([t[0] for t in grid.nodes]) - (5.0 if zeros_outside else 0.0)
    hi = np.array([t[-1] for t in grid.nodes]) + (5.0 if zeros_outside else 0.0)
    pts = lo + rng.random((3000, 3)) * (hi - lo)
    Q = mesh.get_interpolation_matrix(pts, loc, zeros_outside=zeros_outside)
    v, x = rng.standard_normal(Q.shape[1]), rng.standard_normal(Q.shape[0])
    y0, z0, M0 = Q @ v, Q.T @ x, Q.tocsr()
    Q.bucket()
    assert Q.perm is not None and not np.array_equal(Q.perm.cpu().numpy(), np.arange(pts.shape[0]))
    assert np.array_equal(Q @ v, y0)                      # same products, same order of the 8 terms: bitwise
    Q.unpermute = "scatter"
    assert np.array_equal(Q @ v, y0)
    Q.unpermute = "gather"
    assert_close_vec(Q.T @ x, z0)                         # atomics: order of the additions differs
    assert_same_matrix(Q.tocsr(), M0)
    assert_same_matrix(M0, T.interpolation_matrix(grid, pts, loc, zeros_outside=zeros_outside))


def test_model_less_and_scalar_composites_fuse():
    """The reference's own Maxwell composition uses get_face_inner_product() with NO model
    (tests/boundaries/test_boundary_maxwell.py:95-99): it must reach the fused kernels, like scalar models."""
    from discretize_b200.tensor_ops import DCNodalOperator, MaxwellOperator

    mesh, grid = make_pair("wide")
    rng = np.random.default_rng(51)
    sigma = np.exp(rng.standard_normal(grid.nC))
    e, phi = rng.standard_normal(grid.nE), rng.standard_normal(grid.nN)
    C, G = T.edge_curl(grid), T.nodal_gradient(grid)
    Cd, Gd = mesh.edge_curl, mesh.nodal_gradient
    ones = np.ones(grid.nC)
    A = Cd.T @ mesh.get_face_inner_product() @ Cd + mesh.get_edge_inner_product(sigma)
    assert isinstance(A, MaxwellOperator)
    assert_close_vec(A @ e, C.T @ (T.inner_product(grid, "F", ones) @ (C @ e)) + T.inner_product(grid, "E", sigma) @ e)
    B = Cd.T @ mesh.get_face_inner_product(2.5) @ Cd + 3.0 * mesh.get_edge_inner_product()
    assert isinstance(B, MaxwellOperator)
    assert_close_vec(B @ e, C.T @ (T.inner_product(grid, "F", 2.5 * ones) @ (C @ e)) + 3.0 * (T.inner_product(grid, "E", ones) @ e))
    D = Gd.T @ mesh.get_edge_inner_product(0.7) @ Gd
    assert isinstance(D, DCNodalOperator)
    assert_close_vec(D @ phi, G.T @ (T.inner_product(grid, "E", 0.7 * ones) @ (G @ phi)))


def test_frequency_domain_maxwell_operator():
    """A = C^T Mf(mui) C + 1j omega Me(sigma), the system SimPEG's frequency-domain Maxwell solves apply: complex scalars
    in the operator algebra, fused form, complex / real inputs, transpose, adjoint and matrix."""
    from discretize_b200.tensor_ops import ComplexMaxwellOperator

    mesh, grid = make_pair("nonuniform")
    rng = np.random.default_rng(52)
    sigma, mui = np.exp(rng.standard_normal(grid.nC)), np.exp(rng.standard_normal(grid.nC))
    omega = 2 * np.pi * 3.0
    C = T.edge_curl(grid)
    K, M = (C.T @ T.inner_product(grid, "F", mui) @ C).tocsr(), T.inner_product(grid, "E", sigma)
    ref = (K + 1j * omega * M).tocsr()
    Cd = mesh.edge_curl
    A = Cd.T @ mesh.get_face_inner_product(mui) @ Cd + 1j * omega * mesh.get_edge_inner_product(sigma)
    assert isinstance(A, ComplexMaxwellOperator) and A.dtype == np.complex128
    e = rng.standard_normal(grid.nE) + 1j * rng.standard_normal(grid.nE)
    assert_close_vec(A @ e, ref @ e)                        # K9c: one launch on the interleaved complex field
    assert A.apply_device_complex(torch.as_tensor(e).cuda()) is not None
    A.fused = False
    assert_close_vec(A @ e, ref @ e)                        # two-part path: K9 on a*sigma, K5 on b*sigma
    A.fused = True
    assert_close_vec(A @ e.real, ref @ e.real)
    assert_close_vec(A.T @ e, ref.T @ e)
    assert_close_vec(A.H @ e, ref.conj().T @ e)
    et = torch.as_tensor(e).cuda()
    assert_close_vec((A @ et).cpu().numpy(), ref @ e)
    got = A.tocsr()
    assert np.abs((got - ref)).max() <= 1e-12 * np.abs(ref).max()
    # (0.5 - 2j) Me inside a generic (unfused) complex composite, and a complex multiple of a real operator
    Me = mesh.get_edge_inner_product(sigma)
    S = Me + (0.5 - 2j) * Me
    assert S.is_complex
    assert_close_vec(S @ e, (1.5 - 2j) * (M @ e))
    assert_close_vec((S.H @ e), (1.5 + 2j) * (M @ e))
    assert_close_vec(((2j * Cd) @ e), 2j * (C @ e))


def test_dc_nodal_multi_rhs():
    """matmat on the fused DC operator: all right-hand sides in one launch, for numpy (C- and F-ordered) and CUDA blocks."""
    mesh, grid = make_pair("wide")
    rng = np.random.default_rng(61)
    sigma = np.exp(rng.standard_normal(grid.nC))
    G = T.nodal_gradient(grid)
    ref = (G.T @ T.inner_product(grid, "E", sigma) @ G).tocsr()
    Gd = mesh.nodal_gradient
    A = Gd.T @ mesh.get_edge_inner_product(sigma) @ Gd
    X = rng.standard_normal((grid.nN, 5))
    want = ref @ X
    assert_close_vec(A @ X, want)
    assert_close_vec(A.matmat(np.asfortranarray(X)), want)
    assert_close_vec(A.T @ X, want)
    Xt = torch.as_tensor(X).cuda()
    Y = A @ Xt
    assert Y.shape == (grid.nN, 5) and Y.is_cuda
    assert_close_vec(Y.cpu().numpy(), want)
    for c in range(5):                                    # bitwise the single right-hand side kernel
        assert np.array_equal(Y[:, c].cpu().numpy(), (A @ Xt[:, c].contiguous()).cpu().numpy())


@pytest.mark.parametrize("shape", [(40, 24, 36), (33, 18, 7), (64, 31, 12), (5, 4, 3), (70, 9, 37)])
@pytest.mark.parametrize("cfg", [(1, 6), (1, 8), (2, 4)])
def test_complex_maxwell_kernel_shapes(shape, cfg):
    """K9c on even / odd sized meshes (odd nx: super-row maps for the two model arrays), every block shape."""
    import ctypes

    import discretize_b200 as dz
    from discretize_b200._lib import load

    lib = load()
    lib.dzb_tune_maxwell_cplx.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.dzb_tune_maxwell_cplx.restype = None
    rng = np.random.default_rng(71)
    h = [rng.random(n) + 0.5 for n in shape]
    mesh, grid = dz.TensorMesh(h), T.Grid(h)
    sigma, mui = np.exp(rng.standard_normal(grid.nC)), np.exp(rng.standard_normal(grid.nC))
    C = T.edge_curl(grid)
    alpha = 0.3 - 2.5j
    ref = (C.T @ T.inner_product(grid, "F", mui) @ C + alpha * T.inner_product(grid, "E", sigma)).tocsr()
    Cd = mesh.edge_curl
    A = Cd.T @ mesh.get_face_inner_product(mui) @ Cd + alpha * mesh.get_edge_inner_product(sigma)
    e = rng.standard_normal(grid.nE) + 1j * rng.standard_normal(grid.nE)
    lib.dzb_tune_maxwell_cplx(*cfg)
    try:
        y = A.apply_device_complex(torch.as_tensor(e).cuda())
        assert y is not None
        assert_close_vec(y.cpu().numpy(), ref @ e)
    finally:
        lib.dzb_tune_maxwell_cplx(1, 6)

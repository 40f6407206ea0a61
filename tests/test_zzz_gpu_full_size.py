"""The five BASELINE configurations at their FULL sizes on the GPU (named to run last: several minutes, tens of GB).

Small meshes are compared element by element with the oracle elsewhere; here every configuration runs at the size
BASELINE.json names and is checked (a) against an independent second route through this library at full size (the
elementary kernels chained, the host CSR of the tree, a second locate of a subset) and (b) against the ORACLE on a piece
the oracle can assemble in seconds: stencils are local, so a slab of a few cell layers cut out of the big mesh (same
widths, same model and field values) reproduces the big result exactly on its interior planes.

Tolerances as everywhere: values <= 1e-12 relative, interpolation indices and tree counts bit-exact.
"""
import numpy as np
import pytest

from oracle import tensor_ops as T
from tests.helpers import RTOL, assert_same_matrix

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device (no CPU fallback exists)")
    yield
    torch.cuda.empty_cache()


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _randn(n, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    return torch.randn(n, dtype=torch.float64, device=dev, generator=g)


# ------------------------------------------------------------------------------------------------------------
# config 2: TensorMesh 512^3, G^T Me(sigma) G
# ------------------------------------------------------------------------------------------------------------
def test_config2_dc_nodal_512():
    import discretize_b200 as dz

    n, dev = 512, torch.device("cuda")
    mesh = dz.TensorMesh([n, n, n])
    phi = _randn(mesh.nN, 5, dev)
    sigma = torch.exp(_randn(mesh.nC, 6, dev))
    G = mesh.nodal_gradient
    Me = mesh.get_edge_inner_product(sigma)
    A = G.T @ Me @ G
    assert type(A).__name__ == "DCNodalOperator"
    y = A.apply_device(phi)
    chain = G.T.apply_device(Me.apply_device(G.apply_device(phi)))
    assert _rel(y, chain) <= RTOL
    del chain
    # oracle on slabs of 4 cell layers (bottom, middle, top): interior node planes must agree
    plane, layer = (n + 1) ** 2, n * n
    for a in (0, 253, n - 4):
        g = T.Grid([n, n, 4])               # uniform widths 1/512 in x, y; the slab's z widths must be the big mesh's
        g = T.Grid([g.h[0], g.h[1], np.ones(4) / n])
        sig_s = sigma[a * layer:(a + 4) * layer].cpu().numpy()
        phi_s = phi[a * plane:(a + 5) * plane].cpu().numpy()
        Go = T.nodal_gradient(g)
        ref = Go.T @ (T.inner_product(g, "E", sig_s) @ (Go @ phi_s))
        lo, hi = (0 if a == 0 else 1), (5 if a + 4 == n else 4)      # the mesh boundary planes are exact too
        got = y[(a + lo) * plane:(a + hi) * plane].cpu().numpy()
        want = ref[lo * plane:hi * plane]
        assert np.abs(got - want).max() <= RTOL * np.abs(want).max(), (a, np.abs(got - want).max())


# ------------------------------------------------------------------------------------------------------------
# config 5: TensorMesh 1024^3, C^T Mf(mui) C + Me(sigma)
# ------------------------------------------------------------------------------------------------------------
def test_config5_maxwell_1024():
    import ctypes

    import discretize_b200 as dz
    from discretize_b200._lib import load

    n, dev = 1024, torch.device("cuda")
    mesh = dz.TensorMesh([n, n, n])
    e = _randn(mesh.nE, 11, dev)
    sigma = torch.exp(_randn(mesh.nC, 12, dev))
    mui = torch.exp(_randn(mesh.nC, 13, dev))
    C = mesh.edge_curl
    Mf, Me = mesh.get_face_inner_product(mui), mesh.get_edge_inner_product(sigma)
    A = C.T @ Mf @ C + Me
    assert type(A).__name__ == "MaxwellOperator"
    y = A.apply_device(e)
    lib = load()
    lib.dzb_tm_maxwell_last_path.restype = ctypes.c_int
    assert lib.dzb_tm_maxwell_last_path() == 3          # the tensor-map TMA kernel served it
    # full size against the elementary kernels, one temporary at a time (each is 25.8 GB)
    t1 = C.apply_device(e)
    t2 = Mf.apply_device(t1)
    del t1
    chain = C.T.apply_device(t2)
    del t2
    chain += Me.apply_device(e)
    assert _rel(y, chain) <= RTOL
    del chain
    torch.cuda.empty_cache()
    # oracle on slabs of 3 cell layers
    nn = n + 1
    px, py, pz, layer = n * nn, nn * n, nn * nn, n * n
    offy, offz = px * nn, px * nn + py * nn
    hx = np.ones(n) / n
    for a in (0, 510, n - 3):
        g = T.Grid([hx, hx, np.ones(3) / n])
        sl = slice(a * layer, (a + 3) * layer)
        e_s = torch.cat([e[a * px:(a + 4) * px], e[offy + a * py:offy + (a + 4) * py], e[offz + a * pz:offz + (a + 3) * pz]])
        Co = T.edge_curl(g)
        ref = Co.T @ (T.inner_product(g, "F", mui[sl].cpu().numpy()) @ (Co @ e_s.cpu().numpy())) \
            + T.inner_product(g, "E", sigma[sl].cpu().numpy()) @ e_s.cpu().numpy()
        lo, hi = (0 if a == 0 else 1), (4 if a + 3 == n else 3)      # node planes that see all their neighbours
        spx, spy = px * 4, py * 4
        for name, got, want in (
                ("Ex", y[(a + lo) * px:(a + hi) * px], ref[lo * px:hi * px]),
                ("Ey", y[offy + (a + lo) * py:offy + (a + hi) * py], ref[spx + lo * py:spx + hi * py]),
                ("Ez", y[offz + a * pz:offz + (a + 3) * pz], ref[spx + spy:])):
            got = got.cpu().numpy()
            assert np.abs(got - want).max() <= RTOL * np.abs(want).max(), (a, name, np.abs(got - want).max())


def test_frequency_domain_maxwell_512_one_launch():
    """C^T Mf C + i omega Me on a complex 512^3 field (K9c, one launch) against the real kernels applied to both parts."""
    import discretize_b200 as dz

    n, dev = 512, torch.device("cuda")
    mesh = dz.TensorMesh([n, n, n])
    er, ei = _randn(mesh.nE, 21, dev), _randn(mesh.nE, 22, dev)
    sigma = torch.exp(_randn(mesh.nC, 23, dev))
    mui = torch.exp(_randn(mesh.nC, 24, dev))
    omega = 2 * np.pi * 7.0
    C = mesh.edge_curl
    Mf, Me = mesh.get_face_inner_product(mui), mesh.get_edge_inner_product(sigma)
    A = C.T @ Mf @ C + 1j * omega * Me
    assert type(A).__name__ == "ComplexMaxwellOperator"
    y = A.apply_device_complex(torch.complex(er, ei))
    assert y is not None
    K = lambda v: C.T.apply_device(Mf.apply_device(C.apply_device(v)))
    want_r = K(er) - omega * Me.apply_device(ei)
    want_i = K(ei) + omega * Me.apply_device(er)
    scale = float(torch.maximum(want_r.abs().max(), want_i.abs().max()))
    assert float((y.real - want_r).abs().max()) <= RTOL * scale
    assert float((y.imag - want_i).abs().max()) <= RTOL * scale


# ------------------------------------------------------------------------------------------------------------
# config 4: 1e8 random points on TensorMesh 512^3
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("loc", ["CC", "N", "Ex", "Fz"])
def test_config4_interpolation_1e8_points(loc):
    import discretize_b200 as dz

    n, npts, nsub, dev = 512, 100_000_000, 1_000_000, torch.device("cuda")
    mesh = dz.TensorMesh([n, n, n])
    grid = T.Grid([n, n, n])
    g = torch.Generator(device=dev)
    g.manual_seed(10)
    pts = torch.rand((npts, 3), dtype=torch.float64, device=dev, generator=g)
    Q = mesh.get_interpolation_matrix(pts, loc)
    assert Q.shape[0] == npts
    # (a) indices and weights of a 1e6-point subset: the matrix the reference would build, bit-exact pattern
    sub = pts[:nsub].cpu().numpy()
    Qs = mesh.get_interpolation_matrix(sub, loc)
    assert_same_matrix(Qs.tocsr(), T.interpolation_matrix(grid, sub, loc))
    # (b) the rows of the big operator are those rows
    v = _randn(Q.shape[1], 11, dev)
    big = Q.apply_device(v)
    assert torch.equal(big[:nsub], Qs.apply_device(v))
    # (c) every point: a trilinear interpolant reproduces linear fields (away from the half-cell clamp zone)
    kinds = {"CC": "cell_centers", "N": "nodes", "Ex": "edges_x", "Fz": "faces_z"}
    tx, ty, tz = T.location_tensor(grid, kinds[loc])
    X, Y, Z = np.meshgrid(tx, ty, tz, indexing="ij")
    lin = (0.3 + 1.7 * X - 0.9 * Y + 2.3 * Z).reshape(-1, order="F")
    full = np.zeros(Q.shape[1])
    off = {"CC": 0, "N": 0, "Ex": 0, "Fz": grid.vnF[0] + grid.vnF[1]}[loc]
    full[off:off + lin.size] = lin
    got = Q.apply_device(torch.as_tensor(full, device=dev))
    want = 0.3 + 1.7 * pts[:, 0] - 0.9 * pts[:, 1] + 2.3 * pts[:, 2]
    lo = torch.tensor([tx[0], ty[0], tz[0]], device=dev)
    hi = torch.tensor([tx[-1], ty[-1], tz[-1]], device=dev)
    inside = ((pts >= lo) & (pts <= hi)).all(dim=1)
    assert float(inside.double().mean()) > 0.98
    assert float((got - want)[inside].abs().max()) <= 1e-12 * float(want.abs().max())


# ------------------------------------------------------------------------------------------------------------
# config 3: OcTree with 10 075 955 cells
# ------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def big_tree():
    import discretize_b200 as dz

    m = dz.TreeMesh([512, 512, 512], diagonal_balance=True)
    pts = np.random.default_rng(2024).random((48000, 3)) * 0.6 + 0.2
    m.refine_points(pts, -1, [2, 2, 2], finalize=True)
    return m


def test_config3_octree_counts(big_tree):
    m = big_tree
    # the reference's counts for this recipe (measured with the reference build)
    assert (m.nC, m.nF, m.nE, m.nN) == (10_075_955, 27_534_210, 25_042_770, 7_584_516)
    assert (m.nhF, m.nhE, m.nhN) == (7_185_128, 14_313_118, 5_360_277)


@pytest.mark.parametrize("name", ["face_divergence", "edge_curl", "average_face_to_cell", "average_edge_to_cell",
                                  "average_node_to_cell"])
def test_config3_octree_spmv(big_tree, name):
    m, dev = big_tree, torch.device("cuda")
    op = getattr(m, name)
    x = _randn(op.shape[1], 7, dev)
    y = op.apply_device(x).cpu().numpy()
    M = m._host_matrix(name)                        # the reference's matrix (checked cell for cell on small trees)
    ref = M @ x.cpu().numpy()
    assert np.abs(y - ref).max() <= RTOL * np.abs(ref).max()
    w = _randn(op.shape[0], 8, dev)
    yt = op.T.apply_device(w).cpu().numpy()
    reft = M.T @ w.cpu().numpy()
    assert np.abs(yt - reft).max() <= RTOL * np.abs(reft).max()
    for key in ("P_" + name, "H_" + name, "OP_" + name):
        m._cache.pop(key, None)

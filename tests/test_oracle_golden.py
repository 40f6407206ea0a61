"""CPU: the numpy oracle against the committed reference fixtures and the reference's own
known-answer tests (runs everywhere, incl. the GPU box where /root/reference is absent)."""
import numpy as np
import pytest

from oracle import tensor_ops as T
from tests.helpers import ORACLE_OPS, assert_same_matrix, golden_matrix, load_golden


@pytest.fixture(scope="module")
def gold():
    g = load_golden()
    grid = T.Grid([g["hx"], g["hy"], g["hz"]], g["origin"])
    return g, grid


@pytest.mark.parametrize("name", list(ORACLE_OPS))
def test_elementary(gold, name):
    g, grid = gold
    assert_same_matrix(ORACLE_OPS[name](grid), golden_matrix(g, name), 1e-14)


@pytest.mark.parametrize("proj", ["F", "E"])
@pytest.mark.parametrize("mk", ["iso", "aniso", "tensor"])
@pytest.mark.parametrize("im", [False, True])
@pytest.mark.parametrize("ix", [False, True])
def test_mass(gold, proj, mk, im, ix):
    g, grid = gold
    if mk == "tensor" and ix:
        pytest.skip("reference raises: Solver needed to invert A.")
    a = T.inner_product(grid, proj, g["model_" + mk], invert_model=im, invert_matrix=ix)
    assert_same_matrix(a, golden_matrix(g, f"mass_{proj}_{mk}_{int(im)}{int(ix)}"), 1e-13)


@pytest.mark.parametrize("proj", ["F", "E"])
@pytest.mark.parametrize("mk", ["iso", "aniso", "tensor"])
@pytest.mark.parametrize("im", [False, True])
@pytest.mark.parametrize("ix", [False, True])
def test_mass_deriv(gold, proj, mk, im, ix):
    g, grid = gold
    if mk == "tensor" and (ix or im):
        pytest.skip("reference: not implemented for full tensors")
    a = T.inner_product_deriv(grid, proj, g["model_" + mk], g["deriv_v_" + proj], invert_model=im, invert_matrix=ix)
    assert_same_matrix(a, golden_matrix(g, f"deriv_{proj}_{mk}_{int(im)}{int(ix)}"), 1e-13)


def test_composites(gold):
    g, grid = gold
    G, C = T.nodal_gradient(grid), T.edge_curl(grid)
    Me = T.inner_product(grid, "E", g["model_iso"])
    Mf = T.inner_product(grid, "F", g["model_mui"])
    assert_same_matrix(G.T @ Me @ G, golden_matrix(g, "dc_nodal"), 1e-13)
    assert_same_matrix(C.T @ Mf @ C + Me, golden_matrix(g, "maxwell"), 1e-13)


@pytest.mark.parametrize("loc", ["CC", "N", "Fx", "Fy", "Fz", "Ex", "Ey", "Ez", "CCVx", "CCVy", "CCVz"])
def test_interpolation(gold, loc):
    g, grid = gold
    assert_same_matrix(T.interpolation_matrix(grid, g["interp_pts"], loc), golden_matrix(g, "interp_" + loc), 0.0)


# ---- the reference's own known-answer integrals (tests/base/test_tensor_innerproduct.py:31-164) ----
def _analytic_fields(X, Y, Z):
    ex = X**2 + Y * Z
    ey = (Z**2) * X + Y * Z
    ez = Y**2 + X * Z
    return ex, ey, ez


def _sigma(XC, rep):
    x, y, z = XC[:, 0], XC[:, 1], XC[:, 2]
    s1 = x * y + 1
    if rep == 1:
        return s1
    s2, s3 = x * z + 2, 3 + z * y
    if rep == 3:
        return np.c_[s1, s2, s3]
    return np.c_[s1, s2, s3, 0.1 * x * y * z, 0.2 * x * y, 0.1 * z]


KNOWN = {1: 647.0 / 360, 3: 37.0 / 12, 6: 69881.0 / 21600}


def _grid_points(grid, kind):
    t = T.location_tensor(grid, {"Ex": "edges_x", "Ey": "edges_y", "Ez": "edges_z",
                                 "Fx": "faces_x", "Fy": "faces_y", "Fz": "faces_z", "CC": "cell_centers"}[kind])
    m = np.meshgrid(*t, indexing="ij")
    return np.stack([a.reshape(-1, order="F") for a in m], axis=1)


@pytest.mark.parametrize("proj", ["E", "F"])
@pytest.mark.parametrize("rep", [1, 3, 6])
@pytest.mark.parametrize("inv", [False, True])
def test_known_integrals_converge(proj, rep, inv):
    """u^T M(sigma) u -> the sympy-derived integral at second order (reference uses OrderTest)."""
    errs = []
    for n in (8, 16):
        grid = T.Grid([n, n, n])
        comps = []
        for c, ax in enumerate("xyz"):
            P = _grid_points(grid, proj + ax)
            comps.append(_analytic_fields(P[:, 0], P[:, 1], P[:, 2])[c])
        u = np.concatenate(comps)
        sig = _sigma(_grid_points(grid, "CC"), rep)
        if inv:
            M = T.inner_product(grid, proj, T.invert_tensor_model(grid, sig).reshape(sig.shape, order="F")
                                if rep > 1 else 1.0 / sig, invert_model=True)
        else:
            M = T.inner_product(grid, proj, sig)
        errs.append(abs(u @ (M @ u) - KNOWN[rep]))
    order = np.log2(errs[0] / errs[1])
    assert order > 1.9, (errs, order)


def test_mimetic_identities():
    """D C = 0 and C G = 0 (reference tests/base/test_operators.py:770-802)."""
    rng = np.random.default_rng(0)
    grid = T.Grid([rng.random(5) + 0.5, rng.random(4) + 0.5, rng.random(6) + 0.5])
    D, C, G = T.face_divergence(grid), T.edge_curl(grid), T.nodal_gradient(grid)
    assert abs(D @ C).max() < 1e-12
    assert abs(C @ G).max() < 1e-12

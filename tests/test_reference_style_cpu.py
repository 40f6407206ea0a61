"""Tests in the style of the reference's own suite, on the host-assembled matrices of this package:
mimetic identities on an OcTree with hanging entities (tests/tree/test_tree.py:215-233), equivalence of a fully refined
OcTree with the TensorMesh after ``permute_*`` (tests/tree/test_tree.py:161-213), order of accuracy of OcTree
interpolation (tests/tree/test_tree_interpolation.py:39-215) and Taylor tests of the surface / line inner-product
derivatives (tests/base/test_tensor_innerproduct_derivs.py harness).  No GPU: ``_host_matrix`` / ``.tocsr()`` of
host-assembled operators only."""
import numpy as np
import pytest
import scipy.sparse as sp

import discretize_b200 as dz
from oracle import tensor_ops as oracle


def hanging_tree(n=16, levels=(3, 4)):
    m = dz.TreeMesh([np.full(n, 1.0 / n)] * 3)
    m.refine_ball(np.array([[0.3, 0.4, 0.5], [0.7, 0.6, 0.4]]), [0.2, 0.15], list(levels))
    assert m.n_hanging_faces > 0 and m.n_hanging_edges > 0
    return m


def test_tree_mimetic_identities():
    m = hanging_tree()
    D, C, G = m._host_matrix("face_divergence"), m._host_matrix("edge_curl"), m._host_matrix("nodal_gradient")
    scale = abs(D).max() * abs(C).max()
    assert abs(D @ C).max() <= 1e-13 * scale
    assert abs(C @ G).max() <= 1e-13 * abs(C).max() * abs(G).max()


def test_fully_refined_tree_is_the_tensor_mesh_after_permutation():
    h = [np.array([1.0, 2.0, 1.5, 0.5]), np.array([0.7, 1.1, 0.9, 1.3]), np.array([2.0, 1.0, 0.5, 1.5])]
    t = dz.TreeMesh(h)
    t.refine(-1)
    assert t.n_cells == 64 and t.n_hanging_faces == 0
    Pc, Pf, Pe = (t.permute_cells.tocsr(), t.permute_faces.tocsr(), t.permute_edges.tocsr())
    g = oracle.Grid(h)
    vol, area, length = g.cell_volumes(), g.face_areas(), g.edge_lengths()
    assert np.abs(Pc @ t.cell_volumes - vol).max() <= 1e-14 * vol.max()
    assert np.abs(Pf @ t.face_areas - area).max() <= 1e-14 * area.max()
    assert np.abs(Pe @ t.edge_lengths - length).max() <= 1e-14 * length.max()
    D = Pc @ t._host_matrix("face_divergence") @ Pf.T
    C = Pf @ t._host_matrix("edge_curl") @ Pe.T
    assert abs(D - oracle.face_divergence(g)).max() <= 1e-13 * abs(D).max()
    assert abs(C - oracle.edge_curl(g)).max() <= 1e-13 * abs(C).max()
    Ae = Pc @ t._host_matrix("average_edge_to_cell") @ Pe.T
    assert abs(Ae - oracle.average_edge_to_cell(g)).max() <= 1e-14


FIELDS = {
    "CC": lambda x: np.cos(2 * np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) + x[:, 2] ** 2,
    "N": lambda x: np.cos(2 * np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) + x[:, 2] ** 2,
    "Ex": lambda x: np.cos(2 * np.pi * x[:, 1]) + np.sin(np.pi * x[:, 2]) * x[:, 0],
    "Fz": lambda x: np.cos(2 * np.pi * x[:, 0]) * x[:, 1] + np.sin(np.pi * x[:, 2]),
}


@pytest.mark.parametrize("loc", sorted(FIELDS))
def test_tree_interpolation_order_of_accuracy(loc):
    """Max error at random interior points drops ~4x per halving of h on meshes refined to the finest level around the
    probe region (second order), like the reference's OrderTest for OcTree interpolation."""
    rng = np.random.default_rng(9)
    pts = 0.35 + 0.3 * rng.random((200, 3))
    errs = []
    for n in (8, 16, 32):
        m = dz.TreeMesh([np.full(n, 1.0 / n)] * 3)
        m.refine_box(np.array([[0.2, 0.2, 0.2]]), np.array([[0.8, 0.8, 0.8]]), [-1])
        grid = {"CC": m.cell_centers, "N": m.nodes, "Ex": m.edges_x, "Fz": m.faces_z}[loc]
        off = {"CC": 0, "N": 0, "Ex": 0, "Fz": m.n_faces_x + m.n_faces_y}[loc]
        Q = m.get_interpolation_matrix(pts, loc).tocsr()
        f = np.zeros(Q.shape[1])
        f[off:off + grid.shape[0]] = FIELDS[loc](grid)
        errs.append(np.abs(Q @ f - FIELDS[loc](pts)).max())
    orders = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert orders.min() > 1.7, (errs, orders)


@pytest.mark.parametrize("kind", ["get_face_inner_product_surface", "get_edge_inner_product_surface",
                                  "get_edge_inner_product_line"])
@pytest.mark.parametrize("invert_model", [False, True])
@pytest.mark.parametrize("invert_matrix", [False, True])
@pytest.mark.parametrize("mesh_kind", ["tensor", "tree"])
def test_surface_and_line_inner_product_derivatives_taylor(kind, invert_model, invert_matrix, mesh_kind):
    """f(m) = M(m) v: the first-order Taylor remainder with the derivative drops at second order."""
    rng = np.random.default_rng(4)
    mesh = dz.TensorMesh([rng.random(4) + .5, rng.random(3) + .5, rng.random(5) + .5]) if mesh_kind == "tensor" \
        else hanging_tree(8, (2, 3))
    n = mesh.n_edges if kind.endswith("line") else mesh.n_faces
    v = rng.random(mesh.n_faces if kind.startswith("get_face") else mesh.n_edges)
    m0, dm = rng.random(n) + 1.0, rng.random(n) * 0.5

    def f(model):
        return getattr(mesh, kind)(model, invert_model=invert_model, invert_matrix=invert_matrix).tocsr() @ v

    J = getattr(mesh, kind + "_deriv")(m0, invert_model=invert_model, invert_matrix=invert_matrix)(v).tocsr()
    e = [np.linalg.norm(f(m0 + s * dm) - f(m0) - s * (J @ dm)) for s in (1e-1, 1e-2, 1e-3)]
    if max(e) > 1e-13 * np.linalg.norm(f(m0)):         # linear in the model: the remainder is rounding only
        orders = np.log10(np.array(e[:-1]) / np.array(e[1:]))
        assert orders.min() > 1.9, (e, orders)


def test_robin_pieces_reduce_to_neumann_and_scale_with_alpha():
    """beta = 1, alpha = gamma = 0 is the natural (Neumann) condition: B = 0 and b = 0; B is linear in alpha."""
    for mesh in (dz.TensorMesh([5, 4, 3]), hanging_tree(8, (2, 3))):
        B, b = mesh.edge_divergence_weak_form_robin(0.0, 1.0, 0.0)
        assert B.tocsr().nnz == 0 or abs(B.tocsr()).max() == 0.0
        assert np.abs(b).max() == 0.0
        B1, _ = mesh.edge_divergence_weak_form_robin(1.0, 2.0, 0.0)
        B3, _ = mesh.edge_divergence_weak_form_robin(3.0, 2.0, 0.0)
        assert abs(3.0 * B1.tocsr() - B3.tocsr()).max() <= 1e-14 * abs(B3.tocsr()).max()
        # the boundary areas seen through the nodes add up to the surface area of the box
        total = -np.asarray(B1.tocsr().diagonal()).sum() * 2.0
        ext = [np.sum(x) for x in mesh.h]
        assert abs(total - 2 * (ext[0] * ext[1] + ext[1] * ext[2] + ext[0] * ext[2])) <= 1e-12 * total


KNOWN_2D = {1: 144877.0 / 360, 2: 189959.0 / 120, 3: 781427.0 / 360}      # tests/base/test_tensor_innerproduct.py:424-430


def _mesh_2d(kind, n):
    if kind == "tensor":
        return dz.TensorMesh([n, n])
    m = dz.TreeMesh([n, n])
    if kind == "quadtree_uniform":
        m.refine(-1)
    else:                      # finest level in a band, one level coarser elsewhere: hanging edges along the interface
        m.refine(lambda cell: m.max_level if abs(cell.center[1] - 0.5) < 0.25 else m.max_level - 1)
    return m


@pytest.mark.parametrize("kind", ["tensor", "quadtree_uniform", "quadtree_hanging"])
@pytest.mark.parametrize("location", ["edges", "faces"])
@pytest.mark.parametrize("sigma_test", [1, 2, 3])
@pytest.mark.parametrize("invert_model", [False, True])
def test_known_2d_inner_product_integrals(kind, location, sigma_test, invert_model):
    """The reference's own known-answer test in 2-D (TestInnerProducts2D): u^T M(sigma) u converges at second order to
    the sympy-derived integral over the unit square, for isotropic, anisotropic and full-tensor sigma, on a TensorMesh
    and on QuadTrees with and without hanging edges."""
    z = 5.0
    ex = lambda x, y: x ** 2 + y * z                      # noqa: E731
    ey = lambda x, y: (z ** 2) * x + y * z                # noqa: E731
    s1 = lambda x, y: x * y + 1                           # noqa: E731
    s2 = lambda x, y: x * z + 2                           # noqa: E731
    s3 = lambda x, y: 3 + z * y                           # noqa: E731
    errs = []
    for n in (8, 16, 32):
        m = _mesh_2d(kind, n)
        cc = m.cell_centers
        cols = [f(cc[:, 0], cc[:, 1]) for f in (s1, s2, s3)[:sigma_test]]
        sigma = np.c_[tuple(cols)] if sigma_test < 3 else np.r_[tuple(cols)]
        vec = lambda g: np.c_[ex(g[:, 0], g[:, 1]), ey(g[:, 0], g[:, 1])]      # noqa: E731
        if location == "edges":
            u = m.project_edge_vector(np.vstack((vec(m.edges_x), vec(m.edges_y))))
            get = m.get_edge_inner_product
        else:
            u = m.project_face_vector(np.vstack((vec(m.faces_x), vec(m.faces_y))))
            get = m.get_face_inner_product
        if invert_model:
            A = get(dz.utils.inverse_property_tensor(m, sigma), invert_model=True)
        else:
            A = get(sigma)
        errs.append(abs(u @ (A.tocsr() @ u) - KNOWN_2D[sigma_test]))
    orders = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert orders.min() > 0.8 * 2, (errs, orders)          # the reference's OrderTest tolerance (0.85) on order 2


@pytest.mark.parametrize("kind,expected", [("tensor", 2.0), ("quadtree_uniform", 2.0), ("quadtree_hanging", 1.0)])
def test_2d_operator_order_of_accuracy(kind, expected):
    """DIV, GRAD and CURL against analytic fields in the max norm (tests/base/test_operators.py:45-100,
    tests/tree/test_tree_operators.py:110-292): second order on uniform cells, first order across hanging interfaces."""
    tau = 2 * np.pi
    err = {"div": [], "grad": [], "curl": []}
    for n in (16, 32, 64):
        m = _mesh_2d(kind, n)
        fx, fy, ex_, ey_, cc, nd = m.faces_x, m.faces_y, m.edges_x, m.edges_y, m.cell_centers, m.nodes
        D, G, C = (getattr(m, name).tocsr() for name in ("face_divergence", "nodal_gradient", "edge_curl"))
        u = np.r_[np.sin(tau * fx[:, 0]), np.sin(tau * fy[:, 1])]
        err["div"].append(np.abs(D @ u - tau * (np.cos(tau * cc[:, 0]) + np.cos(tau * cc[:, 1]))).max())
        phi = np.cos(tau * nd[:, 0]) * np.sin(tau * nd[:, 1])
        want = np.r_[-tau * np.sin(tau * ex_[:, 0]) * np.sin(tau * ex_[:, 1]), tau * np.cos(tau * ey_[:, 0]) * np.cos(tau * ey_[:, 1])]
        err["grad"].append(np.abs(G @ phi - want).max())
        e = np.r_[np.cos(tau * ex_[:, 1]), np.sin(tau * ey_[:, 0])]               # e = (cos 2 pi y, sin 2 pi x)
        err["curl"].append(np.abs(C @ e - tau * (np.cos(tau * cc[:, 0]) + np.sin(tau * cc[:, 1]))).max())
    for name, e in err.items():
        orders = np.log2(np.array(e[:-1]) / np.array(e[1:]))
        assert orders.min() > 0.85 * expected, (name, e, orders)


@pytest.mark.parametrize("hanging,expected", [(False, 2.0), (True, 1.0)])
def test_octree_operator_order_of_accuracy(hanging, expected):
    """DIV, CURL and GRAD of the OcTree host matrices against analytic fields in the max norm
    (tests/tree/test_tree_operators.py:110-292)."""
    tau = 2 * np.pi
    err = {"div": [], "curl": [], "grad": []}
    for n in ((16, 32, 64) if hanging else (8, 16, 32)):      # the interface error only settles to first order later
        m = dz.TreeMesh([n, n, n])
        if hanging:
            m.refine(lambda cc: np.where(np.abs(cc[:, 2] - 0.5) < 0.25, m.max_level, m.max_level - 1), vectorized=True)
        else:
            m.refine(-1)
        D, C, G = (m._host_matrix(k) for k in ("face_divergence", "edge_curl", "nodal_gradient"))
        fx, fy, fz, cc, nd = m.faces_x, m.faces_y, m.faces_z, m.cell_centers, m.nodes
        ex_, ey_, ez_ = m.edges_x, m.edges_y, m.edges_z
        u = np.r_[np.sin(tau * fx[:, 0]), np.sin(tau * fy[:, 1]), np.sin(tau * fz[:, 2])]
        err["div"].append(np.abs(D @ u - tau * np.cos(tau * cc).sum(axis=1)).max())
        # e = (cos 2 pi y, cos 2 pi z, cos 2 pi x)  ->  curl e = 2 pi (sin 2 pi z, sin 2 pi x, sin 2 pi y)
        e = np.r_[np.cos(tau * ex_[:, 1]), np.cos(tau * ey_[:, 2]), np.cos(tau * ez_[:, 0])]
        want = tau * np.r_[np.sin(tau * fx[:, 2]), np.sin(tau * fy[:, 0]), np.sin(tau * fz[:, 1])]
        err["curl"].append(np.abs(C @ e - want).max())
        phi = np.cos(tau * nd[:, 0]) * np.cos(tau * nd[:, 1]) * np.cos(tau * nd[:, 2])
        g = lambda p, a: -tau * np.prod([np.sin(tau * p[:, b]) if b == a else np.cos(tau * p[:, b]) for b in range(3)], axis=0)  # noqa: E731
        err["grad"].append(np.abs(G @ phi - np.r_[g(ex_, 0), g(ey_, 1), g(ez_, 2)]).max())
    for name, e in err.items():
        orders = np.log2(np.array(e[:-1]) / np.array(e[1:]))
        assert orders.min() > 0.85 * expected, (name, e, orders)

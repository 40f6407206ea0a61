"""CPU: the from-scratch OcTree builder (octree.py + tree_build.py + the host assembly in tree_mesh.py)
against fixtures made by the reference (tests/golden/tree_small.npz) and, where /root/reference exists,
against the reference itself on random trees."""
import numpy as np
import pytest

from discretize_b200.tree_mesh import TreeMesh, TreeMeshNotFinalizedError
from tests.golden.make_golden_tree import COUNTS, TREE_OPS
from tests.helpers import assert_same_matrix, golden_matrix, load_golden


def _mesh_from_fixture(g, tag, how):
    h = [g[f"{tag}_h{a}"] for a in "xyz"]
    m = TreeMesh(h, origin=g[f"{tag}_origin"], diagonal_balance=bool(g[f"{tag}_db"]))
    if how == "refine":
        m.refine_points(g[f"{tag}_pts"], -1, [1, 1], finalize=True)
    else:
        m._set_state(g[f"{tag}_state_indexes"], g[f"{tag}_state_levels"])
        m.finalize()
    return m


@pytest.fixture(scope="module")
def gold():
    return load_golden("tree_small.npz")


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("how", ["refine", "state"])
def test_tree_matches_reference_fixture(gold, tag, how):
    g = gold
    m = _mesh_from_fixture(g, tag, how)
    idx, lev = m._tree.state()
    np.testing.assert_array_equal(idx, g[f"{tag}_state_indexes"])   # same cells, same order
    np.testing.assert_array_equal(lev, g[f"{tag}_state_levels"])
    np.testing.assert_array_equal(np.array([getattr(m, c) for c in COUNTS]), g[f"{tag}_counts"])
    for name in ("cell_volumes", "face_areas", "edge_lengths", "cell_centers", "nodes"):
        np.testing.assert_array_equal(getattr(m, name), g[f"{tag}_{name}"])
    for name in TREE_OPS:
        assert_same_matrix(m._host_matrix(name), golden_matrix(g, f"{tag}_{name}"), 1e-13)


def test_not_finalized_and_errors():
    m = TreeMesh([8, 8, 8])
    with pytest.raises(TreeMeshNotFinalizedError):
        m.n_cells
    with pytest.raises(ValueError, match="power of 2"):
        TreeMesh([6, 8, 8])
    m.refine(2, finalize=True)
    assert m.n_cells == 64 and m.n_hanging_faces == 0 and m.nhN == 0
    with pytest.raises(RuntimeError):
        m.refine(3)


def test_full_refinement_equals_tensor_counts():
    """A fully refined OcTree has the TensorMesh counts (reference tests/tree/test_tree.py:161-213)."""
    m = TreeMesh([4, 4, 4])
    m.refine(-1, finalize=True)
    assert (m.nC, m.nN, m.nF, m.nE) == (64, 125, 240, 300)
    assert m.n_hanging_edges == 0
    D = m._host_matrix("face_divergence")
    C = m._host_matrix("edge_curl")
    G = m._host_matrix("nodal_gradient")
    assert abs(D @ C).max() < 1e-12 and abs(C @ G).max() < 1e-12


def test_averages_reproduce_constants_on_refined_tree(gold):
    # (D C = 0 / C G = 0 only hold on fully refined trees -- that is all the reference tests, too:
    #  tests/tree/test_tree.py:215-233 -- the deflated operators of a hanging tree do not commute.)
    m = _mesh_from_fixture(gold, "a", "state")
    for name in ("average_face_to_cell", "average_edge_to_cell", "average_node_to_cell"):
        A = m._host_matrix(name)
        np.testing.assert_allclose(A @ np.ones(A.shape[1]), 1.0, atol=1e-14)  # constants are reproduced


@pytest.mark.parametrize("db", [False, True])
def test_random_trees_against_reference(reference, db):
    import warnings

    warnings.simplefilter("ignore")
    rng = np.random.default_rng(5)
    for shape in ([16, 16, 16], [32, 16, 8]):
        h = [rng.random(s) + 0.5 for s in shape]
        ext = np.array([x.sum() for x in h])
        pts = (rng.random((12, 3)) * 0.8 + 0.1) * ext
        r = reference.TreeMesh(h, diagonal_balance=db)
        n = TreeMesh(h, diagonal_balance=db)
        for mesh in (r, n):
            mesh.refine_ball(pts[:4], [0.9, 1.3, 0.4, 2.0], [-1, -2, -1, 3], finalize=False)
            mesh.insert_cells(pts[4:], np.array([-1, -1, -2, -1, 2, 3, -1, -1]), finalize=False)
            mesh.refine_points(pts[:3], -1, [2, 1, 1], finalize=True)
        ri, rl = r.__getstate__()
        ni, nl = n._tree.state()
        np.testing.assert_array_equal(ri, ni)
        np.testing.assert_array_equal(rl, nl)
        assert (r.nhN, r.nhEx, r.nhFz, r.nN, r.nE, r.nF) == (n.nhN, n.nhEx, n.nhFz, n.nN, n.nE, n.nF)
        for name in ("face_divergence", "edge_curl", "nodal_gradient", "average_edge_to_cell", "average_node_to_cell"):
            assert_same_matrix(n._host_matrix(name), getattr(r, name), 1e-13)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_tree_node_and_edge_to_face_averages_match_reference_fixture(gold, tag):
    """average_node_to_edge / average_node_to_face / average_edge_to_face (tree_ext.pyx:4350-5078)."""
    from tests.golden.make_golden_tree_n2 import TREE_N2_OPS

    g2 = load_golden("tree_n2_small.npz")
    m = _mesh_from_fixture(gold, tag, "refine")
    for name in TREE_N2_OPS:
        A = m._host_matrix(name)
        assert_same_matrix(A, golden_matrix(g2, f"{tag}_{name}"), 1e-13)
        np.testing.assert_allclose(A @ np.ones(A.shape[1]), 1.0, rtol=1e-13)   # averages reproduce constants


@pytest.mark.parametrize("tag", ["a", "b"])
def test_tree_cell_centred_operators_match_reference_fixture(gold, tag):
    """stencil_cell_gradient*, average_cell_to_face*, average_cell_to_total_face_*, face_*_divergence and the boundary
    face indices (tree_ext.pyx:2824-2920, 3471-3853, 5080-5512), explicit (0, 0) zero of the reference included."""
    from tests.golden.make_golden_tree_n2 import TREE_CC_METHODS, TREE_CC_OPS

    g2 = load_golden("tree_n2_small.npz")
    m = _mesh_from_fixture(gold, tag, "refine")
    for name in TREE_CC_OPS + TREE_CC_METHODS:
        assert_same_matrix(m._host_matrix(name), golden_matrix(g2, f"{tag}_{name}"), 1e-13)
    for k, ind in enumerate(m.face_boundary_indices):
        np.testing.assert_array_equal(ind, g2[f"{tag}_face_boundary_{k}"])
    A = m._host_matrix("average_cell_to_face")
    np.testing.assert_allclose(A @ np.ones(m.nC), 1.0, rtol=1e-13)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_tree_point_location_and_node_interpolation_match_reference_fixture(gold, tag):
    """point2index (tree.cpp:758-766, 1499-1516: ties go high between roots, low inside) and the node interpolation
    matrix with hanging nodes eliminated (tree_ext.pyx:6116-6179), points on nodes / centres / outside included."""
    g2 = load_golden("tree_n2_small.npz")
    m = _mesh_from_fixture(gold, tag, "refine")
    pts = g2[f"{tag}_interp_pts"]
    np.testing.assert_array_equal(m.point2index(pts), g2[f"{tag}_point2index"])
    assert_same_matrix(m._host_node_interpolation(pts, False), golden_matrix(g2, f"{tag}_interp_N"), 1e-13)
    assert_same_matrix(m._host_node_interpolation(pts, True), golden_matrix(g2, f"{tag}_interp_N_zeros_outside"), 1e-13)
    assert m.get_interpolation_matrix(pts, "Ex").shape == (pts.shape[0], m.n_edges)   # details: test_tree_interp_cpu.py
    with pytest.raises(TypeError):
        m.get_interpolation_matrix(pts, "N", locType="N")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_tree_tensor_model_derivative_matches_reference_fixture(gold, tag):
    from discretize_b200.tree_ops import _tree_tensor_mass_deriv_matrix

    g2 = load_golden("tree_n2_small.npz")
    m = _mesh_from_fixture(gold, tag, "state")
    for proj, key, v in (("F", "dMf_tensor", gold[f"{tag}_vF"]), ("E", "dMe_tensor", gold[f"{tag}_vE"])):
        a = _tree_tensor_mass_deriv_matrix(m, proj, v)
        b = golden_matrix(g2, f"{tag}_{key}")
        a.eliminate_zeros(); b.eliminate_zeros()     # the reference's sums of products keep some exact zeros
        assert_same_matrix(a, b, 1e-12)


def test_refine_box_and_bounding_box():
    """refine_box (closed-interval box/cell intersection, geom.cpp:102-112) and refine_bounding_box
    (tree_mesh.py:443-546): cell for cell the reference's tree where it is available; always: every leaf that
    intersects the box is at the requested level."""
    import warnings

    from oracle.refload import load_reference

    rng = np.random.default_rng(8)
    ref = load_reference()
    for shape, db in (((16, 16, 16), True), ((16, 8, 32), False)):
        h = [rng.random(s) + 0.5 for s in shape]
        ext = np.array([x.sum() for x in h])
        x0 = ext * rng.random((2, 3)) * 0.5
        x1 = x0 + ext * rng.random((2, 3)) * 0.4
        pts = ext * (rng.random((20, 3)) * 0.3 + 0.3)
        m1 = TreeMesh(h, diagonal_balance=db)
        m1.refine_box(x0, x1, [-1, -2], finalize=True)
        m2 = TreeMesh(h, diagonal_balance=db)
        m2.refine_bounding_box(pts, -1, [1, 2, 1], finalize=True)
        # property: leaves intersecting box 0 sit at the finest level
        t = m1._t()
        a = np.stack([m1._xs[d][t.centers[:, d] - t.half] for d in range(3)], axis=1)
        b = np.stack([m1._xs[d][t.centers[:, d] + t.half] for d in range(3)], axis=1)
        hit = ((np.maximum(x0[0], x1[0]) >= a) & (np.minimum(x0[0], x1[0]) <= b)).all(axis=1)
        assert hit.any() and (t.half[hit] == 1).all()
        if ref is not None:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r1 = ref.TreeMesh(h, diagonal_balance=db)
                r1.refine_box(x0, x1, [-1, -2], finalize=True)
                r2 = ref.TreeMesh(h, diagonal_balance=db)
                r2.refine_bounding_box(pts, -1, [1, 2, 1], finalize=True)
            for ours, theirs in ((m1, r1), (m2, r2)):
                idx, lev = ours._tree.state()
                ridx, rlev = theirs.__getstate__()
                np.testing.assert_array_equal(idx, ridx)
                np.testing.assert_array_equal(lev, rlev)
    with pytest.raises(ValueError):
        TreeMesh([8, 8, 8]).refine_box(np.zeros((2, 3)), np.ones((3, 3)), [1, 2, 3])


def test_code_table_holds_exact_values_and_overflow_is_signalled():
    """The multiplier table of a coded OcTree operator holds entries of the matrix itself (values that differ by a
    rounding error share a code, to 1e-15 relative -- not the 2e-14 of an absolute rounding), and an operator with more
    distinct weights than a code byte can name raises CodeTableOverflow so that the caller stores the values instead."""
    from discretize_b200.tree_ops import CodeTableOverflow, _group_values, encode_parts

    rng = np.random.default_rng(3)
    base = np.array([1.0 / 6.0, 1.0 / 3.0, 1.0 / 12.0, 0.5, -0.25, 1e-9 / 3.0])
    data = np.r_[base, base * (1 + 2.2e-16), base[rng.integers(0, base.size, 500)]]
    tab, code = _group_values(data)
    assert tab.size == base.size
    assert np.isin(tab, data).all()
    assert np.abs(tab[code] - data).max() <= 1e-15 * np.abs(data).max()
    assert (np.abs(tab[code] - data) <= 1e-15 * np.abs(data)).all()

    mesh = TreeMesh([8, 8, 8])
    mesh.refine_points(np.array([[0.3, 0.3, 0.3]]), -1, finalize=True)
    for name in ("face_divergence", "average_face_to_cell", "average_edge_to_cell", "average_node_to_cell"):
        geoms, parts = mesh._host_parts(name)
        M = mesh._host_matrix(name)
        indptr, indices, codes, tmult, tsel = encode_parts(geoms, parts, M.shape)
        assert np.isin(tmult, np.concatenate([np.asarray(P.data) for _, P in parts])).all()   # exact entries, never rounded
    import scipy.sparse as sp

    wide = sp.csr_matrix(np.diag(rng.random(300) + 0.5))
    with pytest.raises(CodeTableOverflow):
        encode_parts([], [(0, wide)], wide.shape)


def test_geometry_folds_into_the_code_table_on_uniform_base_meshes():
    """On an OcTree over a uniform base mesh the products multiplier * (1/h, 1/L) are a handful of values: they replace
    the per-row geometry arrays (the SpMV kernel then reads none) and reproduce the reference's matrix entry for entry;
    a strongly graded base mesh keeps its geometry arrays."""
    import scipy.sparse as sp

    from discretize_b200.tree_ops import encode_parts, fold_geometry

    mesh = TreeMesh([16, 16, 16])
    mesh.refine_points(np.array([[0.3, 0.3, 0.3], [0.7, 0.6, 0.2]]), -1, finalize=True)
    for name in ("face_divergence", "edge_curl", "nodal_gradient"):
        geoms, parts = mesh._host_parts(name)
        M = mesh._host_matrix(name)
        for by_col in (False, True):
            pp = [(s_, sp.csr_matrix(P.T)) for s_, P in parts] if by_col else parts
            want = sp.csr_matrix(M.T) if by_col else M
            g2, p2 = fold_geometry(geoms, pp, by_col)
            assert g2 == [] and all(s_ == 0 for s_, _ in p2)
            total = sum(P for _, P in p2)
            assert_same_matrix(sp.csr_matrix(total), want)
            indptr, indices, codes, tmult, tsel = encode_parts(g2, p2, want.shape)
            assert not tsel.any() and tmult.size < 64
            got = sp.csr_matrix((tmult[codes], indices, indptr), shape=want.shape)
            assert_same_matrix(got, want)
    rng = np.random.default_rng(5)
    graded = TreeMesh([rng.random(64) + 0.5, rng.random(64) + 0.5, rng.random(8) + 0.5])
    graded.refine(-1, finalize=True)            # every base cell: 64 distinct widths along x and y
    geoms, parts = graded._host_parts("face_divergence")
    g2, p2 = fold_geometry(geoms, parts, False)
    assert g2 is geoms and p2 is parts


def test_point_location_by_keys_equals_the_descent():
    """The device kernel dzb_tree_locate replaces the reference's level-by-level descent (tree.cpp:758-766, 1499-1516) by two
    bisections per axis, a Morton key and a search in the sorted leaf keys.  The same algorithm in numpy against the
    descent (`TreeMesh._containing_cells`, itself checked against the reference fixture): random points, points exactly on
    nodes and cell centres, points outside, on cubic and non-cubic (several roots) trees."""
    from discretize_b200.octree import _spread

    rng = np.random.default_rng(33)
    meshes = []
    m = TreeMesh([16, 16, 16])
    m.refine_points(rng.random((12, 3)), -1, finalize=True)
    meshes.append(m)
    m = TreeMesh([rng.random(32) + 0.5, rng.random(8) + 0.5, rng.random(16) + 0.5], origin=[-2.0, 0.5, 3.0])
    m.refine_points(rng.random((30, 3)) * np.array([20.0, 5.0, 10.0]) + np.array([-2.0, 0.5, 3.0]), -1, finalize=True)
    meshes.append(m)
    for m in meshes:
        tr = m._tree
        lo = np.array([x[0] for x in m._xs])
        hi = np.array([x[-1] for x in m._xs])
        pts = lo + (rng.random((6000, 3)) * 1.4 - 0.2) * (hi - lo)
        for d in range(3):
            pts[:2000, d] = m._xs[d][rng.integers(0, m._xs[d].size, 2000)]
        want = m._containing_cells(pts)
        leaf_lo, _ = tr.leaves()
        leaf_keys = tr.keys(leaf_lo)
        assert (np.diff(leaf_keys) > 0).all()
        root_id = np.zeros(pts.shape[0], dtype=np.int64)
        code = np.zeros(pts.shape[0], dtype=np.int64)
        stride = 1
        for d in range(3):
            nodes = m._xs[d][::2]
            root = np.searchsorted(nodes[tr.rw * np.arange(1, tr.roots[d])], pts[:, d], side="right")
            fine = np.empty(pts.shape[0], dtype=np.int64)
            for r in range(tr.roots[d]):
                sel = root == r
                fine[sel] = np.searchsorted(nodes[tr.rw * r + 1:tr.rw * (r + 1)], pts[sel, d], side="left")
            root_id += stride * root
            stride *= tr.roots[d]
            code |= _spread(fine, tr.nbits, 3) << d
        got = np.searchsorted(leaf_keys, root_id * np.int64(tr.rw) ** 3 + code, side="right") - 1
        assert np.array_equal(got, want)

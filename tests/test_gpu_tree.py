"""GPU parity tests for the OcTree operators (coded-CSR tree_spmv kernel) against the reference fixtures."""
import numpy as np
import pytest

from tests.golden.make_golden_tree import TREE_OPS
from tests.helpers import assert_close_vec, assert_same_matrix, golden_matrix, load_golden
from tests.test_tree_cpu import _mesh_from_fixture

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def gold():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    return load_golden("tree_small.npz")


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("name", TREE_OPS)
def test_tree_operator_apply(gold, tag, name):
    m = _mesh_from_fixture(gold, tag, "refine")
    ref = golden_matrix(gold, f"{tag}_{name}")
    op = getattr(m, name)
    assert op.shape == ref.shape
    rng = np.random.default_rng(7)
    x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
    assert_close_vec(op @ x, ref @ x)
    assert_close_vec(op.T @ y, ref.T @ y)
    assert_same_matrix(op.tocsr(), ref)
    xd = torch.from_numpy(x).cuda()
    assert_close_vec((op @ xd).cpu().numpy(), ref @ x)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_tree_mass_matrices(gold, tag):
    m = _mesh_from_fixture(gold, tag, "state")
    sig, sig3 = gold[f"{tag}_sigma"], gold[f"{tag}_sigma3"]
    rng = np.random.default_rng(8)
    cases = [("Mf_iso", m.get_face_inner_product(sig)), ("Me_iso", m.get_edge_inner_product(sig)),
             ("Me_iso_inv", m.get_edge_inner_product(sig, invert_model=True, invert_matrix=True)),
             ("Mf_aniso", m.get_face_inner_product(sig3)), ("Me_aniso", m.get_edge_inner_product(sig3))]
    for key, op in cases:
        ref = golden_matrix(gold, f"{tag}_{key}")
        x = rng.standard_normal(ref.shape[1])
        assert_close_vec(op @ x, ref @ x)
        assert_same_matrix(op.tocsr(), ref)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_tree_tensor_mass_and_derivs(gold, tag):
    m = _mesh_from_fixture(gold, tag, "state")
    rng = np.random.default_rng(10)
    sig, sig3, sig6 = gold[f"{tag}_sigma"], gold[f"{tag}_sigma3"], gold[f"{tag}_sigma6"]
    vF, vE = gold[f"{tag}_vF"], gold[f"{tag}_vE"]
    cases = [
        ("Mf_tensor", m.get_face_inner_product(sig6)),
        ("Me_tensor_inv", m.get_edge_inner_product(sig6, invert_model=True)),
        ("dMf_iso", m.get_face_inner_product_deriv(sig)(vF)),
        ("dMe_iso_inv", m.get_edge_inner_product_deriv(sig, invert_model=True, invert_matrix=True)(vE)),
        ("dMe_aniso", m.get_edge_inner_product_deriv(sig3)(vE)),
        ("dMf_aniso_invmat", m.get_face_inner_product_deriv(sig3, invert_matrix=True)(vF)),
    ]
    for key, op in cases:
        ref = golden_matrix(gold, f"{tag}_{key}")
        assert op.shape == ref.shape, key
        x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
        assert_close_vec(op @ x, ref @ x)
        assert_close_vec(op.T @ y, ref.T @ y)
        a, b = op.tocsr(), ref.copy()
        a.eliminate_zeros(); b.eliminate_zeros()
        assert_same_matrix(a, b)
    with pytest.raises(Exception, match="Solver needed"):
        m.get_face_inner_product(sig6, invert_matrix=True)


def test_scipy_matrix_in_operator_algebra(gold):
    """A scipy.sparse projection mixed into the algebra becomes a device CSR operator (dzb_csr_spmv)."""
    import scipy.sparse as sp

    m = _mesh_from_fixture(gold, "a", "state")
    rng = np.random.default_rng(11)
    keep = np.sort(rng.choice(m.nN, m.nN // 2, replace=False))
    P = sp.csr_matrix((np.ones(keep.size), (keep, np.arange(keep.size))), shape=(m.nN, keep.size))
    G = m.nodal_gradient
    A = P.T @ (G.T @ m.get_edge_inner_product(gold["a_sigma"]) @ G) @ P
    Gr = golden_matrix(gold, "a_nodal_gradient")
    Ar = P.T @ Gr.T @ golden_matrix(gold, "a_Me_iso") @ Gr @ P
    x = rng.standard_normal(keep.size)
    assert_close_vec(A @ x, Ar @ x)
    assert_close_vec((G @ P) @ x, Gr @ (P @ x))


def test_tree_composite_and_aliases(gold):
    m = _mesh_from_fixture(gold, "a", "refine")
    C, Mf, Me = m.edge_curl, m.get_face_inner_product(gold["a_sigma"]), m.get_edge_inner_product(gold["a_sigma"])
    A = C.T @ Mf @ C + Me
    Cr = golden_matrix(gold, "a_edge_curl")
    Ar = Cr.T @ golden_matrix(gold, "a_Mf_iso") @ Cr + golden_matrix(gold, "a_Me_iso")
    e = np.random.default_rng(9).standard_normal(m.nE)
    assert_close_vec(A @ e, Ar @ e)
    assert m.aveF2CC is m.average_face_to_cell and m.nhF == m.n_hanging_faces


def test_larger_tree_device_vs_host_matrix():
    """~60k-cell tree: the kernel against the host-assembled matrix of the same operator (every entry coded)."""
    from discretize_b200.tree_mesh import TreeMesh

    rng = np.random.default_rng(3)
    m = TreeMesh([64, 64, 64], diagonal_balance=True)
    m.refine_points(rng.random((40, 3)) * 0.6 + 0.2, -1, [2, 2, 2], finalize=True)
    assert m.nC > 10000
    for name in ("face_divergence", "edge_curl", "nodal_gradient", "average_face_to_cell", "average_edge_to_cell",
                 "average_node_to_cell"):
        op, H = getattr(m, name), m._host_matrix(name)
        x, y = rng.standard_normal(H.shape[1]), rng.standard_normal(H.shape[0])
        assert_close_vec(op @ x, H @ x)
        assert_close_vec(op.T @ y, H.T @ y)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_tree_operators_loopback(gold, world):
    """tree_shard on one GPU: every rank's local coded-CSR SpMV (CUDA) on [owned | ghost] vectors, the
    exchange replayed by device copies that follow the plans (the NCCL path: tools/check_tree_shard.py)."""
    from discretize_b200.tree_shard import OPERATOR_KINDS, ShardedTreeOperator, TreeShard

    m = _mesh_from_fixture(gold, "a", "refine")
    shards = [TreeShard(m, r, world) for r in range(world)]
    rng = np.random.default_rng(11)
    for name, (rk, ck) in OPERATOR_KINDS.items():
        for tr in (False, True):
            # the reference's matrix: from the fixture where it holds it, else the host assembly (itself pinned against
            # the reference's fixtures in tests/test_tree_cpu.py)
            ref = golden_matrix(gold, f"a_{name}") if f"a_{name}__data" in gold.files else m._host_matrix(name)
            ref = ref.T.tocsr() if tr else ref
            kin, kout = (rk, ck) if tr else (ck, rk)
            x = rng.standard_normal(ref.shape[1])
            ops = [ShardedTreeOperator(s, name, tr) for s in shards]
            for s, op in zip(shards, ops):
                op.buf.zero_()
                op.buf[:op.plan.n_owned] = torch.from_numpy(s.scatter(kin, x)).cuda()
            for p, op in enumerate(ops):
                for q in range(world):
                    idx = op.plan.send_idx[q]
                    if idx.size:
                        dst = ops[q]
                        lo = dst.plan.n_owned + int(dst.plan.ghost_off[p])
                        dst.buf[lo:lo + idx.size] = op.buf[torch.as_tensor(idx).cuda()]
            y = np.zeros(ref.shape[0])
            y2 = np.zeros(ref.shape[0])
            for s, op in zip(shards, ops):
                op.csr.apply(op.buf, op.out)
                y[s.owned(kout)] = op.out.cpu().numpy()
                if op.csr_early is not None:       # the overlap path: early pass without ghost reads + boundary rows
                    op.out.fill_(float("nan"))
                    op.csr_early.apply(op.buf, op.out)
                    op.csr_bnd.apply(op.buf, op.out_bnd)
                    op.out.index_copy_(0, op.bnd_rows, op.out_bnd)
                y2[s.owned(kout)] = op.out.cpu().numpy()
            assert_close_vec(y, ref @ x)
            assert_close_vec(y2, ref @ x)


@pytest.mark.parametrize("name", ["face_divergence", "average_edge_to_cell", "edge_curl", "nodal_gradient"])
def test_tree_spmv_layouts_agree(gold, name):
    """ELL (constant row length), sliced ELL (natural and column-sorted rows), warp-staged CSR and the lanes-per-row
    kernel are implementations of one SpMV."""
    from discretize_b200.tree_ops import CodedCsr

    m = _mesh_from_fixture(gold, "b", "refine")
    ref = golden_matrix(gold, f"b_{name}")
    rng = np.random.default_rng(13)
    x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
    op = getattr(m, name)
    for tr, vec, r in ((False, x, ref @ x), (True, y, ref.T @ y)):
        csr = op._csr(tr)
        layouts = ["stream", "lpr", "sell"] + (["ell"] if csr.row_len > 0 else []) + (["hyb"] if csr.hyb_K > 0 else [])
        if name == "nodal_gradient" and not tr:
            assert csr.layout == "hyb"               # ragged short rows: ELL of the common width + the overflow rows
        if not tr and name in ("face_divergence", "average_edge_to_cell"):
            assert "ell" in layouts       # every cell-rowed operator has rows of one length
        for layout in layouts:
            csr.use_layout(layout)
            for lpr in ((1, 2, 8) if layout == "lpr" else (0,)):
                csr.lpr = lpr
                out = torch.empty(csr.shape[0], dtype=torch.float64, device="cuda")
                csr.apply(torch.from_numpy(vec).cuda(), out)
                assert_close_vec(out.cpu().numpy(), r)
        # sliced ELL with its rows stored in column-locality order (result scattered through the permutation), with the
        # geometry kept in arrays (fold=False) as on meshes whose products do not fit a code table
        geoms, parts = m._host_parts(name)
        if tr:
            parts = [(s_, P.T.tocsr()) for s_, P in parts]
        for fold in (True, False):
            c2 = CodedCsr(geoms, parts, csr.shape, geom_by_col=tr, sort_rows=True, fold=fold)
            c2.use_layout("sell")
            assert c2.ell_perm is not None
            out = torch.full((c2.shape[0],), float("nan"), dtype=torch.float64, device="cuda")
            c2.apply(torch.from_numpy(vec).cuda(), out)
            assert_close_vec(out.cpu().numpy(), r)
            # hybrid layout (ELL of the common width + overflow rows), geometry by row / by column kept in arrays too
            c3 = CodedCsr(geoms, parts, csr.shape, geom_by_col=tr, fold=fold)
            if c3.hyb_K > 0:
                c3.use_layout("hyb")
                out = torch.full((c3.shape[0],), float("nan"), dtype=torch.float64, device="cuda")
                c3.apply(torch.from_numpy(vec).cuda(), out)
                assert_close_vec(out.cpu().numpy(), r)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_tree_node_and_edge_to_face_averages(gold, tag):
    from tests.golden.make_golden_tree_n2 import TREE_N2_OPS

    g2 = load_golden("tree_n2_small.npz")
    m = _mesh_from_fixture(gold, tag, "refine")
    rng = np.random.default_rng(14)
    for name in TREE_N2_OPS:
        ref = golden_matrix(g2, f"{tag}_{name}")
        op = getattr(m, name)
        x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
        assert_close_vec(op @ x, ref @ x)
        assert_close_vec(op.T @ y, ref.T @ y)
        assert_same_matrix(op.tocsr(), ref)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_tree_cell_centred_operators(gold, tag):
    """OcTree cell-centred side on the GPU: the host-assembled operators through K12 and cell_gradient(_x/_y/_z) =
    -Pa Mf^-1 D^T diag(V) composed lazily from CUDA operators (tree_mesh.py:860-958)."""
    from tests.golden.make_golden_tree_n2 import TREE_CC_COMPOSITES, TREE_CC_METHODS, TREE_CC_OPS

    g2 = load_golden("tree_n2_small.npz")
    m = _mesh_from_fixture(gold, tag, "refine")
    rng = np.random.default_rng(15)
    for name in TREE_CC_OPS + TREE_CC_METHODS + TREE_CC_COMPOSITES:
        ref = golden_matrix(g2, f"{tag}_{name}")
        op = getattr(m, name)
        op = op() if name in TREE_CC_METHODS else op
        assert op.shape == ref.shape, name
        x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
        assert_close_vec(op @ x, ref @ x)
        assert_close_vec(op.T @ y, ref.T @ y)
        a, b = op.tocsr(), ref.copy()
        if name in TREE_CC_COMPOSITES:      # products: compare without stored zeros (scipy prunes differently per route)
            a.eliminate_zeros(); b.eliminate_zeros()
        assert_same_matrix(a, b)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_tree_node_interpolation(gold, tag):
    g2 = load_golden("tree_n2_small.npz")
    m = _mesh_from_fixture(gold, tag, "refine")
    pts = g2[f"{tag}_interp_pts"]
    rng = np.random.default_rng(16)
    for zo, key in ((False, "interp_N"), (True, "interp_N_zeros_outside")):
        ref = golden_matrix(g2, f"{tag}_{key}")
        Q = m.get_interpolation_matrix(pts, "N", zeros_outside=zo)
        x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
        assert_close_vec(Q @ x, ref @ x)
        assert_close_vec(Q.T @ y, ref.T @ y)
        assert_same_matrix(Q.tocsr(), ref)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_tree_tensor_model_derivative(gold, tag):
    g2 = load_golden("tree_n2_small.npz")
    m = _mesh_from_fixture(gold, tag, "state")
    rng = np.random.default_rng(18)
    sig6 = gold[f"{tag}_sigma6"]
    for get, key, v in ((m.get_face_inner_product_deriv, "dMf_tensor", gold[f"{tag}_vF"]),
                        (m.get_edge_inner_product_deriv, "dMe_tensor", gold[f"{tag}_vE"])):
        ref = golden_matrix(g2, f"{tag}_{key}")
        op = get(sig6)(v)
        x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
        assert_close_vec(op @ x, ref @ x)
        assert_close_vec(op.T @ y, ref.T @ y)
    with pytest.raises(NotImplementedError):
        m.get_face_inner_product_deriv(sig6, invert_model=True)


@pytest.mark.parametrize("loc", ["Ex", "Fy", "CC"])
def test_tree_edge_face_cell_interpolation(loc):
    """Edge / face / cell-centre interpolation (tree_interp.py) applied on the GPU against the reference's matrix
    (tests/golden/tree_interp_small.npz, bit-exact on the host in test_tree_interp_cpu.py)."""
    import os
    import sys

    import scipy.sparse as sp

    import discretize_b200 as dz

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import boundary_cases as bc

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tree_interp_small.npz"))
    tag, h, origin = bc.tree_specs()[1]
    m = bc.build_tree(dz.TreeMesh, h, origin)
    ref = sp.csr_matrix((g[f"{tag}|{loc}|data"], g[f"{tag}|{loc}|indices"], g[f"{tag}|{loc}|indptr"]),
                        shape=tuple(g[f"{tag}|{loc}|shape"]))
    Q = m.get_interpolation_matrix(g[f"{tag}|points"], loc)
    rng = np.random.default_rng(19)
    x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
    assert_close_vec(Q @ x, ref @ x)
    assert_close_vec(Q.T @ y, ref.T @ y)


def test_tree_operators_with_many_distinct_weights_fall_back_to_stored_values():
    """average_cell_to_face carries distance weights that depend on the base widths: with random widths there are more
    distinct values than codes, and the operator must then apply the reference's matrix with stored values (ADVICE r1:
    it used to raise)."""
    from discretize_b200.tree_mesh import TreeMesh

    rng = np.random.default_rng(5)
    h = [rng.random(16) + 0.5 for _ in range(3)]
    mesh = TreeMesh(h)
    mesh.refine_points(np.array([[4.0, 5.0, 6.0], [9.0, 3.0, 7.0]]), -1, finalize=True)
    for name in ("average_cell_to_face", "average_cell_vector_to_face", "average_face_to_cell", "face_divergence"):
        op = getattr(mesh, name)
        M = op.tocsr()
        x, y = rng.standard_normal(M.shape[1]), rng.standard_normal(M.shape[0])
        assert_close_vec(op @ x, M @ x)
        assert_close_vec(op.T @ y, M.T @ y)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_device_point_location_matches_the_host_descent(gold, tag):
    """dzb_tree_locate (bisections + Morton key + leaf-key search) against the level-by-level descent that restates
    tree.cpp:758-766, 1499-1516: random points, points exactly on fine nodes and cell centres (the reference's ties: root
    boundaries go high, centres go low), points outside the mesh, NaN; also on a non-cubic mesh (several roots)."""
    import discretize_b200 as dz

    meshes = [_mesh_from_fixture(gold, tag, "refine")]
    rng = np.random.default_rng(21)
    m2 = dz.TreeMesh([rng.random(32) + 0.5, rng.random(16) + 0.5, rng.random(64) + 0.5], origin=[-3.0, 1.0, 0.5])
    m2.refine_points(rng.random((40, 3)) * np.array([20.0, 10.0, 40.0]) + np.array([-3.0, 1.0, 0.5]), -1, finalize=True)
    meshes.append(m2)
    for m in meshes:
        lo = np.array([x[0] for x in m._xs])
        hi = np.array([x[-1] for x in m._xs])
        pts = lo + rng.random((20000, 3)) * (hi - lo)
        k = 3000
        for d in range(3):                                   # coordinates exactly on fine nodes / fine cell centres
            pts[:k, d] = m._xs[d][rng.integers(0, m._xs[d].size, k)]
        pts[k:k + 500] = lo + (rng.random((500, 3)) * 3.0 - 1.0) * (hi - lo)      # inside and outside
        pts[k + 500:k + 510, 1] = np.nan
        m._device_locate_min_points = 1 << 62
        host = m._containing_cells(pts)
        dev = m.containing_cells_device(pts).cpu().numpy()
        assert np.array_equal(host, dev)
        m._device_locate_min_points = 1
        assert np.array_equal(m.point2index(pts), host)      # the public call takes the device path for large inputs

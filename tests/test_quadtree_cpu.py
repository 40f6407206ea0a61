"""QuadTree (2-D TreeMesh) against tests/golden/quadtree_small.npz, generated from the unmodified reference by
tests/golden/make_golden_quadtree.py.  One case list (tests/golden/quadtree_cases.py) is run on both libraries.  All host
assembly -> no GPU needed; the GPU apply is the generic ``dzb_csr_spmv`` path (test_apply_on_gpu below, tests/test_gpu_csr).

Bar: cells, counts and index arrays identical; geometry bit-identical; operator / mass values within 1e-13 relative;
interpolation weights bit-identical; same exception types."""
import os
import sys
import warnings

import numpy as np
import pytest

import discretize_b200 as dz

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import _pack  # noqa: E402
import quadtree_cases as qc  # noqa: E402

GOLD = _pack.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quadtree_small.npz"))
SPECS = {s[0]: s for s in qc.specs()}
EXACT = ("count:", "array:", "levels", "point2index", "interp:", "index:")


def build(tag, **kw):
    _, h, origin, kind = SPECS[tag]
    return qc.build(dz, h, origin, kind, **kw)


@pytest.mark.parametrize("tag", list(SPECS))
def test_quadtree_matches_reference_fixture(tag):
    mesh = build(tag)
    assert type(mesh).__name__ == "QuadTreeMesh" and isinstance(mesh, dz.TreeMesh) and mesh.dim == 2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ours = qc.evaluate(mesh, GOLD[f"{tag}|points"], host_matrix=mesh._host_matrix)
    keys = [k for k in GOLD.files if k.startswith(tag + "|") and not k.endswith("|points")]
    assert sorted(k.split("|", 1)[1] for k in keys) == sorted(ours)
    for k in keys:
        ref, got = GOLD[k], ours[k.split("|", 1)[1]]
        name = k.split("|", 1)[1]
        if ref.dtype.kind in "US":
            assert got.dtype.kind in "US" and str(got) == str(ref), (k, str(ref), str(got)[:80])
            continue
        assert got.dtype.kind not in "US", (k, str(got))
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        if name.startswith(EXACT) or ref.dtype.kind in "ib":
            assert np.array_equal(got, ref), k
        elif ref.size:
            scale = max(1.0, float(np.abs(ref).max()))
            assert float(np.abs(got - ref).max()) <= 1e-13 * scale, (k, float(np.abs(got - ref).max()))


def test_diagonal_balance_gives_the_reference_tree():
    mesh = build("stretched_ball", diagonal_balance=True)
    assert np.array_equal(mesh.cell_centers, GOLD["diag|cell_centers"])
    assert np.array_equal(mesh._cell_levels_by_indexes(), GOLD["diag|levels"])
    assert mesh.n_cells != build("stretched_ball").n_cells


def test_mimetic_identities_and_constants():
    uniform = build("uniform_ball")      # curl grad = 0 needs hanging nodes at the midpoints of their parents
    C, G = uniform._host_matrix("edge_curl"), uniform._host_matrix("nodal_gradient")
    assert uniform.n_hanging_nodes > 0 and abs(C @ G).max() <= 1e-12 * abs(C).max() * abs(G).max()
    mesh = build("stretched_box_points")
    D = mesh._host_matrix("face_divergence")
    for name, n_in in (("average_face_to_cell", mesh.n_faces), ("average_edge_to_cell", mesh.n_edges),
                       ("average_node_to_cell", mesh.n_nodes), ("average_node_to_edge", mesh.n_nodes)):
        A = mesh._host_matrix(name)
        assert np.abs(A @ np.ones(n_in) - 1.0).max() <= 1e-14
    # the divergence of a constant vector field vanishes: u = (1, 0) on faces [x-faces | y-faces]
    u = np.r_[np.ones(mesh.n_faces_x), np.zeros(mesh.n_faces_y)]
    assert np.abs(D @ u).max() <= 1e-12 * abs(D).max()


def test_unfinalized_and_unbuilt_pieces_raise():
    m = dz.TreeMesh([8, 8])
    with pytest.raises(dz.TreeMeshNotFinalizedError):
        m.n_cells
    m.refine(1)
    with pytest.raises(NotImplementedError):
        m.cell_gradient_z
    with pytest.raises(NotImplementedError, match="Unable to interpolate from Z edges/faces in 2D"):
        m.get_interpolation_matrix(np.zeros((1, 2)), "Ez")
    with pytest.raises(ValueError):
        dz.TreeMesh([8, 6])          # not a power of two


@pytest.mark.gpu
def test_apply_on_gpu():
    tag = "stretched_box_points"
    mesh = build(tag)
    rng = np.random.default_rng(1)
    for name in ("face_divergence", "edge_curl", "nodal_gradient", "average_edge_to_cell_vector"):
        ref = GOLD[f"{tag}|op:{name}"]
        op = getattr(mesh, name)
        x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
        assert np.abs(op @ x - ref @ x).max() <= 1e-12 * max(1.0, np.abs(ref @ x).max())
        assert np.abs(op.T @ y - ref.T @ y).max() <= 1e-12 * max(1.0, np.abs(ref.T @ y).max())
    M = mesh.get_edge_inner_product(GOLD[f"{tag}|points"][:0].sum() + 2.5)
    ref = GOLD[f"{tag}|mass:edge:scalar:00"]
    x = rng.standard_normal(ref.shape[1])
    assert np.abs(M @ x - ref @ x).max() <= 1e-12 * np.abs(ref @ x).max()
    Q = mesh.get_interpolation_matrix(GOLD[f"{tag}|points"], "Ex")
    ref = GOLD[f"{tag}|interp:Ex"]
    x = rng.standard_normal(ref.shape[1])
    assert np.abs(Q @ x - ref @ x).max() <= 1e-12 * max(1.0, np.abs(ref @ x).max())

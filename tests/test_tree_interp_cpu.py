"""TreeMesh.get_interpolation_matrix (nodes / edges / faces / cell centres) against tests/golden/tree_interp_small.npz,
generated from the unmodified reference by tests/golden/make_golden_tree_interp.py.  Host assembly -> no GPU needed; the
GPU apply is the generic ``dzb_csr_spmv`` path (tests/test_gpu_csr.py, test_gpu_tree.py::node interpolation).

Bar: identical sparsity pattern (canonical CSR, explicit zeros included) and bit-identical weights."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import discretize_b200 as dz

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import boundary_cases as bc  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tree_interp_small.npz"))
LOCS = ["N", "Ex", "Ey", "Ez", "Fx", "Fy", "Fz", "CC"]
_MESHES = {}


def mesh_for(tag):
    if tag not in _MESHES:
        spec = {s[0]: s for s in bc.tree_specs()}[tag]
        _MESHES[tag] = bc.build_tree(dz.TreeMesh, spec[1], spec[2])
    return _MESHES[tag]


def canon(a):
    a = sp.csr_matrix(a.tocsr()).copy()
    a.sum_duplicates()
    a.sort_indices()
    return a


@pytest.mark.parametrize("loc", LOCS)
@pytest.mark.parametrize("tag", [s[0] for s in bc.tree_specs()])
def test_matches_reference_bit_for_bit(tag, loc):
    m = mesh_for(tag)
    A = canon(m.get_interpolation_matrix(GOLD[f"{tag}|points"], loc))
    assert tuple(A.shape) == tuple(GOLD[f"{tag}|{loc}|shape"])
    assert np.array_equal(A.indptr, GOLD[f"{tag}|{loc}|indptr"])
    assert np.array_equal(A.indices, GOLD[f"{tag}|{loc}|indices"])
    assert np.array_equal(A.data, GOLD[f"{tag}|{loc}|data"])


@pytest.mark.parametrize("loc", ["Ex", "Fz", "CC", "N"])
def test_zeros_outside(loc):
    """Outside points get a zero row; inside points keep their weights (the reference leaves those rows uninitialised
    for edges / faces / cells, tree_ext.pyx:5816-5835, so only its node variant can be compared: see the golden N case)."""
    tag = "o_stretched"
    m = mesh_for(tag)
    pts, inside = GOLD[f"{tag}|points"], GOLD[f"{tag}|inside"]
    assert np.array_equal(m.is_inside(pts), inside) and (~inside).sum() > 50
    A = canon(m.get_interpolation_matrix(pts, loc))
    Z = canon(m.get_interpolation_matrix(pts, loc, zeros_outside=True))
    assert abs(Z[~inside]).sum() == 0.0
    assert abs(A[inside] - Z[inside]).max() == 0.0


def test_partition_of_unity_and_exact_for_linear_fields():
    m = mesh_for("o_uniform")        # hanging nodes sit at the midpoints of their parents only on uniform widths
    pts = GOLD["o_uniform|points"][GOLD["o_uniform|inside"]]
    for loc, grid in (("CC", m.cell_centers), ("Ex", m.edges_x), ("Fy", m.faces_y), ("N", m.nodes)):
        A = m.get_interpolation_matrix(pts, loc).tocsr()
        off = {"CC": 0, "N": 0, "Ex": 0, "Fy": m.n_faces_x}[loc]
        assert np.abs(np.asarray(A.sum(axis=1)).ravel() - 1.0).max() < 1e-14
        assert A.min() >= 0.0
        if loc == "N":      # trilinear on the containing cell: reproduces any linear field
            f = 2.0 * grid[:, 0] - 3.0 * grid[:, 1] + 0.5 * grid[:, 2] + 1.0
            want = 2.0 * pts[:, 0] - 3.0 * pts[:, 1] + 0.5 * pts[:, 2] + 1.0
            assert np.abs(A[:, off:off + grid.shape[0]] @ f - want).max() < 1e-12


def test_the_row_depends_on_the_points_before_it():
    """Cell-centre interpolation carries its axis order from point to point (tree_ext.pyx:6240-6262): reversing the
    point list changes some rows.  The vectorised prefix composition must agree with a plain sequential replay."""
    from discretize_b200 import tree_interp

    m = mesh_for("o_stretched")
    pts = GOLD["o_stretched|points"]
    A = canon(m.get_interpolation_matrix(pts, "CC"))
    Ar = canon(m.get_interpolation_matrix(pts[::-1].copy(), "CC"))[::-1]
    assert abs(A - Ar).max() > 0
    rng = np.random.default_rng(0)
    maps = rng.integers(0, 6, size=(257, 6))
    s, seq = 0, []
    for row in maps:
        s = row[s]
        seq.append(s)
    assert np.array_equal(tree_interp._prefix_states(maps), np.array(seq))


def test_routing_errors():
    m = mesh_for("o_uniform")
    with pytest.raises(ValueError, match="Location must be a grid location, not edges"):
        m.get_interpolation_matrix(np.zeros((1, 3)), "E")
    with pytest.raises(TypeError, match="locType"):
        m.get_interpolation_matrix(np.zeros((1, 3)), locType="CC")
    with pytest.raises(ValueError):
        m.get_interpolation_matrix(np.zeros((1, 3)), "nonsense")

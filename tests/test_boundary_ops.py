"""Weak-form boundary pieces, boundary integrals / projections, surface and line inner products and the small geometry
conveniences against tests/golden/boundary_small.npz (generated from the unmodified reference by
tests/golden/make_golden_boundary.py).  All host assembly -> runs without a GPU; the GPU apply of these operators is
the generic ``dzb_csr_spmv`` path covered by tests/test_gpu_csr.py.

Bar: identical shapes, identical exception types, values within 1e-13 relative (index arrays exactly)."""
import os
import sys
import warnings

import numpy as np
import pytest

import discretize_b200 as dz

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import _pack  # noqa: E402
import boundary_cases as bc  # noqa: E402

GOLD = _pack.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "boundary_small.npz"))
RTOL = 1e-13


def compare(tag, ours):
    keys = [k for k in GOLD.files if k.startswith(tag + "|") and not k.endswith("|n_cells")]
    assert sorted(k.split("|", 1)[1] for k in keys) == sorted(ours), "case lists differ"
    for k in keys:
        ref, got = GOLD[k], ours[k.split("|", 1)[1]]
        if ref.dtype.kind in "US":
            assert got.dtype.kind in "US" and str(got) == str(ref), (k, str(ref), got)
            continue
        assert got.dtype.kind not in "US", (k, str(got))
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        if ref.dtype.kind in "ib":
            assert np.array_equal(got, ref), k
        elif ref.size:
            scale = max(1.0, float(np.abs(ref).max()))
            assert float(np.abs(got - ref).max()) <= RTOL * scale, (k, float(np.abs(got - ref).max()))


@pytest.mark.parametrize("spec", bc.tensor_specs(), ids=lambda s: s[0])
def test_tensor_mesh_boundary_and_surface(spec):
    tag, h, origin = spec
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        compare(tag, bc.evaluate(dz.TensorMesh(h, origin=origin)))


@pytest.mark.parametrize("spec", bc.tree_specs(), ids=lambda s: s[0])
def test_tree_mesh_boundary_and_surface(spec):
    tag, h, origin = spec
    mesh = bc.build_tree(dz.TreeMesh, h, origin)
    assert mesh.n_cells == int(GOLD[tag + "|n_cells"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        compare(tag, bc.evaluate(mesh, tree=True))


def test_removed_names_raise_like_the_reference():
    m = dz.TensorMesh([3, 2, 2])
    with pytest.raises(NotImplementedError, match="TensorMesh.vol has been removed, please use TensorMesh.cell_volumes."):
        m.vol
    with pytest.raises(NotImplementedError, match=r"The nCx property is removed, please access as mesh.shape_cells\[0\]"):
        m.nCx
    t = dz.TreeMesh([4, 4, 4])
    t.refine(1)
    with pytest.raises(NotImplementedError, match="TreeMesh.maxLevel has been removed, please use TreeMesh.max_used_level."):
        t.maxLevel
    with pytest.raises(AttributeError):
        m.not_a_mesh_attribute


def test_boundary_rows_match_the_full_host_matrix():
    """The O(boundary) assembly of P_bf @ A equals selecting the rows of the full host matrix, in every dimension."""
    rng = np.random.default_rng(2)
    for h in ([rng.random(4) + .5, rng.random(3) + .5, rng.random(5) + .5], [rng.random(5) + .5, rng.random(4) + .5],
              [rng.random(6) + .5]):
        m = dz.TensorMesh(h)
        bf = m._boundary_indices()[0]
        for name in ("average_cell_to_face", "average_node_to_face", "average_edge_to_face"):
            full = m._host_matrix(name)[bf]
            assert abs(m._boundary_face_rows(name) - full).max() == 0.0

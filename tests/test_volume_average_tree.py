"""utils.volume_average with a TreeMesh on either side against tests/golden/volume_average_tree_small.npz (generated
from the unmodified reference by tests/golden/make_golden_volume_average_tree.py).

Bar: identical sparsity pattern (zero-weight pairs dropped, as the reference drops them), weights within 1e-13."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import discretize_b200 as dz

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import volume_average_tree_cases as cases  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "volume_average_tree_small.npz"))
TAGS = sorted({k.split("|")[0] for k in GOLD.files})
_PAIRS = {}


def pair(tag):
    if not _PAIRS:
        _PAIRS.update(cases.build(dz))
    return _PAIRS[tag]


def canon(m):
    m = sp.csr_matrix(m).copy()
    m.sum_duplicates()
    m.eliminate_zeros()
    m.sort_indices()
    return m


def test_case_list_matches_fixture():
    assert TAGS == sorted(cases.build(dz))


@pytest.mark.parametrize("tag", TAGS)
def test_matrix_matches_reference_fixture(tag):
    mesh_in, mesh_out = pair(tag)
    A = canon(dz.utils.volume_average(mesh_in, mesh_out).tocsr())
    assert tuple(A.shape) == tuple(GOLD[f"{tag}|shape"]) == (mesh_out.n_cells, mesh_in.n_cells)
    assert np.array_equal(A.indptr, GOLD[f"{tag}|indptr"])
    assert np.array_equal(A.indices, GOLD[f"{tag}|indices"])
    assert np.abs(A.data - GOLD[f"{tag}|data"]).max() <= 1e-13
    # the matrix reproduces the reference's direct application, and averages: rows sum to one
    assert np.abs(A @ GOLD[f"{tag}|values"] - GOLD[f"{tag}|averaged"]).max() <= 1e-13
    assert np.abs(np.asarray(A.sum(axis=1)).ravel() - 1.0).max() <= 1e-13


def test_nested_meshes_conserve_the_integral():
    """Same tensor base, tree -> coarser / finer tree: the volume integral of the model is preserved."""
    mesh_in, mesh_out = pair("tree_tree_same_base")
    A = dz.utils.volume_average(mesh_in, mesh_out).tocsr()
    v = np.random.default_rng(3).random(mesh_in.n_cells)
    assert abs(mesh_out.cell_volumes @ (A @ v) - mesh_in.cell_volumes @ v) < 1e-12 * abs(mesh_in.cell_volumes @ v)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["tree_tree_general", "tensor_tree_general", "tree_tensor_general"])
def test_apply_on_gpu(tag):
    mesh_in, mesh_out = pair(tag)
    got = dz.utils.volume_average(mesh_in, mesh_out, GOLD[f"{tag}|values"])
    assert np.abs(got - GOLD[f"{tag}|averaged"]).max() <= 1e-12
    out = np.zeros(mesh_out.n_cells)
    assert dz.utils.volume_average(mesh_in, mesh_out, GOLD[f"{tag}|values"], out) is out
    assert np.abs(out - GOLD[f"{tag}|averaged"]).max() <= 1e-12

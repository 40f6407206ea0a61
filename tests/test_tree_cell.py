"""``TreeMesh.refine(function(cell))`` and ``TreeCell`` (mesh[i]) against tests/golden/tree_cell_small.npz, generated from
the unmodified reference by tests/golden/make_golden_tree_cell.py.  The reference asks the function in one depth-first
pass while balancing as it goes (tree.cpp:398-419, 459-746), so the resulting tree -- and even the number of calls --
depends on that order; both are reproduced."""
import os
import sys
import warnings

import numpy as np
import pytest

import discretize_b200 as dz

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import tree_cell_cases as tc  # noqa: E402


def test_refine_function_and_tree_cells_match_reference():
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tree_cell_small.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ours = tc.evaluate(dz)
    assert sorted(ours) == sorted(gold.files)
    for k in gold.files:
        assert np.asarray(ours[k]).shape == gold[k].shape, k
        assert np.array_equal(np.asarray(ours[k]), gold[k]), k


def test_vectorized_extension_and_errors():
    m = dz.TreeMesh([16, 16])
    m.refine(lambda cc: np.where(np.linalg.norm(cc - 0.5, axis=1) < 0.45, 4, 1), vectorized=True)
    assert m.n_cells > 16 and m.finalized
    assert m[0].index == 0 and m[-1].index == m.n_cells - 1 and len(list(m)) == m.n_cells
    with pytest.raises(IndexError):
        m[m.n_cells]
    with pytest.raises(TypeError):
        m["a"]
    seen = []
    t = dz.TreeMesh([4, 4, 4])
    t.refine(lambda cell: seen.append(cell.index) or 1)
    assert set(seen) == {-1}                      # cells are not numbered while the tree is being built
    with pytest.raises(dz.TreeMeshNotFinalizedError):
        dz.tree_mesh.TreeCell(t, [0, 0, 0], 2, 1).nodes

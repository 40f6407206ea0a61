"""TensorCell / TensorMesh indexing against tests/golden/tensor_cell_small.npz (generated from the unmodified reference by
tests/golden/make_golden_tensor_cell.py): every cell of a 1-D, a 2-D and a 3-D mesh, tuple / slice / negative indexing."""
import os
import sys

import numpy as np
import pytest

import discretize_b200 as dz

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import tensor_cell_cases as tc  # noqa: E402


def test_tensor_cells_match_reference_fixture():
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tensor_cell_small.npz"))
    ours = tc.evaluate(dz)
    assert sorted(ours) == sorted(gold.files)
    for k in gold.files:
        ref, got = gold[k], np.asarray(ours[k])
        if ref.dtype.kind in "US":
            assert str(got) == str(ref), (k, str(ref), str(got))
        else:
            assert got.shape == ref.shape and np.array_equal(got, ref), k


def test_cell_equality():
    m = dz.TensorMesh([3, 2])
    assert m[0] == m[0] and not (m[0] == m[1])
    with pytest.raises(TypeError, match="Cannot compare"):
        m[0] == 3

"""utils matrix / index / coordinate helpers and the symbolic Zero / Identity against
tests/golden/matrix_utils_small.npz (generated from the unmodified reference by
tests/golden/make_golden_matrix_utils.py; one recipe runs on both libraries).  Values within 1e-13, same exception types."""
import os
import sys
import warnings

import numpy as np

import discretize_b200 as dz

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import matrix_utils_cases as mc  # noqa: E402


def test_matrix_utils_match_reference_fixture():
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matrix_utils_small.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ours = mc.evaluate(dz.utils, dz)
    assert sorted(ours) == sorted(gold.files)
    for k in gold.files:
        ref, got = gold[k], np.asarray(ours[k])
        if ref.dtype.kind in "US":
            assert got.dtype.kind in "US" and str(got) == str(ref), (k, str(ref), str(got)[:80])
            continue
        assert got.dtype.kind not in "US", (k, str(got))
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        if ref.dtype.kind in "ib":
            assert np.array_equal(got, ref), k
        elif ref.size:
            assert np.abs(got - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), (k, np.abs(got - ref).max())


def test_removed_function_names_raise_with_a_pointer():
    import pytest

    with pytest.raises(NotImplementedError, match="sdInv has been removed, please use sdinv."):
        dz.utils.sdInv(1)
    with pytest.raises(NotImplementedError, match="meshTensor has been removed, please use unpack_widths."):
        dz.utils.meshTensor([(1.0, 2)])

"""utils.mesh_builder_xyz / refine_tree_xyz / active_from_xyz / extract_core_mesh and the small array helpers against
tests/golden/mesh_utils_small.npz (generated from the unmodified reference by tests/golden/make_golden_mesh_utils.py; one
case list runs on both libraries).  Bar: identical widths, origins, cells and masks."""
import os
import sys

import numpy as np
import pytest

import discretize_b200 as dz

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import mesh_utils_cases as mc  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_utils_small.npz"))


def test_mesh_utils_match_reference_fixture():
    ours = mc.evaluate(dz.utils)
    assert sorted(ours) == sorted(GOLD.files)
    for k in GOLD.files:
        assert np.asarray(ours[k]).shape == GOLD[k].shape, k
        assert np.array_equal(np.asarray(ours[k]), GOLD[k]), k


def test_errors():
    with pytest.raises(ValueError, match="Revise mesh_type"):
        dz.utils.mesh_builder_xyz(np.zeros((3, 2)), [1.0, 1.0], mesh_type="curvilinear")
    m = dz.TreeMesh([8, 8])
    with pytest.raises(NotImplementedError):
        dz.utils.refine_tree_xyz(m, np.zeros((2, 2)), method="spiral")
    with pytest.raises(ValueError, match="octree_levels_padding"):
        dz.utils.refine_tree_xyz(m, np.zeros((2, 2)), octree_levels=[1, 1], octree_levels_padding=[1])
    with pytest.raises(ValueError, match="grid_reference"):
        dz.utils.active_from_xyz(dz.TensorMesh([4, 4]), np.zeros((2, 2)), grid_reference="F")
    with pytest.raises(Exception, match="Only implemented for class TensorMesh"):
        dz.utils.extract_core_mesh(np.zeros((2, 2)), m)

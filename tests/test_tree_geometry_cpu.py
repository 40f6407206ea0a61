"""Geometric refinement and cell queries of TreeMesh (2-D and 3-D) against tests/golden/tree_geometry_small.npz, generated
from the unmodified reference by tests/golden/make_golden_tree_geometry.py.  Bar: the same cells in the same order (centres
bit-identical, levels equal) and the same index sets from every query."""
import os
import sys

import numpy as np
import pytest

import discretize_b200 as dz

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import tree_geometry_cases as gc  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tree_geometry_small.npz"))
CASES = gc.cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_refinement_and_queries_match_reference(case):
    ours = gc.evaluate(dz, case)
    keys = [k for k in GOLD.files if k.startswith(case[0] + "|")]
    assert sorted(k.split("|", 1)[1] for k in keys) == sorted(ours)
    for k in keys:
        ref, got = GOLD[k], ours[k.split("|", 1)[1]]
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        assert np.array_equal(got, ref), k


IMAGE_CASES = gc.image_cases()


@pytest.mark.parametrize("case", IMAGE_CASES, ids=[c[0] for c in IMAGE_CASES])
def test_refine_image_boundary_cells_and_tensors_match_reference(case):
    ours = gc.evaluate_image(dz, case)
    keys = [k for k in GOLD.files if k.startswith(case[0] + "|")]
    assert sorted(k.split("|", 1)[1] for k in keys) == sorted(ours)
    for k in keys:
        ref, got = GOLD[k], ours[k.split("|", 1)[1]]
        assert got.shape == ref.shape and np.array_equal(got, ref), k


def test_argument_checks():
    m = dz.TreeMesh([8, 8, 8])
    with pytest.raises(ValueError, match="do not broadcast"):
        m.refine_line(np.zeros((4, 3)), [1, 2])
    with pytest.raises(ValueError, match="triangle array must be"):
        m.refine_triangle(np.zeros((2, 4, 3)), 1)
    with pytest.raises(ValueError, match="All heights must be positive."):
        m.refine_vertical_trianglular_prism(np.zeros((1, 3, 3)), -1.0, 1)
    m.refine(1)
    with pytest.raises(RuntimeError):
        m.refine_line(np.zeros((2, 3)), 1)
    with pytest.raises(ValueError, match="A line segment has two points"):
        m.get_cells_on_line(np.zeros((3, 3)))
    with pytest.raises(ValueError, match="image array size"):
        dz.TreeMesh([8, 8]).refine_image(np.zeros(10))
    with pytest.raises(NotImplementedError, match="has been removed"):
        m.readUBC("x")
    with pytest.raises(NotImplementedError, match="Nodal Laplacian"):
        m.nodal_laplacian
    with pytest.raises(NotImplementedError):
        q = dz.TreeMesh([8, 8])
        q.refine(1)
        q.get_cells_in_vertical_trianglular_prism(np.zeros((3, 2)), 1.0)

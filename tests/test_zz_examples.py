"""examples/octree_dc_workflow.py: the mesh-building half runs everywhere, the solve on a GPU box.  (Named to run last.)"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
import octree_dc_workflow as example  # noqa: E402


def test_example_builds_its_mesh():
    mesh, electrodes, active = example.build_mesh()
    assert mesh.finalized and mesh.dim == 3 and mesh.n_hanging_nodes > 0
    assert active.dtype == bool and 0 < active.sum() < mesh.n_cells
    # every electrode sits in a finest-level cell below the surface
    cells = np.atleast_1d(mesh.point2index(electrodes))
    assert np.all(mesh._cell_levels_by_indexes(cells) == mesh.max_level) and active[cells].all()


@pytest.mark.gpu
def test_example_solves_on_the_gpu():
    mesh, residual, voltages, info = example.main(verbose=False)
    assert info == 0 and residual < 1e-6
    assert np.isfinite(voltages).all() and voltages[0] > voltages[-1]      # +1 A at the first, -1 A at the last electrode

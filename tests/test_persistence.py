"""to_dict / save / load_mesh / copy / equals / pickle (base/base_mesh.py:50-197, tree_mesh.py:1085-1118,
utils/io_utils.py:8-40).  tests/golden/saved_*.json were written by the UNMODIFIED reference's ``mesh.save()``
(tests/golden/make_golden_saved_meshes.py): they must load here into the same meshes this package builds itself, and
what this package saves must be the same JSON (apart from the module name)."""
import json
import os
import pickle
import sys

import numpy as np
import pytest

import discretize_b200 as dz

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import boundary_cases as bc  # noqa: E402
import quadtree_cases as qc  # noqa: E402


def own_meshes():
    _, h, origin = bc.tensor_specs()[0]
    yield "saved_tensor.json", dz.TensorMesh(h, origin=origin)
    _, h, origin = bc.tree_specs()[1]
    yield "saved_octree.json", bc.build_tree(dz.TreeMesh, h, origin)
    _, h, origin, kind = qc.specs()[2]
    yield "saved_quadtree.json", qc.build(dz, h, origin, kind)


@pytest.mark.parametrize("file_name,mesh", list(own_meshes()), ids=["tensor", "octree", "quadtree"])
def test_reference_files_load_and_round_trip(file_name, mesh, tmp_path):
    path = os.path.join(HERE, "golden", file_name)
    loaded = dz.utils.load_mesh(path)
    assert type(loaded) is type(mesh) and loaded.equals(mesh) and mesh.equals(loaded)
    assert np.array_equal(loaded.cell_centers, mesh.cell_centers)
    assert np.array_equal(loaded.cell_volumes, mesh.cell_volumes)
    # what we save is what the reference saved
    ours = json.load(open(mesh.save(str(tmp_path / "m.json"))))
    theirs = json.load(open(path))
    assert ours.pop("__class__") in (theirs.pop("__class__"), "QuadTreeMesh")
    ours.pop("__module__"), theirs.pop("__module__")
    assert ours == theirs
    # copy / pickle / deserialize
    for other in (mesh.copy(), pickle.loads(pickle.dumps(mesh)), type(mesh).deserialize(mesh.serialize())):
        assert other is not mesh and other.equals(mesh) and other.validate()
        assert np.array_equal(other.cell_centers, mesh.cell_centers)


def test_equals_and_validate():
    a = dz.TensorMesh([4, 4, 4])
    assert not a.equals(dz.TensorMesh([4, 4, 2])) and not a.equals(dz.TreeMesh([4, 4, 4]))
    t = dz.TreeMesh([8, 8, 8])
    assert not t.validate() and not t.equals(t)           # not finalized
    t.refine(1)
    u = dz.TreeMesh([8, 8, 8])
    u.refine(2)
    assert t.validate() and t.equals(t) and not t.equals(u)
    with pytest.raises(NotImplementedError):
        dz.TensorMesh([2, 2], orientation=[[0.0, 1.0], [1.0, 0.0]])
    with pytest.raises(ValueError, match="Coordinate system"):
        dz.TensorMesh([2, 2], reference_system="polar")
    assert dz.TensorMesh([2, 2], reference_system="cyl").reference_system == "cylindrical"

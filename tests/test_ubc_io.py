"""UBC-GIF mesh / model files (mixins/mesh_io.py).  tests/golden/ubc/* were written by the UNMODIFIED reference
(tests/golden/make_golden_ubc.py): this package must read them into the meshes / models it builds itself and write
byte-identical files."""
import os
import sys

import numpy as np
import pytest

import discretize_b200 as dz

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ubc_cases as uc  # noqa: E402

UBC = os.path.join(HERE, "golden", "ubc")
CASES = {**uc.tensor_cases(dz), **uc.tree_cases(dz)}


@pytest.mark.parametrize("tag", list(CASES))
def test_read_the_reference_files_and_write_them_back(tag, tmp_path):
    mesh = CASES[tag]
    model = uc.model_for(mesh)
    cls = dz.TensorMesh if tag.startswith("tensor") else dz.TreeMesh
    loaded = cls.read_UBC(tag + ".msh", directory=UBC)
    assert type(loaded) is type(mesh) and loaded.equals(mesh)
    got = loaded.read_model_UBC(os.path.join(UBC, tag + ".mod"))
    assert got.shape == model.shape and np.abs(got - model).max() < 1e-12
    mesh.write_UBC(str(tmp_path / "m.msh"), models={str(tmp_path / "m.mod"): model})
    assert open(tmp_path / "m.msh").read() == open(os.path.join(UBC, tag + ".msh")).read()
    assert open(tmp_path / "m.mod").read() == open(os.path.join(UBC, tag + ".mod")).read()


def test_star_notation_comments_and_errors(tmp_path):
    p = tmp_path / "star.msh"
    p.write_text("! a comment\n4 3 2\n0 0 10 ! top south-west corner\n2*5.0 2*10\n3*2.5\n2.0 4.0\n")
    m = dz.TensorMesh.read_UBC(str(p))
    assert np.array_equal(m.h[0], [5, 5, 10, 10]) and np.array_equal(m.h[2], [4, 2]) and m.origin[2] == 4.0
    (tmp_path / "bad.msh").write_text("4 3\n0 0\n1 1\n")
    with pytest.raises(Exception, match="File format not recognized"):
        dz.TensorMesh.read_UBC(str(tmp_path / "bad.msh"))
    t = dz.TreeMesh([np.r_[1.0, 2.0, 1.0, 1.0]] * 3)
    t.refine(1)
    with pytest.raises(Exception, match="variable cell widths"):
        t.write_UBC(str(tmp_path / "t.msh"))
    with pytest.raises(TypeError, match="models must be a dict"):
        CASES["tensor3"].write_UBC(str(tmp_path / "x.msh"), models=[1, 2])

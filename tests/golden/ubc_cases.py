"""Meshes and models for the UBC-GIF file tests; shared by make_golden_ubc.py (reference) and tests/test_ubc_io.py."""
import numpy as np


def tensor_cases(lib):
    return {
        "tensor3": lib.TensorMesh([np.r_[10., 10., 20., 40.], np.r_[5., 5., 5.], np.r_[1., 2., 4., 8., 8.]], origin=[100., -20., -50.]),
        "tensor2": lib.TensorMesh([np.r_[10., 10., 20., 40., 40.], np.r_[1., 2., 4., 4.]], origin=[-30., -11.]),
    }


def tree_cases(lib):
    out = {}
    for tag, hs, o, balls, d in (("octree", [np.full(8, 2.0)] * 3, np.array([10., 20., -30.]), [[4., 4., 4., 3.], [12., 10., 8., 2.5]], 3),
                                 ("quadtree", [np.full(32, 1.5), np.full(16, 1.5)], np.array([-5., 3.]), [[10., 8., 6.], [30., 12., 5.]], 2)):
        t = lib.TreeMesh(hs, origin=o)
        b = np.array(balls)
        t.refine_ball(b[:, :d] + o, b[:, d], [-1, -2])
        out[tag] = t
    return out


def model_for(mesh):
    return np.round(np.random.default_rng(mesh.n_cells).random(mesh.n_cells), 6)

"""Mesh pairings for volume averaging with a TreeMesh on either side; shared by the golden generator (reference) and
tests/test_volume_average_tree.py (discretize_b200).  ``build(lib)`` -> {tag: (mesh_in, mesh_out)}."""
import numpy as np


def _tree(lib, h, origin, balls, levels):
    m = lib.TreeMesh(h, origin=origin)
    b = np.array(balls)
    d = len(h)
    m.refine_ball(b[:, :d] + origin, b[:, d], levels, finalize=True)
    return m


def build(lib):
    rng = np.random.default_rng(21)
    h16 = [np.full(16, 0.25)] * 3
    o0 = np.zeros(3)
    hs = [rng.random(16) * .2 + .15, rng.random(8) * .3 + .3, rng.random(16) * .2 + .2]
    o1 = np.array([-0.3, 0.2, 0.1])
    b1 = [[1., 1., 1., .8], [3., 2.5, 2., .6]]
    b2 = [[2., 2., 2., 1.0], [.5, .5, 3., .5]]
    ht = [rng.random(11) * .3 + .3, rng.random(7) * .5 + .4, rng.random(9) * .4 + .3]
    ot = np.array([-0.5, -0.4, 0.3])
    far = np.array([5.5, -7.0, 0.1])

    def A():
        return _tree(lib, h16, o0, b1, [3, 4])

    def B():
        return _tree(lib, h16, o0, b2, [4, 3])

    def S():
        return _tree(lib, hs, o1, b1, [3, 4])

    # QuadTrees (2-D): the same pairings; drawn after the 3-D widths so those keep their values
    q16 = [np.full(16, 0.25)] * 2
    p0 = np.zeros(2)
    qs = [rng.random(16) * .2 + .15, rng.random(8) * .3 + .3]
    p1 = np.array([-0.3, 0.2])
    c1 = [[1., 1., .8], [3., 2.5, .6]]
    c2 = [[2., 2., 1.0], [.5, .5, .5]]
    qt = [rng.random(11) * .3 + .3, rng.random(7) * .5 + .4]
    pt = np.array([-0.5, -0.4])

    def QA():
        return _tree(lib, q16, p0, c1, [3, 4])

    def QB():
        return _tree(lib, q16, p0, c2, [4, 3])

    def QS():
        return _tree(lib, qs, p1, c1, [3, 4])

    quad = {
        "quad_quad_same_base": (QA(), QB()),
        "tensor_quad_same_base": (lib.TensorMesh(q16, origin=p0), QA()),
        "quad_tensor_same_base": (QA(), lib.TensorMesh(q16, origin=p0)),
        "quad_quad_general": (QS(), QB()),
        "quad_quad_general_rev": (QB(), QS()),
        "tensor_quad_general": (lib.TensorMesh(qt, origin=pt), QA()),
        "quad_tensor_general": (QA(), lib.TensorMesh(qt, origin=pt)),
        "quad_quad_disjoint": (QA(), _tree(lib, qs, np.array([5.5, -7.0]), c1, [3, 4])),
    }
    return {
        **_three_d(lib, A, B, S, h16, o0, ht, ot, hs, far, b1),
        **quad,
    }


def _three_d(lib, A, B, S, h16, o0, ht, ot, hs, far, b1):
    return {
        "tree_tree_same_base": (A(), B()),
        "tree_tree_same_base_rev": (B(), A()),
        "tensor_tree_same_base": (lib.TensorMesh(h16, origin=o0), A()),
        "tree_tensor_same_base": (A(), lib.TensorMesh(h16, origin=o0)),
        "tree_tree_general": (S(), B()),
        "tree_tree_general_rev": (B(), S()),
        "tensor_tree_general": (lib.TensorMesh(ht, origin=ot), A()),
        "tree_tensor_general": (A(), lib.TensorMesh(ht, origin=ot)),
        "tree_tensor_shared_boundary_plane": (A(), lib.TensorMesh(h16, origin=np.array([4.0, 0.0, 0.0]))),
        "tree_tree_disjoint": (A(), _tree(lib, hs, far, b1, [3, 4])),
        "tensor_tree_disjoint": (lib.TensorMesh(ht, origin=ot), _tree(lib, hs, far, b1, [3, 4])),
    }

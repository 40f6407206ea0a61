"""TensorMesh.reshape (base/base_regular_mesh.py:817-1001) over every combination of input kind, x_type, out_type and
return_format on 1-D / 2-D / 3-D meshes: exception types, output structure and values (as SHA-1 digests) pinned by
tests/golden/reshape_digests.npz, generated from the unmodified reference (tests/golden/make_golden_reshape.py)."""
import os
import sys
import warnings

import numpy as np

import discretize_b200 as dz

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import reshape_cases as rc  # noqa: E402


def test_reshape_matches_reference_digests():
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reshape_digests.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        keys, digests = rc.evaluate(dz)
    assert np.array_equal(keys, gold["keys"])
    bad = [(k, a, b) for k, a, b in zip(keys, digests, gold["digests"]) if a != b]
    assert not bad, bad[:5]
    assert sum(not d.startswith("EXC") for d in digests) > 200          # plenty of combinations produce values


def test_reshape_examples():
    m = dz.TensorMesh([3, 4, 2])
    v = np.arange(m.n_cells, dtype=float)
    assert m.reshape(v, "CC", "CC", "M").shape == (3, 4, 2) and m.reshape(v, "CC", "CC", "M")[1, 2, 1] == 1 + 3 * (2 + 4 * 1)
    f = np.arange(m.n_faces, dtype=float)
    fx, fy, fz = m.reshape(f, "F", "F", "M")
    assert fx.shape == (4, 4, 2) and fy.shape == (3, 5, 2) and fz.shape == (3, 4, 3)
    assert np.array_equal(m.reshape(f, "F", "Fy", "V"), f[m.n_faces_x:m.n_faces_x + m.n_faces_y])

"""volume_average (TensorMesh -> TensorMesh): host assembly against the reference-made fixture (CPU) and
the CUDA apply (GPU)."""
import numpy as np
import pytest

import discretize_b200 as dz
from discretize_b200.utils import _tensor_volume_average_matrix, volume_average
from tests.helpers import assert_close_vec, assert_same_matrix, golden_matrix, load_golden

TAGS = ["aligned", "random", "sticks_out"]


def _meshes(g, tag):
    mi = dz.TensorMesh([g[f"{tag}_hin{a}"] for a in "xyz"], origin=g[f"{tag}_oin"])
    mo = dz.TensorMesh([g[f"{tag}_hout{a}"] for a in "xyz"], origin=g[f"{tag}_oout"])
    return mi, mo


@pytest.mark.parametrize("tag", TAGS)
def test_matrix_matches_reference_fixture(tag):
    g = load_golden("volume_average_small.npz")
    mi, mo = _meshes(g, tag)
    A = _tensor_volume_average_matrix(mi, mo)
    assert_same_matrix(A, golden_matrix(g, tag), 1e-13)
    if tag != "sticks_out":      # a conservative average reproduces constants wherever the output lies inside the input
        np.testing.assert_allclose(A @ np.ones(mi.n_cells), 1.0, rtol=1e-12)


def test_argument_checks():
    a, b = dz.TensorMesh([4, 4, 4]), dz.TensorMesh([4, 4])
    with pytest.raises(ValueError):
        volume_average(a, b)
    with pytest.raises(TypeError):
        volume_average(a, object())
    with pytest.raises(ValueError):
        volume_average(a, dz.TensorMesh([2, 2, 2]), values=np.ones(3))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_apply_on_gpu(tag):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    g = load_golden("volume_average_small.npz")
    mi, mo = _meshes(g, tag)
    ref = golden_matrix(g, tag)
    x = np.random.default_rng(2).standard_normal(mi.n_cells)
    op = volume_average(mi, mo)
    assert_close_vec(op @ x, ref @ x)
    assert_close_vec(op.T @ (ref @ x), ref.T @ (ref @ x))
    assert_same_matrix(op.tocsr(), ref)
    out = np.empty(mo.n_cells)
    res = volume_average(mi, mo, x, out)
    assert res is out
    assert_close_vec(out, ref @ x)
    xd = torch.from_numpy(x).cuda()
    assert_close_vec(volume_average(mi, mo, xd).cpu().numpy(), ref @ x)

"""1-D / 2-D TensorMesh (SURVEY 8a "Dimension coverage"): host assembly (lowdim.py) against the reference-made fixture on
the CPU; the same operators through the mesh API on the GPU (generic CSR kernel)."""
import numpy as np
import pytest

import discretize_b200 as dz
from discretize_b200 import lowdim
import copy

from tests.golden.make_golden_lowdim import BCS, LOCS, MASS_CASES, bc_for, ops_for
from tests.helpers import assert_close_vec, assert_same_matrix, golden_matrix, load_golden


@pytest.fixture(scope="module")
def gold():
    return load_golden("tensor_lowdim_small.npz")


def _mesh(g, tag):
    d = int(tag[1])
    return dz.TensorMesh([g[f"{tag}_h{a}"] for a in range(d)], origin=g[f"{tag}_origin"])


@pytest.mark.parametrize("tag", ["d1", "d2"])
def test_host_assembly_matches_reference_fixture(gold, tag):
    m = _mesh(gold, tag)
    for name in ops_for(m.dim):
        if "cell_gradient" not in name:     # those go through lowdim.cell_gradient (boundary conditions)
            assert_same_matrix(lowdim.build(m, name), golden_matrix(gold, f"{tag}_{name}"), 1e-13)
        assert_same_matrix(getattr(m, name).tocsr(), golden_matrix(gold, f"{tag}_{name}"), 1e-13)   # no GPU needed
    for arr in ("cell_volumes", "face_areas", "edge_lengths"):
        np.testing.assert_array_equal(getattr(m, arr), gold[f"{tag}_{arr}"])
    for bct in BCS:
        m.set_cell_gradient_BC(copy.deepcopy(bc_for(m.dim, bct)))
        assert_same_matrix(m.stencil_cell_gradient.tocsr(), golden_matrix(gold, f"{tag}_stencil_cell_gradient_{bct}"), 1e-13)
        assert_same_matrix(m.cell_gradient.tocsr(), golden_matrix(gold, f"{tag}_cell_gradient_{bct}"), 1e-13)
    for proj in "FE":
        for mk, im, ix in MASS_CASES:
            d = lowdim.mass_diagonal(m, proj, gold[f"{tag}_model_{mk}"], im, ix)
            assert_close_vec(d, gold[f"{tag}_mass_{proj}_{mk}_{int(im)}{int(ix)}"], 1e-13)
    for proj in "FE":
        for mk, im, ix in MASS_CASES:
            a = lowdim.mass_deriv_matrix(m, proj, gold[f"{tag}_model_{mk}"], gold[f"{tag}_v{proj}"], im, ix)
            b = golden_matrix(gold, f"{tag}_deriv_{proj}_{mk}_{int(im)}{int(ix)}")
            a.eliminate_zeros(); b.eliminate_zeros()
            assert_same_matrix(a, b, 1e-12)
    if m.dim == 2:      # full-tensor model (nC, 3): the generic corner sum (inner_products.py:147-272, 459-565)
        sig = gold[f"{tag}_model_tensor"]
        for proj in "FE":
            for im in (False, True):
                a = lowdim.mass_tensor_matrix(m, proj, sig, im)
                b = golden_matrix(gold, f"{tag}_mass_{proj}_tensor_{int(im)}0")
                a.eliminate_zeros(); b.eliminate_zeros()
                assert_same_matrix(a, b, 1e-13)
            a = lowdim.mass_tensor_deriv_matrix(m, proj, gold[f"{tag}_v{proj}"])
            b = golden_matrix(gold, f"{tag}_deriv_{proj}_tensor_00")
            a.eliminate_zeros(); b.eliminate_zeros()
            assert_same_matrix(a, b, 1e-13)
        with pytest.raises(Exception, match="Solver needed to invert A."):
            m.get_face_inner_product(sig, invert_matrix=True)
        with pytest.raises(NotImplementedError):
            m.get_edge_inner_product_deriv(sig, invert_model=True)
    for loc in LOCS[m.dim]:
        assert_same_matrix(lowdim.interpolation_matrix(m, gold[f"{tag}_pts"], loc), golden_matrix(gold, f"{tag}_interp_{loc}"), 1e-13)
    assert_same_matrix(lowdim.interpolation_matrix(m, gold[f"{tag}_pts_out"], "CC", zeros_outside=True),
                       golden_matrix(gold, f"{tag}_interp_zo_CC"), 1e-13)
    with pytest.raises(ValueError, match="outside"):
        lowdim.interpolation_matrix(m, gold[f"{tag}_pts_out"], "CC")


def test_composites_assemble_without_a_gpu(gold):
    """Host-assembled operators (CSR and numpy diagonals) compose and materialise on the CPU: system matrices for a direct
    solver do not need the device."""
    m = _mesh(gold, "d2")
    Gr = golden_matrix(gold, "d2_nodal_gradient")
    Dr = golden_matrix(gold, "d2_face_divergence")
    sig = gold["d2_model_iso"]
    A = (m.nodal_gradient.T @ m.get_edge_inner_product(sig) @ m.nodal_gradient).tocsr()
    Ar = Gr.T @ np.diag(gold["d2_mass_E_iso_00"]) @ Gr
    assert np.abs(A.toarray() - Ar).max() <= 1e-13 * np.abs(Ar).max()
    B = (m.face_divergence @ m.get_face_inner_product(sig, invert_matrix=True) @ m.face_divergence.T).tocsr()
    Br = Dr @ np.diag(gold["d2_mass_F_iso_01"]) @ Dr.T
    assert np.abs(B.toarray() - Br).max() <= 1e-13 * np.abs(Br).max()


def test_errors_keep_the_reference_behaviour():
    with pytest.raises(NotImplementedError, match="Edge Curl"):
        dz.TensorMesh([5]).edge_curl
    with pytest.raises(NotImplementedError):
        dz.TensorMesh([5, 4]).cell_gradient_z
    with pytest.raises(NotImplementedError):
        dz.TensorMesh([5]).average_face_y_to_cell
    with pytest.raises(NotImplementedError):
        dz.TensorMesh([5]).get_face_inner_product(np.ones((5, 3)))           # no full tensor in 1-D


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["d1", "d2"])
def test_lowdim_operators_on_gpu(gold, tag):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    m = _mesh(gold, tag)
    rng = np.random.default_rng(3)
    for name in ops_for(m.dim):
        ref = golden_matrix(gold, f"{tag}_{name}")
        op = getattr(m, name)
        x, y = rng.standard_normal(ref.shape[1]), rng.standard_normal(ref.shape[0])
        assert_close_vec(op @ x, ref @ x)
        assert_close_vec(op.T @ y, ref.T @ y)
    n = {"F": m.n_faces, "E": m.n_edges}
    for proj, get in (("F", m.get_face_inner_product), ("E", m.get_edge_inner_product)):
        for mk, im, ix in MASS_CASES:
            M = get(gold[f"{tag}_model_{mk}"], invert_model=im, invert_matrix=ix)
            x = rng.standard_normal(n[proj])
            assert_close_vec(M @ x, gold[f"{tag}_mass_{proj}_{mk}_{int(im)}{int(ix)}"] * x)
    for loc in LOCS[m.dim]:
        ref = golden_matrix(gold, f"{tag}_interp_{loc}")
        Q = m.get_interpolation_matrix(gold[f"{tag}_pts"], loc)
        x = rng.standard_normal(ref.shape[1])
        assert_close_vec(Q @ x, ref @ x)
        assert_same_matrix(Q.tocsr(), ref)
    # a 2-D composite through the operator algebra: G^T Me G
    if m.dim == 2:
        sig = gold[f"{tag}_model_iso"]
        G = m.nodal_gradient
        A = G.T @ m.get_edge_inner_product(sig) @ G
        Gr = golden_matrix(gold, f"{tag}_nodal_gradient")
        Ar = Gr.T @ np.diag(gold[f"{tag}_mass_E_iso_00"]) @ Gr
        x = rng.standard_normal(m.n_nodes)
        assert_close_vec(A @ x, Ar @ x)
        # full-tensor model: host-assembled, applied by the CSR kernel
        ref = golden_matrix(gold, f"{tag}_mass_E_tensor_00")
        M = m.get_edge_inner_product(gold[f"{tag}_model_tensor"])
        x = rng.standard_normal(m.n_edges)
        assert_close_vec(M @ x, ref @ x)

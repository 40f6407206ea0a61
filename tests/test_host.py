"""CPU tests: the C-ABI library loads and exports what include/dzb.h declares; host logic."""
import ctypes
import os
import re

import numpy as np
import pytest

import discretize_b200 as dz
from discretize_b200 import _lib
from oracle import tensor_ops as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dzb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dzb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from discretize_b200._build import build

        build(verbose=False)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"libdzb.so does not export {n}"
        assert n in _lib.SIGNATURES, f"python binding lacks {n}"
    lib.dzb_version.restype = ctypes.c_int
    assert lib.dzb_version() == _lib.ABI_VERSION


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """The boundary is a C ABI: include/dzb.h compiles as C99 with no C++ or torch types, and a C program that calls
    an entry point links against libdzb.so and runs (dzb_version needs no GPU)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "dzb.h"\nint main(void) { printf("%d\\n", dzb_version()); return 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)],
                   check=True)
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-ldzb",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert int(out) == _lib.ABI_VERSION


def test_shapes_through_the_abi():
    lib = _lib.load()
    mesh = dz.TensorMesh([5, 4, 3])
    g = T.Grid([5, 4, 3])
    expect = {
        _lib.OP_DIV: (g.nC, g.nF), _lib.OP_CURL: (g.nF, g.nE), _lib.OP_GRAD: (g.nE, g.nN),
        _lib.OP_AVE_F2CC: (g.nC, g.nF), _lib.OP_AVE_F2CCV: (3 * g.nC, g.nF), _lib.OP_AVE_E2CC: (g.nC, g.nE),
        _lib.OP_AVE_E2CCV: (3 * g.nC, g.nE), _lib.OP_AVE_N2CC: (g.nC, g.nN),
        _lib.OP_AVE_FX2CC: (g.nC, g.vnF[0]), _lib.OP_AVE_EY2CC: (g.nC, g.vnE[1]),
    }
    for op, shp in expect.items():
        nr, nc = ctypes.c_int64(), ctypes.c_int64()
        assert lib.dzb_tm_shape(ctypes.byref(mesh._host_mesh_counts()), op, ctypes.byref(nr), ctypes.byref(nc)) == 0
        assert (nr.value, nc.value) == shp
    nr, nc = ctypes.c_int64(), ctypes.c_int64()
    assert lib.dzb_tm_shape(ctypes.byref(mesh._host_mesh_counts()), 99, ctypes.byref(nr), ctypes.byref(nc)) != 0
    assert b"unknown operator" in lib.dzb_last_error()


def test_mesh_construction_matches_oracle_layout():
    rng = np.random.default_rng(0)
    h = [rng.random(4) + 0.5, rng.random(6) + 0.5, rng.random(5) + 0.5]
    mesh, g = dz.TensorMesh(h, origin=[-1.0, 0.25, 3.0]), T.Grid(h, [-1.0, 0.25, 3.0])
    assert (mesh.nC, mesh.nN, mesh.nF, mesh.nE) == (g.nC, g.nN, g.nF, g.nE)
    assert mesh.vnF == g.vnF and mesh.vnE == g.vnE
    assert mesh.shape_edges_y == g.shapes("Ey") and mesh.shape_faces_z == g.shapes("Fz")
    for a, nm in enumerate("xyz"):
        np.testing.assert_array_equal(getattr(mesh, "nodes_" + nm), g.nodes[a])
        np.testing.assert_array_equal(getattr(mesh, "cell_centers_" + nm), g.centers[a])
    np.testing.assert_array_equal(mesh.cell_volumes, g.cell_volumes())
    np.testing.assert_array_equal(mesh.face_areas, g.face_areas())
    np.testing.assert_array_equal(mesh.edge_lengths, g.edge_lengths())
    assert mesh.cell_centers.shape == (g.nC, 3) and mesh.edges_z.shape == (g.vnE[2], 3)
    np.testing.assert_array_equal(mesh.is_inside(np.array([[0.0, 1.0, 4.0], [100.0, 0, 0]])),
                                  T.is_inside(g, np.array([[0.0, 1.0, 4.0], [100.0, 0, 0]])))


def test_constructor_forms():
    m = dz.TensorMesh([4, 2, 8])
    np.testing.assert_allclose(m.h[0], np.ones(4) / 4)
    m = dz.TensorMesh([[(10.0, 3, -1.5), (10.0, 2), (10.0, 3, 1.5)], [1.0, 2.0], 2], origin="CN0")
    assert m.shape_cells == (8, 2, 2)
    np.testing.assert_allclose(m.h[0], [33.75, 22.5, 15.0, 10, 10, 15, 22.5, 33.75])
    np.testing.assert_allclose(m.origin, [-m.h[0].sum() / 2, -3.0, 0.0])
    assert dz.TensorMesh([3, 3, 3], x0=[1, 2, 3]).origin.tolist() == [1.0, 2.0, 3.0]
    with pytest.raises(TypeError):
        dz.TensorMesh(5)
    with pytest.raises(ValueError):
        dz.TensorMesh([2, 2, 2, 2])
    with pytest.raises(ValueError):
        dz.TensorMesh([2, 2, 2], origin=[0, 0])
    # lower dimensions construct; their operators are host-assembled CSR (lowdim.py), what has no 1-D / 2-D meaning or
    # no kernel refuses loudly
    m2 = dz.TensorMesh([4, 3])
    assert m2.dim == 2 and m2.nC == 12 and m2.nN == 20
    assert m2.face_divergence.shape == (12, 31) and m2.edge_curl.shape == (12, 31)
    assert m2.cell_gradient.shape == (31, 12)
    assert callable(m2.get_face_inner_product_deriv(np.ones((12, 3))))    # 2-D full tensor (lowdim.mass_tensor_deriv_matrix)
    with pytest.raises(NotImplementedError):
        m2.get_face_inner_product_deriv(np.ones((12, 3)), invert_model=True)
    with pytest.raises(NotImplementedError):
        dz.TensorMesh([4]).edge_curl


def test_operator_algebra_shapes_without_gpu():
    mesh = dz.TensorMesh([5, 4, 3])
    D, C, G = mesh.face_divergence, mesh.edge_curl, mesh.nodal_gradient
    assert (D @ C).shape == (mesh.nC, mesh.nE)
    assert (C.T @ C).shape == (mesh.nE, mesh.nE)
    assert (2 * G).shape == G.shape and (G.T).shape == (mesh.nN, mesh.nE)
    assert G.T.T.transposed is False
    with pytest.raises(ValueError):
        D @ G
    with pytest.raises(ValueError):
        D + G
    assert mesh.aveF2CC is mesh.average_face_to_cell
    with pytest.raises(AttributeError):
        mesh.not_a_thing
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(dz.DzbError, match="no CUDA device"):
            D @ np.ones(mesh.nF)


def test_operator_algebra_values_without_gpu():
    """Sums, products, scalings and transposes of host-assembled operators materialise to the scipy expression (the
    same algebra layer composes the CUDA operators; here it runs on QuadTree matrices, which need no device)."""
    import scipy.sparse as sp

    m = dz.TreeMesh([16, 8])
    m.refine_ball(np.array([[0.4, 0.5]]), [0.25], [-1])
    assert m.n_hanging_edges > 0
    rng = np.random.default_rng(0)
    sig = rng.random(m.n_cells) + 0.5
    D, G, C = m.face_divergence, m.nodal_gradient, m.edge_curl
    Ds, Gs, Cs = (sp.csr_matrix(x.tocsr()) for x in (D, G, C))
    Me, Mf = m.get_edge_inner_product(sig), m.get_face_inner_product(sig, invert_matrix=True)
    Mes, Mfs = Me.tocsr(), Mf.tocsr()

    def close(op, ref):
        got = op.tocsr()
        assert got.shape == ref.shape
        assert abs(got - ref).max() <= 1e-13 * max(1.0, abs(ref).max())

    close(G.T @ Me @ G, Gs.T @ Mes @ Gs)
    close(D @ Mf @ D.T, Ds @ Mfs @ Ds.T)
    close(2.5 * (C.T @ C) + Me, 2.5 * (Cs.T @ Cs) + Mes)
    close((C.T @ C + Me).T, (Cs.T @ Cs + Mes).T)
    close(-(D @ D.T) + 3.0 * sp.identity(m.n_cells, format="csr"), -(Ds @ Ds.T) + 3.0 * sp.identity(m.n_cells))
    close(G.T.T, Gs)
    close((G.T @ Me) @ G - 0.5 * (G.T @ Me @ G), 0.5 * (Gs.T @ Mes @ Gs))
    close(Gs.T @ Me @ G, Gs.T @ Mes @ Gs)                      # a scipy matrix on the left joins the algebra
    with pytest.raises(ValueError):
        D @ D                                                  # (nC x nF) @ (nC x nF): shapes do not chain
    assert (G.T @ Me @ G).shape == (m.n_nodes, m.n_nodes)

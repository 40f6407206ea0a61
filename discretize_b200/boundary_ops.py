"""Boundary-condition pieces of the weak forms: projections, Robin terms, boundary integrals.

Reference: ``operators/differential_operators.py`` -- ``edge_divergence_weak_form_robin`` (:772-960),
``cell_gradient_weak_form_robin`` (:1597-1731), ``boundary_edge_vector_integral`` (:2199-2254),
``boundary_node_vector_integral`` (:2257-2301), ``get_BC_projections`` (:2303-2408), ``get_BC_projections_simple``
(:2410-2477).  They apply to TensorMesh and TreeMesh alike, so they are written once against a small mesh protocol:

* ``mesh._boundary_indices()``  -> (faces, edges, nodes): sorted global indices of the boundary entities;
* ``mesh._boundary_face_rows(name)`` -> scipy CSR of ``project_face_to_boundary_face @ mesh.<name>``.

Everything here is O(boundary) host work; the matrices come back as :class:`~discretize_b200.csr.CsrOperator`
(applied on the GPU by ``dzb_csr_spmv``, ``.tocsr()`` for the reference's matrix), the vectors as numpy arrays.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .csr import CsrOperator


def _validate_bc(bc):
    from .tensor_mesh import _validate_BC

    return _validate_BC(bc)


def _restrict_columns(rows, idx):
    """``rows @ P.T`` for the selection ``P`` of the sorted global indices ``idx``; every entry must land on ``idx``."""
    rows = rows.tocoo()
    cols = np.searchsorted(idx, rows.col)
    ok = (cols < idx.size)
    ok[ok] &= idx[cols[ok]] == rows.col[ok]
    # interior columns are dropped exactly like the product with P.T drops them
    return sp.csr_matrix((rows.data[ok], (rows.row[ok], cols[ok])), shape=(rows.shape[0], idx.size))


def _scatter(idx, values, n):
    """``P.T @ values``: boundary values placed into a zero vector (or (n, k) array) over all entities."""
    values = np.asarray(values)
    out = np.zeros((n,) + values.shape[1:], dtype=np.result_type(values.dtype, np.float64))
    out[idx] = values
    return out


def _boundary_areas(mesh, bf):
    return np.asarray(mesh.face_areas)[bf]


# ---- Robin conditions --------------------------------------------------------------------------------------------


def edge_divergence_weak_form_robin(mesh, alpha=0.0, beta=1.0, gamma=0.0):
    """(B, b) with ``<psi, div u> = psi.(-G^T Me G + B) phi + psi.b`` under ``alpha phi + beta dphi/dn = gamma``."""
    alpha, beta, gamma = np.atleast_1d(alpha), np.atleast_1d(beta), np.atleast_1d(gamma)
    if np.any(beta == 0.0):
        raise ValueError("beta cannot have a zero value")
    bf, _, bn = mesh._boundary_indices()
    nbf, nbn = bf.size, bn.size
    if len(alpha) == 1:
        n = len(beta) if len(beta) != 1 else len(gamma) if len(gamma) != 1 else nbf
        alpha = np.full(n, alpha[0])
    if len(beta) == 1 and len(alpha) != 1:
        beta = np.full(len(alpha), beta[0])
    if len(gamma) == 1 and len(alpha) != 1:
        gamma = np.full(len(alpha), gamma[0])
    if len(alpha) != len(beta) or len(beta) != len(gamma):
        raise ValueError("alpha, beta, and gamma must have the same length")
    if len(alpha) not in [nbf, nbn]:
        raise ValueError("The arrays must be of length n_boundary_faces or n_boundary_nodes")
    areas = _boundary_areas(mesh, bf)
    ave = _restrict_columns(mesh._boundary_face_rows("average_node_to_face"), bn)       # boundary faces <- boundary nodes
    # on the boundary  u.n = (gamma - alpha phi) / beta
    if len(alpha) == nbf:
        g = gamma / beta[:, None] * areas[:, None] if gamma.ndim == 2 else gamma / beta * areas
        b = ave.T @ g
        diag = ave.T @ (-alpha / beta * areas)
    else:
        node_areas = ave.T @ areas
        b = gamma / beta[:, None] * node_areas[:, None] if gamma.ndim == 2 else gamma / beta * node_areas
        diag = -alpha / beta * node_areas
    n_nodes = mesh.n_nodes
    B = sp.csr_matrix((diag, (bn, bn)), shape=(n_nodes, n_nodes))
    # the reference returns a full diagonal; keep its (explicit zero) interior entries out of the device matrix
    return CsrOperator(B), _scatter(bn, b, n_nodes)


def cell_gradient_weak_form_robin(mesh, alpha=0.0, beta=1.0, gamma=0.0):
    """(B, b) with ``<u, grad phi> = u.(-D^T Mc + B) phi + u.b`` under ``alpha phi + beta dphi/dn = gamma``."""
    ave = mesh._boundary_face_rows("average_cell_to_face")
    if mesh.dim == 1:
        h = np.abs(np.asarray(mesh.boundary_faces) - ave @ np.asarray(mesh.cell_centers))
    else:
        h = np.linalg.norm(np.asarray(mesh.boundary_faces) - ave @ np.asarray(mesh.cell_centers), axis=1)
    # ghost value  u_k = a u_i + b
    a = beta / h / (alpha + beta / h)
    A = sp.diags(a) @ ave
    gamma = np.asarray(gamma)
    b = gamma / (alpha + beta / h)[:, None] if gamma.ndim > 1 else gamma / (alpha + beta / h)
    M = mesh.boundary_face_scalar_integral.tocsr()
    return CsrOperator((M @ A).tocsr()), M @ b


# ---- boundary integrals ------------------------------------------------------------------------------------------


def _outward_area_vectors(mesh, bf):
    return np.asarray(mesh.boundary_face_outward_normals) * _boundary_areas(mesh, bf)[:, None]


def boundary_edge_vector_integral(mesh):
    """``w^T P u_b`` = integral of ``w . (u x n)`` over the boundary; (n_edges, n_be) in 2-D, (n_edges, 3 n_be) in 3-D."""
    if mesh.dim == 1:
        raise NotImplementedError("boundary_edge_vector_integral: a 1-D mesh has no boundary edges")
    bf, be, _ = mesh._boundary_indices()
    dA = _outward_area_vectors(mesh, bf)
    w = np.asarray(mesh.edge_tangents)[be]
    nbe = be.size
    ave = _restrict_columns(mesh._boundary_face_rows("average_edge_to_face"), be)
    if mesh.dim > 2:
        ave = ave * 2
    av_da = ave.T @ dA
    if mesh.dim == 2:
        cross = av_da[:, 0] * w[:, 1] - av_da[:, 1] * w[:, 0]
        return CsrOperator(sp.csr_matrix((cross, (be, np.arange(nbe))), shape=(mesh.n_edges, nbe)))
    cross = np.cross(av_da, w)
    rows = np.tile(be, 3)
    cols = np.arange(3 * nbe)
    return CsrOperator(sp.csr_matrix((cross.T.reshape(-1), (rows, cols)), shape=(mesh.n_edges, 3 * nbe)))


def boundary_node_vector_integral(mesh):
    """``w^T P u_b`` = integral of ``(w u) . n`` over the boundary; (n_nodes, dim * n_bn)."""
    d = mesh.dim
    if d == 1:
        n = mesh.shape_nodes[0]
        return CsrOperator(sp.csr_matrix(([-1.0, 1.0], ([0, n - 1], [0, 1])), shape=(n, 2)))
    bf, _, bn = mesh._boundary_indices()
    dA = _outward_area_vectors(mesh, bf)
    ave = _restrict_columns(mesh._boundary_face_rows("average_node_to_face"), bn)
    u_dot_ds = ave.T @ dA                                     # (n_bn, dim)
    nbn = bn.size
    rows = np.tile(bn, d)
    return CsrOperator(sp.csr_matrix((u_dot_ds.T.reshape(-1), (rows, np.arange(d * nbn))), shape=(mesh.n_nodes, d * nbn)))


# ---- cell-centred weak-form projections (tensor-product ordering of the faces) -----------------------------------------


def _kron_axes(mats):
    """Tensor product with the FIRST axis fastest (x fastest): kron(last, ..., first)."""
    out = mats[0]
    for m in mats[1:]:
        out = sp.kron(m, out, format="csr")
    return sp.csr_matrix(out)


def _per_axis(mesh, piece):
    """block_diag over face directions of (identity over the other axes) x (``piece(n_a, a)`` along the normal)."""
    n = mesh.shape_cells
    blocks = []
    for a in range(mesh.dim):
        blocks.append(_kron_axes([piece(n[b], a) if b == a else sp.identity(n[b], format="csr") for b in range(mesh.dim)]))
    return sp.block_diag(blocks, format="csr") if len(blocks) > 1 else blocks[0]


def _boundary_face_areas_by_direction(mesh):
    ind = mesh.face_boundary_indices
    mask = np.concatenate([np.asarray(ind[2 * a]) | np.asarray(ind[2 * a + 1]) for a in range(mesh.dim)])
    return np.asarray(mesh.face_areas)[mask]


def _dirichlet_piece(bcs):
    def piece(n, a):
        bc = bcs[a]
        vals = [-1.0 if bc[0] == "dirichlet" else 0.0, 1.0 if bc[1] == "dirichlet" else 0.0]
        return sp.csr_matrix((vals, ([0, n], [0, 1])), shape=(n + 1, 2))
    return piece


def get_BC_projections(mesh, BC, discretization="CC"):
    """(Pbc, Pin, Pout) of the reference's cell-centred weak form; ``BC``: str, or one entry (str or pair) per axis."""
    if discretization != "CC":
        raise NotImplementedError("Boundary conditions only implementedfor CC discretization.")
    if isinstance(BC, str):
        BC = [BC for _ in mesh.shape_cells]
    elif isinstance(BC, list):
        if len(BC) != mesh.dim:
            raise ValueError("BC list must be the size of your mesh")
    else:
        raise TypeError("BC must be a str or a list.")
    for i, bc_i in enumerate(BC):
        BC[i] = _validate_bc(bc_i)

    def neumann_in(n, a):
        keep = np.ones(n + 1, dtype=bool)
        keep[0] = BC[a][0] != "neumann"
        keep[n] = BC[a][1] != "neumann"
        idx = np.nonzero(keep)[0]
        return sp.csr_matrix((np.ones(idx.size), (np.arange(idx.size), idx)), shape=(idx.size, n + 1))

    def neumann_out(n, a):
        vals = [1.0 if BC[a][0] == "neumann" else 0.0, 1.0 if BC[a][1] == "neumann" else 0.0]
        return sp.csr_matrix((vals, ([0, 1], [0, n])), shape=(2, n + 1))

    Pbc = _per_axis(mesh, _dirichlet_piece(BC)) @ sp.diags(_boundary_face_areas_by_direction(mesh))
    return CsrOperator(Pbc.tocsr()), CsrOperator(_per_axis(mesh, neumann_in)), CsrOperator(_per_axis(mesh, neumann_out))


def get_BC_projections_simple(mesh, discretization="CC"):
    """(Pbc, B^T) for mixed conditions: all-Dirichlet ``Pbc`` and the unsigned boundary-face selection."""
    if discretization != "CC":
        raise NotImplementedError("Boundary conditions only implementedfor CC discretization.")
    BC = [["dirichlet", "dirichlet"]] * 3

    def select(n, a):
        return sp.csr_matrix(([1.0, 1.0], ([0, n], [0, 1])), shape=(n + 1, 2))

    Pbc = _per_axis(mesh, _dirichlet_piece(BC)) @ sp.diags(_boundary_face_areas_by_direction(mesh))
    return CsrOperator(Pbc.tocsr()), CsrOperator(_per_axis(mesh, select).T.tocsr())

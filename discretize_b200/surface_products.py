"""Inner products for properties that live on faces or edges (thin sheets, wires) and their model derivatives.

Reference: ``base/base_mesh.py`` -- ``get_face_inner_product_surface`` (:2009-2156), ``get_edge_inner_product_line``
(:2158-2310), their ``_deriv`` forms (:2717-2818, :2820-2925); ``operators/inner_products.py`` --
``get_edge_inner_product_surface`` (:101-139) and ``_deriv`` (:328-398).  All three matrices are diagonal,

    face surface :  diag(face_areas   o tau)
    edge line    :  diag(edge_lengths o tau)
    edge surface :  (dim - 1) diag(Av_e2f^T (face_areas o tau))

so one description ``(W, n_model)`` -- the weight matrix that maps the model to the diagonal -- serves the matrix, its
inverse and the four derivative forms.  Same for TensorMesh and TreeMesh; the mesh supplies ``face_areas``,
``edge_lengths`` and ``_host_matrix("average_edge_to_face")``.  Results are
:class:`~discretize_b200.csr.CsrOperator` (GPU apply via ``dzb_csr_spmv``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .csr import CsrOperator
from .tensor_mesh import is_scalar

_KINDS = {
    # kind: (size attribute of the model, its name in the reference's message)
    "face_surface": ("n_faces", "faces"),
    "edge_surface": ("n_faces", "faces"),
    "edge_line": ("n_edges", "edges"),
}


def _weights(mesh, kind):
    """W with  diagonal(M) = W @ model."""
    if kind == "face_surface":
        return sp.diags(np.asarray(mesh.face_areas, dtype=np.float64), format="csr")
    if kind == "edge_line":
        return sp.diags(np.asarray(mesh.edge_lengths, dtype=np.float64), format="csr")
    av = mesh._host_matrix("average_edge_to_face")
    return ((mesh.dim - 1) * av.T @ sp.diags(np.asarray(mesh.face_areas, dtype=np.float64))).tocsr()


def _diagonal(mesh, kind, model, invert_model, invert_matrix):
    n = getattr(mesh, _KINDS[kind][0])
    if model is None:
        model = np.ones(n)
    if invert_model:
        model = 1.0 / model
    if is_scalar(model):
        model = model * np.ones(n)
    if model.size != n:
        raise ValueError("Unexpected shape of tensor: {}".format(model.shape),
                         "Must be scalar or have length equal to total number of {}.".format(_KINDS[kind][1]))
    d = _weights(mesh, kind) @ np.asarray(model, dtype=np.float64).reshape(-1, order="F")
    return 1.0 / d if invert_matrix else d


def inner_product(mesh, kind, model=None, invert_model=False, invert_matrix=False):
    return CsrOperator(sp.diags(_diagonal(mesh, kind, model, invert_model, invert_matrix), format="csr"))


def inner_product_deriv(mesh, kind, model, invert_model=False, invert_matrix=False):
    """``v -> d(M(m) v)/dm`` as an operator factory, or None when there is no model (like the reference)."""
    n = getattr(mesh, _KINDS[kind][0])
    if model is None:
        return None
    scalar = is_scalar(model)
    if not scalar and model.size != n:
        if kind == "edge_line":
            raise ValueError("Unexpected shape of tensor: {}.".format(model.shape),
                             "Must be scalar or have length equal to total number of edges: {}.".format(n))
        raise ValueError("Unexpected shape of tensor: {}".format(model.shape),
                         "Must be scalar or have length equal to total number of faces.")
    W = _weights(mesh, kind)
    if invert_model and scalar and not isinstance(model, np.ndarray):
        # the reference builds diag(1/model**2) from the bare number and its sdiag refuses it
        raise TypeError("Vector must be a numpy array")
    left = None
    if invert_matrix:
        mi = _diagonal(mesh, kind, model, invert_model, True)
        left = mi**2 if invert_model else -(mi**2)
    d = W
    if scalar and not (invert_matrix and not invert_model):
        # one parameter: sum the columns ...  except the reference's invert_matrix-only branch, which keeps them all
        if invert_model and not invert_matrix:
            raise ValueError("dimension mismatch: the reference multiplies an (n, n) weight by a (1, 1) diagonal")
        d = sp.csr_matrix(W @ np.ones((W.shape[1], 1)))
    if invert_model:
        m2 = 1.0 / np.asarray(model, dtype=np.float64).reshape(-1) ** 2
        d = d @ sp.diags(m2 if invert_matrix else -m2)
    if left is not None:
        d = sp.diags(left) @ d
    d = sp.csr_matrix(d)

    def deriv(v):
        return CsrOperator((sp.diags(np.asarray(v, dtype=np.float64).reshape(-1)) @ d).tocsr())

    return deriv

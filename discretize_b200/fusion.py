"""Recognise the composites SimPEG builds and replace them with single fused kernels.

``G.T @ Me(sigma) @ G``            -> DCNodalOperator  (K8)
``C.T @ Mf(mui) @ C + Me(sigma)``  -> MaxwellOperator  (K9)

Anything else stays a lazy chain of the elementary kernels.
"""
from __future__ import annotations

from . import _lib


def _is_stencil(op, op_id, transposed):
    from .tensor_ops import StencilOperator

    return isinstance(op, StencilOperator) and op.op_id == op_id and op.transposed == transposed


def _plain_iso_mass(op, proj):
    """A diagonal mass matrix of an isotropic nC model with no inversion flags."""
    from .tensor_ops import MassOperator

    return (isinstance(op, MassOperator) and op.proj == proj and op.model.kind == _lib.MODEL_ISO
            and not op.invert_model and not op.invert_matrix)


def _inverse_iso_face_mass(op):
    """Mf(rho)^-1: diagonal face mass matrix of an isotropic nC model, invert_matrix only."""
    from .tensor_ops import MassOperator

    return (isinstance(op, MassOperator) and op.proj == "F" and op.model.kind == _lib.MODEL_ISO
            and not op.invert_model and op.invert_matrix)


def fuse_product(factors):
    from .tensor_ops import DCCellOperator, DCNodalOperator

    if len(factors) == 3:
        a, m, b = factors
        if (_is_stencil(a, _lib.OP_DIV, False) and _is_stencil(b, _lib.OP_DIV, True)
                and _inverse_iso_face_mass(m) and a.mesh is b.mesh is m.mesh):
            return DCCellOperator(m.mesh, m.model.dev)
        if (_is_stencil(a, _lib.OP_GRAD, True) and _is_stencil(b, _lib.OP_GRAD, False)
                and _plain_iso_mass(m, "E") and a.mesh is b.mesh is m.mesh):
            return DCNodalOperator(m.mesh, m.model.dev)
    return None


def _curl_curl(op):
    """(mesh, mui) if op is C.T @ Mf(mui) @ C with an isotropic, un-inverted face mass matrix."""
    from .operators import ProductOperator

    if not isinstance(op, ProductOperator) or len(op.factors) != 3:
        return None
    a, m, b = op.factors
    if (_is_stencil(a, _lib.OP_CURL, True) and _is_stencil(b, _lib.OP_CURL, False)
            and _plain_iso_mass(m, "F") and a.mesh is b.mesh is m.mesh):
        return m.mesh, m.model.dev
    return None


def fuse_sum(terms):
    from .tensor_ops import MaxwellOperator

    if len(terms) == 2:
        for cc, me in ((terms[0], terms[1]), (terms[1], terms[0])):
            hit = _curl_curl(cc)
            if hit is not None and _plain_iso_mass(me, "E") and me.mesh is hit[0]:
                return MaxwellOperator(hit[0], hit[1], me.model.dev)
    return None

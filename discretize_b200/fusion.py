"""Recognise the composites SimPEG builds and replace them with single fused kernels.

``G.T @ Me(sigma) @ G``                    -> DCNodalOperator  (K8)
``C.T @ Mf(mui) @ C + Me(sigma)``          -> MaxwellOperator  (K9)
``C.T @ Mf(mui) @ C + alpha * Me(sigma)``  -> MaxwellOperator with alpha folded into sigma (alpha real), or
                                              ComplexMaxwellOperator (alpha complex: ``1j * omega * Me``, the system
                                              SimPEG's frequency-domain Maxwell solves apply)
``D @ Mf(rho)^-1 @ D.T``                   -> DCCellOperator   (K13)

The mass matrices may carry an isotropic cell model, a scalar, or no model at all (``get_face_inner_product()`` as in
the reference's tests/boundaries/test_boundary_maxwell.py:95-99): scalars are expanded to a constant cell array.
Anything else stays a lazy chain of the elementary kernels.
"""
from __future__ import annotations

import torch

from . import _lib


def _is_stencil(op, op_id, transposed):
    from .tensor_ops import StencilOperator

    return isinstance(op, StencilOperator) and op.op_id == op_id and op.transposed == transposed


def _cell_model(op):
    """The nC device array behind a mass matrix with an isotropic, scalar or absent model."""
    if op.model.kind == _lib.MODEL_ISO:
        return op.model.dev
    from .operators import cuda_device

    return torch.full((op.mesh.n_cells,), float(op.model.scalar), dtype=torch.float64, device=cuda_device())


def _plain_mass(op, proj):
    """A diagonal mass matrix (isotropic nC model, scalar or no model) with no inversion flags."""
    from .tensor_ops import MassOperator

    return (isinstance(op, MassOperator) and op.proj == proj and op.model.kind in (_lib.MODEL_ISO, _lib.MODEL_SCALAR)
            and not op.invert_model and not op.invert_matrix)


def _inverse_face_mass(op):
    """Mf(rho)^-1: diagonal face mass matrix of an isotropic / scalar model, invert_matrix only."""
    from .tensor_ops import MassOperator

    return (isinstance(op, MassOperator) and op.proj == "F" and op.model.kind in (_lib.MODEL_ISO, _lib.MODEL_SCALAR)
            and not op.invert_model and op.invert_matrix)


def fuse_product(factors):
    from .tensor_ops import DCCellOperator, DCNodalOperator

    if len(factors) == 3:
        a, m, b = factors
        if (_is_stencil(a, _lib.OP_DIV, False) and _is_stencil(b, _lib.OP_DIV, True)
                and _inverse_face_mass(m) and a.mesh is b.mesh is m.mesh):
            return DCCellOperator(m.mesh, _cell_model(m))
        if (_is_stencil(a, _lib.OP_GRAD, True) and _is_stencil(b, _lib.OP_GRAD, False)
                and _plain_mass(m, "E") and a.mesh is b.mesh is m.mesh):
            return DCNodalOperator(m.mesh, _cell_model(m))
    return None


def _curl_curl(op):
    """(mesh, mui) if op is C.T @ Mf(mui) @ C with an un-inverted diagonal face mass matrix."""
    from .operators import ProductOperator

    if not isinstance(op, ProductOperator) or len(op.factors) != 3:
        return None
    a, m, b = op.factors
    if (_is_stencil(a, _lib.OP_CURL, True) and _is_stencil(b, _lib.OP_CURL, False)
            and _plain_mass(m, "F") and a.mesh is b.mesh is m.mesh):
        return m.mesh, _cell_model(m)
    return None


def fuse_sum(terms):
    from .operators import ScaledOperator
    from .tensor_ops import ComplexMaxwellOperator, MaxwellOperator

    if len(terms) == 2:
        for cc, me in ((terms[0], terms[1]), (terms[1], terms[0])):
            hit = _curl_curl(cc)
            if hit is None:
                continue
            alpha = 1.0
            if isinstance(me, ScaledOperator):
                me, alpha = me.base, me.alpha
            if not (_plain_mass(me, "E") and me.mesh is hit[0]):
                continue
            sigma = _cell_model(me)
            if isinstance(alpha, complex):
                return ComplexMaxwellOperator(hit[0], hit[1], sigma, alpha)
            return MaxwellOperator(hit[0], hit[1], sigma if alpha == 1.0 else sigma * alpha)
    return None

"""Jacobi-preconditioned conjugate gradients on the fused operators.

Where SimPEG spends its wall time once the apply is fast: ``A x = b`` with A = G^T Me(sigma) G (DC resistivity) or
A = C^T Mf(1/mu) C + Me(sigma) (Maxwell, real SPD form).  One iteration is the operator apply plus THREE launches of
``csrc/solver.cu`` (dot, fused x/r/z update with both reductions, direction update) -- two where the operator computes
p.Ap inside its apply (``DCNodalOperator.apply_with_dot``, K8: 112 instead of 128 bytes per node and iteration); alpha, beta and all dot products
stay on the device, the host reads the residual norm only every ``check_every`` iterations.  The Jacobi diagonal of the
fused operators comes analytically from one apply of the entrywise-squared stencil to the mass diagonal
(``DCNodalOperator.diagonal_device``, ``MaxwellOperator.diagonal_device``).
"""
from __future__ import annotations

import torch

from ._lib import check, load
from .operators import cuda_device, ptr, stream_ptr, to_device


class CGInfo:
    def __init__(self, iterations, residual_norm, rhs_norm, converged, history):
        self.iterations, self.residual_norm, self.rhs_norm = iterations, residual_norm, rhs_norm
        self.converged, self.history = converged, history

    def __repr__(self):
        return (f"CGInfo(iterations={self.iterations}, relative_residual={self.residual_norm / max(self.rhs_norm, 1e-300):.3e}, "
                f"converged={self.converged})")


def pcg(A, b, x0=None, inv_diag=None, rtol=1e-8, maxiter=1000, check_every=10):
    """Solve ``A x = b`` (A symmetric positive definite, an ``Operator``) on the GPU.

    ``b`` / ``x0``: numpy arrays or CUDA tensors; ``inv_diag``: reciprocal of the Jacobi diagonal (CUDA tensor or numpy)
    or None for plain CG.  Returns ``(x, CGInfo)`` with ``x`` a CUDA tensor.  Stops when ||r|| <= rtol ||b|| (checked
    every ``check_every`` iterations, so up to ``check_every - 1`` extra iterations are run)."""
    lib, dev = load(), cuda_device()
    n = A.shape[0]
    if A.shape[0] != A.shape[1]:
        raise ValueError("pcg needs a square operator")
    b = to_device(b).reshape(-1)
    if b.numel() != n:
        raise ValueError(f"b must have {n} entries")
    x = torch.zeros(n, dtype=torch.float64, device=dev) if x0 is None else to_device(x0).reshape(-1).clone()
    dinv = None if inv_diag is None else to_device(inv_diag).reshape(-1)
    r = b.clone()
    Ap = torch.empty(n, dtype=torch.float64, device=dev)
    if x0 is not None:
        A.apply_device(x, out=Ap)
        r -= Ap
    z = r if dinv is None else dinv * r
    p = z.clone()
    scal = torch.zeros(6, dtype=torch.float64, device=dev)
    ws = torch.zeros(int(lib.dzb_solver_workspace_bytes()), dtype=torch.uint8, device=dev)
    s = stream_ptr()
    fused_dot = hasattr(A, "apply_with_dot")
    check(lib.dzb_dot(n, ptr(r), ptr(z), ptr(ws), ptr(scal), 0, s), "dzb_dot")          # scal[0] = r.z
    bnorm = float(torch.linalg.vector_norm(b))
    rnorm = float(torch.linalg.vector_norm(r))
    history = [rnorm]
    it = 0
    if bnorm == 0.0:
        return torch.zeros_like(b), CGInfo(0, 0.0, 0.0, True, history)
    while rnorm > rtol * bnorm and it < maxiter:
        for _ in range(min(check_every, maxiter - it)):
            # p.Ap: inside the apply where the operator offers it (K8: 16 B / node less per iteration), else a dot kernel
            if not (fused_dot and A.apply_with_dot(p, Ap, scal[1:2])):
                A.apply_device(p, out=Ap)
                check(lib.dzb_dot(n, ptr(p), ptr(Ap), ptr(ws), ptr(scal), 1, s), "dzb_dot")
            check(lib.dzb_cg_update(n, ptr(x), ptr(r), ptr(p), ptr(Ap), ptr(dinv), ptr(z), ptr(ws), ptr(scal), s),
                  "dzb_cg_update")
            check(lib.dzb_cg_direction(n, ptr(p), ptr(z), ptr(ws), ptr(scal), s), "dzb_cg_direction")
            it += 1
        rnorm = float(scal[3].sqrt())      # the only host read-back
        history.append(rnorm)
    return x, CGInfo(it, rnorm, bnorm, rnorm <= rtol * bnorm, history)

"""LinearOperator-compatible operators backed by CUDA kernels.

The reference returns ``scipy.sparse`` matrices from its mesh properties and applies
them with ``A @ x`` (scipy ``csr_matvec``).  Here every such property returns an
:class:`Operator`: ``isinstance(op, scipy.sparse.linalg.LinearOperator)`` holds,
``op @ x`` / ``op.T @ y`` / ``op.dot`` / ``op * x`` / ``matvec`` / ``rmatvec`` /
``matmat`` work on numpy arrays (one H2D + one D2H copy) or on torch CUDA float64
tensors (zero copy, the result stays on the device), ``A @ B``, ``A + B``,
``alpha * A`` and ``A.T`` compose lazily -- recognising ``G.T @ Me @ G`` and
``C.T @ Mf @ C + Me`` and replacing them with single fused kernels -- and
``op.tocsr()`` materialises the same sparsity and values as the reference.

No CPU fallback exists: applying an operator without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import LinearOperator

from . import _lib
from ._lib import DzbError

try:  # torch is plumbing: device memory, streams, torch.distributed
    import torch
except Exception as exc:  # pragma: no cover
    raise ImportError("discretize_b200 needs PyTorch for device memory and streams") from exc


# ----------------------------------------------------------------------------------------
# device plumbing
# ----------------------------------------------------------------------------------------
def cuda_device():
    if not torch.cuda.is_available():
        raise DzbError(
            "no CUDA device available: discretize_b200 operators only run on the GPU "
            "(there is deliberately no CPU fallback)"
        )
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def to_device(x, dtype=None):
    """numpy / torch -> contiguous CUDA tensor (float64 unless dtype is given)."""
    dtype = torch.float64 if dtype is None else dtype
    dev = cuda_device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    a = np.ascontiguousarray(x, dtype={torch.float64: np.float64, torch.int64: np.int64,
                                       torch.int32: np.int32, torch.uint8: np.uint8}[dtype])
    return torch.from_numpy(a).to(dev)


def _is_scalar(x):
    return np.isscalar(x) or (isinstance(x, (np.ndarray, torch.Tensor)) and x.ndim == 0)


def _scalar_value(x):
    """float for real scalars, complex for complex ones (1j * omega * Me is what SimPEG's Maxwell systems add)."""
    v = x.item() if isinstance(x, (np.ndarray, np.generic, torch.Tensor)) else x
    return complex(v) if isinstance(v, complex) and v.imag != 0.0 else float(v.real if isinstance(v, complex) else v)


# ----------------------------------------------------------------------------------------
# base class
# ----------------------------------------------------------------------------------------
class Operator(LinearOperator):
    """A float64 linear operator applied by CUDA kernels.

    Subclasses implement ``_dev_apply(x, out)`` and ``_dev_apply_t(x, out)`` on
    contiguous 1-D CUDA float64 tensors and (where the reference has a matrix)
    ``_tocsr()``.
    """

    #: complex-valued operators (a complex scalar times a real operator, and sums / products containing one) have no real
    #: kernel of their own: they implement ``_apply_complex`` on top of the real kernels of their parts
    is_complex = False

    def __init__(self, shape, is_complex=False):
        LinearOperator.__init__(self, np.dtype(np.complex128 if is_complex else np.float64), tuple(int(s) for s in shape))
        self.is_complex = bool(is_complex)

    def _apply_complex(self, x, transpose):  # pragma: no cover - only complex operators define it
        raise NotImplementedError

    # -- kernels -------------------------------------------------------------------------
    def _dev_apply(self, x, out):  # pragma: no cover - abstract
        raise NotImplementedError

    def _dev_apply_t(self, x, out):  # pragma: no cover - abstract
        raise NotImplementedError

    def _tocsr(self):
        raise NotImplementedError(f"{type(self).__name__} has no matrix form")

    # scipy's abstract hooks (solvers call matvec/rmatvec, which we override below)
    def _matvec(self, x):
        return self._apply_any(x, False)

    def _rmatvec(self, x):
        return self._apply_any(x, True)

    def _matmat(self, X):
        return self._apply_any(X, False)

    def _rmatmat(self, X):
        return self._apply_any(X, True)

    # -- dispatch on the input type ---------------------------------------------------------
    def apply_device(self, x, transpose=False, out=None):
        """y = A x (or A^T x) for a 1-D CUDA float64 tensor; result stays on the device."""
        n_out, n_in = (self.shape[1], self.shape[0]) if transpose else self.shape
        if x.numel() != n_in:
            raise ValueError(f"dimension mismatch: operator {self.shape}, transpose={transpose}, vector {tuple(x.shape)}")
        if out is None:
            out = torch.empty(n_out, dtype=torch.float64, device=x.device)
        (self._dev_apply_t if transpose else self._dev_apply)(x, out)
        return out

    @staticmethod
    def _check_out(out, n_out, x):
        """``out`` receives raw 8-byte writes: it must be float64, contiguous, of the right length and of x's kind."""
        if isinstance(out, torch.Tensor):
            if not isinstance(x, torch.Tensor) or out.device.type != x.device.type:
                raise TypeError("out must be the same kind of array as x (numpy / CPU tensor / CUDA tensor)")
            ok = out.dtype == torch.float64 and out.is_contiguous() and out.numel() == n_out
        elif isinstance(out, np.ndarray):
            if isinstance(x, torch.Tensor):
                raise TypeError("out must be the same kind of array as x (numpy / CPU tensor / CUDA tensor)")
            ok = out.dtype == np.float64 and out.flags.c_contiguous and out.size == n_out and out.flags.writeable
        else:
            raise TypeError("out must be a numpy array or a torch tensor")
        if not ok:
            raise ValueError(f"out must be a contiguous float64 array with {n_out} entries")

    def _apply_any(self, x, transpose, out=None):
        if self.is_complex:
            if out is not None:
                raise TypeError("out is not supported for complex operators (the result is a new complex array)")
            return self._apply_complex(x, transpose)
        n_in = self.shape[0] if transpose else self.shape[1]
        is_torch = isinstance(x, torch.Tensor)
        if is_torch and x.is_complex():
            return torch.complex(self._apply_any(x.real.contiguous(), transpose), self._apply_any(x.imag.contiguous(), transpose))
        if not is_torch:
            x = np.asarray(x)
            if np.iscomplexobj(x):
                if out is not None:
                    raise TypeError("out is not supported for complex input (the result is a new complex array)")
                return self._apply_any(x.real, transpose) + 1j * self._apply_any(x.imag, transpose)
        if out is not None:
            self._check_out(out, self.shape[1] if transpose else self.shape[0], x)
        if x.ndim == 1 or (x.ndim == 2 and x.shape[1] == 1 and x.shape[0] == n_in):
            if x.shape[0] != n_in:
                raise ValueError(f"dimension mismatch: operator {self.shape}, input {tuple(x.shape)}")
            if is_torch and x.device.type == "cpu":
                # host tensor in -> host tensor out; pinned buffers make both copies asynchronous DMA.
                # `out` (a preallocated, ideally pinned, CPU tensor) avoids a page-locked allocation per call.
                xd = x.to(device=cuda_device(), dtype=torch.float64, non_blocking=True).reshape(-1)
                yd = self.apply_device(xd, transpose)
                y = out if out is not None else torch.empty(yd.shape, dtype=torch.float64, pin_memory=x.is_pinned())
                y.reshape(-1).copy_(yd, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return y.reshape(-1, 1) if (x.ndim == 2 and out is None) else y
            xd = to_device(x).reshape(-1)
            if is_torch and out is not None:
                return self.apply_device(xd, transpose, out=out)
            yd = self.apply_device(xd, transpose)
            if is_torch:
                y = yd
            elif out is not None:
                torch.from_numpy(out.reshape(-1)).copy_(yd)
                return out
            else:
                y = yd.cpu().numpy()
            return y.reshape(-1, 1) if x.ndim == 2 else y
        if x.ndim == 2 and x.shape[0] == n_in:
            multi = getattr(self, "_dev_apply_t_multi" if transpose else "_dev_apply_multi", None)
            if multi is not None and x.shape[1] > 1:
                # multi-RHS kernels take the right-hand sides one after the other: (k, n) row-major = the columns of a
                # Fortran-ordered (n, k) block, which is passed through without a copy
                dev = cuda_device()
                if is_torch:
                    xt = x.to(device=dev, dtype=torch.float64).T
                else:
                    xt = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)).to(dev)
                xt = xt if xt.is_contiguous() else xt.contiguous()
                yt = torch.empty((xt.shape[0], self.shape[1] if transpose else self.shape[0]), dtype=torch.float64, device=dev)
                multi(xt, yt)
                return yt.T if is_torch else yt.cpu().numpy().T
            xd = to_device(x)
            cols = [self.apply_device(xd[:, c].contiguous(), transpose) for c in range(xd.shape[1])]
            yd = torch.stack(cols, dim=1) if cols else torch.empty(
                (self.shape[1] if transpose else self.shape[0], 0), dtype=torch.float64, device=xd.device)
            return yd if is_torch else yd.cpu().numpy()
        raise ValueError(f"dimension mismatch: operator {self.shape}, input {tuple(x.shape)}")

    # -- public LinearOperator surface ----------------------------------------------------------
    def matvec(self, x, out=None):
        """y = A x.  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out, CPU tensor in -> CPU tensor
        out.  ``out`` optionally receives the result (same kind as ``x``)."""
        return self._apply_any(x, False, out)

    def rmatvec(self, x, out=None):
        return self._apply_any(x, True, out)

    def matmat(self, X):
        return self._apply_any(X, False)

    def rmatmat(self, X):
        return self._apply_any(X, True)

    def dot(self, x):
        if isinstance(x, Operator):
            return ProductOperator.make([self, x])
        if sp.issparse(x):
            return ProductOperator.make([self, as_operator(x)])
        if _is_scalar(x):
            return ScaledOperator(self, _scalar_value(x))
        return self._apply_any(x, False)

    __mul__ = dot

    def __matmul__(self, x):
        if _is_scalar(x):
            raise ValueError("Scalar operands are not allowed, use '*' instead")
        return self.dot(x)

    def __call__(self, x):
        return self.dot(x)

    def __rmul__(self, x):
        if _is_scalar(x):
            return ScaledOperator(self, _scalar_value(x))
        return self.__rmatmul__(x)

    def __rmatmul__(self, x):
        if _is_scalar(x):
            raise ValueError("Scalar operands are not allowed, use '*' instead")
        if sp.issparse(x):
            return ProductOperator.make([as_operator(x), self])
        # x @ A == (A^T x^T)^T
        if isinstance(x, torch.Tensor):
            return self._apply_any(x if x.ndim == 1 else x.T, True) if x.ndim == 1 else self._apply_any(x.T, True).T
        x = np.asarray(x)
        return self._apply_any(x, True) if x.ndim == 1 else self._apply_any(x.T, True).T

    def __add__(self, other):
        if sp.issparse(other):
            other = as_operator(other)
        if isinstance(other, Operator):
            return SumOperator.make([self, other])
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        if sp.issparse(other):
            other = as_operator(other)
        if isinstance(other, Operator):
            return SumOperator.make([self, ScaledOperator(other, -1.0)])
        return NotImplemented

    def __neg__(self):
        return ScaledOperator(self, -1.0)

    def __truediv__(self, other):
        if not _is_scalar(other):
            raise ValueError("Can only divide a linear operator by a scalar.")
        return ScaledOperator(self, 1.0 / _scalar_value(other))

    def _transpose(self):
        return TransposedOperator(self)

    _adjoint = _transpose

    def transpose(self):
        return self._transpose()

    def adjoint(self):
        return self._adjoint()

    @property
    def T(self):
        return self._transpose()

    @property
    def H(self):
        return self._adjoint()        # == .T for the real operators; conjugate transpose for complex ones

    # -- matrix forms -----------------------------------------------------------------------------
    def tocsr(self):
        """The reference's matrix: same sparsity (canonical CSR: sorted, duplicates summed) and values."""
        return self._tocsr()

    def tocsc(self):
        return self.tocsr().tocsc()

    def tocoo(self):
        return self.tocsr().tocoo()

    def toarray(self):
        return self.tocsr().toarray()

    todense = toarray

    def diagonal(self):
        return self.tocsr().diagonal()

    @property
    def nnz(self):
        return self.tocsr().nnz

    def __repr__(self):
        return f"<{self.shape[0]}x{self.shape[1]} {type(self).__name__} (CUDA, float64)>"


def csr_from_device(indptr, indices, data, shape):
    """Device CSR arrays -> scipy.sparse.csr_matrix (int32 indices when they fit, like scipy)."""
    indptr = indptr.cpu().numpy()
    indices = indices.cpu().numpy()
    data = data.cpu().numpy()
    if max(shape[1], int(indptr[-1]) if indptr.size else 0) < 2**31 - 1:
        indptr = indptr.astype(np.int32)
        indices = indices.astype(np.int32)
    m = sp.csr_matrix((data, indices, indptr), shape=shape)
    m.has_sorted_indices = True
    return m


# ----------------------------------------------------------------------------------------
# algebra
# ----------------------------------------------------------------------------------------
class TransposedOperator(Operator):
    def __init__(self, base):
        super().__init__((base.shape[1], base.shape[0]), base.is_complex)
        self.base = base

    def _apply_complex(self, x, transpose):
        return self.base._apply_any(x, not transpose)

    def _dev_apply(self, x, out):
        self.base._dev_apply_t(x, out)

    def _dev_apply_t(self, x, out):
        self.base._dev_apply(x, out)

    def _transpose(self):
        return self.base

    def _adjoint(self):
        return self.base if not self.is_complex else TransposedOperator(self.base.H.T)

    def _tocsr(self):
        return self.base.tocsr().T.tocsr()


class DiagonalOperator(Operator):
    """diag(d) as an operator (one streaming kernel); ``d``: numpy array or CUDA tensor.  A numpy diagonal is uploaded on
    first use, so operators built from host data (1-D / 2-D meshes) can be composed and ``.tocsr()``-ed without a GPU."""

    def __init__(self, diag):
        if isinstance(diag, torch.Tensor):
            self._host, self._dev = None, to_device(diag).reshape(-1)
            n = self._dev.numel()
        else:
            self._host, self._dev = np.ascontiguousarray(np.asarray(diag, dtype=np.float64).reshape(-1)), None
            n = self._host.size
        super().__init__((n, n))

    @property
    def diag(self):
        if self._dev is None:
            self._dev = to_device(self._host).reshape(-1)
        return self._dev

    def _dev_apply(self, x, out):
        from ._lib import check, load

        check(load().dzb_diag_scale(self.shape[0], ptr(self.diag), ptr(x), ptr(out), 0, stream_ptr()), "dzb_diag_scale")

    _dev_apply_t = _dev_apply

    def _transpose(self):
        return self

    _adjoint = _transpose

    def diagonal(self):
        return self._host.copy() if self._host is not None else self._dev.cpu().numpy()

    def _tocsr(self):
        import scipy.sparse as sp

        return sp.diags(self.diagonal(), format="csr")


class ScaledOperator(Operator):
    """alpha * A; alpha real or complex (``1j * omega * Me``: the result is then a complex operator)."""

    def __init__(self, base, alpha):
        if isinstance(base, ScaledOperator):
            base, alpha = base.base, alpha * base.alpha
        alpha = _scalar_value(alpha)
        super().__init__(base.shape, base.is_complex or isinstance(alpha, complex))
        self.base, self.alpha = base, alpha

    def _dev_apply(self, x, out):
        self.base._dev_apply(x, out)
        if self.alpha != 1.0:
            out.mul_(self.alpha)

    def _dev_apply_t(self, x, out):
        self.base._dev_apply_t(x, out)
        if self.alpha != 1.0:
            out.mul_(self.alpha)

    def _apply_complex(self, x, transpose):
        return self.alpha * self.base._apply_any(x, transpose)     # .T is the plain transpose: alpha is not conjugated

    def _transpose(self):
        return ScaledOperator(self.base.T, self.alpha)

    def _adjoint(self):
        return ScaledOperator(self.base.H, self.alpha.conjugate() if isinstance(self.alpha, complex) else self.alpha)

    def _tocsr(self):
        return (self.base.tocsr() * self.alpha).tocsr()


class SumOperator(Operator):
    def __init__(self, terms):
        shape = terms[0].shape
        for t in terms:
            if t.shape != shape:
                raise ValueError(f"cannot add operators of shapes {shape} and {t.shape}")
        super().__init__(shape, any(t.is_complex for t in terms))
        self.terms = list(terms)

    def _apply_complex(self, x, transpose):
        acc = self.terms[0]._apply_any(x, transpose)
        for t in self.terms[1:]:
            acc = acc + t._apply_any(x, transpose)
        return acc

    @staticmethod
    def make(terms):
        flat = []
        for t in terms:
            flat.extend(t.terms if isinstance(t, SumOperator) else [t])
        from .fusion import fuse_sum

        fused = fuse_sum(flat)
        return fused if fused is not None else SumOperator(flat)

    def _scratch(self, like):
        """One temporary per operator and output size, kept between applies (no allocation on the apply path)."""
        t = getattr(self, "_tmp", None)
        if t is None or t.shape != like.shape or t.device != like.device:
            t = self._tmp = torch.empty_like(like)
        return t

    def _dev_apply(self, x, out):
        self.terms[0]._dev_apply(x, out)
        tmp = self._scratch(out)
        for t in self.terms[1:]:
            if isinstance(t, ScaledOperator):          # alpha * B: one fused multiply-add pass instead of scale + add
                t.base._dev_apply(x, tmp)
                out.add_(tmp, alpha=t.alpha)
            else:
                t._dev_apply(x, tmp)
                out.add_(tmp)

    def _dev_apply_t(self, x, out):
        self.terms[0]._dev_apply_t(x, out)
        tmp = self._scratch(out)
        for t in self.terms[1:]:
            if isinstance(t, ScaledOperator):
                t.base._dev_apply_t(x, tmp)
                out.add_(tmp, alpha=t.alpha)
            else:
                t._dev_apply_t(x, tmp)
                out.add_(tmp)

    def _transpose(self):
        return SumOperator.make([t.T for t in self.terms])

    def _adjoint(self):
        return SumOperator.make([t.H for t in self.terms])

    def _tocsr(self):
        acc = self.terms[0].tocsr()
        for t in self.terms[1:]:
            acc = acc + t.tocsr()
        return acc.tocsr()


class ProductOperator(Operator):
    """A1 @ A2 @ ... @ An, applied right to left with one temporary per factor."""

    def __init__(self, factors):
        for a, b in zip(factors[:-1], factors[1:]):
            if a.shape[1] != b.shape[0]:
                raise ValueError(f"cannot multiply operators of shapes {a.shape} and {b.shape}")
        super().__init__((factors[0].shape[0], factors[-1].shape[1]), any(f.is_complex for f in factors))
        self.factors = list(factors)

    def _apply_complex(self, x, transpose):
        cur = x
        for f in (self.factors if transpose else reversed(self.factors)):
            cur = f._apply_any(cur, transpose)
        return cur

    @staticmethod
    def make(factors):
        flat = []
        for f in factors:
            flat.extend(f.factors if isinstance(f, ProductOperator) else [f])
        for a, b in zip(flat[:-1], flat[1:]):
            if a.shape[1] != b.shape[0]:
                raise ValueError(f"cannot multiply operators of shapes {a.shape} and {b.shape}")
        from .fusion import fuse_product

        fused = fuse_product(flat)
        if fused is not None:
            return fused
        return flat[0] if len(flat) == 1 else ProductOperator(flat)

    def _dev_apply(self, x, out):
        cur = x
        for f in reversed(self.factors[1:]):
            cur = f.apply_device(cur, False)
        self.factors[0]._dev_apply(cur, out)

    def _dev_apply_t(self, x, out):
        cur = x
        for f in self.factors[:-1]:
            cur = f.apply_device(cur, True)
        self.factors[-1]._dev_apply_t(cur, out)

    def _transpose(self):
        return ProductOperator.make([f.T for f in reversed(self.factors)])

    def _adjoint(self):
        return ProductOperator.make([f.H for f in reversed(self.factors)])

    def _tocsr(self):
        acc = self.factors[0].tocsr()
        for f in self.factors[1:]:
            acc = acc @ f.tocsr()
        return acc.tocsr()


def as_operator(a):
    """Wrap a scipy.sparse matrix as a device CSR operator (generic SpMV kernel)."""
    if isinstance(a, Operator):
        return a
    if sp.issparse(a):
        from .csr import CsrOperator

        return CsrOperator.from_scipy(a)
    raise TypeError(f"cannot convert {type(a)} to an Operator")

"""Generic device CSR operator with stored values (dzb_csr_spmv).

Used for scipy.sparse matrices that users mix into operator algebra (projections, boundary
conditions) and for operators whose values are not generated on the fly (full-tensor OcTree mass
matrices)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch

from ._lib import check, load
from .operators import Operator, ptr, stream_ptr, to_device


class _DeviceCsr:
    def __init__(self, m):
        m = sp.csr_matrix(m)
        self.shape = m.shape
        self.nnz = int(m.nnz)
        self.is64 = self.nnz >= 2**31 - 1
        if m.shape[1] >= 2**31 - 1:
            raise ValueError("column indices beyond int32 are not supported by dzb_csr_spmv")
        self.indptr = to_device(m.indptr.astype(np.int64 if self.is64 else np.int32),
                                torch.int64 if self.is64 else torch.int32)
        self.indices = to_device(m.indices.astype(np.int32), torch.int32)
        self.data = to_device(m.data.astype(np.float64))
        avg = self.nnz / max(m.shape[0], 1)
        self.lpr = 1 if avg <= 7 else 2 if avg <= 16 else 4 if avg <= 40 else 16

    def apply(self, x, out):
        check(load().dzb_csr_spmv(self.shape[0], ptr(self.indptr), int(self.is64), ptr(self.indices), ptr(self.data),
                                  ptr(x), ptr(out), self.lpr, stream_ptr()), "dzb_csr_spmv")


class CsrOperator(Operator):
    def __init__(self, matrix, _t=None):
        self._host = sp.csr_matrix(matrix)
        super().__init__(self._host.shape)
        self._fwd = None
        self._bwd = None
        self._tr = _t

    @staticmethod
    def from_scipy(a):
        return CsrOperator(a)

    def _dev_apply(self, x, out):
        if self._fwd is None:
            self._fwd = _DeviceCsr(self._host)
        self._fwd.apply(x, out)

    def _dev_apply_t(self, x, out):
        if self._bwd is None:
            self._bwd = _DeviceCsr(self._host.T.tocsr())
        self._bwd.apply(x, out)

    def _tocsr(self):
        return self._host.copy()

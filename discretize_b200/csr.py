"""Generic device CSR operator (used for scipy matrices mixed into operator algebra and,
with on-the-fly values, for TreeMesh connectivity)."""
from __future__ import annotations

from .operators import Operator


class CsrOperator(Operator):
    @staticmethod
    def from_scipy(a):
        raise NotImplementedError("CsrOperator.from_scipy: generic device SpMV not built yet")

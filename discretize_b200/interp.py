"""get_interpolation_matrix as a CUDA operator (K10 locate+weights, K11 apply).

reference: base/base_tensor_mesh.py:684-790 -> utils/interpolation_utils.py:44-172 ->
_extensions/interputils_cython.pyx:42-182.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from ._lib import check, load
from .operators import Operator, cuda_device, ptr, stream_ptr, to_device


def _location_layout(mesh, location_type):
    """(tensor key, column offset, number of columns) of a location type (:729-769)."""
    lt = mesh._parse_location_type(location_type)
    nC = mesh.n_cells
    if lt in ("cell_centers", "nodes"):
        return lt, 0, (nC if lt == "cell_centers" else mesh.n_nodes)
    if lt in ("faces_x", "faces_y", "faces_z"):
        ind = "xyz".index(lt[-1])
        return lt, int(sum(mesh.n_faces_per_direction[:ind])), mesh.n_faces
    if lt in ("edges_x", "edges_y", "edges_z"):
        ind = "xyz".index(lt[-1])
        return lt, int(sum(mesh.n_edges_per_direction[:ind])), mesh.n_edges
    if lt in ("cell_centers_x", "cell_centers_y", "cell_centers_z"):
        return "cell_centers", "xyz".index(lt[-1]) * nC, 3 * nC
    raise NotImplementedError(
        "get_interpolation_matrix: location_type==" + lt + " and mesh.dim==" + str(mesh.dim)
    )


class InterpolationOperator(Operator):
    """Q (npts x ncols), <= 8 entries per row, stored as COMPACT rows on the GPU: the flat index of the first grid
    sample plus three per-axis weights (28 bytes per point; K10c / K11c in csrc/interp.cu).  `cols` / `w` expand them to
    the (npts, 8) tables the reference's COO triplets hold (interputils_cython.pyx:145-180), bit for bit."""

    def __init__(self, mesh, pts_dev, location_type, zeros_outside, row_scale, store=True):
        key, off, ncols = _location_layout(mesh, location_type)
        npts = pts_dev.shape[0]
        super().__init__((npts, ncols))
        self.mesh, self.col_offset, self.row_scale = mesh, off, row_scale
        self.pts = pts_dev
        self.axes = [to_device(t) for t in mesh.get_tensor(key)]
        self.base = self.wx = self.wy = self.wz = None
        self.cols_is64 = ncols > 2**31 - 1      # int32 indices while they fit (like scipy's index arrays)
        # Bucketing (K11): before the first apply of a LARGE operator the rows are sorted once by the flat index of
        # their first grid sample; the gathers then stream through the grid vector instead of paying a 128-byte DRAM
        # fetch per random 8-byte value.  perm[p] = the point row p belongs to.  Done lazily so that building Q stays a
        # single locate pass (BASELINE config 4 times the build), and only where it pays (the sort costs ~10 applies).
        self.perm = None
        self._scale_rows = row_scale            # row_scale in table order
        self.bucket_min_points = 1 << 20
        self.unpermute = "gather"             # or "scatter": the transposed kernel's order, y[perm[p]] by index_copy
        if store:
            self._locate()

    def _axes_args(self):
        a = self.axes
        return (ptr(a[0]), a[0].numel(), ptr(a[1]), a[1].numel(), ptr(a[2]), a[2].numel())

    def _row_args(self):
        return (ptr(self.base), int(self.cols_is64), ptr(self.wx), ptr(self.wy), ptr(self.wz),
                self.axes[0].numel(), self.axes[1].numel())

    def _locate(self):
        npts = self.shape[0]
        dev = cuda_device()
        self.base = torch.empty(npts, dtype=torch.int64 if self.cols_is64 else torch.int32, device=dev)
        self.wx, self.wy, self.wz = (torch.empty(npts, dtype=torch.float64, device=dev) for _ in range(3))
        self.perm, self._scale_rows = None, self.row_scale          # fresh rows are in point order
        check(load().dzb_interp_locate_compact(*self._axes_args(), ptr(self.pts), npts, self.col_offset, ptr(self.base),
                                               int(self.cols_is64), ptr(self.wx), ptr(self.wy), ptr(self.wz),
                                               stream_ptr()), "dzb_interp_locate_compact")

    def tables(self):
        """(cols, w): the (npts, 8) index / weight tables in TABLE order (row p belongs to point perm[p] once bucketed)."""
        if self.base is None:
            self._locate()
        npts = self.shape[0]
        cols = torch.empty((npts, 8), dtype=self.base.dtype, device=self.base.device)
        w = torch.empty((npts, 8), dtype=torch.float64, device=self.base.device)
        check(load().dzb_interp_expand(*self._row_args(), npts, ptr(cols), int(self.cols_is64), ptr(w), stream_ptr()),
              "dzb_interp_expand")
        return cols, w

    @property
    def cols(self):
        return self.tables()[0]

    @property
    def w(self):
        return self.tables()[1]

    def bucket(self):
        """Sort the rows by their first column (x fastest, like the grid vector) and remember the permutation."""
        if self.perm is not None or self.base is None or self.shape[0] >= 2**31 - 1:
            return
        order = torch.argsort(self.base)
        self.base = self.base.index_select(0, order)
        self.wx, self.wy, self.wz = (t.index_select(0, order) for t in (self.wx, self.wy, self.wz))
        if self.row_scale is not None:
            self._scale_rows = self.row_scale.index_select(0, order)
        self.perm = order.to(torch.int32)
        self.inv_perm = torch.empty_like(self.perm)
        self.inv_perm[order] = torch.arange(order.numel(), dtype=torch.int32, device=order.device)
        self._sorted_out = torch.empty(order.numel(), dtype=torch.float64, device=order.device)

    def _maybe_bucket(self):
        if self.perm is None and self.shape[0] >= self.bucket_min_points:
            self.bucket()

    def _dev_apply(self, x, out):
        npts = self.shape[0]
        if self.base is None:  # fused locate + apply, nothing stored
            check(load().dzb_interp_fused(*self._axes_args(), ptr(self.pts), npts, self.col_offset,
                                          ptr(self.row_scale), ptr(x), ptr(out), stream_ptr()), "dzb_interp_fused")
            return
        self._maybe_bucket()
        # bucketed rows: the result comes out in bucket order (coalesced), then one pass of random READS brings it back
        # to point order (a scatter costs a sector fill plus a write-back per value)
        dst = out if self.perm is None else self._sorted_out
        check(load().dzb_interp_apply_compact(*self._row_args(), ptr(self._scale_rows), npts, ptr(x), ptr(dst),
                                              stream_ptr()), "dzb_interp_apply_compact")
        if self.perm is not None:
            if self.unpermute == "scatter":
                out.index_copy_(0, self.perm.long(), dst)
            else:
                check(load().dzb_gather_rows(npts, ptr(dst), ptr(self.inv_perm), ptr(out), stream_ptr()), "dzb_gather_rows")

    def _dev_apply_t(self, x, out):
        if self.base is None:
            self._locate()
        self._maybe_bucket()
        out.zero_()
        check(load().dzb_interp_apply_t_compact(*self._row_args(), ptr(self._scale_rows), ptr(self.perm), self.shape[0],
                                                ptr(x), ptr(out), stream_ptr()), "dzb_interp_apply_t_compact")

    def _tocsr(self):
        """COO -> CSR exactly like the reference: duplicate columns summed, sorted, explicit zeros kept;
        zeros_outside rows become full rows of explicit zeros (base_tensor_mesh.py:771-772)."""
        npts, ncols = self.shape
        cols, w = self.tables()
        cols = cols.cpu().numpy().reshape(-1).astype(np.int64)
        w = w.cpu().numpy().reshape(-1)
        # bucketed tables: row p of the table is row perm[p] of the matrix
        rows = np.repeat(np.arange(npts) if self.perm is None else self.perm.cpu().numpy().astype(np.int64), 8)
        Q = sp.csr_matrix((w, (rows, cols)), shape=(npts, ncols))
        Q.sum_duplicates()
        Q.sort_indices()
        if self.row_scale is not None:
            outside = self.row_scale.cpu().numpy() == 0.0
            if outside.any():
                if outside.sum() * ncols > 5e8:
                    raise MemoryError("zeros_outside rows are dense in the reference's matrix; too large to materialise")
                counts = np.diff(Q.indptr).astype(np.int64)
                counts[outside] = ncols
                indptr = np.r_[0, np.cumsum(counts)]
                indices = np.empty(indptr[-1], dtype=np.int64)
                data = np.zeros(indptr[-1])
                for r in range(npts):
                    lo, hi = indptr[r], indptr[r + 1]
                    if outside[r]:
                        indices[lo:hi] = np.arange(ncols)
                    else:
                        indices[lo:hi] = Q.indices[Q.indptr[r]:Q.indptr[r + 1]]
                        data[lo:hi] = Q.data[Q.indptr[r]:Q.indptr[r + 1]]
                Q = sp.csr_matrix((data, indices, indptr), shape=(npts, ncols))
        return Q


def make_interpolation_operator(mesh, loc, location_type="cell_centers", zeros_outside=False, store=True):
    from .tensor_mesh import as_array_n_by_dim

    if isinstance(loc, torch.Tensor):
        pts = to_device(loc)
        if pts.ndim == 1:
            pts = pts.reshape(1, -1)
        if pts.shape[1] != mesh.dim:
            raise ValueError("pts must be a column vector of shape (nPts, {0:d})".format(mesh.dim))
    else:
        pts = to_device(as_array_n_by_dim(loc, mesh.dim).astype(np.float64))
    npts = pts.shape[0]
    # is_inside is always evaluated against the nodes tensor (base_tensor_mesh.py:638, 720-723)
    nodes = mesh.get_tensor("nodes")
    lo = (C.c_double * 3)(*[float(t.min()) for t in nodes])
    hi = (C.c_double * 3)(*[float(t.max()) for t in nodes])
    tol = (C.c_double * 3)(*[float(np.diff(t).min() * 1.0e-10) for t in nodes])
    inside = torch.empty(npts, dtype=torch.uint8, device=pts.device)
    check(load().dzb_points_inside(ptr(pts), npts, lo, hi, tol, ptr(inside), stream_ptr()), "dzb_points_inside")
    all_inside = bool(inside.all().item()) if npts else True
    row_scale = None
    if not zeros_outside:
        if not all_inside:
            raise ValueError("Points outside of mesh")
    elif not all_inside:
        # outside points are moved to the mesh centre and their rows zeroed (:723-725, 771-772)
        center = torch.tensor([float(v.mean()) for v in mesh.get_tensor("CC")], dtype=torch.float64, device=pts.device)
        mask = inside.to(torch.bool)
        pts = torch.where(mask[:, None], pts, center[None, :]).contiguous()
        row_scale = mask.to(torch.float64)
    return InterpolationOperator(mesh, pts, location_type, zeros_outside, row_scale, store=store)

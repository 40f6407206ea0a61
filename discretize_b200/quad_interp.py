"""QuadTree interpolation from nodes, edges, faces and cell centres (host assembly, applied by ``dzb_csr_spmv``).

Reference: the ``dim == 2`` branches of ``_extensions/tree_ext.pyx`` -- ``_getNodeIntMat`` (:6116-6179),
``_getEdgeIntMat`` (:5765-5905), ``_getFaceIntMat`` (:5907-6114), ``_getCellIntMat`` (:6181-6404).  The scheme is the one
of ``tree_interp.py`` with four entities per point; x-faces are y-edges and y-faces are x-edges.  The cell-centre variant
keeps its two search axes ordered by distance to the cell centre FROM POINT TO POINT (ties keep the previous order), which
is a forward fill over the points; the 2-D face variant has no such state.  ``zeros_outside`` as in ``tree_interp.py``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .tree_build import key3
from .tree_interp import _clip01, _extend, _pair_along, _ratio

_EPS = 100 * np.finfo(float).eps
_UNIT = np.eye(3, dtype=np.int64)


class _Neighbours:
    """Same-level leaf neighbour of a leaf, or -1."""

    def __init__(self, topo):
        self.t = topo
        keys = key3(topo.centers[:, 0], topo.centers[:, 1], topo.centers[:, 2])
        self.order = np.argsort(keys, kind="stable")
        self.keys = keys[self.order]
        self.table = None

    def _lookup(self, cells, axes, sign):
        t = self.t
        axes = np.broadcast_to(np.asarray(axes, dtype=np.int64), cells.shape)
        q = t.centers[cells] + (sign * 2 * t.half[cells])[:, None] * _UNIT[axes]
        inside = np.ones(cells.size, dtype=bool)
        for d in range(2):
            inside &= (q[:, d] > 0) & (q[:, d] < t.shape_half[d])
        q = np.where(inside[:, None], q, t.centers[cells])
        k = key3(q[:, 0], q[:, 1], q[:, 2])
        pos = np.minimum(np.searchsorted(self.keys, k), self.keys.size - 1)
        nb = self.order[pos]
        ok = inside & (self.keys[pos] == k) & (t.half[nb] == t.half[cells])
        return np.where(ok, nb, -1)

    def step(self, cells, axes, sign):
        """Neighbour of each cell along its axis (int or per-cell array) on side ``sign``: a table built once per mesh."""
        if self.table is None:
            everyone = np.arange(self.t.n_cells)
            self.table = np.stack([self._lookup(everyone, a, s) for a in range(2) for s in (-1, 1)], axis=1)
        axes = np.broadcast_to(np.asarray(axes, dtype=np.int64), cells.shape)
        return self.table[cells, 2 * axes + (1 if sign > 0 else 0)]


def _neighbours(mesh):
    if "interp_nb" not in mesh._cache:
        mesh._cache["interp_nb"] = _Neighbours(mesh._t())
    return mesh._cache["interp_nb"]


def _weights4(w1, w2):
    w1, w2 = _clip01(w1), _clip01(w2)
    return np.stack([w1 * w2, (1 - w1) * w2, w1 * (1 - w2), (1 - w1) * (1 - w2)], axis=1)


def _outside(mesh, cell, pts):
    t = mesh._t()
    out = np.zeros(pts.shape[0], dtype=bool)
    for d in range(2):
        lo = mesh._xs[d][t.centers[cell, d] - t.half[cell]]
        hi = mesh._xs[d][t.centers[cell, d] + t.half[cell]]
        out |= (pts[:, d] < lo - _EPS) | (pts[:, d] > hi + _EPS)
    return out


def _rows_csr(cols, vals, ncols, outside):
    n = cols.shape[0]
    if outside is not None:
        cols = np.where(outside[:, None], 0, cols)
        vals = np.where(outside[:, None], 0.0, vals)
    indptr = 4 * np.arange(n + 1, dtype=np.int64)
    return sp.csr_matrix((vals.reshape(-1), cols.reshape(-1).astype(np.int64), indptr), shape=(n, ncols))


def node_interpolation(mesh, pts, zeros_outside):
    t = mesh._t()
    cell = mesh._containing_cells(pts)
    cen, hw = t.centers[cell], t.half[cell]
    w = []
    for d in range(2):
        x0, x1 = mesh._xs[d][cen[:, d] - hw], mesh._xs[d][cen[:, d] + hw]
        w.append((x1 - pts[:, d]) / (x1 - x0))
    outside = None
    if zeros_outside:
        outside = np.zeros(pts.shape[0], dtype=bool)
        for wd in w:
            outside |= (wd < -_EPS) | (wd > 1 + _EPS)
    cols, vals = t.nodes.index[t.cell_nodes[cell]], _weights4(w[0], w[1])
    if outside is not None:
        cols = np.where(outside[:, None], 0, cols)
        vals = np.where(outside[:, None], 0.0, vals)
    # the reference goes through COO here, which sorts each row before the product with the deflation matrix
    rows = np.repeat(np.arange(pts.shape[0]), 4)
    A = sp.csr_matrix((vals.reshape(-1), (rows, cols.reshape(-1))), shape=(pts.shape[0], t.nodes.n_total))
    return (A @ mesh._deflate_nodes()).tocsr()


def _edge_like(mesh, pts, zeros_outside, d, as_faces):
    """Entries from the ``d``-edges of the containing cell and of its same-level neighbour along ``d``: linear across the
    edges (axis ``1 - d``) and, when there is a neighbour, along them.  The edge variant for direction ``d`` and the
    face variant for the normal ``1 - d`` (those faces ARE the ``d``-edges) differ only in where the block sits in the
    row and in the deflation matrix."""
    t = mesh._t()
    e = t.edges[d]
    o = 1 - d
    cell = mesh._containing_cells(pts)
    ccp = mesh.cell_centers[cell]
    eloc = mesh._midpoints(e.loc, e.half, (d,))  # by sorted position
    ce = t.cell_edges[d]
    i0, i1 = _pair_along(_neighbours(mesh), cell, np.full(cell.size, d), pts, ccp)
    e00, e01, e10, e11 = ce[i0, 0], ce[i0, 1], ce[i1, 0], ce[i1, 1]
    w1 = _ratio(eloc[e01, o], pts[:, o], eloc[e00, o])
    w2 = _ratio(eloc[e10, d], pts[:, d], eloc[e00, d], i0 != i1)
    if as_faces:      # faces = [x-faces (y-edges) | y-faces (x-edges)]
        off, R = (0 if d == 1 else t.edges[1].n_total), mesh._deflate_faces()
    else:             # edges = [x-edges | y-edges]
        off, R = (0 if d == 0 else t.edges[0].n_total), mesh._deflate_edges()
    pos = np.stack([e00, e01, e10, e11], axis=1)
    A = _rows_csr(e.index[pos] + off, _weights4(w1, w2), t.edges[0].n_total + t.edges[1].n_total,
                  _outside(mesh, cell, pts) if zeros_outside else None)
    return (A @ R).tocsr()


def edge_interpolation(mesh, pts, zeros_outside, direction):
    return _edge_like(mesh, pts, zeros_outside, "xy".index(direction), False)


def face_interpolation(mesh, pts, zeros_outside, direction):
    return _edge_like(mesh, pts, zeros_outside, 1 - "xy".index(direction), True)


def cell_interpolation(mesh, pts, zeros_outside):
    t = mesh._t()
    cell = mesh._containing_cells(pts)
    cc = mesh.cell_centers
    ccp = cc[cell]
    n = np.arange(cell.size)
    a, b = np.abs(pts[:, 0] - ccp[:, 0]), np.abs(pts[:, 1] - ccp[:, 1])
    decided = a != b
    last = np.maximum.accumulate(np.where(decided, n, -1))
    swapped = np.where(last >= 0, (a < b)[np.maximum(last, 0)], False)
    d0 = np.where(swapped, 1, 0)
    d1 = 1 - d0
    nb = _neighbours(mesh)
    i00, i01 = _pair_along(nb, cell, d0, pts, ccp)
    (i00, i01), (i10, i11) = _extend(nb, [i00, i01], d1, pts, ccp)
    w1 = _ratio(cc[i01][n, d0], pts[n, d0], cc[i00][n, d0], i01 != i00)
    w2 = _ratio(cc[i10][n, d1], pts[n, d1], cc[i00][n, d1], i10 != i00)
    cols = np.stack([i00, i01, i10, i11], axis=1)
    return _rows_csr(cols, _weights4(w1, w2), t.n_cells, _outside(mesh, cell, pts) if zeros_outside else None)

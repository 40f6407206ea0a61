"""TreeMesh interpolation from edges, faces and cell centres (host assembly, applied by ``dzb_csr_spmv``).

Reference: ``_extensions/tree_ext.pyx`` -- ``_getEdgeIntMat`` (:5765-5905), ``_getFaceIntMat`` (:5907-6114),
``_getCellIntMat`` (:6181-6404).  Per point the reference finds the containing leaf, steps to SAME-LEVEL leaf neighbours
on the side the point lies (one step per axis, each step only if every cell gathered so far has such a neighbour) and
interpolates linearly between the 8 entities so collected; total (hanging) entities are then eliminated by the
deflation matrices.  Two behaviours of the reference are data dependent and are reproduced here:

* the face and cell variants reorder the search axes by distance to the cell centre and KEEP that order for the next
  point (the swap is applied to loop-persistent variables), so a row depends on the points before it.  The order is a
  state machine (2 states for faces, 6 permutations for cells) driven by per-point comparisons; it is evaluated for all
  points at once as a prefix composition of per-point state maps (log2(n) vectorised passes) instead of a loop;
* the third swap of the cell variant assigns ``dir0`` twice and therefore changes nothing.

With ``zeros_outside=True`` the reference zeroes the rows of outside points and leaves every other row UNINITIALISED
(the weights are only computed in the ``else`` branch).  There is nothing to reproduce; here inside points get their
regular weights and outside points a zero row.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .tree_build import key3

_EPS = 100 * np.finfo(float).eps
_UNIT = np.eye(3, dtype=np.int64)
_PERMS = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]


class _Neighbours:
    """Same-level leaf neighbour of a leaf, or -1 (what ``cell.neighbors[k] != NULL and is_leaf and same level`` selects)."""

    def __init__(self, topo):
        self.t = topo
        keys = key3(topo.centers[:, 0], topo.centers[:, 1], topo.centers[:, 2])
        self.order = np.argsort(keys, kind="stable")
        self.keys = keys[self.order]
        self.table = None

    def _lookup(self, cells, axes, sign):
        """cells: (n,) leaf indices; axes: int or (n,) ints; sign: -1 / +1."""
        t = self.t
        axes = np.broadcast_to(np.asarray(axes, dtype=np.int64), cells.shape)
        q = t.centers[cells] + (sign * 2 * t.half[cells])[:, None] * _UNIT[axes]
        inside = np.ones(cells.size, dtype=bool)
        for d in range(3):
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
            self.table = np.stack([self._lookup(everyone, a, s) for a in range(3) for s in (-1, 1)], axis=1)
        axes = np.broadcast_to(np.asarray(axes, dtype=np.int64), cells.shape)
        return self.table[cells, 2 * axes + (1 if sign > 0 else 0)]


def _clip01(w):
    return np.minimum(np.maximum(w, 0.0), 1.0)


def _weights8(w1, w2, w3):
    w1, w2, w3 = _clip01(w1), _clip01(w2), _clip01(w3)
    return np.stack([w1 * w2 * w3, (1 - w1) * w2 * w3, w1 * (1 - w2) * w3, (1 - w1) * (1 - w2) * w3,
                     w1 * w2 * (1 - w3), (1 - w1) * w2 * (1 - w3), w1 * (1 - w2) * (1 - w3),
                     (1 - w1) * (1 - w2) * (1 - w3)], axis=1)


def _outside(mesh, cell, pts):
    t = mesh._t()
    out = np.zeros(pts.shape[0], dtype=bool)
    for d in range(3):
        lo = mesh._xs[d][t.centers[cell, d] - t.half[cell]]
        hi = mesh._xs[d][t.centers[cell, d] + t.half[cell]]
        out |= (pts[:, d] < lo - _EPS) | (pts[:, d] > hi + _EPS)
    return out


def _rows_csr(cols, vals, ncols, outside):
    n = cols.shape[0]
    if outside is not None:
        cols = np.where(outside[:, None], 0, cols)
        vals = np.where(outside[:, None], 0.0, vals)
    indptr = 8 * np.arange(n + 1, dtype=np.int64)
    return sp.csr_matrix((vals.reshape(-1), cols.reshape(-1).astype(np.int64), indptr), shape=(n, ncols))


def _pair_along(nb, cell, axes, pts, cc):
    """(low, high) cells of the one-axis step away from ``cell`` toward the point (both ``cell`` if no step)."""
    n = np.arange(cell.size)
    below = pts[n, axes] < cc[n, axes]
    above = pts[n, axes] > cc[n, axes]
    lo_nb = nb.step(cell, axes, -1)
    hi_nb = nb.step(cell, axes, +1)
    take_lo = below & (lo_nb >= 0)
    take_hi = ~take_lo & above & (hi_nb >= 0)
    first = np.where(take_lo, lo_nb, cell)
    second = np.where(take_hi, hi_nb, cell)
    return first, second


def _extend(nb, group, axes, pts, cc):
    """Step a group of cells together along ``axes`` toward the point; all of them must have the neighbour.
    Returns (near group, far group) ordered low -> high like the reference's i0.. / i1.. sets."""
    n = np.arange(group[0].size)
    below = pts[n, axes] < cc[n, axes]
    above = pts[n, axes] > cc[n, axes]
    lo = [nb.step(g, axes, -1) for g in group]
    hi = [nb.step(g, axes, +1) for g in group]
    take_lo = below & np.logical_and.reduce([x >= 0 for x in lo])
    take_hi = ~take_lo & above & np.logical_and.reduce([x >= 0 for x in hi])
    first = [np.where(take_lo, l, g) for l, g in zip(lo, group)]
    second = [np.where(take_hi, h, g) for h, g in zip(hi, group)]
    return first, second


def _ratio(hi, x, lo, differ=None):
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (hi - x) / (hi - lo)
    return w if differ is None else np.where(differ, w, 1.0)


def _prefix_states(maps):
    """maps: (n, S) next-state tables, one per point.  Returns the state AFTER each point, starting from state 0."""
    p = maps.copy()
    shift = 1
    n = p.shape[0]
    while shift < n:
        p[shift:] = np.take_along_axis(p[shift:], p[:-shift], axis=1)
        shift *= 2
    return p[:, 0]


# ---- edges ---------------------------------------------------------------------------------------------------------


def edge_interpolation(mesh, pts, zeros_outside, direction):
    t = mesh._t()
    d = "xyz".index(direction)
    d1, d2 = [a for a in range(3) if a != d]
    cell = mesh._containing_cells(pts)
    cc = mesh.cell_centers
    nb = _neighbours(mesh)
    i0, i1 = _pair_along(nb, cell, np.full(cell.size, d), pts, cc[cell])
    e = t.edges[d]
    eloc = mesh._midpoints(e.loc, e.half, (d,))                 # by sorted position
    E0, E1 = t.cell_edges[d][i0], t.cell_edges[d][i1]
    w1 = _ratio(eloc[E0[:, 1], d1], pts[:, d1], eloc[E0[:, 0], d1])
    w2 = _ratio(eloc[E1[:, 0], d], pts[:, d], eloc[E0[:, 0], d], i0 != i1)
    w3 = _ratio(eloc[E0[:, 2], d2], pts[:, d2], eloc[E0[:, 0], d2])
    pos = np.stack([E0[:, 0], E0[:, 1], E1[:, 0], E1[:, 1], E0[:, 2], E0[:, 3], E1[:, 2], E1[:, 3]], axis=1)
    off = int(sum(f.n_total for f in t.edges[:d]))
    n_total = int(sum(f.n_total for f in t.edges))
    A = _rows_csr(e.index[pos] + off, _weights8(w1, w2, w3), n_total, _outside(mesh, cell, pts) if zeros_outside else None)
    return (A @ mesh._deflate_edges()).tocsr()


# ---- faces ---------------------------------------------------------------------------------------------------------


def face_interpolation(mesh, pts, zeros_outside, direction):
    t = mesh._t()
    d = "xyz".index(direction)
    o1, o2 = [a for a in range(3) if a != d]
    cell = mesh._containing_cells(pts)
    cc = mesh.cell_centers
    ccp = cc[cell]
    n = np.arange(cell.size)
    # search order: the tangential axis with the larger distance to the centre first; ties keep the previous order
    a, b = np.abs(pts[:, o1] - ccp[:, o1]), np.abs(pts[:, o2] - ccp[:, o2])
    decided = a != b
    last = np.maximum.accumulate(np.where(decided, n, -1))
    swapped = np.where(last >= 0, (a < b)[np.maximum(last, 0)], False)
    d1 = np.where(swapped, o2, o1)
    d2 = np.where(swapped, o1, o2)
    nb = _neighbours(mesh)
    i00, i01 = _pair_along(nb, cell, d1, pts, ccp)
    (i00, i01), (i10, i11) = _extend(nb, [i00, i01], d2, pts, ccp)
    f = t.faces[d]
    floc = mesh._face_locations(f, (o1, o2))                    # by sorted position
    cf = t.cell_faces[d]
    f000, f001, f010, f011 = cf[i00, 0], cf[i00, 1], cf[i01, 0], cf[i01, 1]
    f100, f101, f110, f111 = cf[i10, 0], cf[i10, 1], cf[i11, 0], cf[i11, 1]
    w1 = _ratio(floc[f001, d], pts[:, d], floc[f000, d])
    w2 = _ratio(floc[f010, d1], pts[n, d1], floc[f000, d1], i00 != i01)
    w3 = _ratio(floc[f100, d2], pts[n, d2], floc[f000, d2], i10 != i00)
    pos = np.stack([f000, f001, f010, f011, f100, f101, f110, f111], axis=1)
    off = int(sum(x.n_total for x in t.faces[:d]))
    n_total = int(sum(x.n_total for x in t.faces))
    A = _rows_csr(f.index[pos] + off, _weights8(w1, w2, w3), n_total, _outside(mesh, cell, pts) if zeros_outside else None)
    return (A @ mesh._deflate_faces()).tocsr()


# ---- cell centres ----------------------------------------------------------------------------------------------------


def cell_interpolation(mesh, pts, zeros_outside):
    t = mesh._t()
    cell = mesh._containing_cells(pts)
    cc = mesh.cell_centers
    ccp = cc[cell]
    npts = cell.size
    n = np.arange(npts)
    dist = np.abs(pts - ccp)
    # per point and per incoming axis order: (swap 0<->1 if nearer, then swap 1<->2 if nearer) -> outgoing order
    perms = np.array(_PERMS)
    index_of = {p: i for i, p in enumerate(_PERMS)}
    maps = np.empty((npts, 6), dtype=np.int64)
    for s, (p0, p1, p2) in enumerate(_PERMS):
        q0 = np.full(npts, p0)
        q1 = np.full(npts, p1)
        q2 = np.full(npts, p2)
        sw = dist[n, q0] < dist[n, q1]
        q0, q1 = np.where(sw, q1, q0), np.where(sw, q0, q1)
        sw = dist[n, q1] < dist[n, q2]
        q1, q2 = np.where(sw, q2, q1), np.where(sw, q1, q2)
        code = q0 * 9 + q1 * 3 + q2
        lut = np.full(27, -1, dtype=np.int64)
        for p, i in index_of.items():
            lut[p[0] * 9 + p[1] * 3 + p[2]] = i
        maps[:, s] = lut[code]
    state = _prefix_states(maps) if npts else np.zeros(0, dtype=np.int64)
    d0, d1, d2 = perms[state, 0], perms[state, 1], perms[state, 2]
    nb = _neighbours(mesh)
    i000, i001 = _pair_along(nb, cell, d0, pts, ccp)
    (i000, i001), (i010, i011) = _extend(nb, [i000, i001], d1, pts, ccp)
    (i000, i001, i010, i011), (i100, i101, i110, i111) = _extend(nb, [i000, i001, i010, i011], d2, pts, ccp)
    w1 = _ratio(cc[i001][n, d0], pts[n, d0], cc[i000][n, d0], i001 != i000)
    w2 = _ratio(cc[i010][n, d1], pts[n, d1], cc[i000][n, d1], i010 != i000)
    w3 = _ratio(cc[i100][n, d2], pts[n, d2], cc[i000][n, d2], i100 != i000)
    cols = np.stack([i000, i001, i010, i011, i100, i101, i110, i111], axis=1)
    return _rows_csr(cols, _weights8(w1, w2, w3), t.n_cells, _outside(mesh, cell, pts) if zeros_outside else None)


def _neighbours(mesh):
    if "interp_nb" not in mesh._cache:
        mesh._cache["interp_nb"] = _Neighbours(mesh._t())
    return mesh._cache["interp_nb"]

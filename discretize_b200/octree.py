"""Level-set octree: refinement requests and 2:1 balancing on sorted integer arrays.

The reference grows a pointer tree by recursive ``Cell::divide`` calls that balance as they go
(``discretize/_extensions/tree.cpp:459-746``).  This is a different design with the same result.
The tree is stored as, per level, the sorted set of DIVIDED cells (packed integer coordinates at
that level); leaves are derived on demand as "children of divided cells that are not divided".

* A refinement request (a point to insert, a ball, ...) is pushed down from the roots level by
  level as a list of (cell, request) pairs; every cell a request still concerns above its target
  level joins the divided set.  This is exactly the set of cells the reference would divide for
  that request (``Cell::insert_cell`` tree.cpp:383-396, ``refine_geom`` tree.h:149-167).
* Balancing is one sweep from the finest level to the coarsest: a divided cell at level l has
  children at level l+1, so every (face-, or with ``diagonal`` also edge-/corner-) neighbouring box
  of it at level l must exist as a cell, i.e. the PARENT of each neighbouring box joins the divided
  set of level l-1.  Additions at level l-1 only create requirements at level l-2, so a single
  sweep reaches the unique minimal balanced closure -- the tree the reference's ripple of
  ``divide`` calls arrives at, independent of request order.
* Cells are numbered in the reference's order: roots x-fastest, then depth-first with children
  ordered by (x, y, z) bit (tree.cpp:232-344, 748-756, 949-953), i.e. Z-order within a root.
"""
from __future__ import annotations

import numpy as np


def _spread3(v, nbits):
    out = np.zeros_like(v)
    for b in range(nbits):
        out |= ((v >> b) & 1) << (3 * b)
    return out


class LinearOctree:
    def __init__(self, shape_cells):
        n = tuple(int(s) for s in shape_cells)
        if len(n) != 3:
            raise NotImplementedError("only 3-D OcTrees are built (QuadTree: not part of the five configurations)")
        for s in n:
            if s <= 0 or (s & (s - 1)) != 0:
                raise ValueError("length of cell width vectors must be a power of 2")
        self.n = n
        ls = [s.bit_length() - 1 for s in n]
        min_l = min(ls)
        cubic = ls[0] == ls[1] == ls[2]
        self.max_level = min_l + (0 if cubic else 1)      # tree.cpp:789-793
        self.root_level = 0 if cubic else 1               # tree.cpp:874-878
        self.rw = 1 << min_l                              # root width in fine cells
        self.roots = tuple(s // self.rw for s in n)
        self.nbits = min_l
        # divided cells per level (packed coordinates, sorted unique); levels root_level .. max_level-1
        self.div = {l: np.zeros(0, dtype=np.int64) for l in range(self.root_level, self.max_level)}
        self._leaves = None

    # ---- per-level grids ---------------------------------------------------------------------------
    def level_width(self, l):
        return 1 << (self.max_level - l)

    def grid(self, l):
        w = self.level_width(l)
        return tuple(s // w for s in self.n)

    def pack(self, l, c):
        g = self.grid(l)
        return c[:, 0] + g[0] * (c[:, 1] + g[1] * c[:, 2])

    def unpack(self, l, p):
        g = self.grid(l)
        return np.stack([p % g[0], (p // g[0]) % g[1], p // (g[0] * g[1])], axis=1)

    def width(self, lvl):
        return np.left_shift(np.int64(1), (self.max_level - np.asarray(lvl)).astype(np.int64))

    def root_cells(self):
        g = self.grid(self.root_level)
        return self.unpack(self.root_level, np.arange(g[0] * g[1] * g[2], dtype=np.int64))

    def _add_divided(self, l, packed):
        if packed.size:
            self.div[l] = np.union1d(self.div[l], packed)
            self._leaves = None

    # ---- refinement requests ---------------------------------------------------------------------------
    def refine_requests(self, nreq, levels, concerns):
        """``concerns(req_idx, lo, w)`` -> mask: does request req_idx still concern the cell with low
        corner ``lo`` (fine-cell units) and width ``w``?  Every concerned cell below the request's
        target level is divided."""
        levels = np.minimum(np.asarray(levels, dtype=np.int64), self.max_level)
        roots = self.root_cells()
        cells = np.repeat(roots, nreq, axis=0)
        req = np.tile(np.arange(nreq), roots.shape[0])
        l = self.root_level
        offs = np.array([[a, b, c] for c in (0, 1) for b in (0, 1) for a in (0, 1)], dtype=np.int64)
        while cells.shape[0] and l < self.max_level:
            w = self.level_width(l)
            ok = (levels[req] > l) & concerns(req, cells * w, np.full(req.size, w, dtype=np.int64))
            cells, req = cells[ok], req[ok]
            if not cells.shape[0]:
                break
            self._add_divided(l, np.unique(self.pack(l, cells)))
            # push the surviving pairs down to the 8 children
            cells = (2 * cells[:, None, :] + offs[None, :, :]).reshape(-1, 3)
            req = np.repeat(req, 8)
            l += 1

    def divide_leaves(self, leaf_lo, leaf_lvl):
        """Divide the given leaves (arrays), used by ``refine(function)``."""
        for l in np.unique(leaf_lvl):
            if l >= self.max_level:
                continue
            sel = leaf_lvl == l
            self._add_divided(int(l), np.unique(self.pack(int(l), leaf_lo[sel] // self.level_width(int(l)))))

    def set_leaves(self, lo, lvl):
        """Rebuild the divided sets from an explicit leaf list (ancestor closure)."""
        self.div = {l: np.zeros(0, dtype=np.int64) for l in range(self.root_level, self.max_level)}
        for l in range(self.root_level, self.max_level):
            sel = lvl > l
            if sel.any():
                self.div[l] = np.unique(self.pack(l, lo[sel] // self.level_width(l)))
        self._leaves = None

    # ---- balancing -----------------------------------------------------------------------------------------
    def balance(self, diagonal):
        dirs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)]
        if not diagonal:
            dirs = [d for d in dirs if sum(abs(x) for x in d) <= 1]
        dirs = np.array(dirs, dtype=np.int64)  # includes (0,0,0): a divided cell's own parent is divided
        for l in range(self.max_level - 1, self.root_level, -1):
            d = self.div[l]
            if not d.size:
                continue
            c = self.unpack(l, d)
            g = np.array(self.grid(l), dtype=np.int64)
            need = []
            for dv in dirs:
                nb = c + dv[None, :]
                inside = ((nb >= 0) & (nb < g[None, :])).all(axis=1)
                need.append(self.pack(l - 1, nb[inside] >> 1))
            self._add_divided(l - 1, np.unique(np.concatenate(need)))

    # ---- leaves ------------------------------------------------------------------------------------------------
    def leaves(self):
        """(lo, lvl) of all leaves in the reference's cell order."""
        if self._leaves is not None:
            return self._leaves
        offs = np.array([[a, b, c] for c in (0, 1) for b in (0, 1) for a in (0, 1)], dtype=np.int64)
        los, lvls = [], []
        cand = self.root_cells()  # cells existing at the current level
        for l in range(self.root_level, self.max_level + 1):
            if l < self.max_level and self.div[l].size:
                isdiv = np.isin(self.pack(l, cand), self.div[l], assume_unique=False)
            else:
                isdiv = np.zeros(cand.shape[0], dtype=bool)
            leaf = cand[~isdiv]
            los.append(leaf * self.level_width(l))
            lvls.append(np.full(leaf.shape[0], l, dtype=np.int64))
            if not isdiv.any():
                break
            cand = (2 * cand[isdiv][:, None, :] + offs[None, :, :]).reshape(-1, 3)
        lo = np.concatenate(los)
        lvl = np.concatenate(lvls)
        order = np.argsort(self.keys(lo), kind="stable")
        self._leaves = (lo[order], lvl[order])
        return self._leaves

    def keys(self, lo):
        """Position of a fine cell in the reference's cell order (root id, then Z-order code)."""
        root = lo // self.rw
        rel = lo - root * self.rw
        code = _spread3(rel[:, 0], self.nbits) | (_spread3(rel[:, 1], self.nbits) << 1) | (_spread3(rel[:, 2], self.nbits) << 2)
        root_id = root[:, 0] + self.roots[0] * (root[:, 1] + self.roots[1] * root[:, 2])
        return root_id * (np.int64(self.rw) ** 3) + code

    @property
    def n_leaves(self):
        return self.leaves()[1].size

    def state(self):
        """(cell-centre location in half steps, level) in the reference's cell order."""
        lo, lvl = self.leaves()
        w = self.width(lvl)
        return 2 * lo + w[:, None], lvl.copy()

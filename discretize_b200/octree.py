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


def _spread(v, nbits, stride):
    """Bit b of ``v`` moved to bit ``stride * b`` (Morton interleave helper)."""
    out = np.zeros_like(v)
    for b in range(nbits):
        out |= ((v >> b) & 1) << (stride * b)
    return out


def _child_offsets(dim):
    """Children in the reference's order: x bit fastest, then y, (then z)."""
    if dim == 2:
        return np.array([[a, b] for b in (0, 1) for a in (0, 1)], dtype=np.int64)
    return np.array([[a, b, c] for c in (0, 1) for b in (0, 1) for a in (0, 1)], dtype=np.int64)


class LinearOctree:
    def __init__(self, shape_cells):
        n = tuple(int(s) for s in shape_cells)
        if len(n) not in (2, 3):
            raise NotImplementedError("TreeMesh is 2-D (QuadTree) or 3-D (OcTree)")
        self.dim = len(n)
        for s in n:
            if s <= 0 or (s & (s - 1)) != 0:
                raise ValueError("length of cell width vectors must be a power of 2")
        self.n = n
        ls = [s.bit_length() - 1 for s in n]
        min_l = min(ls)
        cubic = all(l == ls[0] for l in ls)
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
        if self.dim == 2:
            return c[:, 0] + g[0] * c[:, 1]
        return c[:, 0] + g[0] * (c[:, 1] + g[1] * c[:, 2])

    def unpack(self, l, p):
        g = self.grid(l)
        if self.dim == 2:
            return np.stack([p % g[0], p // g[0]], axis=1)
        return np.stack([p % g[0], (p // g[0]) % g[1], p // (g[0] * g[1])], axis=1)

    def width(self, lvl):
        return np.left_shift(np.int64(1), (self.max_level - np.asarray(lvl)).astype(np.int64))

    def root_cells(self):
        g = self.grid(self.root_level)
        return self.unpack(self.root_level, np.arange(int(np.prod(g)), dtype=np.int64))

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
        offs = _child_offsets(self.dim)
        while cells.shape[0] and l < self.max_level:
            w = self.level_width(l)
            ok = (levels[req] > l) & concerns(req, cells * w, np.full(req.size, w, dtype=np.int64))
            cells, req = cells[ok], req[ok]
            if not cells.shape[0]:
                break
            self._add_divided(l, np.unique(self.pack(l, cells)))
            # push the surviving pairs down to the 2^dim children
            cells = (2 * cells[:, None, :] + offs[None, :, :]).reshape(-1, self.dim)
            req = np.repeat(req, offs.shape[0])
            l += 1

    def divide_leaves(self, leaf_lo, leaf_lvl):
        """Divide the given leaves (arrays), used by ``refine(function)``."""
        for l in np.unique(leaf_lvl):
            if l >= self.max_level:
                continue
            sel = leaf_lvl == l
            self._add_divided(int(l), np.unique(self.pack(int(l), leaf_lo[sel] // self.level_width(int(l)))))

    def refine_sequential(self, ask, diagonal):
        """The reference's ``refine(function)`` exactly (tree.cpp:398-419 with the in-flight balancing of
        ``Cell::divide`` :459-746): one depth-first pass in cell order; a LEAF reached by the pass is asked for its level
        and divided if it wants more; dividing a cell immediately divides every coarser (face-, with ``diagonal`` also
        edge- / corner-) neighbour, recursively.  Cells that such a ripple creates AHEAD of the pass are asked when the
        pass reaches them, cells it creates behind the pass never are -- the result depends on that order, so this path is
        sequential (it calls a Python function per cell anyway).  ``ask(level, coords_tuple) -> level``."""
        import itertools

        dim = self.dim
        divided = {l: set(self.div[l].tolist()) for l in self.div}
        grids = {l: self.grid(l) for l in range(self.root_level, self.max_level + 1)}
        dirs = [v for v in itertools.product((-1, 0, 1), repeat=dim) if any(v)]
        if not diagonal:
            dirs = [v for v in dirs if sum(abs(x) for x in v) == 1]
        offs = [tuple(int(x) for x in o) for o in _child_offsets(dim)]

        def pack(l, c):
            g = grids[l]
            return c[0] + g[0] * c[1] if dim == 2 else c[0] + g[0] * (c[1] + g[1] * c[2])

        def divide(l, c):
            p = pack(l, c)
            if l >= self.max_level or p in divided[l]:
                return
            divided[l].add(p)
            if l > self.root_level:
                g = grids[l]
                for v in dirs:
                    nb = tuple(a + b for a, b in zip(c, v))
                    if all(0 <= x < n for x, n in zip(nb, g)):
                        divide(l - 1, tuple(x >> 1 for x in nb))       # no-op unless that neighbour is coarser

        def visit(l, c):
            if l == self.max_level:
                return
            if pack(l, c) not in divided[l]:
                want = int(ask(l, c))
                if want < 0:
                    want = (self.max_level + 1) - (abs(want) % (self.max_level + 1))
                if want <= l:
                    return
                divide(l, c)
            for o in offs:
                visit(l + 1, tuple(2 * a + b for a, b in zip(c, o)))

        for root in self.root_cells():
            visit(self.root_level, tuple(int(x) for x in root))
        for l in self.div:
            self.div[l] = np.array(sorted(divided[l]), dtype=np.int64)
        self._leaves = None

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
        import itertools

        dirs = list(itertools.product((-1, 0, 1), repeat=self.dim))
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
        offs = _child_offsets(self.dim)
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
            cand = (2 * cand[isdiv][:, None, :] + offs[None, :, :]).reshape(-1, self.dim)
        lo = np.concatenate(los)
        lvl = np.concatenate(lvls)
        order = np.argsort(self.keys(lo), kind="stable")
        self._leaves = (lo[order], lvl[order])
        return self._leaves

    def keys(self, lo):
        """Position of a fine cell in the reference's cell order (root id, then Z-order code)."""
        root = lo // self.rw
        rel = lo - root * self.rw
        d = self.dim
        code = _spread(rel[:, 0], self.nbits, d)
        root_id = root[:, 0]
        stride = self.roots[0]
        for a in range(1, d):
            code = code | (_spread(rel[:, a], self.nbits, d) << a)
            root_id = root_id + stride * root[:, a]
            stride = stride * self.roots[a]
        return root_id * (np.int64(self.rw) ** d) + code

    @property
    def n_leaves(self):
        return self.leaves()[1].size

    def state(self):
        """(cell-centre location in half steps, level) in the reference's cell order."""
        lo, lvl = self.leaves()
        w = self.width(lvl)
        return 2 * lo + w[:, None], lvl.copy()

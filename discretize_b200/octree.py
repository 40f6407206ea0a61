"""Linear octree of leaf cells: refinement requests and 2:1 balancing on flat arrays.

The reference grows a pointer tree by recursive ``Cell::divide`` calls that balance as they go
(``discretize/_extensions/tree.cpp:459-746``).  This is a different design with the same result:
leaves are rows of flat arrays (low corner in fine-cell units + level) kept in the reference's
cell order (roots x-fastest, then Z-order: tree.cpp:232-344, 748-756, 949-953), refinement
requests are processed level by level as (leaf, request) pair lists, and the 2:1 (optionally
26-neighbour "diagonal") balance is restored afterwards by a ripple over the leaves that changed,
using binary search in the Z-ordered key array to find the leaf that contains a probe point.
The balanced closure of a set of refinement requests is unique, so the result does not depend on
the order in which the reference would have processed the same requests.
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
        # leaves start as the roots, x fastest
        rx, ry, rz = np.meshgrid(np.arange(self.roots[0]), np.arange(self.roots[1]), np.arange(self.roots[2]), indexing="ij")
        lo = np.stack([rx.reshape(-1, order="F"), ry.reshape(-1, order="F"), rz.reshape(-1, order="F")], axis=1) * self.rw
        self.lo = lo.astype(np.int64)
        self.lvl = np.full(self.lo.shape[0], self.root_level, dtype=np.int64)

    # ---- helpers -----------------------------------------------------------------------------
    def width(self, lvl):
        return np.left_shift(np.int64(1), (self.max_level - lvl).astype(np.int64))

    def keys(self, lo):
        """Position of a fine cell in the reference's cell order (root id, then Z-order code)."""
        root = lo // self.rw
        rel = lo - root * self.rw
        code = _spread3(rel[:, 0], self.nbits) | (_spread3(rel[:, 1], self.nbits) << 1) | (_spread3(rel[:, 2], self.nbits) << 2)
        root_id = root[:, 0] + self.roots[0] * (root[:, 1] + self.roots[1] * root[:, 2])
        return root_id * (np.int64(self.rw) ** 3) + code

    def sort(self):
        order = np.argsort(self.keys(self.lo), kind="stable")
        self.lo, self.lvl = self.lo[order], self.lvl[order]

    def split(self, mask):
        """Replace the leaves selected by ``mask`` with their 8 children; returns the children rows."""
        idx = np.nonzero(mask)[0]
        if idx.size == 0:
            return np.zeros(0, dtype=np.int64)
        keep = ~mask
        plo, plvl = self.lo[idx], self.lvl[idx]
        hw = self.width(plvl) // 2
        offs = np.array([[a, b, c] for c in (0, 1) for b in (0, 1) for a in (0, 1)], dtype=np.int64)
        clo = (plo[:, None, :] + hw[:, None, None] * offs[None, :, :]).reshape(-1, 3)
        clvl = np.repeat(plvl + 1, 8)
        n_keep = int(keep.sum())
        self.lo = np.concatenate([self.lo[keep], clo])
        self.lvl = np.concatenate([self.lvl[keep], clvl])
        return np.arange(n_keep, n_keep + clo.shape[0])

    # ---- refinement requests --------------------------------------------------------------------
    def refine_requests(self, nreq, levels, select_children):
        """Generic top-down refinement.

        ``levels[r]`` is the level request r wants; ``select_children(req_idx, child_lo, child_w)``
        returns a boolean mask over candidate (request, child) pairs saying whether the request
        still concerns that child cell.  Requests concern every root initially (filtered by the
        same predicate).  Leaves are split while a request of higher level concerns them.
        """
        levels = np.asarray(levels, dtype=np.int64)
        levels = np.minimum(levels, self.max_level)
        nl = self.lo.shape[0]
        leaf = np.repeat(np.arange(nl), nreq)
        req = np.tile(np.arange(nreq), nl)
        ok = select_children(req, self.lo[leaf], self.width(self.lvl[leaf])) & (levels[req] > self.lvl[leaf])
        leaf, req = leaf[ok], req[ok]
        while leaf.size:
            mask = np.zeros(self.lo.shape[0], dtype=bool)
            mask[leaf] = True
            # children of the split leaves, 8 per parent, parents in increasing row order
            parents = np.nonzero(mask)[0]
            rank = np.cumsum(mask) - 1                      # parent row -> rank among split parents
            child_rows = self.split(mask)                   # rows of children, 8 consecutive per parent
            base = child_rows[0] if child_rows.size else 0
            c_leaf = (base + 8 * rank[leaf])[:, None] + np.arange(8)[None, :]
            c_req = np.repeat(req, 8)
            c_leaf = c_leaf.reshape(-1)
            ok = select_children(c_req, self.lo[c_leaf], self.width(self.lvl[c_leaf])) & (levels[c_req] > self.lvl[c_leaf])
            leaf, req = c_leaf[ok], c_req[ok]
            del parents

    # ---- balancing ---------------------------------------------------------------------------------
    def balance(self, diagonal):
        """Restore the 2:1 condition across faces (and edges / corners when ``diagonal``)."""
        dirs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1) if (a, b, c) != (0, 0, 0)]
        if not diagonal:
            dirs = [d for d in dirs if sum(abs(x) for x in d) == 1]
        dirs = np.array(dirs, dtype=np.int64)
        nmax = np.array(self.n, dtype=np.int64)
        active = np.arange(self.lo.shape[0])
        while active.size:
            self_keys_order = np.argsort(self.keys(self.lo), kind="stable")
            skeys = self.keys(self.lo)[self_keys_order]
            alo, aw = self.lo[active], self.width(self.lvl[active])
            probe = (alo[:, None, :] + aw[:, None, None] * dirs[None, :, :]).reshape(-1, 3)
            owner = np.repeat(active, dirs.shape[0])
            inside = ((probe >= 0) & (probe < nmax[None, :])).all(axis=1)
            probe, owner = probe[inside], owner[inside]
            pos = np.searchsorted(skeys, self.keys(probe), side="right") - 1
            hit = self_keys_order[pos]                       # leaf containing the probe point
            too_coarse = self.lvl[hit] < self.lvl[owner] - 1
            if not too_coarse.any():
                break
            mask = np.zeros(self.lo.shape[0], dtype=bool)
            mask[hit[too_coarse]] = True
            again = np.unique(owner[too_coarse])
            # rows shift when leaves are removed: remember the triggering leaves by key
            again_lo = self.lo[again]
            children = self.split(mask)
            _, again_rows = self._rows_of(again_lo)
            active = np.unique(np.concatenate([children, again_rows]))

    def _rows_of(self, lo):
        order = np.argsort(self.keys(self.lo), kind="stable")
        skeys = self.keys(self.lo)[order]
        pos = np.searchsorted(skeys, self.keys(lo), side="right") - 1
        return skeys[pos], order[pos]

    # ---- state ----------------------------------------------------------------------------------------
    def state(self):
        """(cell-centre location in half steps, level) in the reference's cell order."""
        self.sort()
        w = self.width(self.lvl)
        return 2 * self.lo + w[:, None], self.lvl.copy()

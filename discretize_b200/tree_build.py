"""OcTree connectivity: leaf cells -> nodes / edges / faces, hanging entities, numbering.

A from-scratch, array-oriented builder that reproduces the numbering of the reference's
pointer-based C++ tree (``discretize/_extensions/tree.cpp``) without walking a tree:

* every entity is identified by its integer location in half-step coordinates (cell centres
  odd, nodes even; tree.cpp:26-84, 109-136) and ordered by the double Cantor pairing of that
  location (``key_func``, tree.h:13-19) -- the order of the reference's ``std::map`` containers;
* entities of all leaf cells are generated in bulk and deduplicated by sort-unique on the keys;
* hanging entities are found geometrically instead of by the reference's per-face rule tables
  (tree.cpp:1039-1227): in a 2:1 balanced tree a node is hanging iff it sits on the midpoint
  of a leaf edge or the centre of a leaf face, an edge iff it is half of a leaf edge (parents:
  that edge, twice) or a cross edge of a leaf face whose centre node exists (parents: the two
  parallel face edges), a face iff it is a quarter of a leaf face (parent: that face);
* numbering puts non-hanging entities first, each group in key order (tree.cpp:1343-1451);
  cells are numbered in depth-first Z-order of the roots (tree.cpp:748-756, 949-953).

Only integer arithmetic happens here; geometry (volumes, areas, lengths) is evaluated from the
half-step coordinate arrays exactly like tree.cpp:61-84, 109-136, 176-230.
"""
from __future__ import annotations

import numpy as np


def cantor2(x, y):
    s = x + y
    return (s * (s + 1)) // 2 + y


def key3(x, y, z):
    """tree.h:13-19 key_func(x, y, z) on int64 arrays."""
    return cantor2(cantor2(x, y), z)


def _lookup(sorted_keys, keys):
    """(position, found) of ``keys`` in the sorted unique array ``sorted_keys``."""
    pos = np.searchsorted(sorted_keys, keys)
    pos_c = np.minimum(pos, sorted_keys.size - 1)
    found = sorted_keys[pos_c] == keys
    return pos_c, found


def _unique_keys(keys):
    """np.unique(keys, return_index=True, return_inverse=True); sorts on the GPU when one is present
    (tens of millions of keys per family: ~100x faster there), identical results either way."""
    try:
        import torch

        if torch.cuda.is_available() and keys.size > 2_000_000:
            k = torch.from_numpy(keys).cuda()
            uniq, inv = torch.unique(k, sorted=True, return_inverse=True)
            first = torch.full((uniq.numel(),), keys.size, dtype=torch.int64, device=k.device)
            first.scatter_reduce_(0, inv, torch.arange(keys.size, device=k.device), reduce="amin")
            return uniq.cpu().numpy(), first.cpu().numpy(), inv.cpu().numpy()
    except Exception:  # pragma: no cover - any torch / memory problem: fall back to numpy
        pass
    return np.unique(keys, return_index=True, return_inverse=True)


class Entities:
    """One family of entities (nodes, x-edges, ..., z-faces) in key order."""

    def __init__(self, loc):
        keys = key3(loc[:, 0], loc[:, 1], loc[:, 2])
        self.keys, first, self.inverse = _unique_keys(keys)
        self.loc = loc[first]
        self.n_total = self.keys.size
        self.hanging = np.zeros(self.n_total, dtype=bool)
        self.index = None  # reference numbering (non-hanging first)
        self.half = None   # half extents per entity where meaningful

    def number(self):
        nh = ~self.hanging
        self.n = int(nh.sum())
        self.n_hanging = self.n_total - self.n
        idx = np.empty(self.n_total, dtype=np.int64)
        idx[nh] = np.arange(self.n)
        idx[self.hanging] = self.n + np.arange(self.n_hanging)
        self.index = idx
        self.order = np.argsort(idx, kind="stable")  # numbering index -> sorted position


# offsets (in units of the cell half width) of the local entities of a cell, reference order
_NODE_OFF = np.array([[sx, sy, sz] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)])
_EDGE_OFF = {
    0: np.array([[0, -1, -1], [0, 1, -1], [0, -1, 1], [0, 1, 1]]),   # ex[0..3] (tree.cpp:967-970)
    1: np.array([[-1, 0, -1], [1, 0, -1], [-1, 0, 1], [1, 0, 1]]),   # ey[0..3]
    2: np.array([[-1, -1, 0], [1, -1, 0], [-1, 1, 0], [1, 1, 0]]),   # ez[0..3]
}
_FACE_OFF = {
    0: np.array([[-1, 0, 0], [1, 0, 0]]),
    1: np.array([[0, -1, 0], [0, 1, 0]]),
    2: np.array([[0, 0, -1], [0, 0, 1]]),
}


class TreeTopology:
    """All integer connectivity of a finalized OcTree (3-D)."""

    def __init__(self, centers, half, shape_half):
        """centers: (nC,3) int64 cell-centre locations (half steps), in reference cell order;
        half: (nC,) int64 half widths (= width in fine cells); shape_half: (2nx, 2ny, 2nz)."""
        c = np.asarray(centers, dtype=np.int64)
        w = np.asarray(half, dtype=np.int64)
        self.centers, self.half, self.shape_half = c, w, tuple(int(s) for s in shape_half)
        nC = c.shape[0]
        self.n_cells = nC

        def expand(off):  # (nC, k, 3) locations of local entities
            return c[:, None, :] + w[:, None, None] * off[None, :, :]

        # ---- unique entities -------------------------------------------------------------
        nloc = expand(_NODE_OFF)
        self.nodes = Entities(nloc.reshape(-1, 3))
        self.cell_nodes = self.nodes.inverse.reshape(nC, 8)          # sorted positions
        self.edges, self.cell_edges = [], []
        for d in range(3):
            loc = expand(_EDGE_OFF[d])
            e = Entities(loc.reshape(-1, 3))
            e.half = np.zeros(e.n_total, dtype=np.int64)
            e.half[e.inverse] = np.repeat(w, 4)
            self.edges.append(e)
            self.cell_edges.append(e.inverse.reshape(nC, 4))
        self.faces, self.cell_faces = [], []
        for d in range(3):
            loc = expand(_FACE_OFF[d])
            f = Entities(loc.reshape(-1, 3))
            f.half = np.zeros(f.n_total, dtype=np.int64)
            f.half[f.inverse] = np.repeat(w, 2)
            self.faces.append(f)
            self.cell_faces.append(f.inverse.reshape(nC, 2))

        self._find_hanging()
        for fam in [self.nodes] + self.edges + self.faces:
            fam.number()

    # ------------------------------------------------------------------------------------
    def _find_hanging(self):
        nodes = self.nodes
        node_par = np.full((nodes.n_total, 4), -1, dtype=np.int64)  # sorted node positions

        unit = np.eye(3, dtype=np.int64)
        # hanging nodes on edge midpoints: node key == edge key (an edge's location IS its midpoint)
        for d in range(3):
            e = self.edges[d]
            pos, found = _lookup(nodes.keys, e.keys)
            hit = np.nonzero(found)[0]
            if hit.size:
                n = pos[hit]
                lo = e.loc[hit] - e.half[hit, None] * unit[d]
                hi = e.loc[hit] + e.half[hit, None] * unit[d]
                plo, flo = _lookup(nodes.keys, key3(lo[:, 0], lo[:, 1], lo[:, 2]))
                phi, fhi = _lookup(nodes.keys, key3(hi[:, 0], hi[:, 1], hi[:, 2]))
                assert flo.all() and fhi.all()
                nodes.hanging[n] = True
                node_par[n, 0], node_par[n, 1], node_par[n, 2], node_par[n, 3] = plo, phi, plo, phi
        # hanging nodes on face centres: node key == face key; parents = the 4 face corners
        for d in range(3):
            f = self.faces[d]
            t1, t2 = [a for a in range(3) if a != d]
            pos, found = _lookup(nodes.keys, f.keys)
            hit = np.nonzero(found)[0]
            if hit.size:
                n = pos[hit]
                nodes.hanging[n] = True
                k = 0
                for s2 in (-1, 1):
                    for s1 in (-1, 1):
                        q = f.loc[hit] + f.half[hit, None] * (s1 * unit[t1] + s2 * unit[t2])
                        pq, fq = _lookup(nodes.keys, key3(q[:, 0], q[:, 1], q[:, 2]))
                        assert fq.all()
                        node_par[n, k] = pq
                        k += 1
        self.node_parents = node_par

        # hanging edges
        self.edge_parents = []
        for d in range(3):
            e = self.edges[d]
            par = np.full((e.n_total, 2), -1, dtype=np.int64)
            # (a) halves of a leaf edge whose midpoint node exists
            big = np.nonzero(e.half >= 2)[0]
            if big.size:
                for s in (-1, 1):
                    q = e.loc[big] + s * (e.half[big, None] // 2) * unit[d]
                    pq, fq = _lookup(e.keys, key3(q[:, 0], q[:, 1], q[:, 2]))
                    fq &= e.half[pq] * 2 == e.half[big]
                    ch = pq[fq]
                    e.hanging[ch] = True
                    par[ch, 0] = big[fq]
                    par[ch, 1] = big[fq]
            # (b) cross edges of a leaf face (normal dn != d) whose centre node exists
            for dn in range(3):
                if dn == d:
                    continue
                f = self.faces[dn]
                t = [a for a in range(3) if a not in (d, dn)][0]  # the in-face direction across the edge
                _, has_center = _lookup(nodes.keys, f.keys)
                fb = np.nonzero(has_center & (f.half >= 2))[0]
                if not fb.size:
                    continue
                # the two face edges along d sit at +-half in direction t
                p0 = f.loc[fb] - f.half[fb, None] * unit[t]
                p1 = f.loc[fb] + f.half[fb, None] * unit[t]
                i0, f0 = _lookup(e.keys, key3(p0[:, 0], p0[:, 1], p0[:, 2]))
                i1, f1 = _lookup(e.keys, key3(p1[:, 0], p1[:, 1], p1[:, 2]))
                assert f0.all() and f1.all()
                for s in (-1, 1):
                    q = f.loc[fb] + s * (f.half[fb, None] // 2) * unit[d]
                    pq, fq = _lookup(e.keys, key3(q[:, 0], q[:, 1], q[:, 2]))
                    fq &= e.half[pq] * 2 == f.half[fb]
                    ch = pq[fq]
                    e.hanging[ch] = True
                    par[ch, 0] = i0[fq]
                    par[ch, 1] = i1[fq]
            self.edge_parents.append(par)

        # hanging faces: quarters of a leaf face whose centre node exists
        self.face_parents = []
        for d in range(3):
            f = self.faces[d]
            t1, t2 = [a for a in range(3) if a != d]
            par = np.full(f.n_total, -1, dtype=np.int64)
            _, has_center = _lookup(nodes.keys, f.keys)
            fb = np.nonzero(has_center & (f.half >= 2))[0]
            if fb.size:
                for s2 in (-1, 1):
                    for s1 in (-1, 1):
                        q = f.loc[fb] + (f.half[fb, None] // 2) * (s1 * unit[t1] + s2 * unit[t2])
                        pq, fq = _lookup(f.keys, key3(q[:, 0], q[:, 1], q[:, 2]))
                        fq &= f.half[pq] * 2 == f.half[fb]
                        ch = pq[fq]
                        f.hanging[ch] = True
                        par[ch] = fb[fq]
            self.face_parents.append(par)


def morton_order(centers, half, root_half, roots_shape):
    """Permutation that sorts leaf cells into the reference's cell order: roots x-fastest, then
    depth-first with children ordered x, y, z bit (tree.cpp:232-344, 748-756, 949-953)."""
    c = np.asarray(centers, dtype=np.int64)
    lo = (c - np.asarray(half, dtype=np.int64)[:, None]) // 2  # low corner in fine cells
    rw = int(root_half)                                          # root width in fine cells
    root = lo // rw
    rel = lo - root * rw
    nbits = max(int(rw).bit_length() - 1, 0)
    code = np.zeros(c.shape[0], dtype=np.int64)
    for b in range(nbits):
        for a in range(3):
            code |= ((rel[:, a] >> b) & 1) << (3 * b + a)
    root_id = root[:, 0] + roots_shape[0] * (root[:, 1] + roots_shape[1] * root[:, 2])
    return np.lexsort((code, root_id))

"""TreeMesh (OcTree): the reference's names, connectivity rebuilt from scratch, CUDA operators.

Drop-in for the operator-application part of ``discretize.TreeMesh`` in 3-D
(reference: discretize/tree_mesh.py:106-270 and discretize/_extensions/tree_ext.pyx).

* refinement (``insert_cells``, ``refine_ball``, ``refine_points``, ``refine``) runs on a linear
  octree (``octree.py``); ``finalize()`` derives nodes / edges / faces, hanging entities and the
  reference's numbering with array sort-unique passes (``tree_build.py``);
* every operator is assembled ONCE on the host as "total operator x deflation" exactly like
  tree_ext.pyx:3114-4762 (hanging faces -> parent, hanging edges -> mean of two parents, hanging
  nodes -> mean of four listed parents, repeated until only non-hanging columns remain), then
  uploaded as CSR with one byte per entry that encodes (multiplier, row-geometry selector):
  the values are regenerated on the fly by the ``tree_spmv`` kernel from per-row geometry;
* cell-centred operators, boundary terms, interpolation (``tree_interp.py``) and the QuadTree
  (``quad_mesh.QuadTreeMesh``, returned by ``TreeMesh([hx, hy])``) are host-assembled CSR applied by
  ``dzb_csr_spmv``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .octree import LinearOctree
from .tensor_mesh import TensorMesh, _ALIASES, _MeshPersistence, _REMOVED, as_array_n_by_dim, is_scalar, removed_method, removed_name, unpack_widths
from .tree_build import TreeTopology, key3


class TreeMeshNotFinalizedError(RuntimeError):
    """reference: tree_ext.pyx:7071-7080."""


_TREE_ALIASES = dict(_ALIASES)
_TREE_ALIASES.update({
    "ntN": "n_total_nodes", "ntEx": "n_total_edges_x", "ntEy": "n_total_edges_y", "ntEz": "n_total_edges_z",
    "ntE": "n_total_edges", "ntFx": "n_total_faces_x", "ntFy": "n_total_faces_y", "ntFz": "n_total_faces_z",
    "ntF": "n_total_faces", "nhN": "n_hanging_nodes", "nhEx": "n_hanging_edges_x", "nhEy": "n_hanging_edges_y",
    "nhEz": "n_hanging_edges_z", "nhE": "n_hanging_edges", "nhFx": "n_hanging_faces_x", "nhFy": "n_hanging_faces_y",
    "nhFz": "n_hanging_faces_z", "nhF": "n_hanging_faces",
})


class TreeCell:
    """One cell of a TreeMesh: what ``mesh[i]`` returns and what a ``refine(function)`` callback receives
    (reference: ``TreeCell``, tree_ext.pyx:23-345).  During refinement the cell is not numbered yet: ``index`` is -1 and
    the connectivity properties are unavailable, as they would be meaningless in the reference at that point."""

    __slots__ = ("_mesh", "_lo", "_width", "_lvl", "_idx")

    def __init__(self, mesh, lo, width, level, index=-1):
        self._mesh, self._lo, self._width, self._lvl, self._idx = mesh, np.asarray(lo, dtype=np.int64), int(width), int(level), int(index)

    def _corner(self, side):
        xs = self._mesh._xs
        return np.array([xs[d][2 * (self._lo[d] + side * self._width)] for d in range(self._mesh.dim)])

    dim = property(lambda self: self._mesh.dim)
    index = property(lambda self: self._idx)
    _level = property(lambda self: self._lvl)
    _index_loc = property(lambda self: tuple(int(2 * l + self._width) for l in self._lo))
    origin = property(lambda self: self._corner(0))
    x0 = property(lambda self: self._corner(0))
    h = property(lambda self: self._corner(1) - self._corner(0))
    center = property(lambda self: (self._corner(0) + self._corner(1)) * 0.5)

    @property
    def bounds(self):
        a, b = self._corner(0), self._corner(1)
        return np.stack([a, b], axis=1).reshape(-1)

    def _topology(self):
        if self._idx < 0:
            raise TreeMeshNotFinalizedError("cell connectivity is available once the TreeMesh is finalized")
        return self._mesh._t()

    @property
    def nodes(self):
        t = self._topology()
        return [int(i) for i in t.nodes.index[t.cell_nodes[self._idx]]]

    @property
    def edges(self):
        t = self._topology()
        return [int(i) for d in range(self._mesh.dim) for i in t.edges[d].index[t.cell_edges[d][self._idx]]]

    @property
    def faces(self):
        t = self._topology()
        return [int(i) for d in range(self._mesh.dim) for i in t.faces[d].index[t.cell_faces[d][self._idx]]]

    @property
    def neighbors(self):
        """[-x, +x, -y, +y, (-z, +z)]: a cell index, -1 outside the mesh, or the list of the finer cells across that
        face (lower tangential axis fastest)."""
        self._topology()
        m = self._mesh
        dim = m.dim
        c, h = self.center, self.h
        out = []
        for axis in range(dim):
            tang = [a for a in range(dim) if a != axis]
            for sign in (-1.0, 1.0):
                pts = []
                for corner in range(2 ** (dim - 1)):
                    p = c.copy()
                    p[axis] += sign * 0.75 * h[axis]
                    for k, a in enumerate(tang):
                        p[a] += (0.25 if (corner >> k) & 1 else -0.25) * h[a]
                    pts.append(p)
                pts = np.array(pts)
                if not m.is_inside(pts)[0]:
                    out.append(-1)
                    continue
                idx = m._containing_cells(pts)
                out.append(int(idx[0]) if np.all(idx == idx[0]) else [int(i) for i in idx])
        return out


_TREE_REMOVED = dict(_REMOVED)
_TREE_REMOVED.update({"cellGradStencil": "cell_gradient_stencil", "maxLevel": "max_used_level",
                      "permuteCC": "permute_cells", "permuteE": "permute_edges", "permuteF": "permute_faces"})


def _cuda_ready():
    """True where the CUDA library can run (a device is visible); host-side tree logic works without one."""
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # pragma: no cover
        return False


class TreeMesh(_MeshPersistence):
    """3-D OcTree.  ``TreeMesh([hx, hy])`` gives the 2-D QuadTree (``quad_mesh.QuadTreeMesh``, a subclass that shares
    the constructor and every refinement method, and assembles its operators on the host)."""

    _meshType = "TREE"
    _items = ("h", "origin", "cell_state")
    _LOC_NAMES = TensorMesh._LOC_NAMES
    _parse_location_type = TensorMesh._parse_location_type
    dim = 3

    def __new__(cls, h=None, *args, **kwargs):
        if cls is TreeMesh and h is not None and len(list(h)) == 2:
            from .quad_mesh import QuadTreeMesh

            return object.__new__(QuadTreeMesh)
        return object.__new__(cls)

    def __init__(self, h=None, origin=None, diagonal_balance=None, **kwargs):
        if "x0" in kwargs:
            origin = kwargs.pop("x0")
        cell_state = kwargs.pop("cell_state", None)
        cell_indexes, cell_levels = kwargs.pop("cell_indexes", None), kwargs.pop("cell_levels", None)
        if cell_state is None and cell_indexes is not None and cell_levels is not None:
            cell_state = {"indexes": cell_indexes, "levels": cell_levels}
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        if diagonal_balance is None:
            diagonal_balance = False  # reference default (FutureWarning there: tree_mesh.py:246-256)
        h = list(h)
        if len(h) != self.dim:
            raise ValueError(f"{type(self).__name__} is {self.dim}-D but {len(h)} width vectors were given")
        for i, h_i in enumerate(h):
            if is_scalar(h_i) and not isinstance(h_i, np.ndarray):
                h_i = np.ones(int(h_i)) / int(h_i)
            elif isinstance(h_i, (list, tuple)):
                h_i = unpack_widths(list(h_i))
            h[i] = np.array(h_i, dtype=np.float64)
        self._h = tuple(h)
        self._origin = np.zeros(self.dim)
        if origin is not None:
            value = list(origin)
            for i, (val, h_i) in enumerate(zip(value, self._h)):
                if isinstance(val, str):
                    value[i] = {"C": -h_i.sum() * 0.5, "N": -h_i.sum(), "0": 0.0}[val]
            self._origin = np.asarray(value, dtype=np.float64)
        self._tree = LinearOctree([x.size for x in self._h])  # raises for non powers of two
        self._diagonal_balance = bool(diagonal_balance)
        self._build_coordinates()
        self._finalized = False
        self._topo = None
        self._cache = {}
        self._dev = {}
        if cell_state is not None:       # tree_mesh.py:258-270: rebuild a saved mesh
            self.__setstate__((cell_state["indexes"], cell_state["levels"]))

    def _build_coordinates(self):
        """Half-step coordinate arrays: nodes at even, midpoints at odd positions (tree_ext.pyx:397-421)."""
        self._xs = []
        for a in range(self.dim):
            n = np.cumsum(np.r_[self._origin[a], self._h[a]])
            x = np.empty(2 * self._h[a].size + 1)
            x[::2] = n
            x[1::2] = (x[:-1:2] + x[2::2]) / 2
            self._xs.append(x)

    def _set_origin(self, value):
        """Move the mesh: the cells stay, every coordinate-dependent cache is dropped (tree_ext.pyx:1512-1540)."""
        if not isinstance(value, (list, tuple, np.ndarray)):
            raise ValueError("origin must be a list, tuple or numpy array")
        value = list(value)
        if len(value) != self.dim:
            raise ValueError("Dimension mismatch. len(origin) != len(h)")
        for i, (val, h_i) in enumerate(zip(value, self._h)):
            if isinstance(val, str):
                value[i] = {"C": -h_i.sum() * 0.5, "N": -h_i.sum(), "0": 0.0}[val]
        self._origin = np.asarray(value, dtype=np.float64)
        self._build_coordinates()
        self._cache = {}
        self._dev = {}

    def __getattr__(self, name):
        if name in _TREE_ALIASES:
            return getattr(self, _TREE_ALIASES[name])
        err = removed_name(type(self).__name__, name, _TREE_REMOVED)
        if err is not None:
            raise err
        method = None if name == "r" else removed_method("TreeMesh", name)
        if method is not None:
            return method
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    # ---- description ------------------------------------------------------------------------------
    h = property(lambda self: self._h)
    origin = property(lambda self: self._origin, lambda self, value: self._set_origin(value))
    shape_cells = property(lambda self: tuple(int(x.size) for x in self._h))
    max_level = property(lambda self: self._tree.max_level)
    finalized = property(lambda self: self._finalized)

    def _parse_level(self, level):
        if level < 0:
            level = (self.max_level + 1) - (abs(level) % (self.max_level + 1))
        if level > self.max_level:
            raise IndexError(f"Level beyond max octree level, {self.max_level}")
        return level

    def _require_open(self):
        if self._finalized:
            raise RuntimeError("the TreeMesh is finalized; it cannot be refined any more")

    # ---- refinement ----------------------------------------------------------------------------------
    def _cell_box(self, lo, w):
        a = np.stack([self._xs[d][2 * lo[:, d]] for d in range(self.dim)], axis=1)
        b = np.stack([self._xs[d][2 * (lo[:, d] + w)] for d in range(self.dim)], axis=1)
        return a, b

    def insert_cells(self, points, levels, finalize=True, diagonal_balance=None):
        """reference: tree_ext.pyx:1158-1215 -> tree.cpp:383-396, 910-927."""
        self._require_open()
        points = as_array_n_by_dim(np.asarray(points, dtype=np.float64), self.dim)
        levels = np.atleast_1d(np.asarray(levels, dtype=np.int64))
        # The reference loops over ``levels`` (tree_ext.pyx:1205-1211): a single level with many points
        # inserts only the FIRST point, and a single point with many levels inserts it at each level.
        # Mirrored here so that the same calls give the same mesh.
        if levels.size == 1:
            points = points[:1]
        elif points.shape[0] == 1:
            points = np.repeat(points, levels.size, axis=0)
        elif points.shape[0] != levels.size:
            raise ValueError("points and levels have incompatible lengths")
        levels = np.array([self._parse_level(int(l)) for l in levels], dtype=np.int64)
        db = self._diagonal_balance if diagonal_balance is None else diagonal_balance
        xs = self._xs

        def concerns(req, lo, w):
            # the reference descends by `x > centre of the parent` (ties go low), starting from the root
            # found by `x >= root edge`: a cell is on the path iff, on every axis, its interval holds the
            # point in that sense.  Points outside the mesh clamp to the boundary cells.
            ok = np.ones(req.size, dtype=bool)
            for d in range(self.dim):
                p = points[req, d]
                a = xs[d][2 * lo[:, d]]
                b = xs[d][2 * (lo[:, d] + w)]
                first = lo[:, d] == 0
                last = lo[:, d] + w == self.shape_cells[d]
                hi_root = (lo[:, d] + w) % self._tree.rw == 0
                ok &= ((p > a) | first | self._root_aligned_low(lo[:, d], p, a)) & ((p < b) | last | (~hi_root & (p == b)))
            return ok

        self._tree.refine_requests(points.shape[0], levels, concerns)
        self._tree.balance(db)
        if finalize:
            self.finalize()

    def _root_aligned_low(self, lo_d, p, a):
        # root selection uses `p >= root low edge` (tree.cpp:914-925); inside a root the split uses `p > centre`
        return (lo_d % self._tree.rw == 0) & (p == a)

    def refine_ball(self, points, radii, levels, finalize=True, diagonal_balance=None):
        """reference: tree_ext.pyx:561-648 -> tree.h:149-167 + geom.cpp:35-58 (Ball::intersects_cell)."""
        self._require_open()
        points = as_array_n_by_dim(np.asarray(points, dtype=np.float64), self.dim)
        n = points.shape[0]
        radii = np.broadcast_to(np.atleast_1d(np.asarray(radii, dtype=np.float64)), (n,)).copy()
        levels = np.atleast_1d(np.asarray(levels, dtype=np.int64))
        if levels.size == 1:
            levels = np.full(n, levels[0])
        levels = np.array([self._parse_level(int(l)) for l in levels], dtype=np.int64)
        db = self._diagonal_balance if diagonal_balance is None else diagonal_balance
        rsq = radii * radii

        def concerns(req, lo, w):
            a, b = self._cell_box(lo, w)
            x0 = points[req]
            dx = np.maximum(a, np.minimum(x0, b)) - x0
            return (dx * dx).sum(axis=1) < rsq[req]

        self._tree.refine_requests(n, levels, concerns)
        self._tree.balance(db)
        if finalize:
            self.finalize()

    def refine_box(self, x0s, x1s, levels, finalize=True, diagonal_balance=None):
        """Refine every cell that intersects the axis-aligned box(es) [x0, x1] (closed intervals) to ``levels``.
        reference: tree_ext.pyx:650-730 -> geom.cpp:102-112 (Box::intersects_cell)."""
        self._require_open()
        x0s = as_array_n_by_dim(np.asarray(x0s, dtype=np.float64), self.dim)
        x1s = as_array_n_by_dim(np.asarray(x1s, dtype=np.float64), self.dim)
        levels = np.atleast_1d(np.asarray(levels, dtype=np.int64))
        n = max(x0s.shape[0], x1s.shape[0], levels.size)
        for name, arr_n in (("x0s", x0s.shape[0]), ("x1s", x1s.shape[0]), ("levels", levels.size)):
            if arr_n not in (1, n):
                raise ValueError(f"First dimensions of x0s, x1s, levels are incompatible ({name} has {arr_n}, expected 1 or {n})")
        x0s = np.broadcast_to(x0s, (n, self.dim))
        x1s = np.broadcast_to(x1s, (n, self.dim))
        levels = np.array([self._parse_level(int(l)) for l in np.broadcast_to(levels, (n,))], dtype=np.int64)
        db = self._diagonal_balance if diagonal_balance is None else diagonal_balance
        lo_box, hi_box = np.minimum(x0s, x1s), np.maximum(x0s, x1s)

        def concerns(req, lo, w):
            a, b = self._cell_box(lo, w)
            return ((hi_box[req] >= a) & (lo_box[req] <= b)).all(axis=1)

        self._tree.refine_requests(n, levels, concerns)
        self._tree.balance(db)
        if finalize:
            self.finalize()

    def refine_bounding_box(self, points, level=-1, padding_cells_by_level=None, finalize=True, diagonal_balance=None):
        """Refine within the bounding box of ``points``, padded by a number of cells per level.
        reference: tree_mesh.py:443-546 (a sequence of growing boxes handed to ``refine_box``)."""
        points = np.atleast_2d(np.asarray(points, dtype=np.float64))
        bsw, tnw = points.min(axis=0), points.max(axis=0)
        if level < 0:
            level = (self.max_level + 1) - (abs(level) % (self.max_level + 1))
        if level > self.max_level:
            raise IndexError(f"Level beyond max octree level, {self.max_level}")
        if padding_cells_by_level is None:
            padding_cells_by_level = np.array([0])
        padding_cells_by_level = np.asarray(padding_cells_by_level)
        if padding_cells_by_level.ndim == 0 and level > 1:
            padding_cells_by_level = np.full(level - 1, padding_cells_by_level)
        padding_cells_by_level = np.atleast_1d(padding_cells_by_level)
        if padding_cells_by_level.ndim == 2 and padding_cells_by_level.shape[1] != self.dim:
            raise ValueError("incorrect dimension for padding_cells_by_level.")
        h_min = np.array([h.min() for h in self.h])
        x0, x1, ls = [], [], []
        for lv, n_pad in zip(np.arange(level, 1, -1), padding_cells_by_level):   # stops with the shorter sequence
            pad = n_pad * h_min * 2 ** (self.max_level - lv)
            bsw, tnw = bsw - pad, tnw + pad
            x0.append(bsw); x1.append(tnw); ls.append(lv)
        self.refine_box(np.array(x0), np.array(x1), np.array(ls), finalize=finalize, diagonal_balance=diagonal_balance)

    def refine_points(self, points, level=-1, padding_cells_by_level=None, finalize=True, diagonal_balance=None):
        """reference: tree_mesh.py:548-636."""
        level = self._parse_level(level)
        if padding_cells_by_level is None:
            padding_cells_by_level = np.array([0])
        padding_cells_by_level = np.asarray(padding_cells_by_level)
        if padding_cells_by_level.ndim == 0 and level > 1:
            padding_cells_by_level = np.full(level - 1, padding_cells_by_level)
        padding_cells_by_level = np.atleast_1d(padding_cells_by_level)
        h_min = np.max([h.min() for h in self.h])
        radius_at_level = 0.0
        for lv, n_pad in zip(np.arange(level, 1, -1), padding_cells_by_level):
            radius_at_level += n_pad * h_min * 2 ** (self.max_level - lv)
            if radius_at_level == 0:
                self.insert_cells(points, lv, finalize=False, diagonal_balance=diagonal_balance)
            else:
                self.refine_ball(points, radius_at_level, lv, finalize=False, diagonal_balance=diagonal_balance)
        if finalize:
            self.finalize()

    def refine(self, function, finalize=True, diagonal_balance=None, vectorized=False):
        """Refine to a constant level (int) or with ``function(cell) -> level``, called with a :class:`TreeCell` for the
        cells a depth-first pass reaches as leaves -- the reference's semantics, including its dependence on the order in
        which balancing creates cells (tree_ext.pyx:494-558, tree.cpp:398-419).  ``vectorized=True`` (an extension) calls
        ``function(cell_centers) -> levels`` on all current cells per pass instead and balances at the end."""
        self._require_open()
        db = self._diagonal_balance if diagonal_balance is None else diagonal_balance
        t = self._tree
        if callable(function) and not vectorized:
            def ask(level, coords):
                w = t.level_width(level)
                return function(TreeCell(self, [c * w for c in coords], w, level))

            t.refine_sequential(ask, db)
        else:
            while True:
                lo, lvl = t.leaves()
                w = t.width(lvl)
                if callable(function):
                    c = np.stack([(self._xs[d][2 * lo[:, d]] + self._xs[d][2 * (lo[:, d] + w)]) * 0.5
                                  for d in range(self.dim)], axis=1)
                    want = np.asarray(function(c), dtype=np.int64)
                else:
                    want = np.full(lvl.size, self._parse_level(int(function)), dtype=np.int64)
                mask = (want > lvl) & (lvl < t.max_level)
                if not mask.any():
                    break
                t.divide_leaves(lo[mask], lvl[mask])
            t.balance(db)
        if finalize:
            self.finalize()

    # ---- geometric refinement and cell queries (tree_ext.pyx:735-1150, 1354-1510; predicates in tree_geometry.py) -------
    def _wrap_level(self, level):
        level = int(level)
        return (self.max_level + 1) - (abs(level) % (self.max_level + 1)) if level < 0 else level

    @staticmethod
    def _broadcast_first(**arrays):
        """Number of items of first-dimension broadcasting; the reference's message when they do not agree."""
        n = 1
        for arr in arrays.values():
            k = arr.shape[0]
            if k != 1:
                if n == 1:
                    n = k
                elif k != n:
                    msg = "First dimensions of" + ",".join(f" {key}: {a.shape}" for key, a in arrays.items())
                    raise ValueError(msg + " do not broadcast.")
        return n

    def _array_with_dim(self, name, arr, ndim):
        arr = np.asarray(arr, dtype=np.float64)
        while arr.ndim < ndim:
            arr = arr[None]
        if arr.ndim != ndim or arr.shape[-1] != self.dim:
            raise ValueError(f"Expected the last dimension of {name}.shape={arr.shape} to be {self.dim}.")
        return arr

    def _refine_geometric(self, predicate, n, levels, finalize, diagonal_balance):
        self._require_open()
        levels = np.atleast_1d(np.asarray(levels, dtype=np.int64))
        levels = np.array([self._wrap_level(l) for l in np.broadcast_to(levels, (n,))], dtype=np.int64)
        db = self._diagonal_balance if diagonal_balance is None else diagonal_balance

        def concerns(req, lo, w):
            a, b = self._cell_box(lo, w)
            return predicate(req, a, b)

        self._tree.refine_requests(n, levels, concerns)
        self._tree.balance(db)
        if finalize:
            self.finalize()

    def refine_line(self, path, levels, finalize=True, diagonal_balance=None):
        """Refine every cell a segment of the polyline ``path`` passes through (tree_ext.pyx:735-809)."""
        from . import tree_geometry as tg

        path = self._array_with_dim("path", path, 2)
        levels = np.atleast_1d(np.asarray(levels, dtype=np.int64))
        n = self._broadcast_first(path=path[:-1], levels=levels)
        starts = np.broadcast_to(path[:-1], (n, self.dim)) if path.shape[0] == 2 else path[:-1]
        ends = np.broadcast_to(path[1:], (n, self.dim)) if path.shape[0] == 2 else path[1:]
        self._refine_geometric(tg.line(starts, ends), n, levels, finalize, diagonal_balance)

    def refine_plane(self, origins, normals, levels, finalize=True, diagonal_balance=None):
        """reference: tree_ext.pyx:812-894."""
        from . import tree_geometry as tg

        origins = self._array_with_dim("origins", origins, 2)
        normals = self._array_with_dim("normals", normals, 2)
        levels = np.atleast_1d(np.asarray(levels, dtype=np.int64))
        n = self._broadcast_first(origins=origins, normals=normals, levels=levels)
        pred = tg.plane(np.broadcast_to(origins, (n, self.dim)), np.broadcast_to(normals, (n, self.dim)))
        self._refine_geometric(pred, n, levels, finalize, diagonal_balance)

    def refine_triangle(self, triangle, levels, finalize=True, diagonal_balance=None):
        """reference: tree_ext.pyx:897-972."""
        from . import tree_geometry as tg

        triangle = self._array_with_dim("triangle", triangle, 3)
        if triangle.shape[-2] != 3:
            raise ValueError(f"triangle array must be (N, 3, {self.dim})")
        levels = np.atleast_1d(np.asarray(levels, dtype=np.int64))
        n = self._broadcast_first(triangle=triangle, levels=levels)
        self._refine_geometric(tg.triangle(np.broadcast_to(triangle, (n, 3, self.dim))), n, levels, finalize, diagonal_balance)

    def refine_vertical_trianglular_prism(self, triangle, h, levels, finalize=True, diagonal_balance=None):
        """reference: tree_ext.pyx:975-1069 (3-D only)."""
        from . import tree_geometry as tg

        if self.dim == 2:
            raise NotImplementedError("refine_vertical_trianglular_prism only implemented in 3D.")
        triangle = self._array_with_dim("triangle", triangle, 3)
        if triangle.shape[-2] != 3:
            raise ValueError(f"triangle array must be (N, 3, {self.dim})")
        h = np.atleast_1d(np.asarray(h, dtype=np.float64))
        if np.any(h < 0):
            raise ValueError("All heights must be positive.")
        levels = np.atleast_1d(np.asarray(levels, dtype=np.int64))
        n = self._broadcast_first(triangle=triangle, h=h, levels=levels)
        pred = tg.vertical_triangular_prism(np.broadcast_to(triangle, (n, 3, 3)), np.broadcast_to(h, (n,)))
        self._refine_geometric(pred, n, levels, finalize, diagonal_balance)

    def refine_tetrahedron(self, tetra, levels, finalize=True, diagonal_balance=None):
        """reference: tree_ext.pyx:1072-1150 (3-D; in 2-D the reference refines the triangle)."""
        from . import tree_geometry as tg

        if self.dim == 2:
            return self.refine_triangle(tetra, levels, finalize=finalize, diagonal_balance=diagonal_balance)
        tetra = self._array_with_dim("tetra", tetra, 3)
        if tetra.shape[-2] != 4:
            raise ValueError(f"tetra array must be (N, {self.dim + 1}, {self.dim})")
        levels = np.atleast_1d(np.asarray(levels, dtype=np.int64))
        n = self._broadcast_first(tetra=tetra, levels=levels)
        self._refine_geometric(tg.tetrahedron(np.broadcast_to(tetra, (n, 4, 3))), n, levels, finalize, diagonal_balance)

    def refine_surface(self, xyz, level=-1, padding_cells_by_level=None, pad_up=False, pad_down=True, finalize=True,
                       diagonal_balance=None):
        """Refine along the surface triangulated from ``xyz`` (or ``(xyz, simplices)``), padded per level by stretching
        its bounding box and thickening it downward / upward (tree_mesh.py:638-797)."""
        level = self._parse_level(self._wrap_level(level))
        if padding_cells_by_level is None:
            padding_cells_by_level = np.array([0])
        padding_cells_by_level = np.asarray(padding_cells_by_level)
        if padding_cells_by_level.ndim == 0 and level > 1:
            padding_cells_by_level = np.full(level - 1, padding_cells_by_level)
        padding_cells_by_level = np.atleast_1d(padding_cells_by_level)
        if padding_cells_by_level.ndim == 2 and padding_cells_by_level.shape[1] != self.dim:
            raise ValueError("incorrect dimension for padding_cells_by_level.")
        if self.dim == 2:
            xyz = np.asarray(xyz)
            xyz = xyz[np.argsort(xyz[:, 0])]
            n_ps = len(xyz)
            inds = np.arange(n_ps)
            # each segment of the polyline becomes two triangles between the curve and its thickened copy
            simps = np.r_[np.c_[inds[:-1], inds[1:], inds[:-1]] + [0, 0, n_ps],
                          np.c_[inds[:-1], inds[1:], inds[1:]] + [n_ps, n_ps, 0]]
        else:
            if isinstance(xyz, tuple):
                xyz, simps = np.asarray(xyz[0]), np.asarray(xyz[1])
            else:
                from scipy.spatial import Delaunay

                xyz = np.asarray(xyz)
                simps = Delaunay(xyz[:, :2]).simplices
            n_ps = len(xyz)
        bb_min, bb_max = np.min(xyz, axis=0)[:-1], np.max(xyz, axis=0)[:-1]
        half_width, center = (bb_max - bb_min) / 2, (bb_max + bb_min) / 2
        points = np.empty((n_ps, self.dim))
        points[:, -1] = xyz[:, -1]
        h_min = np.r_[[h.min() for h in self.h]]
        pad = 0.0
        for lv, n_pad in zip(np.arange(level, 1, -1), padding_cells_by_level):
            pad = pad + n_pad * h_min * 2 ** (self.max_level - lv)
            h = 0
            if pad_up:
                h += pad[-1]
            if pad_down:
                h += pad[-1]
                points[:, -1] = xyz[:, -1] - pad[-1]
            points[:, :-1] = (half_width + pad[:-1]) / half_width * (xyz[:, :-1] - center) + center
            if self.dim == 2:
                self.refine_triangle(np.r_[points, points + [0, h]][simps], lv, finalize=False,
                                     diagonal_balance=diagonal_balance)
            else:
                self.refine_vertical_trianglular_prism(points[simps], h, lv, finalize=False,
                                                       diagonal_balance=diagonal_balance)
        if finalize:
            self.finalize()

    def _find_cells(self, predicate):
        """Indices of the leaves whose box satisfies ``predicate`` (one primitive), ascending -- the traversal order of
        ``find_cells_geom`` (tree.h:170-182, 232-242) is the cell order."""
        self._t()
        tr = self._tree
        from .octree import _child_offsets

        offs = _child_offsets(self.dim)
        c = tr.root_cells()
        found = []
        for l in range(tr.root_level, tr.max_level + 1):
            if not c.shape[0]:
                break
            w = tr.level_width(l)
            a, b = self._cell_box(c * w, w)
            c = c[predicate(np.zeros(c.shape[0], dtype=np.int64), a, b)]
            if l < tr.max_level and tr.div[l].size:
                isdiv = np.isin(tr.pack(l, c), tr.div[l])
            else:
                isdiv = np.zeros(c.shape[0], dtype=bool)
            found.append(c[~isdiv] * w)
            c = (2 * c[isdiv][:, None, :] + offs[None, :, :]).reshape(-1, self.dim)
        lo = np.concatenate(found) if found else np.zeros((0, self.dim), dtype=np.int64)
        all_lo, _ = tr.leaves()
        return np.sort(np.searchsorted(tr.keys(all_lo), tr.keys(lo)))

    def get_cells_in_ball(self, center, radius):
        from . import tree_geometry as tg

        return self._find_cells(tg.ball(self._array_with_dim("center", center, 2), np.array([float(radius)])))

    def get_cells_on_line(self, segment):
        from . import tree_geometry as tg

        segment = self._array_with_dim("segment", segment, 2)
        if segment.shape[0] != 2:
            raise ValueError(f"A line segment has two points, not {segment.shape[0]}")
        return self._find_cells(tg.line(segment[:1], segment[1:]))

    def get_cells_in_aabb(self, x_min, x_max):
        from . import tree_geometry as tg

        return self._find_cells(tg.box(self._array_with_dim("x_min", x_min, 2), self._array_with_dim("x_max", x_max, 2)))

    def get_cells_on_plane(self, origin, normal):
        from . import tree_geometry as tg

        return self._find_cells(tg.plane(self._array_with_dim("origin", origin, 2), self._array_with_dim("normal", normal, 2)))

    def get_cells_in_triangle(self, triangle):
        from . import tree_geometry as tg

        triangle = self._array_with_dim("triangle", triangle, 2)
        if triangle.shape[0] != 3:
            raise ValueError(f"Triangle array must have three points, saw {triangle.shape[0]}")
        return self._find_cells(tg.triangle(triangle[None]))

    def get_cells_in_vertical_trianglular_prism(self, triangle, h):
        from . import tree_geometry as tg

        if self.dim == 2:
            raise NotImplementedError("vertical_trianglular_prism only valid in 3D")
        triangle = self._array_with_dim("triangle", triangle, 2)
        if triangle.shape[0] != 3:
            raise ValueError(f"Triangle array must have three points, saw {triangle.shape[0]}")
        if h < 0:
            raise ValueError("h must be greater than or equal to 0")
        return self._find_cells(tg.vertical_triangular_prism(triangle[None], np.array([float(h)])))

    def get_cells_in_tetrahedron(self, tetra):
        from . import tree_geometry as tg

        if self.dim == 2:
            return self.get_cells_in_triangle(tetra)
        tetra = self._array_with_dim("tetra", tetra, 2)
        if tetra.shape[0] != 4:
            raise ValueError(f"A tetrahedron is defined by 4 points in 3D, not {tetra.shape[0]}.")
        return self._find_cells(tg.tetrahedron(tetra[None]))

    def get_containing_cells(self, points):
        """Index of the cell containing each point; a single point gives a scalar (tree_ext.pyx:5596-5628)."""
        return self.point2index(points)

    def get_overlapping_cells(self, rectangle):
        """Cells intersecting the box ``[xmin, xmax, ymin, ymax, (zmin, zmax)]`` (tree_ext.pyx:7055-7068)."""
        r = np.asarray(rectangle, dtype=np.float64)
        return self.get_cells_in_aabb(r[0::2], r[1::2])

    # ---- UBC-GIF files (mixins/mesh_io.py:440-615; ubc_io.py) ----------------------------------------------------
    @classmethod
    def read_UBC(cls, file_name, directory=None):
        from . import ubc_io

        return ubc_io.read_tree_mesh(TreeMesh, file_name, directory)      # TreeMesh(...) picks Quad / OcTree

    def write_UBC(self, file_name, models=None, directory=None):
        from . import ubc_io

        ubc_io.write_tree_mesh(self, file_name, models, directory)

    def read_model_UBC(self, file_name, directory=None):
        from . import ubc_io

        return ubc_io.read_tree_model(self, file_name, directory)

    def write_model_UBC(self, file_name, model, directory=None):
        from . import ubc_io

        ubc_io.write_tree_model(self, file_name, model, directory)

    _ubc_order = property(lambda self: __import__("discretize_b200.ubc_io", fromlist=["x"]).tree_ubc_order(self))

    # ---- remaining small pieces of the reference's TreeMesh ----------------------------------------------------------
    def refine_image(self, image, finalize=True, diagonal_balance=None):
        """Divide every cell whose block of the base-mesh image is not constant (tree_ext.pyx:1217-1264,
        tree.cpp:421-457).  A block that is constant has constant sub-blocks, so the set of divided cells does not depend
        on the traversal: per level, a min / max pooling of the image marks the boxes to divide."""
        self._require_open()
        db = self._diagonal_balance if diagonal_balance is None else diagonal_balance
        image = np.asarray(image, dtype=np.float64)
        n_expected = int(np.prod(self.shape_cells))
        if image.size != n_expected:
            raise ValueError(f"image array size: {image.size} must match the total number of cells in the base tensor "
                             f"mesh: {n_expected}")
        if image.ndim == 1:
            image = image.reshape(self.shape_cells, order="F")
        if image.shape != tuple(self.shape_cells):
            raise ValueError(f"image array shape: {image.shape} must match the base cell shapes: {self.shape_cells}")
        t = self._tree
        for l in range(t.root_level, t.max_level):
            w = t.level_width(l)
            g = t.grid(l)
            blocks = image.reshape([x for n in g for x in (n, w)])
            axes = tuple(range(1, 2 * self.dim, 2))
            mixed = blocks.max(axis=axes) != blocks.min(axis=axes)
            coords = np.stack(np.nonzero(mixed), axis=1).astype(np.int64)
            if coords.shape[0]:
                t._add_divided(l, np.unique(t.pack(l, coords)))
        t.balance(db)
        if finalize:
            self.finalize()

    @property
    def fill(self):
        """Number of cells over the number of cells of the underlying tensor mesh (tree_ext.pyx:1590-1608)."""
        return float(self.n_cells) / float(np.prod(self.shape_cells))

    def number(self):
        """Numbering happens in ``finalize``; kept for scripts that call it (tree_ext.pyx:1316-1318)."""
        self._t()

    def get_tensor(self, key):
        """1-D sample vectors of the UNDERLYING tensor grid for a location type (base_tensor_mesh.py:567-634)."""
        key = self._parse_location_type(key)
        nodes = [self.nodes_x, self.nodes_y, self.nodes_z][:self.dim]
        centers = [(n[1:] + n[:-1]) / 2 for n in nodes]
        if key == "cell_centers":
            on_nodes = ()
        elif key == "nodes":
            on_nodes = tuple(range(self.dim))
        elif key.startswith("faces_"):
            on_nodes = ("xyz".index(key[-1]),)
        elif key.startswith("edges_"):
            on_nodes = tuple(a for a in range(self.dim) if a != "xyz".index(key[-1]))
        else:
            raise ValueError(f"Unrecognized location type {key}")
        return [nodes[a] if a in on_nodes else centers[a] for a in range(self.dim)]

    @property
    def nodal_laplacian(self):
        raise NotImplementedError("Nodal Laplacian has not been implemented for TreeMesh")

    def get_boundary_cells(self, active_ind=None, direction="zu"):
        """Cells on the outer boundary in ``direction`` ('xd', 'xu', 'yd', 'yu', 'zd', 'zu'), or -- with ``active_ind`` -- the
        active cells whose neighbour(s) in that direction are missing or inactive (tree_ext.pyx:2923-2986)."""
        direction = direction.lower()
        if direction[0] == "z" and self.dim == 2:
            direction = "y" + direction[1]
        dir_ind = {"xd": 0, "xu": 1, "yd": 2, "yu": 3, "zd": 4, "zu": 5}[direction]
        if active_ind is None:
            return self.cell_boundary_indices[dir_ind]
        act = np.asarray(active_ind).astype(bool)
        axis, sign = dir_ind // 2, (1.0 if dir_ind % 2 else -1.0)
        cc, hg = self.cell_centers, self.h_gridded
        tang = [a for a in range(self.dim) if a != axis]
        bound = np.zeros(self.n_cells, dtype=bool)
        for corner in range(2 ** (self.dim - 1)):          # the (up to) 2 / 4 finer neighbours across the face
            p = cc.copy()
            p[:, axis] += sign * 0.75 * hg[:, axis]
            for k, a in enumerate(tang):
                p[:, a] += (0.25 if (corner >> k) & 1 else -0.25) * hg[:, a]
            inside = self.is_inside(p)
            nb = self._containing_cells(p)
            bound |= ~inside | ~act[nb]
        return np.where(bound & act)

    def get_cells_along_line(self, x0, x1):
        """Indices of the cells a segment passes through, in order from ``x0`` to ``x1`` (tree_ext.pyx:2989-3112): from
        the cell containing ``x0``, repeatedly leave through the face(s) the segment crosses first -- several at once when
        it leaves through an edge or a corner -- into the neighbour at the same or a coarser level, then descend to the
        leaf nearest the crossing point.  A short sequential walk; cells are handled as (level, integer coordinates)."""
        self._t()
        tr = self._tree
        dim = self.dim
        a = np.asarray(x0, dtype=np.float64).reshape(-1)
        b = np.asarray(x1, dtype=np.float64).reshape(-1)
        if "div_sets" not in self._cache:
            self._cache["div_sets"] = {l: set(tr.div[l].tolist()) for l in tr.div}
        divided = self._cache["div_sets"]
        leaf_lo, _ = tr.leaves()
        leaf_keys = tr.keys(leaf_lo)

        def pack(l, c):
            g = tr.grid(l)
            return c[0] + g[0] * c[1] if dim == 2 else c[0] + g[0] * (c[1] + g[1] * c[2])

        def is_divided(l, c):
            return l < tr.max_level and pack(l, c) in divided[l]

        def exists(l, c):
            return l == tr.root_level or is_divided(l - 1, tuple(x >> 1 for x in c))

        def neighbour(l, c, axis, sign):
            nb = list(c)
            nb[axis] += sign
            if not 0 <= nb[axis] < tr.grid(l)[axis]:
                return None
            nb = tuple(nb)
            while not exists(l, nb):                 # the coarser leaf that covers the box
                l, nb = l - 1, tuple(x >> 1 for x in nb)
            return l, nb

        def index_of(l, c):
            w = tr.level_width(l)
            lo = np.array([[x * w for x in c]], dtype=np.int64)
            return int(np.searchsorted(leaf_keys, tr.keys(lo))[0])

        def cell_of_point(p):
            i = int(self._containing_cells(p[None])[0])
            lo, lvl = tr.leaves()
            w = int(tr.width(lvl[i]))
            return int(lvl[i]), tuple(int(x) // w for x in lo[i])

        cur = cell_of_point(a)
        last = cell_of_point(b)
        out = [index_of(*cur)]
        while cur != last:
            l, c = cur
            w = tr.level_width(l)
            lo = np.array([self._xs[d][2 * c[d] * w] for d in range(dim)])
            hi = np.array([self._xs[d][2 * (c[d] + 1) * w] for d in range(dim)])
            t_axis = np.full(dim, np.inf)
            for d in range(dim):
                if a[d] > b[d]:
                    t_axis[d] = (lo[d] - a[d]) / (b[d] - a[d])
                elif a[d] < b[d]:
                    t_axis[d] = (hi[d] - a[d]) / (b[d] - a[d])
            t = t_axis.min()
            if t >= 1:
                break
            ip = (b - a) * t + a
            nxt = cur
            for d in range(dim):
                if nxt is not None and t == t_axis[d]:
                    nxt = neighbour(nxt[0], nxt[1], d, -1 if a[d] > b[d] else 1)
            if nxt is None:
                break
            l, c = nxt
            while is_divided(l, c):
                w = tr.level_width(l)
                centre = [self._xs[d][2 * c[d] * w + w] for d in range(dim)]
                bits = [int(ip[d] > centre[d] or (ip[d] == centre[d] and a[d] < b[d])) for d in range(dim)]
                l, c = l + 1, tuple(2 * c[d] + bits[d] for d in range(dim))
            cur = (l, c)
            out.append(index_of(l, c))
        return out

    # ---- cells as objects (tree_ext.pyx:6509-6533) ---------------------------------------------------------------
    def __len__(self):
        return self.n_cells

    def __bool__(self):
        return True          # a mesh is not "empty"; len() of an unfinalized tree must not be consulted for truth

    def __getitem__(self, key):
        if isinstance(key, slice):
            return [self[ii] for ii in range(*key.indices(len(self)))]
        if isinstance(key, (int, np.integer)):
            if key < 0:
                key += len(self)
            if key >= len(self):
                raise IndexError("The index ({0:d}) is out of range.".format(key))
            lo, lvl = self._tree.leaves()
            return TreeCell(self, lo[key], self._tree.width(lvl[key]), lvl[key], key)
        raise TypeError("Invalid argument type.")

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    # ---- persistence (tree_mesh.py:1085-1118, tree_ext.pyx:6472-6507) --------------------------------------------
    def __getstate__(self):
        idx, lv = self._tree.state()
        return idx, lv

    def __setstate__(self, state):
        """Rebuild from (cell-centre half-step indexes, levels).  The saved cells already carry whatever balancing the
        mesh had, so they are taken as they are (the reference re-inserts them with ``diagonal_balance=False``)."""
        indexes, levels = state
        self._set_state(np.asarray(indexes, dtype=np.int64).reshape(-1, self.dim), np.asarray(levels, dtype=np.int64))
        self.finalize()

    def __reduce__(self):
        return TreeMesh, (self.h, self.origin, False), self.__getstate__()

    def validate(self):
        return self.finalized

    def equals(self, other):
        try:
            if self.finalized and other.finalized:
                return super().equals(other)
        except AttributeError:
            pass
        return False

    def _set_state(self, indexes, levels):
        """Rebuild from (cell-centre half-step locations, levels), like __setstate__ (tree_ext.pyx:6488-6507)."""
        self._require_open()
        indexes = np.asarray(indexes, dtype=np.int64)
        levels = np.asarray(levels, dtype=np.int64)
        w = self._tree.width(levels)
        self._tree.set_leaves((indexes - w[:, None]) // 2, levels)

    @property
    def cell_state(self):
        idx, lv = self._tree.state()
        return {"indexes": idx.tolist(), "levels": lv.tolist()}

    def finalize(self):
        """reference: tree_ext.pyx:1266-1279 -> tree.cpp:949-1451."""
        if self._finalized:
            return
        centers, levels = self._tree.state()
        half = self._tree.width(levels)
        self._levels = levels
        self._topo = TreeTopology(centers, half, [2 * n for n in self.shape_cells])
        self._finalized = True

    def _t(self):
        if not self._finalized:
            raise TreeMeshNotFinalizedError("TreeMesh must be finalized before this property is available")
        return self._topo

    # ---- counts (tree_ext.pyx:1649-2019) ------------------------------------------------------------------
    n_cells = property(lambda self: self._t().n_cells)
    n_nodes = property(lambda self: self._t().nodes.n)
    n_total_nodes = property(lambda self: self._t().nodes.n_total)
    n_hanging_nodes = property(lambda self: self._t().nodes.n_hanging)
    n_edges_x = property(lambda self: self._t().edges[0].n)
    n_edges_y = property(lambda self: self._t().edges[1].n)
    n_edges_z = property(lambda self: self._t().edges[2].n)
    n_total_edges_x = property(lambda self: self._t().edges[0].n_total)
    n_total_edges_y = property(lambda self: self._t().edges[1].n_total)
    n_total_edges_z = property(lambda self: self._t().edges[2].n_total)
    n_hanging_edges_x = property(lambda self: self._t().edges[0].n_hanging)
    n_hanging_edges_y = property(lambda self: self._t().edges[1].n_hanging)
    n_hanging_edges_z = property(lambda self: self._t().edges[2].n_hanging)
    n_faces_x = property(lambda self: self._t().faces[0].n)
    n_faces_y = property(lambda self: self._t().faces[1].n)
    n_faces_z = property(lambda self: self._t().faces[2].n)
    n_total_faces_x = property(lambda self: self._t().faces[0].n_total)
    n_total_faces_y = property(lambda self: self._t().faces[1].n_total)
    n_total_faces_z = property(lambda self: self._t().faces[2].n_total)
    n_hanging_faces_x = property(lambda self: self._t().faces[0].n_hanging)
    n_hanging_faces_y = property(lambda self: self._t().faces[1].n_hanging)
    n_hanging_faces_z = property(lambda self: self._t().faces[2].n_hanging)
    n_edges = property(lambda self: sum(e.n for e in self._t().edges))
    n_faces = property(lambda self: sum(f.n for f in self._t().faces))
    n_total_edges = property(lambda self: sum(e.n_total for e in self._t().edges))
    n_total_faces = property(lambda self: sum(f.n_total for f in self._t().faces))
    n_hanging_edges = property(lambda self: sum(e.n_hanging for e in self._t().edges))
    n_hanging_faces = property(lambda self: sum(f.n_hanging for f in self._t().faces))
    n_edges_per_direction = property(lambda self: tuple(e.n for e in self._t().edges))
    n_faces_per_direction = property(lambda self: tuple(f.n for f in self._t().faces))

    def _cell_levels_by_indexes(self, index=None):
        """Levels of all cells (``index is None``) or of the given ones; one index gives a scalar (tree_ext.pyx:5630-5649)."""
        self._t()
        if index is None:
            out = self._levels.copy()
        else:
            out = self._levels[np.atleast_1d(np.asarray(index, dtype=np.int64))]
        return out[0] if out.size == 1 else out

    def cell_levels_by_index(self, indices):
        """reference: tree_mesh.py:988-1001."""
        return self._cell_levels_by_indexes(indices)

    # ---- geometry (tree.cpp:61-84, 109-136, 176-230; tree_ext.pyx:2021-2722) ---------------------------------
    def _coords(self, loc):
        return np.stack([self._xs[d][loc[:, d]] for d in range(3)], axis=1)

    def _extent(self, loc, half, d):
        return self._xs[d][loc[:, d] + half] - self._xs[d][loc[:, d] - half]

    def _midpoints(self, loc, half, axes):
        """Entity locations as the reference computes them: the mean of the corner coordinates
        (tree.cpp:73-75, 122-124) -- not the coordinate at the middle index when h is non-uniform."""
        out = self._coords(loc)
        for d in axes:
            out[:, d] = (self._xs[d][loc[:, d] - half] + self._xs[d][loc[:, d] + half]) * 0.5
        return out

    @property
    def cell_centers(self):
        t = self._t()
        return self._midpoints(t.centers, t.half, (0, 1, 2))

    @property
    def h_gridded(self):
        t = self._t()
        return np.stack([self._extent(t.centers, t.half, d) for d in range(3)], axis=1)

    @property
    def cell_volumes(self):
        hg = self.h_gridded
        return (hg[:, 0] * hg[:, 1]) * hg[:, 2]

    def _numbered(self, fam, values, include_hanging):
        out = np.empty(fam.n_total)
        out[fam.index] = values
        return out if include_hanging else out[:fam.n]

    def _edge_lengths_total(self, d):
        e = self._t().edges[d]
        return self._extent(e.loc, e.half, d)

    def _face_areas_total(self, d):
        f = self._t().faces[d]
        t1, t2 = [a for a in range(3) if a != d]
        return self._extent(f.loc, f.half, t1) * self._extent(f.loc, f.half, t2)

    @property
    def edge_lengths(self):
        t = self._t()
        return np.concatenate([self._numbered(t.edges[d], self._edge_lengths_total(d), False) for d in range(3)])

    @property
    def face_areas(self):
        t = self._t()
        return np.concatenate([self._numbered(t.faces[d], self._face_areas_total(d), False) for d in range(3)])

    def _face_locations(self, fam, axes):
        """(p1+p2+p3+p4)*0.25 in the reference's corner order (tree.cpp:109-124)."""
        out = self._coords(fam.loc)
        t1, t2 = axes
        a1, b1 = self._xs[t1][fam.loc[:, t1] - fam.half], self._xs[t1][fam.loc[:, t1] + fam.half]
        a2, b2 = self._xs[t2][fam.loc[:, t2] - fam.half], self._xs[t2][fam.loc[:, t2] + fam.half]
        out[:, t1] = (((a1 + b1) + a1) + b1) * 0.25
        out[:, t2] = (((a2 + a2) + b2) + b2) * 0.25
        return out

    def _grid(self, fam, include_hanging=False, axes=()):
        if len(axes) == 2:
            c = self._face_locations(fam, axes)
        else:
            c = self._midpoints(fam.loc, fam.half, axes) if axes else self._coords(fam.loc)
        out = np.empty_like(c)
        out[fam.index] = c
        return out if include_hanging else out[:fam.n]

    nodes = property(lambda self: self._grid(self._t().nodes))
    hanging_nodes = property(lambda self: self._grid(self._t().nodes, True)[self._t().nodes.n:])
    edges_x = property(lambda self: self._grid(self._t().edges[0], axes=(0,)))
    edges_y = property(lambda self: self._grid(self._t().edges[1], axes=(1,)))
    edges_z = property(lambda self: self._grid(self._t().edges[2], axes=(2,)))
    faces_x = property(lambda self: self._grid(self._t().faces[0], axes=(1, 2)))
    faces_y = property(lambda self: self._grid(self._t().faces[1], axes=(0, 2)))
    faces_z = property(lambda self: self._grid(self._t().faces[2], axes=(0, 1)))

    hanging_edges_x = property(lambda self: self._grid(self._t().edges[0], True, axes=(0,))[self._t().edges[0].n:])
    hanging_edges_y = property(lambda self: self._grid(self._t().edges[1], True, axes=(1,))[self._t().edges[1].n:])
    hanging_edges_z = property(lambda self: self._grid(self._t().edges[2], True, axes=(2,))[self._t().edges[2].n:])
    hanging_faces_x = property(lambda self: self._grid(self._t().faces[0], True, axes=(1, 2))[self._t().faces[0].n:])
    hanging_faces_y = property(lambda self: self._grid(self._t().faces[1], True, axes=(0, 2))[self._t().faces[1].n:])
    hanging_faces_z = property(lambda self: self._grid(self._t().faces[2], True, axes=(0, 1))[self._t().faces[2].n:])
    total_nodes = property(lambda self: np.vstack((self.nodes, self.hanging_nodes)))        # tree_mesh.py:801-814
    faces = property(lambda self: np.r_[self.faces_x, self.faces_y, self.faces_z])
    edges = property(lambda self: np.r_[self.edges_x, self.edges_y, self.edges_z])
    vntF = property(lambda self: [self.n_total_faces_x, self.n_total_faces_y, self.n_total_faces_z])
    vntE = property(lambda self: [self.n_total_edges_x, self.n_total_edges_y, self.n_total_edges_z])
    nodes_x = property(lambda self: self._xs[0][::2].copy())         # the underlying tensor grid (base_tensor_mesh.py:89-155)
    nodes_y = property(lambda self: self._xs[1][::2].copy())
    nodes_z = property(lambda self: self._xs[2][::2].copy())
    cell_centers_x = property(lambda self: (self.nodes_x[1:] + self.nodes_x[:-1]) / 2)
    cell_centers_y = property(lambda self: (self.nodes_y[1:] + self.nodes_y[:-1]) / 2)
    cell_centers_z = property(lambda self: (self.nodes_z[1:] + self.nodes_z[:-1]) / 2)

    @property
    def max_used_level(self):
        """Finest level any cell is refined to (tree_ext.pyx:1611-1626)."""
        self._t()
        return int(self._levels.max())

    def _unit_vectors(self, counts):
        out = np.zeros((int(sum(counts)), 3))
        o = np.r_[0, np.cumsum(counts)]
        for a in range(3):
            out[o[a]:o[a + 1], a] = 1.0
        return out

    face_normals = property(lambda self: self._unit_vectors(self.n_faces_per_direction))
    edge_tangents = property(lambda self: self._unit_vectors(self.n_edges_per_direction))

    @property
    def cell_bounds(self):
        """(n_cells, 6): x1, x2, y1, y2, z1, z2 of every cell (tree_ext.pyx:1302-1314)."""
        t = self._t()
        cols = []
        for d in range(3):
            cols += [self._xs[d][t.centers[:, d] - t.half], self._xs[d][t.centers[:, d] + t.half]]
        return np.stack(cols, axis=1)

    @property
    def cell_nodes(self):
        """(n_cells, 8) node numbers of every cell, hanging nodes included (tree_ext.pyx:6407-6430)."""
        t = self._t()
        return t.nodes.index[t.cell_nodes]

    @property
    def edge_nodes(self):
        """Per direction, (n_total_edges_d, 2) numbers of the two end nodes of every edge (tree_ext.pyx:6433-6470)."""
        t = self._t()
        unit = np.eye(3, dtype=np.int64)
        out = []
        for d in range(self.dim):
            e = t.edges[d]
            ends = np.empty((e.n_total, 2), dtype=np.int64)
            for k, sgn in enumerate((-1, 1)):
                q = e.loc + sgn * e.half[:, None] * unit[d]
                ends[e.index, k] = t.nodes.index[np.searchsorted(t.nodes.keys, key3(q[:, 0], q[:, 1], q[:, 2]))]
            out.append(ends)
        return tuple(out)

    def is_inside(self, pts, location_type="nodes", **kwargs):
        """Whether points are inside the domain (base_tensor_mesh.py:638-682, on the underlying tensor grid)."""
        pts = as_array_n_by_dim(pts, 3)
        tensors = [self.nodes_x, self.nodes_y, self.nodes_z]
        inside = np.ones(pts.shape[0], dtype=bool)
        for i, tensor in enumerate(tensors):
            tol = (tensor.max() - tensor.min()) * 1.0e-10       # TOL of base_tensor_mesh.py
            inside = inside & (pts[:, i] >= tensor.min() - tol) & (pts[:, i] <= tensor.max() + tol)
        return inside

    def _lexsorted(self, grids):
        o, out = 0, []
        for g in grids:
            out.append(np.lexsort(g.T) + o)
            o += g.shape[0]
        p = np.concatenate(out)
        from .csr import CsrOperator

        return CsrOperator(sp.csr_matrix((np.ones(p.size), (np.arange(p.size), p)), shape=(p.size, p.size)))

    # permutations that sort by x, then y, then z (tree_mesh.py:1037-1083)
    permute_cells = property(lambda self: self._lexsorted([self.cell_centers]))
    permute_faces = property(lambda self: self._lexsorted([self.faces_x, self.faces_y, self.faces_z]))
    permute_edges = property(lambda self: self._lexsorted([self.edges_x, self.edges_y, self.edges_z]))

    # ---- boundary (tree_ext.pyx:2104-2125, 2328-2370, 2543-2610, 5514-5628) ------------------------------------------
    def _boundary_masks(self):
        """Masks over the non-hanging faces / edges / nodes on the domain boundary, in the reference's numbering."""
        if "bmask" not in self._cache:
            t = self._t()
            lim = t.shape_half

            def at_limit(fam, axes):
                m = np.zeros(fam.n_total, dtype=bool)
                for a in axes:
                    m |= (fam.loc[:, a] == 0) | (fam.loc[:, a] == lim[a])
                out = np.empty(fam.n_total, dtype=bool)
                out[fam.index] = m
                return out[:fam.n]

            dims = range(self.dim)
            face = np.concatenate([at_limit(t.faces[d], (d,)) for d in dims])
            edge = np.concatenate([at_limit(t.edges[d], [a for a in dims if a != d]) for d in dims])
            node = at_limit(t.nodes, tuple(dims))
            self._cache["bmask"] = (face, edge, node)
        return self._cache["bmask"]

    def _boundary_indices(self):
        return tuple(np.nonzero(m)[0] for m in self._boundary_masks())

    def _selection(self, key, mask):
        if key not in self._cache:
            from .csr import CsrOperator

            idx = np.nonzero(mask)[0]
            self._cache[key] = CsrOperator(sp.csr_matrix((np.ones(idx.size), (np.arange(idx.size), idx)),
                                                         shape=(idx.size, mask.size)))
        return self._cache[key]

    project_face_to_boundary_face = property(lambda self: self._selection("P_bf", self._boundary_masks()[0]))
    project_edge_to_boundary_edge = property(lambda self: self._selection("P_be", self._boundary_masks()[1]))
    project_node_to_boundary_node = property(lambda self: self._selection("P_bn", self._boundary_masks()[2]))
    boundary_faces = property(lambda self: self.faces[self._boundary_masks()[0]])
    boundary_edges = property(lambda self: self.edges[self._boundary_masks()[1]])
    boundary_nodes = property(lambda self: self.nodes[self._boundary_masks()[2]])

    @property
    def boundary_face_outward_normals(self):
        t = self._t()
        normals = self.face_normals
        low = []
        for d in range(self.dim):
            f = t.faces[d]
            m = np.empty(f.n_total, dtype=bool)
            m[f.index] = f.loc[:, d] == 0
            low.append(m[:f.n])
        normals[np.concatenate(low)] *= -1
        return normals[self._boundary_masks()[0]]

    @property
    def boundary_face_scalar_integral(self):
        """diag(face_areas) P^T diag(+-1) (differential_operators.py:2183-2197)."""
        from .csr import CsrOperator

        bf = self._boundary_indices()[0]
        sign = self.boundary_face_outward_normals.sum(axis=1)
        return CsrOperator(sp.csr_matrix((self.face_areas[bf] * sign, (bf, np.arange(bf.size))), shape=(self.n_faces, bf.size)))

    def _boundary_face_rows(self, name):
        return self._host_matrix(name)[self._boundary_indices()[0]]

    # ---- host assembly of the operators (once per mesh) ---------------------------------------------------------
    def _deflate(self, fam, parents, weights):
        """R (n_total x n): identity on non-hanging, parent weights on hanging rows, applied until no
        hanging columns remain (tree_ext.pyx:3855-4087)."""
        nt, n = fam.n_total, fam.n
        hang = np.nonzero(fam.hanging)[0]
        rows = [fam.index[~fam.hanging]]
        cols = [fam.index[~fam.hanging]]
        vals = [np.ones(n)]
        for k, wgt in enumerate(weights):
            rows.append(fam.index[hang])
            cols.append(fam.index[parents[hang, k] if parents.ndim == 2 else parents[hang]])
            vals.append(np.full(hang.size, wgt))
        Rh = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nt, nt))
        R = Rh
        while R[:, n:].nnz:
            R = (R @ Rh).tocsr()
        return R[:, :n].tocsr()

    def _deflate_faces(self):
        if "Rf" not in self._cache:
            t = self._t()
            self._cache["Rf"] = sp.block_diag([self._deflate(t.faces[d], t.face_parents[d], [1.0]) for d in range(3)]).tocsr()
        return self._cache["Rf"]

    def _deflate_edges(self):
        if "Re" not in self._cache:
            t = self._t()
            self._cache["Re"] = sp.block_diag([self._deflate(t.edges[d], t.edge_parents[d], [0.5, 0.5]) for d in range(3)]).tocsr()
        return self._cache["Re"]

    def _deflate_nodes(self):
        if "Rn" not in self._cache:
            t = self._t()
            self._cache["Rn"] = self._deflate(t.nodes, t.node_parents, [0.25] * 4)
        return self._cache["Rn"]

    def _host_parts(self, name):
        """The operator as  sum_s diag(geom_s) @ M_s : ``(geoms, parts)`` with ``geoms`` a list of per-row
        arrays and ``parts`` a list of ``(s, M_s)`` (s = 0: no geometry, s >= 1: geoms[s-1]).  M_s holds
        pure multipliers (signs and products of hanging weights) -- what the SpMV kernel's code table stores."""
        key = "P_" + name
        if key in self._cache:
            return self._cache[key]
        t = self._t()
        nC = t.n_cells
        ntF = [f.n_total for f in t.faces]
        ntE = [e.n_total for e in t.edges]
        offF = np.r_[0, np.cumsum(ntF)]
        offE = np.r_[0, np.cumsum(ntE)]
        cells = np.arange(nC)
        unit = np.eye(3, dtype=np.int64)
        geoms, parts = [], []
        if name == "face_divergence":  # tree_ext.pyx:3199-3234: -+A_f/V_c = -+1/h_d(cell), then deflate faces
            hg = self.h_gridded
            R = self._deflate_faces()
            for d in range(3):
                rows, cols, vals = [], [], []
                for s, sign in ((0, -1.0), (1, 1.0)):
                    pos = t.cell_faces[d][:, s]
                    rows.append(cells); cols.append(offF[d] + t.faces[d].index[pos]); vals.append(np.full(nC, sign))
                M = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nC, offF[3]))
                geoms.append(1.0 / hg[:, d])
                parts.append((d + 1, M @ R))
        elif name == "edge_curl":  # tree_ext.pyx:3236-3359: +-L_e/A_f = +-1/(other side of the face), deflate edges
            nF = [f.n for f in t.faces]
            offFn = np.r_[0, np.cumsum(nF)]
            R = self._deflate_edges()
            g = [np.zeros(offFn[3]), np.zeros(offFn[3])]
            acc = [([], [], []), ([], [], [])]
            for d in range(3):
                f = t.faces[d]
                nh = np.nonzero(~f.hanging)[0]
                a1, a2 = (d + 1) % 3, (d + 2) % 3   # (curl e)_d = d_{a1} e_{a2} - d_{a2} e_{a1}
                r = offFn[d] + f.index[nh]
                g[0][r] = 1.0 / self._extent(f.loc[nh], f.half[nh], a1)
                g[1][r] = 1.0 / self._extent(f.loc[nh], f.half[nh], a2)
                for k, (de, dt, sgn) in enumerate(((a2, a1, 1.0), (a1, a2, -1.0))):
                    e = t.edges[de]
                    for s in (-1, 1):
                        q = f.loc[nh] + s * f.half[nh, None] * unit[dt]
                        pos = np.searchsorted(e.keys, key3(q[:, 0], q[:, 1], q[:, 2]))
                        acc[k][0].append(r); acc[k][1].append(offE[de] + e.index[pos]); acc[k][2].append(np.full(nh.size, sgn * s))
            for k in range(2):
                M = sp.csr_matrix((np.concatenate(acc[k][2]), (np.concatenate(acc[k][0]), np.concatenate(acc[k][1]))),
                                  shape=(offFn[3], offE[3]))
                geoms.append(g[k])
                parts.append((k + 1, M @ R))
        elif name == "nodal_gradient":  # tree_ext.pyx:3361-3464: -+1/L, deflate nodes
            nE = [e.n for e in t.edges]
            offEn = np.r_[0, np.cumsum(nE)]
            rows, cols, vals = [], [], []
            g = np.zeros(offEn[3])
            for d in range(3):
                e = t.edges[d]
                nh = np.nonzero(~e.hanging)[0]
                r = offEn[d] + e.index[nh]
                g[r] = 1.0 / self._extent(e.loc[nh], e.half[nh], d)
                for s in (-1, 1):
                    q = e.loc[nh] + s * e.half[nh, None] * unit[d]
                    pos = np.searchsorted(t.nodes.keys, key3(q[:, 0], q[:, 1], q[:, 2]))
                    rows.append(r); cols.append(t.nodes.index[pos]); vals.append(np.full(nh.size, float(s)))
            M = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(offEn[3], t.nodes.n_total))
            geoms.append(g)
            parts.append((1, M @ self._deflate_nodes()))
        elif name in ("face_x_divergence", "face_y_divergence", "face_z_divergence"):
            # tree_mesh.py:961-982: the column block of one face direction of face_divergence
            d = "xyz".index(name[5])
            g_all, p_all = self._host_parts("face_divergence")
            o = np.r_[0, np.cumsum([f.n for f in t.faces])]
            geoms.append(g_all[d])
            parts.append((1, p_all[d][1][:, o[d]:o[d + 1]]))
        elif name.startswith("stencil_cell_gradient") or name.startswith("average_cell_to_total_face_") \
                or name.startswith("average_cell_to_face") or name == "average_cell_vector_to_face":
            parts.append((0, self._host_cell_to_face(name)))
        elif name[:-2] in ("average_node_to_edge", "average_node_to_face") and name[-2] == "_":
            # tree_ext.pyx:4765-4875, 4920-5040: the row block of one direction
            d = "xyz".index(name[-1])
            o = np.r_[0, np.cumsum([f.n for f in (t.edges if "edge" in name else t.faces)])]
            parts.append((0, self._host_parts(name[:-2])[1][0][1][o[d]:o[d + 1]]))
        elif name == "average_node_to_edge":  # tree_ext.pyx:4765-4918: 1/2 at the two end nodes, deflate nodes
            nE = [e.n for e in t.edges]
            offEn = np.r_[0, np.cumsum(nE)]
            rows, cols = [], []
            for d in range(3):
                e = t.edges[d]
                nh = np.nonzero(~e.hanging)[0]
                r = offEn[d] + e.index[nh]
                for s in (-1, 1):
                    q = e.loc[nh] + s * e.half[nh, None] * unit[d]
                    rows.append(r); cols.append(t.nodes.index[np.searchsorted(t.nodes.keys, key3(q[:, 0], q[:, 1], q[:, 2]))])
            rows, cols = np.concatenate(rows), np.concatenate(cols)
            M = sp.csr_matrix((np.full(rows.size, 0.5), (rows, cols)), shape=(offEn[3], t.nodes.n_total))
            parts.append((0, M @ self._deflate_nodes()))
        elif name == "average_node_to_face":  # tree_ext.pyx:4920-5078: 1/4 at the four corner nodes, deflate nodes
            nF = [f.n for f in t.faces]
            offFn = np.r_[0, np.cumsum(nF)]
            rows, cols = [], []
            for d in range(3):
                f = t.faces[d]
                nh = np.nonzero(~f.hanging)[0]
                a1, a2 = (d + 1) % 3, (d + 2) % 3
                r = offFn[d] + f.index[nh]
                for s1 in (-1, 1):
                    for s2 in (-1, 1):
                        q = f.loc[nh] + s1 * f.half[nh, None] * unit[a1] + s2 * f.half[nh, None] * unit[a2]
                        rows.append(r); cols.append(t.nodes.index[np.searchsorted(t.nodes.keys, key3(q[:, 0], q[:, 1], q[:, 2]))])
            rows, cols = np.concatenate(rows), np.concatenate(cols)
            M = sp.csr_matrix((np.full(rows.size, 0.25), (rows, cols)), shape=(offFn[3], t.nodes.n_total))
            parts.append((0, M @ self._deflate_nodes()))
        elif name == "average_edge_to_face":  # tree_ext.pyx:4350-4440: 1/4 of the four bounding edges, deflate edges
            nF = [f.n for f in t.faces]
            offFn = np.r_[0, np.cumsum(nF)]
            rows, cols = [], []
            for d in range(3):
                f = t.faces[d]
                nh = np.nonzero(~f.hanging)[0]
                a1, a2 = (d + 1) % 3, (d + 2) % 3
                r = offFn[d] + f.index[nh]
                for de, dt in ((a2, a1), (a1, a2)):      # edges along `de`, displaced along `dt`
                    e = t.edges[de]
                    for s in (-1, 1):
                        q = f.loc[nh] + s * f.half[nh, None] * unit[dt]
                        rows.append(r); cols.append(offE[de] + e.index[np.searchsorted(e.keys, key3(q[:, 0], q[:, 1], q[:, 2]))])
            rows, cols = np.concatenate(rows), np.concatenate(cols)
            M = sp.csr_matrix((np.full(rows.size, 0.25), (rows, cols)), shape=(offFn[3], offE[3]))
            parts.append((0, M @ self._deflate_edges()))
        elif name in ("average_face_to_cell", "average_face_to_cell_vector") or name.startswith("average_face_"):
            parts.append((0, self._host_average(name, t.faces, t.cell_faces, offF, self._deflate_faces(), 0.5)))
        elif name in ("average_edge_to_cell", "average_edge_to_cell_vector") or name.startswith("average_edge_"):
            parts.append((0, self._host_average(name, t.edges, t.cell_edges, offE, self._deflate_edges(), 0.25)))
        elif name == "average_node_to_cell":  # tree_ext.pyx:4709-4762
            rows = np.repeat(cells, 8)
            cols = t.nodes.index[t.cell_nodes.reshape(-1)]
            M = sp.csr_matrix((np.full(rows.size, 0.125), (rows, cols)), shape=(nC, t.nodes.n_total)) @ self._deflate_nodes()
            parts.append((0, M))
        else:
            raise AttributeError(name)
        canon = []
        for s_, M in parts:
            M = sp.csr_matrix(M)
            M.sum_duplicates()
            M.sort_indices()
            canon.append((s_, M))
        self._cache[key] = (geoms, canon)
        return self._cache[key]

    def _host_matrix(self, name):
        """scipy CSR of a reference operator (canonical form), from the same parts the kernel uses."""
        key = "H_" + name
        if key not in self._cache:
            geoms, parts = self._host_parts(name)
            acc = None
            for s_, M in parts:
                term = M if s_ == 0 else sp.diags(geoms[s_ - 1]) @ M
                acc = term if acc is None else acc + term
            acc = sp.csr_matrix(acc)
            acc.sum_duplicates()
            acc.sort_indices()
            self._cache[key] = acc
        return self._cache[key]

    def _face_sides(self, d):
        """(minus cell, plus cell) of every TOTAL face of direction d (-1: none).  A hanging face inherits the side it
        lacks from its parent (the big neighbour lists the parent face); a face with hanging children keeps its
        missing side -- the reference writes the rows of the children instead (tree_ext.pyx:3641-3853)."""
        t = self._t()
        f = t.faces[d]
        cells = np.arange(t.n_cells)
        minus = np.full(f.n_total, -1, dtype=np.int64)
        plus = np.full(f.n_total, -1, dtype=np.int64)
        minus[t.cell_faces[d][:, 1]] = cells      # the cell whose + face this is sits on the minus side
        plus[t.cell_faces[d][:, 0]] = cells
        par = t.face_parents[d]
        h = np.nonzero(f.hanging)[0]
        need = h[minus[h] < 0]
        minus[need] = minus[par[need]]
        need = h[plus[h] < 0]
        plus[need] = plus[par[need]]
        return minus, plus

    def _host_cell_to_face(self, name):
        """Cell -> face operators of the cell-centred formulation (tree_ext.pyx:3471-3853, 5080-5512).

        All of them walk the (minus cell, plus cell) pairs across the total faces.  The reference fills
        preallocated zero triplet arrays, so every slot it does not write lands on entry (0, 0) with value 0:
        mirrored by an explicit zero there."""
        t = self._t()
        nC = t.n_cells
        axes = "xyz"[:self.dim]
        if name in ("stencil_cell_gradient", "average_cell_to_face"):
            return sp.vstack([self._host_cell_to_face(f"{name}_{c}") for c in axes]).tocsr()
        if name == "average_cell_vector_to_face":   # tree_ext.pyx:5131-5190
            return sp.block_diag([self._host_cell_to_face(f"average_cell_to_face_{c}") for c in axes]).tocsr()
        d = "xyz".index(name[-1])
        f = t.faces[d]
        minus, plus = self._face_sides(d)
        both = np.nonzero((minus >= 0) & (plus >= 0))[0]
        zero = (np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64), np.zeros(1))
        if name.startswith("stencil_cell_gradient_") or name.startswith("average_cell_to_total_face_"):
            lo, hi = (-1.0, 1.0) if name.startswith("stencil") else (0.5, 0.5)
            rows = np.r_[f.index[both], f.index[both], zero[0]]
            cols = np.r_[minus[both], plus[both], zero[1]]
            vals = np.r_[np.full(both.size, lo), np.full(both.size, hi), zero[2]]
            return sp.csr_matrix((vals, (rows, cols)), shape=(f.n_total, nC))
        # average_cell_to_face_{x,y,z}: distance weights, hanging faces accumulate a quarter (2-D: a half) each into
        # their parent, boundary faces take the adjacent cell
        xc = self.cell_centers[:, d]
        xf = self._xs[d][f.loc[both, d]]
        w = (xc[plus[both]] - xf) / (xc[plus[both]] - xc[minus[both]])
        hang = f.hanging[both]
        row = np.where(hang, f.index[t.face_parents[d][both]], f.index[both])
        k = np.where(hang, float(2 ** (self.dim - 1)), 1.0)
        bnd = np.nonzero(~f.hanging & ((minus < 0) ^ (plus < 0)) & ~self._has_hanging_children(d))[0]
        bcell = np.where(minus[bnd] >= 0, minus[bnd], plus[bnd])
        # reference quirk (tree_ext.pyx:5219-5236): a cell without a + neighbour writes its + face and `continue`s,
        # so the - boundary face of a cell that spans the whole axis is never written
        spans = (minus[bnd] < 0) & (t.centers[bcell, d] + t.half[bcell] == t.shape_half[d])
        bnd, bcell = bnd[~spans], bcell[~spans]
        rows = np.r_[row, row, f.index[bnd], zero[0]]
        cols = np.r_[minus[both], plus[both], bcell, zero[1]]
        vals = np.r_[w / k, (1.0 - w) / k, np.ones(bnd.size), zero[2]]
        return sp.csr_matrix((vals, (rows, cols)), shape=(f.n, nC))

    def _has_hanging_children(self, d):
        t = self._t()
        f = t.faces[d]
        out = np.zeros(f.n_total, dtype=bool)
        out[t.face_parents[d][f.hanging]] = True
        return out

    @property
    def cell_boundary_indices(self):
        """Cells touching the -x, +x, -y, +y, -z, +z boundary, in cell order (tree_ext.pyx:2700-2822)."""
        t = self._t()
        out = []
        for d in range(self.dim):
            out.append(np.nonzero(t.centers[:, d] - t.half == 0)[0])
            out.append(np.nonzero(t.centers[:, d] + t.half == t.shape_half[d])[0])
        return tuple(out)

    @property
    def face_boundary_indices(self):
        """Per-direction indices of the boundary faces, in the order of their cells (tree_ext.pyx:2824-2920)."""
        t = self._t()
        cb = self.cell_boundary_indices
        out = []
        for d in range(self.dim):
            out.append(t.faces[d].index[t.cell_faces[d][cb[2 * d], 0]])
            out.append(t.faces[d].index[t.cell_faces[d][cb[2 * d + 1], 1]])
        return tuple(out)

    def _host_average(self, name, fams, cell_ents, off, R, w):
        """tree_ext.pyx:4089-4707: per cell 2 faces x 1/2 (4 edges x 1/4) per direction, then deflation."""
        t = self._t()
        nC = t.n_cells
        cells = np.arange(nC)
        k = cell_ents[0].shape[1]
        blocks = []
        for d in range(3):
            rows = np.repeat(cells, k)
            cols = fams[d].index[cell_ents[d].reshape(-1)] + off[d]
            blocks.append(sp.csr_matrix((np.full(rows.size, w), (rows, cols)), shape=(nC, off[3])))
        comp = {"x": 0, "y": 1, "z": 2}
        if name.endswith("_to_cell_vector"):
            M = sp.vstack(blocks) @ R
        elif name.endswith("_x_to_cell") or name.endswith("_y_to_cell") or name.endswith("_z_to_cell"):
            d = comp[name[-9]]
            n_nh = [f.n for f in fams]
            o = np.r_[0, np.cumsum(n_nh)]
            M = (blocks[d] @ R)[:, o[d]:o[d + 1]]
        else:
            M = (1.0 / 3.0) * (blocks[0] + blocks[1] + blocks[2]) @ R
        return M

    # ---- operators -------------------------------------------------------------------------------------------
    def _operator(self, name):
        key = "OP_" + name
        if key not in self._cache:
            from .tree_ops import TreeOperator

            self._cache[key] = TreeOperator(self, name)
        return self._cache[key]

    face_divergence = property(lambda self: self._operator("face_divergence"))
    edge_curl = property(lambda self: self._operator("edge_curl"))
    nodal_gradient = property(lambda self: self._operator("nodal_gradient"))
    average_face_to_cell = property(lambda self: self._operator("average_face_to_cell"))
    average_face_to_cell_vector = property(lambda self: self._operator("average_face_to_cell_vector"))
    average_face_x_to_cell = property(lambda self: self._operator("average_face_x_to_cell"))
    average_face_y_to_cell = property(lambda self: self._operator("average_face_y_to_cell"))
    average_face_z_to_cell = property(lambda self: self._operator("average_face_z_to_cell"))
    average_edge_to_cell = property(lambda self: self._operator("average_edge_to_cell"))
    average_edge_to_cell_vector = property(lambda self: self._operator("average_edge_to_cell_vector"))
    average_edge_x_to_cell = property(lambda self: self._operator("average_edge_x_to_cell"))
    average_edge_y_to_cell = property(lambda self: self._operator("average_edge_y_to_cell"))
    average_edge_z_to_cell = property(lambda self: self._operator("average_edge_z_to_cell"))
    average_node_to_cell = property(lambda self: self._operator("average_node_to_cell"))
    face_x_divergence = property(lambda self: self._operator("face_x_divergence"))
    face_y_divergence = property(lambda self: self._operator("face_y_divergence"))
    face_z_divergence = property(lambda self: self._operator("face_z_divergence"))
    stencil_cell_gradient = property(lambda self: self._operator("stencil_cell_gradient"))
    stencil_cell_gradient_x = property(lambda self: self._operator("stencil_cell_gradient_x"))
    stencil_cell_gradient_y = property(lambda self: self._operator("stencil_cell_gradient_y"))
    stencil_cell_gradient_z = property(lambda self: self._operator("stencil_cell_gradient_z"))
    average_cell_to_face = property(lambda self: self._operator("average_cell_to_face"))
    average_cell_to_face_x = property(lambda self: self._operator("average_cell_to_face_x"))
    average_cell_to_face_y = property(lambda self: self._operator("average_cell_to_face_y"))
    average_cell_to_face_z = property(lambda self: self._operator("average_cell_to_face_z"))
    average_cell_vector_to_face = property(lambda self: self._operator("average_cell_vector_to_face"))

    # plain methods in the reference (tree_ext.pyx:3471-3640 carry no @property)
    def average_cell_to_total_face_x(self):
        return self._operator("average_cell_to_total_face_x")

    def average_cell_to_total_face_y(self):
        return self._operator("average_cell_to_total_face_y")

    def average_cell_to_total_face_z(self):
        return self._operator("average_cell_to_total_face_z")

    def _cell_gradient_component(self, d):
        """-Pa Mf^-1 D_d^T diag(V) (tree_mesh.py:860-958): composed lazily from the CUDA operators; d = None: all faces."""
        from .operators import DiagonalOperator

        key = "OP_cell_gradient" + ("" if d is None else "_" + "xyz"[d])
        if key not in self._cache:
            t = self._t()
            nF = [f.n for f in t.faces]
            o = np.r_[0, np.cumsum(nF)]
            fb = self.face_boundary_indices
            mask = np.ones(o[3])
            for dd in range(3):
                mask[o[dd] + fb[2 * dd]] = 0.0
                mask[o[dd] + fb[2 * dd + 1]] = 0.0
            mfi = self.get_face_inner_product(invert_matrix=True).diag           # device diagonal of Mf(1)
            if d is None:
                D, pa, mi = self.face_divergence, mask, 1.0 / mfi
            else:
                D = self._operator("face_%s_divergence" % "xyz"[d])
                pa, mi = mask[o[d]:o[d + 1]], 1.0 / mfi[int(o[d]):int(o[d + 1])]
            self._cache[key] = -(DiagonalOperator(pa) @ DiagonalOperator(mi) @ D.T @ DiagonalOperator(self.cell_volumes))
        return self._cache[key]

    cell_gradient = property(lambda self: self._cell_gradient_component(None))
    cell_gradient_x = property(lambda self: self._cell_gradient_component(0))
    cell_gradient_y = property(lambda self: self._cell_gradient_component(1))
    cell_gradient_z = property(lambda self: self._cell_gradient_component(2))
    average_node_to_edge = property(lambda self: self._operator("average_node_to_edge"))
    average_node_to_face = property(lambda self: self._operator("average_node_to_face"))
    average_edge_to_face = property(lambda self: self._operator("average_edge_to_face"))
    average_node_to_edge_x = property(lambda self: self._operator("average_node_to_edge_x"))
    average_node_to_edge_y = property(lambda self: self._operator("average_node_to_edge_y"))
    average_node_to_edge_z = property(lambda self: self._operator("average_node_to_edge_z"))
    average_node_to_face_x = property(lambda self: self._operator("average_node_to_face_x"))
    average_node_to_face_y = property(lambda self: self._operator("average_node_to_face_y"))
    average_node_to_face_z = property(lambda self: self._operator("average_node_to_face_z"))

    # ---- mass matrices (shared _fastInnerProduct: base/base_tensor_mesh.py:792-863) -----------------------------
    def get_face_inner_product(self, model=None, invert_model=False, invert_matrix=False, do_fast=True, **kwargs):
        return self._inner_product("F", model, invert_model, invert_matrix, do_fast, kwargs)

    def get_edge_inner_product(self, model=None, invert_model=False, invert_matrix=False, do_fast=True, **kwargs):
        return self._inner_product("E", model, invert_model, invert_matrix, do_fast, kwargs)

    def _inner_product(self, proj, model, invert_model, invert_matrix, do_fast, kwargs):
        from .tensor_mesh import TensorMesh
        from .tree_ops import make_tree_mass_operator

        TensorMesh._check_removed_kwargs(kwargs)
        return make_tree_mass_operator(self, proj, model, invert_model, invert_matrix)

    def get_face_inner_product_deriv(self, model, do_fast=True, invert_model=False, invert_matrix=False, **kwargs):
        from .tree_ops import make_tree_mass_deriv

        return make_tree_mass_deriv(self, "F", model, do_fast, invert_model, invert_matrix)

    def get_edge_inner_product_deriv(self, model, do_fast=True, invert_model=False, invert_matrix=False, **kwargs):
        from .tree_ops import make_tree_mass_deriv

        return make_tree_mass_deriv(self, "E", model, do_fast, invert_model, invert_matrix)

    # ---- point location and interpolation (node variant) -----------------------------------------------
    # ---- weak-form boundary pieces and face / edge hosted properties: shared with TensorMesh -----------------------
    def project_face_vector(self, face_vectors):
        if not isinstance(face_vectors, np.ndarray):
            raise Exception("face_vectors must be an ndarray")
        if not (face_vectors.ndim == 2 and face_vectors.shape == (self.n_faces, self.dim)):
            raise Exception("face_vectors must be an ndarray of shape (n_faces, dim)")
        return np.sum(face_vectors * self.face_normals, 1)

    def project_edge_vector(self, edge_vectors):
        if not isinstance(edge_vectors, np.ndarray):
            raise Exception("edge_vectors must be an ndarray")
        if not (edge_vectors.ndim == 2 and edge_vectors.shape == (self.n_edges, self.dim)):
            raise Exception("edge_vectors must be an ndarray of shape (nE, dim)")
        return np.sum(edge_vectors * self.edge_tangents, 1)

    def closest_points_index(self, locations, grid_loc="CC", discard=False):
        from scipy.spatial import KDTree

        locations = as_array_n_by_dim(locations, self.dim)
        grid_loc = self._parse_location_type(grid_loc)
        key = f"_{grid_loc}_tree"
        tree = self._cache.get(key)
        if tree is None:
            tree = KDTree(getattr(self, grid_loc))
        if not discard:
            self._cache[key] = tree
        return tree.query(locations)[1]

    def edge_divergence_weak_form_robin(self, alpha=0.0, beta=1.0, gamma=0.0):
        from . import boundary_ops

        return boundary_ops.edge_divergence_weak_form_robin(self, alpha, beta, gamma)

    def cell_gradient_weak_form_robin(self, alpha=0.0, beta=1.0, gamma=0.0):
        from . import boundary_ops

        return boundary_ops.cell_gradient_weak_form_robin(self, alpha, beta, gamma)

    @property
    def boundary_edge_vector_integral(self):
        from . import boundary_ops

        return boundary_ops.boundary_edge_vector_integral(self)

    @property
    def boundary_node_vector_integral(self):
        from . import boundary_ops

        return boundary_ops.boundary_node_vector_integral(self)

    def get_face_inner_product_surface(self, model=None, invert_model=False, invert_matrix=False, **kwargs):
        from . import surface_products

        return surface_products.inner_product(self, "face_surface", model, invert_model, invert_matrix)

    def get_edge_inner_product_surface(self, model=None, invert_model=False, invert_matrix=False, **kwargs):
        from . import surface_products

        return surface_products.inner_product(self, "edge_surface", model, invert_model, invert_matrix)

    def get_edge_inner_product_line(self, model=None, invert_model=False, invert_matrix=False, **kwargs):
        from . import surface_products

        return surface_products.inner_product(self, "edge_line", model, invert_model, invert_matrix)

    def get_face_inner_product_surface_deriv(self, model, invert_model=False, invert_matrix=False, **kwargs):
        from . import surface_products

        return surface_products.inner_product_deriv(self, "face_surface", model, invert_model, invert_matrix)

    def get_edge_inner_product_surface_deriv(self, model, invert_model=False, invert_matrix=False, **kwargs):
        from . import surface_products

        return surface_products.inner_product_deriv(self, "edge_surface", model, invert_model, invert_matrix)

    def get_edge_inner_product_line_deriv(self, model, invert_model=False, invert_matrix=False, **kwargs):
        from . import surface_products

        return surface_products.inner_product_deriv(self, "edge_line", model, invert_model, invert_matrix)

    def _containing_cells(self, points):
        """Index of the leaf that contains (or is closest to) each point -- tree.cpp:758-766, 1499-1516: the root is
        found with ``x >= upper edge -> next root`` (ties go high), the descent with ``x > centre -> high child``
        (ties go low); points outside the mesh end in a boundary cell.  Vectorised level by level."""
        tr = self._tree
        pts = np.asarray(points, dtype=np.float64)
        n = pts.shape[0]
        if n >= self._device_locate_min_points and self.dim == 3 and _cuda_ready():
            return self.containing_cells_device(pts).cpu().numpy()
        c = np.zeros((n, 3), dtype=np.int64)
        for d in range(3):
            edges = self._xs[d][2 * tr.rw * np.arange(1, tr.roots[d])]          # interior root boundaries
            c[:, d] = np.searchsorted(edges, pts[:, d], side="right")
        lo = np.zeros((n, 3), dtype=np.int64)
        done = np.zeros(n, dtype=bool)
        for l in range(tr.root_level, tr.max_level + 1):
            w = tr.level_width(l)
            act = np.nonzero(~done)[0]
            if not act.size:
                break
            if l < tr.max_level and tr.div[l].size:
                isdiv = np.isin(tr.pack(l, c[act]), tr.div[l])
            else:
                isdiv = np.zeros(act.size, dtype=bool)
            leaf = act[~isdiv]
            lo[leaf] = c[leaf] * w
            done[leaf] = True
            go = act[isdiv]
            if go.size:
                for d in range(3):
                    mid = self._xs[d][2 * c[go, d] * w + w]                       # the node in the middle of the cell
                    c[go, d] = 2 * c[go, d] + (pts[go, d] > mid)
        leaf_lo, _ = tr.leaves()
        return np.searchsorted(tr.keys(leaf_lo), tr.keys(lo))

    _device_locate_min_points = 1 << 16     # below this the host descent is faster than the transfers

    def containing_cells_device(self, points):
        """``_containing_cells`` on the GPU (``dzb_tree_locate``, csrc/tree_locate.cu): int64 CUDA tensor of cell indices
        for an (n, 3) array or CUDA tensor of points.  The reference's level-by-level descent is replaced by two
        bisections per axis, a Morton key and one bisection over the sorted leaf keys -- same cells, ties included."""
        import ctypes as C

        import torch

        from ._lib import check, load
        from .operators import ptr, stream_ptr, to_device

        if self.dim != 3:
            raise NotImplementedError("device point location is built for the OcTree")
        self._t()
        tr = self._tree
        dev = self._dev.get("locate")
        if dev is None:
            leaf_lo, _ = tr.leaves()
            dev = {"nodes": [to_device(np.ascontiguousarray(self._xs[d][::2])) for d in range(3)],
                   "keys": to_device(np.ascontiguousarray(tr.keys(leaf_lo).astype(np.int64)), torch.int64)}
            self._dev["locate"] = dev
        pts = points if isinstance(points, torch.Tensor) else np.ascontiguousarray(points, dtype=np.float64)
        pts = to_device(pts).reshape(-1, 3).contiguous()
        out = torch.empty(pts.shape[0], dtype=torch.int64, device=pts.device)
        roots = (C.c_int32 * 3)(*[int(r) for r in tr.roots])
        check(load().dzb_tree_locate(ptr(dev["nodes"][0]), ptr(dev["nodes"][1]), ptr(dev["nodes"][2]), roots, int(tr.rw),
                                     ptr(dev["keys"]), int(dev["keys"].numel()), ptr(pts), int(pts.shape[0]), ptr(out),
                                     stream_ptr()), "dzb_tree_locate")
        return out

    def point2index(self, locs):
        """reference: tree_mesh.py:984-986 (index of the cell containing each point)."""
        out = self._containing_cells(as_array_n_by_dim(np.asarray(locs, dtype=np.float64), 3))
        return out[0] if out.size == 1 else out      # one point gives a scalar (tree_ext.pyx:1350-1352)

    def get_interpolation_matrix(self, loc, location_type="cell_centers", zeros_outside=False, **kwargs):
        """Interpolation from nodes (``_getNodeIntMat``, tree_ext.pyx:6116-6179), edges, faces or cell centres
        (tree_interp.py; tree_ext.pyx:5765-6404) to arbitrary points, hanging entities eliminated; routing and errors of
        tree_mesh.py:1003-1034.  Assembled on the host, applied on the GPU by ``dzb_csr_spmv``."""
        if "locType" in kwargs:
            raise TypeError("The locType keyword argument has been removed, please use location_type. "
                            "This will be removed in discretize 1.0.0")
        if "zerosOutside" in kwargs:
            raise TypeError("The zerosOutside keyword argument has been removed, please use zeros_outside. "
                            "This will be removed in discretize 1.0.0")
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        from . import tree_interp
        from .csr import CsrOperator

        location_type = self._parse_location_type(location_type)
        pts = as_array_n_by_dim(np.asarray(loc, dtype=np.float64), 3)
        if location_type == "nodes":
            A = self._host_node_interpolation(pts, zeros_outside)
        elif location_type in ("edges_x", "edges_y", "edges_z"):
            A = tree_interp.edge_interpolation(self, pts, zeros_outside, location_type[-1])
        elif location_type in ("faces_x", "faces_y", "faces_z"):
            A = tree_interp.face_interpolation(self, pts, zeros_outside, location_type[-1])
        elif location_type == "cell_centers":
            A = tree_interp.cell_interpolation(self, pts, zeros_outside)
        else:
            raise ValueError("Location must be a grid location, not {}".format(location_type))
        return CsrOperator(A)

    def _host_node_interpolation(self, loc, zeros_outside):
        t = self._t()
        pts = as_array_n_by_dim(np.asarray(loc, dtype=np.float64), 3)
        npts = pts.shape[0]
        cell = self._containing_cells(pts)
        cen, hw = t.centers[cell], t.half[cell]
        w = []
        for d in range(3):
            x0, x1 = self._xs[d][cen[:, d] - hw], self._xs[d][cen[:, d] + hw]
            w.append((x1 - pts[:, d]) / (x1 - x0))
        eps = 100 * np.finfo(float).eps
        outside = np.zeros(npts, dtype=bool)
        if zeros_outside:
            for wd in w:
                outside |= (wd < -eps) | (wd > 1 + eps)
        w = [np.clip(wd, 0.0, 1.0) for wd in w]
        rows = np.repeat(np.arange(npts), 8)
        cols = t.nodes.index[t.cell_nodes[cell]]                                  # (npts, 8), x fastest
        vals = np.empty((npts, 8))
        for ii in range(8):
            fx = w[0] if not (ii & 1) else 1.0 - w[0]
            fy = w[1] if not (ii & 2) else 1.0 - w[1]
            fz = w[2] if not (ii & 4) else 1.0 - w[2]
            vals[:, ii] = fx * fy * fz
        cols = np.where(outside[:, None], 0, cols)
        vals = np.where(outside[:, None], 0.0, vals)
        A = sp.csr_matrix((vals.reshape(-1), (rows, cols.reshape(-1))), shape=(npts, t.nodes.n_total))
        return (A @ self._deflate_nodes()).tocsr()

    def __repr__(self):
        n = self._topo.n_cells if self._finalized else self._tree.n_leaves
        return f"<discretize_b200.TreeMesh {n} cells, max_level {self.max_level}>"

    def __str__(self):
        """Cells per level, extent and cell widths per axis; usable before ``finalize`` (counts from the leaf list)."""
        kind = "OcTree" if self.dim == 3 else "QuadTree"
        lo, lvl = self._tree.leaves()
        state = "" if self._finalized else ", not finalized"
        fill = 100.0 * lvl.size / float(np.prod(self.shape_cells))
        lines = [f"TreeMesh ({kind}): {lvl.size:,} cells, {fill:.2f}% of the {' x '.join(str(n) for n in self.shape_cells)} "
                 f"base grid{state}", "  level   cells"]
        for level, count in zip(*np.unique(lvl, return_counts=True)):
            lines.append(f"  {int(level):>5}{int(count):>8,}")
        lines.append(f"  {'axis':<5}{'from':>14}{'to':>14}{'h min':>12}{'h max':>12}")
        w = self._tree.width(lvl)
        for a, name in enumerate("xyz"[:self.dim]):
            x = self._xs[a]
            h = x[2 * (lo[:, a] + w)] - x[2 * lo[:, a]]
            lines.append(f"  {name:<5}{x[0]:>14.6g}{x[-1]:>14.6g}{h.min():>12.6g}{h.max():>12.6g}")
        return "\n".join(lines)

"""QuadTree: the 2-D ``TreeMesh``.

Reference: ``discretize/tree_mesh.py`` + the ``dim == 2`` branches of ``_extensions/tree_ext.pyx`` and ``tree.cpp``.
``TreeMesh([hx, hy])`` returns this class.  The constructor and every refinement method are the dimension-generic ones
of :class:`~discretize_b200.tree_mesh.TreeMesh` (``octree.LinearOctree`` handles both 2-D and 3-D); connectivity comes from
``quad_build.QuadTopology``.  2-D problems are small, so -- like 1-D / 2-D ``TensorMesh`` (``lowdim.py``) -- the operators
have no kernels of their own: each matrix is assembled on the host with array passes, exactly as the reference defines it
(entity numbering, hanging entities eliminated by the deflation matrices), and applied on the GPU by ``dzb_csr_spmv``
(:class:`~discretize_b200.csr.CsrOperator`); ``.tocsr()`` is the reference's matrix.

Layout facts of the reference that everything below follows (``tree_ext.pyx:1949-2019, 2655-2656, 3173-3197, 3976-3979``):
x-faces ARE y-edges and y-faces ARE x-edges, so ``faces = [edges_y; edges_x]``, ``face_areas`` are edge lengths,
``_deflate_faces = blockdiag(R_ey, R_ex)``; ``edge_curl`` maps edges to CELLS.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .quad_build import QuadTopology
from .tensor_mesh import as_array_n_by_dim, is_scalar
from .tree_build import key3
from .tree_mesh import TreeMesh


def _csr(rows, cols, vals, shape):
    return sp.csr_matrix((np.asarray(vals, dtype=np.float64), (rows, cols)), shape=shape)


def _canon(m):
    m = sp.csr_matrix(m)
    m.sum_duplicates()
    m.sort_indices()
    return m


class QuadTreeMesh(TreeMesh):
    dim = 2

    def __new__(cls, *args, **kwargs):
        return object.__new__(cls)

    # ---- finalize ---------------------------------------------------------------------------------------------------------
    def finalize(self):
        """reference: tree_ext.pyx:1266-1279 -> tree.cpp:949-1451 (2-D branches)."""
        if self._finalized:
            return
        centers, levels = self._tree.state()
        self._levels = levels
        self._topo = QuadTopology(centers, self._tree.width(levels), [2 * n for n in self.shape_cells])
        self._finalized = True

    # ---- counts (tree_ext.pyx:1649-2019) ------------------------------------------------------------------------------------
    n_edges_z = n_total_edges_z = n_hanging_edges_z = property(lambda self: 0)
    n_faces_x = property(lambda self: self._t().edges[1].n)
    n_faces_y = property(lambda self: self._t().edges[0].n)
    n_faces_z = property(lambda self: 0)
    n_total_faces_x = property(lambda self: self._t().edges[1].n_total)
    n_total_faces_y = property(lambda self: self._t().edges[0].n_total)
    n_total_faces_z = property(lambda self: 0)
    n_hanging_faces_x = property(lambda self: self._t().edges[1].n_hanging)
    n_hanging_faces_y = property(lambda self: self._t().edges[0].n_hanging)
    n_hanging_faces_z = property(lambda self: 0)
    n_faces = property(lambda self: sum(e.n for e in self._t().edges))
    n_total_faces = property(lambda self: sum(e.n_total for e in self._t().edges))
    n_hanging_faces = property(lambda self: sum(e.n_hanging for e in self._t().edges))
    # the reference keeps a trailing zero for the missing z direction in 2-D
    n_faces_per_direction = property(lambda self: (self.n_faces_x, self.n_faces_y, 0))
    n_edges_per_direction = property(lambda self: (self.n_edges_x, self.n_edges_y, 0))
    n_edges = property(lambda self: sum(e.n for e in self._t().edges))
    vntF = property(lambda self: [self.n_total_faces_x, self.n_total_faces_y])
    vntE = property(lambda self: [self.n_total_edges_x, self.n_total_edges_y])

    # ---- geometry ---------------------------------------------------------------------------------------------------------------
    def _coords(self, loc):
        return np.stack([self._xs[d][loc[:, d]] for d in range(2)], axis=1)

    @property
    def h_gridded(self):
        t = self._t()
        return np.stack([self._extent(t.centers, t.half, d) for d in range(2)], axis=1)

    @property
    def cell_volumes(self):
        hg = self.h_gridded
        return hg[:, 0] * hg[:, 1]

    @property
    def cell_centers(self):
        t = self._t()
        return self._midpoints(t.centers, t.half, (0, 1))

    @property
    def edge_lengths(self):
        t = self._t()
        return np.concatenate([self._numbered(t.edges[d], self._edge_lengths_total(d), False) for d in range(2)])

    @property
    def face_areas(self):
        t = self._t()
        return np.concatenate([self._numbered(t.edges[d], self._edge_lengths_total(d), False) for d in (1, 0)])

    faces_x = property(lambda self: self.edges_y)
    faces_y = property(lambda self: self.edges_x)
    hanging_faces_x = property(lambda self: self.hanging_edges_y)
    hanging_faces_y = property(lambda self: self.hanging_edges_x)
    faces = property(lambda self: np.r_[self.faces_x, self.faces_y])
    edges = property(lambda self: np.r_[self.edges_x, self.edges_y])
    nodes_z = cell_centers_z = property(lambda self: None)

    @property
    def edges_z(self):
        return np.zeros((0, 2))

    @property
    def faces_z(self):
        return self.cell_centers

    permute_faces = property(lambda self: self._lexsorted([self.faces_x, self.faces_y]))
    permute_edges = property(lambda self: self._lexsorted([self.edges_x, self.edges_y]))

    def _unit_vectors(self, counts):
        out = np.zeros((int(sum(counts)), 2))
        o = np.r_[0, np.cumsum(counts)]
        for a in range(2):
            out[o[a]:o[a + 1], a] = 1.0
        return out

    @property
    def cell_bounds(self):
        t = self._t()
        cols = []
        for d in range(2):
            cols += [self._xs[d][t.centers[:, d] - t.half], self._xs[d][t.centers[:, d] + t.half]]
        return np.stack(cols, axis=1)

    def is_inside(self, pts, location_type="nodes", **kwargs):
        pts = as_array_n_by_dim(pts, 2)
        inside = np.ones(pts.shape[0], dtype=bool)
        for i, tensor in enumerate((self.nodes_x, self.nodes_y)):
            tol = (tensor.max() - tensor.min()) * 1.0e-10
            inside = inside & (pts[:, i] >= tensor.min() - tol) & (pts[:, i] <= tensor.max() + tol)
        return inside

    # ---- deflation (tree_ext.pyx:3855-4087) ------------------------------------------------------------------------------------------
    def _deflate_edges(self):
        if "Re" not in self._cache:
            t = self._t()
            self._cache["Re"] = sp.block_diag(
                [self._deflate(t.edges[d], t.edge_parents[d], [0.5, 0.5]) for d in range(2)]).tocsr()
        return self._cache["Re"]

    def _deflate_faces(self):
        if "Rf" not in self._cache:
            t = self._t()
            self._cache["Rf"] = sp.block_diag(
                [self._deflate(t.edges[d], t.edge_parents[d], [0.5, 0.5]) for d in (1, 0)]).tocsr()
        return self._cache["Rf"]

    # ---- host matrices ------------------------------------------------------------------------------------------------------------------
    def _host_matrix(self, name):
        key = "H_" + name
        if key not in self._cache:
            self._cache[key] = _canon(self._assemble(name))
        return self._cache[key]

    def _assemble(self, name):
        t = self._t()
        nC = t.n_cells
        cells = np.arange(nC)
        ex, ey = t.edges
        ntE = [ex.n_total, ey.n_total]
        offE = np.r_[0, np.cumsum(ntE)]                 # total edges: [x | y]
        offF = np.r_[0, ey.n_total, ey.n_total + ex.n_total]   # total faces: [x-faces = y-edges | y-faces = x-edges]
        hg = self.h_gridded
        cx = [ex.index[t.cell_edges[0][:, s]] for s in (0, 1)]   # the cell's x-edges at low / high y (total numbering)
        cy = [ey.index[t.cell_edges[1][:, s]] for s in (0, 1)]   # the cell's y-edges at low / high x
        if name == "face_divergence":      # tree_ext.pyx:3173-3197: -+ L/V = -+ 1/h across the face, deflate faces
            rows = np.tile(cells, 4)
            cols = np.r_[cx[0] + offF[1], cx[1] + offF[1], cy[0], cy[1]]
            vals = np.r_[-1.0 / hg[:, 1], 1.0 / hg[:, 1], -1.0 / hg[:, 0], 1.0 / hg[:, 0]]
            return _csr(rows, cols, vals, (nC, offF[2])) @ self._deflate_faces()
        if name in ("face_x_divergence", "face_y_divergence"):      # tree_mesh.py:961-982: one direction's column block
            o = np.r_[0, np.cumsum(self.n_faces_per_direction)]
            d = "xy".index(name[5])
            return self._host_matrix("face_divergence")[:, o[d]:o[d + 1]]
        if name == "edge_curl":            # tree_ext.pyx:3342-3359: per cell d_x e_y - d_y e_x, deflate edges
            rows = np.tile(cells, 4)
            cols = np.r_[cy[0] + offE[1], cx[1], cy[1] + offE[1], cx[0]]
            vals = np.r_[-1.0 / hg[:, 0], -1.0 / hg[:, 1], 1.0 / hg[:, 0], 1.0 / hg[:, 1]]
            return _csr(rows, cols, vals, (nC, offE[2])) @ self._deflate_edges()
        if name in ("nodal_gradient", "average_node_to_edge", "average_node_to_face"):
            unit = np.eye(3, dtype=np.int64)
            order = (1, 0) if name == "average_node_to_face" else (0, 1)      # faces are listed [y-edges | x-edges]
            n_nh = [t.edges[d].n for d in order]
            off = np.r_[0, np.cumsum(n_nh)]
            rows, cols, vals = [], [], []
            for k, d in enumerate(order):
                e = t.edges[d]
                nh = np.nonzero(~e.hanging)[0]
                r = off[k] + e.index[nh]
                length = self._extent(e.loc[nh], e.half[nh], d)
                for s in (-1, 1):
                    q = e.loc[nh] + s * e.half[nh, None] * unit[d]
                    rows.append(r)
                    cols.append(t.nodes.index[np.searchsorted(t.nodes.keys, key3(q[:, 0], q[:, 1], q[:, 2]))])
                    vals.append(s / length if name == "nodal_gradient" else np.full(nh.size, 0.5))
            M = _csr(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), (off[2], t.nodes.n_total))
            return M @ self._deflate_nodes()
        if name == "average_node_to_cell":
            rows = np.repeat(cells, 4)
            cols = t.nodes.index[t.cell_nodes.reshape(-1)]
            return _csr(rows, cols, np.full(rows.size, 0.25), (nC, t.nodes.n_total)) @ self._deflate_nodes()
        if name.startswith("average_edge_") and name.endswith(("_to_cell", "_to_cell_vector")):
            bx = _csr(np.tile(cells, 2), np.r_[cx[0], cx[1]], np.full(2 * nC, 0.5), (nC, offE[2]))
            by = _csr(np.tile(cells, 2), np.r_[cy[0], cy[1]] + offE[1], np.full(2 * nC, 0.5), (nC, offE[2]))
            return self._combine(name, bx, by, self._deflate_edges(), self.n_edges_per_direction)
        if name.startswith("average_face_") and name.endswith(("_to_cell", "_to_cell_vector")):
            bx = _csr(np.tile(cells, 2), np.r_[cy[0], cy[1]], np.full(2 * nC, 0.5), (nC, offF[2]))
            by = _csr(np.tile(cells, 2), np.r_[cx[0], cx[1]] + offF[1], np.full(2 * nC, 0.5), (nC, offF[2]))
            return self._combine(name, bx, by, self._deflate_faces(), self.n_faces_per_direction)
        if name == "average_edge_to_face":      # faces ARE the other direction's edges: a permutation (tree_ext.pyx:4350-4360)
            nx, ny = self.n_edges_x, self.n_edges_y
            return sp.diags([1.0, 1.0], [-ny, nx], shape=(nx + ny, nx + ny)).tocsr()
        if name[:-2] in ("average_node_to_edge", "average_node_to_face") and name[-2:] in ("_x", "_y"):
            d = "xy".index(name[-1])
            per = self.n_edges_per_direction if "edge" in name else self.n_faces_per_direction
            o = np.r_[0, np.cumsum(per)]
            return self._host_matrix(name[:-2])[o[d]:o[d + 1]]
        if name.startswith(("stencil_cell_gradient", "average_cell_to_total_face_", "average_cell_to_face")) \
                or name == "average_cell_vector_to_face":
            return self._host_cell_to_face(name)       # the (minus cell, plus cell) walk of tree_mesh.py, 2-D faces
        if name.startswith("cell_gradient"):           # -Pa Mf^-1 D^T diag(V) (tree_mesh.py:860-958)
            o = np.r_[0, np.cumsum(self.n_faces_per_direction)]
            mask = np.ones(o[2])
            fb = self.face_boundary_indices
            for dd in range(2):
                mask[o[dd] + fb[2 * dd]] = 0.0
                mask[o[dd] + fb[2 * dd + 1]] = 0.0
            mf = 2.0 * (self._host_matrix("average_face_to_cell").T @ self.cell_volumes)
            G = -(sp.diags(mask / mf) @ self._host_matrix("face_divergence").T @ sp.diags(self.cell_volumes))
            if name == "cell_gradient":
                return G
            d = "xy".index(name[-1])
            return G.tocsr()[o[d]:o[d + 1]]
        raise NotImplementedError(f"discretize_b200 QuadTree: {name} is not built")

    @staticmethod
    def _combine(name, bx, by, R, per_direction):
        if name.endswith("_to_cell_vector"):
            return sp.vstack([bx, by]) @ R
        if name.endswith(("_x_to_cell", "_y_to_cell")):
            d = "xy".index(name[-9])
            o = np.r_[0, np.cumsum(per_direction)]
            return ((bx, by)[d] @ R)[:, o[d]:o[d + 1]]
        return 0.5 * (bx + by) @ R

    def _operator(self, name):
        key = "OP_" + name
        if key not in self._cache:
            from .csr import CsrOperator

            self._cache[key] = CsrOperator(self._host_matrix(name))
        return self._cache[key]

    def _cell_gradient_component(self, d):
        return self._operator("cell_gradient" if d is None else "cell_gradient_" + "xyz"[d])

    # ---- mass matrices (base_tensor_mesh.py:792-988 with the tree's averages; inner_products.py:147-272, 459-565) ----------
    def _mass_setup(self, proj):
        base = "average_face_to_cell" if proj == "F" else "average_edge_to_cell"
        return self._host_matrix(base), self._host_matrix(base + "_vector"), (self.n_faces if proj == "F" else self.n_edges)

    def _corner_projections(self, proj):
        """The four (2 nC x n) matrices that pick, per cell, the x- and y-entity meeting at each corner, scaled by
        sqrt(V / 4) (inner_products.py:221-272 with tree_ext.pyx:5651-5763)."""
        t = self._t()
        nC = t.n_cells
        ex, ey = t.edges
        w = np.sqrt(0.25 * self.cell_volumes)
        out = []
        for sy in (0, 1):
            for sx in (0, 1):
                if proj == "F":      # x-face (a y-edge) on the corner's x side, y-face (an x-edge) on its y side
                    a = ey.index[t.cell_edges[1][:, sx]]
                    b = ex.index[t.cell_edges[0][:, sy]] + ey.n_total
                    R = self._deflate_faces()
                else:                # x-edge on the corner's y side, y-edge on its x side
                    a = ex.index[t.cell_edges[0][:, sy]]
                    b = ey.index[t.cell_edges[1][:, sx]] + ex.n_total
                    R = self._deflate_edges()
                P = _csr(np.arange(2 * nC), np.r_[a, b], np.r_[w, w], (2 * nC, R.shape[0]))
                out.append((P @ R).tocsr())
        return out

    def _inner_product(self, proj, model, invert_model, invert_matrix, do_fast, kwargs):
        from .csr import CsrOperator
        from .tensor_mesh import TensorMesh

        TensorMesh._check_removed_kwargs(kwargs)
        nC = self.n_cells
        av, avv, n = self._mass_setup(proj)
        V = self.cell_volumes
        if model is None:
            model = np.ones(nC)
        if is_scalar(model):
            model = np.full(nC, float(np.asarray(model).reshape(-1)[0]))
        model = np.asarray(model, dtype=np.float64)
        if model.size == 3 * nC:          # full tensor [xx, yy, xy]
            m = model.reshape(nC, 3, order="F")
            xx, yy, xy = m[:, 0], m[:, 1], m[:, 2]
            if invert_model:
                det = xx * yy - xy * xy
                xx, yy, xy = yy / det, xx / det, -xy / det
            if invert_matrix:
                raise Exception("Solver needed to invert A.")
            Mu = sp.bmat([[sp.diags(xx), sp.diags(xy)], [sp.diags(xy), sp.diags(yy)]]).tocsr()
            A = None
            for P in self._corner_projections(proj):
                term = P.T @ Mu @ P
                A = term if A is None else A + term
            return CsrOperator(_canon(A))
        if invert_model:
            model = 1.0 / model
        if model.size == nC:
            diag = 2.0 * (av.T @ (V * model.reshape(-1)))
        elif model.size == 2 * nC:
            diag = avv.T @ (np.tile(V, 2) * model.reshape(-1, order="F"))
        else:
            raise Exception("Unexpected shape of tensor")
        if invert_matrix:
            diag = 1.0 / diag
        return CsrOperator(sp.diags(diag, format="csr"))

    def get_face_inner_product_deriv(self, model, do_fast=True, invert_model=False, invert_matrix=False, **kwargs):
        return self._inner_product_deriv("F", model, do_fast, invert_model, invert_matrix, kwargs)

    def get_edge_inner_product_deriv(self, model, do_fast=True, invert_model=False, invert_matrix=False, **kwargs):
        return self._inner_product_deriv("E", model, do_fast, invert_model, invert_matrix, kwargs)

    def _inner_product_deriv(self, proj, model, do_fast, invert_model, invert_matrix, kwargs):
        from .csr import CsrOperator
        from .tensor_mesh import TensorMesh

        TensorMesh._check_removed_kwargs(kwargs)
        if model is None:
            if invert_model or invert_matrix:      # no fast form for "no model": the generic path refuses inversions
                raise NotImplementedError("inverting the property or the matrix is not yet implemented for this "
                                          "mesh/tensorType. You should write it!")
            return None
        nC = self.n_cells
        av, avv, n = self._mass_setup(proj)
        V = self.cell_volumes
        scalar = is_scalar(model)
        m = np.asarray(model, dtype=np.float64)
        if m.size == 3 * nC:
            if invert_model or invert_matrix:
                raise NotImplementedError("inverting the property or the matrix is not yet implemented for this "
                                          "mesh/tensorType. You should write it!")
            Ps = self._corner_projections(proj)
            Z = sp.csr_matrix((nC, nC))

            def tensor_deriv(v=None):
                if v is None:
                    raise Exception("v must be supplied for this implementation.")
                v = np.asarray(v, dtype=np.float64).reshape(-1)
                blocks = [None, None, None]
                for P in Ps:
                    y = P @ v
                    y1, y2 = sp.diags(y[:nC]), sp.diags(y[nC:])
                    terms = (P.T @ sp.vstack((y1, Z)), P.T @ sp.vstack((Z, y2)), P.T @ sp.vstack((y2, y1)))
                    blocks = [t_ if b is None else b + t_ for b, t_ in zip(blocks, terms)]
                return CsrOperator(_canon(sp.hstack(blocks)))

            return tensor_deriv
        mm = np.full(nC, float(m.reshape(-1)[0])) if scalar else m.reshape(-1, order="F")
        if mm.size == nC:
            core = 2.0 * av.T @ sp.diags(V)
        elif mm.size == 2 * nC:
            core = avv.T @ sp.diags(np.tile(V, 2))
        else:
            raise Exception("Unexpected shape of tensor")
        if invert_matrix:
            mi = self._inner_product(proj, model, invert_model, True, True, {}).tocsr().diagonal()
            core = sp.diags((1.0 if invert_model else -1.0) * mi * mi) @ core
        if invert_model:
            if scalar:
                raise TypeError("Vector must be a numpy array")
            core = core @ sp.diags((1.0 if invert_matrix else -1.0) / (mm * mm))
        if scalar and not (invert_matrix and not invert_model):
            core = sp.csr_matrix(core @ np.ones((nC, 1)))
        core = sp.csr_matrix(core)

        def deriv(v=None):
            if v is None:
                raise Exception("v must be supplied for this implementation.")
            return CsrOperator((sp.diags(np.asarray(v, dtype=np.float64).reshape(-1)) @ core).tocsr())

        return deriv

    # ---- point location and interpolation (tree.cpp:758-766, 1499-1516; tree_ext.pyx:5765-6404, dim == 2) -------------------
    def _containing_cells(self, points):
        tr = self._tree
        pts = np.asarray(points, dtype=np.float64)
        n = pts.shape[0]
        c = np.zeros((n, 2), dtype=np.int64)
        for d in range(2):
            edges = self._xs[d][2 * tr.rw * np.arange(1, tr.roots[d])]
            c[:, d] = np.searchsorted(edges, pts[:, d], side="right")
        lo = np.zeros((n, 2), dtype=np.int64)
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
                for d in range(2):
                    mid = self._xs[d][2 * c[go, d] * w + w]
                    c[go, d] = 2 * c[go, d] + (pts[go, d] > mid)
        leaf_lo, _ = tr.leaves()
        return np.searchsorted(tr.keys(leaf_lo), tr.keys(lo))

    def point2index(self, locs):
        out = self._containing_cells(as_array_n_by_dim(np.asarray(locs, dtype=np.float64), 2))
        return out[0] if out.size == 1 else out      # one point gives a scalar (tree_ext.pyx:1350-1352)

    def get_interpolation_matrix(self, loc, location_type="cell_centers", zeros_outside=False, **kwargs):
        if "locType" in kwargs:
            raise TypeError("The locType keyword argument has been removed, please use location_type. "
                            "This will be removed in discretize 1.0.0")
        if "zerosOutside" in kwargs:
            raise TypeError("The zerosOutside keyword argument has been removed, please use zeros_outside. "
                            "This will be removed in discretize 1.0.0")
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        from . import quad_interp
        from .csr import CsrOperator

        location_type = self._parse_location_type(location_type)
        if "z" in location_type:
            raise NotImplementedError("Unable to interpolate from Z edges/faces in 2D")
        pts = as_array_n_by_dim(np.asarray(loc, dtype=np.float64), 2)
        if location_type == "nodes":
            A = quad_interp.node_interpolation(self, pts, zeros_outside)
        elif location_type in ("edges_x", "edges_y"):
            A = quad_interp.edge_interpolation(self, pts, zeros_outside, location_type[-1])
        elif location_type in ("faces_x", "faces_y"):
            A = quad_interp.face_interpolation(self, pts, zeros_outside, location_type[-1])
        elif location_type == "cell_centers":
            A = quad_interp.cell_interpolation(self, pts, zeros_outside)
        else:
            raise ValueError("Location must be a grid location, not {}".format(location_type))
        return CsrOperator(A)

    # ---- everything inherited from the OcTree that has no 2-D form here -----------------------------------------------------
    def _not_built(name):      # noqa: N805 - helper evaluated at class creation
        def raiser(self, *args, **kwargs):
            raise NotImplementedError(f"discretize_b200 QuadTree: {name} is not built")
        return raiser

    for _n in ("stencil_cell_gradient_z", "average_cell_to_face_z", "cell_gradient_z", "average_face_z_to_cell",
               "average_edge_z_to_cell", "face_z_divergence", "average_node_to_edge_z", "average_node_to_face_z",
               "hanging_edges_z", "hanging_faces_z"):
        locals()[_n] = property(_not_built(_n))
    for _n in ("average_cell_to_total_face_z",):
        locals()[_n] = _not_built(_n)
    del _n, _not_built

    def __repr__(self):
        n = self._topo.n_cells if self._finalized else self._tree.n_leaves
        return f"<discretize_b200.TreeMesh (QuadTree) {n} cells, max_level {self.max_level}>"

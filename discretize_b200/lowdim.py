"""1-D and 2-D TensorMesh operators: host assembly, applied on the GPU by the generic CSR kernel.

The hand-written stencil kernels (K1-K9) are 3-D.  The same reference entry points also serve 1-D and 2-D tensor
meshes ("Dimension coverage"); those are small problems, so they are routed to ``dzb_csr_spmv``: every operator
is written down here ONCE as a sum of tensor products of 1-D two-point stencils -- by index arithmetic on the tensor
grid, in any dimension -- and uploaded as CSR.  Layouts and special cases follow the reference:

* x fastest; faces [Fx; Fy], edges [Ex; Ey]; in 2-D an x-edge has the extents of a y-face and vice versa
  (``base/base_regular_mesh.py:601-762``); in 1-D edges have the extents of cells;
* 2-D ``edge_curl`` maps edges to CELLS: ``diag(1/V) [-D_y, D_x] diag(L)`` (``operators/differential_operators.py:
  2134-2140, 2176-2181``); 1-D has no curl;
* diagonal mass matrices: ``diag(dim * Av^T (V o sigma))`` / ``diag(AvV^T (V o sigma_cc))``
  (``base/base_tensor_mesh.py:792-863``);
* interpolation: ``interpolation_utils.py:44-172`` (1-D linear, 2-D bilinear; entry order of ``_interpmat1D/2D``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# ---- 1-D two-point stencils as (row, col, value) triplets ----------------------------------------------------------


def _eye(n):
    i = np.arange(n)
    return i, i, np.ones(n)


def _ddx(n):          # cells <- nodes: (-1, +1)
    i = np.arange(n)
    return np.r_[i, i], np.r_[i, i + 1], np.r_[-np.ones(n), np.ones(n)]


def _av(n):           # cells <- nodes: (1/2, 1/2)
    i = np.arange(n)
    return np.r_[i, i], np.r_[i, i + 1], np.full(2 * n, 0.5)


def _av_extrap(n):    # nodes <- cells: (1/2, 1/2), weight 1 on the two boundary nodes
    i = np.arange(1, n)
    return np.r_[0, i, i, n], np.r_[0, i - 1, i, n - 1], np.r_[1.0, np.full(2 * (n - 1), 0.5), 1.0]


class _Grid:
    """Extents and offsets of the dof blocks of a tensor mesh of any dimension."""

    def __init__(self, mesh):
        self.mesh, self.dim, self.n = mesh, mesh.dim, mesh.shape_cells
        d = self.dim
        self.cell = (False,) * d
        self.node = (True,) * d
        self.face = [tuple(a == b for b in range(d)) for a in range(d)]           # node extent along the normal only
        self.edge = [tuple(a != b for b in range(d)) for a in range(d)]           # node extent across the edge

    def shape(self, loc):
        return tuple(n + int(f) for n, f in zip(self.n, loc))

    def count(self, loc):
        return int(np.prod(self.shape(loc)))

    def offsets(self, locs):
        return np.r_[0, np.cumsum([self.count(l) for l in locs])]

    def term(self, row_loc, col_loc, stencils, scale=1.0, row_shape=None):
        """COO triplets (local to the two blocks) of the tensor product of per-axis stencils, x fastest."""
        rs, cs = (row_shape or self.shape(row_loc)), self.shape(col_loc)
        rows = np.zeros(1, dtype=np.int64)
        cols = np.zeros(1, dtype=np.int64)
        vals = np.full(1, float(scale))
        rstride = cstride = 1
        for a in range(self.dim):
            r, c, v = stencils[a]
            rows = (rows[None, :] + rstride * r[:, None]).reshape(-1)
            cols = (cols[None, :] + cstride * c[:, None]).reshape(-1)
            vals = (vals[None, :] * v[:, None]).reshape(-1)
            rstride *= rs[a]
            cstride *= cs[a]
        return rows, cols, vals


def _assemble(shape, blocks):
    rows = np.concatenate([b[0] for b in blocks])
    cols = np.concatenate([b[1] for b in blocks])
    vals = np.concatenate([b[2] for b in blocks])
    m = sp.csr_matrix((vals, (rows, cols)), shape=shape)
    m.sum_duplicates()
    m.sort_indices()
    return m


def _axis_stencils(g, special, stencil_of):
    """identity on every axis (extent of the row location) except those in ``special`` (axis -> stencil)."""
    return [special[a] if a in special else stencil_of(a) for a in range(g.dim)]


# ---- geometry ----------------------------------------------------------------------------------------------------


def _outer(vs):
    out = np.asarray(vs[0], dtype=np.float64)
    for v in vs[1:]:
        out = np.outer(out, v).reshape(-1, order="F")
    return out


def cell_volumes(mesh):
    return _outer(list(mesh.h))


def face_areas(mesh):
    """1-D: ones; 2-D: lengths of the faces; per direction, stacked (tensor_mesh.py:312-403)."""
    d, h = mesh.dim, mesh.h
    out = []
    for a in range(d):
        out.append(_outer([np.ones(h[b].size + 1) if b == a else h[b] for b in range(d)]))
    return np.concatenate(out)


def edge_lengths(mesh):
    d, h = mesh.dim, mesh.h
    out = []
    for a in range(d):
        out.append(_outer([h[b] if b == a else np.ones(h[b].size + 1) for b in range(d)]))
    return np.concatenate(out)


# ---- operators ----------------------------------------------------------------------------------------------------


def build(mesh, name):
    """scipy CSR of ``mesh.<name>`` for a 1-D or 2-D TensorMesh (canonical form)."""
    g = _Grid(mesh)
    d, n, h = g.dim, g.n, mesh.h
    rh = [1.0 / x for x in h]
    nC = g.count(g.cell)
    offF, offE = g.offsets(g.face), g.offsets(g.edge)

    def ident(loc):
        return lambda a: _eye(g.shape(loc)[a])

    def scaled(st, w):           # scale the rows of a (cells <- nodes) stencil by a per-cell weight
        r, c, v = st
        return r, c, v * w[r]

    if name in ("face_divergence", "face_x_divergence", "face_y_divergence"):
        blocks = []
        axes = range(d) if name == "face_divergence" else ["xyz".index(name[5])]
        for a in axes:
            st = _axis_stencils(g, {a: scaled(_ddx(n[a]), rh[a])}, ident(g.cell))
            r, c, v = g.term(g.cell, g.face[a], st)
            blocks.append((r, c + (offF[a] if name == "face_divergence" else 0), v))
        ncols = offF[-1] if name == "face_divergence" else g.count(g.face[axes[0]])
        return _assemble((nC, ncols), blocks)
    if name == "nodal_gradient":
        blocks = []
        for a in range(d):
            st = _axis_stencils(g, {a: scaled(_ddx(n[a]), rh[a])}, ident(g.edge[a]))
            r, c, v = g.term(g.edge[a], g.node, st)
            blocks.append((r + offE[a], c, v))
        return _assemble((offE[-1], g.count(g.node)), blocks)
    if name == "edge_curl":
        if d != 2:
            raise NotImplementedError("Edge Curl only programed for 2 or 3D.")
        # (curl e)_cell = d_x e_y - d_y e_x = (L_y e_y)|_{i+1} - (L_y e_y)|_i - ... all over V  ->  +-1/h of the OTHER axis
        sx = [_eye(n[0]), scaled(_ddx(n[1]), rh[1])]            # -d_y on x-edges (nx, ny+1)
        r0, c0, v0 = g.term(g.cell, g.edge[0], sx, -1.0)
        sy = [scaled(_ddx(n[0]), rh[0]), _eye(n[1])]            # +d_x on y-edges (nx+1, ny)
        r1, c1, v1 = g.term(g.cell, g.edge[1], sy)
        return _assemble((nC, offE[-1]), [(r0, c0, v0), (r1, c1 + offE[1], v1)])
    ave = {
        "average_face_to_cell": (g.face, offF, "scalar"), "average_face_to_cell_vector": (g.face, offF, "vector"),
        "average_edge_to_cell": (g.edge, offE, "scalar"), "average_edge_to_cell_vector": (g.edge, offE, "vector"),
    }
    comp = {"average_face_x_to_cell": (g.face, 0), "average_face_y_to_cell": (g.face, 1),
            "average_edge_x_to_cell": (g.edge, 0), "average_edge_y_to_cell": (g.edge, 1)}

    def to_cell(loc):            # average along every node-extent axis of the source location
        return [_av(n[a]) if loc[a] else _eye(n[a]) for a in range(d)]

    if name in ave:
        locs, off, kind = ave[name]
        blocks = []
        for a in range(d):
            r, c, v = g.term(g.cell, locs[a], to_cell(locs[a]), (1.0 / d) if kind == "scalar" else 1.0)
            blocks.append((r + (a * nC if kind == "vector" else 0), c + off[a], v))
        return _assemble(((d if kind == "vector" else 1) * nC, off[-1]), blocks)
    if name in comp and comp[name][1] < d:
        locs, a = comp[name]
        return _assemble((nC, g.count(locs[a])), [g.term(g.cell, locs[a], to_cell(locs[a]))])
    if name == "average_node_to_cell":
        return _assemble((nC, g.count(g.node)), [g.term(g.cell, g.node, to_cell(g.node))])
    if name in ("average_cell_to_face", "average_cell_vector_to_face"):
        blocks = []
        vec = name == "average_cell_vector_to_face" and d > 1
        for a in range(d):
            st = _axis_stencils(g, {a: _av_extrap(n[a])}, ident(g.cell))
            r, c, v = g.term(g.face[a], g.cell, st)
            blocks.append((r + offF[a], c + (a * nC if vec else 0), v))
        return _assemble((offF[-1], (d if vec else 1) * nC), blocks)
    if name == "average_node_to_edge":
        blocks = []
        for a in range(d):
            st = _axis_stencils(g, {a: _av(n[a])}, ident(g.edge[a]))
            r, c, v = g.term(g.edge[a], g.node, st)
            blocks.append((r + offE[a], c, v))
        return _assemble((offE[-1], g.count(g.node)), blocks)
    if name == "average_node_to_face":
        blocks = []
        for a in range(d):
            st = [_eye(n[b] + 1) if b == a else _av(n[b]) for b in range(d)]
            r, c, v = g.term(g.face[a], g.node, st)
            blocks.append((r + offF[a], c, v))
        return _assemble((offF[-1], g.count(g.node)), blocks)
    if name == "average_cell_to_edge":       # av_extrap across the edge (differential_operators.py:2867-2892); 1-D: identity
        if d == 1:
            return sp.identity(n[0], format="csr")
        blocks = []
        for a in range(d):
            st = [_eye(n[b]) if b == a else _av_extrap(n[b]) for b in range(d)]
            r, c, v = g.term(g.edge[a], g.cell, st)
            blocks.append((r + offE[a], c, v))
        return _assemble((offE[-1], nC), blocks)
    if name == "average_edge_to_face":       # 1-D: average_cell_to_face; 2-D: x-faces ARE y-edges and vice versa (:3205-3217)
        if d == 1:
            return build(mesh, "average_cell_to_face")
        if d == 3:           # a face <- its four edges, 1/4 each: the two tangent directions, averaged across the face
            blocks = []
            for a in range(d):
                for b in range(d):
                    if b != a:
                        st = [_eye(n[c] + 1) if c == a else _eye(n[c]) if c == b else _av(n[c]) for c in range(d)]
                        r, c, v = g.term(g.face[a], g.edge[b], st, 0.5)
                        blocks.append((r + offF[a], c + offE[b], v))
            return _assemble((offF[-1], offE[-1]), blocks)
        nFx, nFy = g.count(g.face[0]), g.count(g.face[1])
        return sp.diags([1.0, 1.0], [-nFx, nFy], shape=(nFx + nFy, offE[-1])).tocsr()
    raise NotImplementedError(f"discretize_b200.TensorMesh.{name} is not available for mesh.dim == {d}")


def _end_rows(st, n_rows):
    """Keep the first and last row of a 1-D stencil, renumbered 0 and 1."""
    r, c, v = st
    keep = (r == 0) | (r == n_rows - 1)
    return (r[keep] != 0).astype(np.int64), c[keep], v[keep]


def boundary_face_rows(mesh, name):
    """``project_face_to_boundary_face @ mesh.<name>`` for the three averages the boundary integrals need, any dimension,
    assembled from the boundary rows only (O(boundary) work): the 1-D stencil along the face normal is cut down to its
    two end rows, which keeps the boundary faces in their global order."""
    g = _Grid(mesh)
    d, n = g.dim, g.n
    blocks, row0 = [], 0
    if name == "average_edge_to_face" and d == 2:       # a boundary x-face IS a y-edge and vice versa
        offE = g.offsets(g.edge)
        for a in range(d):
            st = [_end_rows(_eye(n[b] + 1), n[b] + 1) if b == a else _eye(n[b]) for b in range(d)]
            shape = tuple(2 if b == a else n[b] for b in range(d))
            r, c, v = g.term(g.face[a], g.edge[1 - a], st, row_shape=shape)
            blocks.append((r + row0, c + offE[1 - a], v))
            row0 += int(np.prod(shape))
        return _assemble((row0, offE[-1]), blocks)
    for a in range(d):
        shape = tuple(2 if b == a else n[b] for b in range(d))
        if name == "average_cell_to_face":
            st = [_end_rows(_av_extrap(n[b]), n[b] + 1) if b == a else _eye(n[b]) for b in range(d)]
            r, c, v = g.term(g.face[a], g.cell, st, row_shape=shape)
            blocks.append((r + row0, c, v))
        elif name == "average_node_to_face":
            st = [_end_rows(_eye(n[b] + 1), n[b] + 1) if b == a else _av(n[b]) for b in range(d)]
            r, c, v = g.term(g.face[a], g.node, st, row_shape=shape)
            blocks.append((r + row0, c, v))
        elif name == "average_edge_to_face":
            if d == 1:
                return boundary_face_rows(mesh, "average_cell_to_face")
            offE = g.offsets(g.edge)
            for b in range(d):       # the two edge directions tangent to the face, two edges of each
                if b == a:
                    continue
                st = [_end_rows(_eye(n[c] + 1), n[c] + 1) if c == a else _eye(n[c]) if c == b else _av(n[c]) for c in range(d)]
                r, c, v = g.term(g.face[a], g.edge[b], st, 0.5, row_shape=shape)
                blocks.append((r + row0, c + offE[b], v))
        else:
            raise KeyError(name)
        row0 += int(np.prod(shape))
    ncols = {"average_cell_to_face": g.count(g.cell), "average_node_to_face": g.count(g.node),
             "average_edge_to_face": int(g.offsets(g.edge)[-1])}[name]
    return _assemble((row0, ncols), blocks)


def _cell_grad_1d(n, bc):
    """nodes <- cells: (-1, +1); ghost-point boundary rows: Dirichlet +-2, Neumann an explicit 0
    (operators/differential_operators.py:41-82)."""
    i = np.arange(1, n)
    first = {"dirichlet": 2.0, "neumann": 0.0}[bc[0]]
    last = {"dirichlet": -2.0, "neumann": 0.0}[bc[1]]
    return np.r_[0, i, i, n], np.r_[0, i - 1, i, n - 1], np.r_[first, np.full(n - 1, -1.0), np.full(n - 1, 1.0), last]


def cell_gradient(mesh, bcs, scaled, axis=None):
    """stencil_cell_gradient (``scaled=False``) / cell_gradient = diag(S / aveCC2F V) stencil (:1413-1595); ``axis``: one
    component block with Neumann conditions (:1049-1411, 1762-2083).  The product drops the Neumann rows' explicit zeros."""
    g = _Grid(mesh)
    d, n = g.dim, g.n
    offF = g.offsets(g.face)
    axes = range(d) if axis is None else [axis]
    blocks = []
    for a in axes:
        bc = bcs[a] if axis is None else ["neumann", "neumann"]
        st = [_cell_grad_1d(n[b], bc) if b == a else _eye(n[b]) for b in range(d)]
        r, c, v = g.term(g.face[a], g.cell, st)
        blocks.append((r + (offF[a] if axis is None else 0), c, v))
    nrows = offF[-1] if axis is None else g.count(g.face[axis])
    G = _assemble((nrows, g.count(g.cell)), blocks)
    if not scaled:
        return G
    scale = face_areas(mesh) / (build(mesh, "average_cell_to_face") @ cell_volumes(mesh))
    if axis is not None:
        scale = scale[offF[axis]:offF[axis + 1]]
    G = (sp.diags(scale) @ G).tocsr()
    G.eliminate_zeros()
    G.sort_indices()
    return G


# ---- diagonal mass matrices ----------------------------------------------------------------------------------------


def mass_diagonal(mesh, proj, model, invert_model=False, invert_matrix=False):
    """Diagonal of get_face_inner_product / get_edge_inner_product for scalar, isotropic and anisotropic models
    (base/base_tensor_mesh.py:792-863): dim * Av^T (V sigma) or AvV^T (V sigma_cc)."""
    d, nC = mesh.dim, mesh.n_cells
    V = cell_volumes(mesh)
    if model is None:
        model = np.ones(nC)
    model = np.asarray(model, dtype=np.float64)
    if model.size == 1:
        model = np.full(nC, float(model.reshape(-1)[0]))
    if invert_model:
        model = 1.0 / model
    base = "average_face_to_cell" if proj == "F" else "average_edge_to_cell"
    if model.size == nC:
        diag = d * (build(mesh, base).T @ (V * model.reshape(-1)))
    elif model.size == nC * d:
        diag = build(mesh, base + "_vector").T @ (np.tile(V, d) * model.reshape(-1, order="F"))
    else:
        raise NotImplementedError("full-tensor models on 1-D / 2-D meshes are not built")
    return 1.0 / diag if invert_matrix else diag


# ---- full-tensor models in 2-D (operators/inner_products.py:147-272, 459-565; 2-D tables :599-669, 750-790) ---------


def _corner_entities_2d(mesh, proj):
    """For each of the four cell corners (x low/high, y low/high): the global numbers of the x- and the y-entity
    (face or edge) of every cell that touch that corner, in the reference's corner order 000, 100, 010, 110."""
    nx, ny = mesh.shape_cells
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    i, j = i.reshape(-1, order="F"), j.reshape(-1, order="F")
    out = []
    for sy in (0, 1):
        for sx in (0, 1):
            if proj == "F":      # x-face on the corner's x side, y-face on its y side
                a = (i + sx) + (nx + 1) * j
                b = (nx + 1) * ny + i + nx * (j + sy)
            else:                # x-edge on the corner's y side, y-edge on its x side
                a = i + nx * (j + sy)
                b = nx * (ny + 1) + (i + sx) + (nx + 1) * j
            out.append((a, b))
    return out


def _tensor_model_2d(mesh, model, invert_model):
    m = np.asarray(model, dtype=np.float64).reshape(mesh.n_cells, 3, order="F")
    xx, yy, xy = m[:, 0], m[:, 1], m[:, 2]
    if invert_model:             # closed-form inverse of [[xx, xy], [xy, yy]] per cell (utils/matrix_utils.py:1246-1434)
        det = xx * yy - xy * xy
        xx, yy, xy = yy / det, xx / det, -xy / det
    return xx, yy, xy


def mass_tensor_matrix(mesh, proj, model, invert_model=False):
    """sum over corners of P^T sqrt(V/4) Sigma sqrt(V/4) P for a 2-D model given as (nC, 3) = [xx, yy, xy]."""
    if mesh.dim != 2:
        raise NotImplementedError("full-tensor models exist for 2-D and 3-D meshes only")
    xx, yy, xy = _tensor_model_2d(mesh, model, invert_model)
    w = 0.25 * cell_volumes(mesh)
    n = mesh.n_faces if proj == "F" else mesh.n_edges
    rows, cols, vals = [], [], []
    for a, b in _corner_entities_2d(mesh, proj):
        rows += [a, a, b, b]
        cols += [a, b, a, b]
        vals += [w * xx, w * xy, w * xy, w * yy]
    return _assemble((n, n), [(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))])


def mass_tensor_deriv_matrix(mesh, proj, v):
    """d(M(m) v)/dm for the 2-D full-tensor model: n x 3 nC, column blocks [xx | yy | xy]."""
    if mesh.dim != 2:
        raise NotImplementedError("full-tensor models exist for 2-D and 3-D meshes only")
    nC = mesh.n_cells
    n = mesh.n_faces if proj == "F" else mesh.n_edges
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    w = 0.25 * cell_volumes(mesh)
    c = np.arange(nC)
    rows, cols, vals = [], [], []
    for a, b in _corner_entities_2d(mesh, proj):
        rows += [a, b, a, b]
        cols += [c, c + nC, c + 2 * nC, c + 2 * nC]
        vals += [w * v[a], w * v[b], w * v[b], w * v[a]]
    return _assemble((n, 3 * nC), [(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))])


# ---- interpolation -------------------------------------------------------------------------------------------------


def _bisect_weights(x, xp):
    """interputils_cython.pyx:42-67 (_bisect_right + _get_inds_ws), vectorised."""
    ind = np.searchsorted(x, xp, side="right")
    i2 = np.clip(ind, 0, x.size - 1)
    i1 = np.clip(ind - 1, 0, x.size - 1)
    same = i1 == i2
    den = np.where(same, 1.0, x[i2] - x[i1])
    w1 = np.where(same, 0.5, (x[i2] - xp) / den)
    return i1, i2, w1, 1.0 - w1


def interpolation_matrix(mesh, pts, location_type, zeros_outside=False):
    """get_interpolation_matrix for 1-D / 2-D tensor meshes (base_tensor_mesh.py:684-790 ->
    interpolation_utils.py:44-172 -> interputils_cython.pyx:70-126), same entry order, same is_inside rule."""
    from .interp import _location_layout
    from .tensor_mesh import as_array_n_by_dim

    d = mesh.dim
    pts = as_array_n_by_dim(np.asarray(pts, dtype=np.float64), d).copy()
    key, off, ncols = _location_layout(mesh, location_type)
    inside = mesh.is_inside(pts)
    if not zeros_outside:
        if not inside.all():
            raise ValueError("Points outside of mesh")
    else:
        center = np.array([v.mean() for v in mesh.get_tensor("CC")])
        pts[~inside] = center
    axes = mesh.get_tensor(key)
    npts = pts.shape[0]
    hits = [_bisect_weights(axes[a], pts[:, a]) for a in range(d)]
    rows, cols, vals = [], [], []
    if d == 1:
        i1, i2, w1, w2 = hits[0]
        for i, w in ((i1, w1), (i2, w2)):
            rows.append(np.arange(npts)); cols.append(off + i); vals.append(w)
    else:
        nx = axes[0].size
        (x1, x2, wx1, wx2), (y1, y2, wy1, wy2) = hits
        # entry order of _interpmat2D: (x1,y1), (x1,y2), (x2,y1), (x2,y2)
        for ix, iy, w in ((x1, y1, wx1 * wy1), (x1, y2, wx1 * wy2), (x2, y1, wx2 * wy1), (x2, y2, wx2 * wy2)):
            rows.append(np.arange(npts)); cols.append(off + ix + nx * iy); vals.append(w)
    Q = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(npts, ncols))
    Q.sum_duplicates()
    Q.sort_indices()
    if zeros_outside and (~inside).any():
        # the reference's ``Q[outside, :] = 0`` on a CSR matrix stores a whole row of explicit zeros (:771-772)
        counts = np.diff(Q.indptr).astype(np.int64)
        counts[~inside] = ncols
        indptr = np.r_[0, np.cumsum(counts)]
        indices = np.empty(indptr[-1], dtype=np.int64)
        data = np.zeros(indptr[-1])
        for r in range(npts):
            lo, hi = indptr[r], indptr[r + 1]
            if inside[r]:
                indices[lo:hi] = Q.indices[Q.indptr[r]:Q.indptr[r + 1]]
                data[lo:hi] = Q.data[Q.indptr[r]:Q.indptr[r + 1]]
            else:
                indices[lo:hi] = np.arange(ncols)
        Q = sp.csr_matrix((data, indices, indptr), shape=(npts, ncols))
    return Q


# ---- model derivatives of the diagonal mass matrices -----------------------------------------------------------------


def mass_deriv_matrix(mesh, proj, model, v, invert_model=False, invert_matrix=False):
    """d(M(m) v)/dm for scalar / isotropic / anisotropic models on 1-D / 2-D meshes (base/base_tensor_mesh.py:865-988):
    ``diag(v) @ dMdprop`` with ``dMdprop = r (k Av^T) diag(V) c`` -- r = +-MI^2 rows for ``invert_matrix``, c = +-1/m^2
    columns for ``invert_model``; a scalar model gives an n x 1 matrix (n x nC in the reference's invert_matrix-only
    branch, which skips the sum over cells)."""
    d, nC = mesh.dim, mesh.n_cells
    V = cell_volumes(mesh)
    model = np.asarray(model, dtype=np.float64)
    scalar = model.size == 1
    m = np.full(nC, float(model.reshape(-1)[0])) if scalar else model.reshape(-1, order="F")
    base = "average_face_to_cell" if proj == "F" else "average_edge_to_cell"
    if m.size == nC:
        core = d * build(mesh, base).T @ sp.diags(V)
    elif m.size == nC * d:
        core = build(mesh, base + "_vector").T @ sp.diags(np.tile(V, d))
    else:
        raise NotImplementedError("full-tensor models on 1-D / 2-D meshes are not built")
    if invert_matrix:
        mi = mass_diagonal(mesh, proj, model, invert_model, True)
        core = sp.diags((1.0 if invert_model else -1.0) * mi * mi) @ core
    if invert_model:
        if scalar:
            raise TypeError("Vector must be a numpy array")     # the reference fails here, too (sdiag of a bare float)
        core = core @ sp.diags((1.0 if invert_matrix else -1.0) / (m * m))
    if scalar and not (invert_matrix and not invert_model):
        core = sp.csr_matrix(core @ np.ones((nC, 1)))
    return (sp.diags(np.asarray(v, dtype=np.float64).reshape(-1)) @ core).tocsr()

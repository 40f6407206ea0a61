"""Utilities that sit next to the operator path (reference: ``discretize.utils``).

``volume_average`` -- conservative model transfer between meshes: every output cell takes the
overlap-volume weighted mean of the input cells it intersects (``utils/interpolation_utils.py:175-287``).

* TensorMesh -> TensorMesh (``_extensions/interputils_cython.pyx:186-332``): the operator is separable, so it is assembled
  from three 1-D overlap lists with array passes (no per-entry loop);
* any pairing with a TreeMesh (``_extensions/tree_ext.pyx:6573-7053``): output boxes are clamped to the input domain, the
  input cells overlapping each box are collected for ALL boxes at once -- index ranges on a tensor input, a level-by-level
  descent of the linear octree on a tree input -- and weighted by the product of their 1-D overlaps, normalised per row.
  Meshes on the same underlying tensor grid take the reference's shortcut (volume ratios / containing cells).

Host assembly, applied on the GPU by the generic CSR kernel.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _axis_overlaps(x_in, x_out):
    """1-D overlap segments of two node arrays: (length, input cell, output cell) of every segment whose centre lies
    inside the OUTPUT mesh.  Segments outside the input mesh take the nearest input cell (the reference clamps its
    running index the same way: interputils_cython.pyx:290-332)."""
    xs = np.unique(np.r_[x_in, x_out])
    centers = 0.5 * (xs[:-1] + xs[1:])
    keep = (x_out[0] <= centers) & (centers <= x_out[-1])
    c = centers[keep]
    h = np.diff(xs)[keep]

    def cell_of(x, c):
        i = np.minimum(np.searchsorted(x, c, side="right"), x.size - 1)     # first node beyond the centre, capped
        return np.clip(i - 1, 0, x.size - 1)

    return h, cell_of(x_in, c), cell_of(x_out, c)


def _tensor_volume_average_matrix(mesh_in, mesh_out):
    dim = mesh_in.dim
    nodes_in = [mesh_in._nodes_1d(a) for a in range(dim)]
    nodes_out = [mesh_out._nodes_1d(a) for a in range(dim)]
    w, i_in, i_out = zip(*[_axis_overlaps(a, b) for a, b in zip(nodes_in, nodes_out)])
    n_in, n_out = mesh_in.shape_cells, mesh_out.shape_cells
    # tensor product of the per-axis lists, x fastest (the reference's loop nest and flat indices)
    W, Ii, Io = w[0], i_in[0], i_out[0]
    sin, sout = n_in[0], n_out[0]
    for a in range(1, dim):
        W = (w[a][:, None] * W[None, :]).reshape(-1)
        Ii = (i_in[a][:, None] * sin + Ii[None, :]).reshape(-1)
        Io = (i_out[a][:, None] * sout + Io[None, :]).reshape(-1)
        sin *= n_in[a]
        sout *= n_out[a]
    W = W / mesh_out.cell_volumes[Io]
    return sp.csr_matrix((W, (Io, Ii)), shape=(mesh_out.n_cells, mesh_in.n_cells))


# ---- pairings that involve a TreeMesh --------------------------------------------------------------------------------


def _base_nodes(mesh):
    if mesh._meshType == "TREE":
        return [mesh.nodes_x, mesh.nodes_y, mesh.nodes_z][:mesh.dim]
    return [mesh._nodes_1d(a) for a in range(mesh.dim)]


def _same_base(mesh_a, mesh_b):
    try:
        return all(np.allclose(a, b) for a, b in zip(_base_nodes(mesh_a), _base_nodes(mesh_b)))
    except ValueError:
        return False


def _expand_ranges(lo, hi):
    """All index triples of the boxes [lo, hi) (one box per row), x fastest within a box: (box id, (n, 3) indices)."""
    ext = np.maximum(hi - lo, 0)
    cnt = np.prod(ext, axis=1)
    box = np.repeat(np.arange(lo.shape[0]), cnt)
    local = np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    cols = []
    for d in range(lo.shape[1]):
        e = ext[box, d]
        cols.append(local % e)
        local = local // e
    return box, np.stack(cols, axis=1) + lo[box]


def _touching(lo, hi, a, b, active):
    """Box [lo, hi] against cell [a, b], per axis: a genuine overlap where the axis is active, contact where the box
    was collapsed onto the domain boundary (the reference tests closed intervals everywhere and then drops the
    zero-volume pairs; this keeps the same pairs)."""
    return np.where(active, (hi > a) & (lo < b), (hi >= a) & (lo <= b)).all(axis=1)


def _tree_candidates(mesh_in, lo, hi, active):
    """(box id, input cell) of every leaf of ``mesh_in`` that overlaps each box: descend the octree with all boxes
    at once, one level per pass (tree.h:170-182 does the same per box, recursively)."""
    tr = mesh_in._tree
    xs = mesh_in._xs
    dim = mesh_in.dim

    def bounds(c, w):
        a = np.stack([xs[d][2 * c[:, d] * w] for d in range(dim)], axis=1)
        b = np.stack([xs[d][2 * (c[:, d] + 1) * w] for d in range(dim)], axis=1)
        return a, b

    w = tr.level_width(tr.root_level)
    g = tr.grid(tr.root_level)
    # roots: index ranges per axis (closed contact), then the exact test
    r_lo = np.empty(lo.shape, dtype=np.int64)
    r_hi = np.empty(lo.shape, dtype=np.int64)
    for d in range(dim):
        nodes = xs[d][::2 * w]
        r_lo[:, d] = np.clip(np.searchsorted(nodes, lo[:, d], side="left") - 1, 0, g[d])
        r_hi[:, d] = np.clip(np.searchsorted(nodes, hi[:, d], side="right"), 0, g[d])
    box, c = _expand_ranges(r_lo, r_hi)
    from .octree import _child_offsets

    offs = _child_offsets(dim)
    out_box, out_lo = [], []
    for l in range(tr.root_level, tr.max_level + 1):
        if not box.size:
            break
        w = tr.level_width(l)
        a, b = bounds(c, w)
        keep = _touching(lo[box], hi[box], a, b, active[box])
        box, c = box[keep], c[keep]
        if l < tr.max_level and tr.div[l].size:
            isdiv = np.isin(tr.pack(l, c), tr.div[l])
        else:
            isdiv = np.zeros(box.size, dtype=bool)
        out_box.append(box[~isdiv])
        out_lo.append(c[~isdiv] * w)
        box = np.repeat(box[isdiv], offs.shape[0])
        c = (2 * c[isdiv][:, None, :] + offs[None, :, :]).reshape(-1, dim)
    box = np.concatenate(out_box)
    leaf_lo = np.concatenate(out_lo)
    all_lo, _ = tr.leaves()
    return box, np.searchsorted(tr.keys(all_lo), tr.keys(leaf_lo))


def _tensor_candidates(mesh_in, lo, hi, active):
    nodes = _base_nodes(mesh_in)
    n = mesh_in.shape_cells
    i_lo = np.empty(lo.shape, dtype=np.int64)
    i_hi = np.empty(lo.shape, dtype=np.int64)
    dim = mesh_in.dim
    for d in range(dim):       # tree_ext.pyx:6969-6979
        i_lo[:, d] = np.maximum(np.searchsorted(nodes[d], lo[:, d], side="left") - 1, 0)
        i_hi[:, d] = np.minimum(np.searchsorted(nodes[d], hi[:, d], side="right"), n[d])
    box, idx = _expand_ranges(i_lo, i_hi)
    a = np.stack([nodes[d][idx[:, d]] for d in range(dim)], axis=1)
    b = np.stack([nodes[d][idx[:, d] + 1] for d in range(dim)], axis=1)
    keep = _touching(lo[box], hi[box], a, b, active[box])
    idx = idx[keep]
    flat, stride = idx[:, 0].copy(), n[0]
    for d in range(1, dim):
        flat += stride * idx[:, d]
        stride *= n[d]
    return box[keep], flat


def _overlap_volume_average_matrix(mesh_in, mesh_out):
    if mesh_in.dim not in (2, 3):
        raise NotImplementedError("volume_average with a TreeMesh needs 2-D or 3-D meshes")
    base_in = _base_nodes(mesh_in)
    origin = np.array([x[0] for x in base_in])
    x_end = np.array([x[-1] for x in base_in])
    bounds_out = np.asarray(mesh_out.cell_bounds)
    lo = np.minimum(bounds_out[:, 0::2], x_end)           # a box beyond the input mesh collapses onto its boundary
    hi = np.maximum(bounds_out[:, 1::2], origin)
    active = (lo < x_end) & (hi > origin)
    if mesh_in._meshType == "TREE":
        box, cell = _tree_candidates(mesh_in, lo, hi, active)
    else:
        box, cell = _tensor_candidates(mesh_in, lo, hi, active)
    bounds_in = np.asarray(mesh_in.cell_bounds)
    a, b = bounds_in[cell, 0::2], bounds_in[cell, 1::2]
    length = np.where(active[box], np.minimum(hi[box], b) - np.maximum(lo[box], a), 1.0)
    weight = np.prod(length, axis=1) if length.shape[1] == 2 else length[:, 0] * length[:, 1] * length[:, 2]
    nz = weight != 0.0
    box, cell, weight = box[nz], cell[nz], weight[nz]
    total = np.bincount(box, weights=weight, minlength=mesh_out.n_cells)
    return sp.csr_matrix((weight / total[box], (box, cell)), shape=(mesh_out.n_cells, mesh_in.n_cells))


def _cells_of(mesh, points):
    """Containing cell of each point, always as an array (TreeMesh.point2index gives a scalar for one point)."""
    return np.atleast_1d(mesh.point2index(points))


def _same_base_volume_average_matrix(mesh_in, mesh_out):
    """Meshes refined from one tensor grid: nested cells, so weights are volume ratios or 1 (tree_ext.pyx:6627-6680,
    6762-6775, 6900-6912)."""
    n_in, n_out = mesh_in.n_cells, mesh_out.n_cells
    if mesh_in._meshType == "TENSOR":                      # fine tensor cells -> the tree cell that contains them
        rows = _cells_of(mesh_out, mesh_in.cell_centers)
        w = mesh_in.cell_volumes / mesh_out.cell_volumes[rows]
        return sp.csr_matrix((w, (rows, np.arange(n_in))), shape=(n_out, n_in))
    if mesh_out._meshType == "TENSOR":                     # every fine tensor cell sits inside one tree cell
        cols = _cells_of(mesh_in, mesh_out.cell_centers)
        return sp.csr_matrix((np.ones(n_out), (np.arange(n_out), cols)), shape=(n_out, n_in))
    # tree -> tree: input cells finer than their output cell contribute their volume share; output cells no input
    # cell is finer than copy the input cell that contains them
    rows = _cells_of(mesh_out, mesh_in.cell_centers)
    lvl_in, lvl_out = mesh_in._levels, mesh_out._levels
    finer = lvl_out[rows] < lvl_in
    w = np.where(finer, mesh_in.cell_volumes / mesh_out.cell_volumes[rows], 0.0)
    P = sp.csr_matrix((w, (rows, np.arange(n_in))), shape=(n_out, n_in))
    visited = np.zeros(n_out, dtype=bool)
    visited[rows[finer]] = True
    rest = np.nonzero(~visited)[0]
    if rest.size:
        cols = _cells_of(mesh_in, mesh_out.cell_centers[rest])
        P = P + sp.csr_matrix((np.ones(rest.size), (rest, cols)), shape=(n_out, n_in))
    return sp.csr_matrix(P)


def volume_average(mesh_in, mesh_out, values=None, output=None):
    """Volume-averaging interpolation between two meshes (reference: utils/interpolation_utils.py:175-287).

    ``values is None``: returns the operator (n_cells_out x n_cells_in) as a CUDA ``LinearOperator`` whose ``.tocsr()`` is
    the reference's matrix; otherwise applies it to ``values`` (numpy in -> numpy out, CUDA tensor in -> CUDA tensor out)
    and, like the reference, writes into ``output`` when one is given."""
    try:
        in_type, out_type = mesh_in._meshType, mesh_out._meshType
    except AttributeError:
        raise TypeError("Both input and output mesh must be valid discetize meshes")
    if in_type not in ("TENSOR", "TREE") or out_type not in ("TENSOR", "TREE"):
        raise NotImplementedError(
            f"Volume averaging is only implemented for TensorMesh and TreeMesh, "
            f"not {type(mesh_in).__name__} and/or {type(mesh_out).__name__}")
    if mesh_in.dim != mesh_out.dim:
        raise ValueError("Both meshes must have the same dimension")
    if values is not None and len(values) != mesh_in.n_cells:
        raise ValueError("Input array does not have the same length as the number of cells in input mesh")
    if output is not None and len(output) != mesh_out.n_cells:
        raise ValueError("Output array does not have the same length as the number of cells in output mesh")
    from .csr import CsrOperator

    if in_type == "TENSOR" and out_type == "TENSOR":
        matrix = _tensor_volume_average_matrix(mesh_in, mesh_out)
    elif _same_base(mesh_in, mesh_out):
        matrix = _same_base_volume_average_matrix(mesh_in, mesh_out)
    else:
        matrix = _overlap_volume_average_matrix(mesh_in, mesh_out)
    op = CsrOperator(matrix)
    if values is None:
        return op
    result = op @ values
    if output is not None:
        output[...] = result
        return output
    return result


def load_mesh(file_name):
    """Load a mesh saved by ``mesh.save()`` of this package OR of the reference (utils/io_utils.py:8-40): the JSON names
    the class; ``discretize.TensorMesh`` / ``discretize.TreeMesh`` map to the classes of this package."""
    import json

    from . import TensorMesh, TreeMesh

    with open(file_name, "r") as f:
        items = json.load(f)
    items.pop("__module__", None)
    name = items.pop("__class__")
    classes = {"TensorMesh": TensorMesh, "TreeMesh": TreeMesh, "QuadTreeMesh": TreeMesh}
    if name not in classes:
        raise NotImplementedError(f"discretize_b200.load_mesh: {name} is not a mesh type of this package")
    if "_n" in items:
        items["shape_cells"] = items.pop("_n")
    return classes[name](**items)


from .mesh_utils import as_array_n_by_dim, is_scalar, mkvc, ndgrid, unpack_widths  # noqa: E402,F401
from .matrix_utils import (Identity, TensorType, Zero, av, av_extrap, cart2cyl, cartesian_to_cylindrical, cross2d,  # noqa: E402,F401
                           cyl2cart, cylindrical_to_cartesian, ddx, get_subarray, ind2sub, interpolation_matrix,
                           inverse_2x2_block_diagonal, inverse_3x3_block_diagonal, inverse_property_tensor, invert_blocks,
                           kron3, make_boundary_bool, make_property_tensor, rotate_points_from_normals,
                           rotation_matrix_from_normals, sdiag, sdinv, speye, spzeros, sub2ind)


# camelCase names the reference has removed: calling them raises with the pointer to the new name
# (utils/code_utils.py: deprecate_function(..., error=True))
_REMOVED_FUNCTIONS = {
    "isScalar": "is_scalar", "asArray_N_x_Dim": "as_array_n_by_dim", "sdInv": "sdinv", "getSubArray": "get_subarray",
    "inv3X3BlockDiagonal": "inverse_3x3_block_diagonal", "inv2X2BlockDiagonal": "inverse_2x2_block_diagonal",
    "makePropertyTensor": "make_property_tensor", "invPropertyTensor": "inverse_property_tensor",
    "meshTensor": "unpack_widths",
    "interpmat": "interpolation_matrix", "rotationMatrixFromNormals": "rotation_matrix_from_normals",
    "rotatePointsFromNormals": "rotate_points_from_normals",
}


def _removed_function(old, new):
    def removed(*args, **kwargs):
        raise NotImplementedError(f"{old} has been removed, please use {new}.")

    removed.__name__ = old
    return removed


for _old, _new in _REMOVED_FUNCTIONS.items():
    globals()[_old] = _removed_function(_old, _new)
del _old, _new

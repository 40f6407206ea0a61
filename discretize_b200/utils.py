"""Utilities that sit next to the operator path (reference: ``discretize.utils``).

``volume_average`` -- conservative model transfer between two TensorMeshes: every output cell takes the
overlap-volume weighted mean of the input cells it intersects (``utils/interpolation_utils.py:175-287`` ->
``_extensions/interputils_cython.pyx:186-332``).  The operator is separable, so it is assembled on the host from three
1-D overlap lists with array passes (no per-entry loop) and applied on the GPU by the generic CSR kernel.
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
    if in_type != "TENSOR" or out_type != "TENSOR":
        raise NotImplementedError("discretize_b200.utils.volume_average: only TensorMesh -> TensorMesh is built")
    from .csr import CsrOperator

    op = CsrOperator(_tensor_volume_average_matrix(mesh_in, mesh_out))
    if values is None:
        return op
    result = op @ values
    if output is not None:
        output[...] = result
        return output
    return result

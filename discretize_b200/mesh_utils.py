"""The two array helpers the operator path itself uses (``mkvc``, ``ndgrid``) plus the width / scalar helpers of
``tensor_mesh``.  The reference's mesh-construction conveniences (``mesh_builder_xyz``, ``refine_tree_xyz``,
``active_from_xyz``, ``extract_core_mesh``, ``random_model``) are outside the hot-path scope and
are not part of this package."""
from __future__ import annotations

import numpy as np

from .tensor_mesh import TensorMesh, as_array_n_by_dim, is_scalar, unpack_widths  # noqa: F401  (re-exported)


def mkvc(x, n_dims=1, **kwargs):
    """Flatten in column-major order to a vector with ``n_dims`` dimensions (matrix_utils.py:35-110)."""
    if isinstance(x, np.matrix):
        x = np.array(x)
    if hasattr(x, "tovec"):
        x = x.tovec()
    if not isinstance(x, np.ndarray):
        raise TypeError("Vector must be a numpy array")
    if n_dims == 1:
        return x.flatten(order="F")
    if n_dims == 2:
        return x.flatten(order="F")[:, np.newaxis]
    if n_dims == 3:
        return x.flatten(order="F")[:, np.newaxis, np.newaxis]
    raise ValueError("n_dims must be 1, 2 or 3")


def ndgrid(*args, vector=True, order="F"):
    """Gridded locations of 1-D vectors, first axis fastest (matrix_utils.py:307-400)."""
    if not isinstance(vector, bool):
        raise TypeError("'vector' must be a boolean")
    xin = args[0] if isinstance(args[0], list) else list(args)
    if len(xin) == 1:
        return xin[0]
    grids = np.meshgrid(*xin, indexing="ij")
    if vector:
        return np.stack([g.reshape(-1, order=order) for g in grids], axis=1)
    return tuple(grids)

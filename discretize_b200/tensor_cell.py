"""One cell of a TensorMesh as an object: what ``mesh[i]``, ``mesh[i, j, k]`` and iteration return (reference:
``discretize/tensor_cell.py``, ``tensor_mesh.py:166-310``).  All numbering comes from one rule: an entity family is the
cell grid extended by one along its "node axes", and a cell touches the entities at its own index plus 0 / 1 along those
axes, listed with the lower axis changing fastest."""
from __future__ import annotations

import itertools

import numpy as np


class TensorCell:
    def __init__(self, h, origin, index_unraveled, mesh_shape):
        self._h, self._origin, self._index_unraveled, self._mesh_shape = h, origin, index_unraveled, mesh_shape

    h = property(lambda self: self._h)
    origin = property(lambda self: self._origin)
    index_unraveled = property(lambda self: self._index_unraveled)
    mesh_shape = property(lambda self: self._mesh_shape)
    dim = property(lambda self: len(self._h))
    index = property(lambda self: np.ravel_multi_index(self._index_unraveled, dims=self._mesh_shape, order="F"))
    center = property(lambda self: np.array(self._origin) + np.array(self._h) / 2)

    def __repr__(self):
        return "TensorCell(" + ", ".join(f"{a}={getattr(self, a)}" for a in ("h", "origin", "index", "mesh_shape")) + ")"

    def __eq__(self, other):
        if not isinstance(other, TensorCell):
            raise TypeError(f"Cannot compare an object of type '{other.__class__.__name__}' with a TensorCell")
        return bool(np.all(self.h == other.h) and np.all(self.origin == other.origin) and self.index == other.index
                    and self.mesh_shape == other.mesh_shape)

    __hash__ = None

    @property
    def bounds(self):
        return np.array([o + f * w for o, w in zip(self._origin, self._h) for f in (0, 1)])

    @property
    def neighbors(self):
        """Indices of the cells sharing a face: -x, +x, -y, +y, ... (those that exist)."""
        out = []
        for axis in range(self.dim):
            for delta in (-1, 1):
                idx = list(self._index_unraveled)
                idx[axis] += delta
                if 0 <= idx[axis] < self._mesh_shape[axis]:
                    out.append(np.ravel_multi_index(idx, dims=self._mesh_shape, order="F"))
        return out

    def _touching(self, node_axes, offset=0):
        """Numbers of the entities of one family (cell grid + 1 along ``node_axes``) that touch this cell."""
        shape = [n + 1 if a in node_axes else n for a, n in enumerate(self._mesh_shape)]
        out = []
        for deltas in itertools.product((0, 1), repeat=len(node_axes)):
            idx = list(self._index_unraveled)
            for a, dlt in zip(reversed(node_axes), deltas):      # product varies its LAST factor fastest
                idx[a] += dlt
            out.append(offset + np.ravel_multi_index(idx, dims=shape, order="F"))
        return out

    def _family_size(self, node_axes):
        return int(np.prod([n + 1 if a in node_axes else n for a, n in enumerate(self._mesh_shape)]))

    @property
    def nodes(self):
        return self._touching(tuple(range(self.dim)))

    @property
    def edges(self):
        if self.dim == 1:
            return [self.index]
        out, offset = [], 0
        for d in range(self.dim):
            axes = tuple(a for a in range(self.dim) if a != d)
            out += self._touching(axes, offset)
            offset += self._family_size(axes)
        return out

    @property
    def faces(self):
        if self.dim == 1:
            return [self.index, self.index + 1]
        out, offset = [], 0
        for d in range(self.dim):
            out += self._touching((d,), offset)
            offset += self._family_size((d,))
        return out

    def get_neighbors(self, mesh):
        return [mesh[index] for index in self.neighbors]

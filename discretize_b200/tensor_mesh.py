"""TensorMesh: the reference's names, B200 operators.

Drop-in for the operator-application part of ``discretize.TensorMesh``
(reference: discretize/tensor_mesh.py:15-22 and its mixins).  Geometry and counts
are host-side numpy exactly like the reference; every operator property returns a
CUDA-backed :class:`~discretize_b200.operators.Operator`.

Only 3-D meshes have kernels (the five target configurations are 3-D); 1-D/2-D
meshes construct (counts, geometry) but their operators raise NotImplementedError.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import _lib
from .operators import to_device, cuda_device

_ALIASES = {
    # reference: discretize/base/base_mesh.py:24-48 and operators/differential_operators.py:147-154
    "nC": "n_cells", "nN": "n_nodes", "nE": "n_edges", "nF": "n_faces",
    "nEx": "n_edges_x", "nEy": "n_edges_y", "nEz": "n_edges_z",
    "nFx": "n_faces_x", "nFy": "n_faces_y", "nFz": "n_faces_z",
    "vnC": "shape_cells", "vnN": "shape_nodes",
    "vnEx": "shape_edges_x", "vnEy": "shape_edges_y", "vnEz": "shape_edges_z",
    "vnFx": "shape_faces_x", "vnFy": "shape_faces_y", "vnFz": "shape_faces_z",
    "vnE": "n_edges_per_direction", "vnF": "n_faces_per_direction",
    "x0": "origin", "vol": "cell_volumes", "area": "face_areas", "edge": "edge_lengths",
    "gridCC": "cell_centers", "gridN": "nodes",
    "vectorNx": "nodes_x", "vectorNy": "nodes_y", "vectorNz": "nodes_z",
    "vectorCCx": "cell_centers_x", "vectorCCy": "cell_centers_y", "vectorCCz": "cell_centers_z",
    "hx": "h_x", "hy": "h_y", "hz": "h_z",
    "aveF2CC": "average_face_to_cell", "aveF2CCV": "average_face_to_cell_vector",
    "aveFx2CC": "average_face_x_to_cell", "aveFy2CC": "average_face_y_to_cell", "aveFz2CC": "average_face_z_to_cell",
    "aveE2CC": "average_edge_to_cell", "aveE2CCV": "average_edge_to_cell_vector",
    "aveEx2CC": "average_edge_x_to_cell", "aveEy2CC": "average_edge_y_to_cell", "aveEz2CC": "average_edge_z_to_cell",
    "aveN2CC": "average_node_to_cell",
    "aveCC2F": "average_cell_to_face", "aveCCV2F": "average_cell_vector_to_face",
    "aveN2E": "average_node_to_edge", "aveN2F": "average_node_to_face",
    "nodalGrad": "nodal_gradient", "edgeCurl": "edge_curl", "faceDiv": "face_divergence",
    "getFaceInnerProduct": "get_face_inner_product", "getEdgeInnerProduct": "get_edge_inner_product",
    "getFaceInnerProductDeriv": "get_face_inner_product_deriv",
    "getEdgeInnerProductDeriv": "get_edge_inner_product_deriv",
    "getInterpolationMat": "get_interpolation_matrix",
    "isInside": "is_inside", "getTensor": "get_tensor",
}


def _validate_BC(bc):
    """reference: operators/differential_operators.py:21-38."""
    if isinstance(bc, str):
        bc = [bc, bc]
    if not isinstance(bc, list):
        raise TypeError("bc must be a single string or list of strings")
    if not len(bc) == 2:
        raise TypeError("bc list must have two elements, one for each side")
    for bc_i in bc:
        if not isinstance(bc_i, str):
            raise TypeError("each bc must be a string")
        if bc_i not in ["dirichlet", "neumann"]:
            raise ValueError("each bc must be either, 'dirichlet' or 'neumann'")
    return bc


def is_scalar(f):
    """reference: utils/code_utils.py:9-34 (python/numpy scalars and 1-element arrays)."""
    if isinstance(f, (int, float, complex, np.number)):
        return True
    if isinstance(f, np.ndarray) and f.size == 1 and isinstance(f.reshape(-1)[0], (np.number, np.bool_)):
        return True
    return False


def unpack_widths(value):
    """[w, (w, n), (w, n, factor)] -> widths (reference: utils/mesh_utils.py:117-185)."""
    if type(value) is not list:
        raise Exception("unpack_widths must be a list of scalars and tuples.")
    out = []
    for v in value:
        if is_scalar(v):
            out.append(float(v))
        elif type(v) is tuple and len(v) == 2:
            out.extend([float(v[0])] * int(v[1]))
        elif type(v) is tuple and len(v) == 3:
            w0, num, fac = float(v[0]), int(v[1]), float(v[2])
            pad = (np.abs(fac) * np.ones(num)) ** (np.arange(num) + 1) * w0
            out.extend((pad[::-1] if fac < 0 else pad).tolist())
        else:
            raise Exception("unpack_widths must contain only scalars and len(2) or len(3) tuples.")
    return np.array(out)


def as_array_n_by_dim(pts, dim):
    """reference: utils/code_utils.py:37-77."""
    if type(pts) is list:
        pts = np.array(pts)
    if not isinstance(pts, np.ndarray):
        raise TypeError("pts must be a numpy array")
    if dim > 1:
        pts = np.atleast_2d(pts)
    elif len(pts.shape) == 1:
        pts = pts[:, np.newaxis]
    if pts.shape[1] != dim:
        raise ValueError("pts must be a column vector of shape (nPts, {0:d}) not ({1:d}, {2:d})".format(*((dim,) + pts.shape)))
    return pts


class TensorMesh:
    """Tensor-product mesh with CUDA operators (reference: discretize/tensor_mesh.py)."""

    _meshType = "TENSOR"

    def __init__(self, h, origin=None, **kwargs):
        # reference: base/base_tensor_mesh.py:86-114
        if "x0" in kwargs:
            origin = kwargs.pop("x0")
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        try:
            h = list(h)
        except TypeError:
            raise TypeError("h must be an iterable object, not {}".format(type(h)))
        if len(h) == 0 or len(h) > 3:
            raise ValueError("h must be of dimension 1, 2, or 3 not {}".format(len(h)))
        for i, h_i in enumerate(h):
            if is_scalar(h_i) and not isinstance(h_i, np.ndarray):
                h_i = np.ones(int(h_i)) / int(h_i)
            elif isinstance(h_i, (list, tuple)):
                h_i = unpack_widths(list(h_i))
            if not isinstance(h_i, np.ndarray):
                raise TypeError("h[{0:d}] is not a numpy array.".format(i))
            if len(h_i.shape) != 1:
                raise ValueError("h[{0:d}] must be a 1D numpy array.".format(i))
            h[i] = np.array(h_i, dtype=np.float64)
        self._h = tuple(h)
        self._origin = np.zeros(len(h))
        if origin is not None:
            self.origin = origin
        self._cache = {}
        self._dev = None

    def __getattr__(self, name):
        # alias table, like BaseMesh.__getattr__ (base/base_mesh.py:24-48)
        if name in _ALIASES:
            return getattr(self, _ALIASES[name])
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    # ---- basic description ---------------------------------------------------------------
    @property
    def h(self):
        return self._h

    @property
    def h_x(self):
        return self._h[0]

    @property
    def h_y(self):
        return None if self.dim < 2 else self._h[1]

    @property
    def h_z(self):
        return None if self.dim < 3 else self._h[2]

    @property
    def dim(self):
        return len(self._h)

    @property
    def origin(self):
        return self._origin

    @origin.setter
    def origin(self, value):
        # reference: base/base_tensor_mesh.py:138-153
        try:
            value = list(value)
        except TypeError:
            raise TypeError("origin must be iterable")
        if len(value) != self.dim:
            raise ValueError("Dimension mismatch. len(origin) != len(h)")
        for i, (val, h_i) in enumerate(zip(value, self.h)):
            if isinstance(val, str):
                if val == "C":
                    value[i] = -h_i.sum() * 0.5
                elif val == "N":
                    value[i] = -h_i.sum()
                elif val == "0":
                    value[i] = 0.0
        self._origin = np.asarray(value, dtype=np.float64)
        self._cache = {}

    @property
    def shape_cells(self):
        return tuple(int(x.size) for x in self._h)

    @property
    def shape_nodes(self):
        return tuple(n + 1 for n in self.shape_cells)

    def _shape(self, node_axes):
        return tuple(n + (1 if a in node_axes else 0) for a, n in enumerate(self.shape_cells))

    @property
    def shape_faces_x(self):
        return self._shape((0,))

    @property
    def shape_faces_y(self):
        return None if self.dim < 2 else self._shape((1,))

    @property
    def shape_faces_z(self):
        return None if self.dim < 3 else self._shape((2,))

    @property
    def shape_edges_x(self):
        return self._shape((1, 2))

    @property
    def shape_edges_y(self):
        return None if self.dim < 2 else self._shape((0, 2))

    @property
    def shape_edges_z(self):
        return None if self.dim < 3 else self._shape((0, 1))

    @property
    def n_cells(self):
        return int(np.prod(self.shape_cells))

    @property
    def n_nodes(self):
        return int(np.prod(self.shape_nodes))

    @property
    def n_faces_x(self):
        return int(np.prod(self.shape_faces_x))

    @property
    def n_faces_y(self):
        return None if self.dim < 2 else int(np.prod(self.shape_faces_y))

    @property
    def n_faces_z(self):
        return None if self.dim < 3 else int(np.prod(self.shape_faces_z))

    @property
    def n_edges_x(self):
        return int(np.prod(self.shape_edges_x))

    @property
    def n_edges_y(self):
        return None if self.dim < 2 else int(np.prod(self.shape_edges_y))

    @property
    def n_edges_z(self):
        return None if self.dim < 3 else int(np.prod(self.shape_edges_z))

    @property
    def n_faces_per_direction(self):
        return tuple(x for x in (self.n_faces_x, self.n_faces_y, self.n_faces_z) if x is not None)

    @property
    def n_edges_per_direction(self):
        return tuple(x for x in (self.n_edges_x, self.n_edges_y, self.n_edges_z) if x is not None)

    @property
    def n_faces(self):
        return int(sum(self.n_faces_per_direction))

    @property
    def n_edges(self):
        return int(sum(self.n_edges_per_direction))

    # ---- 1-D vectors and grids (base/base_tensor_mesh.py:155-240, 567-634) ------------------
    def _nodes_1d(self, a):
        key = ("nodes", a)
        if key not in self._cache:
            self._cache[key] = np.r_[self._origin[a], self._h[a]].cumsum()
        return self._cache[key]

    def _centers_1d(self, a):
        n = self._nodes_1d(a)
        return (n[1:] + n[:-1]) / 2

    nodes_x = property(lambda self: self._nodes_1d(0))
    nodes_y = property(lambda self: None if self.dim < 2 else self._nodes_1d(1))
    nodes_z = property(lambda self: None if self.dim < 3 else self._nodes_1d(2))
    cell_centers_x = property(lambda self: self._centers_1d(0))
    cell_centers_y = property(lambda self: None if self.dim < 2 else self._centers_1d(1))
    cell_centers_z = property(lambda self: None if self.dim < 3 else self._centers_1d(2))

    _LOC_NAMES = {
        "CC": "cell_centers", "N": "nodes", "Fx": "faces_x", "Fy": "faces_y", "Fz": "faces_z",
        "Ex": "edges_x", "Ey": "edges_y", "Ez": "edges_z",
        "CCVx": "cell_centers_x", "CCVy": "cell_centers_y", "CCVz": "cell_centers_z",
    }

    def _parse_location_type(self, location_type):
        # reference: base/base_mesh.py:4182-4203
        if len(location_type) == 0:
            return location_type
        if location_type in self._LOC_NAMES.values():
            return location_type
        if location_type in self._LOC_NAMES:
            return self._LOC_NAMES[location_type]
        raise ValueError(f"Unrecognized location type {location_type}")

    def get_tensor(self, key):
        """The 1-D sample vectors of a location grid (base/base_tensor_mesh.py:567-634)."""
        key = self._parse_location_type(key)
        if key == "cell_centers":
            node_axes = ()
        elif key == "nodes":
            node_axes = (0, 1, 2)
        elif key.startswith("faces_"):
            node_axes = ("xyz".index(key[-1]),)
        elif key.startswith("edges_"):
            node_axes = tuple(a for a in range(3) if a != "xyz".index(key[-1]))
        else:
            raise ValueError(f"Unrecognized location type {key}")
        return [self._nodes_1d(a) if a in node_axes else self._centers_1d(a) for a in range(self.dim)]

    def _grid(self, key):
        t = self.get_tensor(key)
        mesh = np.meshgrid(*t, indexing="ij")
        return np.stack([m.reshape(-1, order="F") for m in mesh], axis=1)

    cell_centers = property(lambda self: self._grid("cell_centers"))
    nodes = property(lambda self: self._grid("nodes"))
    faces_x = property(lambda self: self._grid("faces_x"))
    faces_y = property(lambda self: self._grid("faces_y"))
    faces_z = property(lambda self: self._grid("faces_z"))
    edges_x = property(lambda self: self._grid("edges_x"))
    edges_y = property(lambda self: self._grid("edges_y"))
    edges_z = property(lambda self: self._grid("edges_z"))

    def is_inside(self, pts, location_type="nodes"):
        """reference: base/base_tensor_mesh.py:638-682."""
        pts = as_array_n_by_dim(pts, self.dim)
        tensors = self.get_tensor(location_type)
        inside = np.ones(pts.shape[0], dtype=bool)
        for i, tensor in enumerate(tensors):
            tol = np.diff(tensor).min() * 1.0e-10
            inside = inside & (pts[:, i] >= tensor.min() - tol) & (pts[:, i] <= tensor.max() + tol)
        return inside

    # ---- geometry vectors (tensor_mesh.py:296-517), host numpy like the reference ------------
    def _outer(self, *vs):
        out = vs[0]
        for v in vs[1:]:
            out = np.outer(out, v).reshape(-1, order="F")
        return out

    @property
    def cell_volumes(self):
        return self._outer(*self._h)

    @property
    def face_areas(self):
        if self.dim < 3:
            from . import lowdim

            return lowdim.face_areas(self)
        hx, hy, hz = self._h
        nx, ny, nz = self.shape_cells
        return np.r_[self._outer(np.ones(nx + 1), self._outer(hy, hz)),
                     self._outer(hx, self._outer(np.ones(ny + 1), hz)),
                     self._outer(hx, self._outer(hy, np.ones(nz + 1)))]

    @property
    def edge_lengths(self):
        if self.dim < 3:
            from . import lowdim

            return lowdim.edge_lengths(self)
        hx, hy, hz = self._h
        nx, ny, nz = self.shape_cells
        return np.r_[self._outer(hx, np.ones((ny + 1) * (nz + 1))),
                     self._outer(np.ones(nx + 1), self._outer(hy, np.ones(nz + 1))),
                     self._outer(np.ones((nx + 1) * (ny + 1)), hz)]

    # ---- device state ---------------------------------------------------------------------------
    def _require_3d(self, what):
        if self.dim != 3:
            raise NotImplementedError(
                f"discretize_b200.TensorMesh.{what}: only 3-D meshes have CUDA kernels (mesh.dim == {self.dim})"
            )

    def device_mesh(self):
        """DzbTensorMesh struct with the 1-D widths (and reciprocals) resident on the GPU."""
        self._require_3d("operators")
        dev = cuda_device()
        if self._dev is None or self._dev[0] != dev:
            hs = [to_device(h) for h in self._h]
            rhs = [to_device(1.0 / h) for h in self._h]
            nx, ny, nz = self.shape_cells
            st = _lib.DzbTensorMesh(nx, ny, nz, hs[0].data_ptr(), hs[1].data_ptr(), hs[2].data_ptr(),
                                    rhs[0].data_ptr(), rhs[1].data_ptr(), rhs[2].data_ptr())
            self._dev = (dev, st, hs, rhs)
        return self._dev[1]

    def _host_mesh_counts(self):
        nx, ny, nz = self.shape_cells
        return _lib.DzbTensorMesh(nx, ny, nz, 0, 0, 0, 0, 0, 0)

    # ---- differential operators and averaging ------------------------------------------------------
    def _lowdim_operator(self, name):
        """1-D / 2-D meshes: the operator assembled on the host (lowdim.py) and applied by the generic CSR kernel."""
        key = ("lowdim", name)
        if key not in self._cache:
            from . import lowdim
            from .csr import CsrOperator

            self._cache[key] = CsrOperator(lowdim.build(self, name))
        return self._cache[key]

    def _lowdim_cell_gradient(self, name):
        from . import lowdim
        from .csr import CsrOperator

        axis = "xyz".index(name[-1]) if name[-2:] in ("_x", "_y", "_z") else None
        if axis is not None and axis >= self.dim:
            raise NotImplementedError(f"discretize_b200.TensorMesh.{name} is not available for mesh.dim == {self.dim}")
        bcs = self.set_cell_gradient_BC(self._cell_gradient_BC_list)
        key = ("lowdim", name, str(bcs))
        if key not in self._cache:
            self._cache[key] = CsrOperator(lowdim.cell_gradient(self, bcs, not name.startswith("stencil"), axis))
        return self._cache[key]

    def _stencil(self, name, op_id, flags=0):
        if self.dim < 3:
            if "cell_gradient" in name:
                return self._lowdim_cell_gradient(name)
            return self._lowdim_operator(name)
        key = name if flags == 0 else (name, flags)
        if key not in self._cache:
            from .tensor_ops import StencilOperator

            self._cache[key] = StencilOperator(self, op_id, name=name, flags=flags)
        return self._cache[key]

    face_divergence = property(lambda self: self._stencil("face_divergence", _lib.OP_DIV))
    nodal_gradient = property(lambda self: self._stencil("nodal_gradient", _lib.OP_GRAD))
    average_face_to_cell = property(lambda self: self._stencil("average_face_to_cell", _lib.OP_AVE_F2CC))
    average_face_to_cell_vector = property(lambda self: self._stencil("average_face_to_cell_vector", _lib.OP_AVE_F2CCV))
    average_face_x_to_cell = property(lambda self: self._stencil("average_face_x_to_cell", _lib.OP_AVE_FX2CC))
    average_face_y_to_cell = property(lambda self: self._stencil("average_face_y_to_cell", _lib.OP_AVE_FY2CC))
    average_face_z_to_cell = property(lambda self: self._stencil("average_face_z_to_cell", _lib.OP_AVE_FZ2CC))
    average_edge_to_cell = property(lambda self: self._stencil("average_edge_to_cell", _lib.OP_AVE_E2CC))
    average_edge_to_cell_vector = property(lambda self: self._stencil("average_edge_to_cell_vector", _lib.OP_AVE_E2CCV))
    average_edge_x_to_cell = property(lambda self: self._stencil("average_edge_x_to_cell", _lib.OP_AVE_EX2CC))
    average_edge_y_to_cell = property(lambda self: self._stencil("average_edge_y_to_cell", _lib.OP_AVE_EY2CC))
    average_edge_z_to_cell = property(lambda self: self._stencil("average_edge_z_to_cell", _lib.OP_AVE_EZ2CC))
    average_node_to_cell = property(lambda self: self._stencil("average_node_to_cell", _lib.OP_AVE_N2CC))

    # ---- cell-centred side (operators/differential_operators.py:236-591, 975-2083, 2789-2892,
    # 3205-3382): same matrix-free row generators, boundary conditions as runtime flags ------------------------
    face_x_divergence = property(lambda self: self._stencil("face_x_divergence", _lib.OP_DIV_X))
    face_y_divergence = property(lambda self: self._stencil("face_y_divergence", _lib.OP_DIV_Y))
    face_z_divergence = property(lambda self: self._stencil("face_z_divergence", _lib.OP_DIV_Z))
    average_cell_to_face = property(lambda self: self._stencil("average_cell_to_face", _lib.OP_AVE_CC2F))
    average_cell_vector_to_face = property(lambda self: self._stencil("average_cell_vector_to_face", _lib.OP_AVE_CCV2F))
    average_cell_to_edge = property(lambda self: self._stencil("average_cell_to_edge", _lib.OP_AVE_CC2E))
    average_edge_to_face = property(lambda self: self._stencil("average_edge_to_face", _lib.OP_AVE_E2F))
    average_node_to_edge = property(lambda self: self._stencil("average_node_to_edge", _lib.OP_AVE_N2E))
    average_node_to_face = property(lambda self: self._stencil("average_node_to_face", _lib.OP_AVE_N2F))

    _cell_gradient_BC_list = "neumann"

    def set_cell_gradient_BC(self, BC):
        """reference: operators/differential_operators.py:975-1046 (zero Dirichlet / Neumann per side)."""
        if isinstance(BC, str):
            BC = [BC] * self.dim
        if isinstance(BC, list):
            if len(BC) != self.dim:
                raise ValueError("BC list must be the size of your mesh")
        else:
            raise TypeError("BC must be a str or a list.")
        for i, bc_i in enumerate(BC):
            BC[i] = _validate_BC(bc_i)
        self._cell_gradient_BC_list = BC
        return BC

    def _cell_gradient_flags(self):
        BC = self.set_cell_gradient_BC(self._cell_gradient_BC_list)
        flags = 0
        for a, (lo, hi) in enumerate(BC):
            flags |= (1 << (2 * a)) if lo == "dirichlet" else 0
            flags |= (2 << (2 * a)) if hi == "dirichlet" else 0
        return flags

    @property
    def stencil_cell_gradient(self):
        return self._stencil("stencil_cell_gradient", _lib.OP_CELL_GRAD, self._cell_gradient_flags())

    @property
    def cell_gradient(self):
        """diag(S / (aveCC2F V)) @ stencil_cell_gradient (:1445-1595).  NOTE: the reference caches the first
        matrix it builds until set_cell_gradient_BC is called again; here the current BC always applies."""
        return self._stencil("cell_gradient", _lib.OP_CELL_GRAD, self._cell_gradient_flags() | _lib.CG_SCALE)

    # the component operators always use Neumann conditions (:1167, 1296, 1400)
    stencil_cell_gradient_x = property(lambda self: self._stencil("stencil_cell_gradient_x", _lib.OP_CELL_GRAD_X))
    stencil_cell_gradient_y = property(lambda self: self._stencil("stencil_cell_gradient_y", _lib.OP_CELL_GRAD_Y))
    stencil_cell_gradient_z = property(lambda self: self._stencil("stencil_cell_gradient_z", _lib.OP_CELL_GRAD_Z))
    cell_gradient_x = property(lambda self: self._stencil("cell_gradient_x", _lib.OP_CELL_GRAD_X, _lib.CG_SCALE))
    cell_gradient_y = property(lambda self: self._stencil("cell_gradient_y", _lib.OP_CELL_GRAD_Y, _lib.CG_SCALE))
    cell_gradient_z = property(lambda self: self._stencil("cell_gradient_z", _lib.OP_CELL_GRAD_Z, _lib.CG_SCALE))

    # ---- boundary (base/base_tensor_mesh.py:300-310, 423-560; operators/differential_operators.py:2183-2197,
    # 3384-3430): index selections, assembled on the host and applied by the generic CSR kernel ----------------
    @staticmethod
    def _boundary_mask(shape, axes):
        """Flattened (x fastest) mask of the entries of a tensor of ``shape`` that sit on its first or last index along
        any of ``axes`` (utils/matrix_utils.py: make_boundary_bool)."""
        mask = np.zeros(shape, dtype=bool)
        for a in axes:
            sel = [slice(None)] * len(shape)
            for end in (0, -1):
                sel[a] = end
                mask[tuple(sel)] = True
        return mask.reshape(-1, order="F")

    def _boundary_masks(self):
        d = self.dim
        face = np.concatenate([self._boundary_mask(self._shape((a,)), (a,)) for a in range(d)])
        edge = None
        if d > 1:
            edge = np.concatenate([self._boundary_mask(self._shape(tuple(b for b in range(d) if b != a)),
                                                       tuple(b for b in range(d) if b != a)) for a in range(d)])
        node = self._boundary_mask(self.shape_nodes, tuple(range(d)))
        return face, edge, node

    def _selection(self, key, mask):
        if key not in self._cache:
            from .csr import CsrOperator

            idx = np.nonzero(mask)[0]
            P = sp.csr_matrix((np.ones(idx.size), (np.arange(idx.size), idx)), shape=(idx.size, mask.size))
            self._cache[key] = CsrOperator(P)
        return self._cache[key]

    @property
    def project_face_to_boundary_face(self):
        return self._selection("P_bf", self._boundary_masks()[0])

    @property
    def project_edge_to_boundary_edge(self):
        if self.dim == 1:
            return None  # no edges are on the boundary in 1-D
        return self._selection("P_be", self._boundary_masks()[1])

    @property
    def project_node_to_boundary_node(self):
        return self._selection("P_bn", self._boundary_masks()[2])

    @property
    def boundary_nodes(self):
        if self.dim == 1:
            return self.nodes_x[[0, -1]]
        return self.nodes[self._boundary_masks()[2]]

    def _face_grids(self):
        return [self._grid("faces_" + "xyz"[a]) for a in range(self.dim)]

    @property
    def boundary_faces(self):
        if self.dim == 1:
            return self.nodes_x[[0, -1]]
        return np.concatenate(self._face_grids())[self._boundary_masks()[0]]

    @property
    def boundary_edges(self):
        if self.dim == 1:
            return None
        grids = np.concatenate([self._grid("edges_" + "xyz"[a]) for a in range(self.dim)])
        return grids[self._boundary_masks()[1]]

    @property
    def boundary_face_outward_normals(self):
        d = self.dim
        if d == 1:
            return np.array([-1, 1])
        out = []
        for a in range(d):
            shp = self._shape((a,))
            idx = np.indices(shp)[a].reshape(-1, order="F")
            on = self._boundary_mask(shp, (a,))
            nrm = np.zeros((int(on.sum()), d))
            nrm[:, a] = np.where(idx[on] == 0, -1.0, 1.0)
            out.append(nrm)
        return np.concatenate(out)

    @property
    def boundary_face_scalar_integral(self):
        """diag(face_areas) P^T diag(+-1): nF x n_boundary_faces (differential_operators.py:2183-2197)."""
        from .csr import CsrOperator

        if self.dim == 1:
            return CsrOperator(sp.csr_matrix(([-1.0, 1.0], ([0, self.n_faces_x - 1], [0, 1])), shape=(self.n_faces_x, 2)))
        P = self.project_face_to_boundary_face.tocsr()
        sign = self.boundary_face_outward_normals.sum(axis=1)       # the one non-zero component: -1 (low side) or +1
        return CsrOperator((sp.diags(self.face_areas) @ P.T @ sp.diags(sign)).tocsr())

    @property
    def nodal_laplacian(self):
        # the reference's own 3-D assembly fails with a dimension mismatch (:719-748 multiply an n_edges_x
        # diagonal into an n_nodes-column stencil), so there is nothing to be compatible with
        raise NotImplementedError("nodal_laplacian: the reference implementation raises for 3-D tensor meshes")

    @property
    def edge_curl(self):
        if self.dim <= 1:
            raise NotImplementedError("Edge Curl only programed for 2 or 3D.")
        return self._stencil("edge_curl", _lib.OP_CURL)

    # ---- mass matrices (operators/inner_products.py:35-99) ---------------------------------------------
    def get_face_inner_product(self, model=None, invert_model=False, invert_matrix=False, do_fast=True, **kwargs):
        return self._inner_product("F", model, invert_model, invert_matrix, do_fast, kwargs)

    def get_edge_inner_product(self, model=None, invert_model=False, invert_matrix=False, do_fast=True, **kwargs):
        return self._inner_product("E", model, invert_model, invert_matrix, do_fast, kwargs)

    @staticmethod
    def _check_removed_kwargs(kwargs):
        # reference: inner_products.py:44-58, 176-190
        for old, new in (("invProp", "invert_model"), ("invMat", "invert_matrix"), ("doFast", "do_fast")):
            if old in kwargs:
                raise TypeError(
                    f"The {old} keyword argument has been removed, please use {new}. "
                    "This will be removed in discretize 1.0.0"
                )
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")

    def _inner_product(self, proj, model, invert_model, invert_matrix, do_fast, kwargs):
        self._check_removed_kwargs(kwargs)
        if self.dim < 3:
            import torch

            from . import lowdim
            from .operators import DiagonalOperator

            host = model.cpu().numpy() if isinstance(model, torch.Tensor) else model
            return DiagonalOperator(lowdim.mass_diagonal(self, proj, host, invert_model, invert_matrix))
        from .tensor_ops import make_mass_operator

        return make_mass_operator(self, proj, model, invert_model, invert_matrix, do_fast)

    def get_face_inner_product_deriv(self, model, do_fast=True, invert_model=False, invert_matrix=False, **kwargs):
        return self._inner_product_deriv("F", model, do_fast, invert_model, invert_matrix, kwargs)

    def get_edge_inner_product_deriv(self, model, do_fast=True, invert_model=False, invert_matrix=False, **kwargs):
        return self._inner_product_deriv("E", model, do_fast, invert_model, invert_matrix, kwargs)

    def _inner_product_deriv(self, proj, model, do_fast, invert_model, invert_matrix, kwargs):
        self._check_removed_kwargs(kwargs)
        if self.dim < 3:
            if model is None:
                return None
            import torch

            from . import lowdim
            from .csr import CsrOperator

            host = model.cpu().numpy() if isinstance(model, torch.Tensor) else model
            if np.asarray(host).size not in (1, self.n_cells, self.n_cells * self.dim):
                raise NotImplementedError("full-tensor model derivatives are only built for 3-D meshes")

            def inner_product_deriv(v=None):
                if v is None:
                    raise Exception("v must be supplied for this implementation.")
                vh = v.cpu().numpy() if isinstance(v, torch.Tensor) else v
                return CsrOperator(lowdim.mass_deriv_matrix(self, proj, host, vh, invert_model, invert_matrix))

            return inner_product_deriv
        from .tensor_ops import make_mass_deriv

        return make_mass_deriv(self, proj, model, do_fast, invert_model, invert_matrix)

    # ---- interpolation (base/base_tensor_mesh.py:684-790) -------------------------------------------------
    def get_interpolation_matrix(self, loc, location_type="cell_centers", zeros_outside=False, **kwargs):
        if "locType" in kwargs:
            raise TypeError(
                "The locType keyword argument has been removed, please use location_type. "
                "This will be removed in discretize 1.0.0"
            )
        if "zerosOutside" in kwargs:
            raise TypeError(
                "The zerosOutside keyword argument has been removed, please use zeros_outside. "
                "This will be removed in discretize 1.0.0"
            )
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        if self.dim < 3:
            import torch

            from . import lowdim
            from .csr import CsrOperator

            host = loc.cpu().numpy() if isinstance(loc, torch.Tensor) else loc
            return CsrOperator(lowdim.interpolation_matrix(self, host, location_type, zeros_outside))
        from .interp import make_interpolation_operator

        return make_interpolation_operator(self, loc, location_type, zeros_outside)

    def __repr__(self):
        return f"<discretize_b200.TensorMesh {'x'.join(str(n) for n in self.shape_cells)} cells>"

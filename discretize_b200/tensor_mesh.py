"""TensorMesh: the reference's names, B200 operators.

Drop-in for the operator-application part of ``discretize.TensorMesh``
(reference: discretize/tensor_mesh.py:15-22 and its mixins).  Geometry and counts
are host-side numpy exactly like the reference; every operator property returns a
CUDA-backed :class:`~discretize_b200.operators.Operator`.

The hand-written stencil kernels are 3-D (the five target configurations are 3-D); 1-D/2-D
meshes and the boundary / surface / line terms are assembled on the host (lowdim.py,
boundary_ops.py, surface_products.py) and applied on the GPU by the generic CSR kernel.
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
    "x0": "origin",
    "gridCC": "cell_centers", "gridN": "nodes",
    "aveF2CC": "average_face_to_cell", "aveF2CCV": "average_face_to_cell_vector",
    "aveFx2CC": "average_face_x_to_cell", "aveFy2CC": "average_face_y_to_cell", "aveFz2CC": "average_face_z_to_cell",
    "aveE2CC": "average_edge_to_cell", "aveE2CCV": "average_edge_to_cell_vector",
    "aveEx2CC": "average_edge_x_to_cell", "aveEy2CC": "average_edge_y_to_cell", "aveEz2CC": "average_edge_z_to_cell",
    "aveN2CC": "average_node_to_cell",
    "aveCC2F": "average_cell_to_face", "aveCCV2F": "average_cell_vector_to_face",
    "aveN2E": "average_node_to_edge", "aveN2F": "average_node_to_face",
    "gridFx": "faces_x", "gridFy": "faces_y", "gridFz": "faces_z", "gridEx": "edges_x", "gridEy": "edges_y", "gridEz": "edges_z",
    "setCellGradBC": "set_cell_gradient_BC",
    "projectFaceVector": "project_face_vector", "projectEdgeVector": "project_edge_vector",
    "getBCProjWF": "get_BC_projections", "getBCProjWF_simple": "get_BC_projections_simple",
    "getFaceInnerProduct": "get_face_inner_product", "getEdgeInnerProduct": "get_edge_inner_product",
    "getFaceInnerProductDeriv": "get_face_inner_product_deriv",
    "getEdgeInnerProductDeriv": "get_edge_inner_product_deriv",
    "getInterpolationMat": "get_interpolation_matrix",
    "isInside": "is_inside", "getTensor": "get_tensor",
}

# names the reference has removed: they raise NotImplementedError pointing at the new name (utils/code_utils.py:225-245)
_REMOVED = {
    "area": "face_areas", "areaFx": "face_x_areas", "areaFy": "face_y_areas", "areaFz": "face_z_areas",
    "cellBoundaryInd": "cell_boundary_indices", "cellGrad": "cell_gradient", "cellGradBC": "cell_gradient_BC",
    "cellGradx": "cell_gradient_x", "cellGrady": "cell_gradient_y", "cellGradz": "cell_gradient_z",
    "edge": "edge_lengths", "edgeCurl": "edge_curl", "edgeEx": "edge_x_lengths", "edgeEy": "edge_y_lengths",
    "edgeEz": "edge_z_lengths", "faceBoundaryInd": "face_boundary_indices", "faceDiv": "face_divergence",
    "faceDivx": "face_x_divergence", "faceDivy": "face_y_divergence", "faceDivz": "face_z_divergence",
    "nodalGrad": "nodal_gradient", "nodalLaplacian": "nodal_laplacian", "normals": "face_normals",
    "tangents": "edge_tangents", "vectorCCx": "cell_centers_x", "vectorCCy": "cell_centers_y",
    "vectorCCz": "cell_centers_z", "vectorNx": "nodes_x", "vectorNy": "nodes_y", "vectorNz": "nodes_z",
    "vol": "cell_volumes",
}
_REMOVED_INDEXED = {       # base/base_regular_mesh.py / base_tensor_mesh.py: removed per-axis shorthands
    **{"h" + c: f"mesh.h[{i}]" for i, c in enumerate("xyz")},
    **{"nC" + c: f"mesh.shape_cells[{i}]" for i, c in enumerate("xyz")},
    **{"nN" + c: f"mesh.shape_nodes[{i}]" for i, c in enumerate("xyz")},
}


# removed METHODS: the attribute exists, calling it raises (utils/code_utils.py: deprecate_method(..., error=True))
_REMOVED_METHODS = {"readUBC": "read_UBC", "readModelUBC": "read_model_UBC", "writeUBC": "write_UBC",
                    "writeModelUBC": "write_model_UBC", "r": "reshape"}


def removed_method(cls_name, name):
    """A callable that raises like the reference's removed methods, or None."""
    if name not in _REMOVED_METHODS:
        return None

    def removed(*args, **kwargs):
        raise NotImplementedError(f"{cls_name}.{name} has been removed, please use {cls_name}.{_REMOVED_METHODS[name]}.")

    return removed


def removed_name(cls_name, name, table=_REMOVED):
    """NotImplementedError for a name the reference has removed, or None."""
    if name in table:
        return NotImplementedError(f"{cls_name}.{name} has been removed, please use {cls_name}.{table[name]}.")
    if name in _REMOVED_INDEXED:
        return NotImplementedError(f"The {name} property is removed, please access as {_REMOVED_INDEXED[name]}. "
                                   "This message will be removed in discretize 1.0.0.")
    if name in ("axis_u", "axis_v", "axis_w"):      # base_regular_mesh.py: the first message carries the reference's typo
        word = "rmoved" if name == "axis_u" else "removed"
        return NotImplementedError(f"The {name} property is {word}, please access as self.orientation[{'uvw'.index(name[-1])}]. "
                                   "This will be removed in discretize 1.0.0.")
    return None


def _check_reference_system(value):
    """reference: base/base_regular_mesh.py:238-262 (names and abbreviations)."""
    if value is None:
        return "cartesian"
    choices = {"cartesian": "cartesian", "car": "cartesian", "cartesian": "cartesian", "cylindrical": "cylindrical",
               "cyl": "cylindrical", "spherical": "spherical", "sph": "spherical"}
    if not isinstance(value, str) or value.lower() not in choices:
        raise ValueError("Coordinate system ({}) unknown.".format(value))
    return choices[value.lower()]


def _check_orientation(value, dim):
    """Only the identity orientation is supported: the kernels address an axis-aligned grid."""
    if value is None:
        return
    value = np.asarray(value, dtype=np.float64)
    if value.shape != (dim, dim):
        raise ValueError(f"orientation must have shape ({dim}, {dim}), got {value.shape}")
    if not np.allclose(value, np.eye(dim)):
        raise NotImplementedError("discretize_b200: rotated meshes (orientation != identity) are not supported")


class _MeshPersistence:
    """to_dict / serialize / deserialize / save / copy / equals of the reference (base/base_mesh.py:50-197), shared by
    TensorMesh and TreeMesh.  The dictionaries (and the JSON written by ``save``) are the reference's, so meshes saved by
    either library load in the other (``utils.load_mesh``)."""

    _items = ()
    _ref_module = "discretize"

    orientation = property(lambda self: np.eye(self.dim))
    rotation_matrix = property(lambda self: np.eye(self.dim))
    reference_is_rotated = property(lambda self: False)
    reference_system = property(lambda self: getattr(self, "_reference_system", "cartesian"))

    def to_dict(self):
        out = {"__module__": type(self).__module__, "__class__": type(self).__name__}
        for item in self._items:
            attr = getattr(self, item, None)
            if attr is None:
                continue
            if isinstance(attr, np.ndarray):
                attr = attr.tolist()
            elif isinstance(attr, tuple):
                attr = [a.tolist() if isinstance(a, np.ndarray) else a for a in attr]
            out[item] = attr
        return out

    serialize = to_dict

    @classmethod
    def deserialize(cls, items, **kwargs):
        items = dict(items)
        items.pop("__module__", None)
        items.pop("__class__", None)
        return cls(**items)

    def save(self, file_name="mesh.json", verbose=False, **kwargs):
        import json
        import os

        if "filename" in kwargs:
            raise TypeError("The filename keyword argument has been removed, please use file_name. "
                            "This will be removed in discretize 1.0.0")
        f = os.path.abspath(file_name)
        with open(f, "w") as outfile:
            json.dump(self.to_dict(), outfile)
        if verbose:
            print("Saved {}".format(f))
        return f

    def copy(self):
        return type(self).deserialize(self.to_dict())

    def validate(self):
        return True

    def equals(self, other_mesh):
        if type(self) is not type(other_mesh):
            return False
        a, b = self.to_dict(), other_mesh.to_dict()
        if a.keys() != b.keys():
            return False
        for key in a:
            try:
                same = np.array_equal(np.asarray(a[key], dtype=object if key == "cell_state" else None),
                                      np.asarray(b[key], dtype=object if key == "cell_state" else None)) \
                    if not isinstance(a[key], (str, dict)) else a[key] == b[key]
            except Exception:
                same = False
            if key == "h":
                same = len(a[key]) == len(b[key]) and all(np.array_equal(x, y) for x, y in zip(a[key], b[key]))
            if not same:
                return False
        return True


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


class TensorMesh(_MeshPersistence):
    """Tensor-product mesh with CUDA operators (reference: discretize/tensor_mesh.py)."""

    _items = ("h", "origin", "shape_cells", "orientation", "reference_system")

    _meshType = "TENSOR"

    def __init__(self, h, origin=None, **kwargs):
        # reference: base/base_tensor_mesh.py:86-114
        if "x0" in kwargs:
            origin = kwargs.pop("x0")
        kwargs.pop("shape_cells", None)                 # written by to_dict; implied by h (base_tensor_mesh.py:110-111)
        orientation = kwargs.pop("orientation", None)
        self._reference_system = _check_reference_system(kwargs.pop("reference_system", None))
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
        _check_orientation(orientation, len(h))

    # ---- cells as objects (tensor_mesh.py:160-310; tensor_cell.py) -----------------------------------------------------
    def __len__(self):
        return self.n_cells

    def __bool__(self):
        return True          # a mesh is not "empty"; len() of an unfinalized tree must not be consulted for truth

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def _sanitize_indices(self, indices, dim=None):
        if isinstance(indices, tuple):
            return tuple(i if i >= 0 else i + self.shape_cells[a] for a, i in enumerate(indices))
        if indices < 0:
            return indices + (self.n_cells if dim is None else self.shape_cells[dim])
        return indices

    def _get_cell(self, indices):
        from .tensor_cell import TensorCell

        assert all(index >= 0 for index in indices)
        nodes = [self._nodes_1d(a) for a in range(self.dim)]
        origin = np.array([nodes[a][i] for a, i in enumerate(indices)])
        h = np.array([nodes[a][i + 1] - nodes[a][i] for a, i in enumerate(indices)])
        return TensorCell(h, origin, tuple(indices), self.shape_cells)

    def __getitem__(self, indices):
        """A TensorCell for an int (ravelled) or a tuple of ints; a list of cells when slices are involved."""
        import itertools

        def span(sl, end):
            start = 0 if sl.start is None else sl.start
            stop = end if sl.stop is None else sl.stop
            step = 1 if sl.step is None else sl.step
            start, stop = (start + end if start < 0 else start), (stop + end if stop < 0 else stop)
            return reversed(range(start, stop, abs(step))) if step < 0 else range(start, stop, step)

        if isinstance(indices, slice):
            return [self[i] for i in span(indices, len(self))]
        if np.issubdtype(type(indices), np.integer):
            indices = np.unravel_index(self._sanitize_indices(indices), self.shape_cells, order="F")
        if not isinstance(indices, tuple):
            raise ValueError(f"Invalid indices '{indices}'. It should be an int, a slice or a tuple of int and slices.")
        if len(indices) != self.dim:
            raise ValueError(f"Invalid number of indices '{len(indices)}'. "
                             f"It should match the number of dimensions of the mesh ({self.dim}).")
        if all(np.issubdtype(type(i), np.integer) for i in indices):
            return self._get_cell(self._sanitize_indices(tuple(indices)))
        per_axis = [list(span(i, self.shape_cells[a])) if isinstance(i, slice) else [self._sanitize_indices(i, dim=a)]
                    for a, i in enumerate(indices)]
        cells = [self._get_cell(i[::-1]) for i in itertools.product(*per_axis[::-1])]      # first axis fastest
        return cells if cells else None

    def reshape(self, x, x_type="cell_centers", out_type="cell_centers", return_format="V", **kwargs):
        """Reshape a discrete quantity between vectors and tensor-shaped arrays, or pick a component of a face / edge
        vector (base/base_regular_mesh.py:817-1001).  Same argument checks and messages as the reference; like there, a
        list input is refused (its element check tests the list itself)."""
        if "xType" in kwargs:
            raise TypeError("The xType keyword argument has been removed, please use x_type. "
                            "This will be removed in discretize 1.0.0")
        if "outType" in kwargs:
            raise TypeError("The outType keyword argument has been removed, please use out_type. "
                            "This will be removed in discretize 1.0.0")
        if "format" in kwargs:
            raise TypeError("The format keyword argument has been removed, please use return_format. "
                            "This will be removed in discretize 1.0.0")
        x_type = self._parse_location_type(x_type)
        out_type = self._parse_location_type(out_type)
        allowed = ["cell_centers", "nodes", "faces", "faces_x", "faces_y", "faces_z", "edges", "edges_x", "edges_y", "edges_z"]
        if not isinstance(x, (list, np.ndarray)):
            raise TypeError("x must be either a list or a ndarray")
        if x_type not in allowed:
            raise ValueError("x_type must be either '" + "', '".join(allowed) + "'")
        if out_type not in allowed:
            raise ValueError("out_type must be either '" + "', '".join(allowed) + "'")
        if return_format not in ["M", "V"]:
            raise ValueError("return_format must be either 'M' or 'V'")
        if out_type[: len(x_type)] != x_type:
            raise ValueError("You cannot change types when reshaping.")
        if x_type not in out_type:
            raise ValueError("You cannot change type of components.")
        if isinstance(x, list):
            raise TypeError("x[0] must be a numpy array")
        x = x[:]
        d = self.dim
        family = "faces" if x_type.startswith("faces") else "edges"
        shapes = [getattr(self, f"shape_{family}_{c}", None) if i < d else None for i, c in enumerate("xyz")]

        def shaped(xx, nn):
            return xx.reshape(nn, order="F") if return_format == "M" else xx.reshape(-1, order="F")

        def one(xx, out_type):
            if x_type in ("cell_centers", "nodes"):
                nn = self.shape_cells if x_type == "cell_centers" else self.shape_nodes
                if xx.size != np.prod(nn):
                    raise ValueError("Number of elements must not change.")
                return shaped(xx, nn)
            if x_type in ("faces", "edges"):          # a component of a full face / edge vector
                xx = xx.reshape(-1, order="F")
                counts = (self.n_faces_per_direction if x_type == "faces" else self.n_edges_per_direction)[:d]
                off = np.r_[0, np.cumsum(counts)]
                for a, c in enumerate("xyz"):
                    if c in out_type:
                        if d <= a:
                            raise ValueError("Dimensions of mesh not great enough for {}_{}".format(x_type, c))
                        if xx.size != off[-1]:
                            raise ValueError("Vector is not the right size.")
                        return shaped(xx[off[a]:off[a + 1]], shapes[a])
                return None
            a = 0 if "x" in x_type else 1 if "y" in x_type else 2     # faces_x ... edges_z: one component on its own
            nn = shapes[a]
            if xx.size != np.prod(nn):
                raise ValueError(f"Vector is not the right size. Expected {np.prod(nn)}, got {xx.size}")
            return shaped(xx, nn)

        is_vector_quantity = x.ndim == 2 and x.shape[1] == d
        if out_type in ("faces", "edges"):
            if is_vector_quantity:
                raise ValueError("Not sure what to do with a vector vector quantity..")
            return tuple(one(x, out_type + "_" + c) for c in "xyz"[:d])
        if is_vector_quantity:
            return tuple(one(x[:, k], out_type) for k in range(x.shape[1]))
        return one(x, out_type)

    # ---- UBC-GIF files (mixins/mesh_io.py:15-420; ubc_io.py) ---------------------------------------------------------
    @classmethod
    def read_UBC(cls, file_name, directory=None):
        from . import ubc_io

        return ubc_io.read_tensor_mesh(cls, file_name, directory)

    def write_UBC(self, file_name, models=None, directory=None, comment_lines=""):
        from . import ubc_io

        ubc_io.write_tensor_mesh(self, file_name, models, directory, comment_lines)

    def read_model_UBC(self, file_name, directory=None):
        from . import ubc_io

        return ubc_io.read_tensor_model(self, file_name, directory)

    def write_model_UBC(self, file_name, model, directory=None):
        from . import ubc_io

        ubc_io.write_tensor_model(self, file_name, model, directory)

    def __reduce__(self):
        # pickle / copy rebuild from the defining data; cached operators and device arrays are not carried along
        return type(self), (self.h, self.origin)

    def __getattr__(self, name):
        # alias table, like BaseMesh.__getattr__ (base/base_mesh.py:24-48)
        if name in _ALIASES:
            return getattr(self, _ALIASES[name])
        err = removed_name(type(self).__name__, name)
        if err is not None:
            raise err
        method = removed_method(type(self).__name__, name)
        if method is not None:
            return method
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
        """Short names -> long names by their first letter; anything else passes through (base/base_mesh.py:4182-4203)."""
        if len(location_type) == 0:
            return location_type
        first = location_type[0]
        if first == "F":
            return "faces_" + location_type[-1] if len(location_type) > 1 else "faces"
        if first == "E":
            return "edges_" + location_type[-1] if len(location_type) > 1 else "edges"
        if first == "N":
            return "nodes"
        if first == "C":
            return "cell_centers_" + location_type[-1] if len(location_type) > 2 else "cell_centers"
        return location_type

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
        if self.dim == 1:  # the reference returns flat arrays for 1-D meshes (base_tensor_mesh.py:735-741)
            return np.asarray(t[0], dtype=np.float64).copy()
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

    # ---- small host-side conveniences of the reference (base_regular_mesh.py:130-330, base_tensor_mesh.py:311-356,
    # tensor_mesh.py:312-517, 518-760) ----------------------------------------------------------------------

    def _split(self, vec, counts, a, what="faces"):
        if a >= self.dim:   # tensor_mesh.py:366-367, 398-399, 466-467, 499-500 (the z-edge message says "y-edges" there too)
            label = "xyz"[a] if what == "faces" else "y"
            raise Exception("{}D meshes do not have {}-{}".format(self.dim, label, what.capitalize() if what == "faces" else what))
        o = np.r_[0, np.cumsum(counts)]
        return vec[o[a]:o[a + 1]]

    face_x_areas = property(lambda self: self._split(self.face_areas, self.n_faces_per_direction, 0))
    face_y_areas = property(lambda self: self._split(self.face_areas, self.n_faces_per_direction, 1))
    face_z_areas = property(lambda self: self._split(self.face_areas, self.n_faces_per_direction, 2))
    edge_x_lengths = property(lambda self: self._split(self.edge_lengths, self.n_edges_per_direction, 0, "edges"))
    edge_y_lengths = property(lambda self: self._split(self.edge_lengths, self.n_edges_per_direction, 1, "edges"))
    edge_z_lengths = property(lambda self: self._split(self.edge_lengths, self.n_edges_per_direction, 2, "edges"))

    @property
    def h_gridded(self):
        """(n_cells, dim) widths of every cell."""
        if self.dim == 1:
            return self.h[0][:, None]
        grids = np.meshgrid(*self.h, indexing="ij")
        return np.stack([g.reshape(-1, order="F") for g in grids], axis=1)

    @property
    def faces(self):
        if self.dim == 1:
            return self._grid("faces_x")
        return np.concatenate([self._grid("faces_" + "xyz"[a]) for a in range(self.dim)])

    @property
    def edges(self):
        if self.dim == 1:
            return self._grid("edges_x")
        return np.concatenate([self._grid("edges_" + "xyz"[a]) for a in range(self.dim)])

    def _unit_vectors(self, counts):
        if self.dim == 1:
            return None  # the reference defines these for 2-D and 3-D only (base_rectangular_mesh.py:588-660)
        out = np.zeros((int(sum(counts)), self.dim))
        o = np.r_[0, np.cumsum(counts)]
        for a in range(self.dim):
            out[o[a]:o[a + 1], a] = 1.0
        return out

    face_normals = property(lambda self: self._unit_vectors(self.n_faces_per_direction))
    edge_tangents = property(lambda self: self._unit_vectors(self.n_edges_per_direction))

    @property
    def cell_boundary_indices(self):
        """Boolean masks of the cells touching the -x, +x, (-y, +y, (-z, +z)) boundary (tensor_mesh.py:677-760)."""
        idx = np.indices(self.shape_cells)
        out = []
        for a in range(self.dim):
            ia = idx[a].reshape(-1, order="F")
            out += [ia == 0, ia == self.shape_cells[a] - 1]
        return tuple(out)

    @property
    def face_boundary_indices(self):
        """Boolean masks over the faces of each direction that lie on the lower / upper boundary (tensor_mesh.py:518-594)."""
        out = []
        for a in range(self.dim):
            shp = self._shape((a,))
            ia = np.indices(shp)[a].reshape(-1, order="F")
            out += [ia == 0, ia == shp[a] - 1]
        return tuple(out)

    @property
    def cell_bounds(self):
        """(n_cells, 2 dim): x1, x2, (y1, y2, (z1, z2)) of every cell (tensor_mesh.py:597-614)."""
        cols = []
        for a in range(self.dim):
            x = self.get_tensor("nodes")[a]
            for side in (x[:-1], x[1:]):
                grids = np.meshgrid(*[side if b == a else np.zeros(self.shape_cells[b]) for b in range(self.dim)], indexing="ij")
                cols.append(grids[a].reshape(-1, order="F"))
        return np.stack(cols, axis=1)

    @property
    def cell_nodes(self):
        """(n_cells, 2**dim) node numbers of every cell, x fastest within a cell (tensor_mesh.py:617-674)."""
        idx = np.arange(self.n_nodes).reshape(self.shape_nodes, order="F")
        corners = []
        for c in range(2 ** self.dim):
            sl = tuple(slice(1, None) if (c >> a) & 1 else slice(None, -1) for a in range(self.dim))
            corners.append(idx[sl].reshape(-1, order="F"))
        return np.stack(corners, axis=-1)

    def point2index(self, locs):
        """Index of the cell containing (or nearest to) each point (tensor_mesh.py:759-784)."""
        locs = as_array_n_by_dim(locs, self.dim)
        tensors = self.get_tensor("nodes")
        multi = tuple(np.clip(np.searchsorted(n, p) - 1, 0, len(n) - 2) for n, p in zip(tensors, locs.T))
        if self.dim == 1:
            return multi[0]
        return np.ravel_multi_index(multi, self.shape_cells, order="F")

    def closest_points_index(self, locations, grid_loc="CC", discard=False):
        """Index of the closest grid location to each point (base/base_mesh.py:3999-4064)."""
        from scipy.spatial import KDTree

        locations = as_array_n_by_dim(locations, self.dim)
        grid_loc = self._parse_location_type(grid_loc)
        key = f"_{grid_loc}_tree"
        tree = self._cache.get(key)
        if tree is None:
            tree = KDTree(as_array_n_by_dim(getattr(self, grid_loc), self.dim))
        if not discard:
            self._cache[key] = tree
        return tree.query(locations)[1]

    def project_face_vector(self, face_vectors):
        """Dot product of one vector per face with the face normals (base/base_mesh.py:627-655)."""
        if not isinstance(face_vectors, np.ndarray):
            raise Exception("face_vectors must be an ndarray")
        if not (face_vectors.ndim == 2 and face_vectors.shape == (self.n_faces, self.dim)):
            raise Exception("face_vectors must be an ndarray of shape (n_faces, dim)")
        return np.sum(face_vectors * self.face_normals, 1)

    def project_edge_vector(self, edge_vectors):
        """Dot product of one vector per edge with the edge tangents (base/base_mesh.py:657-685)."""
        if not isinstance(edge_vectors, np.ndarray):
            raise Exception("edge_vectors must be an ndarray")
        if not (edge_vectors.ndim == 2 and edge_vectors.shape == (self.n_edges, self.dim)):
            raise Exception("edge_vectors must be an ndarray of shape (nE, dim)")
        return np.sum(edge_vectors * self.edge_tangents, 1)

    # ---- host matrices for the host-assembled operators (boundary terms, surface / line inner products) ------------
    def _host_matrix(self, name):
        """scipy CSR of ``self.<name>`` assembled on the host (any dimension), cached; no GPU involved."""
        key = ("host", name)
        if key not in self._cache:
            from . import lowdim

            self._cache[key] = lowdim.build(self, name)
        return self._cache[key]

    def _boundary_indices(self):
        face, edge, node = self._boundary_masks()
        return np.nonzero(face)[0], (None if edge is None else np.nonzero(edge)[0]), np.nonzero(node)[0]

    def _boundary_face_rows(self, name):
        from . import lowdim

        return lowdim.boundary_face_rows(self, name)

    # ---- weak-form boundary pieces (boundary_ops.py) ------------------------------------------------------------
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

    def get_BC_projections(self, BC, discretization="CC"):
        from . import boundary_ops

        return boundary_ops.get_BC_projections(self, BC, discretization)

    def get_BC_projections_simple(self, discretization="CC"):
        from . import boundary_ops

        return boundary_ops.get_BC_projections_simple(self, discretization)

    # ---- face / edge hosted properties (surface_products.py) -------------------------------------------------------
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
            if self.dim == 2 and host is not None and np.asarray(host).size == 3 * self.n_cells:
                from .csr import CsrOperator

                if invert_matrix:
                    raise Exception("Solver needed to invert A.")       # inner_products.py:216-217
                return CsrOperator(lowdim.mass_tensor_matrix(self, proj, host, invert_model))
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
            if self.dim == 2 and np.asarray(host).size == 3 * self.n_cells:
                if invert_model or invert_matrix:
                    raise NotImplementedError("inverting the property or the matrix is not yet implemented for this "
                                              "mesh/tensorType. You should write it!")

                def tensor_deriv(v=None):
                    if v is None:
                        raise Exception("v must be supplied for this implementation.")
                    vh = v.cpu().numpy() if isinstance(v, torch.Tensor) else v
                    return CsrOperator(lowdim.mass_tensor_deriv_matrix(self, proj, vh))

                return tensor_deriv
            if np.asarray(host).size not in (1, self.n_cells, self.n_cells * self.dim):
                raise NotImplementedError("unexpected model size for a %d-D mesh" % self.dim)

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

    def __str__(self):
        """One line per axis: cells, extent, smallest / largest width, largest ratio between neighbouring widths."""
        lines = [f"TensorMesh: {self.n_cells:,} cells ({' x '.join(str(n) for n in self.shape_cells)})",
                 f"  {'axis':<5}{'cells':>8}{'from':>14}{'to':>14}{'h min':>12}{'h max':>12}{'max ratio':>11}"]
        for a, name in enumerate("xyz"[:self.dim]):
            x, h = self._nodes_1d(a), self._h[a]
            ratio = 1.0 if h.size < 2 else float(np.max(np.r_[h[:-1] / h[1:], h[1:] / h[:-1]]))
            lines.append(f"  {name:<5}{h.size:>8}{x[0]:>14.6g}{x[-1]:>14.6g}{h.min():>12.6g}{h.max():>12.6g}{ratio:>11.3f}")
        return "\n".join(lines)

"""Small matrix, index and coordinate helpers of ``discretize.utils`` that user code (and SimPEG) calls next to the mesh
operators.  Reference: ``utils/matrix_utils.py`` (sdiag .. inverse_property_tensor, Zero / Identity / TensorType),
``utils/coordinate_utils.py``, ``utils/interpolation_utils.py:44-172``.  Host-side numpy / scipy; every function is checked
against the reference (tests/test_matrix_utils.py, fixture generated from it)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .tensor_mesh import as_array_n_by_dim, is_scalar


def _vec(x):
    from .mesh_utils import mkvc

    return mkvc(x)


# ---- symbolic zero / identity ------------------------------------------------------------------------------------------


class _EndlessShape(tuple):
    """Shape of an operand that fits anything: every index gives the same value."""

    def __new__(cls, val=None):
        self = super().__new__(cls)
        self._val = val
        return self

    def __getitem__(self, key):
        return _EndlessShape(self._val) if isinstance(key, slice) else self._val

    def __len__(self):
        return 0

    def __repr__(self):
        return f"({self._val}, {self._val}, ...)"


class _Shapeless:
    __array_ufunc__ = None          # numpy must hand arithmetic over to these classes
    __numpy_ufunc__ = True
    ndim = property(lambda self: None)
    shape = property(lambda self: _EndlessShape(None))
    T = property(lambda self: self)

    def transpose(self):
        return self


class Zero(_Shapeless):
    """Arithmetic with "nothing": adding it changes nothing, multiplying by it gives it back (matrix_utils.py:1462-1616)."""

    def __repr__(self):
        return "Zero"

    def __bool__(self):
        return False

    def __add__(self, v):
        return v

    __radd__ = __iadd__ = __rsub__ = __isub__ = __add__

    def __sub__(self, v):
        return -v

    def __mul__(self, v):
        return self

    __rmul__ = __matmul__ = __rmatmul__ = __div__ = __truediv__ = __getitem__ = __mul__

    def __rtruediv__(self, v):
        raise ZeroDivisionError("Cannot divide by zero.")

    __rdiv__ = __rfloordiv__ = __rtruediv__

    def __pos__(self):
        return self

    __neg__ = __pos__

    def __lt__(self, v):
        return 0 < v

    def __le__(self, v):
        return 0 <= v

    def __eq__(self, v):
        return v == 0

    def __ne__(self, v):
        return not (0 == v)

    def __ge__(self, v):
        return 0 >= v

    def __gt__(self, v):
        return 0 > v

    __hash__ = None


class Identity(_Shapeless):
    """Arithmetic with +-1 that becomes a sparse identity when added to a sparse matrix (matrix_utils.py:1619-1780)."""

    def __init__(self, positive=True):
        self._positive = positive

    _one = property(lambda self: 1 if self._positive else -1)

    def __repr__(self):
        return "I" if self._positive else "-I"

    def __bool__(self):
        return True

    def __pos__(self):
        return self

    def __neg__(self):
        return Identity(not self._positive)

    def __add__(self, v):
        if sp.issparse(v):
            eye = sp.identity(v.shape[0], format="csr")
            return v + eye if self._positive else v - eye
        return v + 1 if self._positive else v - 1

    __radd__ = __add__

    def __sub__(self, v):
        return self + -v

    def __rsub__(self, v):
        return -self + v

    def __mul__(self, v):
        return v if self._positive else -v

    __rmul__ = __matmul__ = __rmatmul__ = __rdiv__ = __rtruediv__ = __mul__

    def __truediv__(self, v):
        if sp.issparse(v):
            raise NotImplementedError("Sparse arrays not divisibile.")
        return 1.0 / v if self._positive else -1.0 / v

    __div__ = __truediv__

    def __floordiv__(self, v):
        return self._one // v

    def __rfloordiv__(self, v):
        return v // self._one

    def __lt__(self, v):
        return self._one < v

    def __le__(self, v):
        return self._one <= v

    def __eq__(self, v):
        return v == self._one

    def __ne__(self, v):
        return not (self._one == v)

    def __ge__(self, v):
        return self._one >= v

    def __gt__(self, v):
        return self._one > v

    __hash__ = None


# ---- sparse building blocks ----------------------------------------------------------------------------------------------


def sdiag(v):
    """Sparse diagonal matrix of a vector (matrix_utils.py:85-114)."""
    if isinstance(v, Zero):
        return Zero()
    return sp.spdiags(_vec(v), 0, v.size, v.size, format="csr")


def sdinv(M):
    """Inverse of a sparse diagonal matrix."""
    return sdiag(1.0 / M.diagonal())


def speye(n):
    return sp.identity(n, format="csr")


def spzeros(n1, n2):
    return sp.dia_matrix((n1, n2))


def kron3(A, B, C):
    """A (x) B (x) C as CSR."""
    return sp.kron(sp.kron(A, B), C, format="csr")


def ddx(n):
    """Nodes -> centres difference stencil (-1, +1), n x (n + 1)."""
    i = np.arange(n)
    return sp.csr_matrix((np.r_[-np.ones(n), np.ones(n)], (np.r_[i, i], np.r_[i, i + 1])), shape=(n, n + 1))


def av(n):
    """Nodes -> centres average (1/2, 1/2), n x (n + 1)."""
    i = np.arange(n)
    return sp.csr_matrix((np.full(2 * n, 0.5), (np.r_[i, i], np.r_[i, i + 1])), shape=(n, n + 1))


def av_extrap(n):
    """Centres -> nodes average, the two end nodes take their only neighbour with weight 1, (n + 1) x n."""
    i = np.arange(1, n)
    rows = np.r_[0, i, i, n]
    cols = np.r_[0, i - 1, i, n - 1]
    vals = np.r_[1.0, np.full(2 * (n - 1), 0.5), 1.0]
    return sp.csr_matrix((vals, (rows, cols)), shape=(n + 1, n))


def make_boundary_bool(shape, bdir="xyz", **kwargs):
    """Flattened (first axis fastest) mask of the entries on the first / last index along the axes in ``bdir``."""
    import warnings

    old = kwargs.pop("dir", None)
    if old is not None:
        warnings.warn("The `dir` keyword argument has been renamed to `bdir` to avoid shadowing the builtin variable `dir`. "
                      "This will be removed in discretize 1.0.0", FutureWarning, stacklevel=2)
        bdir = old
    mask = np.zeros(shape, dtype=bool, order="F")
    for axis, name in enumerate("xyz"[:len(shape)]):
        if name in bdir:
            sel = [slice(None)] * len(shape)
            sel[axis] = [0, -1]
            mask[tuple(sel)] = True
    return mask.reshape(-1, order="F")


def ind2sub(shape, inds):
    return np.unravel_index(inds, shape, order="F")


def sub2ind(shape, subs):
    if len(shape) == 1:
        return subs
    subs = np.atleast_2d(subs)
    if subs.shape[1] != len(shape):
        raise ValueError("Indexing must be done as a column vectors. e.g. [[3,6],[6,2],...]")
    return _vec(np.ravel_multi_index(subs.T, shape, order="F"))


def get_subarray(A, ind):
    if not isinstance(ind, list):
        raise TypeError("ind must be a list of vectors")
    if len(A.shape) != len(ind):
        raise ValueError("ind must have the same length as the dimension of A")
    if len(A.shape) == 2:
        return A[ind[0], :][:, ind[1]]
    if len(A.shape) == 3:
        return A[ind[0], :, :][:, ind[1], :][:, :, ind[2]]
    raise Exception("get_subarray does not support dimension asked.")


# ---- small dense block inverses --------------------------------------------------------------------------------------------


def _deprecated_return_matrix(kwargs, return_matrix):
    import warnings

    if "returnMatrix" in kwargs:
        warnings.warn("The returnMatrix keyword argument has been deprecated, please use return_matrix. "
                      "This will be removed in discretize 1.0.0", FutureWarning, stacklevel=3)
        return kwargs["returnMatrix"]
    return return_matrix


def inverse_3x3_block_diagonal(a11, a12, a13, a21, a22, a23, a31, a32, a33, return_matrix=True, **kwargs):
    """Element vectors (or the block matrix) of the inverses of many 3 x 3 matrices given by their element vectors:
    adjugate over determinant (matrix_utils.py:615-776)."""
    return_matrix = _deprecated_return_matrix(kwargs, return_matrix)
    a11, a12, a13, a21, a22, a23, a31, a32, a33 = (_vec(np.asarray(x)) for x in (a11, a12, a13, a21, a22, a23, a31, a32, a33))
    det = (a31 * a12 * a23 - a31 * a13 * a22 - a21 * a12 * a33 + a21 * a13 * a32 + a11 * a22 * a33 - a11 * a23 * a32)
    b11 = +(a22 * a33 - a23 * a32) / det
    b12 = -(a12 * a33 - a13 * a32) / det
    b13 = +(a12 * a23 - a13 * a22) / det
    b21 = +(a31 * a23 - a21 * a33) / det
    b22 = -(a31 * a13 - a11 * a33) / det
    b23 = +(a21 * a13 - a11 * a23) / det
    b31 = -(a31 * a22 - a21 * a32) / det
    b32 = +(a31 * a12 - a11 * a32) / det
    b33 = -(a21 * a12 - a11 * a22) / det
    if not return_matrix:
        return b11, b12, b13, b21, b22, b23, b31, b32, b33
    return sp.vstack((sp.hstack((sdiag(b11), sdiag(b12), sdiag(b13))), sp.hstack((sdiag(b21), sdiag(b22), sdiag(b23))),
                      sp.hstack((sdiag(b31), sdiag(b32), sdiag(b33)))))


def inverse_2x2_block_diagonal(a11, a12, a21, a22, return_matrix=True, **kwargs):
    return_matrix = _deprecated_return_matrix(kwargs, return_matrix)
    a11, a12, a21, a22 = (_vec(np.asarray(x)) for x in (a11, a12, a21, a22))
    det = a11 * a22 - a21 * a12
    b11, b12, b21, b22 = +a22 / det, -a12 / det, -a21 / det, +a11 / det
    if not return_matrix:
        return b11, b12, b21, b22
    return sp.vstack((sp.hstack((sdiag(b11), sdiag(b12))), sp.hstack((sdiag(b21), sdiag(b22)))))


def invert_blocks(A):
    """Inverse of the trailing 2 x 2 or 3 x 3 blocks of an array (closed form), numpy.linalg.inv beyond."""
    if A.shape[-1] != A.shape[-2]:
        raise ValueError(f"Last two dimensions are not equal, got {A.shape}")
    n = A.shape[-1]
    if n == 2:
        parts = inverse_2x2_block_diagonal(A[..., 0, 0].reshape(-1), A[..., 0, 1].reshape(-1), A[..., 1, 0].reshape(-1),
                                           A[..., 1, 1].reshape(-1), return_matrix=False)
    elif n == 3:
        parts = inverse_3x3_block_diagonal(*(A[..., i, j].reshape(-1) for i in range(3) for j in range(3)),
                                           return_matrix=False)
    else:
        return np.linalg.inv(A)
    B = np.empty_like(A)
    for k, part in enumerate(parts):
        B[..., k // n, k % n] = part.reshape(A.shape[:-2])
    return B


# ---- physical property tensors ------------------------------------------------------------------------------------------------


class TensorType:
    """-1 none, 0 scalar, 1 isotropic, 2 anisotropic, 3 full tensor -- comparable with ints (matrix_utils.py:982-1092)."""

    _names = {-1: "none", 0: "scalar", 1: "isotropic", 2: "anisotropic", 3: "tensor"}

    def __init__(self, mesh, tensor):
        nC, d = mesh.n_cells, mesh.dim
        if tensor is None:
            self._tt = -1
        elif is_scalar(tensor):
            self._tt = 0
        elif tensor.size == nC:
            self._tt = 1
        elif (d == 2 and tensor.size == nC * 2) or (d == 3 and tensor.size == nC * 3):
            self._tt = 2
        elif (d == 2 and tensor.size == nC * 3) or (d == 3 and tensor.size == nC * 6):
            self._tt = 3
        else:
            raise Exception("Unexpected shape of tensor: {}".format(tensor.shape))
        self._tts = self._names[self._tt]

    def __str__(self):
        return "TensorType[{0:d}]: {1!s}".format(self._tt, self._tts)

    def __eq__(self, v):
        return self._tt == v

    def __le__(self, v):
        return self._tt <= v

    def __ge__(self, v):
        return self._tt >= v

    def __lt__(self, v):
        return self._tt < v

    def __gt__(self, v):
        return self._tt > v

    __hash__ = None


def make_property_tensor(mesh, tensor):
    """The (dim nC x dim nC) block matrix Sigma of a scalar / isotropic / anisotropic / full-tensor cell property."""
    nC, d = mesh.n_cells, mesh.dim
    if tensor is None:
        tensor = np.ones(nC)
    if is_scalar(tensor):
        tensor = tensor * np.ones(nC)
    kind = TensorType(mesh, tensor)
    if kind == 1:
        return sp.kron(sp.identity(d), sdiag(_vec(tensor)))
    if kind == 2:
        return sdiag(_vec(tensor))
    if d == 2 and tensor.size == nC * 3:
        t = tensor.reshape((nC, 3), order="F")
        return sp.vstack((sp.hstack((sdiag(t[:, 0]), sdiag(t[:, 2]))), sp.hstack((sdiag(t[:, 2]), sdiag(t[:, 1])))))
    if d == 3 and tensor.size == nC * 6:
        t = tensor.reshape((nC, 6), order="F")
        return sp.vstack((sp.hstack((sdiag(t[:, 0]), sdiag(t[:, 3]), sdiag(t[:, 4]))),
                          sp.hstack((sdiag(t[:, 3]), sdiag(t[:, 1]), sdiag(t[:, 5]))),
                          sp.hstack((sdiag(t[:, 4]), sdiag(t[:, 5]), sdiag(t[:, 2])))))
    raise Exception("Unexpected shape of tensor")


def inverse_property_tensor(mesh, tensor, return_matrix=False, **kwargs):
    """The property that gives Sigma^-1: reciprocal for the diagonal kinds, per-cell 2 x 2 / 3 x 3 inverse otherwise."""
    return_matrix = _deprecated_return_matrix(kwargs, return_matrix)
    nC, d = mesh.n_cells, mesh.dim
    kind = TensorType(mesh, tensor)
    if is_scalar(tensor):
        T = 1.0 / tensor
    elif kind < 3:
        T = 1.0 / _vec(tensor)
    elif d == 2 and tensor.size == nC * 3:
        t = tensor.reshape((nC, 3), order="F")
        b11, b12, _, b22 = inverse_2x2_block_diagonal(t[:, 0], t[:, 2], t[:, 2], t[:, 1], return_matrix=False)
        T = np.r_[b11, b22, b12]
    elif d == 3 and tensor.size == nC * 6:
        t = tensor.reshape((nC, 6), order="F")
        b = inverse_3x3_block_diagonal(t[:, 0], t[:, 3], t[:, 4], t[:, 3], t[:, 1], t[:, 5], t[:, 4], t[:, 5], t[:, 2],
                                       return_matrix=False)
        T = np.r_[b[0], b[4], b[8], b[1], b[2], b[5]]
    else:
        raise Exception("Unexpected shape of tensor")
    return make_property_tensor(mesh, T) if return_matrix else T


def cross2d(x, y):
    """z component of the cross product of 2-D vectors (rows)."""
    x, y = as_array_n_by_dim(x, 2), as_array_n_by_dim(y, 2)
    return x[:, 0] * y[:, 1] - x[:, 1] * y[:, 0]


# ---- interpolation on a tensor grid ----------------------------------------------------------------------------------------------


def interpolation_matrix(locs, x, y=None, z=None):
    """Linear / bilinear / trilinear interpolation matrix from the tensor grid (x, y, z) to ``locs``
    (interpolation_utils.py:44-172 -> interputils_cython.pyx:42-182): per axis the bracketing samples by bisection, both
    weights 1/2 on the same index outside the grid; entries in the reference's order, duplicates summed."""
    from .lowdim import _bisect_weights

    axes = [np.asarray(a, dtype=np.float64) for a in (x, y, z) if a is not None]
    dim = len(axes)
    locs = np.asarray(locs, dtype=np.float64)
    if dim == 1:
        locs = locs.reshape(-1, 1)
    npts = locs.shape[0]
    hits = [_bisect_weights(axes[a], locs[:, a]) for a in range(dim)]
    shape = [a.size for a in axes]
    rows, cols, vals = [], [], []
    # x index fastest in the column number; entries ordered with the LAST listed axis toggling fastest per the reference:
    # 1-D (x1, x2); 2-D (x1y1, x1y2, x2y1, x2y2); 3-D (x1y1z1, x1y2z1, x2y1z1, x2y2z1, then the same with z2)
    if dim == 1:
        order = [(0,), (1,)]
    elif dim == 2:
        order = [(0, 0), (0, 1), (1, 0), (1, 1)]
    else:
        order = [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 1)]
    for pick in order:
        col = np.zeros(npts, dtype=np.int64)
        w = np.ones(npts)
        stride = 1
        for a in range(dim):
            col = col + stride * hits[a][pick[a]]
            w = w * hits[a][2 + pick[a]]
            stride *= shape[a]
        rows.append(np.arange(npts))
        cols.append(col)
        vals.append(w)
    Q = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(npts, int(np.prod(shape))))
    return Q


# ---- coordinate transforms -----------------------------------------------------------------------------------------------------------


def cylindrical_to_cartesian(grid, vec=None):
    """(r, theta, z) points -> (x, y, z); with ``vec`` the cylindrical vector components at those points instead."""
    from .mesh_utils import mkvc

    grid = np.atleast_2d(grid)
    if vec is None:
        return np.hstack([mkvc(grid[:, 0] * np.cos(grid[:, 1]), 2), mkvc(grid[:, 0] * np.sin(grid[:, 1]), 2),
                          mkvc(grid[:, 2], 2)])
    vec = np.asanyarray(vec)
    if len(vec.shape) == 1 or vec.shape[1] == 1:
        vec = vec.reshape(grid.shape, order="F")
    c, s = np.cos(grid[:, 1]), np.sin(grid[:, 1])
    parts = [vec[:, 0] * c - vec[:, 1] * s, vec[:, 0] * s + vec[:, 1] * c]
    if grid.shape[1] == 3:
        parts.append(vec[:, 2])
    return np.vstack(parts).T


def cartesian_to_cylindrical(grid, vec=None):
    from .mesh_utils import mkvc

    grid = as_array_n_by_dim(grid, 3)
    theta = np.arctan2(grid[:, 1], grid[:, 0])
    if vec is None:
        return np.c_[np.linalg.norm(grid[:, :2], axis=-1), theta, grid[:, 2]]
    vec = as_array_n_by_dim(vec, 3)
    return np.hstack([mkvc(np.cos(theta) * vec[:, 0] + np.sin(theta) * vec[:, 1], 2),
                      mkvc(-np.sin(theta) * vec[:, 0] + np.cos(theta) * vec[:, 1], 2), mkvc(vec[:, 2], 2)])


cyl2cart = cylindrical_to_cartesian
cart2cyl = cartesian_to_cylindrical


def rotation_matrix_from_normals(v0, v1, tol=1e-20):
    """Rodrigues rotation that turns direction ``v0`` into ``v1``."""
    if len(v0) != 3:
        raise ValueError("Length of n0 should be 3")
    if len(v1) != 3:
        raise ValueError("Length of n1 should be 3")
    n0 = v0 * 1.0 / np.linalg.norm(v0)
    n1 = v1 * 1.0 / np.linalg.norm(v1)
    dot = n0.dot(n1)
    axis = np.cross(n0, n1)
    if np.linalg.norm(axis) < tol:
        return np.eye(3, dtype=float)
    axis *= 1.0 / np.linalg.norm(axis)
    cos_t = dot / (np.linalg.norm(n0) * np.linalg.norm(n1))
    sin_t = np.sqrt(1.0 - dot ** 2)
    K = np.array([[0.0, -axis[2], axis[1]], [axis[2], 0.0, -axis[0]], [-axis[1], axis[0], 0.0]], dtype=float)
    return np.eye(3, dtype=float) + sin_t * K + (1.0 - cos_t) * (K.dot(K))


def rotate_points_from_normals(xyz, v0, v1, x0=np.r_[0.0, 0.0, 0.0]):
    R = rotation_matrix_from_normals(v0, v1)
    if xyz.shape[1] != 3:
        raise ValueError("Grid of xyz points should be n x 3")
    if len(x0) != 3:
        raise ValueError("x0 should have length 3")
    X0 = np.ones([xyz.shape[0], 1]) * _vec(np.asarray(x0))
    return (xyz - X0).dot(R.T) + X0

"""Geometry helpers for general hexahedral / quadrilateral cells (reference: ``utils/curvilinear_utils.py``): node-index
selection on a structured grid, tetrahedron volumes, average normals and areas of (possibly non-planar) quadrilateral
faces, and the example node grids.  Pure numpy; no mesh class here uses them, scripts written for the reference do."""
from __future__ import annotations

import numpy as np

from .mesh_utils import ndgrid

_CORNER = {"A": (0, 0, 0), "B": (0, 1, 0), "C": (1, 1, 0), "D": (1, 0, 0),
           "E": (0, 0, 1), "F": (0, 1, 1), "G": (1, 1, 1), "H": (1, 0, 1)}


def example_curvilinear_grid(nC, exType):
    """Node coordinate arrays (one per axis) of a unit box grid, optionally twisted ("rotate") or inflated to a ball
    ("sphere")."""
    if not isinstance(nC, list):
        raise TypeError("nC must be a list containing the number of nodes")
    if len(nC) != 2 and len(nC) != 3:
        raise ValueError("nC must either two or three dimensions")
    if exType not in ("rect", "rotate", "sphere"):
        raise TypeError("Not a possible example type.")
    axes = [np.cumsum(np.r_[0, np.ones(n) / n]) for n in nC]
    if exType == "rect":
        return list(ndgrid(axes, vector=False))
    if exType == "sphere":
        nodes = 2 * np.stack(ndgrid([a - 0.5 for a in axes], vector=False), axis=-1)
        r_inf = np.linalg.norm(nodes, ord=np.inf, axis=-1)
        r_2 = np.linalg.norm(nodes, axis=-1)
        r_inf[r_inf == 0.0] = 1.0
        r_2[r_2 == 0.0] = 1.0
        nodes = nodes * (r_inf / r_2)[..., None]
        return [nodes[..., a] for a in range(len(nC))]
    grids = ndgrid(axes, vector=False)
    amount = 0.5 - np.sqrt(sum((g - 0.5) ** 2 for g in grids))
    amount[amount < 0] = 0
    if len(nC) == 2:
        X, Y = grids
        return [X + (-(Y - 0.5)) * amount, Y + (+(X - 0.5)) * amount]
    X, Y, Z = grids
    return [X + (-(Y - 0.5)) * amount, Y + (-(Z - 0.5)) * amount, Z + (-(X - 0.5)) * amount]


def volume_tetrahedron(xyz, A, B, C, D):
    """|det[A - D, B - D, C - D]| / 6 for index arrays A .. D into the points ``xyz``."""
    a, b, c = xyz[A, :] - xyz[D, :], xyz[B, :] - xyz[D, :], xyz[C, :] - xyz[D, :]
    det = ((b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]) * a[:, 2] - (b[:, 0] * c[:, 2] - b[:, 2] * c[:, 0]) * a[:, 1]
           + (b[:, 1] * c[:, 2] - b[:, 2] * c[:, 1]) * a[:, 0])
    return np.abs(det / 6)


def index_cube(nodes, grid_shape, n=None):
    """Node numbers of the named corners ('A'..'D' in 2-D, 'A'..'H' in 3-D) of every cell of a structured node grid."""
    if not isinstance(nodes, str):
        raise TypeError("Nodes must be a str variable: e.g. 'ABCD'")
    nodes = nodes.upper()
    try:
        dim = len(grid_shape)
        if n is None:
            n = tuple(x - 1 for x in grid_shape)
    except TypeError:
        return TypeError("grid_shape must be iterable")
    allowed = "ABCD" if dim == 2 else "ABCDEFGH"
    for node in nodes:
        if node not in allowed:
            raise ValueError("Nodes must be chosen from: '{0!s}'".format(allowed))
    if dim not in (2, 3):
        raise Exception("Only 2 and 3 dimensions supported.")
    cells = ndgrid(*[np.arange(k) for k in n[:dim]])
    out = ()
    for node in nodes:
        shifted = cells + np.array(_CORNER[node][:dim])
        out += (np.ravel_multi_index(shifted.T, grid_shape, order="F").reshape(-1),)
    return out


def face_info(xyz, A, B, C, D, average=True, normalize_normals=True, **kwargs):
    """Normal(s) and area of the quadrilaterals A-B-C-D: the four corner normals (cross products of the adjacent edges),
    averaged and normalised, or returned one per corner; the area is the mean corner parallelogram."""
    if "normalizeNormals" in kwargs:
        raise TypeError("The normalizeNormals keyword argument has been removed, please use normalize_normals. "
                        "This will be removed in discretize 1.0.0")
    if not isinstance(average, bool):
        raise TypeError("average must be a boolean")
    if not isinstance(normalize_normals, bool):
        raise TypeError("normalize_normals must be a boolean")
    AB, BC, CD, DA = xyz[B, :] - xyz[A, :], xyz[C, :] - xyz[B, :], xyz[D, :] - xyz[C, :], xyz[A, :] - xyz[D, :]
    corner_normals = [np.cross(AB, DA), np.cross(BC, AB), np.cross(CD, BC), np.cross(DA, CD)]

    def length(v):
        return np.sqrt(v[:, 0] ** 2 + v[:, 1] ** 2 + v[:, 2] ** 2)

    def unit(v):
        return v / length(v)[:, None]

    if average:
        N = unit((corner_normals[0] + corner_normals[1] + corner_normals[2] + corner_normals[3]) / 4)
    else:
        N = [unit(v) for v in corner_normals] if normalize_normals else corner_normals
    area = (length(corner_normals[0]) + length(corner_normals[1]) + length(corner_normals[2]) + length(corner_normals[3])) / 4
    return N, area

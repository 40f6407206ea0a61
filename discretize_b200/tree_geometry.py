"""Cell-versus-primitive intersection tests for TreeMesh refinement and cell queries, vectorised over (primitive, cell) pairs.

Reference: ``_extensions/geom.cpp`` -- ``Ball`` (:35-44), ``Line`` (:60-91), ``Box`` (:102-112), ``Plane`` (:124-136),
``Triangle`` (:167-300), ``VerticalTriangularPrism`` (:312-450), ``Tetrahedron`` (:495-565); used by
``Cell::refine_geom`` / ``find_cells_geom`` (``tree.h:149-182``).  The tests are the classic ones -- closest point for a
ball, slab clipping for a segment, centre / radius projection for a plane, and separating-axis tests (box axes, edge x
box-axis cross products, face normals) for the polytopes -- restated here on whole arrays: every function returns a
``predicate(req, a, b)`` that takes the index of the primitive for each pair and the low / high corners ``a, b``
``(n_pairs, dim)`` of the cells, and answers with a boolean mask.  Operand order inside each test follows the reference so
that cells touching a primitive exactly fall on the same side.
"""
from __future__ import annotations

import numpy as np


def _center_half(a, b):
    center = 0.5 * (b + a)
    return center, center - a


def ball(centers, radii):
    centers = np.asarray(centers, dtype=np.float64)
    rsq = np.asarray(radii, dtype=np.float64) ** 2

    def predicate(req, a, b):
        x0 = centers[req]
        dx = np.maximum(a, np.minimum(x0, b)) - x0
        return (dx * dx).sum(axis=1) < rsq[req]

    return predicate


def box(x0s, x1s):
    lo, hi = np.minimum(x0s, x1s), np.maximum(x0s, x1s)

    def predicate(req, a, b):
        return ((hi[req] >= a) & (lo[req] <= b)).all(axis=1)

    return predicate


def line(x0s, x1s):
    """Segments x0 -> x1: clip the parameter range [0, 1] against the slab of every axis."""
    x0s, x1s = np.asarray(x0s, dtype=np.float64), np.asarray(x1s, dtype=np.float64)
    with np.errstate(divide="ignore"):
        inv = 1.0 / (x1s - x0s)

    def predicate(req, a, b):
        x0, x1, inv_dx = x0s[req], x1s[req], inv[req]
        ok = np.ones(req.size, dtype=bool)
        t_near = np.full(req.size, -np.inf)
        t_far = np.full(req.size, np.inf)
        for i in range(a.shape[1]):
            flat = x0[:, i] == x1[:, i]
            ok &= ~(flat & ((x0[:, i] < a[:, i]) | (x0[:, i] > b[:, i])))
            ok &= ~(np.maximum(x0[:, i], x1[:, i]) < a[:, i])
            ok &= ~(np.minimum(x0[:, i], x1[:, i]) > b[:, i])
            with np.errstate(invalid="ignore"):
                t0 = (a[:, i] - x0[:, i]) * inv_dx[:, i]
                t1 = (b[:, i] - x0[:, i]) * inv_dx[:, i]
            lo, hi = np.minimum(t0, t1), np.maximum(t0, t1)
            t_near = np.where(flat, t_near, np.maximum(t_near, lo))
            t_far = np.where(flat, t_far, np.minimum(t_far, hi))
            ok &= flat | ~((t_near > t_far) | (t_far < 0) | (t_near > 1))
        return ok

    return predicate


def plane(origins, normals):
    origins, normals = np.asarray(origins, dtype=np.float64), np.asarray(normals, dtype=np.float64)

    def predicate(req, a, b):
        center, half = _center_half(a, b)
        n = normals[req]
        r = np.zeros(req.size)
        s = np.zeros(req.size)
        for i in range(a.shape[1]):
            r = r + half[:, i] * np.abs(n[:, i])
            s = s + n[:, i] * (center[:, i] - origins[req, i])
        return np.abs(s) <= r

    return predicate


# cross products of an edge e with the box axes: (name, projection of a vertex v, radius from the box half widths)
def _cross_z(e, v):
    return e[:, 1] * v[:, 0] - e[:, 0] * v[:, 1]


def _cross_x(e, v):
    return e[:, 2] * v[:, 1] - e[:, 1] * v[:, 2]


def _cross_y(e, v):
    return -e[:, 2] * v[:, 0] + e[:, 0] * v[:, 2]


def _rad_z(e, half):
    return np.abs(e[:, 1]) * half[:, 0] + np.abs(e[:, 0]) * half[:, 1]


def _rad_x(e, half):
    return np.abs(e[:, 2]) * half[:, 1] + np.abs(e[:, 1]) * half[:, 2]


def _rad_y(e, half):
    return np.abs(e[:, 2]) * half[:, 0] + np.abs(e[:, 0]) * half[:, 2]


def _separated(projections, rad):
    pmin = projections[0]
    pmax = projections[0]
    for p in projections[1:]:
        pmin = np.minimum(pmin, p)
        pmax = np.maximum(pmax, p)
    return (pmin > rad) | (pmax < -rad)


def _bounding_box_miss(vs, half, extra_top=None):
    vmin = vs[0]
    vmax = vs[0]
    for v in vs[1:]:
        vmin = np.minimum(vmin, v)
        vmax = np.maximum(vmax, v)
    if extra_top is not None:
        vmax = vmax.copy()
        vmax[:, 2] += extra_top
    return ((vmin > half) | (vmax < -half)).any(axis=1)


def triangle(tris):
    """``tris``: (n, 3, dim).  Box axes, then edge x box-axis cross products (only the two vertices whose projections
    can differ are projected, as in the reference), then -- in 3-D -- the triangle's normal."""
    tris = np.asarray(tris, dtype=np.float64)
    dim = tris.shape[2]
    e_all = (tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 1], tris[:, 2] - tris[:, 0])
    normal = np.cross(e_all[0], e_all[1]) if dim == 3 else None
    z_pairs = ((1, 2), (0, 1), (1, 2))        # vertices projected for edge 0 / 1 / 2 against z
    xy_pairs = ((0, 2), (0, 2), (0, 1))       # ... against x and y

    def predicate(req, a, b):
        center, half = _center_half(a, b)
        v = [tris[req, k] - center for k in range(3)]
        out = ~_bounding_box_miss(v, half)
        for k in range(3):
            e = e_all[k][req]
            i, j = z_pairs[k]
            out &= ~_separated([_cross_z(e, v[i]), _cross_z(e, v[j])], _rad_z(e, half))
        if dim == 3:
            for k in range(3):
                e = e_all[k][req]
                i, j = xy_pairs[k]
                out &= ~_separated([_cross_x(e, v[i]), _cross_x(e, v[j])], _rad_x(e, half))
                out &= ~_separated([_cross_y(e, v[i]), _cross_y(e, v[j])], _rad_y(e, half))
            n = normal[req]
            pmin = np.zeros(req.size)
            pmax = np.zeros(req.size)
            for i in range(3):
                pos = n[:, i] > 0
                lo = n[:, i] * (-half[:, i] - v[0][:, i])
                hi = n[:, i] * (half[:, i] - v[0][:, i])
                pmin = pmin + np.where(pos, lo, hi)
                pmax = pmax + np.where(pos, hi, lo)
            out &= ~((pmin > 0) | (pmax < 0))
        return out

    return predicate


def vertical_triangular_prism(tris, heights):
    """The triangle swept upward (+z) by ``heights``: triangle axes with the top copies of the projected vertices."""
    tris = np.asarray(tris, dtype=np.float64)
    heights = np.asarray(heights, dtype=np.float64)
    e_all = (tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 1], tris[:, 2] - tris[:, 0])
    normal = np.cross(e_all[0], e_all[1])
    z_pairs = ((1, 2), (0, 1), (1, 2))
    xy_pairs = ((0, 2), (0, 2), (0, 1))

    def predicate(req, a, b):
        center, half = _center_half(a, b)
        h = heights[req]
        v = [tris[req, k] - center for k in range(3)]
        top = []
        for vk in v:
            t = vk.copy()
            t[:, 2] = vk[:, 2] + h
            top.append(t)
        out = ~_bounding_box_miss(v, half, extra_top=h)
        for k in range(3):
            e = e_all[k][req]
            i, j = z_pairs[k]
            out &= ~_separated([_cross_z(e, v[i]), _cross_z(e, v[j])], _rad_z(e, half))
        for k in range(3):
            e = e_all[k][req]
            i, j = xy_pairs[k]
            out &= ~_separated([_cross_x(e, v[i]), _cross_x(e, top[i]), _cross_x(e, v[j]), _cross_x(e, top[j])],
                               _rad_x(e, half))
            out &= ~_separated([_cross_y(e, v[i]), _cross_y(e, top[i]), _cross_y(e, v[j]), _cross_y(e, top[j])],
                               _rad_y(e, half))
        n = normal[req]
        p0 = n[:, 0] * v[0][:, 0] + n[:, 1] * v[0][:, 1] + n[:, 2] * v[0][:, 2]
        p1 = n[:, 0] * v[0][:, 0] + n[:, 1] * v[0][:, 1] + n[:, 2] * (v[0][:, 2] + h)
        rad = np.abs(n[:, 0]) * half[:, 0] + np.abs(n[:, 1]) * half[:, 1] + np.abs(n[:, 2]) * half[:, 2]
        out &= ~_separated([p0, p1], rad)
        return out

    return predicate


def tetrahedron(tets):
    """``tets``: (n, 4, 3).  Box axes, six edges x three box axes, four face normals; all four vertices projected."""
    tets = np.asarray(tets, dtype=np.float64)
    x0, x1, x2, x3 = (tets[:, k] for k in range(4))
    edges = (x1 - x0, x2 - x0, x2 - x1, x3 - x0, x3 - x1, x3 - x2)
    normals = (np.cross(edges[0], edges[1]), np.cross(edges[0], edges[3]), np.cross(edges[1], edges[3]),
               np.cross(edges[2], edges[5]))

    def predicate(req, a, b):
        center, half = _center_half(a, b)
        v = [tets[req, k] - center for k in range(4)]
        out = ~_bounding_box_miss(v, half)
        for ek in edges:
            e = ek[req]
            out &= ~_separated([_cross_x(e, vk) for vk in v], _rad_x(e, half))
            out &= ~_separated([_cross_y(e, vk) for vk in v], _rad_y(e, half))
            out &= ~_separated([_cross_z(e, vk) for vk in v], _rad_z(e, half))
        for nk in normals:
            n = nk[req]
            proj = [n[:, 0] * vk[:, 0] + n[:, 1] * vk[:, 1] + n[:, 2] * vk[:, 2] for vk in v]
            rad = np.abs(n[:, 0]) * half[:, 0] + np.abs(n[:, 1]) * half[:, 1] + np.abs(n[:, 2]) * half[:, 2]
            out &= ~_separated(proj, rad)
        return out

    return predicate

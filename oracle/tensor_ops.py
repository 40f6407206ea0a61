"""CPU oracle for the TensorMesh operator hot path.  TEST INFRASTRUCTURE ONLY.

A numpy/scipy *restatement* (not a copy) of the algorithm the reference uses to
build its 3-D TensorMesh operators.  The reference assembles every operator as
Kronecker products of 1-D stencils scaled by geometry diagonals; this file writes
the same matrices down directly from the stencil definitions by index arithmetic
(COO triplets -> CSR), so it is an independent statement of the same math.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product path
(``discretize_b200``) never does.

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` compares every
function here with the unmodified reference (imported from /root/reference with
its extensions compiled by ``oracle/build_ref.sh``) on uniform, non-uniform and
non-cubic meshes (pattern bit-exact after canonicalisation, values <= 1e-14
relative), and ``tests/golden/*.npz`` (made by ``tests/golden/make_golden.py``
from the reference itself) pin it where the reference is absent (the GPU box).

Conventions (reference: discretize/base/base_regular_mesh.py:601-762,
discretize/utils/matrix_utils.py:77-82): x fastest (Fortran order);
faces = [Fx;Fy;Fz] with shapes (nx+1,ny,nz),(nx,ny+1,nz),(nx,ny,nz+1);
edges = [Ex;Ey;Ez] with shapes (nx,ny+1,nz+1),(nx+1,ny,nz+1),(nx+1,ny+1,nz);
nodes (nx+1,ny+1,nz+1).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class Grid:
    """Minimal 3-D tensor grid (reference: base/base_tensor_mesh.py:86-114, 155-240)."""

    def __init__(self, h, origin=None):
        hs = []
        for hi in h:
            if np.isscalar(hi):
                hi = np.ones(int(hi)) / int(hi)  # base_tensor_mesh.py:98-100
            hs.append(np.asarray(hi, dtype=np.float64).copy())
        if len(hs) != 3:
            raise NotImplementedError("oracle covers 3-D tensor meshes")
        self.h = hs
        self.origin = np.zeros(3) if origin is None else np.asarray(origin, dtype=np.float64)
        self.n = tuple(int(x.size) for x in hs)
        nx, ny, nz = self.n
        self.nC = nx * ny * nz
        self.vnF = ((nx + 1) * ny * nz, nx * (ny + 1) * nz, nx * ny * (nz + 1))
        self.vnE = (nx * (ny + 1) * (nz + 1), (nx + 1) * ny * (nz + 1), (nx + 1) * (ny + 1) * nz)
        self.nF = sum(self.vnF)
        self.nE = sum(self.vnE)
        self.nN = (nx + 1) * (ny + 1) * (nz + 1)
        # base_tensor_mesh.py:174 (nodes = cumsum([origin, h])), :239-240 (centres)
        self.nodes = [np.r_[self.origin[a], hs[a]].cumsum() for a in range(3)]
        self.centers = [(x[1:] + x[:-1]) / 2 for x in self.nodes]

    # ---- geometry vectors (tensor_mesh.py:296-517) -------------------------------
    def cell_volumes(self):
        hx, hy, hz = self.h
        a = np.outer(hx, hy).reshape(-1, order="F")
        return np.outer(a, hz).reshape(-1, order="F")

    def face_areas(self):
        hx, hy, hz = self.h
        nx, ny, nz = self.n
        ax = np.outer(np.ones(nx + 1), np.outer(hy, hz).reshape(-1, order="F")).reshape(-1, order="F")
        ay = np.outer(hx, np.outer(np.ones(ny + 1), hz).reshape(-1, order="F")).reshape(-1, order="F")
        az = np.outer(hx, np.outer(hy, np.ones(nz + 1)).reshape(-1, order="F")).reshape(-1, order="F")
        return np.r_[ax, ay, az]

    def edge_lengths(self):
        hx, hy, hz = self.h
        nx, ny, nz = self.n
        lx = np.outer(hx, np.ones((ny + 1) * (nz + 1))).reshape(-1, order="F")
        ly = np.outer(np.ones(nx + 1), np.outer(hy, np.ones(nz + 1)).reshape(-1, order="F")).reshape(-1, order="F")
        lz = np.outer(np.ones((nx + 1) * (ny + 1)), hz).reshape(-1, order="F")
        return np.r_[lx, ly, lz]

    # ---- flat index helpers -------------------------------------------------------
    def _ijk(self, shape):
        """Flattened (F-order) i, j, k index arrays of a (sx,sy,sz) block."""
        sx, sy, sz = shape
        i, j, k = np.meshgrid(np.arange(sx), np.arange(sy), np.arange(sz), indexing="ij")
        return (i.reshape(-1, order="F"), j.reshape(-1, order="F"), k.reshape(-1, order="F"))

    def shapes(self, kind):
        nx, ny, nz = self.n
        return {
            "C": (nx, ny, nz),
            "N": (nx + 1, ny + 1, nz + 1),
            "Fx": (nx + 1, ny, nz), "Fy": (nx, ny + 1, nz), "Fz": (nx, ny, nz + 1),
            "Ex": (nx, ny + 1, nz + 1), "Ey": (nx + 1, ny, nz + 1), "Ez": (nx + 1, ny + 1, nz),
        }[kind]

    def offset(self, kind):
        return {"C": 0, "N": 0, "Fx": 0, "Fy": self.vnF[0], "Fz": self.vnF[0] + self.vnF[1],
                "Ex": 0, "Ey": self.vnE[0], "Ez": self.vnE[0] + self.vnE[1]}[kind]

    def flat(self, kind, i, j, k):
        sx, sy, _ = self.shapes(kind)
        return self.offset(kind) + i + sx * (j + sy * k)


def _csr(rows, cols, vals, shape):
    m = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=shape).tocsr()
    m.sum_duplicates()
    m.sort_indices()
    return m


def canonical(m):
    """Canonical CSR used for every parity comparison (explicit zeros are kept)."""
    m = sp.csr_matrix(m).copy()
    m.sum_duplicates()
    m.sort_indices()
    return m


# ------------------------------------------------------------------------------------
# differential operators
# ------------------------------------------------------------------------------------
def face_divergence(g: Grid):
    """D = diag(1/V) . Dstencil . diag(S)  (operators/differential_operators.py:161-235)."""
    V, S = g.cell_volumes(), g.face_areas()
    i, j, k = g._ijk(g.shapes("C"))
    c = g.flat("C", i, j, k)
    rows, cols, vals = [], [], []
    for kind, d in (("Fx", (1, 0, 0)), ("Fy", (0, 1, 0)), ("Fz", (0, 0, 1))):
        lo = g.flat(kind, i, j, k)
        hi = g.flat(kind, i + d[0], j + d[1], k + d[2])
        rows += [c, c]
        cols += [lo, hi]
        vals += [(1.0 / V[c]) * -1.0 * S[lo], (1.0 / V[c]) * 1.0 * S[hi]]
    return _csr(rows, cols, vals, (g.nC, g.nF))


def nodal_gradient(g: Grid):
    """G = diag(1/L) . Gstencil (differential_operators.py:593-664)."""
    L = g.edge_lengths()
    rows, cols, vals = [], [], []
    for kind, d in (("Ex", (1, 0, 0)), ("Ey", (0, 1, 0)), ("Ez", (0, 0, 1))):
        i, j, k = g._ijk(g.shapes(kind))
        e = g.flat(kind, i, j, k)
        rows += [e, e]
        cols += [g.flat("N", i, j, k), g.flat("N", i + d[0], j + d[1], k + d[2])]
        vals += [(1.0 / L[e]) * -1.0, (1.0 / L[e]) * 1.0]
    return _csr(rows, cols, vals, (g.nE, g.nN))


def edge_curl(g: Grid):
    """C = diag(1/S) . Cstencil . diag(L) (differential_operators.py:2091-2181).

    Fx rows:  d_y(e_z) - d_z(e_y);  Fy rows: d_z(e_x) - d_x(e_z);  Fz rows: d_x(e_y) - d_y(e_x).
    """
    S, L = g.face_areas(), g.edge_lengths()
    rows, cols, vals = [], [], []

    def add(f, e, sign):
        rows.append(f)
        cols.append(e)
        vals.append((1.0 / S[f]) * sign * L[e])

    i, j, k = g._ijk(g.shapes("Fx"))
    f = g.flat("Fx", i, j, k)
    add(f, g.flat("Ez", i, j + 1, k), 1.0); add(f, g.flat("Ez", i, j, k), -1.0)
    add(f, g.flat("Ey", i, j, k + 1), -1.0); add(f, g.flat("Ey", i, j, k), 1.0)
    i, j, k = g._ijk(g.shapes("Fy"))
    f = g.flat("Fy", i, j, k)
    add(f, g.flat("Ex", i, j, k + 1), 1.0); add(f, g.flat("Ex", i, j, k), -1.0)
    add(f, g.flat("Ez", i + 1, j, k), -1.0); add(f, g.flat("Ez", i, j, k), 1.0)
    i, j, k = g._ijk(g.shapes("Fz"))
    f = g.flat("Fz", i, j, k)
    add(f, g.flat("Ey", i + 1, j, k), 1.0); add(f, g.flat("Ey", i, j, k), -1.0)
    add(f, g.flat("Ex", i, j + 1, k), -1.0); add(f, g.flat("Ex", i, j, k), 1.0)
    return _csr(rows, cols, vals, (g.nF, g.nE))


# ------------------------------------------------------------------------------------
# averaging (differential_operators.py:2478-3253; utils/matrix_utils.py:227-253)
# ------------------------------------------------------------------------------------
_FACE_D = {"Fx": (1, 0, 0), "Fy": (0, 1, 0), "Fz": (0, 0, 1)}
# the 4 edges of direction a around a cell: offsets in the two transverse axes
_EDGE_D = {
    "Ex": [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1)],
    "Ey": [(0, 0, 0), (1, 0, 0), (0, 0, 1), (1, 0, 1)],
    "Ez": [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0)],
}


def _ave_component(g: Grid, kind, weight, row_offset=0):
    i, j, k = g._ijk(g.shapes("C"))
    c = g.flat("C", i, j, k) + row_offset
    rows, cols, vals = [], [], []
    if kind[0] == "F":
        d = _FACE_D[kind]
        offs = [(0, 0, 0), d]
        w = 0.5 * weight
    elif kind[0] == "E":
        offs = _EDGE_D[kind]
        w = 0.25 * weight
    else:
        offs = [(a, b, cc) for cc in (0, 1) for b in (0, 1) for a in (0, 1)]
        w = 0.125 * weight
    for d in offs:
        rows.append(c)
        cols.append(g.flat(kind, i + d[0], j + d[1], k + d[2]))
        vals.append(np.full(c.size, w))
    return rows, cols, vals


def average_component_to_cell(g: Grid, kind):
    """aveFx2CC ... aveEz2CC: nC x (nFx | ... ) (differential_operators.py:2510-2787, 2926-3203)."""
    r, c, v = _ave_component(g, kind, 1.0)
    off = g.offset(kind)
    sx, sy, sz = g.shapes(kind)
    return _csr(r, [x - off for x in c], v, (g.nC, sx * sy * sz))


def average_face_to_cell(g: Grid):
    """(1/3) [aveFx aveFy aveFz] (differential_operators.py:2478-2492)."""
    r, c, v = [], [], []
    for kind in ("Fx", "Fy", "Fz"):
        a, b, d = _ave_component(g, kind, 1.0 / 3.0)
        r += a; c += b; v += d
    return _csr(r, c, v, (g.nC, g.nF))


def average_face_to_cell_vector(g: Grid):
    """blockdiag(aveFx, aveFy, aveFz): 3nC x nF (differential_operators.py:2494-2508)."""
    r, c, v = [], [], []
    for n, kind in enumerate(("Fx", "Fy", "Fz")):
        a, b, d = _ave_component(g, kind, 1.0, row_offset=n * g.nC)
        r += a; c += b; v += d
    return _csr(r, c, v, (3 * g.nC, g.nF))


def average_edge_to_cell(g: Grid):
    """(1/3) [aveEx aveEy aveEz] (differential_operators.py:2894-2908)."""
    r, c, v = [], [], []
    for kind in ("Ex", "Ey", "Ez"):
        a, b, d = _ave_component(g, kind, 1.0 / 3.0)
        r += a; c += b; v += d
    return _csr(r, c, v, (g.nC, g.nE))


def average_edge_to_cell_vector(g: Grid):
    """blockdiag(aveEx, aveEy, aveEz): 3nC x nE (differential_operators.py:2910-2924)."""
    r, c, v = [], [], []
    for n, kind in enumerate(("Ex", "Ey", "Ez")):
        a, b, d = _ave_component(g, kind, 1.0, row_offset=n * g.nC)
        r += a; c += b; v += d
    return _csr(r, c, v, (3 * g.nC, g.nE))


def average_node_to_cell(g: Grid):
    """mean of the 8 corner nodes (differential_operators.py:3236-3253)."""
    r, c, v = _ave_component(g, "N", 1.0)
    return _csr(r, c, v, (g.nC, g.nN))


# ------------------------------------------------------------------------------------
# mass matrices
# ------------------------------------------------------------------------------------
def _model_kind(g: Grid, model):
    """TensorType (utils/matrix_utils.py:1048-1068): -1 none, 0 scalar, 1 iso, 2 aniso, 3 tensor."""
    if model is None:
        return -1
    if np.isscalar(model) or (isinstance(model, np.ndarray) and model.ndim == 0):
        return 0
    model = np.asarray(model)
    if model.size == g.nC:
        return 1
    if model.size == 3 * g.nC:
        return 2
    if model.size == 6 * g.nC:
        return 3
    raise Exception("Unexpected shape of tensor: {}".format(model.shape))


def invert_tensor_model(g: Grid, model):
    """inverse_property_tensor (utils/matrix_utils.py:1246-1434, 3x3 closed form :734-754)."""
    kind = _model_kind(g, model)
    if kind == 0:
        return 1.0 / model
    if kind < 3:
        return 1.0 / np.asarray(model, dtype=np.float64).reshape(-1, order="F")
    t = np.asarray(model, dtype=np.float64).reshape((g.nC, 6), order="F")
    a11, a22, a33, a12, a13, a23 = (t[:, q] for q in range(6))
    a21, a31, a32 = a12, a13, a23
    det = (a11 * a22 * a33 + a12 * a23 * a31 + a13 * a21 * a32
           - a13 * a22 * a31 - a12 * a21 * a33 - a11 * a23 * a32)
    b11 = +(a22 * a33 - a23 * a32) / det
    b12 = -(a12 * a33 - a13 * a32) / det
    b13 = +(a12 * a23 - a13 * a22) / det
    b22 = +(a11 * a33 - a13 * a31) / det
    b23 = -(a11 * a23 - a13 * a21) / det
    b33 = +(a11 * a22 - a12 * a21) / det
    return np.r_[b11, b22, b33, b12, b13, b23]


def _corner_members(g: Grid, proj):
    """For each of the 8 cell corners: flat indices of the x/y/z face (edge) touching it.

    Faces: inner_products.py:247-269 + _getFacePxxx :671-738.
    Edges: inner_products.py:258-268 + _getEdgePxxx :792-863.
    Corner n = (a,b,c) in {0,1}^3, order 000,100,010,110,001,101,011,111.
    """
    i, j, k = g._ijk(g.shapes("C"))
    out = []
    for c in (0, 1):
        for b in (0, 1):
            for a in (0, 1):
                if proj == "F":
                    m = (g.flat("Fx", i + a, j, k), g.flat("Fy", i, j + b, k), g.flat("Fz", i, j, k + c))
                else:
                    m = (g.flat("Ex", i, j + b, k + c), g.flat("Ey", i + a, j, k + c), g.flat("Ez", i + a, j + b, k))
                out.append(m)
    return out


def inner_product(g: Grid, proj, model=None, invert_model=False, invert_matrix=False, do_fast=True):
    """get_face_inner_product / get_edge_inner_product (inner_products.py:35-219).

    Fast path (diagonal): base/base_tensor_mesh.py:792-863.
    Generic path: sum over the 8 corners of P^T sqrt(V/8) Sigma sqrt(V/8) P.
    """
    if proj not in ("F", "E"):
        raise TypeError("projection_type must be 'F' for faces or 'E' for edges")
    n = g.nF if proj == "F" else g.nE
    V = g.cell_volumes()
    kind = _model_kind(g, model)
    if do_fast and kind != 3:
        m = np.ones(g.nC) if model is None else model
        if invert_model:
            m = 1.0 / m
        if kind == 0:
            m = m * np.ones(g.nC)
        m = np.asarray(m, dtype=np.float64)
        if m.size == g.nC:
            Av = average_face_to_cell(g) if proj == "F" else average_edge_to_cell(g)
            diag = 3 * (Av.T @ (V * m.reshape(-1, order="F")))
        else:
            Av = average_face_to_cell_vector(g) if proj == "F" else average_edge_to_cell_vector(g)
            diag = Av.T @ (np.tile(V, 3) * m.reshape(-1, order="F"))
        if invert_matrix:
            diag = 1.0 / diag
        return sp.diags(diag, format="csr")

    # generic path
    if invert_model:
        model = invert_tensor_model(g, model)
        kind = _model_kind(g, model)
    if model is None:
        model = np.ones(g.nC)
    if kind <= 0:
        model = model * np.ones(g.nC)
    model = np.asarray(model, dtype=np.float64)
    if kind <= 1:
        t = model.reshape(-1, order="F")
        sig = {(0, 0): t, (1, 1): t, (2, 2): t}
    elif kind == 2:
        t = model.reshape((g.nC, 3), order="F")
        sig = {(0, 0): t[:, 0], (1, 1): t[:, 1], (2, 2): t[:, 2]}
    else:
        t = model.reshape((g.nC, 6), order="F")
        sig = {(0, 0): t[:, 0], (1, 1): t[:, 1], (2, 2): t[:, 2],
               (0, 1): t[:, 3], (1, 0): t[:, 3], (0, 2): t[:, 4], (2, 0): t[:, 4],
               (1, 2): t[:, 5], (2, 1): t[:, 5]}
    sq = np.sqrt(0.125 * V)
    acc = None
    for mem in _corner_members(g, proj):
        rows, cols, vals = [], [], []
        for (a, b), s in sig.items():
            rows.append(mem[a]); cols.append(mem[b]); vals.append((sq * s) * sq)
        part = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
        acc = part if acc is None else acc + part
    acc = canonical(acc)
    acc.eliminate_zeros()
    if invert_matrix:
        if kind == 3:
            raise Exception("Solver needed to invert A.")
        acc = sp.diags(1.0 / acc.diagonal(), format="csr")
    return acc


def inner_product_deriv(g: Grid, proj, model, v, invert_model=False, invert_matrix=False, do_fast=True):
    """get_{face,edge}_inner_product_deriv(model, ...)(v): n x (nC*k) sparse.

    Fast path: base/base_tensor_mesh.py:865-988.  Tensor path: inner_products.py:459-565.
    """
    n = g.nF if proj == "F" else g.nE
    V = g.cell_volumes()
    kind = _model_kind(g, model)
    v = np.asarray(v, dtype=np.float64)
    if kind <= 2 and do_fast:
        MI = None
        if invert_model or invert_matrix:
            MI = inner_product(g, proj, model, invert_model=invert_model, invert_matrix=invert_matrix).diagonal()
        if kind <= 1:
            Av = average_face_to_cell(g) if proj == "F" else average_edge_to_cell(g)
            base = 3 * (Av.T @ sp.diags(V))
            ncol = g.nC
        else:
            Av = average_face_to_cell_vector(g) if proj == "F" else average_edge_to_cell_vector(g)
            base = Av.T @ sp.diags(np.tile(V, 3))
            ncol = 3 * g.nC
        if kind == 0:
            m = model
            ones = sp.csr_matrix(np.ones((g.nC, 1)))
            if not invert_model and not invert_matrix:
                d = base @ ones
            elif invert_model and invert_matrix:
                d = sp.diags(MI ** 2) @ base @ ones * (1.0 / m ** 2)
            elif invert_model:
                d = base @ sp.diags(-1.0 / (m * np.ones(g.nC)) ** 2)
            else:
                d = sp.diags(-MI ** 2) @ base
        else:
            m = np.asarray(model, dtype=np.float64).reshape(-1, order="F")
            if not invert_model and not invert_matrix:
                d = base
            elif invert_model and invert_matrix:
                d = sp.diags(MI ** 2) @ base @ sp.diags(1.0 / m ** 2)
            elif invert_model:
                d = base @ sp.diags(-1.0 / m ** 2)
            else:
                d = sp.diags(-MI ** 2) @ base
        return canonical(sp.diags(v) @ d)

    if invert_model or invert_matrix:
        raise NotImplementedError(
            "inverting the property or the matrix is not yet implemented for this mesh/tensorType. You should write it!"
        )
    # generic: columns blocks per parameter (xx,yy,zz,xy,xz,yz); inner_products.py:520-563
    sq = np.sqrt(0.125 * V)
    c = np.arange(g.nC)
    if kind == 0:
        nblk, pairs = 1, [[(0, 0), (1, 1), (2, 2)]]
    elif kind == 1:
        nblk, pairs = 1, [[(0, 0), (1, 1), (2, 2)]]
    elif kind == 2:
        nblk, pairs = 3, [[(0, 0)], [(1, 1)], [(2, 2)]]
    else:
        nblk, pairs = 6, [[(0, 0)], [(1, 1)], [(2, 2)], [(0, 1), (1, 0)], [(0, 2), (2, 0)], [(1, 2), (2, 1)]]
    rows, cols, vals = [], [], []
    for mem in _corner_members(g, proj):
        y = [sq * v[mem[a]] for a in range(3)]  # P v, per component
        for blk, pl in enumerate(pairs):
            for (a, b) in pl:
                # d/dsigma_ab of  P^T Sigma (P v):  row = member a, weight sq * y_b
                rows.append(mem[a])
                cols.append((np.zeros(g.nC, dtype=np.int64) if kind == 0 else c + blk * g.nC))
                vals.append(sq * y[b])
    ncol = 1 if kind == 0 else nblk * g.nC
    return _csr(rows, cols, vals, (n, ncol))


# ------------------------------------------------------------------------------------
# interpolation (base/base_tensor_mesh.py:638-790; utils/interpolation_utils.py:44-172;
# _extensions/interputils_cython.pyx:42-182)
# ------------------------------------------------------------------------------------
_LOC_ALIASES = {
    "CC": "cell_centers", "N": "nodes", "Fx": "faces_x", "Fy": "faces_y", "Fz": "faces_z",
    "Ex": "edges_x", "Ey": "edges_y", "Ez": "edges_z",
    "CCVx": "cell_centers_x", "CCVy": "cell_centers_y", "CCVz": "cell_centers_z",
}


def location_tensor(g: Grid, loc):
    """get_tensor (base_tensor_mesh.py:567-634): the three 1-D sample arrays of a location grid."""
    N, C = g.nodes, g.centers
    return {
        "cell_centers": (C[0], C[1], C[2]), "nodes": (N[0], N[1], N[2]),
        "faces_x": (N[0], C[1], C[2]), "faces_y": (C[0], N[1], C[2]), "faces_z": (C[0], C[1], N[2]),
        "edges_x": (C[0], N[1], N[2]), "edges_y": (N[0], C[1], N[2]), "edges_z": (N[0], N[1], C[2]),
    }[loc]


def is_inside(g: Grid, pts):
    """base_tensor_mesh.py:638-682: always against the nodes tensor, TOL = 1e-10*min(diff)."""
    inside = np.ones(pts.shape[0], dtype=bool)
    for a in range(3):
        t = g.nodes[a]
        tol = np.diff(t).min() * 1.0e-10
        inside &= (pts[:, a] >= t.min() - tol) & (pts[:, a] <= t.max() + tol)
    return inside


def _axis_inds_ws(x, xp):
    """_get_inds_ws (interputils_cython.pyx:56-67) vectorised; bisect_right == searchsorted(side='right')."""
    nx = x.size
    ind = np.searchsorted(x, xp, side="right").astype(np.int64)
    i2 = np.clip(ind, 0, nx - 1)
    i1 = np.clip(ind - 1, 0, nx - 1)
    same = i1 == i2
    den = np.where(same, 1.0, x[i2] - x[i1])
    w1 = np.where(same, 0.5, (x[i2] - xp) / den)
    return i1, i2, w1, 1 - w1


def interp_entries(pts, x, y, z):
    """_interpmat3D (interputils_cython.pyx:128-182): (npts,8) flat columns (F-order) and weights."""
    xi1, xi2, xw1, xw2 = _axis_inds_ws(x, pts[:, 0])
    yi1, yi2, yw1, yw2 = _axis_inds_ws(y, pts[:, 1])
    zi1, zi2, zw1, zw2 = _axis_inds_ws(z, pts[:, 2])
    nx, ny = x.size, y.size
    cols, vals = [], []
    for (zi, zw) in ((zi1, zw1), (zi2, zw2)):
        for (xi, xw) in ((xi1, xw1), (xi2, xw2)):
            for (yi, yw) in ((yi1, yw1), (yi2, yw2)):
                cols.append(xi + nx * (yi + ny * zi))
                vals.append(xw * yw * zw)
    return np.stack(cols, axis=1), np.stack(vals, axis=1)


def interpolation_matrix(g: Grid, pts, location_type="cell_centers", zeros_outside=False):
    """TensorMesh.get_interpolation_matrix (base_tensor_mesh.py:684-790)."""
    pts = np.atleast_2d(np.asarray(pts, dtype=np.float64))
    loc = _LOC_ALIASES.get(location_type, location_type)
    inside = is_inside(g, pts)
    if not zeros_outside:
        if not np.all(inside):
            raise ValueError("Points outside of mesh")
    else:
        pts = pts.copy()
        pts[~inside, :] = np.array([v.mean() for v in g.centers])
    npts = pts.shape[0]
    if loc in ("cell_centers", "nodes"):
        base, off, ncols = loc, 0, (g.nC if loc == "cell_centers" else g.nN)
    elif loc.startswith("faces_") or loc.startswith("edges_"):
        base = loc
        kind = ("F" if loc[0] == "f" else "E") + loc[-1]
        off, ncols = g.offset(kind), (g.nF if loc[0] == "f" else g.nE)
    elif loc in ("cell_centers_x", "cell_centers_y", "cell_centers_z"):
        base, off, ncols = "cell_centers", "xyz".index(loc[-1]) * g.nC, 3 * g.nC
    else:
        raise NotImplementedError("get_interpolation_matrix: location_type==" + loc + " and mesh.dim==3")
    cols, vals = interp_entries(pts, *location_tensor(g, base))
    rows = np.repeat(np.arange(npts), 8)
    Q = sp.csr_matrix((vals.reshape(-1), (rows, cols.reshape(-1) + off)), shape=(npts, ncols))
    Q.sum_duplicates()
    Q.sort_indices()
    if zeros_outside and (~inside).any():
        # `Q[indZeros, :] = 0` on a CSR matrix stores a *full row of explicit zeros*
        # for every outside point (base_tensor_mesh.py:771-772).
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


# ------------------------------------------------------------------------------------
# cell-centred operators: tensor products of 1-D two-point stencils
# ------------------------------------------------------------------------------------
def _s_eye(n):
    i = np.arange(n)
    return i, i, np.ones(n)


def _s_av(n):
    """cells <- nodes, (1/2, 1/2)  (utils/matrix_utils.py:227-253)."""
    i = np.arange(n)
    return np.r_[i, i], np.r_[i, i + 1], np.full(2 * n, 0.5)


def _s_av_extrap(n):
    """nodes <- cells, (1/2, 1/2) inside and weight 1 on the two boundary nodes (matrix_utils.py:256-287)."""
    i = np.arange(1, n)
    r = np.r_[0, i, i, n]
    c = np.r_[0, i - 1, i, n - 1]
    v = np.r_[1.0, np.full(2 * (n - 1), 0.5), 1.0]
    return r, c, v


def _s_cell_grad(n, bc):
    """nodes <- cells, (-1, +1); boundary rows by ghost point: dirichlet +-2, neumann an explicit 0
    (operators/differential_operators.py:41-82)."""
    i = np.arange(1, n)
    first = {"dirichlet": 2.0, "neumann": 0.0}[bc[0]]
    last = {"dirichlet": -2.0, "neumann": 0.0}[bc[1]]
    r = np.r_[0, i, i, n]
    c = np.r_[0, i - 1, i, n - 1]
    v = np.r_[first, np.full(n - 1, -1.0), np.full(n - 1, 1.0), last]
    return r, c, v


def _tensor_term(g: Grid, row_kind, col_kind, sx, sy, sz, scale=1.0, row_offset=0, col_offset=None):
    """COO triplets of kron3(sz, sy, sx) between two locations (x fastest)."""
    (rx, cx, vx), (ry, cy, vy), (rz, cz, vz) = sx, sy, sz
    ax, ay, az = np.meshgrid(np.arange(rx.size), np.arange(ry.size), np.arange(rz.size), indexing="ij")
    ax, ay, az = ax.ravel(), ay.ravel(), az.ravel()
    rsx, rsy, _ = g.shapes(row_kind)
    csx, csy, _ = g.shapes(col_kind)
    rows = row_offset + rx[ax] + rsx * (ry[ay] + rsy * rz[az])
    coff = g.offset(col_kind) if col_offset is None else col_offset
    cols = coff + cx[ax] + csx * (cy[ay] + csy * cz[az])
    return rows, cols, scale * vx[ax] * vy[ay] * vz[az]


def _stack(g: Grid, terms, shape):
    rows, cols, vals = zip(*terms)
    return _csr(list(rows), list(cols), list(vals), shape)


_NORMAL_AXIS = {"Fx": 0, "Fy": 1, "Fz": 2}


def _normal_stencils(g: Grid, axis, s_normal):
    """(sx, sy, sz) with ``s_normal`` along ``axis`` and identities (cell extents) elsewhere."""
    s = [_s_eye(n) for n in g.n]
    s[axis] = s_normal
    return s


def average_cell_to_face(g: Grid):
    """vstack of kron3(I, I, av_extrap) per direction: nF x nC (differential_operators.py:2789-2828)."""
    terms = []
    for kind, a in _NORMAL_AXIS.items():
        terms.append(_tensor_term(g, kind, "C", *_normal_stencils(g, a, _s_av_extrap(g.n[a])), row_offset=g.offset(kind)))
    return _stack(g, terms, (g.nF, g.nC))


def average_cell_vector_to_face(g: Grid):
    """block_diag of the three components: nF x 3nC (differential_operators.py:2830-2865)."""
    terms = []
    for kind, a in _NORMAL_AXIS.items():
        terms.append(_tensor_term(g, kind, "C", *_normal_stencils(g, a, _s_av_extrap(g.n[a])), row_offset=g.offset(kind),
                                  col_offset=a * g.nC))
    return _stack(g, terms, (g.nF, 3 * g.nC))


def average_cell_to_edge(g: Grid):
    """av_extrap on the two axes transverse to the edge: nE x nC (differential_operators.py:2867-2892)."""
    terms = []
    for a, kind in enumerate(("Ex", "Ey", "Ez")):
        s = [_s_av_extrap(n) for n in g.n]
        s[a] = _s_eye(g.n[a])
        terms.append(_tensor_term(g, kind, "C", *s, row_offset=g.offset(kind)))
    return _stack(g, terms, (g.nE, g.nC))


def average_edge_to_face(g: Grid):
    """0.5 * [[0, Ey->Fx, Ez->Fx], [Ex->Fy, 0, Ez->Fy], [Ex->Fz, Ey->Fz, 0]]: each face takes the mean of the
    four edges that bound it (differential_operators.py:3205-3234)."""
    terms = []
    for fa, fk in enumerate(("Fx", "Fy", "Fz")):
        for ea, ek in enumerate(("Ex", "Ey", "Ez")):
            if ea == fa:
                continue
            third = 3 - fa - ea            # the axis that is averaged (node extent on the edge, cell extent on the face)
            fs, es = g.shapes(fk), g.shapes(ek)
            s = []
            for ax in range(3):
                s.append(_s_av(g.n[ax]) if ax == third else _s_eye(fs[ax]))
                assert ax == third or fs[ax] == es[ax]
            terms.append(_tensor_term(g, fk, ek, *s, scale=0.5, row_offset=g.offset(fk)))
    return _stack(g, terms, (g.nF, g.nE))


def average_node_to_edge(g: Grid):
    """mean of the two end nodes: nE x nN (differential_operators.py:3254-3319)."""
    terms = []
    sn = g.shapes("N")
    for a, kind in enumerate(("Ex", "Ey", "Ez")):
        s = [_s_eye(sn[ax]) for ax in range(3)]
        s[a] = _s_av(g.n[a])
        terms.append(_tensor_term(g, kind, "N", *s, row_offset=g.offset(kind)))
    return _stack(g, terms, (g.nE, g.nN))


def average_node_to_face(g: Grid):
    """mean of the four corner nodes: nF x nN (differential_operators.py:3321-3382)."""
    terms = []
    sn = g.shapes("N")
    for a, kind in enumerate(("Fx", "Fy", "Fz")):
        s = [_s_av(g.n[ax]) for ax in range(3)]
        s[a] = _s_eye(sn[a])
        terms.append(_tensor_term(g, kind, "N", *s, row_offset=g.offset(kind)))
    return _stack(g, terms, (g.nF, g.nN))


def normalize_bc(bc):
    """set_cell_gradient_BC / _validate_BC (differential_operators.py:21-38, 975-1046)."""
    if isinstance(bc, str):
        bc = [bc] * 3
    out = []
    for b in bc:
        b = [b, b] if isinstance(b, str) else list(b)
        assert len(b) == 2 and all(x in ("dirichlet", "neumann") for x in b)
        out.append(b)
    return out


def stencil_cell_gradient(g: Grid, bc="neumann", axis=None):
    """vstack of kron3(I, I, _ddxCellGrad(n, bc)) per direction: nF x nC (differential_operators.py:1413-1443);
    ``axis`` selects one component block with Neumann conditions (stencil_cell_gradient_x/y/z, :1167-1181)."""
    bc = normalize_bc(bc)
    if axis is not None:
        kind = ("Fx", "Fy", "Fz")[axis]
        t = _tensor_term(g, kind, "C", *_normal_stencils(g, axis, _s_cell_grad(g.n[axis], ["neumann", "neumann"])))
        return _stack(g, [t], (g.vnF[axis], g.nC))
    terms = []
    for kind, a in _NORMAL_AXIS.items():
        terms.append(_tensor_term(g, kind, "C", *_normal_stencils(g, a, _s_cell_grad(g.n[a], bc[a])), row_offset=g.offset(kind)))
    return _stack(g, terms, (g.nF, g.nC))


def cell_gradient(g: Grid, bc="neumann", axis=None):
    """diag(S / (aveCC2F V)) @ stencil (differential_operators.py:1445-1595, 1762-2083); the product drops the
    explicit zeros of the Neumann rows."""
    S = g.face_areas()
    Vf = average_cell_to_face(g) @ g.cell_volumes()
    scale = S / Vf
    if axis is not None:
        lo = g.offset(("Fx", "Fy", "Fz")[axis])
        scale = scale[lo:lo + g.vnF[axis]]
    G = sp.diags(scale).tocsr() @ stencil_cell_gradient(g, bc, axis)
    G.eliminate_zeros()
    return canonical(G)


def face_component_divergence(g: Grid, axis):
    """face_x/y/z_divergence: diag(1/V) Dstencil_axis diag(S_axis), nC x nF_axis (differential_operators.py:236-591)."""
    kind = ("Fx", "Fy", "Fz")[axis]
    D = face_divergence(g)
    lo = g.offset(kind)
    return canonical(D[:, lo:lo + g.vnF[axis]])

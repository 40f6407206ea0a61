"""TensorMesh operators: thin host objects over the libdzb kernels."""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from ._lib import check, load
from .operators import Operator, csr_from_device, cuda_device, ptr, stream_ptr, to_device


def _csr_two_pass(nrows, ncols, count, fill):
    """count(row_nnz) ; indptr = exclusive scan ; fill(indptr, indices, data) -> scipy CSR."""
    dev = cuda_device()
    row_nnz = torch.zeros(nrows, dtype=torch.int64, device=dev)
    count(row_nnz)
    indptr = torch.zeros(nrows + 1, dtype=torch.int64, device=dev)
    torch.cumsum(row_nnz, 0, out=indptr[1:])
    nnz = int(indptr[-1].item())
    indices = torch.empty(nnz, dtype=torch.int64, device=dev)
    data = torch.empty(nnz, dtype=torch.float64, device=dev)
    fill(indptr, indices, data)
    return csr_from_device(indptr, indices, data, (nrows, ncols))


class StencilOperator(Operator):
    """DIV / CURL / GRAD / averaging, matrix-free (K1-K4).

    reference: operators/differential_operators.py:225-235 (face_divergence), :2169-2181
    (edge_curl), :658-664 (nodal_gradient), :2478-3253 (average_*).
    """

    def __init__(self, mesh, op_id, transpose=False, name=None, flags=0):
        self.flags = int(flags)
        nr, nc = C.c_int64(), C.c_int64()
        check(load().dzb_tm_shape(C.byref(mesh._host_mesh_counts()), op_id, C.byref(nr), C.byref(nc)), "dzb_tm_shape")
        shape = (nc.value, nr.value) if transpose else (nr.value, nc.value)
        super().__init__(shape)
        self.mesh, self.op_id, self.transposed, self.name = mesh, op_id, bool(transpose), name

    def _launch(self, x, out, transpose):
        check(load().dzb_tm_apply_flags(C.byref(self.mesh.device_mesh()), self.op_id, int(transpose), self.flags, ptr(x),
                                        ptr(out), stream_ptr()), "dzb_tm_apply")

    def _dev_apply(self, x, out):
        self._launch(x, out, self.transposed)

    def _dev_apply_t(self, x, out):
        self._launch(x, out, not self.transposed)

    def _transpose(self):
        return StencilOperator(self.mesh, self.op_id, not self.transposed, self.name, self.flags)

    _adjoint = _transpose

    def _tocsr(self):
        lib, dm, tr, fl = load(), self.mesh.device_mesh(), int(self.transposed), self.flags
        return _csr_two_pass(
            self.shape[0], self.shape[1],
            lambda rn: check(lib.dzb_tm_csr_count_flags(C.byref(dm), self.op_id, tr, fl, ptr(rn), stream_ptr()), "csr_count"),
            lambda ip, ix, da: check(lib.dzb_tm_csr_fill_flags(C.byref(dm), self.op_id, tr, fl, ptr(ip), ptr(ix), ptr(da),
                                                               stream_ptr()), "csr_fill"),
        )

    def __repr__(self):
        t = ".T" if self.transposed else ""
        return f"<{self.shape[0]}x{self.shape[1]} StencilOperator {self.name}{t} (CUDA, matrix-free)>"


# ----------------------------------------------------------------------------------------
# models
# ----------------------------------------------------------------------------------------
class _Model:
    """Classified physical property (utils/matrix_utils.py:1048-1068 TensorType)."""

    def __init__(self, mesh, model):
        from .tensor_mesh import is_scalar

        nC = mesh.n_cells
        self.scalar = 0.0
        self.dev = None
        if model is None:
            self.tt, self.kind, self.scalar = -1, _lib.MODEL_SCALAR, 1.0
        elif is_scalar(model) and not isinstance(model, torch.Tensor):
            self.tt, self.kind, self.scalar = 0, _lib.MODEL_SCALAR, float(np.asarray(model).reshape(-1)[0])
        else:
            size = model.numel() if isinstance(model, torch.Tensor) else np.asarray(model).size
            if size == nC:
                self.tt, self.kind, ncol = 1, _lib.MODEL_ISO, 1
            elif size == nC * 3:
                self.tt, self.kind, ncol = 2, _lib.MODEL_ANISO, 3
            elif size == nC * 6:
                self.tt, self.kind, ncol = 3, _lib.MODEL_TENSOR, 6
            else:
                shp = tuple(model.shape)
                raise Exception("Unexpected shape of tensor: {}".format(shp))
            # Fortran-ordered columns, like mkvc / reshape(order="F") in the reference
            if isinstance(model, torch.Tensor):
                t = model.to(torch.float64)
                if t.ndim == 2:
                    t = t.T
                self.dev = to_device(t).reshape(-1)
            else:
                a = np.asarray(model, dtype=np.float64)
                self.dev = to_device(a.reshape(-1, order="F"))
        self.n_params = {-1: 1, 0: 1, 1: nC, 2: 3 * nC, 3: 6 * nC}[self.tt]


# ----------------------------------------------------------------------------------------
# mass matrices
# ----------------------------------------------------------------------------------------
class MassOperator(Operator):
    """get_face_inner_product / get_edge_inner_product as an operator (K5 / K6).

    Symmetric; the matrix is never stored: entries are generated from the model,
    cell widths and (for full tensors) the corner tables inside the kernel.
    """

    def __init__(self, mesh, proj, model, invert_model, invert_matrix):
        n = mesh.n_faces if proj == "F" else mesh.n_edges
        super().__init__((n, n))
        self.mesh, self.proj, self.model = mesh, proj, model
        self.invert_model, self.invert_matrix = bool(invert_model), bool(invert_matrix)
        self.flags = ((_lib.FLAG_EDGES if proj == "E" else 0)
                      | (_lib.FLAG_INVERT_MODEL if invert_model else 0)
                      | (_lib.FLAG_INVERT_MATRIX if invert_matrix else 0))
        # The model is fixed for the life of the operator, so the divisions are taken out of the apply path (an
        # FP64 division is ~10 dependent instructions; with 1/model AND 1/M in the kernel the apply ran at 0.48 of
        # HBM): 1/model is formed once here (same IEEE division as in the kernel), and with invert_matrix the
        # inverse diagonal is cached on first use and applied as a plain diagonal scale.
        self._apply_model, self._apply_scalar, self._apply_flags = model.dev, model.scalar, self.flags
        if self.is_diagonal and invert_model:
            if model.kind == _lib.MODEL_SCALAR:
                self._apply_scalar = 1.0 / model.scalar
            else:
                self._apply_model = 1.0 / model.dev
            self._apply_flags &= ~_lib.FLAG_INVERT_MODEL
        if not self.is_diagonal and invert_model:
            # full tensors: Sigma^-1 of every cell once (the closed form the kernels use), so that the apply takes the
            # marching kernels K6 like a plain tensor model
            inv = torch.empty_like(model.dev)
            check(load().dzb_tm_tensor_invert(mesh.n_cells, ptr(model.dev), ptr(inv), stream_ptr()), "dzb_tm_tensor_invert")
            self._apply_model = inv
            self._apply_flags &= ~_lib.FLAG_INVERT_MODEL
        self._inv_diag = None

    @property
    def is_diagonal(self):
        return self.model.kind != _lib.MODEL_TENSOR

    def _dev_apply(self, x, out):
        if self.is_diagonal and self.invert_matrix:
            if self._inv_diag is None:
                self._inv_diag = torch.empty(self.shape[0], dtype=torch.float64, device=cuda_device())
                check(load().dzb_tm_mass_diag(C.byref(self.mesh.device_mesh()), self._apply_flags, self.model.kind,
                                              ptr(self._apply_model), self._apply_scalar, ptr(self._inv_diag),
                                              stream_ptr()), "dzb_tm_mass_diag")
            check(load().dzb_diag_scale(self.shape[0], ptr(self._inv_diag), ptr(x), ptr(out), 0, stream_ptr()),
                  "dzb_diag_scale")
            return
        check(load().dzb_tm_mass_apply(C.byref(self.mesh.device_mesh()), self._apply_flags, self.model.kind,
                                       ptr(self._apply_model), self._apply_scalar, ptr(x), ptr(out), stream_ptr()),
              "dzb_tm_mass_apply")

    _dev_apply_t = _dev_apply

    def _transpose(self):
        return self

    _adjoint = _transpose

    def diagonal_device(self):
        """Diagonal of a diagonal mass matrix as a CUDA tensor."""
        if not self.is_diagonal:
            raise ValueError("full-tensor mass matrices are not diagonal")
        out = torch.empty(self.shape[0], dtype=torch.float64, device=cuda_device())
        check(load().dzb_tm_mass_diag(C.byref(self.mesh.device_mesh()), self.flags, self.model.kind,
                                      ptr(self.model.dev), self.model.scalar, ptr(out), stream_ptr()),
              "dzb_tm_mass_diag")
        return out

    def diagonal(self):
        if self.is_diagonal:
            return self.diagonal_device().cpu().numpy()
        return self.tocsr().diagonal()

    def _tocsr(self):
        if self.is_diagonal:
            return sp.diags(self.diagonal(), format="csr")
        lib, dm = load(), self.mesh.device_mesh()
        m = _csr_two_pass(
            self.shape[0], self.shape[1],
            lambda rn: check(lib.dzb_tm_mass_tensor_csr_count(C.byref(dm), self.flags, ptr(rn), stream_ptr()),
                             "mass_tensor_csr_count"),
            lambda ip, ix, da: check(lib.dzb_tm_mass_tensor_csr_fill(C.byref(dm), self.flags, ptr(self.model.dev),
                                                                     ptr(ip), ptr(ix), ptr(da), stream_ptr()),
                                     "mass_tensor_csr_fill"),
        )
        m.eliminate_zeros()  # scipy's sparse products / sums in the reference never store exact zeros
        return m


def make_mass_operator(mesh, proj, model, invert_model, invert_matrix, do_fast):
    """reference: operators/inner_products.py:147-219."""
    if proj not in ("F", "E"):
        raise TypeError("projection_type must be 'F' for faces or 'E' for edges")
    md = _Model(mesh, model)
    if md.kind == _lib.MODEL_TENSOR and invert_matrix:
        raise Exception("Solver needed to invert A.")
    # do_fast=False takes the corner-sum route in the reference; for scalar / isotropic /
    # anisotropic models it yields the same diagonal matrix, so both routes share K5.
    return MassOperator(mesh, proj, md, invert_model, invert_matrix)


# ----------------------------------------------------------------------------------------
# derivatives of the mass matrices with respect to the model
# ----------------------------------------------------------------------------------------
class MassDerivOperator(Operator):
    """F(v) = d(M(model) v)/d(model): n x n_params (K7).

    reference: base/base_tensor_mesh.py:865-988 (fast path), inner_products.py:459-565 (tensor).
    ``F(v) @ dm`` and ``F(v).T @ w`` are matrix-free kernels.
    """

    def __init__(self, mesh, proj, model, v, invert_model, invert_matrix, ncols, scalar_mode=None):
        n = mesh.n_faces if proj == "F" else mesh.n_edges
        super().__init__((n, ncols))
        self.mesh, self.proj, self.model = mesh, proj, model
        self.v = to_device(v).reshape(-1)
        if self.v.numel() != n:
            raise ValueError(f"v must have {n} entries")
        self.scalar_mode = scalar_mode  # None | "broadcast": n x 1 built from the isotropic kernel
        self.flags = ((_lib.FLAG_EDGES if proj == "E" else 0)
                      | (_lib.FLAG_INVERT_MODEL if invert_model else 0)
                      | (_lib.FLAG_INVERT_MATRIX if invert_matrix else 0))
        self._iso_model = None
        if model.kind == _lib.MODEL_SCALAR:
            self._iso_model = torch.full((mesh.n_cells,), model.scalar, dtype=torch.float64, device=cuda_device())

    # With inversions the reference's matrix is  diag(v a) [n_el Av^T V] diag(b),  a = +-MI^2 (MI = diagonal of the inverted
    # mass matrix), b = -+1/m^2 (base_tensor_mesh.py:917-978).  v and the model are fixed per operator, so u = v a and b
    # are formed once and every apply is the inversion-free marching kernel on (u, b o dm) -- the same bytes as without
    # inversions plus one pass over the cells -- instead of the gather kernels with the inversions inside (0.24-0.33 of HBM).
    _fast_inverse = True

    def _inverse_factors(self):
        if getattr(self, "_u", None) is None:
            kind, mp = self.model.kind, self.model.dev
            if kind == _lib.MODEL_SCALAR:
                kind, mp = _lib.MODEL_ISO, self._iso_model
            inv_model = bool(self.flags & _lib.FLAG_INVERT_MODEL)
            inv_matrix = bool(self.flags & _lib.FLAG_INVERT_MATRIX)
            u = self.v
            if inv_matrix:
                mi = torch.empty_like(self.v)
                check(load().dzb_tm_mass_diag(C.byref(self.mesh.device_mesh()), self.flags, kind, ptr(mp), 0.0, ptr(mi),
                                              stream_ptr()), "dzb_tm_mass_diag")
                u = self.v * mi * mi
                if not inv_model:
                    u = -u
            b = None
            if inv_model:
                b = 1.0 / (mp * mp)
                if not inv_matrix:
                    b = -b
            self._u, self._b, self._kind = u, b, kind
        return self._u, self._b, self._kind

    def _call(self, transpose, x, out):
        kind, mp = self.model.kind, self.model.dev
        if kind == _lib.MODEL_SCALAR:
            kind, mp = _lib.MODEL_ISO, self._iso_model
        inverted = self.flags & (_lib.FLAG_INVERT_MODEL | _lib.FLAG_INVERT_MATRIX)
        if inverted and self._fast_inverse and kind in (_lib.MODEL_ISO, _lib.MODEL_ANISO):
            u, b, kind = self._inverse_factors()
            plain = self.flags & _lib.FLAG_EDGES
            mesh = C.byref(self.mesh.device_mesh())
            if not transpose:
                dm = x if b is None else x * b
                check(load().dzb_tm_mass_apply(mesh, plain, kind, ptr(dm), 0.0, ptr(u), ptr(out), stream_ptr()),
                      "dzb_tm_mass_apply")
            else:
                check(load().dzb_tm_mass_deriv_apply(mesh, plain, kind, ptr(mp), 0.0, ptr(u), 1, ptr(x), ptr(out),
                                                     stream_ptr()), "dzb_tm_mass_deriv_apply")
                if b is not None:
                    out.mul_(b)
            return
        check(load().dzb_tm_mass_deriv_apply(C.byref(self.mesh.device_mesh()), self.flags, kind, ptr(mp), 0.0,
                                             ptr(self.v), int(transpose), ptr(x), ptr(out), stream_ptr()),
              "dzb_tm_mass_deriv_apply")

    def _dev_apply(self, x, out):
        if self.scalar_mode == "broadcast":
            x = x.expand(self.mesh.n_cells).contiguous()
        self._call(False, x, out)

    def _dev_apply_t(self, x, out):
        if self.scalar_mode == "broadcast":
            tmp = torch.empty(self.mesh.n_cells, dtype=torch.float64, device=x.device)
            self._call(True, x, tmp)
            out.copy_(tmp.sum().reshape(1))
        else:
            self._call(True, x, out)

    def _tocsr(self):
        # columns generated by applying F(v) to unit vectors is O(n_params) launches; instead use
        # the structure: every column block is  diag-scaled incidence, recovered from F(v)^T e_f.
        # Small meshes only (parity / debugging): probe with the identity in blocks.
        n, k = self.shape
        if n * 1.0 * k > 4e8:
            raise MemoryError("MassDerivOperator.tocsr() is meant for small meshes")
        eye = torch.eye(k, dtype=torch.float64, device=cuda_device())
        cols = [self.apply_device(eye[:, c].contiguous(), False) for c in range(k)]
        dense = torch.stack(cols, dim=1).cpu().numpy() if cols else np.zeros((n, 0))
        return sp.csr_matrix(dense)


def make_mass_deriv(mesh, proj, model, do_fast, invert_model, invert_matrix):
    """Returns the callable ``v -> operator`` (inner_products.py:405-457)."""
    md = _Model(mesh, model)
    nC = mesh.n_cells
    if md.tt == -1:
        return None
    generic = (md.kind == _lib.MODEL_TENSOR) or not do_fast
    if generic and (invert_model or invert_matrix):
        raise NotImplementedError(
            "inverting the property or the matrix is not yet implemented for this mesh/tensorType. You should write it!"
        )
    if md.tt == 0:
        if invert_model and not invert_matrix and not generic:
            # the reference raises here (sdiag of a bare float / shape mismatch); keep it an error
            raise TypeError("Vector must be a numpy array")
        if invert_matrix and not invert_model and not generic:
            ncols, mode = nC, None  # reference quirk: n x nC (base_tensor_mesh.py:927-928)
        else:
            ncols, mode = 1, "broadcast"
    else:
        ncols, mode = md.n_params, None

    def inner_product_deriv(v=None):
        if v is None:
            raise Exception("v must be supplied for this implementation.")
        return MassDerivOperator(mesh, proj, md, v, invert_model, invert_matrix, ncols, mode)

    return inner_product_deriv


# ----------------------------------------------------------------------------------------
# fused composites
# ----------------------------------------------------------------------------------------
class DCNodalOperator(Operator):
    """G^T Me(sigma) G in one kernel (K8); symmetric positive semi-definite.

    reference composition: tutorials/pde/3_nodal_dirichlet_poisson.py:146-160.
    """

    def __init__(self, mesh, sigma_dev):
        n = mesh.n_nodes
        super().__init__((n, n))
        self.mesh, self.sigma = mesh, sigma_dev

    def _dev_apply(self, x, out):
        check(load().dzb_tm_dc_nodal(C.byref(self.mesh.device_mesh()), ptr(self.sigma), ptr(x), ptr(out), stream_ptr()),
              "dzb_tm_dc_nodal")

    _dev_apply_t = _dev_apply

    def _dev_apply_multi(self, xs, outs):
        """``outs[r] = A xs[r]`` for the rows of a (k, n) block in ONE launch (``matmat``): sigma is read from DRAM once
        per tile instead of once per right-hand side (the k blocks of a tile run side by side and share it through L2)."""
        # measured at 512^3 (tools/bench_multi_rhs.py): 2 right-hand sides +4.5 %, 4: +7.7 %, 8: none -- beyond 4 the tiles of a
        # group no longer stay in L2 together, so larger blocks go in groups of 4
        for r0 in range(0, xs.shape[0], 4):
            k = min(4, xs.shape[0] - r0)
            check(load().dzb_tm_dc_nodal_multi(C.byref(self.mesh.device_mesh()), ptr(self.sigma), ptr(xs[r0:r0 + k]),
                                               ptr(outs[r0:r0 + k]), k, int(xs.shape[1]), stream_ptr()), "dzb_tm_dc_nodal_multi")

    _dev_apply_t_multi = _dev_apply_multi

    def diagonal_device(self):
        """diag(G^T Me G) = (G o G)^T diag(Me): one apply of the entrywise-squared gradient to the edge mass
        diagonal (the Jacobi preconditioner of ``solvers.pcg``)."""
        me = torch.empty(self.mesh.n_edges, dtype=torch.float64, device=cuda_device())
        check(load().dzb_tm_mass_diag(C.byref(self.mesh.device_mesh()), _lib.FLAG_EDGES, _lib.MODEL_ISO, ptr(self.sigma),
                                      0.0, ptr(me), stream_ptr()), "dzb_tm_mass_diag")
        return StencilOperator(self.mesh, _lib.OP_GRAD_SQ, transpose=True, name="grad_sq").apply_device(me)

    def diagonal(self):
        return self.diagonal_device().cpu().numpy()

    def apply_planes(self, x, out, k_begin, k_end):
        """Write only node planes k_begin <= k < k_end of y = A x (z-slab building block)."""
        check(load().dzb_tm_dc_nodal_range(C.byref(self.mesh.device_mesh()), ptr(self.sigma), ptr(x), ptr(out),
                                           int(k_begin), int(k_end), stream_ptr()), "dzb_tm_dc_nodal_range")

    def apply_with_dot(self, x, out, dot):
        """``out = A x`` and ``dot[0] = x . out`` in one pass (``dzb_tm_dc_nodal_dot``): the p.Ap of a CG iteration without
        a second sweep over p and Ap.  ``dot``: a one-element CUDA tensor (or a view of one).  Returns False where the
        fused form does not apply (the caller then uses ``apply_device`` + a dot kernel)."""
        ws = getattr(self, "_dot_ws", None)
        if ws is None:
            nx, ny, nz = self.mesh.shape_cells
            blocks = ((nx + 1 + 539) // 540) * ((ny + 1 + 3) // 4) * 256       # z-chunks are at most 256
            ws = self._dot_ws = torch.empty(min(blocks * 18, 1 << 24), dtype=torch.float64, device=cuda_device())
        rc = load().dzb_tm_dc_nodal_dot(C.byref(self.mesh.device_mesh()), ptr(self.sigma), ptr(x), ptr(out), ptr(ws),
                                        int(ws.numel()), ptr(dot), stream_ptr())
        if rc == -2:
            return False
        check(rc, "dzb_tm_dc_nodal_dot")
        return True

    def apply_planes_peer(self, x, out, k_begin, k_end, lower=None, upper=None, sync=None):
        """Node planes [k_begin, k_end) of a z-slab whose halo planes are read from the neighbours' memory by the kernel.
        ``lower`` = (device pointer of the lower neighbour's local nodal vector, its cell layers, local halo plane, plane
        in the neighbour's array) or None; ``upper`` likewise.  ``sync`` = (this rank's sync words, the lower neighbour's
        signal word, the upper neighbour's) for ordering inside the kernel, or None (the caller put a barrier in front)."""
        lo = lower if lower is not None else (0, 0, -1, -1)
        hi = upper if upper is not None else (0, 0, -1, -1)
        sy = sync if sync is not None else (0, 0, 0)
        check(load().dzb_tm_dc_nodal_range_peer(C.byref(self.mesh.device_mesh()), ptr(self.sigma), ptr(x),
                                                C.c_void_p(int(lo[0])), int(lo[1]), int(lo[2]), int(lo[3]),
                                                C.c_void_p(int(hi[0])), int(hi[1]), int(hi[2]), int(hi[3]),
                                                C.c_void_p(int(sy[0])), C.c_void_p(int(sy[1])), C.c_void_p(int(sy[2])),
                                                ptr(out), int(k_begin), int(k_end), stream_ptr()),
              "dzb_tm_dc_nodal_range_peer")

    def matvec(self, x, out=None):
        # Host (CPU tensor) input: stream z-chunks through the GPU so that the H2D copy of chunk c+1, the
        # kernel on chunk c and the D2H copy of chunk c-1 overlap (PCIe is full duplex); the apply then
        # costs about one direction of the transfer instead of both plus the kernel.
        if isinstance(x, torch.Tensor) and x.device.type == "cpu" and x.ndim == 1 and x.is_pinned():
            return self._matvec_streamed(x, out)
        return super().matvec(x, out)

    rmatvec = matvec

    def _matvec_streamed(self, x, out, nchunks=32):   # tools/sweep_e2e.py: 8 -> 27.7 ms, 32 -> 25.2 ms at 512^3 (raw H2D alone: 19.4 ms)
        n = self.shape[0]
        if x.numel() != n:
            raise ValueError(f"dimension mismatch: operator {self.shape}, input {tuple(x.shape)}")
        dev = cuda_device()
        nz1 = self.mesh.shape_cells[2] + 1
        plane = n // nz1
        if out is None:
            out = torch.empty(n, dtype=torch.float64, pin_memory=True)
        if getattr(self, "_stage", None) is None or self._stage[0].device != dev:
            self._stage = (torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.float64, device=dev),
                           torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        xd, yd, s_in, s_out = self._stage
        main = torch.cuda.current_stream()
        nchunks = max(1, min(nchunks, nz1))
        cuts = [round(c * nz1 / nchunks) for c in range(nchunks + 1)]
        s_in.wait_stream(main)
        s_out.wait_stream(main)
        copied = []
        with torch.cuda.stream(s_in):
            for c in range(nchunks):
                a, b = cuts[c] * plane, cuts[c + 1] * plane
                xd[a:b].copy_(x[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
                copied.append(ev)
        for c in range(nchunks):
            main.wait_event(copied[min(c + 1, nchunks - 1)])   # chunk c needs the first plane of chunk c+1
            self.apply_planes(xd, yd, cuts[c], cuts[c + 1])
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                a, b = cuts[c] * plane, cuts[c + 1] * plane
                out[a:b].copy_(yd[a:b], non_blocking=True)
        main.wait_stream(s_out)
        main.synchronize()
        return out

    def _transpose(self):
        return self

    _adjoint = _transpose

    def _tocsr(self):
        G = self.mesh.nodal_gradient.tocsr()
        md = _Model(self.mesh, self.sigma)
        Me = MassOperator(self.mesh, "E", md, False, False).tocsr()
        return (G.T @ Me @ G).tocsr()


class DCCellOperator(Operator):
    """D Mf(rho)^-1 D^T in one kernel (cell-centred DC, SimPEG's default formulation); symmetric.

    reference composition: examples/plot_dc_resistivity.py:23-30.  ``neumann=True`` drops the boundary faces
    (zero-flux boundary); the plain product keeps them (homogeneous Dirichlet)."""

    def __init__(self, mesh, rho_dev, neumann=False, cache_inverse_mass=True):
        n = mesh.n_cells
        super().__init__((n, n))
        self.mesh, self.rho, self.neumann = mesh, rho_dev, bool(neumann)
        # cache_inverse_mass: keep 1/diag(Mf) (nF values) and run the division-free kernel (40 B/cell); False: generate
        # the face masses on the fly from rho (24 B/cell + five FP64 divisions per cell, no extra memory)
        self.cache_inverse_mass = bool(cache_inverse_mass)
        self._mi = None

    def _dev_apply(self, x, out):
        dm = C.byref(self.mesh.device_mesh())
        if not self.cache_inverse_mass:
            check(load().dzb_tm_dc_cell(dm, ptr(self.rho), int(self.neumann), ptr(x), ptr(out), stream_ptr()), "dzb_tm_dc_cell")
            return
        if self._mi is None:
            self._mi = torch.empty(self.mesh.n_faces, dtype=torch.float64, device=cuda_device())
            check(load().dzb_tm_mass_diag(dm, _lib.FLAG_INVERT_MATRIX, _lib.MODEL_ISO, ptr(self.rho), 0.0, ptr(self._mi),
                                          stream_ptr()), "dzb_tm_mass_diag")
        check(load().dzb_tm_dc_cell_mi(dm, ptr(self._mi), int(self.neumann), ptr(x), ptr(out), stream_ptr()),
              "dzb_tm_dc_cell_mi")

    _dev_apply_t = _dev_apply

    def _transpose(self):
        return self

    _adjoint = _transpose

    def _tocsr(self):
        D = self.mesh.face_divergence.tocsr()
        mfi = MassOperator(self.mesh, "F", _ModelFromDevice(self.mesh, self.rho), False, True).diagonal()   # 1 / diag(Mf)
        if self.neumann:  # boundary faces (a single adjacent cell) carry no flux
            mfi = np.where(np.diff(D.tocsc().indptr) == 2, mfi, 0.0)
        A = (D @ sp.diags(mfi) @ D.T).tocsr()
        if self.neumann:
            A.eliminate_zeros()
        A.sort_indices()
        return A


def _ModelFromDevice(mesh, dev_tensor):
    md = _Model(mesh, None)
    md.tt, md.kind, md.dev, md.scalar = 1, _lib.MODEL_ISO, dev_tensor, 0.0
    md.n_params = mesh.n_cells
    return md


class MaxwellOperator(Operator):
    """C^T Mf(mui) C + Me(sigma) in one kernel (K9); symmetric.

    reference composition: tests/boundaries/test_boundary_maxwell.py:95-99.
    """

    def __init__(self, mesh, mui_dev, sigma_dev):
        n = mesh.n_edges
        super().__init__((n, n))
        self.mesh, self.mui, self.sigma = mesh, mui_dev, sigma_dev

    def _dev_apply(self, x, out):
        check(load().dzb_tm_maxwell(C.byref(self.mesh.device_mesh()), ptr(self.mui), ptr(self.sigma), ptr(x), ptr(out),
                                    stream_ptr()), "dzb_tm_maxwell")

    _dev_apply_t = _dev_apply

    def diagonal_device(self):
        """diag(C^T Mf C + Me) = (C o C)^T diag(Mf) + diag(Me)."""
        dm = C.byref(self.mesh.device_mesh())
        mf = torch.empty(self.mesh.n_faces, dtype=torch.float64, device=cuda_device())
        me = torch.empty(self.mesh.n_edges, dtype=torch.float64, device=cuda_device())
        check(load().dzb_tm_mass_diag(dm, 0, _lib.MODEL_ISO, ptr(self.mui), 0.0, ptr(mf), stream_ptr()), "dzb_tm_mass_diag")
        check(load().dzb_tm_mass_diag(dm, _lib.FLAG_EDGES, _lib.MODEL_ISO, ptr(self.sigma), 0.0, ptr(me), stream_ptr()),
              "dzb_tm_mass_diag")
        d = StencilOperator(self.mesh, _lib.OP_CURL_SQ, transpose=True, name="curl_sq").apply_device(mf)
        return d.add_(me)

    def diagonal(self):
        return self.diagonal_device().cpu().numpy()

    def apply_planes(self, x, out, k_begin, k_end):
        """Write only Ex, Ey on node planes / Ez in cell layers k_begin <= k < k_end (z-slab building block)."""
        check(load().dzb_tm_maxwell_range(C.byref(self.mesh.device_mesh()), ptr(self.mui), ptr(self.sigma), ptr(x),
                                          ptr(out), int(k_begin), int(k_end), stream_ptr()), "dzb_tm_maxwell_range")

    def apply_planes2(self, x, out, a0, a1, b0, b1):
        """Two disjoint plane ranges [a0, a1) and [b0, b1) in one launch (the boundary planes of a z-slab)."""
        check(load().dzb_tm_maxwell_range2(C.byref(self.mesh.device_mesh()), ptr(self.mui), ptr(self.sigma), ptr(x),
                                           ptr(out), int(a0), int(a1), int(b0), int(b1), stream_ptr()),
              "dzb_tm_maxwell_range2")

    def apply_planes_peer(self, x, out, k_begin, k_end, lower=None, upper=None, sync=None):
        """Planes [k_begin, k_end) of a z-slab whose halo planes are read from the neighbours' memory by the kernel itself.
        ``lower`` = (device pointer of the lower neighbour's local edge vector, its cell layers, local halo plane, plane in
        the neighbour's array) or None; ``upper`` likewise.  ``sync`` = (this rank's sync words, the lower neighbour's
        signal word, the upper neighbour's) for ordering inside the kernel, or None (the caller put a barrier in front).
        Raises where the tensor-map path does not apply."""
        lo = lower if lower is not None else (0, 0, -1, -1)
        hi = upper if upper is not None else (0, 0, -1, -1)
        sy = sync if sync is not None else (0, 0, 0)
        check(load().dzb_tm_maxwell_range_peer_sync(C.byref(self.mesh.device_mesh()), ptr(self.mui), ptr(self.sigma), ptr(x),
                                                    C.c_void_p(int(lo[0])), int(lo[1]), int(lo[2]), int(lo[3]),
                                                    C.c_void_p(int(hi[0])), int(hi[1]), int(hi[2]), int(hi[3]),
                                                    C.c_void_p(int(sy[0])), C.c_void_p(int(sy[1])), C.c_void_p(int(sy[2])),
                                                    ptr(out), int(k_begin), int(k_end), stream_ptr()),
              "dzb_tm_maxwell_range_peer")

    def matvec(self, x, out=None):
        # pinned host input: stream z-chunks (H2D of chunk c+1, kernel on chunk c, D2H of chunk c-1 overlap)
        if isinstance(x, torch.Tensor) and x.device.type == "cpu" and x.ndim == 1 and x.is_pinned():
            return self._matvec_streamed(x, out)
        return super().matvec(x, out)

    rmatvec = matvec

    def _matvec_streamed(self, x, out, nchunks=32):
        n = self.shape[0]
        if x.numel() != n:
            raise ValueError(f"dimension mismatch: operator {self.shape}, input {tuple(x.shape)}")
        dev = cuda_device()
        nx, ny, nz = self.mesh.shape_cells
        px, py, pz = nx * (ny + 1), (nx + 1) * ny, (nx + 1) * (ny + 1)     # Ex, Ey entries per node plane, Ez per layer
        off_y, off_z = px * (nz + 1), px * (nz + 1) + py * (nz + 1)
        if out is None:
            out = torch.empty(n, dtype=torch.float64, pin_memory=True)
        if getattr(self, "_stage", None) is None or self._stage[0].device != dev:
            self._stage = (torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.float64, device=dev),
                           torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        xd, yd, s_in, s_out = self._stage
        main = torch.cuda.current_stream()
        nchunks = max(1, min(nchunks, nz + 1))
        cuts = [round(c * (nz + 1) / nchunks) for c in range(nchunks + 1)]

        def segments(a, b):   # node planes [a, b) of Ex and Ey, cell layers [a, min(b, nz)) of Ez
            segs = [(a * px, b * px), (off_y + a * py, off_y + b * py)]
            if min(b, nz) > a:
                segs.append((off_z + a * pz, off_z + min(b, nz) * pz))
            return segs

        s_in.wait_stream(main)
        s_out.wait_stream(main)
        copied = []
        with torch.cuda.stream(s_in):
            for c in range(nchunks):
                for lo, hi in segments(cuts[c], cuts[c + 1]):
                    xd[lo:hi].copy_(x[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
                copied.append(ev)
        for c in range(nchunks):
            main.wait_event(copied[min(c + 1, nchunks - 1)])   # chunk c needs the first plane of chunk c+1
            self.apply_planes(xd, yd, cuts[c], cuts[c + 1])
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                for lo, hi in segments(cuts[c], cuts[c + 1]):
                    out[lo:hi].copy_(yd[lo:hi], non_blocking=True)
        main.wait_stream(s_out)
        main.synchronize()
        return out

    def _transpose(self):
        return self

    _adjoint = _transpose

    def _tocsr(self):
        Cm = self.mesh.edge_curl.tocsr()
        Mf = MassOperator(self.mesh, "F", _Model(self.mesh, self.mui), False, False).tocsr()
        Me = MassOperator(self.mesh, "E", _Model(self.mesh, self.sigma), False, False).tocsr()
        return (Cm.T @ Mf @ Cm + Me).tocsr()


class ComplexMaxwellOperator(Operator):
    """C^T Mf(mui) C + (a + ib) Me(sigma): the frequency-domain Maxwell operator SimPEG applies (``alpha = 1j * omega``).

    With e = e_r + i e_i:  y_r = (K + a M) e_r - b M e_i,  y_i = (K + a M) e_i + b M e_r,  K = C^T Mf C, M = Me(sigma).
    ``K + a M`` is the fused kernel K9 on the model a*sigma, ``b M`` the edge mass kernel K5 on b*sigma: two launches of
    each per complex apply, nothing assembled.  Complex symmetric (``.T`` is the operator itself, ``.H`` its conjugate).
    """

    def __init__(self, mesh, mui_dev, sigma_dev, alpha):
        n = mesh.n_edges
        super().__init__((n, n), is_complex=True)
        alpha = complex(alpha)
        self.mesh, self.mui, self.sigma, self.alpha = mesh, mui_dev, sigma_dev, alpha
        self._kam = MaxwellOperator(mesh, mui_dev, sigma_dev * alpha.real)
        self._bm = MassOperator(mesh, "E", _Model(mesh, sigma_dev * alpha.imag), False, False)

    #: the one-launch kernel on interleaved complex edges (K9c); False forces the two-part path below (tests compare them)
    fused = True

    def apply_device_complex(self, xc, out=None):
        """y = A x for a complex128 CUDA tensor in ONE launch (K9c); returns None where the tensor-map path does not apply."""
        if out is None:
            out = torch.empty_like(xc)
        xr_, yr_ = torch.view_as_real(xc), torch.view_as_real(out)          # interleaved (re, im) storage
        rc = load().dzb_tm_maxwell_cplx(C.byref(self.mesh.device_mesh()), ptr(self.mui), ptr(self.sigma), self.alpha.real,
                                        self.alpha.imag, ptr(xr_), ptr(yr_), stream_ptr())
        if rc == -2:
            return None
        check(rc, "dzb_tm_maxwell_cplx")
        return out

    def _apply_complex(self, x, transpose):
        is_torch = isinstance(x, torch.Tensor)
        xa = x if is_torch else np.asarray(x)
        if xa.ndim == 2 and xa.shape[1] != 1:
            cols = [self._apply_complex(xa[:, c], transpose) for c in range(xa.shape[1])]
            return torch.stack(cols, dim=1) if is_torch else np.stack(cols, axis=1)
        flat = xa.reshape(-1)
        if flat.shape[0] != self.shape[1]:
            raise ValueError(f"dimension mismatch: operator {self.shape}, input {tuple(xa.shape)}")
        cplx = flat.is_complex() if is_torch else np.iscomplexobj(flat)
        if self.fused:
            dev = cuda_device()
            xc = (flat if is_torch else torch.from_numpy(np.ascontiguousarray(flat))).to(device=dev, dtype=torch.complex128).contiguous()
            yc = self.apply_device_complex(xc)
            if yc is not None:
                if is_torch:
                    yc = yc if x.device.type == "cuda" else yc.cpu()
                else:
                    yc = yc.cpu().numpy()
                return yc.reshape(xa.shape) if xa.ndim == 2 else yc
        xr = to_device(flat.real if cplx else flat)
        yr, yi = self._kam.apply_device(xr), self._bm.apply_device(xr)
        if cplx:
            xi = to_device(flat.imag)
            yr -= self._bm.apply_device(xi)
            yi += self._kam.apply_device(xi)
        y = torch.complex(yr, yi)
        if is_torch:
            y = y if x.device.type == "cuda" else y.cpu()
        else:
            y = y.cpu().numpy()
        return y.reshape(xa.shape) if xa.ndim == 2 else y

    def _transpose(self):
        return self

    def _adjoint(self):
        return ComplexMaxwellOperator(self.mesh, self.mui, self.sigma, self.alpha.conjugate())

    def _tocsr(self):
        return (self._kam.tocsr() + 1j * self._bm.tocsr()).tocsr()

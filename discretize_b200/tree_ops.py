"""TreeMesh operators on the GPU: coded CSR connectivity + the tree_spmv kernel (K12)."""
from __future__ import annotations


import numpy as np
import scipy.sparse as sp
import torch

from ._lib import check, load
from .operators import Operator, cuda_device, ptr, stream_ptr, to_device


def _unique_small(values):
    """np.unique(values, return_inverse=True) for arrays that hold only a handful of distinct values (the multipliers
    of an OcTree operator: signs times products of 1/2 hanging weights): candidates from a sample, a binary search
    of every entry in that tiny table, and a second round for whatever the sample missed -- O(n log k) instead of a
    full sort of 10^8 entries."""
    if values.size == 0:
        return values.copy(), np.zeros(0, dtype=np.int64)
    cand = np.unique(values[: min(values.size, 1 << 16)])
    while True:
        pos = np.minimum(np.searchsorted(cand, values), cand.size - 1)
        miss = cand[pos] != values
        if not miss.any():
            return cand, pos
        cand = np.unique(np.r_[cand, np.unique(values[miss])])


class CodeTableOverflow(ValueError):
    """More distinct (multiplier, selector) pairs than a code byte can name: the caller stores the values instead."""


MAX_CODES = 255          # code 255 marks the padding slots of the ELL layout


def _group_values(data):
    """(table values, code per entry) for the multipliers of one operator part.

    The table holds EXACT entries of the matrix.  Values that differ by a rounding error only (1/6 computed along two
    paths) share a code: neighbours of the sorted distinct values are merged when they agree to 1e-15 relative, and the
    smaller one represents the group -- so the applied values are within 1e-15 of the reference's, not within the 2e-14
    an absolute np.round(., 14) used to leave, and tiny values are not rounded away."""
    vals, inv = _unique_small(np.asarray(data, dtype=np.float64))
    if vals.size <= 1:
        return vals, inv
    gap = np.abs(np.diff(vals)) > 1e-15 * np.maximum(np.abs(vals[:-1]), np.abs(vals[1:]))
    group = np.r_[0, np.cumsum(gap)]                      # group id of every distinct value (sorted)
    rep = vals[np.r_[True, gap]]                          # first (smallest) member represents its group
    return rep, group[inv]


def fold_geometry(geoms, parts, geom_by_col):
    """Multiply the row (or column) geometry into the entries where the result still fits a code table.

    An operator is  sum_s diag(geom_s) @ M_s  (transposed: M_s^T @ diag(geom_s)).  On an OcTree over a UNIFORM base mesh
    -- what discretize's octrees are in practice -- 1/h, 1/L take one value per level, so the products
    multiplier * geometry are a handful of numbers: they go into the table as they are (the same IEEE product the kernel
    would form, so nothing changes bitwise) and the kernel reads no geometry arrays at all (face_divergence: 24 of 88
    bytes per cell).  Returns (geoms, parts) unchanged if any part would need more than MAX_CODES values in total."""
    if not any(s_ for s_, _ in parts):
        return geoms, parts
    folded, total = [], 0
    for s_, M in parts:
        M = sp.csr_matrix(M)
        if s_:
            g = np.asarray(geoms[s_ - 1], dtype=np.float64)
            if geom_by_col:
                data = M.data * g[M.indices]
            else:
                data = M.data * np.repeat(g, np.diff(M.indptr))
            M = sp.csr_matrix((data, M.indices, M.indptr), shape=M.shape)
        # distinct values of this part, without a full sort where they are few
        sample = np.unique(M.data[: min(M.data.size, 1 << 18)])
        if sample.size > MAX_CODES:
            return geoms, parts
        vals, _ = _unique_small(M.data)
        total += vals.size
        if total > MAX_CODES:
            return geoms, parts
        folded.append((0, M))
    return [], folded


def encode_parts_raw(geoms, parts, shape):
    """Like ``encode_parts`` for parts that hold DUPLICATE (row, column) pairs on purpose (ghost columns redirected by
    ``tree_shard``): entries are concatenated row by row, never summed."""
    tab_mult, tab_sel = [], []
    rows, cols, codes = [], [], []
    for s_, M in parts:
        M = sp.csr_matrix(M)
        vals, inv = _group_values(M.data)
        base = len(tab_mult)
        tab_mult.extend(vals.tolist())
        tab_sel.extend([s_] * vals.size)
        if len(tab_mult) > MAX_CODES:
            raise CodeTableOverflow(f"operator needs more than {MAX_CODES} distinct (multiplier, selector) pairs")
        rows.append(np.repeat(np.arange(M.shape[0]), np.diff(M.indptr)))
        cols.append(M.indices)
        codes.append(inv + base)
    rows, cols, codes = np.concatenate(rows), np.concatenate(cols), np.concatenate(codes)
    order = np.lexsort((cols, rows))
    indptr = np.r_[0, np.cumsum(np.bincount(rows, minlength=shape[0]))].astype(np.int64)
    return (indptr, cols[order].astype(np.int32), codes[order].astype(np.uint8), np.array(tab_mult, dtype=np.float64),
            np.array(tab_sel, dtype=np.uint8))


def encode_parts(geoms, parts, shape):
    """Merge ``sum_s diag(geom_s) @ M_s`` into one CSR with a code byte per entry.

    Returns (indptr, indices int32, codes uint8, tab_mult, tab_sel).  The parts of every reference
    operator live in disjoint column blocks (x / y / z faces or edges), so the merge is a sparse sum of
    "code matrices"; if two parts ever shared an entry the sum would be caught by the nnz check."""
    tab_mult, tab_sel = [], []
    acc = None
    for s_, M in parts:
        M = sp.csr_matrix(M)
        vals, inv = _group_values(M.data)
        base = len(tab_mult)
        tab_mult.extend(vals.tolist())
        tab_sel.extend([s_] * vals.size)
        if len(tab_mult) > MAX_CODES:
            raise CodeTableOverflow(f"operator needs more than {MAX_CODES} distinct (multiplier, selector) pairs")
        code_m = sp.csr_matrix(((inv + base + 1).astype(np.float64), M.indices, M.indptr), shape=M.shape)
        acc = code_m if acc is None else acc + code_m
    acc = sp.csr_matrix(acc)
    if acc.nnz != sum(sp.csr_matrix(M).nnz for _, M in parts):
        raise ValueError("operator parts overlap; cannot encode one code per entry")
    acc.sort_indices()
    codes = (np.rint(acc.data).astype(np.int64) - 1).astype(np.uint8)
    return (acc.indptr.astype(np.int64), acc.indices.astype(np.int32), codes, np.array(tab_mult, dtype=np.float64),
            np.array(tab_sel, dtype=np.uint8))


class CodedCsr:
    """Device-resident coded CSR (one direction of an operator)."""

    def __init__(self, geoms, parts, shape, geom_by_col, canonical=True, sort_rows=False, fold=True):
        if fold:
            geoms, parts = fold_geometry(geoms, parts, geom_by_col)
        indptr, indices, codes, tmult, tsel = (encode_parts if canonical else encode_parts_raw)(geoms, parts, shape)
        self.shape = shape
        self.nnz = int(indptr[-1])
        self.is64 = self.nnz >= 2**31 - 1
        self._host = (indptr, indices, codes)
        self.indptr = self.indices = self.codes = None
        self.tab_mult = to_device(tmult)
        self.tab_sel = to_device(tsel, torch.uint8)
        self.ntab = int(tmult.size)
        self.geoms = [to_device(g) for g in geoms] + [None] * (3 - len(geoms))
        self.geom_by_col = int(geom_by_col)
        # layout (ncu, profiles/r1b_k12_*: thread-per-row CSR is L1-bound on strided index loads):
        #   "ell"    every row has the same length (all cell-rowed operators of a balanced OcTree): column-major
        #            entries, no row pointers, perfectly coalesced;
        #   "stream" anything else: CSR whose entries a warp stages through shared memory with coalesced loads;
        #   "lpr"    the first kernel (LPR lanes per row straight from global memory), kept for comparison.
        lens = np.diff(indptr)
        self.row_len, self.padded = -1, 0
        if lens.size and self.ntab <= 255:
            kmax, avg_len = int(lens.max()), float(lens.mean())
            # one common length (all cell-rowed operators), or mildly variable rows (edge_curl: 4..6) padded
            if kmax > 0 and (lens.min() == kmax or (kmax <= 12 and kmax <= 1.5 * avg_len)):
                self.row_len = kmax
            self.padded = int(lens.min() != kmax)
        # rows of different lengths (transposed operators, nodal_gradient): sliced ELL; "lpr" / "stream" (CSR) stay
        # selectable for comparison (tools/sweep_tree.py)
        self.layout = "ell" if self.row_len > 0 else "lpr"
        # ragged rows (transposed operators, nodal_gradient): HYBRID = ELL of width K0 on the first K0 entries of every row
        # + a small CSR of what the few longer rows (parent faces / edges / nodes with hanging children) have beyond K0.
        # K0 minimises  padded ELL slots + 3 x overflow entries  (an overflow entry costs a row index, a row pointer and
        # uncoalesced loads); used when the overflow stays below 30 % of the entries.
        self.hyb_K = 0
        self.hyb = None
        if self.row_len <= 0 and lens.size and self.ntab <= 255 and lens.max() > 0:
            hist = np.bincount(lens)
            ks = np.arange(1, hist.size)
            longer = np.array([(hist[k + 1:] * (np.arange(k + 1, hist.size) - k)).sum() for k in ks])
            cost = lens.size * ks + 3 * longer
            k0 = int(ks[np.argmin(cost)])
            # measured on the 10 M-cell config-3 tree (tools/bench_tree.py) against the CSR kernel: D.T 320 -> 283 us,
            # C.T 404 -> 344, G 300 -> 285, aveF2CC.T 320 -> 281, aveE2CC.T 429 -> 397 (K0 = 2 or 4); the node-rowed
            # transposes (K0 = 6, 8: G.T 264 -> 301, aveN2CC.T 285 -> 359) lose, so wide rows stay on the CSR kernel
            if longer[k0 - 1] <= 0.3 * max(self.nnz, 1):
                self.hyb_K = k0
                if k0 <= 4:
                    self.layout = "hyb"
        if self.layout == "lpr" and lens.size and self.ntab <= 255:
            # sliced ELL only where its slices are nearly uniform: on the 10 M-cell config-3 tree every transposed
            # operator has a long row (a parent face / edge with hanging children) in most slices of 32 -- 50..87 %
            # padding, and the CSR kernel is 20 % faster there (tools/bench_tree.py)
            lp = np.zeros((lens.size + 31) // 32 * 32, dtype=np.int64)
            lp[:lens.size] = lens
            padded_entries = int(lp.reshape(-1, 32).max(axis=1).sum()) * 32
            if padded_entries <= 1.1 * max(self.nnz, 1):
                self.layout = "sell"
        avg = self.nnz / max(shape[0], 1)
        self.lpr = 1 if avg <= 7 else 2 if avg <= 16 else 4 if avg <= 40 else 16
        self.ell_indices = self.ell_codes = None
        self.sell_ptr = self.sell_indices = self.sell_codes = None
        self.sell_padding = 0.0
        # rows stored in column-locality order (ELL / SELL): see dzb_tree_ell_spmv_perm
        self.sort_rows = bool(sort_rows)
        self.ell_perm = None
        self.ell_geoms = self.geoms
        self.use_layout(self.layout)

    def use_layout(self, layout):
        """Select (and upload on first use) one of the device layouts: "ell" (constant row length only), "hyb" (ELL +
        overflow rows), "sell" (sliced ELL, any rows), "stream", "lpr"."""
        indptr, indices, codes = self._host
        if layout == "ell":
            if self.row_len <= 0:
                raise ValueError("the ELL layout needs rows of one common length")
            if self.ell_indices is None:
                n, K = self.shape[0], self.row_len
                self.stride = (n + 31) // 32 * 32
                ie = np.zeros((K, self.stride), dtype=np.int32)
                ce = np.full((K, self.stride), 255, dtype=np.uint8)      # 255 = padding slot
                lens = np.diff(indptr)
                rows = np.repeat(np.arange(n), lens)
                slot = np.arange(indices.size) - np.repeat(indptr[:-1], lens)
                ie[slot, rows] = indices
                ce[slot, rows] = codes
                if self.sort_rows and n > 1 and not self.geom_by_col:
                    first = np.where(ce[:, :n] != 255, ie[:, :n], np.iinfo(np.int32).max).min(axis=0)
                    order = np.argsort(first, kind="stable")
                    ie[:, :n] = ie[:, :n][:, order]
                    ce[:, :n] = ce[:, :n][:, order]
                    self.ell_perm = to_device(order.astype(np.int32), torch.int32)
                    dorder = to_device(order.astype(np.int64), torch.int64)
                    self.ell_geoms = [None if g is None else g.index_select(0, dorder) for g in self.geoms]
                self.ell_indices = to_device(ie.reshape(-1), torch.int32)
                self.ell_codes = to_device(ce.reshape(-1), torch.uint8)
        elif layout == "hyb":
            if self.hyb_K <= 0:
                raise ValueError("the hybrid layout needs a width K0 (ragged rows with a short overflow)")
            if self.hyb is None:
                n, K = self.shape[0], self.hyb_K
                self.stride = (n + 31) // 32 * 32
                lens = np.diff(indptr)
                rows = np.repeat(np.arange(n), lens)
                slot = np.arange(indices.size) - np.repeat(indptr[:-1], lens)
                head = slot < K
                ie = np.zeros((K, self.stride), dtype=np.int32)
                ce = np.full((K, self.stride), 255, dtype=np.uint8)      # 255 = padding slot
                ie[slot[head], rows[head]] = indices[head]
                ce[slot[head], rows[head]] = codes[head]
                over = np.nonzero(lens > K)[0]
                rptr = np.r_[0, np.cumsum(lens[over] - K)].astype(np.int64)
                self.hyb = {
                    "indices": to_device(ie.reshape(-1), torch.int32), "codes": to_device(ce.reshape(-1), torch.uint8),
                    "rows": to_device(over.astype(np.int32), torch.int32), "nrem": int(over.size),
                    "ptr": to_device(rptr.astype(np.int32), torch.int32),
                    "rem_indices": to_device(indices[~head], torch.int32), "rem_codes": to_device(codes[~head], torch.uint8),
                    "padded": int((lens < K).any()),
                }
        elif layout == "sell":
            if self.sell_indices is None:
                n = self.shape[0]
                lens = np.diff(indptr)
                order = None
                if self.sort_rows and n > 1:
                    first = np.full(n, np.iinfo(np.int32).max, dtype=np.int64)
                    has = lens > 0
                    first[has] = indices[indptr[:-1][has]]          # canonical CSR: the smallest column comes first
                    order = np.argsort(first, kind="stable")
                l = lens if order is None else lens[order]
                starts = indptr[:-1] if order is None else indptr[:-1][order]
                nsl = (n + 31) // 32
                lp = np.zeros(nsl * 32, dtype=np.int64)
                lp[:n] = l
                K = lp.reshape(nsl, 32).max(axis=1)
                sptr = np.r_[0, np.cumsum(K * 32)].astype(np.int64)
                total = int(sptr[-1])
                ie = np.zeros(total, dtype=np.int32)
                ce = np.full(total, 255, dtype=np.uint8)           # 255 = padding slot
                new_row = np.repeat(np.arange(n, dtype=np.int64), l)
                slot = np.arange(int(l.sum()), dtype=np.int64) - np.repeat(np.cumsum(l) - l, l)
                src = np.repeat(starts, l) + slot
                dst = sptr[new_row >> 5] + slot * 32 + (new_row & 31)
                ie[dst] = indices[src]
                ce[dst] = codes[src]
                self.sell_padding = total / max(self.nnz, 1) - 1.0
                self.sell_is64 = total >= 2**31 - 1
                self.sell_ptr = to_device(sptr if self.sell_is64 else sptr.astype(np.int32),
                                          torch.int64 if self.sell_is64 else torch.int32)
                self.sell_indices = to_device(ie, torch.int32)
                self.sell_codes = to_device(ce, torch.uint8)
                if order is not None:
                    self.ell_perm = to_device(order.astype(np.int32), torch.int32)
                    if not self.geom_by_col:
                        dorder = to_device(order.astype(np.int64), torch.int64)
                        self.ell_geoms = [None if g is None else g.index_select(0, dorder) for g in self.geoms]
        elif layout in ("stream", "lpr"):
            if self.indices is None:
                self.indptr = to_device(indptr if self.is64 else indptr.astype(np.int32),
                                        torch.int64 if self.is64 else torch.int32)
                self.indices = to_device(indices, torch.int32)
                self.codes = to_device(codes, torch.uint8)
        else:
            raise ValueError(f"unknown layout {layout!r}")
        self.layout = layout

    def apply(self, x, out):
        g = self.geoms
        if self.layout == "ell":
            g = self.ell_geoms
            check(load().dzb_tree_ell_spmv_perm(self.shape[0], self.row_len, self.stride, ptr(self.ell_indices),
                                                ptr(self.ell_codes), ptr(self.tab_mult), ptr(self.tab_sel), self.ntab,
                                                ptr(g[0]), ptr(g[1]), ptr(g[2]), self.geom_by_col, self.padded,
                                                ptr(self.ell_perm), ptr(x), ptr(out), stream_ptr()), "dzb_tree_ell_spmv")
            return
        if self.layout == "hyb":
            h = self.hyb
            check(load().dzb_tree_ell_spmv_perm(self.shape[0], self.hyb_K, self.stride, ptr(h["indices"]), ptr(h["codes"]),
                                                ptr(self.tab_mult), ptr(self.tab_sel), self.ntab, ptr(g[0]), ptr(g[1]),
                                                ptr(g[2]), self.geom_by_col, 1, None, ptr(x), ptr(out), stream_ptr()),
                  "dzb_tree_ell_spmv")
            if h["nrem"]:
                check(load().dzb_tree_rows_add(h["nrem"], ptr(h["rows"]), ptr(h["ptr"]), 0, ptr(h["rem_indices"]),
                                               ptr(h["rem_codes"]), ptr(self.tab_mult), ptr(self.tab_sel), self.ntab,
                                               ptr(g[0]), ptr(g[1]), ptr(g[2]), self.geom_by_col, ptr(x), ptr(out),
                                               stream_ptr()), "dzb_tree_rows_add")
            return
        if self.layout == "sell":
            g = self.ell_geoms
            check(load().dzb_tree_sell_spmv(self.shape[0], ptr(self.sell_ptr), int(self.sell_is64), ptr(self.sell_indices),
                                            ptr(self.sell_codes), ptr(self.tab_mult), ptr(self.tab_sel), self.ntab,
                                            ptr(g[0]), ptr(g[1]), ptr(g[2]), self.geom_by_col, ptr(self.ell_perm),
                                            ptr(x), ptr(out), stream_ptr()), "dzb_tree_sell_spmv")
            return
        lpr = 0 if self.layout == "stream" else self.lpr
        check(load().dzb_tree_spmv(self.shape[0], ptr(self.indptr), int(self.is64), ptr(self.indices), ptr(self.codes),
                                   ptr(self.tab_mult), ptr(self.tab_sel), self.ntab, ptr(g[0]), ptr(g[1]), ptr(g[2]),
                                   self.geom_by_col, ptr(x), ptr(out), lpr, stream_ptr()), "dzb_tree_spmv")

    def bytes_per_apply(self):
        """Algorithmic bytes: vectors + int32 columns + code bytes + indptr + row geometry."""
        n_geom = sum(g is not None for g in self.geoms)
        gsize = (self.shape[1] if self.geom_by_col else self.shape[0]) * n_geom
        ptr_bytes = 0 if self.layout in ("ell", "sell", "hyb") else (8 if self.is64 else 4) * (self.shape[0] + 1)
        return 8 * (self.shape[0] + self.shape[1]) + 5 * self.nnz + ptr_bytes + 8 * gsize   # padding slots not counted


# cell-rowed operators whose ELL rows are stored in column-locality order (measured on the 10 M-cell config-3 tree,
# tools/sweep_tree.py: cells follow the tree's Z-order, their faces / edges the Cantor key)
SORTED_ROW_OPERATORS = {"face_divergence", "average_face_to_cell", "average_edge_to_cell", "average_node_to_cell"}


class _StoredCsr:
    """Stored-value fallback with the interface TreeOperator needs of a CodedCsr."""

    layout = "stored"

    def __init__(self, dev_csr):
        self.m = dev_csr
        self.shape, self.nnz = dev_csr.shape, dev_csr.nnz

    def apply(self, x, out):
        self.m.apply(x, out)

    def bytes_per_apply(self):
        return 8 * (self.shape[0] + self.shape[1]) + 12 * self.nnz + 4 * (self.shape[0] + 1)


class TreeOperator(Operator):
    """A deflated OcTree operator (reference: tree_ext.pyx:3114-4762), applied by tree_spmv."""

    def __init__(self, mesh, name, transposed=False, _shared=None):
        M = mesh._host_parts(name)[1][0][1]
        shape = M.shape
        super().__init__((shape[1], shape[0]) if transposed else shape)
        self.mesh, self.name, self.transposed = mesh, name, transposed
        self._fwd_shape = shape
        self._dev = _shared if _shared is not None else {}

    def _csr(self, transposed):
        key = "T" if transposed else "N"
        if key not in self._dev:
            geoms, parts = self.mesh._host_parts(self.name)
            if transposed:
                parts = [(s_, M.T.tocsr()) for s_, M in parts]
                shape = (self._fwd_shape[1], self._fwd_shape[0])
            else:
                shape = self._fwd_shape
            try:
                label = self.name + (".T" if transposed else "")
                self._dev[key] = CodedCsr(geoms, parts, shape, geom_by_col=transposed,
                                          sort_rows=label in SORTED_ROW_OPERATORS)
            except CodeTableOverflow:
                # geometry-dependent weights (the distance weights of average_cell_to_face on non-uniform base widths):
                # apply the reference's matrix with stored values, 12 B per entry instead of 5
                from .csr import _DeviceCsr

                M = self.mesh._host_matrix(self.name)
                self._dev[key] = _StoredCsr(_DeviceCsr(M.T.tocsr() if transposed else M))
        return self._dev[key]

    def _dev_apply(self, x, out):
        self._csr(self.transposed).apply(x, out)

    def _dev_apply_t(self, x, out):
        self._csr(not self.transposed).apply(x, out)

    def _transpose(self):
        return TreeOperator(self.mesh, self.name, not self.transposed, self._dev)

    _adjoint = _transpose

    def _tocsr(self):
        M = self.mesh._host_matrix(self.name)
        return M.T.tocsr() if self.transposed else M.copy()

    def bytes_per_apply(self):
        return self._csr(self.transposed).bytes_per_apply()

    def __repr__(self):
        t = ".T" if self.transposed else ""
        return f"<{self.shape[0]}x{self.shape[1]} TreeOperator {self.name}{t} (CUDA, coded CSR)>"


class TreeMassDiag(Operator):
    """Diagonal mass matrix on a TreeMesh: M = diag(dim * Av^T (V sigma)) (isotropic) or
    diag(AvV^T (V sigma_cc)) (anisotropic), base/base_tensor_mesh.py:792-863 with the tree averages."""

    def __init__(self, mesh, proj, model, invert_model, invert_matrix):
        from .tensor_ops import _Model
        from . import _lib

        n = mesh.n_faces if proj == "F" else mesh.n_edges
        super().__init__((n, n))
        md = _Model(mesh, model)
        if md.kind == _lib.MODEL_TENSOR:
            raise NotImplementedError("full-tensor mass matrices on TreeMesh are not built yet (SURVEY T5 / K13)")
        dev = cuda_device()
        nC = mesh.n_cells
        V = to_device(mesh.cell_volumes)
        if md.kind == _lib.MODEL_SCALAR:
            m = torch.full((nC,), md.scalar, dtype=torch.float64, device=dev)
        else:
            m = md.dev
        if invert_model:
            m = 1.0 / m
        base = "average_face_to_cell" if proj == "F" else "average_edge_to_cell"
        if m.numel() == nC:
            Av = mesh._operator(base)
            d = Av.apply_device(V * m, transpose=True)
            d.mul_(3.0)
        else:
            Av = mesh._operator(base + "_vector")
            d = Av.apply_device(V.repeat(3) * m, transpose=True)
        self.invert_matrix = bool(invert_matrix)
        self.diag = d

    def _dev_apply(self, x, out):
        check(load().dzb_diag_scale(self.shape[0], ptr(self.diag), ptr(x), ptr(out), int(self.invert_matrix), stream_ptr()),
              "dzb_diag_scale")

    _dev_apply_t = _dev_apply

    def _transpose(self):
        return self

    _adjoint = _transpose

    def diagonal(self):
        d = self.diag.cpu().numpy()
        return 1.0 / d if self.invert_matrix else d

    def _tocsr(self):
        return sp.diags(self.diagonal(), format="csr")


def _tree_corner_projections(mesh, proj):
    """The 8 corner projection matrices P_n (3nC x n), deflated: tree_ext.pyx:5651-5763 (_getFaceP /
    _getEdgeP) with the corner tables of operators/inner_products.py:247-269.  Corner n = (a,b,c):
    faces  fx(a), fy(b), fz(c);  edges  ex(b,c), ey(a,c), ez(a,b)."""
    t = mesh._t()
    nC = t.n_cells
    fams = t.faces if proj == "F" else t.edges
    cell_ents = t.cell_faces if proj == "F" else t.cell_edges
    off = np.r_[0, np.cumsum([f.n_total for f in fams])]
    R = mesh._deflate_faces() if proj == "F" else mesh._deflate_edges()
    rows = np.arange(3 * nC)
    Ps = []
    for c in (0, 1):
        for b in (0, 1):
            for a in (0, 1):
                if proj == "F":
                    local = (a, b, c)                         # which of the 2 faces per direction
                else:
                    local = (b + 2 * c, a + 2 * c, a + 2 * b)  # which of the 4 edges per direction
                cols = np.concatenate([off[d] + fams[d].index[cell_ents[d][:, local[d]]] for d in range(3)])
                P = sp.csr_matrix((np.ones(3 * nC), (rows, cols)), shape=(3 * nC, off[3]))
                Ps.append((P @ R).tocsr())
    return Ps


def _tree_cell_entities(mesh, proj):
    """(ent int32 [slots, nC] on the device, R, R^T as device CSR, n_total) for the matrix-free tensor mass kernels:
    the TOTAL entity numbers of every cell in the slot order of dzb_tree_mass_tensor, and the deflation matrix."""
    key = "TMT_" + proj
    if key not in mesh._cache:
        from .csr import _DeviceCsr

        t = mesh._t()
        fams = t.faces if proj == "F" else t.edges
        cell_ents = t.cell_faces if proj == "F" else t.cell_edges
        off = np.r_[0, np.cumsum([f.n_total for f in fams])]
        per = 2 if proj == "F" else 4
        ent = np.stack([off[d] + fams[d].index[cell_ents[d][:, l]] for d in range(3) for l in range(per)])
        R = (mesh._deflate_faces() if proj == "F" else mesh._deflate_edges()).tocsr()
        mesh._cache[key] = (to_device(ent.astype(np.int32), torch.int32), _DeviceCsr(R), _DeviceCsr(R.T.tocsr()), int(off[3]))
    return mesh._cache[key]


def _invert_tensor6(t):
    """Closed-form inverse of the symmetric 3x3 tensors [xx, yy, zz, xy, xz, yz] (utils/matrix_utils.py:1246-1434)."""
    a11, a22, a33, a12, a13, a23 = (t[q] for q in range(6))
    det = (a11 * a22 * a33 + a12 * a23 * a13 + a13 * a12 * a23 - a13 * a22 * a13 - a12 * a12 * a33 - a11 * a23 * a23)
    return torch.stack([(a22 * a33 - a23 * a23) / det, (a11 * a33 - a13 * a13) / det, (a11 * a22 - a12 * a12) / det,
                        -(a12 * a33 - a13 * a23) / det, (a12 * a23 - a13 * a22) / det, -(a11 * a23 - a13 * a12) / det])


class TreeMassTensor(Operator):
    """get_face / get_edge_inner_product with a full-tensor model on an OcTree, matrix-free (kernel K13t):
    M x = R^T A(Sigma) (R x) -- deflation, one cell loop with the 8 corner contractions, inflation.  Symmetric."""

    def __init__(self, mesh, proj, model_dev6, invert_model):
        n = mesh.n_faces if proj == "F" else mesh.n_edges
        super().__init__((n, n))
        self.mesh, self.proj, self.invert_model = mesh, proj, bool(invert_model)
        self.ent, self.R, self.RT, self.n_total = _tree_cell_entities(mesh, proj)
        self.model_raw = model_dev6                                   # [6, nC]: xx yy zz xy xz yz
        self.model = _invert_tensor6(model_dev6) if invert_model else model_dev6
        self.model = self.model.contiguous()
        self.vol = to_device(mesh.cell_volumes)

    def apply_total(self, model6, x, out):
        """out = R^T A(model6) R x"""
        dev = x.device
        xt = torch.empty(self.n_total, dtype=torch.float64, device=dev)
        yt = torch.zeros(self.n_total, dtype=torch.float64, device=dev)
        self.R.apply(x, xt)
        check(load().dzb_tree_mass_tensor(self.mesh.n_cells, int(self.proj == "E"), ptr(self.ent), ptr(model6), ptr(self.vol),
                                          ptr(xt), ptr(yt), stream_ptr()), "dzb_tree_mass_tensor")
        self.RT.apply(yt, out)

    def _dev_apply(self, x, out):
        self.apply_total(self.model, x, out)

    _dev_apply_t = _dev_apply

    def _transpose(self):
        return self

    _adjoint = _transpose

    def _tocsr(self):
        host = self.model_raw.cpu().numpy().T                       # (nC, 6)
        return _tree_tensor_mass_matrix(self.mesh, self.proj, host.reshape(-1, order="F"), self.invert_model)


class TreeMassTensorDeriv(Operator):
    """F(v) = d(M(m) v)/dm for a full-tensor model on an OcTree (n x 6 nC), matrix-free: forward = the mass kernel with dm
    as the model, adjoint = the cell-wise contraction kernel (operators/inner_products.py:459-565)."""

    def __init__(self, mesh, proj, v):
        n = mesh.n_faces if proj == "F" else mesh.n_edges
        super().__init__((n, 6 * mesh.n_cells))
        self.mesh, self.proj = mesh, proj
        self.ent, self.R, self.RT, self.n_total = _tree_cell_entities(mesh, proj)
        self.v = to_device(v).reshape(-1).clone()
        self.vol = to_device(mesh.cell_volumes)
        self.vt = torch.empty(self.n_total, dtype=torch.float64, device=self.v.device)
        self.R.apply(self.v, self.vt)

    def _dev_apply(self, x, out):          # x = dm (6 nC, columns xx yy zz xy xz yz in F order = [6, nC] row-major)
        yt = torch.zeros(self.n_total, dtype=torch.float64, device=x.device)
        check(load().dzb_tree_mass_tensor(self.mesh.n_cells, int(self.proj == "E"), ptr(self.ent), ptr(x), ptr(self.vol),
                                          ptr(self.vt), ptr(yt), stream_ptr()), "dzb_tree_mass_tensor")
        self.RT.apply(yt, out)

    def _dev_apply_t(self, x, out):        # x = w (n) -> 6 nC
        wt = torch.empty(self.n_total, dtype=torch.float64, device=x.device)
        self.R.apply(x, wt)
        check(load().dzb_tree_mass_tensor_adj(self.mesh.n_cells, int(self.proj == "E"), ptr(self.ent), ptr(self.vol),
                                              ptr(self.vt), ptr(wt), ptr(out), stream_ptr()), "dzb_tree_mass_tensor_adj")

    def _tocsr(self):
        return _tree_tensor_mass_deriv_matrix(self.mesh, self.proj, self.v.cpu().numpy())


def _tree_tensor_mass_matrix(mesh, proj, model, invert_model):
    """M = sum_n P_n^T sqrt(V/8) Sigma sqrt(V/8) P_n  (operators/inner_products.py:147-272), host assembly."""
    nC = mesh.n_cells
    t = np.asarray(model, dtype=np.float64).reshape((nC, 6), order="F")
    if invert_model:
        a11, a22, a33, a12, a13, a23 = (t[:, q] for q in range(6))
        det = (a11 * a22 * a33 + a12 * a23 * a13 + a13 * a12 * a23 - a13 * a22 * a13 - a12 * a12 * a33 - a11 * a23 * a23)
        t = np.stack([(a22 * a33 - a23 * a23) / det, (a11 * a33 - a13 * a13) / det, (a11 * a22 - a12 * a12) / det,
                      -(a12 * a33 - a13 * a23) / det, (a12 * a23 - a13 * a22) / det, -(a11 * a23 - a13 * a12) / det], axis=1)
    d = lambda v: sp.diags(v)
    Sigma = sp.bmat([[d(t[:, 0]), d(t[:, 3]), d(t[:, 4])],
                     [d(t[:, 3]), d(t[:, 1]), d(t[:, 5])],
                     [d(t[:, 4]), d(t[:, 5]), d(t[:, 2])]]).tocsr()
    sq = sp.diags(np.tile(np.sqrt(0.125 * mesh.cell_volumes), 3))
    acc = None
    for P in _tree_corner_projections(mesh, proj):
        VP = sq @ P
        term = VP.T @ Sigma @ VP
        acc = term if acc is None else acc + term
    acc = sp.csr_matrix(acc)
    acc.sum_duplicates()
    acc.sort_indices()
    return acc


def _tree_tensor_mass_deriv_matrix(mesh, proj, v):
    """d(M(m) v)/dm for a full-tensor model on a TreeMesh: n x 6 nC (operators/inner_products.py:459-565, 3-D tensorType 3,
    with the OcTree corner projections of tree_ext.pyx:5651-5763), host assembly.  With VP_n = sqrt(V/8) P_n and
    (y1, y2, y3) = VP_n v the six column blocks are  VP_n^T [y1;0;0], [0;y2;0], [0;0;y3], [y2;y1;0], [y3;0;y1], [0;y3;y2]."""
    nC = mesh.n_cells
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    sq = sp.diags(np.tile(np.sqrt(0.125 * mesh.cell_volumes), 3))
    d = lambda a: sp.diags(a)
    Z = sp.csr_matrix((nC, nC))
    blocks = [None] * 6
    for P in _tree_corner_projections(mesh, proj):
        VP = (sq @ P).tocsr()
        Y = VP @ v
        y1, y2, y3 = Y[:nC], Y[nC:2 * nC], Y[2 * nC:]
        stacks = [sp.vstack((d(y1), Z, Z)), sp.vstack((Z, d(y2), Z)), sp.vstack((Z, Z, d(y3))),
                  sp.vstack((d(y2), d(y1), Z)), sp.vstack((d(y3), Z, d(y1))), sp.vstack((Z, d(y3), d(y2)))]
        for q in range(6):
            term = VP.T @ stacks[q]
            blocks[q] = term if blocks[q] is None else blocks[q] + term
    out = sp.hstack(blocks).tocsr()
    out.sum_duplicates()
    out.sort_indices()
    return out


class TreeMassDerivDiag(Operator):
    """F(v) = d(M(m) v)/dm for diagonal (scalar / isotropic / anisotropic) models on a TreeMesh
    (base/base_tensor_mesh.py:865-988 with the tree averages): F(v) = diag(v r) (k Av^T) diag(V c)."""

    def __init__(self, mesh, proj, md, v, invert_model, invert_matrix, ncols, scalar_mode):
        from . import _lib

        n = mesh.n_faces if proj == "F" else mesh.n_edges
        super().__init__((n, ncols))
        dev = cuda_device()
        nC = mesh.n_cells
        self.scalar_mode = scalar_mode
        base = "average_face_to_cell" if proj == "F" else "average_edge_to_cell"
        aniso = md.kind == _lib.MODEL_ANISO
        self.Av = mesh._operator(base + ("_vector" if aniso else ""))
        self.k = 1.0 if aniso else 3.0
        V = to_device(mesh.cell_volumes)
        m = torch.full((nC,), md.scalar, dtype=torch.float64, device=dev) if md.kind == _lib.MODEL_SCALAR else md.dev
        self.colscale = V.repeat(3) if aniso else V.clone()
        if invert_model:
            self.colscale = self.colscale * ((1.0 if invert_matrix else -1.0) / (m * m))
        self.rowscale = to_device(v).reshape(-1).clone()
        if invert_matrix:
            MI = TreeMassDiag(mesh, proj, md.scalar if md.kind == _lib.MODEL_SCALAR else md.dev, invert_model, False).diag
            mi2 = 1.0 / (MI * MI)
            self.rowscale = self.rowscale * (mi2 if invert_model else -mi2)
        self.nC = nC

    def _dev_apply(self, x, out):
        if self.scalar_mode == "broadcast":
            x = x.expand(self.nC)
        tmp = self.Av.apply_device(self.colscale * x, transpose=True)
        torch.mul(tmp, self.rowscale, out=out)
        out.mul_(self.k)

    def _dev_apply_t(self, x, out):
        tmp = self.Av.apply_device(self.rowscale * x, transpose=False)
        tmp.mul_(self.colscale).mul_(self.k)
        if self.scalar_mode == "broadcast":
            out.copy_(tmp.sum().reshape(1))
        else:
            out.copy_(tmp)

    def _tocsr(self):
        Av = self.Av.tocsr()
        M = sp.diags(self.rowscale.cpu().numpy()) @ (self.k * Av.T) @ sp.diags(self.colscale.cpu().numpy())
        if self.scalar_mode == "broadcast":
            M = sp.csr_matrix(M @ np.ones((self.nC, 1)))
        return sp.csr_matrix(M)


def make_tree_mass_operator(mesh, proj, model, invert_model, invert_matrix):
    from . import _lib
    from .csr import CsrOperator
    from .tensor_ops import _Model

    if proj not in ("F", "E"):
        raise TypeError("projection_type must be 'F' for faces or 'E' for edges")
    md = _Model(mesh, model)
    if md.kind == _lib.MODEL_TENSOR:
        if invert_matrix:
            raise Exception("Solver needed to invert A.")
        return TreeMassTensor(mesh, proj, md.dev.reshape(6, mesh.n_cells), invert_model)      # md.dev: F-ordered columns
    return TreeMassDiag(mesh, proj, model, invert_model, invert_matrix)


def make_tree_mass_deriv(mesh, proj, model, do_fast, invert_model, invert_matrix):
    from . import _lib
    from .tensor_ops import _Model

    md = _Model(mesh, model)
    if md.tt == -1:
        return None
    if md.kind == _lib.MODEL_TENSOR:
        if invert_model or invert_matrix:
            raise NotImplementedError(
                "inverting the property or the matrix is not yet implemented for this mesh/tensorType. You should write it!")
        def tensor_deriv(v=None):
            if v is None:
                raise Exception("v must be supplied for this implementation.")
            return TreeMassTensorDeriv(mesh, proj, v)

        return tensor_deriv
    nC = mesh.n_cells
    if md.tt == 0:
        if invert_model and not invert_matrix:
            raise TypeError("Vector must be a numpy array")  # the reference raises here, too
        ncols, mode = (nC, None) if (invert_matrix and not invert_model) else (1, "broadcast")
    else:
        ncols, mode = md.n_params, None

    def inner_product_deriv(v=None):
        if v is None:
            raise Exception("v must be supplied for this implementation.")
        return TreeMassDerivDiag(mesh, proj, md, v, invert_model, invert_matrix, ncols, mode)

    return inner_product_deriv

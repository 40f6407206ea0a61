"""TreeMesh operators partitioned across the GPUs of one box ("TreeMesh" row).

One process per GPU.  Every rank holds the (replicated, deterministic) host tree; the *vectors* and the
operator rows are distributed:

* cells: contiguous ranges of the tree's cell ordering (depth-first Z-order, reference
  ``tree.cpp:748-756, 949-953``), so a rank's cells are a spatially compact blob;
* faces / edges / nodes (non-hanging, i.e. the reference's vector entries): owned by the rank of the
  lowest-indexed cell they touch -- read off the column patterns of the averaging operators, NOT by index
  ranges (the reference numbers entities by Cantor key, ``tree.h:13-19``, which scatters neighbours);
* an operator's rows follow the ownership of its output entities.  The columns a rank's rows touch are
  renumbered locally as ``[owned entries | ghosts grouped by owner rank]`` so the local SpMV is an ordinary
  coded-CSR ``tree_spmv`` (K12) on a compact vector, and the ghosts of one peer arrive as ONE contiguous
  receive.

Data path of an apply: pack (one gather) -> point-to-point ``send/recv`` with every neighbouring rank
(``batch_isend_irecv``: NCCL over NVLink on CUDA tensors, gloo on CPU tensors) -> local SpMV.  No collective.
The communication plan is computed once per operator, on the host, from the CSR pattern alone and without any
communication (each rank derives every rank's needs from the replicated pattern).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

# (row kind, column kind) of the operators that shard; kinds: C cells, F faces, E edges, N nodes,
# C3 = three stacked cell blocks (the *_vector averages)
OPERATOR_KINDS = {
    "face_divergence": ("C", "F"),
    "edge_curl": ("F", "E"),
    "nodal_gradient": ("E", "N"),
    "average_face_to_cell": ("C", "F"),
    "average_edge_to_cell": ("C", "E"),
    "average_node_to_cell": ("C", "N"),
    "average_face_to_cell_vector": ("C3", "F"),
    "average_edge_to_cell_vector": ("C3", "E"),
    "average_node_to_edge": ("E", "N"),
    "average_node_to_face": ("F", "N"),
    "average_edge_to_face": ("F", "E"),
    "average_cell_to_face": ("F", "C"),
}


def cell_cuts(n_cells: int, world: int):
    """Balanced contiguous cell ranges: rank r owns cells [cuts[r], cuts[r+1])."""
    return np.array([(n_cells * r) // world for r in range(world + 1)], dtype=np.int64)


def _first_row_per_column(A):
    """Lowest row index with a stored entry in every column of CSR ``A`` (every column must have one)."""
    Ac = sp.csc_matrix(A)
    Ac.sort_indices()
    if (np.diff(Ac.indptr) == 0).any():
        raise ValueError("an entity touches no cell; cannot assign an owner")
    return Ac.indices[Ac.indptr[:-1]]


class ExchangePlan:
    """Who sends which owned entries to whom, for one (operator, direction) on one rank."""

    def __init__(self, rank, world, n_owned, send_idx, ghost_counts):
        self.rank, self.world, self.n_owned = rank, world, int(n_owned)
        self.send_idx = send_idx              # per peer: LOCAL owned positions to send, ascending global order
        self.ghost_counts = ghost_counts      # per peer: number of ghosts received from it
        self.ghost_off = np.r_[0, np.cumsum(ghost_counts)].astype(np.int64)
        self.n_ghost = int(self.ghost_off[-1])
        self.send_off = np.r_[0, np.cumsum([s.size for s in send_idx])].astype(np.int64)
        self._send_all = None

    @property
    def n_local(self):
        return self.n_owned + self.n_ghost

    def peers(self):
        return [q for q in range(self.world) if q != self.rank and (self.send_idx[q].size or self.ghost_counts[q])]

    def send_index_tensor(self, device):
        if self._send_all is None or self._send_all.device != torch.device(device):
            cat = np.concatenate(self.send_idx) if self.world > 1 else np.zeros(0, dtype=np.int64)
            self._send_all = torch.as_tensor(cat.astype(np.int64), device=device)
        return self._send_all

    def exchange(self, buf, send_buf=None, group=None):
        """Fill the ghost tail of ``buf`` (length n_local; the owned head already holds this rank's entries).
        Returns the outstanding requests (empty when there is no peer)."""
        if self.world == 1 or not self.peers():
            return []
        idx = self.send_index_tensor(buf.device)
        if send_buf is None:
            send_buf = torch.empty(idx.numel(), dtype=buf.dtype, device=buf.device)
        torch.index_select(buf[:self.n_owned], 0, idx, out=send_buf)
        ops = []
        for q in self.peers():
            if self.send_idx[q].size:
                ops.append(dist.P2POp(dist.isend, send_buf[self.send_off[q]:self.send_off[q + 1]], q, group))
            if self.ghost_counts[q]:
                lo = self.n_owned + self.ghost_off[q]
                ops.append(dist.P2POp(dist.irecv, buf[lo:lo + self.ghost_counts[q]], q, group))
        return dist.batch_isend_irecv(ops)


class LocalOperator:
    """Host description of one rank's share of an operator: local parts / geometry and the exchange plan."""

    def __init__(self, geoms, parts, shape, plan, geom_by_col):
        self.geoms, self.parts, self.shape, self.plan, self.geom_by_col = geoms, parts, shape, plan, geom_by_col

    def host_matrix(self):
        """scipy CSR in LOCAL numbering (rows: owned outputs; columns: [owned | ghosts]) -- test / tocsr use."""
        acc = None
        for s_, M in self.parts:
            if s_ == 0:
                term = M
            elif self.geom_by_col:
                term = M @ sp.diags(self.geoms[s_ - 1])
            else:
                term = sp.diags(self.geoms[s_ - 1]) @ M
            acc = term if acc is None else acc + term
        return sp.csr_matrix(acc)


class TreeShard:
    """Ownership of cells / faces / edges / nodes of a finalized TreeMesh for ``rank`` of ``world``."""

    def __init__(self, mesh, rank=None, world=None):
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        if not 0 <= rank < world:
            raise ValueError("rank outside [0, world)")
        if world > mesh.n_cells:
            raise ValueError("more ranks than cells")
        self.mesh, self.rank, self.world = mesh, int(rank), int(world)
        nC = mesh.n_cells
        self.cuts = cell_cuts(nC, world)
        cell_owner = (np.searchsorted(self.cuts[1:], np.arange(nC), side="right")).astype(np.int32)
        self._owner = {"C": cell_owner, "C3": np.tile(cell_owner, 3)}
        self._owned = {}
        self._local_ops = {}

    # ---- ownership ---------------------------------------------------------------------------------
    def owner(self, kind):
        """int32 array: owning rank of every global entry of ``kind``."""
        if kind not in self._owner:
            name = {"F": "average_face_to_cell", "E": "average_edge_to_cell", "N": "average_node_to_cell"}[kind]
            first = _first_row_per_column(self.mesh._host_parts(name)[1][0][1])
            self._owner[kind] = self._owner["C"][first]
        return self._owner[kind]

    def owned(self, kind):
        """Ascending global indices of the entries of ``kind`` this rank owns."""
        if kind not in self._owned:
            self._owned[kind] = np.nonzero(self.owner(kind) == self.rank)[0]
        return self._owned[kind]

    def n_owned(self, kind):
        return int(self.owned(kind).size)

    def scatter(self, kind, global_vec):
        """This rank's part of a global (replicated) vector."""
        idx = self.owned(kind)
        if isinstance(global_vec, torch.Tensor):
            return global_vec[torch.as_tensor(idx, device=global_vec.device)]
        return np.asarray(global_vec)[idx]

    def gather(self, kind, local_vec, group=None):
        """All ranks' parts assembled into the global vector (collective; for checks and I/O, not the data path)."""
        n = int(self.owner(kind).size)
        out = torch.zeros(n, dtype=local_vec.dtype, device=local_vec.device)
        out[torch.as_tensor(self.owned(kind), device=local_vec.device)] = local_vec
        if self.world > 1:
            dist.all_reduce(out, group=group)   # supports are disjoint, so the sum is the assembly
        return out

    # ---- operators ------------------------------------------------------------------------------------
    def local_operator(self, name, transposed=False) -> LocalOperator:
        key = (name, bool(transposed))
        if key in self._local_ops:
            return self._local_ops[key]
        if name not in OPERATOR_KINDS:
            raise NotImplementedError(f"{name} has no sharded form")
        rk, ck = OPERATOR_KINDS[name]
        geoms, parts = self.mesh._host_parts(name)
        if transposed:
            rk, ck = ck, rk
            parts = [(s_, sp.csr_matrix(M.T)) for s_, M in parts]
        row_owner, col_owner = self.owner(rk), self.owner(ck)
        rows = self.owned(rk)
        cols_owned = self.owned(ck)
        n_cols_global = col_owner.size
        # --- the communication pattern of ALL ranks from the replicated pattern: unique (dest, column) pairs
        pat = None
        for _, M in parts:
            P = sp.csr_matrix((np.ones(M.nnz, dtype=np.int8), M.indices, M.indptr), shape=M.shape)
            pat = P if pat is None else pat + P
        pat = sp.csr_matrix(pat)
        dest = np.repeat(row_owner, np.diff(pat.indptr)).astype(np.int64)
        src = col_owner[pat.indices]
        off = dest != src
        pairs = np.unique(dest[off] * n_cols_global + pat.indices[off].astype(np.int64))   # sorted by (dest, column)
        p_dest, p_col = pairs // n_cols_global, pairs % n_cols_global
        p_src = col_owner[p_col]
        # ghosts of this rank, grouped by owner then ascending column (np.unique already sorted columns)
        mine = p_dest == self.rank
        g_cols, g_src = p_col[mine], p_src[mine]
        order = np.argsort(g_src, kind="stable")
        g_cols, g_src = g_cols[order], g_src[order]
        ghost_counts = np.bincount(g_src, minlength=self.world).astype(np.int64)
        # what this rank sends: its owned columns other ranks need, per destination, ascending column
        g2l = np.full(n_cols_global, -1, dtype=np.int64)
        g2l[cols_owned] = np.arange(cols_owned.size)
        send_idx = []
        outgoing = p_src == self.rank
        for q in range(self.world):
            sel = outgoing & (p_dest == q)
            send_idx.append(g2l[p_col[sel]])
        plan = ExchangePlan(self.rank, self.world, cols_owned.size, send_idx, ghost_counts)
        # --- local matrix: owned rows, columns renumbered [owned | ghosts]
        g2l[g_cols] = cols_owned.size + np.arange(g_cols.size)
        local_cols_global = np.r_[cols_owned, g_cols]
        lparts = []
        for s_, M in parts:
            Ml = sp.csr_matrix(M[rows])
            li = g2l[Ml.indices]
            if li.size and li.min() < 0:
                raise AssertionError("a needed column is neither owned nor a ghost")
            Ml = sp.csr_matrix((Ml.data, li, Ml.indptr), shape=(rows.size, local_cols_global.size))
            Ml.sort_indices()
            lparts.append((s_, Ml))
        if transposed:      # geometry belongs to the columns (the forward operator's rows)
            lgeoms = [np.ascontiguousarray(g[local_cols_global]) for g in geoms]
        else:
            lgeoms = [np.ascontiguousarray(g[rows]) for g in geoms]
        lo = LocalOperator(lgeoms, lparts, (rows.size, local_cols_global.size), plan, bool(transposed))
        # rows that read a ghost column: the only ones that have to wait for the exchange
        touches = np.zeros(rows.size, dtype=bool)
        for _, Ml in lparts:
            ghost_entry = Ml.indices >= cols_owned.size
            touches |= np.add.reduceat(np.r_[ghost_entry, False].astype(np.int64), Ml.indptr[:-1])[:rows.size] * (np.diff(Ml.indptr) > 0) > 0
        lo.boundary_rows = np.nonzero(touches)[0]
        self._local_ops[key] = lo
        return lo

    def operator(self, name, transposed=False, overlap=True):
        """The device operator of this rank (CUDA): ``ShardedTreeOperator``."""
        return ShardedTreeOperator(self, name, transposed, overlap)


def _graphed_mixin():
    from .slab import _GraphedStep

    return _GraphedStep


class ShardedTreeOperator(_graphed_mixin()):
    """y_owned = (A x)_owned for a distributed x: ghost exchange + local coded-CSR SpMV (K12).

    ``apply(x_owned, out=None)`` takes and returns CUDA float64 tensors holding this rank's owned entries
    (``TreeShard.scatter`` / ``gather`` convert from / to replicated global vectors).  ``step()`` applies the
    operator to the owned head of ``self.buf`` (write the input there, ``self.x_owned``) into ``self.out``;
    ``step_graphed()`` replays pack + exchange + SpMV as one CUDA graph, because at a few hundred microseconds
    per apply the Python enqueue cost of the three pieces is not negligible."""

    def __init__(self, shard: TreeShard, name, transposed=False, overlap=True):
        from .operators import cuda_device
        from .tree_ops import CodedCsr

        self.shard, self.name, self.transposed = shard, name, bool(transposed)
        lo = shard.local_operator(name, transposed)
        self.plan = lo.plan
        self.shape_local = lo.shape
        dev = cuda_device()
        self.csr = CodedCsr(lo.geoms, lo.parts, lo.shape, geom_by_col=lo.geom_by_col)
        self.buf = torch.zeros(lo.shape[1], dtype=torch.float64, device=dev)
        self.send_buf = torch.empty(int(lo.plan.send_off[-1]), dtype=torch.float64, device=dev)
        self.out = torch.empty(lo.shape[0], dtype=torch.float64, device=dev)
        # Overlap of the ghost exchange with the SpMV: `csr_early` is the local operator with every ghost column
        # redirected to column 0 (same row lengths, so the ELL layout survives; it never reads the ghost tail), run
        # while the exchange is in flight; the few rows that really touch a ghost (`bnd_rows`, a surface effect) are
        # then recomputed by `csr_bnd` and scattered over the early result.
        self.csr_early = self.csr_bnd = self.bnd_rows = None
        nb = lo.boundary_rows.size
        if overlap and lo.plan.n_ghost > 0 and 0 < nb < lo.shape[0]:
            n_owned = lo.plan.n_owned
            early = []
            for s_, M in lo.parts:
                Me = sp.csr_matrix((M.data, np.where(M.indices >= n_owned, 0, M.indices), M.indptr), shape=M.shape)
                early.append((s_, Me))
            self.csr_early = CodedCsr(lo.geoms, early, lo.shape, geom_by_col=lo.geom_by_col, canonical=False)
            br = lo.boundary_rows
            bgeoms = lo.geoms if lo.geom_by_col else [np.ascontiguousarray(g[br]) for g in lo.geoms]
            self.csr_bnd = CodedCsr(bgeoms, [(s_, sp.csr_matrix(M[br])) for s_, M in lo.parts], (nb, lo.shape[1]),
                                    geom_by_col=lo.geom_by_col)
            self.bnd_rows = torch.as_tensor(br.astype(np.int64), device=dev)
            self.out_bnd = torch.empty(nb, dtype=torch.float64, device=dev)
            self.comm_stream = torch.cuda.Stream(device=dev)

    def apply(self, x_owned, out=None, group=None):
        if out is None:
            out = self.out
        n = self.plan.n_owned
        if x_owned.numel() != n:
            raise ValueError(f"expected this rank's {n} owned entries, got {x_owned.numel()}")
        self.buf[:n].copy_(x_owned)
        if self.csr_early is not None:
            res = self.step(group)
            if out is not self.out:
                out.copy_(res)
            return out
        for r in self.plan.exchange(self.buf, self.send_buf, group):
            r.wait()
        self.csr.apply(self.buf, out)
        return out

    @property
    def x_owned(self):
        return self.buf[:self.plan.n_owned]

    def step(self, group=None):
        if self.csr_early is None:
            for r in self.plan.exchange(self.buf, self.send_buf, group):
                r.wait()
            self.csr.apply(self.buf, self.out)
            return self.out
        main = torch.cuda.current_stream()
        self.comm_stream.wait_stream(main)
        with torch.cuda.stream(self.comm_stream):
            for r in self.plan.exchange(self.buf, self.send_buf, group):
                r.wait()
        self.csr_early.apply(self.buf, self.out)              # all rows, ghost columns redirected: no ghost reads
        main.wait_stream(self.comm_stream)
        self.csr_bnd.apply(self.buf, self.out_bnd)             # the rows that need the ghosts
        self.out.index_copy_(0, self.bnd_rows, self.out_bnd)
        return self.out

    def bytes_per_apply(self):
        return self.csr.bytes_per_apply() + 16 * int(self.plan.send_off[-1])

    @property
    def T(self):
        return ShardedTreeOperator(self.shard, self.name, not self.transposed, self.csr_early is not None)

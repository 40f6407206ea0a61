"""z-slab decomposition of TensorMesh operators across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  A TensorMesh shards
naturally in z: rank p owns the cell layers ``[c0, c1)`` and the node planes ``[c0, c1)`` (the
last rank also owns the top plane ``nz``), x and y stay whole so that rows remain contiguous.
The only data-path communication of an apply is one nearest-neighbour exchange of boundary
planes (``send/recv`` pairs, no collective); static cell models exchange their halo layer once
at construction.

The local problem of a rank is itself a TensorMesh (its layers plus one halo layer per
neighbour), so the single-GPU kernels run unchanged on it; ``dzb_tm_dc_nodal_range`` restricts
the output to the owned planes and lets interior planes be computed while halos are in flight.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def slab_bounds(nz: int, world: int, rank: int):
    """Cell layers [c0, c1) owned by ``rank`` (balanced, contiguous)."""
    base, rem = divmod(nz, world)
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


class SlabLayout:
    """Index bookkeeping of one rank's z-slab (pure host logic, no device needed)."""

    def __init__(self, nx, ny, nz, rank, world):
        if world > nz:
            raise ValueError("more ranks than cell layers")
        self.nx, self.ny, self.nz, self.rank, self.world = nx, ny, nz, rank, world
        self.c0, self.c1 = slab_bounds(nz, world, rank)
        self.has_lower, self.has_upper = rank > 0, rank < world - 1
        # local cell layers: owned + one halo layer below (for the node plane c0)
        self.l0 = self.c0 - (1 if self.has_lower else 0)
        self.l1 = self.c1
        self.nz_local = self.l1 - self.l0
        # local node planes l0 .. l1 (nz_local + 1 of them); owned ones:
        self.p_own0 = self.c0 - self.l0
        self.p_own1 = self.p_own0 + (self.c1 - self.c0) + (0 if self.has_upper else 1)
        self.plane = (nx + 1) * (ny + 1)
        self.layer = nx * ny

    @property
    def n_owned_nodes(self):
        return (self.p_own1 - self.p_own0) * self.plane

    def owned_node_slice(self):
        """Slice of the GLOBAL nodal vector owned by this rank."""
        g0 = self.c0
        g1 = self.c1 + (0 if self.has_upper else 1)
        return slice(g0 * self.plane, g1 * self.plane)


def exchange_node_halos(local, layout: SlabLayout, group=None):
    """Fill the halo node planes of ``local`` (shape [(nz_local+1) * plane]) from the neighbours.

    Sends the lowest owned plane down and the highest owned plane up; receives the neighbours'
    planes into the halo slots.  Works on CPU tensors (gloo) and CUDA tensors (NCCL).
    """
    P = layout.plane
    ops = []
    if layout.has_lower:
        ops.append(dist.P2POp(dist.isend, local[layout.p_own0 * P:(layout.p_own0 + 1) * P], layout.rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, local[(layout.p_own0 - 1) * P:layout.p_own0 * P], layout.rank - 1, group))
    if layout.has_upper:
        ops.append(dist.P2POp(dist.isend, local[(layout.p_own1 - 1) * P:layout.p_own1 * P], layout.rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, local[layout.p_own1 * P:(layout.p_own1 + 1) * P], layout.rank + 1, group))
    if not ops:
        return []
    return dist.batch_isend_irecv(ops)


class _GraphedStep:
    """Mixin: replay one distributed apply (NCCL halo exchange + kernels) as a CUDA graph.

    At 8 GPUs a 512^3 slab apply is ~70 us of kernel time; enqueueing the exchange and three launches from
    Python costs more than that, so the whole step is captured once and replayed (falls back to eager
    launches if the capture is refused)."""

    _graph = None
    _graph_failed = False

    def step_graphed(self):
        if self._graph_failed:
            return self.step()
        if self._graph is None:
            try:
                self.step()                      # warm up: NCCL communicators, kernel attributes
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.step()
                self._graph = g
            except Exception as exc:             # pragma: no cover - depends on the NCCL / driver
                self._graph_failed = True
                self._graph_error = repr(exc)
                torch.cuda.synchronize()
                return self.step()
        self._graph.replay()
        return self.out


def plane_field(seed, k, n, device, log_normal=False):
    """Plane k of a synthetic global field: identical for every decomposition (keyed by k)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 1000003 + int(k))
    x = torch.randn(n, dtype=torch.float64, device=device, generator=gen)
    return torch.exp(x) if log_normal else x


def _local_cell_model(model, layout, seed, dev):
    """This rank's cell layers [l0, l1) of a cell model as a device tensor.

    ``model``: None (synthetic log-normal field keyed by the layer index, the benchmark's choice), or the GLOBAL model in
    the reference's cell order (x fastest; numpy array or torch tensor with nx*ny*nz entries) -- every rank passes the
    same array (or a memory map of it) and keeps only its own layers plus the halo layer."""
    L = layout
    if model is None:
        return torch.cat([plane_field(seed, k, L.layer, dev, log_normal=True) for k in range(L.l0, L.l1)])
    flat = model.reshape(-1) if isinstance(model, torch.Tensor) else np.asarray(model).reshape(-1)
    if flat.shape[0] != L.nx * L.ny * L.nz:
        raise ValueError(f"cell model must have {L.nx * L.ny * L.nz} entries, got {flat.shape[0]}")
    part = flat[L.l0 * L.layer:L.l1 * L.layer]
    part = part if isinstance(part, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64))
    return part.to(device=dev, dtype=torch.float64).contiguous()


def _symmetric_vector(local_size, nx, ny, nz, rank, world, dev):
    """A rank's local vector in symmetric memory (``torch.distributed._symmetric_memory``: every rank learns the device
    addresses of the others' buffers) + what the fused kernels need to read the halo planes from the neighbours:
    ``lower`` / ``upper`` = (neighbour's buffer address, its local cell layers, local halo plane, that plane's index in the
    neighbour's array) or None.  ``local_size(layout)`` = entries of a rank's local vector."""
    import torch.distributed._symmetric_memory as symm_mem

    layouts = [SlabLayout(nx, ny, nz, r, world) for r in range(world)]
    n_max = max(local_size(l) for l in layouts)              # symmetric allocations have one size on every rank
    n_max = (n_max + 1) // 2 * 2
    buf = symm_mem.empty(n_max + 8, dtype=torch.float64, device=dev)     # + 64 bytes of sync words (dzb.h)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    buf.zero_()
    torch.cuda.synchronize()
    dist.barrier()                                           # nobody signals into words that are yet to be cleared
    L = layouts[rank]
    lower = upper = None
    sync_off = n_max * 8
    lo_signal = hi_signal = 0
    if L.has_lower:
        Ll = layouts[rank - 1]
        lower = (int(hdl.buffer_ptrs[rank - 1]), Ll.nz_local, L.p_own0 - 1, Ll.p_own1 - 1)
        lo_signal = int(hdl.buffer_ptrs[rank - 1]) + sync_off + 4      # the lower neighbour's "set by my upper" word
    if L.has_upper:
        Lu = layouts[rank + 1]
        upper = (int(hdl.buffer_ptrs[rank + 1]), Lu.nz_local, L.p_own1, Lu.p_own0)
        hi_signal = int(hdl.buffer_ptrs[rank + 1]) + sync_off           # the upper neighbour's "set by my lower" word
    return {"buffer": buf, "handle": hdl, "lower": lower, "upper": upper,
            "sync": int(hdl.buffer_ptrs[rank]) + sync_off, "lo_signal": lo_signal, "hi_signal": hi_signal,
            "sync_words": buf[n_max:].view(torch.int32)}


class SlabDCNodal(_GraphedStep):
    """y = G^T Me(sigma) G phi on a z-slab decomposed TensorMesh([nx, ny, nz]) (unit cube widths).

    ``step()`` performs one distributed apply: halo exchange of phi (side stream) overlapped with
    the interior planes, then the two boundary planes.
    """

    def __init__(self, nx, ny, nz, rank, world, seed_phi=5, seed_sigma=6, h=None, sigma=None, peer_halo=False):
        """``sigma``: the global conductivity model (nC entries, reference order) or None for the synthetic benchmark
        field; ``h``: the three global width arrays (default: unit cube).  ``peer_halo``: keep phi in symmetric memory so
        that the fused kernel reads the halo planes from the neighbours' memory itself (``step_peer``: one launch per
        apply, no send / recv)."""
        import discretize_b200 as dz

        self.layout = L = SlabLayout(nx, ny, nz, rank, world)
        dev = torch.device("cuda", torch.cuda.current_device())
        hx, hy, hz = h if h is not None else (np.ones(nx) / nx, np.ones(ny) / ny, np.ones(nz) / nz)
        self.mesh = dz.TensorMesh([np.asarray(hx), np.asarray(hy), np.asarray(hz)[L.l0:L.l1]])
        # static model incl. its halo layer
        self.sigma = _local_cell_model(sigma, L, seed_sigma, dev)
        n_local = (L.nz_local + 1) * L.plane
        self.peer = None
        if peer_halo and world > 1:
            self.peer = _symmetric_vector(lambda l: (l.nz_local + 1) * l.plane, nx, ny, nz, rank, world, dev)
            self.phi = self.peer["buffer"][:n_local]
            self.phi.zero_()
        else:
            self.phi = torch.zeros(n_local, dtype=torch.float64, device=dev)
        for lp in range(L.p_own0, L.p_own1):
            self.phi[lp * L.plane:(lp + 1) * L.plane] = plane_field(seed_phi, L.l0 + lp, L.plane, dev)
        self.out = torch.zeros_like(self.phi)
        G = self.mesh.nodal_gradient
        self.op = G.T @ self.mesh.get_edge_inner_product(self.sigma) @ G
        self.comm_stream = torch.cuda.Stream(device=dev)
        self.launches_per_step = 3 if world > 1 else 1

    def step_peer(self, sync="kernel"):
        """One distributed apply as ONE launch of the fused kernel over all owned planes; its producer warps copy the two
        halo planes from the neighbours' memory over NVLink.  ``sync="kernel"``: the launches order themselves -- each
        announces its epoch to the neighbours and waits, where it needs a halo plane, for theirs (dzb.h); nothing else
        runs between two applies.  ``sync="barrier"``: a device-side barrier across all ranks in front of the launch.
        (Whoever overwrites ``phi`` afterwards orders that after the neighbours' kernels, e.g. by the all-reduce of a
        dot product.)"""
        L, P = self.layout, self.peer
        if P is None:
            return self.step()
        if sync == "barrier":
            P["handle"].barrier()
            self.op.apply_planes_peer(self.phi, self.out, L.p_own0, L.p_own1, P["lower"], P["upper"])
        else:
            self.op.apply_planes_peer(self.phi, self.out, L.p_own0, L.p_own1, P["lower"], P["upper"],
                                      sync=(P["sync"], P["lo_signal"], P["hi_signal"]))
        return self.out

    def peer_sync_timed_out(self):
        """True if a halo wait inside a kernel gave up (the neighbour never announced its epoch): results are invalid."""
        return self.peer is not None and bool(int(self.peer["sync_words"][4].item()))

    def step(self):
        L, op = self.layout, self.op
        main = torch.cuda.current_stream()
        if L.world == 1:
            op.apply_planes(self.phi, self.out, L.p_own0, L.p_own1)
            return self.out
        # halo exchange on the side stream ...
        self.comm_stream.wait_stream(main)
        with torch.cuda.stream(self.comm_stream):
            reqs = exchange_node_halos(self.phi, L)
            for r in reqs:
                r.wait()
        # ... overlapped with the planes that do not touch a halo
        lo = L.p_own0 + (1 if L.has_lower else 0)
        hi = L.p_own1 - (1 if L.has_upper else 0)
        if hi > lo:
            op.apply_planes(self.phi, self.out, lo, hi)
        main.wait_stream(self.comm_stream)
        if L.has_lower and lo > L.p_own0:
            op.apply_planes(self.phi, self.out, L.p_own0, min(lo, L.p_own1))
        if L.has_upper and hi < L.p_own1 and hi >= lo:
            op.apply_planes(self.phi, self.out, max(hi, lo), L.p_own1)
        return self.out

    def owned_output(self):
        L = self.layout
        return self.out[L.p_own0 * L.plane:L.p_own1 * L.plane]

    def set_field(self, phi_global):
        """Load this rank's owned node planes from a GLOBAL nodal vector (numpy / torch, nN entries)."""
        L = self.layout
        part = phi_global[L.owned_node_slice()]
        part = part if isinstance(part, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64))
        self.phi[L.p_own0 * L.plane:L.p_own1 * L.plane].copy_(part)

    def matvec_global(self, phi_global):
        """y = A phi for a global nodal vector every rank holds: returns (slice of the global result, owned values)."""
        self.set_field(phi_global)
        self.step()
        return self.layout.owned_node_slice(), self.owned_output()

    def step_from_host(self, host_in, host_out, nchunks=16):
        """One distributed apply with this rank's owned planes in PINNED HOST memory (``host_in`` -> ``host_out``, both
        ``n_owned_nodes`` long): the two outermost owned planes go up first so that the halo exchange can start,
        then z-chunks stream through -- H2D of chunk c+1, kernel on chunk c, D2H of chunk c-1 overlap."""
        L, op, P = self.layout, self.op, self.layout.plane
        if getattr(self, "_host_streams", None) is None:
            dev = self.phi.device
            self._host_streams = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        s_in, s_out = self._host_streams
        main = torch.cuda.current_stream()
        o0, o1 = L.p_own0, L.p_own1
        nchunks = max(1, min(nchunks, o1 - o0))
        cuts = [o0 + round(c * (o1 - o0) / nchunks) for c in range(nchunks + 1)]

        def h2d(a, b):   # local planes [a, b)
            self.phi[a * P:b * P].copy_(host_in[(a - o0) * P:(b - o0) * P], non_blocking=True)

        s_in.wait_stream(main)
        copied = []
        with torch.cuda.stream(s_in):
            h2d(o0, o0 + 1)
            if o1 - 1 > o0:
                h2d(o1 - 1, o1)
            edge = torch.cuda.Event()
            edge.record(s_in)
        if L.world > 1:
            self.comm_stream.wait_event(edge)
            with torch.cuda.stream(self.comm_stream):
                for r in exchange_node_halos(self.phi, L):
                    r.wait()
        with torch.cuda.stream(s_in):
            for c in range(nchunks):
                h2d(cuts[c], cuts[c + 1])
                ev = torch.cuda.Event()
                ev.record(s_in)
                copied.append(ev)
        s_out.wait_stream(main)
        for c in range(nchunks):
            main.wait_event(copied[min(c + 1, nchunks - 1)])     # chunk c reads the first plane of chunk c + 1
            if L.world > 1 and (c == 0 or c == nchunks - 1):
                main.wait_stream(self.comm_stream)               # the halo planes border the first / last chunk
            op.apply_planes(self.phi, self.out, cuts[c], cuts[c + 1])
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                host_out[(cuts[c] - o0) * P:(cuts[c + 1] - o0) * P].copy_(self.out[cuts[c] * P:cuts[c + 1] * P],
                                                                          non_blocking=True)
        main.wait_stream(s_out)
        return host_out


# ---------------------------------------------------------------------------------------------------
# Maxwell:  y = C^T Mf(mui) C e + Me(sigma) e   (BASELINE config 5)
# ---------------------------------------------------------------------------------------------------
class EdgeSlab:
    """Offsets of the local stacked edge vector [Ex; Ey; Ez] of a rank's sub-mesh."""

    def __init__(self, layout: SlabLayout):
        L = layout
        nx, ny, nzl = L.nx, L.ny, L.nz_local
        self.px = nx * (ny + 1)            # Ex entries per node plane
        self.py = (nx + 1) * ny            # Ey entries per node plane
        self.pz = (nx + 1) * (ny + 1)      # Ez entries per cell layer
        self.off_y = self.px * (nzl + 1)
        self.off_z = self.off_y + self.py * (nzl + 1)
        self.n = self.off_z + self.pz * nzl

    def ex(self, v, k):
        return v[k * self.px:(k + 1) * self.px]

    def ey(self, v, k):
        return v[self.off_y + k * self.py:self.off_y + (k + 1) * self.py]

    def ez(self, v, k):
        return v[self.off_z + k * self.pz:self.off_z + (k + 1) * self.pz]


def exchange_edge_halos(local, layout: SlabLayout, es: EdgeSlab, group=None):
    """Halo exchange of an edge field: Ex, Ey plane c0 go down; Ex, Ey plane c1-1 and Ez layer c1-1 go up."""
    L = layout
    ops = []
    if L.has_lower:
        lo = L.p_own0
        ops += [dist.P2POp(dist.isend, es.ex(local, lo), L.rank - 1, group),
                dist.P2POp(dist.isend, es.ey(local, lo), L.rank - 1, group),
                dist.P2POp(dist.irecv, es.ex(local, lo - 1), L.rank - 1, group),
                dist.P2POp(dist.irecv, es.ey(local, lo - 1), L.rank - 1, group),
                dist.P2POp(dist.irecv, es.ez(local, lo - 1), L.rank - 1, group)]
    if L.has_upper:
        hi = L.p_own1 - 1
        ops += [dist.P2POp(dist.isend, es.ex(local, hi), L.rank + 1, group),
                dist.P2POp(dist.isend, es.ey(local, hi), L.rank + 1, group),
                dist.P2POp(dist.isend, es.ez(local, hi), L.rank + 1, group),
                dist.P2POp(dist.irecv, es.ex(local, hi + 1), L.rank + 1, group),
                dist.P2POp(dist.irecv, es.ey(local, hi + 1), L.rank + 1, group)]
    return dist.batch_isend_irecv(ops) if ops else []


class SlabMaxwell(_GraphedStep):
    """y = C^T Mf(mui) C e + Me(sigma) e on a z-slab decomposed TensorMesh (fused kernel K9 per rank)."""

    def __init__(self, nx, ny, nz, rank, world, seed_e=11, seed_sigma=12, seed_mui=13, h=None, sigma=None, mui=None,
                 peer_halo=False):
        """``sigma`` / ``mui``: the global cell models (nC entries, reference order) or None for the synthetic benchmark
        fields; ``h``: the three global width arrays (default: unit cube).  ``peer_halo``: allocate the edge vector in
        symmetric memory (``torch.distributed._symmetric_memory``) so that the fused kernel fetches the halo planes from
        the neighbours' memory itself (``step_peer``: one launch per apply, no send / recv)."""
        import discretize_b200 as dz

        self.layout = L = SlabLayout(nx, ny, nz, rank, world)
        dev = torch.device("cuda", torch.cuda.current_device())
        hx, hy, hz = h if h is not None else (np.ones(nx) / nx, np.ones(ny) / ny, np.ones(nz) / nz)
        self.mesh = dz.TensorMesh([np.asarray(hx), np.asarray(hy), np.asarray(hz)[L.l0:L.l1]])
        self.es = es = EdgeSlab(L)
        self.sigma = _local_cell_model(sigma, L, seed_sigma, dev)
        self.mui = _local_cell_model(mui, L, seed_mui, dev)
        self.peer = None
        if peer_halo and world > 1:
            self.peer = self._setup_peer(nx, ny, nz, rank, world, dev)
            self.e = self.peer["buffer"][:es.n]
            self.e.zero_()
        else:
            self.e = torch.zeros(es.n, dtype=torch.float64, device=dev)
        n_own_layers = L.c1 - L.c0
        for lp in range(L.p_own0, L.p_own1):      # owned node planes: Ex, Ey
            g = L.l0 + lp
            es.ex(self.e, lp).copy_(plane_field(seed_e, 3 * g, es.px, dev))
            es.ey(self.e, lp).copy_(plane_field(seed_e, 3 * g + 1, es.py, dev))
        for lp in range(L.p_own0, L.p_own0 + n_own_layers):   # owned cell layers: Ez
            es.ez(self.e, lp).copy_(plane_field(seed_e, 3 * (L.l0 + lp) + 2, es.pz, dev))
        self.out = torch.zeros_like(self.e)
        Cm = self.mesh.edge_curl
        self.op = Cm.T @ self.mesh.get_face_inner_product(self.mui) @ Cm + self.mesh.get_edge_inner_product(self.sigma)
        assert type(self.op).__name__ == "MaxwellOperator"
        self.comm_stream = torch.cuda.Stream(device=dev)
        self.launches_per_step = 2 if world > 1 else 1
        self.n_owned_edges = ((L.p_own1 - L.p_own0) * (es.px + es.py) + n_own_layers * es.pz)

    def _setup_peer(self, nx, ny, nz, rank, world, dev):
        """Edge vector in symmetric memory + the neighbours' addresses and slab layouts."""
        return _symmetric_vector(lambda l: EdgeSlab(l).n, nx, ny, nz, rank, world, dev)

    def step_peer(self, sync="kernel"):
        """One distributed apply as ONE launch of the fused kernel over all owned planes, which reads the halo planes from
        the neighbours' memory over NVLink.  ``sync="kernel"``: the launches order themselves (each announces its epoch to
        the neighbours and waits, before its first copy out of a neighbour's memory, for that neighbour's announcement:
        dzb.h) -- nothing else runs between two applies; ``sync="barrier"``: a device-side barrier across all ranks in
        front of the launch.  (Whoever overwrites ``e`` afterwards orders that after the neighbours' kernels, e.g. by the
        all-reduce of a dot product.)"""
        L, op, P = self.layout, self.op, self.peer
        if P is None:
            return self.step()
        if sync == "barrier":
            P["handle"].barrier()
            op.apply_planes_peer(self.e, self.out, L.p_own0, L.p_own1, P["lower"], P["upper"])
        else:
            op.apply_planes_peer(self.e, self.out, L.p_own0, L.p_own1, P["lower"], P["upper"],
                                 sync=(P["sync"], P["lo_signal"], P["hi_signal"]))
        return self.out

    def peer_sync_timed_out(self):
        """True if a halo wait inside a kernel gave up (the neighbour never announced its epoch): results are invalid."""
        return self.peer is not None and bool(int(self.peer["sync_words"][4].item()))

    def step(self):
        L, op = self.layout, self.op
        main = torch.cuda.current_stream()
        if L.world == 1:
            op.apply_planes(self.e, self.out, L.p_own0, L.p_own1)
            return self.out
        self.comm_stream.wait_stream(main)
        with torch.cuda.stream(self.comm_stream):
            for r in exchange_edge_halos(self.e, L, self.es):
                r.wait()
        lo = L.p_own0 + (1 if L.has_lower else 0)
        hi = L.p_own1 - (1 if L.has_upper else 0)
        if hi > lo:
            op.apply_planes(self.e, self.out, lo, hi)
        main.wait_stream(self.comm_stream)
        # the planes next to a halo, once it has arrived: both ends in ONE launch
        a0, a1 = (L.p_own0, min(lo, L.p_own1)) if (L.has_lower and lo > L.p_own0) else (0, 0)
        b0, b1 = (max(hi, lo), L.p_own1) if (L.has_upper and hi < L.p_own1 and hi >= lo) else (0, 0)
        if a1 > a0 or b1 > b0:
            op.apply_planes2(self.e, self.out, a0, a1, b0, b1)
        return self.out

    def owned_parts(self, v):
        """(Ex planes, Ey planes, Ez layers) owned by this rank, as views of a local edge vector."""
        L, es = self.layout, self.es
        n_own_layers = L.c1 - L.c0
        return (v[L.p_own0 * es.px:L.p_own1 * es.px],
                v[es.off_y + L.p_own0 * es.py:es.off_y + L.p_own1 * es.py],
                v[es.off_z + L.p_own0 * es.pz:es.off_z + (L.p_own0 + n_own_layers) * es.pz])

    # --- global vectors -----------------------------------------------------------------------------------
    def owned_global_slices(self):
        """Where this rank's (Ex planes, Ey planes, Ez layers) sit in the GLOBAL stacked edge vector [Ex; Ey; Ez]."""
        L = self.layout
        nx, ny, nz = L.nx, L.ny, L.nz
        px, py, pz = nx * (ny + 1), (nx + 1) * ny, (nx + 1) * (ny + 1)
        offy, offz = px * (nz + 1), px * (nz + 1) + py * (nz + 1)
        g0, g1 = L.c0, L.c1 + (0 if L.has_upper else 1)
        return (slice(g0 * px, g1 * px), slice(offy + g0 * py, offy + g1 * py), slice(offz + L.c0 * pz, offz + L.c1 * pz))

    def set_field(self, e_global):
        """Load this rank's owned edges from a GLOBAL edge vector (numpy / torch, nE entries)."""
        for dst, sl in zip(self.owned_parts(self.e), self.owned_global_slices()):
            part = e_global[sl]
            part = part if isinstance(part, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64))
            dst.copy_(part)

    def matvec_global(self, e_global):
        """y = A e for a global edge vector every rank holds: returns (three slices of the global result, owned values)."""
        self.set_field(e_global)
        self.step()
        return self.owned_global_slices(), self.owned_parts(self.out)

    # --- end to end from / to pinned host memory ------------------------------------------------------------
    def owned_sizes(self):
        """Lengths of the three owned parts (Ex planes, Ey planes, Ez layers) -- the layout of the host buffers."""
        L, es = self.layout, self.es
        npl, nly = L.p_own1 - L.p_own0, L.c1 - L.c0
        return npl * es.px, npl * es.py, nly * es.pz

    def step_from_host(self, host_in, host_out, nchunks=16):
        """One distributed apply with this rank's owned edges in PINNED HOST memory.

        ``host_in`` / ``host_out`` hold ``[Ex owned planes | Ey owned planes | Ez owned layers]`` (``owned_sizes``).  The
        planes the neighbours need go up first so that the halo exchange can start, then z-chunks stream through:
        H2D of chunk c+1, the fused kernel on chunk c and D2H of chunk c-1 overlap (PCIe is full duplex)."""
        L, op, es = self.layout, self.op, self.es
        if getattr(self, "_host_streams", None) is None:
            dev = self.e.device
            self._host_streams = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        s_in, s_out = self._host_streams
        main = torch.cuda.current_stream()
        o0, o1 = L.p_own0, L.p_own1
        nly = L.c1 - L.c0                       # owned Ez layers: local layers o0 .. o0 + nly - 1
        sx, sy, _ = self.owned_sizes()
        hx_in, hy_in, hz_in = host_in[:sx], host_in[sx:sx + sy], host_in[sx + sy:]
        hx_out, hy_out, hz_out = host_out[:sx], host_out[sx:sx + sy], host_out[sx + sy:]
        nchunks = max(1, min(nchunks, o1 - o0))
        cuts = [o0 + round(c * (o1 - o0) / nchunks) for c in range(nchunks + 1)]

        def h2d(a, b):   # local node planes [a, b) of Ex, Ey and cell layers [a, min(b, top)) of Ez
            self.e[a * es.px:b * es.px].copy_(hx_in[(a - o0) * es.px:(b - o0) * es.px], non_blocking=True)
            self.e[es.off_y + a * es.py:es.off_y + b * es.py].copy_(hy_in[(a - o0) * es.py:(b - o0) * es.py], non_blocking=True)
            bz = min(b, o0 + nly)
            if bz > a:
                self.e[es.off_z + a * es.pz:es.off_z + bz * es.pz].copy_(hz_in[(a - o0) * es.pz:(bz - o0) * es.pz],
                                                                          non_blocking=True)

        def d2h(a, b):
            hx_out[(a - o0) * es.px:(b - o0) * es.px].copy_(self.out[a * es.px:b * es.px], non_blocking=True)
            hy_out[(a - o0) * es.py:(b - o0) * es.py].copy_(self.out[es.off_y + a * es.py:es.off_y + b * es.py], non_blocking=True)
            bz = min(b, o0 + nly)
            if bz > a:
                hz_out[(a - o0) * es.pz:(bz - o0) * es.pz].copy_(self.out[es.off_z + a * es.pz:es.off_z + bz * es.pz],
                                                                  non_blocking=True)

        s_in.wait_stream(main)
        copied = []
        with torch.cuda.stream(s_in):
            if L.world > 1:
                h2d(o0, o0 + 1)
                if o1 - 1 > o0:
                    h2d(o1 - 1, o1)
                edge = torch.cuda.Event()
                edge.record(s_in)
        if L.world > 1:
            self.comm_stream.wait_event(edge)
            with torch.cuda.stream(self.comm_stream):
                for r in exchange_edge_halos(self.e, L, es):
                    r.wait()
        with torch.cuda.stream(s_in):
            for c in range(nchunks):
                h2d(cuts[c], cuts[c + 1])
                ev = torch.cuda.Event()
                ev.record(s_in)
                copied.append(ev)
        s_out.wait_stream(main)
        for c in range(nchunks):
            main.wait_event(copied[min(c + 1, nchunks - 1)])     # chunk c reads the first plane of chunk c + 1
            if L.world > 1 and (c == 0 or c == nchunks - 1):
                main.wait_stream(self.comm_stream)               # the halo planes border the first / last chunk
            op.apply_planes(self.e, self.out, cuts[c], cuts[c + 1])
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                d2h(cuts[c], cuts[c + 1])
        main.wait_stream(s_out)
        return host_out

    def owned_to_host(self, v, host):
        """Copy the owned parts of a local edge vector into a host buffer laid out like ``owned_sizes``."""
        sx, sy, _ = self.owned_sizes()
        px, py, pz = self.owned_parts(v)
        host[:sx].copy_(px)
        host[sx:sx + sy].copy_(py)
        host[sx + sy:].copy_(pz)
        return host

"""CPU (gloo, world_size 2 and 3): the z-slab bookkeeping and halo exchange of discretize_b200.slab."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from discretize_b200.slab import SlabLayout, exchange_node_halos, slab_bounds


def test_bounds_partition_everything():
    for nz in (7, 64, 513):
        for world in (1, 2, 3, 8):
            if world > nz:
                continue
            cuts = [slab_bounds(nz, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == nz
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_layout_covers_all_planes_once():
    nx, ny, nz = 5, 4, 11
    for world in (1, 2, 4):
        owned = np.zeros((nz + 1) * (nx + 1) * (ny + 1), dtype=int)
        for r in range(world):
            L = SlabLayout(nx, ny, nz, r, world)
            owned[L.owned_node_slice()] += 1
            assert L.nz_local == L.l1 - L.l0
            assert L.p_own1 - L.p_own0 == (L.c1 - L.c0) + (0 if L.has_upper else 1)
            # every owned plane has both neighbouring planes available locally (or is a mesh boundary)
            assert L.p_own0 >= (1 if L.has_lower else 0)
            assert L.p_own1 <= L.nz_local + (0 if L.has_upper else 1)
        assert (owned == 1).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, ny, nz, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = SlabLayout(nx, ny, nz, rank, world)
        P = L.plane
        glob = torch.arange((nz + 1) * P, dtype=torch.float64)  # global nodal field = its own index
        local = torch.full(((L.nz_local + 1) * P,), -1.0, dtype=torch.float64)
        for lp in range(L.p_own0, L.p_own1):
            g = L.l0 + lp
            local[lp * P:(lp + 1) * P] = glob[g * P:(g + 1) * P]
        for r in exchange_node_halos(local, L):
            r.wait()
        ok = True
        # after the exchange every local plane except a discarded outer one holds the global values
        lo = L.p_own0 - (1 if L.has_lower else 0)
        hi = L.p_own1 + (1 if L.has_upper else 0)
        for lp in range(lo, hi):
            g = L.l0 + lp
            ok = ok and bool(torch.equal(local[lp * P:(lp + 1) * P], glob[g * P:(g + 1) * P]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 6, 5, 9, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(world))
    assert all(results[r] for r in range(world)), results


def _edge_worker(rank, world, port, nx, ny, nz, q):
    from discretize_b200.slab import EdgeSlab, exchange_edge_halos

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = SlabLayout(nx, ny, nz, rank, world)
        es = EdgeSlab(L)
        # global edge field: value encodes (component, global plane / layer, index in plane)
        def gval(comp, g, n):
            return comp * 1e6 + g * 1e3 + torch.arange(n, dtype=torch.float64) * 1e-3
        local = torch.full((es.n,), -1.0, dtype=torch.float64)
        for lp in range(L.p_own0, L.p_own1):
            es.ex(local, lp).copy_(gval(1, L.l0 + lp, es.px))
            es.ey(local, lp).copy_(gval(2, L.l0 + lp, es.py))
        for lp in range(L.p_own0, L.p_own0 + (L.c1 - L.c0)):
            es.ez(local, lp).copy_(gval(3, L.l0 + lp, es.pz))
        for r in exchange_edge_halos(local, L, es):
            r.wait()
        ok = True
        if L.has_lower:
            lp = L.p_own0 - 1
            ok &= bool(torch.equal(es.ex(local, lp), gval(1, L.l0 + lp, es.px)))
            ok &= bool(torch.equal(es.ey(local, lp), gval(2, L.l0 + lp, es.py)))
            ok &= bool(torch.equal(es.ez(local, lp), gval(3, L.l0 + lp, es.pz)))
        if L.has_upper:
            lp = L.p_own1
            ok &= bool(torch.equal(es.ex(local, lp), gval(1, L.l0 + lp, es.px)))
            ok &= bool(torch.equal(es.ey(local, lp), gval(2, L.l0 + lp, es.py)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_edge_halo_exchange_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_edge_worker, args=(r, world, port, 5, 4, 7, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(world))
    assert all(results[r] for r in range(world)), results


def test_interpolation_point_ranges_partition_all_points():
    from discretize_b200.interp_shard import point_range

    for npts in (0, 1, 7, 100, 10**8 + 3):
        for world in (1, 2, 3, 8):
            cuts = [point_range(npts, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == npts
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1

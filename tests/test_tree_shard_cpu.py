"""CPU (gloo, world_size 2 and 3): ownership, local renumbering and ghost exchange of discretize_b200.tree_shard.

The local SpMV itself is the CUDA kernel (tests/test_gpu_tree.py, tools/check_tree_shard.py); here the local
host matrix stands in for it, so what is checked is everything AROUND the kernel: that owned rows x
[owned | ghost] columns, fed by the point-to-point exchange, reproduce the global operator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from discretize_b200.tree_mesh import TreeMesh
from discretize_b200.tree_shard import OPERATOR_KINDS, TreeShard, cell_cuts


def _mesh():
    m = TreeMesh([16, 16, 8], diagonal_balance=True)
    pts = np.random.default_rng(5).random((6, 3)) * np.array([0.8, 0.8, 0.8]) + 0.1
    m.refine_points(pts, -1, [1, 1], finalize=True)
    return m


def test_cuts_and_ownership_partition():
    m = _mesh()
    for world in (1, 2, 3, 5):
        cuts = cell_cuts(m.n_cells, world)
        assert cuts[0] == 0 and cuts[-1] == m.n_cells and (np.diff(cuts) > 0).all()
        counts = {k: np.zeros(n, dtype=int) for k, n in (("C", m.nC), ("F", m.nF), ("E", m.nE), ("N", m.nN))}
        for r in range(world):
            s = TreeShard(m, r, world)
            for k in counts:
                counts[k][s.owned(k)] += 1
        assert all((c == 1).all() for c in counts.values())


def test_single_process_loopback_all_operators():
    """All ranks' shards in one process; the exchange is replayed by direct copies following the plans."""
    m = _mesh()
    world = 3
    shards = [TreeShard(m, r, world) for r in range(world)]
    rng = np.random.default_rng(0)
    for name, (rk, ck) in OPERATOR_KINDS.items():
        for tr in (False, True):
            A = m._host_matrix(name)
            A = A.T.tocsr() if tr else A
            kin, kout = (rk, ck) if tr else (ck, rk)
            x = rng.standard_normal(A.shape[1])
            los = [s.local_operator(name, tr) for s in shards]
            bufs = []
            for s, lo in zip(shards, los):
                b = np.zeros(lo.shape[1])
                b[:lo.plan.n_owned] = x[s.owned(kin)]
                bufs.append(b)
            for p, lo in enumerate(los):           # p sends to q
                for q in range(world):
                    idx = lo.plan.send_idx[q]
                    if idx.size:
                        dst = los[q].plan
                        assert dst.ghost_counts[p] == idx.size
                        lo_ = dst.n_owned + dst.ghost_off[p]
                        bufs[q][lo_:lo_ + idx.size] = bufs[p][idx]
            y = np.zeros(A.shape[0])
            for s, lo, b in zip(shards, los, bufs):
                y[s.owned(kout)] = lo.host_matrix() @ b
            ref = A @ x
            assert np.abs(y - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1.0), (name, tr)
    # ghosts are a surface effect: far fewer than owned entries even on this small tree
    lo = shards[1].local_operator("edge_curl")
    assert lo.plan.n_ghost < lo.plan.n_owned


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _mesh()
        s = TreeShard(m)                       # rank / world from the process group
        ok = True
        rng = np.random.default_rng(1)
        for name, tr in (("face_divergence", False), ("edge_curl", False), ("edge_curl", True),
                         ("nodal_gradient", False), ("average_edge_to_cell", True)):
            rk, ck = OPERATOR_KINDS[name]
            kin, kout = (rk, ck) if tr else (ck, rk)
            A = m._host_matrix(name)
            A = A.T.tocsr() if tr else A
            x = rng.standard_normal(A.shape[1])          # same stream on every rank
            lo = s.local_operator(name, tr)
            buf = torch.zeros(lo.shape[1], dtype=torch.float64)
            buf[:lo.plan.n_owned] = torch.from_numpy(s.scatter(kin, x))
            for r in lo.plan.exchange(buf):
                r.wait()
            y_loc = torch.from_numpy(lo.host_matrix() @ buf.numpy())
            y = s.gather(kout, y_loc).numpy()
            ref = A @ x
            ok = ok and bool(np.abs(y - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1.0))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_operators_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(world))
    assert all(results[r] for r in range(world)), results

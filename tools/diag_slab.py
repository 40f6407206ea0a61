"""Where a distributed Maxwell apply spends its time: torchrun --nproc-per-node N tools/diag_slab.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from discretize_b200.slab import SlabMaxwell, exchange_edge_halos

world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
op = SlabMaxwell(n, n, n, rank, world)
L = op.layout


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


lo = L.p_own0 + (1 if L.has_lower else 0)
hi = L.p_own1 - (1 if L.has_upper else 0)


def interior():
    op.op.apply_planes(op.e, op.out, lo, hi)


def boundary():
    if L.has_lower:
        op.op.apply_planes(op.e, op.out, L.p_own0, lo)
    if L.has_upper:
        op.op.apply_planes(op.e, op.out, hi, L.p_own1)


def exchange():
    for r in exchange_edge_halos(op.e, L, op.es):
        r.wait()


def whole():
    op.op.apply_planes(op.e, op.out, L.p_own0, L.p_own1)


peer = SlabMaxwell(n, n, n, rank, world, peer_halo=True) if world > 1 else None
res = {"step_peer_read_one_launch": timeit(peer.step_peer) if peer is not None else 0.0, "step_eager": timeit(op.step), "step_graph": timeit(op.step_graphed), "interior": timeit(interior),
       "boundary_launches": timeit(boundary), "exchange": timeit(exchange), "one_launch_all_owned_planes": timeit(whole)}
if rank == 0:
    print({k: round(v, 4) for k, v in res.items()}, "ms; planes per rank", L.p_own1 - L.p_own0, flush=True)
torch.cuda.synchronize()
os._exit(0)

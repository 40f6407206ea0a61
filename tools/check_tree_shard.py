"""torchrun --nproc-per-node N tools/check_tree_shard.py [--pts 2000] [--time]
The partitioned OcTree operators (ghost exchange over NCCL + local tree_spmv) == the single-GPU operators."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import discretize_b200 as dz
from discretize_b200.tree_shard import OPERATOR_KINDS, TreeShard

ap = argparse.ArgumentParser()
ap.add_argument("--base", type=int, default=128)
ap.add_argument("--pts", type=int, default=2000)
ap.add_argument("--time", action="store_true")
ap.add_argument("--ops", default="face_divergence,edge_curl,nodal_gradient,average_face_to_cell,average_edge_to_cell,average_node_to_cell")
ap.add_argument("--no-transposes", action="store_true")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
m = dz.TreeMesh([a.base] * 3, diagonal_balance=True)
pts = np.random.default_rng(2024).random((a.pts, 3)) * 0.6 + 0.2
m.refine_points(pts, -1, [2, 2, 2], finalize=True)
shard = TreeShard(m, rank, world)
ok = True
for name in a.ops.split(","):
    rk, ck = OPERATOR_KINDS[name]
    for tr in ((False,) if a.no_transposes else (False, True)):
        kin, kout = (rk, ck) if tr else (ck, rk)
        t0 = time.time()
        op = shard.operator(name, tr)
        t_plan = time.time() - t0
        whole = getattr(m, name)
        whole = whole.T if tr else whole
        gen = torch.Generator(device=dev)
        gen.manual_seed(7)
        x = torch.randn(whole.shape[1], dtype=torch.float64, device=dev, generator=gen)   # same on every rank
        ref = whole.apply_device(x)
        y_loc = op.apply(shard.scatter(kin, x))
        got = shard.gather(kout, y_loc)
        err = float((got - ref).abs().max() / ref.abs().max())
        rec = {"rank": rank, "world": world, "op": name + (".T" if tr else ""), "rel_err": err, "n_owned_in": op.plan.n_owned,
               "n_ghost": op.plan.n_ghost, "peers": len(op.plan.peers()), "plan_s": round(t_plan, 2)}
        if a.time:
            for _ in range(3):
                op.apply(op.buf[:op.plan.n_owned])
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                op.apply(op.buf[:op.plan.n_owned])
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            for _ in range(3):
                op.step_graphed()
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(50):
                op.step_graphed()
            e1.record()
            torch.cuda.synchronize()
            tg = torch.tensor([e0.elapsed_time(e1) / 50], device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            err_g = float((shard.gather(kout, op.out) - ref).abs().max() / ref.abs().max())
            rec.update(ms_sharded_graph_max=float(tg), graph=op._graph is not None, rel_err_graph=err_g)
            ok = ok and err_g <= 1e-12
            e0.record()
            for _ in range(20):
                whole.apply_device(x, out=ref)
            e1.record()
            torch.cuda.synchronize()
            rec.update(ms_sharded_max=float(t), ms_single=e0.elapsed_time(e1) / 20, rows=int(whole.shape[0]))
        if rank == 0 or err > 1e-12:
            print(json.dumps(rec), flush=True)
        ok = ok and err <= 1e-12
        del op, whole
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

"""torchrun --nproc-per-node N tools/check_interp_shard.py : point-range partitioned interpolation == single GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import discretize_b200 as dz
from discretize_b200.interp_shard import ShardedInterpolation

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
mesh = dz.TensorMesh([256, 256, 256])
gen = torch.Generator(device=dev); gen.manual_seed(10)
pts = torch.rand((npts, 3), dtype=torch.float64, device=dev, generator=gen)      # same on every rank
ok = True
for loc in ("CC", "Ex", "Fz"):
    S = ShardedInterpolation(mesh, pts, loc, rank=rank, world=world)
    whole = mesh.get_interpolation_matrix(pts, loc)
    v = torch.randn(whole.shape[1], dtype=torch.float64, device=dev, generator=gen)
    x = torch.randn(npts, dtype=torch.float64, device=dev, generator=gen)
    ref, ref_t = whole.apply_device(v), whole.apply_device(x, transpose=True)
    got = S.gather_rows(S.apply(v))
    got_t = S.apply_t(x[S.r0:S.r1].contiguous())
    e1 = float((got - ref).abs().max() / ref.abs().max())
    e2 = float((got_t - ref_t).abs().max() / ref_t.abs().max())     # atomics: summation order differs
    e0, e1ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    y = torch.empty(S.r1 - S.r0, dtype=torch.float64, device=dev)
    for _ in range(2):
        S.apply(v, out=y)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        S.apply(v, out=y)
    e1ev.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1ev) / 5], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e0.record()
    for _ in range(5):
        whole.apply_device(v, out=ref)
    e1ev.record(); torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"world": world, "loc": loc, "npts": npts, "rel_err_apply": e1, "rel_err_apply_t": e2,
                          "ms_sharded_max": float(t), "ms_single": e0.elapsed_time(e1ev) / 5}), flush=True)
    ok = ok and e1 == 0.0 and e2 <= 1e-12
    del S, whole
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

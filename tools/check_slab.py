"""torchrun --nproc-per-node N tools/check_slab.py : slab-decomposed K8 == single-GPU K8 (bitwise)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from discretize_b200.slab import SlabDCNodal

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for (nx, ny, nz) in ((40, 33, 29), (70, 20, 64), (512, 64, 96)):
    rng = np.random.default_rng(1)
    h = (rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(nz) + 0.5)
    part = SlabDCNodal(nx, ny, nz, rank, world, h=h)
    for _ in range(2):
        part.step()
    torch.cuda.synchronize()
    whole = SlabDCNodal(nx, ny, nz, 0, 1, h=h)  # every rank recomputes the undecomposed answer
    whole.step()
    torch.cuda.synchronize()
    ref = whole.out[part.layout.owned_node_slice()]
    got = part.owned_output()
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"rank {rank}/{world} mesh {nx}x{ny}x{nz} owned planes {part.layout.p_own1 - part.layout.p_own0} rel err {err:.2e}", flush=True)
    ok = ok and err == 0.0
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

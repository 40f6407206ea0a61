"""torchrun --nproc-per-node 2 tools/diag_symm.py : does the fused Maxwell kernel run slower on symmetric memory?
Times the SAME local kernel (no halos involved) with the edge vector in symmetric memory and in ordinary device memory,
then the peer-read launch, on a 1024 x 1024 x 512 slab."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from discretize_b200.slab import SlabMaxwell

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
part = SlabMaxwell(n, n, n, rank, world, peer_halo=True)
L = part.layout
e_sym = part.e
e_reg = part.e.clone()
out = part.out


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for rnd in range(2):
    t_reg = timeit(lambda: part.op.apply_planes(e_reg, out, L.p_own0, L.p_own1))
    t_sym = timeit(lambda: part.op.apply_planes(e_sym, out, L.p_own0, L.p_own1))
    t_peer = timeit(lambda: part.step_peer())
    t_peer_b = timeit(lambda: part.step_peer(sync="barrier"))
    out_sym = torch.empty_like(e_sym)   # output in ordinary memory, input symmetric (already the case); now the reverse
    print(f"rank {rank} round {rnd}: local kernel, e in ordinary memory {t_reg:.3f} ms | e in symmetric memory {t_sym:.3f} ms | "
          f"peer launch (in-kernel ordering) {t_peer:.3f} ms | peer launch (barrier) {t_peer_b:.3f} ms", flush=True)
dist.barrier()
dist.destroy_process_group()

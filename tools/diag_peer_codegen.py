"""One GPU: the PEER=true instantiation of K9 without any neighbour (pure code path) against the plain kernel, and with a
'neighbour' that is a local allocation (peer maps, local reads)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from discretize_b200.slab import SlabMaxwell

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
parts = [SlabMaxwell(n, n, n, r, 2) for r in range(2)]
p = parts[1]
L = p.layout
other = parts[0]
lower = (other.e.data_ptr(), other.layout.nz_local, L.p_own0 - 1, other.layout.p_own1 - 1)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for rnd in range(2):
    t0 = timeit(lambda: p.op.apply_planes(p.e, p.out, L.p_own0, L.p_own1))
    t1 = timeit(lambda: p.op.apply_planes_peer(p.e, p.out, L.p_own0, L.p_own1, None, None))
    t2 = timeit(lambda: p.op.apply_planes_peer(p.e, p.out, L.p_own0, L.p_own1, lower, None))
    print(f"round {rnd}: plain kernel {t0:.3f} ms | PEER instantiation, no neighbour {t1:.3f} ms | PEER, lower halo from a local "
          f"'neighbour' array {t2:.3f} ms", flush=True)

"""Host-to-host apply of the fused DC operator (bench.py's e2e leg): number of z-chunks in the H2D / kernel / D2H
pipeline, next to the raw pinned-memory copy rates of the box."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mesh = dz.TensorMesh([n, n, n])
dev = torch.device("cuda")
sig = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
G = mesh.nodal_gradient
A = G.T @ mesh.get_edge_inner_product(sig) @ G
x = torch.randn(mesh.nN, dtype=torch.float64).pin_memory()
y = torch.empty(mesh.nN, dtype=torch.float64).pin_memory()
xd = torch.empty(mesh.nN, dtype=torch.float64, device=dev)


def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


gb = 8 * mesh.nN / 1e9
print(json.dumps({"h2d_ms": wall(lambda: xd.copy_(x, non_blocking=True)), "d2h_ms": wall(lambda: y.copy_(xd, non_blocking=True)),
                  "gb_each_way": gb}), flush=True)
for nch in (2, 4, 8, 16, 32, 64):
    ms = wall(lambda: A._matvec_streamed(x, y, nchunks=nch))
    print(json.dumps({"nchunks": nch, "ms": ms, "gdof_s": mesh.nN / ms / 1e6, "gb_s_each_way": gb / ms * 1e3}), flush=True)

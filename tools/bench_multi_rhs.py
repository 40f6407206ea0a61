"""Multi-RHS fused DC apply (K8): k right-hand sides in one launch against k separate applies, 512^3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mesh = dz.TensorMesh([n, n, n])
dev = torch.device("cuda")
sig = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
G = mesh.nodal_gradient
A = G.T @ mesh.get_edge_inner_product(sig) @ G
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in (1, 2, 4, 8):
    X = torch.randn((k, mesh.nN), dtype=torch.float64, device=dev)
    Y = torch.empty_like(X)

    def multi():
        A._dev_apply_multi(X, Y)

    def loop():
        for r in range(k):
            A.apply_device(X[r], out=Y[r])

    res = []
    for fn in (loop, multi):
        for _ in range(3):
            fn()
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10)
    alg = 8.0 * (2 * mesh.nN * k + mesh.nC)
    print(f"k={k}: {k} applies {res[0]:.3f} ms | one multi-RHS launch {res[1]:.3f} ms = {mesh.nN * k / res[1] / 1e6:.1f} Gdof/s, "
          f"{alg / res[1] / 1e6:.0f} GB/s of {8.0 * (2 * mesh.nN + mesh.nC / k) / 1e0 :.0f} B-per-rhs algorithmic", flush=True)

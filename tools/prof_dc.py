"""A few launches of the fused DC kernel K8 at n^3 for ncu:  python tools/prof_dc.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mesh = dz.TensorMesh([n, n, n])
dev = torch.device("cuda")
phi = torch.randn(mesh.nN, dtype=torch.float64, device=dev)
sig = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
G = mesh.nodal_gradient
A = G.T @ mesh.get_edge_inner_product(sig) @ G
out = torch.empty(mesh.nN, dtype=torch.float64, device=dev)
for _ in range(3):
    A.apply_device(phi, out=out)
torch.cuda.synchronize()
print("ok")

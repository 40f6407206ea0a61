"""One launch (after a warm-up launch) of each kernel added late in round 2, for ncu:
K10c / K11c (compact interpolation rows, 1e8 points on 512^3) + k_gather_rows, K6 (full-tensor mass, 512^3), K9c (complex
Maxwell, 512^3), K7 with inversions (marching kernels on pre-scaled vectors).  python tools/prof_round2.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import discretize_b200 as dz

dev = torch.device("cuda")
g = torch.Generator(device=dev)
g.manual_seed(10)
mesh = dz.TensorMesh([512, 512, 512])

# K10c / K11c
pts = torch.rand((100_000_000, 3), dtype=torch.float64, device=dev, generator=g)
Q = mesh.get_interpolation_matrix(pts, "CC")
v = torch.randn(Q.shape[1], dtype=torch.float64, device=dev, generator=g)
out = torch.empty(Q.shape[0], dtype=torch.float64, device=dev)
for _ in range(2):
    Q.apply_device(v, out=out)
xt = torch.randn(Q.shape[0], dtype=torch.float64, device=dev, generator=g)
ot = torch.empty(Q.shape[1], dtype=torch.float64, device=dev)
Q.apply_device(xt, transpose=True, out=ot)
del pts, Q, v, out, xt, ot
torch.cuda.empty_cache()

# K6
sig = torch.rand((mesh.nC, 6), dtype=torch.float64, device=dev, generator=g)
sig[:, :3] += 2.0
sig[:, 3:] -= 0.5
for get, n in ((mesh.get_face_inner_product, mesh.nF), (mesh.get_edge_inner_product, mesh.nE)):
    M = get(sig)
    x = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    y = torch.empty_like(x)
    for _ in range(2):
        M.apply_device(x, out=y)
    del M, x, y
del sig
torch.cuda.empty_cache()

# K9c
sigma = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev, generator=g))
mui = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev, generator=g))
C = mesh.edge_curl
A = C.T @ mesh.get_face_inner_product(mui) @ C + 1j * 2 * np.pi * 7.0 * mesh.get_edge_inner_product(sigma)
e = torch.complex(torch.randn(mesh.nE, dtype=torch.float64, device=dev, generator=g),
                  torch.randn(mesh.nE, dtype=torch.float64, device=dev, generator=g))
yc = torch.empty_like(e)
for _ in range(2):
    assert A.apply_device_complex(e, out=yc) is not None
del A, e, yc
torch.cuda.synchronize()
print("ok")

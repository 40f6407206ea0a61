"""A few bucketed interpolation applies for ncu:  python tools/prof_interp.py [npts] [loc]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
loc = sys.argv[2] if len(sys.argv) > 2 else "CC"
mesh = dz.TensorMesh([512, 512, 512])
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(10)
pts = torch.rand((npts, 3), dtype=torch.float64, device=dev, generator=g)
Q = mesh.get_interpolation_matrix(pts, loc)
v = torch.randn(Q.shape[1], dtype=torch.float64, device=dev)
out = torch.empty(npts, dtype=torch.float64, device=dev)
for _ in range(3):
    Q.apply_device(v, out=out)
x = torch.randn(npts, dtype=torch.float64, device=dev)
o2 = torch.empty(Q.shape[1], dtype=torch.float64, device=dev)
for _ in range(2):
    Q.apply_device(x, transpose=True, out=o2)
torch.cuda.synchronize()
print("ok")

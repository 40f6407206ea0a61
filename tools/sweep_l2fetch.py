"""L2 fetch granularity (cudaLimitMaxL2FetchGranularity) vs the random-gather kernels: interpolation apply /
fused / transpose at 2e7 points on 512^3, and the OcTree SpMV."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import discretize_b200 as dz
from discretize_b200.interp import make_interpolation_operator
from tools.sweep_elementary import timeit

lib = C.CDLL(os.path.join(os.path.dirname(dz.__file__), "libdzb.so"))
dev = torch.device("cuda")
torch.zeros(1, device=dev)
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
tree_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mesh = dz.TensorMesh([512] * 3)
gen = torch.Generator(device=dev); gen.manual_seed(10)
pts = torch.rand((npts, 3), dtype=torch.float64, device=dev, generator=gen)
Q = mesh.get_interpolation_matrix(pts, "Ex")
Qf = make_interpolation_operator(mesh, pts, "Ex", store=False)
v = torch.randn(Q.shape[1], dtype=torch.float64, device=dev)
w = torch.randn(npts, dtype=torch.float64, device=dev)
o1 = torch.empty(npts, dtype=torch.float64, device=dev)
o2 = torch.empty(Q.shape[1], dtype=torch.float64, device=dev)
tree_ops = []
if tree_pts:
    m = dz.TreeMesh([512, 512, 512], diagonal_balance=True)
    m.refine_points(np.random.default_rng(2024).random((tree_pts, 3)) * 0.6 + 0.2, -1, [2, 2, 2], finalize=True)
    for name in ("face_divergence", "edge_curl"):
        op = getattr(m, name)
        tree_ops.append((name, op, torch.randn(op.shape[1], dtype=torch.float64, device=dev),
                         torch.empty(op.shape[0], dtype=torch.float64, device=dev)))
for g in (128, 64, 32):
    got = lib.dzb_tune_l2_fetch(g)
    row = {"l2_fetch_requested": g, "l2_fetch_reported": got,
           "locate_ms": timeit(Q._locate, 5, 1),
           "apply_ms": timeit(lambda: Q.apply_device(v, out=o1), 5, 1),
           "fused_ms": timeit(lambda: Qf.apply_device(v, out=o1), 5, 1),
           "apply_t_ms": timeit(lambda: Q.apply_device(w, transpose=True, out=o2), 5, 1)}
    for name, op, x, y in tree_ops:
        row["tree_" + name + "_ms"] = timeit(lambda: op.apply_device(x, out=y), 20, 3)
    print(json.dumps(row), flush=True)

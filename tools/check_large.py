"""64-bit index paths: the marching kernels against the generic row kernels at 1024^3 (3.2e9 faces / edges > 2^31),
through the public operators.  Two independent CUDA implementations must agree to rounding."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz
from discretize_b200._lib import LIB_PATH, load

load()
lib = C.CDLL(LIB_PATH)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mesh = dz.TensorMesh([n, n, n])
dev = torch.device("cuda")
gen = torch.Generator(device=dev); gen.manual_seed(1)
big = torch.randn(mesh.nE, dtype=torch.float64, device=dev, generator=gen)
ok = True
sig = torch.rand(mesh.nC, dtype=torch.float64, device=dev, generator=gen) + 0.5
ops = [("D", mesh.face_divergence), ("D.T", mesh.face_divergence.T), ("G", mesh.nodal_gradient), ("G.T", mesh.nodal_gradient.T),
       ("C", mesh.edge_curl), ("C.T", mesh.edge_curl.T), ("aveE2CC", mesh.average_edge_to_cell),
       ("aveCC2F", mesh.average_cell_to_face), ("cellGrad", mesh.cell_gradient), ("Me(iso)", mesh.get_edge_inner_product(sig))]
for name, op in ops:
    x = big[:op.shape[1]]
    lib.dzb_tune_march(-1); lib.dzb_tune_mass(1)
    y1 = op.apply_device(x)
    lib.dzb_tune_march(0); lib.dzb_tune_mass(0)
    y0 = op.apply_device(x)
    lib.dzb_tune_march(-1); lib.dzb_tune_mass(1)
    err = float((y1 - y0).abs().max() / y0.abs().max())
    tail = float((y1[-1000:] - y0[-1000:]).abs().max())
    print(json.dumps({"n": n, "op": name, "rows": int(op.shape[0]), "cols": int(op.shape[1]), "rel_err_march_vs_generic": err,
                      "abs_err_last_1000_rows": tail}), flush=True)
    ok = ok and err <= 1e-13
    del y0, y1
sys.exit(0 if ok else 1)

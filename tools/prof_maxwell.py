"""A few launches of one K9 configuration, for ncu:  python tools/prof_maxwell.py n variant a b [kchunk] [launches]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz
from discretize_b200._lib import load

# python tools/prof_maxwell.py n variant xs r nslot nprod [kchunk] [launches]   (variant 1: xs = consumer warps)
n, variant, a, b, ns, np_ = (int(v) for v in sys.argv[1:7])
kc = int(sys.argv[7]) if len(sys.argv) > 7 else 0
launches = int(sys.argv[8]) if len(sys.argv) > 8 else 3
mesh = dz.TensorMesh([n, n, n])
dev = torch.device("cuda")
e = torch.randn(mesh.nE, dtype=torch.float64, device=dev)
sig = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
mui = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
C = mesh.edge_curl
A = C.T @ mesh.get_face_inner_product(mui) @ C + mesh.get_edge_inner_product(sig)
out = torch.empty(mesh.nE, dtype=torch.float64, device=dev)
lib = load()
lib.dzb_tune_maxwell.argtypes = [ctypes.c_int, ctypes.c_int]
lib.dzb_tune_maxwell.restype = None
lib.dzb_tune_maxwell_rows.argtypes = [ctypes.c_int] * 4
lib.dzb_tune_maxwell_rows.restype = None
lib.dzb_tune_maxwell(variant, kc)
lib.dzb_tune_maxwell_tmap.argtypes = [ctypes.c_int] * 3
lib.dzb_tune_maxwell_tmap.restype = None
if variant == 2:
    lib.dzb_tune_maxwell_rows(a, b, ns, np_)
else:
    lib.dzb_tune_maxwell_tmap(a, b, ns)
for _ in range(launches):
    A.apply_device(e, out=out)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))

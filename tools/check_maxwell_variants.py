"""K9 variants against the elementary chain C^T (Mf (C e)) + Me e on small / odd-shaped meshes (GPU)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import discretize_b200 as dz
from discretize_b200._lib import load

lib = load()
for name, n in (("dzb_tune_maxwell", 2), ("dzb_tune_maxwell_rows", 4), ("dzb_tune_maxwell_tmap", 3)):
    getattr(lib, name).argtypes = [ctypes.c_int] * n
    getattr(lib, name).restype = None
rng = np.random.default_rng(0)
dev = torch.device("cuda")
worst = 0.0
for shape in ((40, 24, 36), (64, 64, 64), (33, 17, 9), (70, 5, 3), (2, 2, 2), (1, 1, 1), (129, 66, 20), (96, 31, 12)):
    h = [rng.random(s) + 0.5 for s in shape]
    mesh = dz.TensorMesh(h)
    e = torch.tensor(rng.standard_normal(mesh.nE), device=dev)
    sig = torch.tensor(np.exp(rng.standard_normal(mesh.nC)), device=dev)
    mui = torch.tensor(np.exp(rng.standard_normal(mesh.nC)), device=dev)
    C = mesh.edge_curl
    Mf, Me = mesh.get_face_inner_product(mui), mesh.get_edge_inner_product(sig)
    A = C.T @ Mf @ C + Me
    ref = C.T.apply_device(Mf.apply_device(C.apply_device(e))) + Me.apply_device(e)
    for variant, cfgs in ((0, [None]), (2, [(2, 4, 4, 2), (1, 8, 4, 2)]), (3, [(1, 4, 4), (1, 5, 6), (1, 6, 4), (1, 6, 8), (1, 8, 4), (2, 4, 4), (2, 6, 6), (2, 8, 4), (4, 4, 4), (1, 8, 6)])):
        for cfg in cfgs:
            lib.dzb_tune_maxwell(variant, 0)
            if variant == 2:
                lib.dzb_tune_maxwell_rows(*cfg)
            if variant == 3:
                lib.dzb_tune_maxwell_tmap(*cfg)
            out = torch.full((mesh.nE,), float("nan"), dtype=torch.float64, device=dev)
            A.apply_device(e, out=out)
            torch.cuda.synchronize()
            err = float((out - ref).abs().max() / ref.abs().max())
            path = lib.dzb_tm_maxwell_last_path()
            worst = max(worst, err if err == err else 1.0)
            flag = "" if err < 1e-12 else "   <-- FAIL"
            print(f"{shape} variant {variant} {cfg}: path {path} rel err {err:.2e}{flag}", flush=True)
lib.dzb_tune_maxwell(3, 0)
print("worst", worst)
sys.exit(0 if worst < 1e-12 else 1)

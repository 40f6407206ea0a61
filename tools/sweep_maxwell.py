"""Tuning sweep for the fused Maxwell kernel (K9): variant / warps per block / z-chunk."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz
from discretize_b200._lib import load

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
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
lib.dzb_tune_maxwell_warps.argtypes = [ctypes.c_int]
lib.dzb_tune_maxwell_warps.restype = None
alg = 8.0 * (2 * mesh.nE + 2 * mesh.nC)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
cases = ([(0, 0, 64)] if n <= 512 else []) + [(1, w, kc) for w in ((9, 12, 18) if n <= 512 else (6, 9)) for kc in ((32, 64, 128) if n <= 512 else (32, 64, 128, 256))]
for variant, w, kc in cases:
    lib.dzb_tune_maxwell(variant, kc)
    if w:
        lib.dzb_tune_maxwell_warps(w)
    for _ in range(2):
        A.apply_device(e, out=out)
    reps = 3 if variant == 0 else 10
    e0.record()
    for _ in range(reps):
        A.apply_device(e, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if ref is None:
        ref = out.clone()
    err = float((out - ref).abs().max() / ref.abs().max())
    print(f"variant={variant} warps={w:2d} kchunk={kc:3d} err={err:.1e} {ms:.3f} ms {mesh.nE/ms/1e6:.1f} Gdof/s {alg/ms/1e6:.0f} GB/s frac {alg/ms/1e6/6541.1:.3f}", flush=True)
lib.dzb_tune_maxwell(1, 64); lib.dzb_tune_maxwell_warps(12)

"""Tuning sweep for the fused DC nodal kernel (K8) on the 512^3 workload."""
import ctypes
import sys

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz
from discretize_b200._lib import load

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mesh = dz.TensorMesh([n, n, n])
dev = torch.device("cuda")
phi = torch.randn(mesh.nN, dtype=torch.float64, device=dev)
sigma = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
G = mesh.nodal_gradient
A = G.T @ mesh.get_edge_inner_product(sigma) @ G
out = torch.empty(mesh.nN, dtype=torch.float64, device=dev)
lib = load()
lib.dzb_tune_dc.argtypes = [ctypes.c_int, ctypes.c_int]
lib.dzb_tune_dc.restype = None
alg = 8.0 * (2 * mesh.nN + mesh.nC)
# copy bandwidth reference
b = torch.empty_like(phi)
for _ in range(3):
    b.copy_(phi)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    b.copy_(phi)
e1.record(); torch.cuda.synchronize()
print("copy GB/s", 16.0 * mesh.nN / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9)
lib.dzb_tune_dc_variant.argtypes = [ctypes.c_int]
lib.dzb_tune_dc_variant.restype = None
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
ref = None
for variant, ry, kc in [(v, r, k) for v in variants for r in (2, 4, 8) for k in (16, 32, 64, 128)]:
    if True:
        lib.dzb_tune_dc_variant(variant)
        lib.dzb_tune_dc(ry, kc)
        for _ in range(3):
            A.apply_device(phi, out=out)
        e0.record()
        for _ in range(10):
            A.apply_device(phi, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if ref is None:
            ref = out.clone()
        err = float((out - ref).abs().max() / ref.abs().max())
        print(f"v={variant} err={err:.1e} ry={ry} kchunk={kc:4d}  {ms:.3f} ms  {mesh.nN/ms/1e6:.1f} Gdof/s  {alg/ms/1e6:.0f} GB/s  frac {alg/ms/1e6/6541.1:.3f}", flush=True)

"""Frequency-domain Maxwell operator C^T Mf C + i omega Me on a complex field: one launch (K9c) against the two-part path.

  python tools/bench_complex.py [n]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz
from discretize_b200._lib import load

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = load()
lib.dzb_tune_maxwell_cplx.argtypes = [ctypes.c_int, ctypes.c_int]
lib.dzb_tune_maxwell_cplx.restype = None
mesh = dz.TensorMesh([n, n, n])
dev = torch.device("cuda")
sig = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
mui = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
C = mesh.edge_curl
A = C.T @ mesh.get_face_inner_product(mui) @ C + 1j * 18.85 * mesh.get_edge_inner_product(sig)
e = torch.complex(torch.randn(mesh.nE, dtype=torch.float64, device=dev), torch.randn(mesh.nE, dtype=torch.float64, device=dev))
out = torch.empty_like(e)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
alg = 8.0 * (4 * mesh.nE + 2 * mesh.nC)        # complex in + out, the two models once


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for cfg in ((1, 6), (1, 8), (2, 4)):
    lib.dzb_tune_maxwell_cplx(*cfg)
    ms = timeit(lambda: A.apply_device_complex(e, out=out))
    print(f"K9c shape {cfg}: {ms:.3f} ms  {mesh.nE / ms / 1e6:.1f} complex Gdof/s  {alg / ms / 1e6:.0f} GB/s  frac {alg / ms / 1e6 / 6541.1:.3f}", flush=True)
lib.dzb_tune_maxwell_cplx(1, 6)
ref = out.clone()
if n <= 512:
    A.fused = False
    ms2 = timeit(lambda: A @ e, 3)
    err = float((A @ e - ref).abs().max() / ref.abs().max())
    print(f"two-part path (2 x K9 + 2 x K5 + splits): {ms2:.3f} ms, rel diff {err:.1e}", flush=True)

"""Tuning sweep for the fused Maxwell kernel (K9): variant / block shape / z-chunk.

variant 2 = row-per-warp ring fed by row-wise bulk copies, variant 3 = the same ring fed by tensor-map TMA.
  python tools/sweep_maxwell.py [n] [v2|v3|all] [debug mode]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz
from discretize_b200._lib import load

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
which = sys.argv[2] if len(sys.argv) > 2 else "all"
dbg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
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
lib.dzb_tune_maxwell_debug.argtypes = [ctypes.c_int]
lib.dzb_tune_maxwell_debug.restype = None
lib.dzb_tune_maxwell_debug(dbg)
alg = 8.0 * (2 * mesh.nE + 2 * mesh.nC)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cases = []
# reference result: the elementary chain C^T (Mf (C e)) + Me e (separate kernels, pinned against the oracle by the tests)
Mf, Me = mesh.get_face_inner_product(mui), mesh.get_edge_inner_product(sig)
ref = C.T.apply_device(Mf.apply_device(C.apply_device(e)))
ref += Me.apply_device(e)
torch.cuda.synchronize()
if which in ("v2", "all"):
    shapes = ((2, 4, 4, 1), (2, 4, 4, 2), (2, 4, 4, 4), (2, 4, 6, 1), (2, 4, 6, 2), (4, 2, 4, 1), (4, 2, 4, 2), (4, 2, 4, 4),
              (4, 4, 4, 2), (4, 4, 4, 4), (4, 4, 6, 4), (1, 8, 4, 2))
    cases += [(2, sh, 0, 0) for sh in shapes]
promo = int(os.environ.get("DZB_L2PROMO", "128"))
lib.dzb_tune_maxwell_l2promo.argtypes = [ctypes.c_int]
lib.dzb_tune_maxwell_l2promo.restype = None
lib.dzb_tune_maxwell_l2promo(promo)
lib.dzb_tune_maxwell_l2hint.argtypes = [ctypes.c_int]
lib.dzb_tune_maxwell_l2hint.restype = None
lib.dzb_tune_maxwell_l2hint(int(os.environ.get("DZB_L2HINT", "1")))
if which in ("v3", "all"):
    lib.dzb_tune_maxwell_tmap.argtypes = [ctypes.c_int] * 3
    lib.dzb_tune_maxwell_tmap.restype = None
    shapes3 = ((2, 8, 6), (1, 12, 6), (1, 16, 4), (1, 16, 6), (2, 10, 4))
    cases += [(3, sh, 0, 0) for sh in shapes3]
for variant, a, b, kc in cases:
    lib.dzb_tune_maxwell(variant, kc)
    if variant == 2:
        lib.dzb_tune_maxwell_rows(*a)
    else:
        lib.dzb_tune_maxwell_tmap(*a)
    for _ in range(2):
        A.apply_device(e, out=out)
    reps = 10
    e0.record()
    for _ in range(reps):
        A.apply_device(e, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out.fill_(float("nan"))
    A.apply_device(e, out=out)
    torch.cuda.synchronize()
    err = float((out - ref).abs().max() / ref.abs().max())
    print(f"variant={variant} shape={a} kchunk={kc:3d} err={err:.1e} {ms:.3f} ms {mesh.nE/ms/1e6:.1f} Gdof/s {alg/ms/1e6:.0f} GB/s frac {alg/ms/1e6/6541.1:.3f}", flush=True)
lib.dzb_tune_maxwell(3, 0)

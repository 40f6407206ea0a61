"""Throughput of the full-tensor mass matrices: TensorMesh K6 (512^3) and OcTree K13t (config-3 tree).

  python tools/bench_tensor_mass.py [tree refinement points, default 48000; 0 = skip the tree]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import discretize_b200 as dz

PEAK = 6541.1
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def spd6(nC, dev, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    t = torch.rand((nC, 6), dtype=torch.float64, device=dev, generator=g)
    t[:, :3] += 2.0          # diagonally dominant: symmetric positive definite
    t[:, 3:] -= 0.5
    return t


dev = torch.device("cuda")
mesh = dz.TensorMesh([512, 512, 512])
sig = spd6(mesh.nC, dev, 1)
for proj, n in (("F", mesh.nF), ("E", mesh.nE)):
    M = mesh.get_face_inner_product(sig) if proj == "F" else mesh.get_edge_inner_product(sig)
    x = torch.randn(n, dtype=torch.float64, device=dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    ms = timeit(lambda: M.apply_device(x, out=out))
    alg = 8.0 * (2 * n + 6 * mesh.nC)
    print(json.dumps({"op": f"K6 TensorMesh 512^3 M{proj.lower()}(full tensor) @ x", "dofs": n, "ms": ms, "gdof_s": n / ms / 1e6,
                      "alg_gb_s": alg / ms / 1e6, "frac_of_measured_hbm": alg / ms / 1e6 / PEAK}), flush=True)
    del M, x, out
del sig, mesh
torch.cuda.empty_cache()
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 48000
if npts:
    t0 = time.time()
    m = dz.TreeMesh([512, 512, 512], diagonal_balance=True)
    m.refine_points(np.random.default_rng(2024).random((npts, 3)) * 0.6 + 0.2, -1, [2, 2, 2], finalize=True)
    print(json.dumps({"tree_cells": m.nC, "build_s": time.time() - t0}), flush=True)
    sig = spd6(m.nC, dev, 2)
    for proj, n in (("F", m.nF), ("E", m.nE)):
        t0 = time.time()
        M = m.get_face_inner_product(sig) if proj == "F" else m.get_edge_inner_product(sig)
        x = torch.randn(n, dtype=torch.float64, device=dev)
        out = torch.empty(n, dtype=torch.float64, device=dev)
        M.apply_device(x, out=out)
        torch.cuda.synchronize()
        setup = time.time() - t0
        ms = timeit(lambda: M.apply_device(x, out=out))
        slots = 6 if proj == "F" else 12
        # vectors + model + volumes + entity table + the two deflation SpMVs (values 8 + index 4 per entry, ~1.3 per total entity)
        alg = 8.0 * (2 * n + 7 * m.nC) + 4.0 * slots * m.nC + 2 * (12.0 * M.R.nnz + 8.0 * (n + M.n_total))
        print(json.dumps({"op": f"K13t OcTree {m.nC} cells M{proj.lower()}(full tensor) @ x", "dofs": n, "ms": ms,
                          "gdof_s": n / ms / 1e6, "alg_gb_s": alg / ms / 1e6, "frac_of_measured_hbm": alg / ms / 1e6 / PEAK,
                          "setup_s": setup}), flush=True)
        F = (m.get_face_inner_product_deriv(sig) if proj == "F" else m.get_edge_inner_product_deriv(sig))(x)
        dm = torch.randn(6 * m.nC, dtype=torch.float64, device=dev)
        o1 = torch.empty(n, dtype=torch.float64, device=dev)
        o2 = torch.empty(6 * m.nC, dtype=torch.float64, device=dev)
        ms_f = timeit(lambda: F.apply_device(dm, out=o1))
        ms_t = timeit(lambda: F.apply_device(x, transpose=True, out=o2))
        print(json.dumps({"op": f"K13t OcTree dM{proj.lower()}(full tensor)(v): forward / adjoint", "ms_forward": ms_f, "ms_adjoint": ms_t}), flush=True)
        del M, F, x, out, dm, o1, o2

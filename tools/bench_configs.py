"""Throughput of the five BASELINE configurations on one GPU (Gdof/s, algorithmic GB/s, roofline fraction).

  python tools/bench_configs.py [--configs 1,3,4,5] [--tree-pts 12000] [--interp-pts 100000000] [--maxwell-n 512]
"""
import argparse
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


def timeit(fn, reps=20, warm=3):
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


def report(rows, name, ndof, nbytes, ms, extra=None):
    r = {"op": name, "dofs": int(ndof), "ms": ms, "gdof_s": ndof / ms / 1e6, "alg_gb_s": nbytes / ms / 1e6,
         "frac_of_measured_hbm": nbytes / ms / 1e6 / PEAK}
    if extra:
        r.update(extra)
    rows.append(r)
    print(json.dumps(r), flush=True)


def config1(rows):
    mesh = dz.TensorMesh([64, 64, 64])
    dev = torch.device("cuda")
    u = torch.randn(mesh.nF, dtype=torch.float64, device=dev)
    e = torch.randn(mesh.nE, dtype=torch.float64, device=dev)
    sig = torch.rand(mesh.nC, dtype=torch.float64, device=dev) + 0.5
    D, C, Mf = mesh.face_divergence, mesh.edge_curl, mesh.get_face_inner_product(sig)
    oD, oC, oM = (torch.empty(n, dtype=torch.float64, device=dev) for n in (mesh.nC, mesh.nF, mesh.nF))
    report(rows, "cfg1 64^3 face_divergence @ u (L2 resident)", mesh.nC, 8 * (mesh.nF + mesh.nC), timeit(lambda: D.apply_device(u, out=oD), 200))
    report(rows, "cfg1 64^3 edge_curl @ e (L2 resident)", mesh.nF, 8 * (mesh.nE + mesh.nF), timeit(lambda: C.apply_device(e, out=oC), 200))
    report(rows, "cfg1 64^3 Mf(sigma) @ u (L2 resident)", mesh.nF, 8 * (2 * mesh.nF + mesh.nC), timeit(lambda: Mf.apply_device(u, out=oM), 200))
    # the same elementary kernels out of cache
    mesh = dz.TensorMesh([512, 512, 512])
    u = torch.randn(mesh.nF, dtype=torch.float64, device=dev)
    oD = torch.empty(mesh.nC, dtype=torch.float64, device=dev)
    D = mesh.face_divergence
    report(rows, "512^3 face_divergence @ u", mesh.nC, 8 * (mesh.nF + mesh.nC), timeit(lambda: D.apply_device(u, out=oD), 10))
    del u, oD
    phi = torch.randn(mesh.nN, dtype=torch.float64, device=dev)
    oG = torch.empty(mesh.nE, dtype=torch.float64, device=dev)
    G = mesh.nodal_gradient
    report(rows, "512^3 nodal_gradient @ phi", mesh.nE, 8 * (mesh.nE + mesh.nN), timeit(lambda: G.apply_device(phi, out=oG), 10))
    del phi
    oC = torch.empty(mesh.nF, dtype=torch.float64, device=dev)
    C = mesh.edge_curl
    report(rows, "512^3 edge_curl @ e", mesh.nF, 8 * (mesh.nE + mesh.nF), timeit(lambda: C.apply_device(oG, out=oC), 10))


def config3(rows, npts):
    t0 = time.time()
    m = dz.TreeMesh([512, 512, 512], diagonal_balance=True)
    pts = np.random.default_rng(2024).random((npts, 3)) * 0.6 + 0.2
    m.refine_points(pts, -1, [2, 2, 2], finalize=True)
    t_build = time.time() - t0
    print(json.dumps({"tree": f"512^3 base, {npts} points, pad [2,2,2]", "nC": m.nC, "nF": m.nF, "nE": m.nE, "nN": m.nN,
                      "hanging_F_E_N": [m.nhF, m.nhE, m.nhN], "build_s": t_build}), flush=True)
    dev = torch.device("cuda")
    for name in ("face_divergence", "edge_curl", "average_face_to_cell", "average_edge_to_cell", "average_node_to_cell",
                 "nodal_gradient"):
        t0 = time.time()
        op = getattr(m, name)
        x = torch.randn(op.shape[1], dtype=torch.float64, device=dev)
        out = torch.empty(op.shape[0], dtype=torch.float64, device=dev)
        op.apply_device(x, out=out)
        torch.cuda.synchronize()
        t_op = time.time() - t0
        csr = op._csr(False)
        ms = timeit(lambda: op.apply_device(x, out=out), 20)
        report(rows, f"cfg3 OcTree {m.nC} cells {name}", op.shape[0], op.bytes_per_apply(), ms,
               {"nnz": csr.nnz, "lanes_per_row": csr.lpr, "host_assembly_s": t_op})
        m._cache.pop("P_" + name, None); m._cache.pop("H_" + name, None); m._cache.pop("OP_" + name, None)
        del op, x, out, csr


def config4(rows, npts):
    from discretize_b200.interp import make_interpolation_operator

    mesh = dz.TensorMesh([512, 512, 512])
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev); gen.manual_seed(10)
    pts = torch.rand((npts, 3), dtype=torch.float64, device=dev, generator=gen)
    for loc in ("CC", "N", "Ex", "Fz"):
        t0 = time.time()
        Q = mesh.get_interpolation_matrix(pts, loc)
        torch.cuda.synchronize()
        first = time.time() - t0
        ms_build = timeit(Q._locate, 5, 1)
        report(rows, f"cfg4 interp build (locate+weights) {loc} {npts} pts", npts, 120.0 * npts, ms_build,
               {"unit": "points", "first_call_s": first})
        v = torch.randn(Q.shape[1], dtype=torch.float64, device=dev)
        out = torch.empty(npts, dtype=torch.float64, device=dev)
        ms_apply = timeit(lambda: Q.apply_device(v, out=out), 5, 1)
        report(rows, f"cfg4 interp apply Q@v {loc}", npts, 136.0 * npts + 8.0 * 32 * npts, ms_apply,
               {"unit": "points", "bytes_note": "tables 128 B/pt + 8 B out + 8 gathers x 32 B sectors"})
        Qf = make_interpolation_operator(mesh, pts, loc, store=False)
        ms_fused = timeit(lambda: Qf.apply_device(v, out=out), 5, 1)
        report(rows, f"cfg4 interp fused locate+apply {loc}", npts, 32.0 * npts + 8.0 * 32 * npts, ms_fused,
               {"unit": "points", "bytes_note": "24 B pt + 8 B out + 8 gathers x 32 B sectors"})
        del Q, Qf, v, out


def config5(rows, n):
    mesh = dz.TensorMesh([n, n, n])
    dev = torch.device("cuda")
    e = torch.randn(mesh.nE, dtype=torch.float64, device=dev)
    sig = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
    mui = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
    C = mesh.edge_curl
    A = C.T @ mesh.get_face_inner_product(mui) @ C + mesh.get_edge_inner_product(sig)
    out = torch.empty(mesh.nE, dtype=torch.float64, device=dev)
    ms = timeit(lambda: A.apply_device(e, out=out), 5, 2)
    report(rows, f"cfg5 {n}^3 Maxwell C^T Mf C + Me fused ({type(A).__name__})", mesh.nE, 8.0 * (2 * mesh.nE + 2 * mesh.nC), ms)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,4,5,3")
    ap.add_argument("--tree-pts", type=int, default=12000)
    ap.add_argument("--interp-pts", type=int, default=100_000_000)
    ap.add_argument("--maxwell-n", type=int, default=512)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rows = []
    for c in a.configs.split(","):
        if c == "1": config1(rows)
        if c == "3": config3(rows, a.tree_pts)
        if c == "4": config4(rows, a.interp_pts)
        if c == "5": config5(rows, a.maxwell_n)
        torch.cuda.empty_cache()
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)

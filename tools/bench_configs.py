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


def cpu_ms_per_dof(A, reps=3):
    """The reference's CPU apply path: scipy csr_matvec (one thread) on the materialised matrix; ns per output row."""
    x = np.random.default_rng(0).standard_normal(A.shape[1])
    A @ x
    t0 = time.perf_counter()
    for _ in range(reps):
        A @ x
    return (time.perf_counter() - t0) / reps * 1e9 / max(A.shape[0], 1)


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
    # the same kernels out of cache: every elementary operator, mass matrix and model derivative at 512^3
    mesh = dz.TensorMesh([512, 512, 512])
    big = torch.randn(mesh.nE, dtype=torch.float64, device=dev)
    out = torch.empty(mesh.nE, dtype=torch.float64, device=dev)
    sig = torch.rand(mesh.nC, dtype=torch.float64, device=dev) + 0.5
    sig3 = torch.rand((mesh.nC, 3), dtype=torch.float64, device=dev) + 0.5
    D, G, C = mesh.face_divergence, mesh.nodal_gradient, mesh.edge_curl
    dMf = mesh.get_face_inner_product_deriv(sig)(big[:mesh.nF].clone())
    dMe3 = mesh.get_edge_inner_product_deriv(sig3)(big.clone())
    nF, nE, nC = mesh.nF, mesh.nE, mesh.nC
    ops = [("face_divergence @ u", D, None), ("face_divergence.T @ c", D.T, None), ("nodal_gradient @ phi", G, None),
           ("nodal_gradient.T @ e", G.T, None), ("edge_curl @ e", C, None), ("edge_curl.T @ u", C.T, None),
           ("average_face_to_cell @ u", mesh.average_face_to_cell, None),
           ("average_edge_to_cell @ e", mesh.average_edge_to_cell, None),
           ("average_node_to_cell @ phi", mesh.average_node_to_cell, None),
           ("average_face_to_cell_vector.T @ c3 (generic kernel)", mesh.average_face_to_cell_vector.T, None),
           ("average_cell_to_face @ c", mesh.average_cell_to_face, None),
           ("cell_gradient @ c (Neumann)", mesh.cell_gradient, None),
           ("Mf(sigma iso) @ u", mesh.get_face_inner_product(sig), 8.0 * (2 * nF + nC)),
           ("Me(sigma iso) @ e", mesh.get_edge_inner_product(sig), 8.0 * (2 * nE + nC)),
           ("Me(sigma aniso) @ e", mesh.get_edge_inner_product(sig3), 8.0 * (2 * nE + 3 * nC)),
           ("Mf(1/sigma)^-1 @ u (invert_model, invert_matrix)",
            mesh.get_face_inner_product(sig, invert_model=True, invert_matrix=True), 8.0 * (2 * nF + nC)),
           ("dMf(sigma iso)(v) @ dm", dMf, 8.0 * (2 * nF + nC)), ("dMf(sigma iso)(v).T @ w", dMf.T, 8.0 * (2 * nF + nC)),
           ("dMe(sigma aniso)(v) @ dm", dMe3, 8.0 * (2 * nE + 3 * nC)),
           ("dMe(sigma aniso)(v).T @ w", dMe3.T, 8.0 * (2 * nE + 3 * nC))]
    # CPU baseline beside every line: the same operator materialised at 96^3 (.tocsr() == the reference's matrix),
    # applied by scipy csr_matvec on one host core, scaled per output row
    small = dz.TensorMesh([96, 96, 96])
    ssig = np.random.default_rng(1).random(small.nC) + 0.5
    ssig3 = np.random.default_rng(2).random((small.nC, 3)) + 0.5
    sv = np.random.default_rng(3).standard_normal(small.nE)
    small_ops = {"face_divergence @ u": small.face_divergence, "face_divergence.T @ c": small.face_divergence.T,
                 "nodal_gradient @ phi": small.nodal_gradient, "nodal_gradient.T @ e": small.nodal_gradient.T,
                 "edge_curl @ e": small.edge_curl, "edge_curl.T @ u": small.edge_curl.T,
                 "average_face_to_cell @ u": small.average_face_to_cell,
                 "average_edge_to_cell @ e": small.average_edge_to_cell,
                 "average_node_to_cell @ phi": small.average_node_to_cell,
                 "average_cell_to_face @ c": small.average_cell_to_face,
                 "cell_gradient @ c (Neumann)": small.cell_gradient,
                 "Mf(sigma iso) @ u": small.get_face_inner_product(ssig),
                 "Me(sigma iso) @ e": small.get_edge_inner_product(ssig),
                 "Me(sigma aniso) @ e": small.get_edge_inner_product(ssig3),
                 }
    import scipy.sparse as sp
    AvF = small.average_face_to_cell.tocsr()
    # the reference's own formula for the isotropic derivative: diag(v) (3 Av^T diag(V))  (base_tensor_mesh.py:865-988)
    dM_small = (sp.diags(sv[:small.nF]) @ (3.0 * AvF.T) @ sp.diags(small.cell_volumes)).tocsr()
    small_csr = {"dMf(sigma iso)(v) @ dm": dM_small, "dMf(sigma iso)(v).T @ w": dM_small.T.tocsr()}
    for name, op, nbytes in ops:
        x, y = big[:op.shape[1]], out[:op.shape[0]]
        ms = timeit(lambda: op.apply_device(x, out=y), 10)
        extra = None
        if (name in small_ops or name in small_csr) and not os.environ.get("DZB_NO_CPU"):
            ns = cpu_ms_per_dof(small_csr[name] if name in small_csr else small_ops[name].tocsr())
            cpu_gdofs = 1.0 / ns
            extra = {"cpu_scipy_gdof_s": cpu_gdofs, "cpu_sample": "96^3, scipy csr_matvec, 1 core",
                     "gpu_over_cpu": op.shape[0] / ms / 1e6 / cpu_gdofs}
        report(rows, "512^3 " + name, op.shape[0], nbytes if nbytes else 8.0 * (op.shape[0] + op.shape[1]), ms, extra)


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
        extra = {"nnz": csr.nnz, "layout": csr.layout, "host_assembly_s": t_op}
        if not os.environ.get("DZB_NO_CPU"):
            ns = cpu_ms_per_dof(m._host_matrix(name))
            extra.update(cpu_scipy_gdof_s=1.0 / ns, cpu_sample="the same CSR, scipy csr_matvec, 1 core, full size",
                         cpu_ms=ns * op.shape[0] * 1e-6, gpu_over_cpu=ns * op.shape[0] * 1e-6 / ms)
        report(rows, f"cfg3 OcTree {m.nC} cells {name}", op.shape[0], op.bytes_per_apply(), ms, extra)
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
        extra = {"unit": "points", "bytes_note": "tables 96 B/pt + 8 B out + 8 gathers x 32 B sectors"}
        if not os.environ.get("DZB_NO_CPU") and loc == "CC":
            Qs = mesh.get_interpolation_matrix(pts[:1_000_000], loc).tocsr()
            ns = cpu_ms_per_dof(Qs)
            extra.update(cpu_scipy_gpts_s=1.0 / ns, cpu_sample="1e6 of the points, scipy csr_matvec on Q.tocsr(), 1 core",
                         gpu_over_cpu=npts / ms_apply / 1e6 * ns)
        report(rows, f"cfg4 interp apply Q@v {loc} (bucketed tables, gather un-permute)", npts, 104.0 * npts + 8.0 * 32 * npts, ms_apply, extra)
        Q.unpermute = "scatter"
        report(rows, f"cfg4 interp apply Q@v {loc} (bucketed tables, scatter)", npts, 104.0 * npts + 8.0 * 32 * npts,
               timeit(lambda: Q.apply_device(v, out=out), 5, 1))
        xt = torch.randn(npts, dtype=torch.float64, device=dev)
        ot = torch.empty(Q.shape[1], dtype=torch.float64, device=dev)
        report(rows, f"cfg4 interp apply Q.T@x {loc} (bucketed tables)", npts, 104.0 * npts + 8.0 * 32 * npts,
               timeit(lambda: Q.apply_device(xt, transpose=True, out=ot), 5, 1))
        del xt, ot
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
    extra = None
    if not os.environ.get("DZB_NO_CPU"):
        # the reference cannot assemble this operator at 1024^3 (C alone is ~206 GB): assembled at 64^3, scaled per row
        sm = dz.TensorMesh([64, 64, 64])
        rng = np.random.default_rng(5)
        Cs = sm.edge_curl.tocsr()
        As = (Cs.T @ sm.get_face_inner_product(np.exp(rng.standard_normal(sm.nC))).tocsr() @ Cs
              + sm.get_edge_inner_product(np.exp(rng.standard_normal(sm.nC))).tocsr()).tocsr()
        ns = cpu_ms_per_dof(As)
        extra = {"cpu_scipy_gdof_s": 1.0 / ns, "cpu_sample": "64^3, assembled C^T Mf C + Me, scipy csr_matvec, 1 core",
                 "gpu_over_cpu": mesh.nE / ms / 1e6 * ns}
    report(rows, f"cfg5 {n}^3 Maxwell C^T Mf C + Me fused ({type(A).__name__})", mesh.nE, 8.0 * (2 * mesh.nE + 2 * mesh.nC), ms, extra)
    # the 512^3 DC operator of config 2 (the bench.py workload) beside it
    if n == 1024:
        return
    phi = torch.randn(mesh.nN, dtype=torch.float64, device=dev)
    G = mesh.nodal_gradient
    Adc = G.T @ mesh.get_edge_inner_product(sig) @ G
    o2 = torch.empty(mesh.nN, dtype=torch.float64, device=dev)
    ms = timeit(lambda: Adc.apply_device(phi, out=o2), 10, 3)
    report(rows, f"cfg2 {n}^3 DC G^T Me G fused ({type(Adc).__name__})", mesh.nN, 8.0 * (2 * mesh.nN + mesh.nC), ms)
    D = mesh.face_divergence
    Acc = D @ mesh.get_face_inner_product(sig, invert_matrix=True) @ D.T
    xc = torch.randn(mesh.nC, dtype=torch.float64, device=dev)
    o3 = torch.empty(mesh.nC, dtype=torch.float64, device=dev)
    ms = timeit(lambda: Acc.apply_device(xc, out=o3), 10, 3)
    report(rows, f"{n}^3 cell-centred DC D Mf(rho)^-1 D^T fused, cached 1/Mf ({type(Acc).__name__})", mesh.nC,
           16.0 * mesh.nC + 8.0 * mesh.nF, ms)
    Acc.cache_inverse_mass = False
    ms = timeit(lambda: Acc.apply_device(xc, out=o3), 10, 3)
    report(rows, f"{n}^3 cell-centred DC D Mf(rho)^-1 D^T fused, masses on the fly", mesh.nC, 24.0 * mesh.nC, ms)
    MfI = mesh.get_face_inner_product(sig, invert_matrix=True)
    t1 = torch.empty(mesh.nF, dtype=torch.float64, device=dev)
    t2 = torch.empty(mesh.nF, dtype=torch.float64, device=dev)

    def chain():
        D.apply_device(xc, transpose=True, out=t1)
        MfI.apply_device(t1, out=t2)
        D.apply_device(t2, out=o3)

    ms = timeit(chain, 10, 3)
    report(rows, f"{n}^3 cell-centred DC as three applies (D, Mf^-1, D^T)", mesh.nC, 24.0 * mesh.nC, ms)


def config_solver(rows, n):
    """Jacobi-PCG on the fused DC operator: time per iteration = apply + dot + update + direction."""
    from discretize_b200.solvers import pcg

    mesh = dz.TensorMesh([n, n, n])
    dev = torch.device("cuda")
    sig = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev))
    G = mesh.nodal_gradient
    A = G.T @ mesh.get_edge_inner_product(sig) @ G
    xt = torch.randn(mesh.nN, dtype=torch.float64, device=dev)
    b = A.apply_device(xt)
    dinv = 1.0 / A.diagonal_device()
    del xt
    iters = 60
    pcg(A, b, inv_diag=dinv, rtol=0.0, maxiter=10, check_every=10)          # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    x, info = pcg(A, b, inv_diag=dinv, rtol=0.0, maxiter=iters, check_every=iters)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # per iteration: apply 24 B/node + dot 16 + update 64 (x rw, r rw, p, Ap, dinv, z) + direction 24
    report(rows, f"cfg2 {n}^3 Jacobi-PCG iteration on fused G^T Me G (apply + 3 vector kernels)", mesh.nN,
           8.0 * (2 * mesh.nN + mesh.nC) + 104.0 * mesh.nN, ms,
           {"iterations": info.iterations, "residual_drop": info.history[-1] / info.history[0]})


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
        if c == "s": config_solver(rows, 512)
        if c == "5": config5(rows, a.maxwell_n)
        torch.cuda.empty_cache()
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)

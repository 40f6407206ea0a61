"""One launch of every non-fused kernel family at out-of-cache sizes, for `ncu --set full -k regex:...`.

  python tools/prof_kernels.py [--which elem,mass,interp,tree] [--n 512] [--tree-pts 48000] [--interp-pts 20000000]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import discretize_b200 as dz


def once(op, x):
    out = torch.empty(op.shape[0], dtype=torch.float64, device=x.device)
    for _ in range(1):
        op.apply_device(x, out=out)
    torch.cuda.synchronize()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="elem,mass,interp,tree")
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--tree-pts", type=int, default=48000)
    ap.add_argument("--interp-pts", type=int, default=20_000_000)
    a = ap.parse_args()
    which = a.which.split(",")
    dev = torch.device("cuda")
    mesh = dz.TensorMesh([a.n] * 3)
    if "elem" in which:
        u = torch.randn(mesh.nF, dtype=torch.float64, device=dev)
        once(mesh.face_divergence, u)                       # k_apply<DIV>
        e = once(mesh.face_divergence.T, torch.randn(mesh.nC, dtype=torch.float64, device=dev))  # k_apply3<DIV^T>
        del e
        phi = torch.randn(mesh.nN, dtype=torch.float64, device=dev)
        e = once(mesh.nodal_gradient, phi)                  # k_apply3<GRAD>
        del phi
        once(mesh.edge_curl, e)                             # k_apply3<CURL>
        once(mesh.nodal_gradient.T, e)                      # k_apply<GRAD^T>
        once(mesh.average_edge_to_cell, e)
        del e, u
        torch.cuda.empty_cache()
    if "mass" in which:
        sig = torch.rand(mesh.nC, dtype=torch.float64, device=dev) + 0.5
        u = torch.randn(mesh.nF, dtype=torch.float64, device=dev)
        once(mesh.get_face_inner_product(sig), u)
        once(mesh.get_face_inner_product(torch.rand((mesh.nC, 3), dtype=torch.float64, device=dev) + 0.5), u)
        dM = mesh.get_face_inner_product_deriv(sig)(u)
        once(dM, torch.randn(mesh.nC, dtype=torch.float64, device=dev))
        once(dM.T, u)
        del sig, u, dM
        torch.cuda.empty_cache()
    if "interp" in which:
        from discretize_b200.interp import make_interpolation_operator

        gen = torch.Generator(device=dev)
        gen.manual_seed(10)
        pts = torch.rand((a.interp_pts, 3), dtype=torch.float64, device=dev, generator=gen)
        Q = mesh.get_interpolation_matrix(pts, "Ex")
        Q._locate()
        v = torch.randn(Q.shape[1], dtype=torch.float64, device=dev)
        once(Q, v)
        once(Q.T, torch.randn(a.interp_pts, dtype=torch.float64, device=dev))
        once(make_interpolation_operator(mesh, pts, "Ex", store=False), v)
        del Q, v, pts
        torch.cuda.empty_cache()
    if "tree" in which:
        m = dz.TreeMesh([512, 512, 512], diagonal_balance=True)
        pts = np.random.default_rng(2024).random((a.tree_pts, 3)) * 0.6 + 0.2
        m.refine_points(pts, -1, [2, 2, 2], finalize=True)
        for name in ("face_divergence", "edge_curl", "average_node_to_cell", "nodal_gradient"):
            op = getattr(m, name)
            once(op, torch.randn(op.shape[1], dtype=torch.float64, device=dev))
            m._cache.pop("P_" + name, None); m._cache.pop("H_" + name, None); m._cache.pop("OP_" + name, None)
            del op
    print("prof_kernels done")


if __name__ == "__main__":
    main()

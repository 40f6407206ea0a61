"""Launch-configuration sweep of the elementary stencil kernels (K1-K4) at 512^3, and the L2 fetch granularity
experiment for the gather kernels.  Build with DZB_EXTRA_NVCC_FLAGS=-DDZB_APPLY_SWEEP for all configurations."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz
from discretize_b200._lib import load

PEAK = 6541.1
CFGS = {0: "32x4x4 zt1", 1: "32x4x4 zt2", 2: "32x4x4 zt4", 3: "64x4x2 zt4", 4: "128x2x2 zt4", 5: "32x8x2 zt4",
        6: "64x2x2 zt4", 7: "32x4x2 zt4", 8: "32x4x4 zt8", 9: "32x2x4 zt4"}


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    cfgs = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(CFGS)
    lib = C.CDLL(os.path.join(os.path.dirname(dz.__file__), "libdzb.so"))
    load()
    mesh = dz.TensorMesh([n] * 3)
    dev = torch.device("cuda")
    names = [("D", mesh.face_divergence), ("D.T", mesh.face_divergence.T), ("G", mesh.nodal_gradient),
             ("G.T", mesh.nodal_gradient.T), ("C", mesh.edge_curl), ("C.T", mesh.edge_curl.T),
             ("aveE2CC", mesh.average_edge_to_cell), ("aveF2CCV.T", mesh.average_face_to_cell_vector.T),
             ("aveN2CC", mesh.average_node_to_cell), ("aveF2CC", mesh.average_face_to_cell), ("cellGrad", mesh.cell_gradient)]
    sig = torch.rand(mesh.nC, dtype=torch.float64, device=dev) + 0.5
    sig3 = torch.rand((mesh.nC, 3), dtype=torch.float64, device=dev) + 0.5
    mass = [("Mf(iso)", mesh.get_face_inner_product(sig), 8.0 * (2 * mesh.nF + mesh.nC)),
            ("Me(iso)", mesh.get_edge_inner_product(sig), 8.0 * (2 * mesh.nE + mesh.nC)),
            ("Me(aniso)", mesh.get_edge_inner_product(sig3), 8.0 * (2 * mesh.nE + 3 * mesh.nC)),
            ("Mf(iso)^-1", mesh.get_face_inner_product(sig, invert_model=True, invert_matrix=True), 8.0 * (2 * mesh.nF + mesh.nC))]
    vF = torch.randn(mesh.nF, dtype=torch.float64, device=dev)
    dMf = mesh.get_face_inner_product_deriv(sig)(vF)
    dMe3 = mesh.get_edge_inner_product_deriv(sig3)(torch.randn(mesh.nE, dtype=torch.float64, device=dev))
    mass += [("dMf(iso)", dMf, 8.0 * (2 * mesh.nF + mesh.nC)), ("dMf(iso).T", dMf.T, 8.0 * (2 * mesh.nF + mesh.nC)),
             ("dMe(aniso)", dMe3, 8.0 * (2 * mesh.nE + 3 * mesh.nC)), ("dMe(aniso).T", dMe3.T, 8.0 * (2 * mesh.nE + 3 * mesh.nC))]
    big = torch.randn(mesh.nE, dtype=torch.float64, device=dev)
    out = torch.empty(mesh.nE, dtype=torch.float64, device=dev)
    ref = {}
    march = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [-1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]
    MARCH = {-1: "per-operator tuned", 1: "by4 kc32", 2: "by8 kc32", 3: "by8 kc16", 4: "by8 kc16 minb8", 5: "by8 kc32 minb8", 6: "by4 kc32 minb16",
             7: "by4 kc16 minb16", 8: "by8 kc8 minb8", 9: "by16 kc16 minb4", 10: "by8 kc16 minb6", 11: "by8 kc64 minb8",
             12: "by8 kc16 minb4", 13: "by8 kc16 minb2", 14: "by4 kc16 minb1"}
    lib.dzb_tune_march(0)
    NM = [n for n in names if n[0] in ("D", "D.T", "G", "G.T", "C", "C.T", "aveE2CC", "aveN2CC", "aveF2CC")]
    for name, op in NM:
        x, y = big[:op.shape[1]], out[:op.shape[0]]
        op.apply_device(x, out=y)
        ref["m" + name] = y.clone() if name in ("D", "G.T") else float(y.abs().sum())
    for mc in march:
        lib.dzb_tune_march(mc)
        row = {"march": mc, "shape": MARCH.get(mc, "?")}
        for name, op in NM:
            x, y = big[:op.shape[1]], out[:op.shape[0]]
            ms = timeit(lambda: op.apply_device(x, out=y))
            row[name] = f"{ms:.3f}ms {8.0 * (op.shape[0] + op.shape[1]) / ms / 1e6 / PEAK:.2f}"
            r = ref["m" + name]
            err = float((y - r).abs().max() / r.abs().max()) if torch.is_tensor(r) else abs(float(y.abs().sum()) - r) / r
            if err > 1e-12:
                row[name] += f" ERR {err:.1e}"
        print(json.dumps(row), flush=True)
    for mm in (0, 1):
        lib.dzb_tune_mass(mm)
        row = {"mass_march": mm}
        for name, op, nbytes in mass:
            x, y = big[:op.shape[1]], out[:op.shape[0]]
            ms = timeit(lambda: op.apply_device(x, out=y))
            row[name] = f"{ms:.3f}ms {nbytes / ms / 1e6 / PEAK:.2f}"
        print(json.dumps(row), flush=True)
    lib.dzb_tune_march(0)
    for cfg in cfgs:
        lib.dzb_tune_elementary(cfg)
        row = {"cfg": cfg, "shape": CFGS.get(cfg, "?")}
        for name, op in names:
            x, y = big[:op.shape[1]], out[:op.shape[0]]
            ms = timeit(lambda: op.apply_device(x, out=y))
            nbytes = 8.0 * (op.shape[0] + op.shape[1])
            row[name] = f"{ms:.3f}ms {nbytes / ms / 1e6 / PEAK:.2f}"
            chk = float(y.double().abs().sum())
            if name in ref and ref[name] != chk:
                row[name] += " MISMATCH"
            ref.setdefault(name, chk)
        print(json.dumps(row), flush=True)
    lib.dzb_tune_elementary(2)
    lib.dzb_tune_march(-1)


if __name__ == "__main__":
    main()

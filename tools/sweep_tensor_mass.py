"""K6 (TensorMesh full-tensor mass apply): marching kernels (configurations 1..6) against the generic row generator (0).
python tools/sweep_tensor_mass.py [n] [configs ...]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz
from discretize_b200._lib import load

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfgs = [int(a) for a in sys.argv[2:]] or [0, 1, 2, 3, 4, 5, 6]
lib = load()
lib.dzb_tune_mass_tensor.argtypes = [ctypes.c_int]
lib.dzb_tune_mass_tensor.restype = None
dev = torch.device("cuda")
PEAK = 6541.1
for shape in ([n, n, n], [n // 2 + 3, n // 2 - 5, 37]):
    mesh = dz.TensorMesh(shape) if shape[0] == n else dz.TensorMesh([torch.rand(s).double().numpy() + 0.5 for s in shape])
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    sig = torch.rand((mesh.nC, 6), dtype=torch.float64, device=dev, generator=g)
    sig[:, :3] += 2.0
    sig[:, 3:] -= 0.5
    for proj, nd in (("F", mesh.nF), ("E", mesh.nE)):
        for inv in (False, True):
            get = mesh.get_face_inner_product if proj == "F" else mesh.get_edge_inner_product
            x = torch.randn(nd, dtype=torch.float64, device=dev, generator=g)
            ref = None
            for c in cfgs:
                lib.dzb_tune_mass_tensor(c)
                M = get(sig, invert_model=inv)
                if c == 0 and inv:          # generic generator inverts inside the kernel
                    M._apply_model, M._apply_flags = M.model.dev, M.flags
                out = torch.full((nd,), float("nan"), dtype=torch.float64, device=dev)
                for _ in range(2):
                    M.apply_device(x, out=out)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    M.apply_device(x, out=out)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                if ref is None:
                    ref = out.clone()
                err = float((out - ref).abs().max() / ref.abs().max())
                alg = 8.0 * (2 * nd + 6 * mesh.nC)
                print(f"{shape} M{proj.lower()} invert_model={inv!s:5} cfg {c}: {ms:8.3f} ms  {alg / ms / 1e6:6.0f} GB/s  "
                      f"frac {alg / ms / 1e6 / PEAK:.2f}  rel diff vs first {err:.1e}", flush=True)
    lib.dzb_tune_mass_tensor(1)

"""K7 with inversions: d(M(m)^-1 v)/dm etc. through the inversion-free marching kernels (fast) against the gather kernels
with the inversions inside (generic).  python tools/bench_deriv.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz
from discretize_b200.tensor_ops import MassDerivOperator

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
mesh = dz.TensorMesh([n, n, n])
dev = torch.device("cuda")
g = torch.Generator(device=dev)
g.manual_seed(1)
peak = 6541.1


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for proj, ndof in (("E", mesh.nE), ("F", mesh.nF)):
    for ncomp in (1, 3):
        model = torch.exp(torch.randn(mesh.nC * ncomp, dtype=torch.float64, device=dev, generator=g))
        v = torch.randn(ndof, dtype=torch.float64, device=dev, generator=g)
        dm = torch.randn(mesh.nC * ncomp, dtype=torch.float64, device=dev, generator=g)
        w = torch.randn(ndof, dtype=torch.float64, device=dev, generator=g)
        for im, ix in ((False, False), (True, False), (False, True), (True, True)):
            get = mesh.get_edge_inner_product_deriv if proj == "E" else mesh.get_face_inner_product_deriv
            F = get(model.reshape(ncomp, -1).T if ncomp == 3 else model, invert_model=im, invert_matrix=ix)(v)
            y, z = torch.empty(ndof, dtype=torch.float64, device=dev), torch.empty(mesh.nC * ncomp, dtype=torch.float64, device=dev)
            res = {}
            for fast in (True, False):
                MassDerivOperator._fast_inverse = fast
                tf = timeit(lambda: F.apply_device(dm, out=y))
                ta = timeit(lambda: F.apply_device(w, transpose=True, out=z))
                res[fast] = (tf, ta, y.clone(), z.clone())
            MassDerivOperator._fast_inverse = True
            ef = float((res[True][2] - res[False][2]).abs().max() / res[False][2].abs().max())
            ea = float((res[True][3] - res[False][3]).abs().max() / res[False][3].abs().max())
            bytes_ = 8.0 * (2 * ndof + 2 * mesh.nC * ncomp)
            print(f"{proj} {'iso  ' if ncomp == 1 else 'aniso'} invert_model={im!s:5} invert_matrix={ix!s:5}  "
                  f"F dm {res[True][0]:.3f} ms ({bytes_ / res[True][0] / 1e6 / peak:.2f}) [generic {res[False][0]:.3f}]  "
                  f"F.T w {res[True][1]:.3f} ms ({bytes_ / res[True][1] / 1e6 / peak:.2f}) [generic {res[False][1]:.3f}]  "
                  f"rel diff {ef:.1e} / {ea:.1e}", flush=True)

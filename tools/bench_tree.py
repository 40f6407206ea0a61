"""Config 3 (OcTree, 10 075 955 cells): K12 applies as shipped -- time, algorithmic bytes, fraction of the HBM peak.
python tools/bench_tree.py [refinement points] [operator ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import discretize_b200 as dz

args = sys.argv[1:]
npts = int(args[0]) if args and args[0].isdigit() else 48000
names = [a for a in args if not a.isdigit()] or ["face_divergence", "edge_curl", "nodal_gradient", "average_face_to_cell",
                                                 "average_edge_to_cell", "average_node_to_cell"]
peak = 6541.1
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
t0 = time.time()
m = dz.TreeMesh([512, 512, 512], diagonal_balance=True)
m.refine_points(np.random.default_rng(2024).random((npts, 3)) * 0.6 + 0.2, -1, [2, 2, 2], finalize=True)
print(f"cells {m.nC} (built in {time.time() - t0:.1f} s)", flush=True)
dev = torch.device("cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
from discretize_b200 import tree_ops

# DZB_SORT=face_divergence.T,nodal_gradient,... : time those operators a second time with their rows stored in
# column-locality order; DZB_LAYOUT=lpr : the CSR kernel instead of sliced ELL, for comparison
extra_sorted = set(a for a in os.environ.get("DZB_SORT", "").split(",") if a)
if "all" in extra_sorted:
    extra_sorted = set(n + t for n in names for t in ("", ".T"))
force_layout = os.environ.get("DZB_LAYOUT")
shipped = set(tree_ops.SORTED_ROW_OPERATORS)


def run(label, o):
    x = torch.randn(o.shape[1], dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    out = torch.empty(o.shape[0], dtype=torch.float64, device=dev)
    csr = o._csr(o.transposed)
    if force_layout and csr.layout != "ell":
        csr.use_layout(force_layout)
    for _ in range(3):
        o.apply_device(x, out=out)
    e0.record()
    for _ in range(20):
        o.apply_device(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    b = csr.bytes_per_apply()
    ng = sum(g is not None for g in getattr(csr, "geoms", []))
    print(f"{label:26s} {ms * 1e3:7.1f} us  {b / 1e6:7.1f} MB  {b / ms / 1e6:6.0f} GB/s  frac {b / ms / 1e6 / peak:.2f}  "
          f"layout {csr.layout} geometry arrays {ng} codes {getattr(csr, 'ntab', '-')} padding {getattr(csr, 'sell_padding', 0.0):.3f} "
          f"sorted {getattr(csr, 'ell_perm', None) is not None}", flush=True)
    return out


for name in names:
    for t in ("", ".T"):
        tree_ops.SORTED_ROW_OPERATORS = set(shipped)
        m._cache.pop("OP_" + name, None)
        op = getattr(m, name)
        ref = run(name + t, op.T if t else op)
        if name + t in extra_sorted and name + t not in shipped:
            tree_ops.SORTED_ROW_OPERATORS = shipped | {name + t}
            m._cache.pop("OP_" + name, None)
            op = getattr(m, name)
            got = run(name + t + " [sorted]", op.T if t else op)
            print(f"{'':26s} rel diff sorted / natural {float((got - ref).abs().max() / ref.abs().max()):.1e}", flush=True)
        del op
    for key in ("P_" + name, "H_" + name, "OP_" + name):
        m._cache.pop(key, None)
    torch.cuda.empty_cache()

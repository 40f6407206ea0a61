"""Lanes-per-row sweep for the coded-CSR tree SpMV (K12)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import discretize_b200 as dz

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
m = dz.TreeMesh([512, 512, 512], diagonal_balance=True)
m.refine_points(np.random.default_rng(2024).random((npts, 3)) * 0.6 + 0.2, -1, [2, 2, 2], finalize=True)
print("cells", m.nC, flush=True)
dev = torch.device("cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name in ("face_divergence", "edge_curl", "average_edge_to_cell", "average_node_to_cell", "nodal_gradient"):
    op = getattr(m, name)
    x = torch.randn(op.shape[1], dtype=torch.float64, device=dev)
    out = torch.empty(op.shape[0], dtype=torch.float64, device=dev)
    csr = op._csr(False)
    res = []
    variants = [("lpr", 1), ("lpr", 2), ("lpr", 4), ("stream", 0)] + ([("ell", 0)] if csr.row_len > 0 else [])
    for layout, lpr in variants:
        csr.use_layout(layout)
        csr.lpr = lpr
        for _ in range(3):
            op.apply_device(x, out=out)
        e0.record()
        for _ in range(20):
            op.apply_device(x, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res.append((layout + (str(lpr) if layout == "lpr" else ""), ms, csr.bytes_per_apply() / ms / 1e6))
    print(name, "rows", op.shape[0], "nnz/row %.2f" % (csr.nnz / op.shape[0]),
          " ".join(f"{l}:{t*1e3:.0f}us({b:.0f}GB/s)" for l, t, b in res), flush=True)

"""K12 (coded OcTree SpMV): natural against column-locality row order of the ELL layout, on the config-3 tree.

  python tools/sweep_tree.py [refinement points, default 48000 = the 10 075 955-cell BASELINE tree]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import discretize_b200 as dz
from discretize_b200 import tree_ops

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 48000
m = dz.TreeMesh([512, 512, 512], diagonal_balance=True)
m.refine_points(np.random.default_rng(2024).random((npts, 3)) * 0.6 + 0.2, -1, [2, 2, 2], finalize=True)
print("cells", m.nC, flush=True)
dev = torch.device("cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
names = ("face_divergence", "average_face_to_cell", "average_edge_to_cell", "average_node_to_cell", "edge_curl")
for name in names:
    res = []
    ref = None
    for sort in (False, True):
        tree_ops.SORTED_ROW_OPERATORS = set(names) if sort else set()
        for key in ("P_" + name, "H_" + name, "OP_" + name):
            m._cache.pop(key, None)
        op = getattr(m, name)
        x = torch.randn(op.shape[1], dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
        out = torch.empty(op.shape[0], dtype=torch.float64, device=dev)
        csr = op._csr(False)
        for _ in range(3):
            op.apply_device(x, out=out)
        e0.record()
        for _ in range(20):
            op.apply_device(x, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        if ref is None:
            ref = out.clone()
        err = float((out - ref).abs().max() / ref.abs().max())
        res.append(f"{'sorted ' if sort else 'natural'} rows: {ms * 1e3:.0f} us ({csr.bytes_per_apply() / ms / 1e6:.0f} GB/s, {csr.layout}, err {err:.1e})")
        del op, csr
    print(f"{name:24s}", " | ".join(res), flush=True)

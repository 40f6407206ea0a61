"""Config 4 (10^8 points on 512^3): build, apply, transpose of the compact interpolation rows.
python tools/bench_interp.py [npts] [loc ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import discretize_b200 as dz
from discretize_b200._lib import check, load
from discretize_b200.operators import ptr, stream_ptr


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


npts = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
locs = sys.argv[2:] or ["CC", "Ex"]
mesh = dz.TensorMesh([512, 512, 512])
dev = torch.device("cuda")
g = torch.Generator(device=dev)
g.manual_seed(10)
pts = torch.rand((npts, 3), dtype=torch.float64, device=dev, generator=g)
for loc in locs:
    Q = mesh.get_interpolation_matrix(pts, loc)
    print(f"{loc}: build (locate, compact rows) {timeit(Q._locate):.3f} ms", flush=True)
    v = torch.randn(Q.shape[1], dtype=torch.float64, device=dev)
    out = torch.empty(npts, dtype=torch.float64, device=dev)
    Q.bucket_min_points = 1 << 62
    print(f"{loc}: apply, rows in point order {timeit(lambda: Q.apply_device(v, out=out)):.3f} ms", flush=True)
    ref = out.clone()
    Q.bucket_min_points = 1 << 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); Q.bucket(); b.record(); torch.cuda.synchronize()
    print(f"{loc}: bucket (argsort + 4 gathers) {a.elapsed_time(b):.1f} ms", flush=True)
    ms = timeit(lambda: Q.apply_device(v, out=out))
    print(f"{loc}: apply bucketed + un-permute {ms:.3f} ms  ({npts / ms / 1e6:.1f} Gpts/s)  same={torch.equal(out, ref)}", flush=True)
    ms_k = timeit(lambda: check(load().dzb_interp_apply_compact(*Q._row_args(), None, npts, ptr(v), ptr(Q._sorted_out),
                                                                stream_ptr()), "apply"))
    print(f"{loc}:   kernel alone (bucket order) {ms_k:.3f} ms; un-permute {ms - ms_k:.3f} ms", flush=True)
    x = torch.randn(npts, dtype=torch.float64, device=dev)
    o2 = torch.empty(Q.shape[1], dtype=torch.float64, device=dev)
    print(f"{loc}: Q.T @ x bucketed {timeit(lambda: Q.apply_device(x, transpose=True, out=o2)):.3f} ms", flush=True)
    del Q, v, out, ref, x, o2
    torch.cuda.empty_cache()

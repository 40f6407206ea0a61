#!/usr/bin/env python
"""bench.py -- operator-apply throughput (Gdof/s) and HBM-roofline fraction on B200.

Workload at N=1 = BASELINE.json configs[1]: TensorMesh 512^3 DC-resistivity nodal operator
y = G^T Me(sigma) G phi, applied matrix-free by ONE fused kernel (K8).  A "step" is one
apply over one synthetic field (phi ~ N(0,1), sigma log-normal).  "dof" = one output row.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 512]

Under torchrun (N>1) every rank owns a z-slab: by default WEAK scaling -- each GPU keeps the N=1 workload (a 512^3
slab of a 512 x 512 x 512N mesh) and exchanges one node plane with each neighbour per apply over NCCL; `--scaling
strong` splits the 512^3 mesh itself; `--workload maxwell` is BASELINE config 5 (1024^3, strong scaling by
definition).  Timing = barrier + cuda sync, CUDA events, max over ranks.
`--impl reference` times the reference's own CPU path (scipy csr_matvec on the assembled
G^T Me G) on a bounded sample of the same workload on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "operator_apply_gdof_per_s"
UNIT = "Gdof/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's scipy path on a bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_reference_arm(n_sample, steps, warmup):
    """scipy csr_matvec on the assembled A = G^T Me(sigma) G (what the reference/SimPEG applies)."""
    from oracle import tensor_ops as T  # the checker; allowed here (cpu_baseline leg only)

    g = T.Grid([n_sample] * 3)
    sigma = np.exp(np.random.default_rng(6).standard_normal(g.nC))
    phi = np.random.default_rng(5).standard_normal(g.nN)
    G = T.nodal_gradient(g)
    Me = T.inner_product(g, "E", sigma)
    A = (G.T @ Me @ G).tocsr()
    for _ in range(warmup):
        A @ phi
    t0 = time.perf_counter()
    for _ in range(steps):
        A @ phi
    dt = (time.perf_counter() - t0) / steps
    return g.nN / dt / 1e9, dt, g.nN


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.cpu_n
    val, dt, ndof = cpu_reference_arm(n_sample, max(args.steps, 1), max(args.warmup, 1))
    sample = (f"TensorMesh {n_sample}^3 (of the 512^3 workload), assembled A=G^T Me(sigma) G as scipy CSR, "
              f"A @ phi via scipy csr_matvec, 1 thread (scipy SpMV is single-threaded) of {os.cpu_count()} cores")
    line = {
        "metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TensorMesh 512^3 DC nodal G^T Me(sigma) G apply (CPU sample at {n_sample}^3)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import discretize_b200 as dz

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback on the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    if args.workload == "maxwell":
        # BASELINE config 5: C^T Mf(mui) C + Me(sigma) on n^3 (1024^3 by default for this workload), z-slabs
        from discretize_b200.slab import SlabMaxwell

        op = SlabMaxwell(n, n, n, rank, world)
        step = op.step_graphed if (world > 1 and not args.eager) else op.step
        ndof_total = n * (n + 1) * (n + 1) * 3
        alg_bytes = 8.0 * (2 * ndof_total + 2 * n ** 3) / world
        launches_per_step = op.launches_per_step
        parallelism = "1 GPU" if world == 1 else f"z-slabs x{world}, NCCL halo send/recv"
        workload_name = f"TensorMesh {n}^3 Maxwell C^T Mf(mui) C + Me(sigma) fused matrix-free apply"
        kernel_name = "k_maxwell_ring<2,NWC,PARX,4>"
    elif world == 1:
        mesh = dz.TensorMesh([n, n, n])
        gen = torch.Generator(device=dev); gen.manual_seed(5)
        phi = torch.randn(mesh.nN, dtype=torch.float64, device=dev, generator=gen)
        gen.manual_seed(6)
        sigma = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev, generator=gen))
        G = mesh.nodal_gradient
        A = G.T @ mesh.get_edge_inner_product(sigma) @ G  # fuses into the single K8 kernel
        assert type(A).__name__ == "DCNodalOperator"
        out = torch.empty(mesh.nN, dtype=torch.float64, device=dev)
        step = lambda: A.apply_device(phi, out=out)
        ndof_total = mesh.nN
        alg_bytes = 8.0 * (2 * mesh.nN + mesh.nC)
        launches_per_step = 1
        parallelism = "1 GPU"
        workload_name = f"TensorMesh {n}^3 DC-resistivity nodal operator G^T Me(sigma) G matrix-free apply"
        kernel_name = "k_dc_nodal_ring<4,NWC,PARX>"
    else:
        from discretize_b200.slab import SlabDCNodal

        # weak scaling (default): every rank owns the N=1 workload, a 512^3 slab of a 512 x 512 x (512 N) mesh, and
        # exchanges one node plane with each neighbour per apply; --scaling strong splits the 512^3 mesh itself
        # (66 us of kernel time per rank at N=8: latency-bound, profiles/r1_scale_dc_512_n1248.jsonl)
        nz_global = n * world if args.scaling == "weak" else n
        op = SlabDCNodal(n, n, nz_global, rank, world, seed_phi=5, seed_sigma=6)
        step = op.step_graphed if not args.eager else op.step
        ndof_total = (n + 1) ** 2 * (nz_global + 1)
        alg_bytes = 8.0 * (2 * ndof_total + n * n * nz_global) / world
        launches_per_step = op.launches_per_step
        parallelism = f"z-slabs x{world} ({args.scaling} scaling), NCCL halo send/recv"
        workload_name = (f"TensorMesh {n}x{n}x{nz_global} DC-resistivity nodal operator G^T Me(sigma) G matrix-free apply"
                         f" ({n}^3 per GPU)" if args.scaling == "weak" else
                         f"TensorMesh {n}^3 DC-resistivity nodal operator G^T Me(sigma) G matrix-free apply")
        kernel_name = "k_dc_nodal_ring<4,NWC,PARX>"

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = ndof_total / (ms_step * 1e-3) / 1e9

    # ---- end-to-end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if world == 1 and args.workload == "dc":
        host_in = torch.empty(mesh.nN, dtype=torch.float64).pin_memory()
        host_in.copy_(phi)
        host_out = torch.empty(mesh.nN, dtype=torch.float64).pin_memory()
        e2e_steps = max(2, min(args.steps, 10))
        for _ in range(2):  # warm-up (public API: pinned CPU tensor in -> pinned CPU tensor out)
            y = A.matvec(host_in, out=host_out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            y = A.matvec(host_in, out=host_out)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        assert y.device.type == "cpu" and y.shape[0] == mesh.nN
        e2e = {"value": ndof_total / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(8 * mesh.nN),
               "d2h_bytes_per_step": int(8 * mesh.nN), "ms_per_step": dt * 1e3,
               "api": "A = G.T @ mesh.get_edge_inner_product(sigma) @ G; A.matvec(x_host_pinned, out=y_host_pinned)"}

    e2e_error = None
    if world > 1 and args.workload == "dc":
        try:
            # every rank streams ITS slab from / to pinned host memory: H2D of the owned planes, the distributed apply
            # (halo exchange + kernels, eager launches), D2H of the owned result -- wall clock, max over ranks
            L = op.layout
            own = slice(L.p_own0 * L.plane, L.p_own1 * L.plane)
            host_in = torch.empty(op.phi[own].numel(), dtype=torch.float64).pin_memory()
            host_in.copy_(op.phi[own])
            host_out = torch.empty_like(host_in).pin_memory()
            e2e_steps = max(2, min(args.steps, 10))

            def e2e_step():
                op.step_from_host(host_in, host_out)

            for _ in range(2):
                e2e_step()
            sync()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            sync()
            tt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
            nb = torch.tensor([8.0 * host_in.numel()], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(nb)
            dt = float(tt.item())
            e2e = {"value": ndof_total / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(nb.item()),
                   "d2h_bytes_per_step": int(nb.item()), "ms_per_step": dt * 1e3,
                   "api": "per rank: SlabDCNodal.step_from_host(x_host_pinned, y_host_pinned): outer planes up, NCCL halo "
                          "exchange, then H2D / K8 / D2H pipelined over z-chunks"}
        except Exception as exc:  # the device-timed line must survive a failure of the host-path leg
            e2e, e2e_error = None, repr(exc)[:300]

    def shutdown():
        # A process group that was captured into a CUDA graph can hang in destroy_process_group();
        # every rank has its result by now, so leave without tearing NCCL down.
        if world > 1:
            sys.stdout.flush()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        shutdown()
        return

    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9  # per-GPU GB/s of the dominant kernel's algorithmic bytes
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name,
                   "dofs_per_step": int(ndof_total), "parallelism": parallelism,
                   "l2": "inputs (>= 2 GB per GPU at N=1) exceed the 126 MB L2; no flush needed"},
        "gpu_launches": launches_per_step * args.steps,
        "launch_mode": ("cuda-graph replay" if (world > 1 and not args.eager and not getattr(op, "_graph_failed", False))
                        else "eager"),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": 3.22e9 if (args.workload == "dc" and n == 512 and world == 1) else None,
                     "traffic_source": "ncu --set full, profiles/r1_k8_dc_nodal_ring_ncu_full_summary.txt (dram read+write per launch)",
                     "peak_source": peak_src, "kernel": kernel_name,
                     "algorithmic_bytes_per_launch": alg_bytes},
    }
    if e2e is not None:
        line["e2e"] = e2e
    elif e2e_error is not None:
        line["e2e_error"] = e2e_error
    if world == 1 and not args.no_cpu and args.workload == "dc":
        val, dt, _ = cpu_reference_arm(args.cpu_n, 3, 1)
        line["cpu_baseline"] = {
            "value": val, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{args.cpu_n}^3 sample of the workload: scipy csr_matvec on the assembled G^T Me G "
                      f"(oracle port of the reference build), 1 of {os.cpu_count()} host cores",
        }
    print(json.dumps(line), flush=True)
    shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=None, help="cells per axis (default 512 for dc = BASELINE configs[1], 1024 for maxwell)")
    ap.add_argument("--workload", default="dc", choices=["dc", "maxwell"],
                    help="dc = configs[1] (the N=1 headline); maxwell = configs[4], the 1024^3 z-slab scaling case")
    ap.add_argument("--cpu-n", type=int, default=128, help="cells per axis of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N > 1: weak = every GPU owns the N=1 workload (default for dc), strong = the N=1 mesh is split "
                         "(default for maxwell: BASELINE config 5 is the 1024^3 mesh on 2/4/8 GPUs)")
    ap.add_argument("--eager", action="store_true", help="multi-GPU: launch every step from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.n is None:
        args.n = 512 if args.workload == "dc" else 1024
    if args.scaling is None:
        args.scaling = "weak" if args.workload == "dc" else "strong"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

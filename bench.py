#!/usr/bin/env python
"""bench.py -- operator-apply throughput (Gdof/s) and HBM-roofline fraction on B200.

Workload (every N) = BASELINE.json configs[4], the configuration the north-star targets are quoted on:
TensorMesh 1024^3 Maxwell operator  y = C^T Mf(mui) C e + Me(sigma) e  (tests/boundaries/test_boundary_maxwell.py:95-99
in the reference), applied matrix-free by ONE fused kernel (K9) per rank.  It fits one B200 (68.8 GB), so N=1 runs the
same mesh and N = 2/4/8 split it into z-slabs (STRONG scaling: efficiency = v_N / (N v_1) is the north-star figure).
A "step" is one apply over one synthetic field (e ~ N(0,1), sigma and mui log-normal); "dof" = one output row.
At N=1 the line also carries `secondary`: BASELINE configs[1] (512^3 DC nodal operator G^T Me(sigma) G, kernel K8), and
`secondary_complex`: the frequency-domain form C^T Mf C + i omega Me on a complex 512^3 field (kernel K9c, one launch).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload maxwell|dc] [--n N]

Timing: W warm-up steps, then R (--repeats, default 5) blocks of EXACTLY K steps, each block bracketed by a barrier +
cuda synchronize, CUDA events on the launching stream, max over ranks per block; `ms_per_step` is the MEDIAN block
(best and all blocks are reported next to it).  Inputs (>= 8 GB per GPU) exceed the 126 MB L2, so no flush is needed.
`--impl reference` times the reference's own CPU path (scipy csr_matvec, one thread) on a bounded sample of the same
workload on the host cores: both the chain C^T (Mf (C e)) + Me e and the pre-assembled A.
"""
from __future__ import annotations

import argparse
import glob
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


def ncu_traffic(kernel_key, n):
    """dram read + write bytes per launch of the dominant kernel at mesh size n, read from the committed ncu capture
    (profiles/r2_*ncu*.json, written by tools/ncu_to_json.py from the .ncu-rep), or (None, reason)."""
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_*ncu*.json"))):
        try:
            with open(path) as f:
                d = json.load(f)
        except Exception:
            continue
        for k in d.get("kernels", []):
            if kernel_key in k.get("name", "") and int(k.get("n", -1)) == int(n):
                best = (float(k["dram_bytes_read"]) + float(k["dram_bytes_write"]), os.path.relpath(path, ROOT))
    return best if best else (None, "no committed ncu capture for this kernel / size")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
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
def cpu_maxwell_arm(n_sample, steps, warmup):
    """scipy csr_matvec on the reference's matrices for C^T Mf(mui) C + Me(sigma): chain and pre-assembled A."""
    from oracle import tensor_ops as T  # the checker; allowed here (cpu_baseline / --impl reference legs only)

    g = T.Grid([n_sample] * 3)
    sigma = np.exp(np.random.default_rng(12).standard_normal(g.nC))
    mui = np.exp(np.random.default_rng(13).standard_normal(g.nC))
    e = np.random.default_rng(11).standard_normal(g.nE)
    t0 = time.perf_counter()
    Cm = T.edge_curl(g)
    Mf, Me = T.inner_product(g, "F", mui), T.inner_product(g, "E", sigma)
    CT = Cm.T.tocsr()
    A = (CT @ Mf @ Cm + Me).tocsr()
    t_build = time.perf_counter() - t0

    def timed(fn):
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        return (time.perf_counter() - t0) / steps

    dt_chain = timed(lambda: CT @ (Mf @ (Cm @ e)) + Me @ e)
    dt_asm = timed(lambda: A @ e)
    return {"ndof": g.nE, "dt_chain": dt_chain, "dt_assembled": dt_asm, "build_s": t_build,
            "gdof_chain": g.nE / dt_chain / 1e9, "gdof_assembled": g.nE / dt_asm / 1e9}


def cpu_dc_arm(n_sample, steps, warmup):
    """scipy csr_matvec for G^T Me(sigma) G: the 3-SpMV chain reference users run and the pre-assembled A."""
    from oracle import tensor_ops as T

    g = T.Grid([n_sample] * 3)
    sigma = np.exp(np.random.default_rng(6).standard_normal(g.nC))
    phi = np.random.default_rng(5).standard_normal(g.nN)
    t0 = time.perf_counter()
    G = T.nodal_gradient(g)
    Me = T.inner_product(g, "E", sigma)
    GT = G.T.tocsr()
    A = (GT @ Me @ G).tocsr()
    t_build = time.perf_counter() - t0

    def timed(fn):
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        return (time.perf_counter() - t0) / steps

    dt_chain = timed(lambda: GT @ (Me @ (G @ phi)))
    dt_asm = timed(lambda: A @ phi)
    return {"ndof": g.nN, "dt_chain": dt_chain, "dt_assembled": dt_asm, "build_s": t_build,
            "gdof_chain": g.nN / dt_chain / 1e9, "gdof_assembled": g.nN / dt_asm / 1e9}


def cpu_baseline_object(res, what, n_sample, n_full):
    import scipy

    return {
        "value": res["gdof_assembled"], "unit": UNIT, "cores": 1, "kind": "port",
        "chain_value": res["gdof_chain"], "chain_ms": res["dt_chain"] * 1e3, "assembled_ms": res["dt_assembled"] * 1e3,
        "assembly_s": res["build_s"],
        "sample": (f"{n_sample}^3 sample of the {n_full}^3 workload (per-dof normalised; the reference cannot assemble "
                   f"{n_full}^3): {what}; value = pre-assembled A @ x, chain_value = the SpMV chain reference users run; "
                   f"scipy {scipy.__version__} csr_matvec is single-threaded: 1 of {os.cpu_count()} host cores; matrices "
                   f"assembled by the oracle port of the reference (oracle/tensor_ops.py, pinned against the reference)"),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.cpu_n
    steps, warmup = max(args.steps, 1), max(args.warmup, 1)
    if args.workload == "maxwell":
        res = cpu_maxwell_arm(n_sample, steps, warmup)
        what = "C^T Mf(mui) C + Me(sigma) as scipy CSR"
        workload = f"TensorMesh {args.n}^3 Maxwell C^T Mf(mui) C + Me(sigma) apply (CPU sample at {n_sample}^3)"
    else:
        res = cpu_dc_arm(n_sample, steps, warmup)
        what = "G^T Me(sigma) G as scipy CSR"
        workload = f"TensorMesh {args.n}^3 DC nodal G^T Me(sigma) G apply (CPU sample at {n_sample}^3)"
    cb = cpu_baseline_object(res, what, n_sample, args.n)
    line = {
        "metric": METRIC, "value": cb["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["dt_assembled"] * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "dofs_per_step": int(res["ndof"])},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def timed_blocks(step, steps, repeats, sync, world, dev):
    """R blocks of exactly `steps` steps; returns per-block ms (max over ranks)."""
    import torch
    import torch.distributed as dist

    out = []
    for _ in range(repeats):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        ev0.record()
        for _ in range(steps):
            step()
        ev1.record()
        sync()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        out.append(ms)
    return out


def secondary_complex_maxwell(args, dev, peak):
    """The frequency-domain form of the benched operator, C^T Mf C + i omega Me on a COMPLEX 512^3 field: one launch (K9c)."""
    import torch

    import discretize_b200 as dz

    n = 512
    mesh = dz.TensorMesh([n, n, n])
    gen = torch.Generator(device=dev); gen.manual_seed(21)
    e = torch.complex(torch.randn(mesh.nE, dtype=torch.float64, device=dev, generator=gen),
                      torch.randn(mesh.nE, dtype=torch.float64, device=dev, generator=gen))
    sigma = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev, generator=gen))
    mui = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev, generator=gen))
    C = mesh.edge_curl
    A = C.T @ mesh.get_face_inner_product(mui) @ C + 1j * (2 * np.pi * 10.0) * mesh.get_edge_inner_product(sigma)
    assert type(A).__name__ == "ComplexMaxwellOperator"
    out = torch.empty_like(e)
    if A.apply_device_complex(e, out=out) is None:
        raise RuntimeError("the one-launch complex kernel does not apply to this mesh")
    step = lambda: A.apply_device_complex(e, out=out)
    for _ in range(3):
        step()
    blocks = timed_blocks(step, args.steps, 3, torch.cuda.synchronize, 1, dev)
    ms = float(np.median(blocks)) / args.steps
    alg = 8.0 * (4 * mesh.nE + 2 * mesh.nC)          # complex e in and out, the two models once
    return {
        "workload": "TensorMesh 512^3 frequency-domain Maxwell C^T Mf(mui) C + i omega Me(sigma) on a complex field, one launch",
        "value": mesh.nE / (ms * 1e-3) / 1e9, "unit": "complex " + UNIT, "ms_per_step": ms, "steps": args.steps,
        "block_ms": blocks, "dofs_per_step": int(mesh.nE), "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "kernel": "k_maxwell_cplx<1,6,PARX,4,2>", "algorithmic_bytes_per_launch": alg},
    }


def secondary_dc(args, dev, peak):
    """BASELINE configs[1] on one GPU: 512^3 DC nodal operator, device-resident + end to end."""
    import torch

    import discretize_b200 as dz

    n = 512
    mesh = dz.TensorMesh([n, n, n])
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    phi = torch.randn(mesh.nN, dtype=torch.float64, device=dev, generator=gen)
    gen.manual_seed(6)
    sigma = torch.exp(torch.randn(mesh.nC, dtype=torch.float64, device=dev, generator=gen))
    G = mesh.nodal_gradient
    A = G.T @ mesh.get_edge_inner_product(sigma) @ G
    assert type(A).__name__ == "DCNodalOperator"
    out = torch.empty(mesh.nN, dtype=torch.float64, device=dev)
    step = lambda: A.apply_device(phi, out=out)
    for _ in range(3):
        step()
    blocks = timed_blocks(step, args.steps, 3, torch.cuda.synchronize, 1, dev)
    ms = float(np.median(blocks)) / args.steps
    alg = 8.0 * (2 * mesh.nN + mesh.nC)
    host_in = torch.empty(mesh.nN, dtype=torch.float64).pin_memory()
    host_in.copy_(phi)
    host_out = torch.empty(mesh.nN, dtype=torch.float64).pin_memory()
    for _ in range(2):
        A.matvec(host_in, out=host_out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        A.matvec(host_in, out=host_out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    traffic, src = ncu_traffic("k_dc_nodal_ring", 512)
    return {
        "workload": "TensorMesh 512^3 DC-resistivity nodal operator G^T Me(sigma) G matrix-free apply (BASELINE configs[1])",
        "value": mesh.nN / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": args.steps,
        "block_ms": blocks, "dofs_per_step": int(mesh.nN),
        "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_source": src,
                     "kernel": "k_dc_nodal_ring<4,NWC,PARX>", "algorithmic_bytes_per_launch": alg},
        "e2e": {"value": mesh.nN / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": int(8 * mesh.nN),
                "d2h_bytes_per_step": int(8 * mesh.nN),
                "api": "A = G.T @ mesh.get_edge_inner_product(sigma) @ G; A.matvec(x_host_pinned, out=y_host_pinned)"},
    }


def run_ours(args):
    import torch
    import torch.distributed as dist

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
        # BASELINE configs[4]: C^T Mf(mui) C + Me(sigma) on n^3, z-slabs across the ranks (strong scaling)
        from discretize_b200.slab import SlabMaxwell

        op, halo_mode = None, "none"
        if world > 1 and args.halo.startswith("peer"):
            try:        # halo planes fetched by the kernel from the neighbours' memory (symmetric allocation): one launch per apply
                op = SlabMaxwell(n, n, n, rank, world, peer_halo=True)
                op.step_peer()
                torch.cuda.synchronize()
                halo_mode = "peer"
            except Exception as exc:
                sys.stderr.write(f"[bench] peer-read halos unavailable ({exc!r}); using NCCL send/recv\n")
                op = None
        if op is None:
            op = SlabMaxwell(n, n, n, rank, world)
            halo_mode = "nccl" if world > 1 else "none"
        ndof_total = n * (n + 1) * (n + 1) * 3
        alg_bytes = 8.0 * (2 * ndof_total + 2 * n ** 3) / world
        workload_name = f"TensorMesh {n}^3 Maxwell C^T Mf(mui) C + Me(sigma) fused matrix-free apply (BASELINE configs[4])"
        kernel_name, kernel_key = "k_maxwell_tmap<XS,R,PARX,NSLOT> (tensor-map TMA ring)", "k_maxwell_tmap"
    else:
        from discretize_b200.slab import SlabDCNodal

        op, halo_mode = None, "none"
        if world > 1 and args.halo.startswith("peer"):
            try:
                op = SlabDCNodal(n, n, n, rank, world, seed_phi=5, seed_sigma=6, peer_halo=True)
                op.step_peer()
                torch.cuda.synchronize()
                halo_mode = "peer"
            except Exception as exc:
                sys.stderr.write(f"[bench] peer-read halos unavailable ({exc!r}); using NCCL send/recv\n")
                op = None
        if op is None:
            op = SlabDCNodal(n, n, n, rank, world, seed_phi=5, seed_sigma=6)
            halo_mode = "nccl" if world > 1 else "none"
        ndof_total = (n + 1) ** 3
        alg_bytes = 8.0 * (2 * ndof_total + n ** 3) / world
        workload_name = f"TensorMesh {n}^3 DC-resistivity nodal operator G^T Me(sigma) G matrix-free apply (BASELINE configs[1])"
        kernel_name, kernel_key = "k_dc_nodal_ring<4,NWC,PARX>", "k_dc_nodal_ring"
    if halo_mode == "peer":
        step, launches_per_step = op.step_peer, 1
        if args.halo == "peer-barrier":      # ordering by a device-side barrier in front of every launch (comparison)
            step = lambda: op.step_peer(sync="barrier")
        parallelism = (f"z-slabs x{world} (strong scaling); halo planes read from the neighbours' memory over NVLink by the "
                       "fused kernel itself (TMA copies from symmetric memory), ordered by epoch signals the launches write into each "
                       "other's memory: one launch per apply, no send/recv, no barrier, no collective")
    else:
        step = op.step_graphed if (world > 1 and not args.eager) else op.step
        launches_per_step = op.launches_per_step
        parallelism = "1 GPU" if world == 1 else f"z-slabs x{world} (strong scaling), NCCL halo send/recv, no collective"

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    blocks = timed_blocks(step, args.steps, args.repeats, sync, world, dev)
    if halo_mode == "peer" and args.halo == "peer":
        # a halo wait inside a kernel that gave up (2 s) invalidates the run on EVERY rank: measure again with the
        # barrier-ordered launches, and say so
        bad = torch.tensor([1.0 if op.peer_sync_timed_out() else 0.0], device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if float(bad.item()) > 0:
            sys.stderr.write("[bench] a halo wait inside the kernel timed out; re-measuring with a barrier per apply\n")
            step = lambda: op.step_peer(sync="barrier")
            parallelism = parallelism.replace("ordered by epoch signals the launches write into each other's memory",
                                              "ordered by a device-side barrier in front of each launch")
            for _ in range(warmup):
                step()
            sync()
            blocks = timed_blocks(step, args.steps, args.repeats, sync, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = float(np.median(blocks)) / args.steps
    value = ndof_total / (ms_step * 1e-3) / 1e9

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e, e2e_error = None, None
    if not args.no_e2e:
        try:
            e2e_steps = max(2, min(args.steps, args.e2e_steps))
            if world == 1:
                A = op.op                                      # the operator a user builds: C.T @ Mf @ C + Me (fused)
                nvec = A.shape[0]
                host_in = torch.empty(nvec, dtype=torch.float64).pin_memory()
                host_in.copy_(op.e if args.workload == "maxwell" else op.phi)
                host_out = torch.empty(nvec, dtype=torch.float64).pin_memory()
                A.matvec(host_in, out=host_out)                # warm-up (allocates the device staging buffers)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    y = A.matvec(host_in, out=host_out)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / e2e_steps
                assert y.device.type == "cpu" and y.shape[0] == nvec
                # the streamed result must be the device-resident one
                dev_out = op.out if world == 1 else None
                chk = float((host_out[:4096].to(dev) - dev_out[:4096]).abs().max())
                api = ("A = C.T @ mesh.get_face_inner_product(mui) @ C + mesh.get_edge_inner_product(sigma); "
                       "A.matvec(x_host_pinned, out=y_host_pinned)") if args.workload == "maxwell" else \
                      "A = G.T @ mesh.get_edge_inner_product(sigma) @ G; A.matvec(x_host_pinned, out=y_host_pinned)"
                e2e = {"value": ndof_total / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(8 * nvec),
                       "d2h_bytes_per_step": int(8 * nvec), "ms_per_step": dt * 1e3, "steps": e2e_steps, "api": api,
                       "max_abs_diff_vs_device_resident_first_4096": chk}
                del host_in, host_out
            else:
                if args.workload == "maxwell":
                    nown = sum(op.owned_sizes())
                    host_in = torch.empty(nown, dtype=torch.float64).pin_memory()
                    op.owned_to_host(op.e, host_in)
                else:
                    L = op.layout
                    own = slice(L.p_own0 * L.plane, L.p_own1 * L.plane)
                    nown = op.phi[own].numel()
                    host_in = torch.empty(nown, dtype=torch.float64).pin_memory()
                    host_in.copy_(op.phi[own])
                host_out = torch.empty_like(host_in).pin_memory()
                op.step_from_host(host_in, host_out)
                sync()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    op.step_from_host(host_in, host_out)
                sync()
                tt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
                nb = torch.tensor([8.0 * nown], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dist.all_reduce(nb)
                dt = float(tt.item())
                e2e = {"value": ndof_total / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(nb.item()),
                       "d2h_bytes_per_step": int(nb.item()), "ms_per_step": dt * 1e3, "steps": e2e_steps,
                       "api": "per rank: Slab operator .step_from_host(x_host_pinned, y_host_pinned): boundary planes up, NCCL "
                              "halo exchange, then H2D / fused kernel / D2H pipelined over z-chunks"}
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
    traffic, traffic_src = ncu_traffic(kernel_key, n) if world == 1 else (None, "single-GPU captures only")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name,
                   "dofs_per_step": int(ndof_total), "parallelism": parallelism,
                   "l2": "inputs (>= 8 GB per GPU) exceed the 126 MB L2; no flush needed"},
        "timing": {"blocks": len(blocks), "steps_per_block": args.steps, "block_ms": blocks,
                   "ms_per_step_median": ms_step, "ms_per_step_best": min(blocks) / args.steps,
                   "note": "each block = exactly `steps` applies between barrier + synchronize; median block reported"},
        "gpu_launches": launches_per_step * args.steps,
        "launch_mode": ("eager" if (world == 1 or args.eager or halo_mode == "peer"
                                   or getattr(op, "_graph_failed", False)) else "cuda-graph replay"),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "kernel": kernel_name,
                     "algorithmic_bytes_per_launch": alg_bytes},
    }
    if e2e is not None:
        line["e2e"] = e2e
    elif e2e_error is not None:
        line["e2e_error"] = e2e_error
    if world == 1:
        del op
        torch.cuda.empty_cache()
        if args.workload == "maxwell" and not args.no_secondary:
            try:
                line["secondary"] = secondary_dc(args, dev, peak)
                torch.cuda.empty_cache()
                line["secondary_complex"] = secondary_complex_maxwell(args, dev, peak)
            except Exception as exc:
                line["secondary_error"] = repr(exc)[:300]
        if not args.no_cpu:
            if args.workload == "maxwell":
                res = cpu_maxwell_arm(args.cpu_n_inline, 3, 1)
                line["cpu_baseline"] = cpu_baseline_object(res, "C^T Mf(mui) C + Me(sigma) as scipy CSR", args.cpu_n_inline, n)
            else:
                res = cpu_dc_arm(args.cpu_n_inline, 3, 1)
                line["cpu_baseline"] = cpu_baseline_object(res, "G^T Me(sigma) G as scipy CSR", args.cpu_n_inline, n)
    print(json.dumps(line), flush=True)
    shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--repeats", type=int, default=5, help="timed blocks of `steps` steps (median reported)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="maxwell", choices=["maxwell", "dc"],
                    help="maxwell = BASELINE configs[4] (default, every N); dc = configs[1]")
    ap.add_argument("--n", type=int, default=None, help="cells per axis (default 1024 for maxwell, 512 for dc)")
    ap.add_argument("--cpu-n", type=int, default=256, help="cells per axis of the CPU sample of --impl reference")
    ap.add_argument("--cpu-n-inline", type=int, default=128, help="cells per axis of the cpu_baseline leg of the default run")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the 512^3 DC object at N=1")
    ap.add_argument("--halo", default="peer", choices=["peer", "peer-barrier", "nccl"],
                    help="N > 1, maxwell: peer = the kernel reads the halo planes from the neighbours' memory (default); "
                         "nccl = send/recv exchange overlapped with the interior planes + one launch for the boundary planes")
    ap.add_argument("--eager", action="store_true", help="multi-GPU: launch every step from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.n is None:
        args.n = 1024 if args.workload == "maxwell" else 512
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

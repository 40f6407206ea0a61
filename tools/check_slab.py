"""torchrun --nproc-per-node N tools/check_slab.py : slab-decomposed K8 == single-GPU K8 (bitwise)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from discretize_b200.slab import SlabDCNodal, SlabMaxwell

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for (nx, ny, nz) in ((40, 33, 29), (70, 20, 64), (512, 64, 96)):
    rng = np.random.default_rng(1)
    h = (rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(nz) + 0.5)
    part = SlabDCNodal(nx, ny, nz, rank, world, h=h)
    for _ in range(2):
        part.step()
    torch.cuda.synchronize()
    whole = SlabDCNodal(nx, ny, nz, 0, 1, h=h)  # every rank recomputes the undecomposed answer
    whole.step()
    torch.cuda.synchronize()
    ref = whole.out[part.layout.owned_node_slice()]
    got = part.owned_output()
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"rank {rank}/{world} mesh {nx}x{ny}x{nz} owned planes {part.layout.p_own1 - part.layout.p_own0} rel err {err:.2e}", flush=True)
    ok = ok and err == 0.0
    # the same apply with the owned planes living in pinned host memory (bench.py's e2e leg at N > 1)
    L = part.layout
    own = slice(L.p_own0 * L.plane, L.p_own1 * L.plane)
    host_in = torch.empty(part.phi[own].numel(), dtype=torch.float64).pin_memory()
    host_in.copy_(part.phi[own])
    host_out = torch.empty_like(host_in).pin_memory()
    part.phi.zero_()
    part.out.zero_()
    torch.cuda.synchronize()
    part.step_from_host(host_in, host_out, nchunks=5)
    torch.cuda.synchronize()
    err_h = float((host_out - ref.cpu()).abs().max() / ref.abs().max())
    print(f"rank {rank}/{world} mesh {nx}x{ny}x{nz} step_from_host rel err {err_h:.2e}", flush=True)
    ok = ok and err_h == 0.0
# K8 with peer-read halos: one launch per apply, the producer warp copies the halo planes from the neighbours' memory
for (nx, ny, nz) in ((40, 33, 29), (70, 20, 64), (41, 32, 30), (512, 64, 96)):
    rng = np.random.default_rng(1)
    h = (rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(nz) + 0.5)
    part = SlabDCNodal(nx, ny, nz, rank, world, h=h, peer_halo=True)
    whole = SlabDCNodal(nx, ny, nz, 0, 1, h=h)
    whole.step()
    torch.cuda.synchronize()
    ref = whole.out[part.layout.owned_node_slice()]
    for mode in ("barrier", "kernel"):          # ordering by a barrier in front / by signals inside the kernel
        part.out.zero_()
        for _ in range(3):
            part.step_peer(sync=mode)
        torch.cuda.synchronize()
        err = float((part.owned_output() - ref).abs().max() / ref.abs().max())
        print(f"rank {rank}/{world} dc PEER-READ halos ({mode}) {nx}x{ny}x{nz} rel err {err:.2e} "
              f"timed out {part.peer_sync_timed_out()}", flush=True)
        ok = ok and err == 0.0 and not part.peer_sync_timed_out()
    # a field that CHANGES between the applies: the in-kernel ordering must deliver the neighbours' new planes
    for it in range(4):
        dist.barrier()                          # the neighbours have finished reading the previous field
        part.phi.mul_(1.0 + 0.25 * (it + 1))
        whole.phi.mul_(1.0 + 0.25 * (it + 1))
        part.step_peer(sync="kernel")
        whole.step()
        torch.cuda.synchronize()
        ref2 = whole.out[part.layout.owned_node_slice()]
        err = float((part.owned_output() - ref2).abs().max() / ref2.abs().max())
        ok = ok and err == 0.0
    print(f"rank {rank}/{world} dc PEER-READ halos, field rescaled between applies: last rel err {err:.2e}", flush=True)
    dist.barrier()
    del part
for (nx, ny, nz) in ((40, 33, 29), (70, 20, 64), (300, 64, 96)):
    rng = np.random.default_rng(2)
    h = (rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(nz) + 0.5)
    part = SlabMaxwell(nx, ny, nz, rank, world, h=h)
    for _ in range(2):
        part.step()
    torch.cuda.synchronize()
    whole = SlabMaxwell(nx, ny, nz, 0, 1, h=h)
    whole.step()
    torch.cuda.synchronize()
    L, esw = part.layout, whole.es
    g0, g1 = L.c0, L.c1 + (0 if L.has_upper else 1)
    ref = (whole.out[g0 * esw.px:g1 * esw.px], whole.out[esw.off_y + g0 * esw.py:esw.off_y + g1 * esw.py],
           whole.out[esw.off_z + L.c0 * esw.pz:esw.off_z + L.c1 * esw.pz])
    errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(part.owned_parts(part.out), ref)]
    print(f"rank {rank}/{world} maxwell {nx}x{ny}x{nz} rel err Ex/Ey/Ez {errs}", flush=True)
    ok = ok and max(errs) <= 1e-14  # FMA contraction differs between the unrolled ring slots
# peer-read halos (one launch per apply, no send / recv): even-sized meshes only (tensor-map path)
for (nx, ny, nz) in ((40, 32, 29), (64, 20, 64), (256, 64, 96)):
    rng = np.random.default_rng(4)
    h = (rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(nz) + 0.5)
    part = SlabMaxwell(nx, ny, nz, rank, world, h=h, peer_halo=True)
    whole = SlabMaxwell(nx, ny, nz, 0, 1, h=h)
    L, esw = part.layout, whole.es
    g0, g1 = L.c0, L.c1 + (0 if L.has_upper else 1)

    def owned_of_whole():
        return (whole.out[g0 * esw.px:g1 * esw.px], whole.out[esw.off_y + g0 * esw.py:esw.off_y + g1 * esw.py],
                whole.out[esw.off_z + L.c0 * esw.pz:esw.off_z + L.c1 * esw.pz])

    whole.step()
    torch.cuda.synchronize()
    for mode in ("barrier", "kernel"):          # ordering by a barrier in front / by signals inside the kernel
        part.out.zero_()
        for _ in range(3):
            part.step_peer(sync=mode)
        torch.cuda.synchronize()
        errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(part.owned_parts(part.out), owned_of_whole())]
        print(f"rank {rank}/{world} maxwell PEER-READ halos ({mode}) {nx}x{ny}x{nz} rel err Ex/Ey/Ez {errs} "
              f"timed out {part.peer_sync_timed_out()}", flush=True)
        ok = ok and max(errs) <= 1e-14 and not part.peer_sync_timed_out()
    for it in range(4):                         # a field that changes between the applies
        dist.barrier()
        part.e.mul_(1.0 + 0.25 * (it + 1))
        whole.e.mul_(1.0 + 0.25 * (it + 1))
        part.step_peer(sync="kernel")
        whole.step()
        torch.cuda.synchronize()
        errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(part.owned_parts(part.out), owned_of_whole())]
        ok = ok and max(errs) <= 1e-14
    print(f"rank {rank}/{world} maxwell PEER-READ halos, field rescaled between applies: last rel err {errs}", flush=True)
    dist.barrier()
    del part
# the same operators with USER models and a global field handed over by every rank (the mesh-level API of a
# distributed apply): compared with the single-GPU operator of the public TensorMesh API
import discretize_b200 as dz

for (nx, ny, nz) in ((36, 20, 17), (64, 32, 24)):
    rng = np.random.default_rng(3)
    h = (rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(nz) + 0.5)
    mesh = dz.TensorMesh(list(h))
    sigma, mui = np.exp(rng.standard_normal(mesh.nC)), np.exp(rng.standard_normal(mesh.nC))
    e, phi = rng.standard_normal(mesh.nE), rng.standard_normal(mesh.nN)
    C, G = mesh.edge_curl, mesh.nodal_gradient
    ref_m = (C.T @ mesh.get_face_inner_product(mui) @ C + mesh.get_edge_inner_product(sigma)) @ e
    ref_d = (G.T @ mesh.get_edge_inner_product(sigma) @ G) @ phi
    pm = SlabMaxwell(nx, ny, nz, rank, world, h=h, sigma=sigma, mui=mui)
    slices, parts = pm.matvec_global(e)
    errs = [float(np.abs(p.cpu().numpy() - ref_m[sl]).max() / np.abs(ref_m).max()) for sl, p in zip(slices, parts)]
    pd = SlabDCNodal(nx, ny, nz, rank, world, h=h, sigma=sigma)
    sl, own = pd.matvec_global(phi)
    errs.append(float(np.abs(own.cpu().numpy() - ref_d[sl]).max() / np.abs(ref_d).max()))
    print(f"rank {rank}/{world} user models {nx}x{ny}x{nz} rel err Ex/Ey/Ez/dc {errs}", flush=True)
    ok = ok and max(errs) <= 1e-14
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

"""get_interpolation_matrix partitioned across the GPUs of one box ("Interpolation" row).

Rows of Q are independent points, so the partition is by contiguous point ranges: rank p locates and applies
its own rows; the grid vector ``v`` is replicated (1-3 GB at 512^3).  ``Q @ v`` needs no communication at all;
``Q.T @ x`` produces a grid-sized partial result per rank, summed by ONE all-reduce (NCCL; NVLS on NVSwitch).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def point_range(npts: int, world: int, rank: int):
    """Rows [r0, r1) of rank ``rank`` (balanced, contiguous)."""
    base, rem = divmod(int(npts), int(world))
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


class ShardedInterpolation:
    """This rank's rows of ``mesh.get_interpolation_matrix(points, location_type)``.

    ``points`` is the full (npts, 3) array (numpy or tensor; every rank passes the same); only the rank's own
    range is uploaded and located.  ``apply(v)`` -> values at this rank's points (CUDA tensor, no communication);
    ``apply_t(x_local)`` -> the full grid-sized ``Q.T @ x`` on every rank (one all-reduce)."""

    def __init__(self, mesh, points, location_type="cell_centers", zeros_outside=False, rank=None, world=None,
                 store=True):
        from .interp import make_interpolation_operator

        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank, self.world = int(rank), int(world)
        self.npts = int(points.shape[0])
        self.r0, self.r1 = point_range(self.npts, self.world, self.rank)
        self.Q = make_interpolation_operator(mesh, points[self.r0:self.r1], location_type, zeros_outside, store=store)
        self.shape = (self.npts, self.Q.shape[1])

    def apply(self, v, out=None):
        return self.Q.apply_device(v, out=out)

    def apply_t(self, x_local, out=None, group=None):
        y = self.Q.apply_device(x_local, transpose=True, out=out)
        if self.world > 1:
            dist.all_reduce(y, group=group)
        return y

    def gather_rows(self, y_local, group=None):
        """All ranks' values concatenated in point order (collective; checks and output, not the data path)."""
        if self.world == 1:
            return y_local
        out = torch.zeros(self.npts, dtype=y_local.dtype, device=y_local.device)
        out[self.r0:self.r1] = y_local
        dist.all_reduce(out, group=group)
        return out

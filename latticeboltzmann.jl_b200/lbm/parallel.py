"""y-slab decomposition plumbing: one process per GPU, `torch.distributed` for the host-side
exchange of small objects (NCCL id, partial sums).  The halo exchange itself happens inside
liblbm_b200.so (NCCL send/recv on a side stream); nothing here touches population data.
"""
import os

import numpy as np


def slab_rows(ny, rank, world):
    """Rows owned by `rank`: (y0, ny_local).  Must match lbm_create (csrc/lbm_b200.cu)."""
    base, rem = divmod(ny, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def halo_rows_per_direction(q):
    """Row-equivalents sent to each neighbour per step: sum of c_y over populations with c_y > 0."""
    cy = q.abscissae[1]
    return int(cy[cy > 0].sum())


class SlabComm:
    """Wraps an initialised torch.distributed process group (nccl on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = int(os.environ.get("LOCAL_RANK", self.rank))

    def broadcast_bytes(self, payload, src=0):
        obj = [payload if self.rank == src else None]
        self.dist.broadcast_object_list(obj, src=src, group=self.group)
        return obj[0]

    def allreduce_sum(self, values):
        import torch
        t = torch.as_tensor(np.asarray(values, dtype=np.float64))
        backend = self.dist.get_backend(self.group)
        if backend == "nccl":
            t = t.cuda(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def barrier(self):
        self.dist.barrier(group=self.group)

    def nccl_id(self):
        from . import _abi
        return self.broadcast_bytes(_abi.nccl_unique_id() if self.rank == 0 else None)

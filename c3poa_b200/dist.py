"""Multi-GPU plumbing: one process per GPU, reads sharded by rank, NO data-path collective
(every read is independent end to end: /root/reference/C3POa.py:112,245).  torch.distributed is
used only for the barrier and the max/sum reductions of the timing scalars."""
from __future__ import annotations

import os


def env_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class Group:
    """Thin wrapper; world == 1 never touches torch."""

    def __init__(self, backend="nccl", device=None):
        self.rank, self.local_rank, self.world = env_world()
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch
            import torch.distributed as dist
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self.device = torch.device("cuda", self.local_rank)
                dist.init_process_group("nccl", device_id=self.device)
            else:
                self.device = torch.device("cpu")
                dist.init_process_group(backend)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x, op):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def allmax(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX if self.dist else None)

    def allsum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM if self.dist else None)

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None

"""One process per GPU: rank discovery, barrier and max/sum reductions for the benchmark.

The hot path itself needs no collective: chunk pairs are independent, every rank owns a shard and
returns its own match list.  torch.distributed is plumbing for timing only."""
from __future__ import annotations

import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_seed(base_seed: int, rank: int) -> int:
    """Weak scaling: every rank generates its own pairs from a rank-specific seed."""
    return base_seed * 1000 + rank


def shard_range(n_total: int, rank: int, world: int):
    """Strong-scaling helper: contiguous slice [lo, hi) of n_total units for this rank."""
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def shard_blocks_by_target(blocks, n_targets: int, rank: int, world: int):
    """Block mode: every rank owns a contiguous range of target chunks -- and keeps the spectra of
    exactly those resident in HBM -- so a t_pair block (tFrom, tTo, qFrom, qTo, fast), inclusive ranges, is clipped
    to the rank's range; blocks that straddle a boundary are split between the neighbours.  The union over the
    ranks is the original work list, pair for pair, without duplicates."""
    lo, hi = shard_range(n_targets, rank, world)
    out = []
    for b in blocks:
        t0, t1 = max(int(b[0]), lo), min(int(b[1]), hi - 1)
        if t0 <= t1:
            out.append((t0, t1, int(b[2]), int(b[3]), int(b[4]) if len(b) > 4 else 0))
    return out


class Group:
    """Thin wrapper: no-ops when world == 1."""

    def __init__(self, backend: str = "nccl", device=None):
        self.rank, self.local_rank, self.world = env_rank()
        self.backend = backend
        self.device = device
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            if not dist.is_initialized():
                kw = {}
                if backend == "nccl" and device is not None:
                    kw["device_id"] = device
                dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world, **kw)
            self.dist = dist

    def _tensor(self, v):
        import torch

        dev = self.device if (self.backend == "nccl" and self.device is not None) else "cpu"
        return torch.tensor([float(v)], dtype=torch.float64, device=dev)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, v: float) -> float:
        if self.dist is None:
            return float(v)
        t = self._tensor(v)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v: float) -> float:
        if self.dist is None:
            return float(v)
        t = self._tensor(v)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()

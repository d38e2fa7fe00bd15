"""Multi-GPU host logic.  Environments are independent, so the batch shards trivially: rank r of W
owns one contiguous slice of global env ids (cut on 32-env state blocks) and steps it with its own
`BatchedSmartNanogridEnv(env_gid0=lo)`; schedule sampling is keyed by global env id, so results do not
depend on W.  Nothing is exchanged on the step path.  The only collective is the optional reduction of
episode-return statistics below (a handful of floats per logging interval)."""
from __future__ import annotations

import dataclasses
import math

import torch
import torch.distributed as dist

BLOCK = 32   # envs per state block (include/sng.h sng_layout.env_block)


def shard_range(total_envs: int, world_size: int, rank: int):
    """[lo, hi) of global env ids owned by `rank`: equal shares rounded to whole state blocks."""
    blocks = (total_envs + BLOCK - 1) // BLOCK
    base, extra = divmod(blocks, world_size)
    b_lo = rank * base + min(rank, extra)
    b_hi = b_lo + base + (1 if rank < extra else 0)
    return min(b_lo * BLOCK, total_envs), min(b_hi * BLOCK, total_envs)


@dataclasses.dataclass
class ReturnStats:
    """Sufficient statistics of finished-episode returns: count, sum, sum of squares, min, max."""
    count: float
    total: float
    total_sq: float
    min: float
    max: float

    @classmethod
    def from_returns(cls, returns: torch.Tensor) -> "ReturnStats":
        r = returns.detach().double()
        if r.numel() == 0:
            return cls(0.0, 0.0, 0.0, math.inf, -math.inf)
        return cls(float(r.numel()), float(r.sum()), float((r * r).sum()), float(r.min()), float(r.max()))

    def all_reduce(self, group=None, device=None) -> "ReturnStats":
        """Combine over all ranks (SUM for the moments, MIN / MAX for the extremes): two tiny collectives.
        With NCCL pass the rank's CUDA device; without an initialised process group this is a no-op."""
        if not (dist.is_available() and dist.is_initialized()):
            return self
        sums = torch.tensor([self.count, self.total, self.total_sq], dtype=torch.float64, device=device)
        ext = torch.tensor([-self.min, self.max], dtype=torch.float64, device=device)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
        return ReturnStats(float(sums[0]), float(sums[1]), float(sums[2]), -float(ext[0]), float(ext[1]))

    @property
    def mean(self) -> float:
        return self.total / self.count if self.count else float("nan")

    @property
    def std(self) -> float:
        if not self.count:
            return float("nan")
        return math.sqrt(max(self.total_sq / self.count - self.mean ** 2, 0.0))

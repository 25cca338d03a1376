"""Data-parallel plumbing: the path shards by window; ranks exchange ONE buffer per step.

The reference has no distributed code (run_all.sh only farms seeds to GPUs).  Here rank r takes an equal
shard of the global (already shuffled) index list, runs the fused kernels with GLOBAL loss denominators
(sum_b w[y_b] over the whole batch -- labels are replicated, so no communication), and the ranks all-reduce
(sum) gbuf = [G | private grads | loss | correct] before the deterministic solve + clip + SGD that every rank
repeats identically.  CAGrad is non-linear in the per-task gradients, so the exchange must carry G, not the
combined gradient."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Equal shards, remainder dropped (mean of means == global mean; weargait semantics are global-batch)."""
    per = n // world
    return rank * per, (rank + 1) * per


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def global_denominators(ys_global, cls_weights) -> torch.Tensor:
    """sum_b w[y_b] per stream over the GLOBAL label vectors (what gaitk_loss_denominators computes on device)."""
    return torch.stack([w[y].sum() if w is not None else torch.tensor(float(y.numel())) for y, w in zip(ys_global, cls_weights)])

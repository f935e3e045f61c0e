"""Data-parallel plumbing for the CLR path.

The batch shards across GPUs; every prototype is ``S_r / N_r`` with sums that decompose over any
partition of the pixels, so the only exchange is ONE all-reduce(sum) of the packed
``[domains][2K][C+1]`` buffer per step (8 KB at C=256, K=2).  The backward needs no collective:
``dL/dmu`` is identical on every rank (the losses are computed from replicated prototypes), so each
rank writes ``grad_x = sum_r (g_r / N_r^global) w_r`` for its own pixels, scaled by ``grad_scale``
(= world size under DDP, whose gradient *averaging* then reproduces the single-process sum).

One process per GPU; ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests)
carries the message.  The reference itself is single-GPU (no collective anywhere).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

_STATE = {"enabled": False, "group": None, "grad_scale": None}


def enable(group=None, grad_scale: Optional[float] = None) -> None:
    """Turn on the cross-rank all-reduce of the packed sums.

    ``grad_scale=None`` -> world size (DDP averages gradients); pass 1.0 when gradients are summed.
    """
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _STATE.update(enabled=True, group=group, grad_scale=grad_scale)


def disable() -> None:
    _STATE.update(enabled=False, group=None, grad_scale=None)


def enabled() -> bool:
    return bool(_STATE["enabled"])


def world_size() -> int:
    if not enabled():
        return 1
    import torch.distributed as dist
    return dist.get_world_size(_STATE["group"])


def grad_scale() -> float:
    gs = _STATE["grad_scale"]
    return float(world_size()) if gs is None else float(gs)


def all_reduce_sums(packed: torch.Tensor) -> torch.Tensor:
    """In-place sum of the packed ``[...][2K][C+1]`` buffer over the group (no-op when disabled)."""
    if enabled():
        import torch.distributed as dist
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=_STATE["group"])
    return packed


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard ``[lo, hi)`` of ``n`` samples for ``rank`` (remainder to the low ranks)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    per, rem = divmod(n, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int):
    """Slice every ``[B, ...]`` tensor to this rank's shard of the batch."""
    out = []
    for t in tensors:
        lo, hi = shard_bounds(t.shape[0], rank, world)
        out.append(t[lo:hi])
    return out

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

_STATE = {"enabled": False, "group": None, "grad_scale": None, "peer": False}


def enable(group=None, grad_scale: Optional[float] = None) -> None:
    """Turn on the cross-rank all-reduce of the packed sums.

    ``grad_scale=None`` -> world size (DDP averages gradients); pass 1.0 when gradients are summed.
    """
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _STATE.update(enabled=True, group=group, grad_scale=grad_scale, peer=False)


def disable() -> None:
    _STATE.update(enabled=False, group=None, grad_scale=None, peer=False)


# ------------------------------------------------------------------------------------------------- in-kernel exchange
# Instead of two NCCL all-reduces per step (launch + ~10 us each, and a kernel boundary on either side), the ranks of
# one NVLink domain can let the step's finish stages exchange the packed sums themselves: every rank owns a small
# receive buffer that all peers map through CUDA IPC; senders store (value, sequence) words straight into it over
# NVLink and each rank sums the contributions in rank order (include/clr_b200.h, clr_step_args.world ...).
_PEER = {}      # (K, C, world) -> dict(local=ptr, ptrs=[...], seq=int)


def enable_peer(group=None, grad_scale: Optional[float] = None) -> None:
    """Like :func:`enable`, but the fused step (``CLRStep`` / ``CLRPlan``) exchanges its sums inside its own kernels
    over peer-mapped memory (single node, world size <= 8).  The drop-in ops keep using the NCCL all-reduce."""
    enable(group, grad_scale)
    _STATE["peer"] = True


def peer_enabled() -> bool:
    return bool(_STATE.get("peer")) and enabled() and world_size() > 1


def rank() -> int:
    import torch.distributed as dist
    return dist.get_rank(_STATE["group"]) if enabled() else 0


def peer_buffers(K: int, C: int, device) -> dict:
    """Collective (every rank must call it with the same K, C): allocate this rank's receive buffer, exchange the IPC
    handles through the process group and map every peer's buffer.  Cached per (K, C, world)."""
    import ctypes
    import torch.distributed as dist
    from . import _lib
    world = world_size()
    key = (K, C, world)
    if key in _PEER:
        return _PEER[key]
    lib = _lib.load()
    if world > _lib.CLR_MAX_WORLD:
        raise RuntimeError("in-kernel exchange supports at most %d ranks (got %d): use dist.enable()" % (_lib.CLR_MAX_WORLD, world))
    nbytes = lib.clr_step_xchg_bytes(world, K, C)
    with torch.cuda.device(device):
        local = ctypes.c_void_p()
        _lib.check(lib.clr_peer_alloc(nbytes, ctypes.byref(local)), "clr_peer_alloc")
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.clr_peer_export(local, handle), "clr_peer_export")
        mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=device)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=_STATE["group"])
        me = rank()
        ptrs = []
        for q in range(world):
            if q == me:
                ptrs.append(local.value)
            else:
                pq = ctypes.c_void_p()
                _lib.check(lib.clr_peer_open(bytes(gathered[q].cpu().tolist()), ctypes.byref(pq)), "clr_peer_open")
                ptrs.append(pq.value)
        torch.cuda.synchronize(device)
        dist.barrier(group=_STATE["group"])      # nobody sends before every buffer is mapped and zeroed
    _PEER[key] = dict(local=local.value, ptrs=ptrs, seq=0, world=world, rank=me, bytes=nbytes)
    return _PEER[key]


def close_peer() -> None:
    """Unmap every peer's receive buffer and free the local ones (collective in spirit: call it on every rank once no
    step is in flight, before ``destroy_process_group``).  Plans built on the freed buffers must not run afterwards."""
    from . import _lib
    if not _PEER:
        return
    lib = _lib.load()
    torch.cuda.synchronize()
    for peer in _PEER.values():
        for q, pq in enumerate(peer["ptrs"]):
            if q != peer["rank"] and pq:
                lib.clr_peer_close(pq)
        lib.clr_peer_free(peer["local"])
    _PEER.clear()


class local:
    """Context manager: run the enclosed steps as a single-process job (no exchange), e.g. a rank that recomputes the
    whole batch as a parity reference next to the sharded step."""

    def __enter__(self):
        self._saved = dict(_STATE)
        _STATE.update(enabled=False, peer=False)
        return self

    def __exit__(self, *exc):
        _STATE.update(self._saved)
        return False


def next_seq(peer: dict) -> int:
    """Sequence number of the next exchange on these buffers (identical on every rank: one per fused step)."""
    peer["seq"] = (peer["seq"] % 0xFFFFFFFE) + 1
    return peer["seq"]


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

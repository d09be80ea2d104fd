"""Image sharding across the GPUs of one box: images of a batch never interact (GroupNorm, LayerNorm, SCA
and attention all reduce within one image), so rank r of G simply takes a contiguous slice of the batch, runs the whole
path on it with replicated weights, and the finished (b,1,H,W) planes are all-gathered once.  There is no collective
inside the sampler loop.  Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's slice: the first (batch % world) ranks take one extra image; ranks past the batch get an empty slice."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_outputs(y_local: torch.Tensor, batch: int, rank: int, world: int, group=None) -> torch.Tensor:
    """All-gather the per-rank output slices back into the full (batch, ...) tensor, in batch order, on every rank.
    Slices may be ragged (batch % world != 0): every rank contributes a buffer padded to the largest slice."""
    import torch.distributed as dist
    if world == 1:
        return y_local
    sizes = [shard_bounds(batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def run_sharded(model: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor, rank: int, world: int, group=None) -> torch.Tensor:
    """model(x) computed as G independent slices + one output gather."""
    xs = shard_batch(x, rank, world)
    ys = model(xs) if xs.shape[0] > 0 else x.new_zeros((0,) + tuple(x.shape[1:]))
    return gather_outputs(ys, x.shape[0], rank, world, group)

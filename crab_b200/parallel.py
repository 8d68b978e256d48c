"""Data-parallel plumbing for the one place the path shards: samples are independent, so each rank
runs the full path on its slice with a full weight replica and the only exchange is a final all-gather of the
generated ids / last-step logits.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the transport."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_range(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of `n_samples` for `rank`; the first n % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_samples, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(items: Sequence, rank: int, world: int) -> List:
    lo, hi = shard_range(len(items), rank, world)
    return list(items[lo:hi])


def gather_rows(local: torch.Tensor, n_total: int, rank: int, world: int, pad_value: int = 0) -> torch.Tensor:
    """All-gather per-rank result rows (ids (b_local, n) or logits (b_local, V)) back into sample order.  Ragged shards
    are padded to the largest shard for the collective and trimmed afterwards."""
    import torch.distributed as dist

    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    buf = local
    if local.shape[0] < mx:
        pad = torch.full((mx - local.shape[0],) + tuple(local.shape[1:]), pad_value, dtype=local.dtype, device=local.device)
        buf = torch.cat([local, pad], 0)
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf.contiguous())
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(world)], 0)

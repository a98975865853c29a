"""Data-parallel inference across the GPUs of one box: one process per GPU, contiguous blocks of clips per rank,
replicated weights, no collective inside the forward; NCCL (or gloo in CPU tests) only gathers frame logits.

The reference's own multi-GPU inference is `torch.nn.DataParallel` (speechbrain/core.py:1164-1169): the batch is
split into per-GPU sub-batches, each normalised over its own sub-batch by the whole-tensor layer norms.  That is
exactly the semantics of sharding by clip here ("DP-equivalent", SURVEY.md 8e option i)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of items owned by `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_logits(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank logits (n_local, T, C) into (n_total, T, C) in clip order on every rank.
    Ranks may own different clip counts; blocks are padded to the largest one for the collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    counts = [shard_range(n_total, r, world) for r in range(world)]
    n_max = max(b - a for a, b in counts)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: counts[r][1] - counts[r][0]] for r in range(world)], dim=0)


def gather_notes(local_notes: List, group=None) -> List:
    """Gather python note lists (one entry per locally decoded song) from every rank, in rank order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(local_notes)
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, list(local_notes), group=group)
    return [song for part in out for song in part]

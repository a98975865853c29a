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


def gather_ragged(local: List[torch.Tensor], n_total: int, group=None) -> List[torch.Tensor]:
    """All-gather a list of per-item tensors whose first dimension differs from item to item (e.g. the frame logits of a
    song's windows: the last window is shorter) -> the n_total tensors in global item order on every rank.
    `local` holds this rank's items of shard_range(n_total, rank, world), all with the same trailing dimensions."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return list(local)
    rank = dist.get_rank(group)
    a, b = shard_range(n_total, rank, world)
    if len(local) != b - a:
        raise ValueError(f"rank {rank} owns items [{a}, {b}) but passed {len(local)} tensors")
    ref = local[0] if local else None
    # lengths first (one small collective), then the payload padded to the longest item
    if ref is not None:
        dev = ref.device
    else:  # a rank that owns no item still takes part in the collectives
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    lens = torch.zeros(n_total, dtype=torch.int64, device=dev)
    for i, t in enumerate(local):
        lens[a + i] = t.shape[0]
    dist.all_reduce(lens, op=dist.ReduceOp.SUM, group=group)
    meta = [None] * world
    dist.all_gather_object(meta, None if ref is None else (tuple(ref.shape[1:]), str(ref.dtype)), group=group)
    tail, dtype_name = next(m for m in meta if m is not None)
    dtype = getattr(torch, dtype_name.replace("torch.", ""))
    n_max = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    l_max = int(lens.max().item())
    pad = torch.zeros((n_max, l_max) + tuple(tail), dtype=dtype, device=dev)
    for i, t in enumerate(local):
        pad[i, : t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out: List[torch.Tensor] = []
    for r in range(world):
        ra, rb = shard_range(n_total, r, world)
        for i in range(rb - ra):
            out.append(bufs[r][i, : int(lens[ra + i].item())])
    return out

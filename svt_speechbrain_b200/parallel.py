"""Data-parallel inference across the GPUs of one box: one process per GPU, contiguous blocks of clips per rank,
replicated weights, no collective inside the forward; NCCL (or gloo in CPU tests) only gathers frame logits.

The reference's own multi-GPU inference is `torch.nn.DataParallel` (speechbrain/core.py:1164-1169): the batch is
split into per-GPU sub-batches, each normalised over its own sub-batch by the whole-tensor layer norms.  That is
exactly the semantics of sharding by clip here ("DP-equivalent", SURVEY.md 8e option i)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of items owned by `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def balanced_shares(n_total: int, seconds_per_item: Sequence[float]) -> List[int]:
    """Split n_total items over ranks in proportion to their measured speed (1 / seconds_per_item), largest-remainder
    rounding, at least one item per rank.  The GPUs of one box do not run at one speed under the power cap (8 GPUs of one
    B200 box: 25.5 ... 27.7 ms for the same 64 clips): an even split runs every step at the pace of the slowest GPU, a
    speed-proportional split at the pace of their mean."""
    rates = [1.0 / max(float(t), 1e-12) for t in seconds_per_item]
    total = sum(rates)
    ideal = [n_total * r / total for r in rates]
    shares = [max(1, int(x)) for x in ideal]
    order = sorted(range(len(rates)), key=lambda i: ideal[i] - int(ideal[i]), reverse=True)
    i = 0
    while sum(shares) < n_total:
        shares[order[i % len(order)]] += 1
        i += 1
    while sum(shares) > n_total:  # only after the max(1, .) floor: take back from the most over-served rank that can give
        cand = [k for k in range(len(shares)) if shares[k] > 1]
        if not cand:
            raise ValueError(f"cannot give each of {len(shares)} ranks one of {n_total} items")
        j = max(cand, key=lambda k: shares[k] - ideal[k])
        shares[j] -= 1
    return shares


def gather_logits(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank logits (n_local, T, C) into (n_total, T, C) in clip order on every rank.
    Equal blocks (n_total % world == 0) go through one `all_gather_into_tensor` straight into the result; ragged blocks
    are padded to the largest one for the collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    counts = [shard_range(n_total, r, world) for r in range(world)]
    if n_total % world == 0:
        if local.shape[0] != n_total // world:
            raise ValueError(f"rank owns {n_total // world} clips but passed {local.shape[0]}")
        out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    n_max = max(b - a for a, b in counts)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: counts[r][1] - counts[r][0]] for r in range(world)], dim=0)


class LogitsGatherer:
    """Serving-loop gather of equal per-rank logits blocks that never stalls the next forward.

    The reference's multi-GPU inference (nn.DataParallel, speechbrain/core.py:1164-1169) gathers the outputs of every step
    on the host thread, a global rendezvous per step.  Here step k's `ncclAllGather` is issued asynchronously (c10d runs it
    on its own stream, ordered after the producing kernels by an event) into one of `depth` pre-allocated
    (world * n_local, T, C) buffers, and the compute stream only waits for it when the result is consumed or when the
    buffer comes round again -- so step k + 1's forward is queued right behind step k's and a rank that is momentarily
    slower delays nobody until `depth` steps later.

        g = LogitsGatherer((B, T, 20), depth=2, device=dev)
        for k, wav in enumerate(batches):
            eng.forward(wav, want_feats=False, want_logits=True, logits_out=g.local(k))
            g.submit(k)                      # asynchronous all-gather of slot k % depth
            if k >= 1: use(g.result(k - 1))  # (world * B, T, 20), blocks the CURRENT STREAM only, not the host
    """

    def __init__(self, local_shape, depth: int = 2, device=None, dtype=torch.float32, group=None):
        """local_shape[0] is the LARGEST per-rank block; ranks with a smaller share (balanced_shares) write only their first
        rows (`local(k, n)`) and the consumer reads rank r's block at rows [r * local_shape[0], + share_r) of `result(k)`."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.depth = depth
        self._local = [torch.empty(tuple(local_shape), dtype=dtype, device=device) for _ in range(depth)]
        self._out = [torch.empty((self.world * local_shape[0],) + tuple(local_shape[1:]), dtype=dtype, device=device)
                     for _ in range(depth)] if self.world > 1 else self._local
        self._work = [None] * depth
        self._step = [-1] * depth

    def local(self, k: int, n: int = None) -> torch.Tensor:
        """The buffer step k's forward writes its logits into (its first n rows when this rank's share is smaller than the
        largest one).  Re-using a slot first drains the gather that read it."""
        s = k % self.depth
        self._drain(s)
        return self._local[s] if n is None else self._local[s][:n]

    def submit(self, k: int) -> None:
        s = k % self.depth
        self._step[s] = k
        if self.world > 1:
            self._work[s] = dist.all_gather_into_tensor(self._out[s], self._local[s], group=self.group, async_op=True)

    def result(self, k: int) -> torch.Tensor:
        s = k % self.depth
        if self._step[s] != k:
            raise ValueError(f"step {k} is not in flight (slot holds step {self._step[s]})")
        self._drain(s)
        return self._out[s]

    def _drain(self, s: int) -> None:
        if self._work[s] is not None:
            self._work[s].wait()  # stream-level wait for NCCL; host-blocking for gloo (CPU tests)
            self._work[s] = None

    def finish(self) -> None:
        for s in range(self.depth):
            self._drain(s)


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

"""world_size-2 gloo test of the multi-GPU host logic (clip sharding + logits / note gathers)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from svt_speechbrain_b200.parallel import gather_logits, gather_notes, shard_range

    a, b = shard_range(n_total, rank, world)
    full = torch.arange(n_total * 3 * 20, dtype=torch.float32).view(n_total, 3, 20)
    got = gather_logits(full[a:b].clone(), n_total)
    notes = gather_notes([[[0.0, 1.0, 60 + i]] for i in range(a, b)])
    q.put((rank, torch.equal(got, full), [n[0][2] for n in notes]))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gather_logits_uneven_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_total, world = 7, 2
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, pitches in res:
        assert ok, rank
        assert pitches == [60 + i for i in range(n_total)]

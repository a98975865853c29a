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


def _gatherer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from svt_speechbrain_b200.parallel import LogitsGatherer, gather_logits

    B, T, C, steps = 3, 5, 20, 5
    g = LogitsGatherer((B, T, C), depth=2, device="cpu")
    ok = True
    for k in range(steps):
        g.local(k).copy_(torch.full((B, T, C), float(100 * k + rank)))   # what the forward of step k would write
        g.submit(k)
        if k >= 1:  # consume one step behind: step k - 1's gather completes while step k is being produced
            got = g.result(k - 1)
            want = torch.cat([torch.full((B, T, C), float(100 * (k - 1) + r)) for r in range(world)])
            ok = ok and torch.equal(got, want)
    ok = ok and torch.equal(g.result(steps - 1)[B:], torch.full((B, T, C), float(100 * (steps - 1) + 1)))
    g.finish()
    try:
        g.result(0)  # its slot was reused by steps 2 and 4
        ok = False
    except ValueError:
        pass
    even = gather_logits(torch.full((2, T, C), float(rank)), 4)  # equal blocks: the all_gather_into_tensor path
    ok = ok and torch.equal(even, torch.cat([torch.zeros(2, T, C), torch.ones(2, T, C)]))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_double_buffered_gatherer_and_even_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gatherer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def test_balanced_shares_apportionment():
    """Speed-proportional sharding of a step's clips (bench.py at N > 1): shares sum to the total, follow the measured
    speeds, never drop a rank, and make the slowest rank's step shorter than the even split does."""
    from svt_speechbrain_b200.parallel import balanced_shares

    t = [26.22, 25.99, 26.11, 25.47, 27.58, 26.86, 27.72, 26.90]   # ms per 64 clips on the 8 GPUs of one box (round-2 run)
    s = balanced_shares(512, t)
    assert sum(s) == 512 and min(s) >= 1
    assert s[3] == max(s) and s[6] == min(s)                        # fastest GPU gets the most clips, slowest the fewest
    assert max(a * b / 64 for a, b in zip(s, t)) < max(t) - 0.5     # the step no longer runs at the slowest GPU's pace
    assert balanced_shares(128, [1.0, 1.0]) == [64, 64]
    assert balanced_shares(7, [1.0, 1.0]) in ([4, 3], [3, 4])
    assert balanced_shares(3, [1.0, 100.0, 100.0]) == [1, 1, 1]
    assert balanced_shares(5, [1.0, 100.0, 100.0]) == [3, 1, 1]
    s = balanced_shares(10, [1.0, 1000.0])
    assert sum(s) == 10 and s[1] >= 1

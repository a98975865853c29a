"""Long-form windowing / stitching arithmetic (BASELINE config 5) and the ragged multi-rank gather, CPU only."""
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svt_speechbrain_b200.amt import FRAME_FIELD, FRAME_HOP, split_song_overlapped, stitch_plan


@pytest.mark.parametrize("n,dur,overlap", [(16000 * 300 + 1234, 10.0, 1.0), (16000 * 30, 10.0, 0.0), (16000 * 47 + 7, 5.0, 2.5),
                                           (16000 * 3, 10.0, 1.0), (16000 * 20 + 399, 10.0, 0.5)])
def test_windows_tile_the_frame_grid(n, dur, overlap):
    w = split_song_overlapped(n, 16000, dur, overlap)
    assert w[0][0] == 0 and w[-1][1] == n and all(a % FRAME_HOP == 0 for a, _ in w)
    assert all(b - a == int(dur * 16000) for a, b in w[:-1])
    plan = stitch_plan(w)
    if overlap == 0.0:
        # the reference's evaluation rule: every window keeps all of its frames (one frame per boundary is never computed)
        assert plan == [(0, (b - a - FRAME_FIELD) // FRAME_HOP + 1) for a, b in w]
        return
    g = 0
    for (a, b), (lo, hi) in zip(w, plan):
        nf = max((b - a - FRAME_FIELD) // FRAME_HOP + 1, 0)
        assert 0 <= lo <= hi <= nf
        if hi > lo:
            assert a // FRAME_HOP + lo == g      # every global frame exactly once, in order
            g = a // FRAME_HOP + hi
    assert g == (n - FRAME_FIELD) // FRAME_HOP + 1  # as many frames as one pass over the whole song would give


def test_kept_frames_stay_away_from_window_edges():
    w = split_song_overlapped(16000 * 100, 16000, 10.0, 2.0)
    plan = stitch_plan(w)
    for i, ((a, b), (lo, hi)) in enumerate(zip(w, plan)):
        nf = (b - a - FRAME_FIELD) // FRAME_HOP + 1
        if 0 < i < len(w) - 1:
            assert lo >= 45 and nf - hi >= 45     # ~1 s = half the overlap dropped on both sides


def _worker(rank, world, port, n_total, q):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from svt_speechbrain_b200.parallel import gather_ragged, shard_range
    lens = [5 + (3 * i) % 4 for i in range(n_total)]
    items = [torch.full((lens[i], 20), float(i)) + torch.arange(lens[i])[:, None] / 100 for i in range(n_total)]
    a, b = shard_range(n_total, rank, world)
    got = gather_ragged([t.clone() for t in items[a:b]], n_total)
    ok = len(got) == n_total and all(torch.equal(x, y) for x, y in zip(got, items))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 1])
def test_gather_ragged_two_ranks(n_total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_tail_shorter_than_one_frame_is_merged_into_the_last_window():
    """A song that ends less than one receptive field after a window boundary: the remainder joins the previous window (as
    the reference's last utterance takes the remainder, prepare_benchmarks.py:124-127) instead of becoming a window without
    frames, which one rank alone would have had to reject."""
    w = split_song_overlapped(320100, 16000, 10.0, 0.0)
    assert w == [(0, 160000), (160000, 320100)]
    w = split_song_overlapped(16000 * 19 + 100, 16000, 10.0, 1.0)   # hop 9 s: third window would start at 18 s with 1.006 s left
    assert all(b - a >= FRAME_FIELD for a, b in w) and w[-1][1] == 16000 * 19 + 100
    n = 16000 * 18 + 200                                            # ... and here with 200 samples left
    w = split_song_overlapped(n, 16000, 10.0, 1.0)
    assert w == [(0, 160000), (144000, n)]
    plan = stitch_plan(w)
    g = 0
    for (a, b), (lo, hi) in zip(w, plan):
        assert a // FRAME_HOP + lo == g
        g = a // FRAME_HOP + hi
    assert g == (n - FRAME_FIELD) // FRAME_HOP + 1


def test_evaluation_driver_batch_plan():
    """plan_song_batches (the host half of AMTTranscriber.transcribe_songs): every utterance of every song exactly once,
    batches hold one clip length each and at most batch_clips clips, utterances follow the reference chunk rule."""
    from svt_speechbrain_b200.amt import AMTHparams, plan_song_batches, split_song

    hp = AMTHparams()  # 5-s utterances
    lengths = [16000 * 12 + 1234, 16000 * 5, 16000 * 30, 16000 * 7 + 3, 16000 * 10]
    n_utt, plan = plan_song_batches(lengths, hp, None, batch_clips=4)
    assert n_utt == [len(split_song(n, hp)) for n in lengths]
    seen = set()
    for batch in plan:
        assert 1 <= len(batch) <= 4
        assert len({b - a for _, _, a, b in batch}) == 1
        for si, ui, a, b in batch:
            assert (si, ui) not in seen and split_song(lengths[si], hp)[ui] == (a, b)
            seen.add((si, ui))
    assert seen == {(si, ui) for si, k in enumerate(n_utt) for ui in range(k)}
    # 80 000-sample utterances: 1 (song 0) + 1 (song 1) + 6 (song 2) + 2 (song 4) = 10 -> batches of 4, 4, 2
    assert sorted(len(b) for b in plan if b[0][3] - b[0][2] == 80000) == [2, 4, 4]

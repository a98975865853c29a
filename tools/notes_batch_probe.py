#!/usr/bin/env python
"""wav -> notes (AMTTranscriber.transcribe_songs, 64 host songs of 10 s) for several batch_clips: smaller batches let the
host decode of one batch overlap the GPU work of the next inside ONE call.   python tools/notes_batch_probe.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import svt_speechbrain_b200 as svt  # noqa: E402
from oracle import wav2vec2_oracle as wo  # noqa: E402  (seeded weights only)

dev = torch.device("cuda", 0)
cfg = wo.W2V2Config.large()
d = bench._lobe_dir(cfg)
lobe = svt.HuggingFaceWav2Vec2(source=d, save_path=d, pretrain=False, output_norm=True, freeze=True)
lobe.load_state_dict(wo.random_weights(cfg, seed=0), strict=True)
lin = svt.Linear(n_neurons=20, input_size=cfg.hidden_size)
lin.load_state_dict(wo.random_head(cfg.hidden_size, 20, seed=0))
tr = svt.AMTTranscriber(lobe.to(dev), lin.to(dev), svt.AMTHparams(dur_threshold=10.0), device=dev)
host = torch.randn(64, 160000).pin_memory()
songs = [host[c] for c in range(64)]
ref = None
for rep in range(2):
    for bc in (64, 32, 16):
        for _ in range(2):
            notes = tr.transcribe_songs(songs, dur=10.0, batch_clips=bc)
        ts = []
        for _ in range(7):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            notes = tr.transcribe_songs(songs, dur=10.0, batch_clips=bc)
            ts.append(1e3 * (time.perf_counter() - t0))
        ts.sort()
        if ref is None:
            ref = notes
        same = all((a.shape == b.shape and (a == b).all()) for a, b in zip(ref, notes))
        print(f"rep {rep} batch_clips={bc}: median {ts[3]:.2f} ms per 64 songs = {640 / ts[3] * 1e3:.0f} audio-s/s, all {[round(t, 2) for t in ts]}, "
              f"notes equal to batch_clips=64: {same}", flush=True)

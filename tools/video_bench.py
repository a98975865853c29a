#!/usr/bin/env python
"""Timing of the AV-HuBERT video stream (BASELINE config 4 per-GPU share: 4 clips x 500 frames of 88 x 88)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import avhubert_oracle as av  # noqa: E402  (seeded weights only)
from svt_speechbrain_b200._lib import VideoConfig, lib  # noqa: E402
from svt_speechbrain_b200.engine import VideoEngine  # noqa: E402

dev = torch.device("cuda", 0)
cfg = av.AVHubertConfig()
sd = av.random_weights(cfg, seed=0)
eng = VideoEngine(VideoConfig(1024, 24, 16, 4096, 128, 16, 1e-5, 0, 1), dev)
eng.load(sd)
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 500
video = torch.randn(B, 1, T, 88, 88, device=dev)
for _ in range(2):
    eng.forward(video)
torch.cuda.synchronize()
n0 = lib().svt_debug_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for _ in range(n):
    eng.forward(video)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
gf = 653.7 * B  # GFLOP per 10-s clip
print(f"video stream B={B} T={T}: {ms:.2f} ms/step = {B * 10 / ms * 1e3:.0f} video-s/s, {gf / ms:.0f} TFLOP/s algorithmic, "
      f"{(lib().svt_debug_launch_count() - n0) // n} launches/step")

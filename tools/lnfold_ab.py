#!/usr/bin/env python
"""A/B of option "ln_fold" inside one process (same box, same thermal state): full step (B=64 x 10 s, large model)
with the per-layer LayerNorms as separate kernels vs folded into the neighbouring GEMMs, alternated."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wav2vec2_oracle as wo  # noqa: E402
from svt_speechbrain_b200._lib import check, lib  # noqa: E402
from svt_speechbrain_b200.engine import EncoderEngine, encoder_config_from_hf  # noqa: E402
from transformers import Wav2Vec2Config  # noqa: E402

dev = torch.device("cuda", 0)
cfg = wo.W2V2Config.large()
eng = EncoderEngine(encoder_config_from_hf(Wav2Vec2Config(**cfg.hf_kwargs()), True, True), dev)
sd = wo.random_weights(cfg, seed=0)
head = wo.random_head(cfg.hidden_size, 20, seed=0)
eng.load(sd, head["w.weight"], head["w.bias"])
B, L = 64, 160000
wavs = [torch.randn(B, L, device=dev) for _ in range(4)]


def timeit(n=6, warm=2):
    for i in range(warm):
        eng.forward(wavs[i % 4], want_feats=False, want_logits=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        eng.forward(wavs[i % 4], want_feats=False, want_logits=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rnd in range(3):
    for fold in (0, 1):
        check(lib().svt_set_option(b"ln_fold", fold))
        print(f"round {rnd} ln_fold={fold}: {timeit():.3f} ms/step", flush=True)
check(lib().svt_set_option(b"ln_fold", 1))

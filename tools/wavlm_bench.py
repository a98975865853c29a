#!/usr/bin/env python
"""WavLM-large step (64 x 10-s clips): attention with the gated relative position bias in the tcgen05 kernel vs the
mma.sync kernel (attention_impl 2 / 1), and wav2vec2-large for scale."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wav2vec2_oracle as wo  # noqa: E402  (seeded weights only)
from svt_speechbrain_b200._lib import check, lib  # noqa: E402
from svt_speechbrain_b200.engine import EncoderEngine, encoder_config_from_hf  # noqa: E402
from transformers import WavLMConfig  # noqa: E402

dev = torch.device("cuda", 0)
cfg = wo.W2V2Config.wavlm_large()
eng = EncoderEngine(encoder_config_from_hf(WavLMConfig(**cfg.hf_kwargs()), True, True), dev)
head = wo.random_head(cfg.hidden_size, 20, seed=0)
eng.load(wo.random_weights(cfg, seed=0), head["w.weight"], head["w.bias"])
B, L = 64, 160000
wavs = [torch.randn(B, L, device=dev) for _ in range(4)]


def timeit(n=5, warm=2):
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


for impl, name in ((2, "tcgen05 attention + bias"), (1, "mma.sync attention + bias")):
    check(lib().svt_set_option(b"attention_impl", impl))
    ms = timeit()
    print(f"WavLM-large B=64 x 10 s, {name}: {ms:.2f} ms/step = {B * 10 / ms * 1e3:.0f} audio-s/s", flush=True)
check(lib().svt_set_option(b"attention_impl", 0))

#!/usr/bin/env python
"""How much of the step is launch gaps?  Eager svt_encoder_forward vs the same call captured in a CUDA graph."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wav2vec2_oracle as wo  # noqa: E402
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
static_wav = torch.empty(B, L, device=dev)


def timeit(fn, n=8, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


eager = timeit(lambda i: eng.forward(wavs[i % 4], want_feats=False, want_logits=True))
print(f"eager: {eager:.3f} ms/step")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    static_wav.copy_(wavs[0])
    out = eng.forward(static_wav, want_feats=False, want_logits=True)[1]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        out = eng.forward(static_wav, want_feats=False, want_logits=True)[1]
torch.cuda.synchronize()


def replay(i):
    static_wav.copy_(wavs[i % 4])
    g.replay()


graph = timeit(replay)
print(f"graph: {graph:.3f} ms/step (includes a 41 MB device copy of the input)")

#!/usr/bin/env python
"""BASELINE config 4 per-GPU share: 4 clips of 10 s audio + 500 lip frames through wav2vec2-large + AV-HuBERT-large +
FusionRCA + head (AVTranscriber.logits), with the time of each stage.   python tools/av_bench.py [clips]"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import svt_speechbrain_b200 as svt  # noqa: E402
from oracle import avhubert_oracle as av  # noqa: E402  (seeded weights only)
from oracle import make_golden as mg  # noqa: E402
from oracle import wav2vec2_oracle as wo  # noqa: E402
from test_gpu_e2e import _build  # noqa: E402
from test_gpu_video import _lobe  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
alobe, lin, _, _ = _build(wo.W2V2Config.large())
vcfg = av.AVHubertConfig()
vlobe = _lobe(vcfg, av.random_weights(vcfg, seed=0)).cuda()
fus = svt.FusionRCA()
full = dict(fus.state_dict())
full.update(mg.random_fusion_weights(1024, 3072, seed=3))
fus.load_state_dict(full, strict=True)
fus = fus.cuda()
tr = svt.AVTranscriber(alobe, vlobe, fus, lin)
wav = torch.randn(B, 160000, device="cuda")
video = torch.randn(B, 1, 500, 88, 88, device="cuda")


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


a = alobe(wav)
v = vlobe({"video": video, "audio": None})
t_all = timeit(lambda: tr.logits(wav, video))
tr.concurrent_streams = False
t_seq = timeit(lambda: tr.logits(wav, video))
ref = tr.logits(wav, video)
tr.concurrent_streams = True
assert torch.equal(tr.logits(wav, video), ref)
t_a = timeit(lambda: alobe(wav))
t_v = timeit(lambda: vlobe({"video": video, "audio": None}))
t_f = timeit(lambda: lin(fus(a, v)))
print(f"AV pipeline B={B} x 10 s: {t_all:.2f} ms/step = {B * 10 / t_all * 1e3:.0f} audio-s/s per GPU "
      f"with the two encoders on two streams, {t_seq:.2f} ms back to back "
      f"(audio lobe {t_a:.2f} ms, video lobe {t_v:.2f} ms, fusion + head {t_f:.2f} ms)")


def per_step(fn, n=20, warm=3):
    """Per-step device times (events around every step) and the host time to ISSUE one step."""
    import time
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    issue = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts, issue * 1e3


ts, issue = per_step(lambda: tr.logits(wav, video))
print(f"per step: median {ts[len(ts) // 2]:.2f} ms, min {ts[0]:.2f}, max {ts[-1]:.2f}, mean {sum(ts) / len(ts):.2f}; "
      f"host time per issued step {issue:.2f} ms (a lower bound on nothing: the launch queue throttles the host)", flush=True)

#!/usr/bin/env python
"""BASELINE config 5 on N GPUs: a 5-minute synthetic song cut into overlapping 10-s windows, windows sharded over the
ranks (NCCL gather of the ragged frame logits), stitched and decoded; checked against the same call on one GPU.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/longform_check.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svt_speechbrain_b200 as svt  # noqa: E402
from oracle import wav2vec2_oracle as wo  # noqa: E402  (seeded weights only)
from transformers import Wav2Vec2Config  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cfg = wo.W2V2Config.large()
import tempfile  # noqa: E402
from transformers import Wav2Vec2FeatureExtractor  # noqa: E402

d = os.path.join(tempfile.mkdtemp(), "wav2vec2-longform")
os.makedirs(d)
Wav2Vec2Config(**cfg.hf_kwargs()).save_pretrained(d)
Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True,
                         return_attention_mask=True).save_pretrained(d)
lobe = svt.HuggingFaceWav2Vec2(source=d, save_path=d, pretrain=False, output_norm=True, freeze=True)
lobe.load_state_dict(wo.random_weights(cfg, seed=0), strict=True)
lin = svt.Linear(n_neurons=20, input_size=cfg.hidden_size)
lin.load_state_dict(wo.random_head(cfg.hidden_size, 20, seed=0))
tr = svt.AMTTranscriber(lobe.to(dev), lin.to(dev), device=dev)
wav = torch.randn(16000 * 300 + 4321, generator=torch.Generator().manual_seed(0)) * 0.1
single = tr.long_form_logits(wav, dur=10.0, overlap=1.0).cpu()  # before the process group exists: all windows here
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
for _ in range(2):
    sharded = tr.long_form_logits(wav, dur=10.0, overlap=1.0)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
sharded = tr.long_form_logits(wav, dur=10.0, overlap=1.0)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
notes = tr.decode(sharded)
same = torch.equal(sharded.cpu(), single)
diff = float((sharded.cpu() - single).abs().max())
rel = float((sharded.cpu() - single).norm() / single.norm())
print(f"rank {rank}/{world}: frames {tuple(sharded.shape)}, {len(notes)} notes, max |sharded - single| = {diff:.3e} "
      f"rel-L2 {rel:.3e} (bit-identical: {same}), {wav.numel() / 16000 / dt:.0f} audio-s/s for this song", flush=True)
# per-clip statistics make a window's result independent of the batch / rank it ran in; the double-precision atomics that
# accumulate them can still reorder, so allow the bf16-noise level rather than demanding bit equality
assert sharded.shape == single.shape and rel < 5e-3
if same:
    assert np.array_equal(notes, tr.decode(single.to(dev)))
if world > 1:
    dist.destroy_process_group()

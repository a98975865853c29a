#!/usr/bin/env python
"""Where the time of AMTTranscriber.transcribe_songs goes (host wav -> host notes), stage by stage, at N ranks.
    python tools/notes_profile.py            or under torch.distributed.run with N processes"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import svt_speechbrain_b200 as svt  # noqa: E402
from oracle import wav2vec2_oracle as wo  # noqa: E402  (seeded weights only)

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
cfg = wo.W2V2Config.large()
d = bench._lobe_dir(cfg)
lobe = svt.HuggingFaceWav2Vec2(source=d, save_path=d, pretrain=False, output_norm=True, freeze=True)
lobe.load_state_dict(wo.random_weights(cfg, seed=0), strict=True)
lin = svt.Linear(n_neurons=20, input_size=cfg.hidden_size)
lin.load_state_dict(wo.random_head(cfg.hidden_size, 20, seed=0))
hp = svt.AMTHparams(dur_threshold=10.0)
tr = svt.AMTTranscriber(lobe.to(dev), lin.to(dev), hp, device=dev)
B, L = 64, 160000
host = torch.randn(B, L).pin_memory()
songs = [host[c] for c in range(B)]
print(f"rank {rank}: OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')} torch threads {torch.get_num_threads()} "
      f"view pinned {songs[3].is_pinned()}", flush=True)
for _ in range(2):
    tr.transcribe_songs(songs, dur=10.0, batch_clips=B)


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(3):
    t0 = T()
    dsongs = [w.to(dev, torch.float32, non_blocking=True).reshape(-1) for w in songs]
    t1 = T()
    lgs = tr._clip_logits(dsongs, B, True)
    t2 = T()
    from svt_speechbrain_b200.amt import _pack_frames, _unpack_frames
    packed = _pack_frames(torch.cat(lgs, dim=0), hp)
    t3 = T()
    h = torch.empty(packed.shape, dtype=torch.float32, pin_memory=True)
    h.copy_(packed, non_blocking=True)
    t4 = T()
    p_on, p_off, octv, pc = _unpack_frames(h)
    t5 = T()
    n = 0
    for c in range(B):
        a = c * 499
        n += len(svt.decode_arrays(p_on[a:a + 499], p_off[a:a + 499], octv[a:a + 499], pc[a:a + 499], 0.4, 0.5, 1 / 49.8))
    t6 = T()
    t7 = T()
    tr.transcribe_songs(songs, dur=10.0, batch_clips=B)
    t8 = T()
    print(f"rank {rank} rep {rep}: H2D {1e3 * (t1 - t0):.2f}  forward(+stack) {1e3 * (t2 - t1):.2f}  pack {1e3 * (t3 - t2):.2f}  "
          f"D2H {1e3 * (t4 - t3):.2f}  sigmoid {1e3 * (t5 - t4):.2f}  decode {1e3 * (t6 - t5):.2f} ({n} notes)  | whole call {1e3 * (t8 - t7):.2f} ms",
          flush=True)
if world > 1:
    dist.destroy_process_group()

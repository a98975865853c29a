#!/usr/bin/env python
"""Phase timeline of CTA 0 of the tcgen05 attention kernel (clock64 stamps, svt_debug_attention_trace)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr  # noqa: E402

dev = torch.device("cuda", 0)
clips, heads, T, Ta, dh = 64, 16, 499, 500, 64
D = heads * dh
qkv = torch.randn(clips * Ta, 3 * D, device=dev)
qkv[:, :D] *= dh ** -0.5
qkv = qkv.bfloat16()
o = torch.zeros(clips * Ta, D, device=dev, dtype=torch.bfloat16)
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
check(lib().svt_set_option(b"attention_impl", 2))


def run():
    check(lib().svt_op_attention(ptr(q), ptr(k), ptr(v), ptr(o), 3 * D, 3 * D, 3 * D, D, T, T, Ta, Ta, clips, heads, dh, current_stream_ptr()))


for _ in range(3):
    run()
torch.cuda.synchronize()
buf = torch.zeros(4 * 256, dtype=torch.int64, device=dev)
lib().svt_debug_attention_trace(ptr(buf))
run()
torch.cuda.synchronize()
lib().svt_debug_attention_trace(None)
b = buf.cpu().view(4, 256)
t0 = int(b[b > 0].min())
names = ["loop", "s_full", "ldtm+free", "max+o_full(+output)", "exps+sts", "fence+arrive"]
NS = 6
for rec in (0, 1):
    ev = [int(x) - t0 for x in b[rec] if x > 0]
    print(f"--- softmax tile {rec} (warp {4 + 4 * rec}): {NS} stamps per key block")
    for i in range(0, min(len(ev), NS * 12), NS):
        e = ev[i:i + NS]
        if len(e) < NS:
            break
        d = [e[0]] + [e[k + 1] - e[k] for k in range(NS - 1)]
        print(f"blk {i // NS:2d} start {e[0]:7d} | " + " ".join(f"{n}={x}" for n, x in zip(names[1:], d[1:])) + f" | total {e[NS - 1] - e[0]}")
ev = [int(x) - t0 for x in b[2] if x > 0]
print("--- S issuer stamps (after s_free wait), per (g, t):", ev[:24])
ev = [int(x) - t0 for x in b[3] if x > 0]
print("--- PV issuer stamps (start, end) per (g, t):", ev[:32])

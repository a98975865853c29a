#!/usr/bin/env python
"""One launch of the fused conv1 + LayerNorm(512) + GELU GEMM (for `ncu --set full` captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr  # noqa: E402

dev = torch.device("cuda", 0)
Mc, Kc = int(sys.argv[1]) if len(sys.argv) > 1 else 256000, 1536
a = torch.randn(2 * Mc + 8, 512, device=dev).bfloat16()
w = (torch.randn(512, Kc, device=dev) / Kc ** 0.5).bfloat16()
bias, gam, bet = torch.zeros(512, device=dev), torch.ones(512, device=dev), torch.zeros(512, device=dev)
o = torch.empty(Mc, 512, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    check(lib().svt_op_gemm_rowln(ptr(a), 1024, 512, ptr(w), ptr(bias), ptr(gam), ptr(bet), 1e-5, 1, ptr(o), Mc, 512, Kc, current_stream_ptr()))
torch.cuda.synchronize()

#!/usr/bin/env python
"""Two launches of the FFN-1 GEMM (M = 32000, N = 4096, K = 1024, bias + GELU, bf16 out) for ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr  # noqa: E402

dev = torch.device("cuda", 0)
M, N, K = 32000, 4096, 1024
act = int(sys.argv[1]) if len(sys.argv) > 1 else 1
a = torch.randn(M, K, device=dev).bfloat16()
w = torch.randn(N, K, device=dev).bfloat16()
bias = torch.zeros(N, device=dev)
o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    check(lib().svt_op_gemm(ptr(a), K, K, ptr(w), ptr(bias), None, None, ptr(o), M, N, K, N, act, current_stream_ptr()))
torch.cuda.synchronize()

#!/usr/bin/env python
"""Three launches of the out-projection GEMM (M = 32000, N = K = 1024, fp32 residual in place + bf16 copy + row statistics)
for ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr  # noqa: E402

dev = torch.device("cuda", 0)
M, N, K = 32000, 1024, 1024
a = torch.randn(M, K, device=dev).bfloat16()
w = (torch.randn(N, K, device=dev) * 0.03).bfloat16()
bias = torch.zeros(N, device=dev)
h = torch.randn(M, N, device=dev)
hb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
st = torch.empty(M, N // 128, 2, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    check(lib().svt_op_gemm_ln(ptr(a), ptr(w), ptr(bias), None, None, 1e-5, ptr(st), ptr(h), ptr(h), ptr(hb), M, N, K, 0, current_stream_ptr()))
torch.cuda.synchronize()

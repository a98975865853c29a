#!/usr/bin/env python
"""conv0 of a 64 x 10 s batch through svt_op_conv0 with both kernels (for ncu captures: -k regex:conv0_)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr  # noqa: E402

dev = torch.device("cuda", 0)
Bc, Lc, ta = 64, 160000, 32000
wav = torch.randn(Bc, Lc, device=dev)
w_kc = (torch.randn(10, 512, device=dev) * 0.4).contiguous()
bias, gam, bet = torch.randn(512, device=dev) * 0.1, torch.ones(512, device=dev), torch.zeros(512, device=dev)
o = torch.empty(Bc, ta, 512, device=dev, dtype=torch.bfloat16)
scratch = torch.zeros(4, dtype=torch.float64, device=dev)
for impl in (1, 0, 1, 0):
    check(lib().svt_set_option(b"conv0_impl", impl))
    check(lib().svt_op_conv0(ptr(wav), Bc, Lc, ptr(w_kc), ptr(bias), ptr(gam), ptr(bet), 1, ptr(o), ta, ptr(scratch), current_stream_ptr()))
    torch.cuda.synchronize()

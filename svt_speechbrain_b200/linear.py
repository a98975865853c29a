"""Drop-in for `speechbrain.nnet.linear.Linear` as used for the 20-way AMT head
(speechbrain/nnet/linear.py:15-76; yaml `model: !new:speechbrain.nnet.linear.Linear`).
Parameters live in `self.w` (an nn.Linear) so state_dict keys are `w.weight` / `w.bias` like the reference."""
from __future__ import annotations

import torch
from torch import nn

from ._lib import check, current_stream_ptr, lib, ptr


class Linear(nn.Module):
    def __init__(self, n_neurons, input_shape=None, input_size=None, bias=True, combine_dims=False):
        super().__init__()
        self.combine_dims = combine_dims
        if input_shape is None and input_size is None:
            raise ValueError("Expected one of input_shape or input_size")  # reference :52-53
        if input_size is None:
            input_size = input_shape[-1]
            if len(input_shape) == 4 and self.combine_dims:
                input_size = input_shape[2] * input_shape[3]
        self.w = nn.Linear(input_size, n_neurons, bias=bias)

    def forward(self, x):
        if x.dim() == 4 and self.combine_dims:
            x = x.reshape(x.shape[0], x.shape[1], x.shape[2] * x.shape[3])
        if not x.is_cuda:
            raise RuntimeError("svt_speechbrain_b200.Linear runs on CUDA (sm_100a) only; no CPU fallback")
        n, D = self.w.weight.shape
        if n > 32 or D % 128 != 0:
            raise NotImplementedError("the B200 head kernel covers n_neurons <= 32 and input sizes that are multiples of 128")
        x2 = x.to(torch.float32).contiguous().view(-1, D)
        y = torch.empty(x2.shape[0], n, dtype=torch.float32, device=x.device)
        w = self.w.weight.detach().to(x.device, torch.float32).contiguous()
        b = None if self.w.bias is None else self.w.bias.detach().to(x.device, torch.float32).contiguous()
        with torch.cuda.device(x.device):
            check(lib().svt_op_linear_small(ptr(x2), x2.shape[0], D, ptr(w), ptr(b), n, ptr(y), current_stream_ptr()))
        return y.view(*x.shape[:-1], n)

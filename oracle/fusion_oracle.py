"""fp32 CPU restatement of FusionRCA (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows N20EMv2/audio_visual/fusion.py:192-210 (FusionRCA.forward), :54-79 (RCANet.forward),
:137-183 (RCALayer.forward, normalize_before=False), with the building blocks
speechbrain/nnet/attention.py:686-695,750-778 (nn.MultiheadAttention, packed in_proj, q scaled
by d_h^-0.5 before QK^T), :823-838 (PositionalwiseFeedForward: Linear, ReLU, Dropout, Linear),
speechbrain/nnet/normalization.py:208-222 (LayerNorm, eps=1e-6 as passed at fusion.py:131-132)
and speechbrain/lobes/models/transformer/Transformer.py:200-222 (sinusoidal PE table).
State-dict keys are the reference's (`fusion.layer{1,2}.self_att.att.in_proj_weight`, ...).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def positional_encoding(T: int, D: int, max_len: int = 2500) -> torch.Tensor:
    """Transformer.py:200-211: pe[pos,2i]=sin(pos*exp(-2i ln(1e4)/D)), pe[pos,2i+1]=cos(...)."""
    pe = torch.zeros(max_len, D)
    positions = torch.arange(0, max_len).unsqueeze(1).float()
    denominator = torch.exp(torch.arange(0, D, 2).float() * -(math.log(10000.0) / D))
    pe[:, 0::2] = torch.sin(positions * denominator)
    pe[:, 1::2] = torch.cos(positions * denominator)
    return pe[:T]


def _mha(sd, p, query, kv, nhead):
    """torch nn.MultiheadAttention forward (no masks, eval): packed in_proj rows [0:D]=Wq,[D:2D]=Wk,[2D:3D]=Wv."""
    B, Tq, D = query.shape
    Tk = kv.shape[1]
    dh = D // nhead
    W = sd[p + "in_proj_weight"]
    b = sd[p + "in_proj_bias"]
    q = F.linear(query, W[:D], b[:D]).view(B, Tq, nhead, dh).transpose(1, 2)
    k = F.linear(kv, W[D : 2 * D], b[D : 2 * D]).view(B, Tk, nhead, dh).transpose(1, 2)
    v = F.linear(kv, W[2 * D :], b[2 * D :]).view(B, Tk, nhead, dh).transpose(1, 2)
    a = torch.softmax(torch.matmul(q * (dh ** -0.5), k.transpose(2, 3)), dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, Tq, D)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def rca_layer(sd, p, src_kv, src_q, nhead=8, alpha=0.5):
    """fusion.py:137-183 with normalize_before=False; self- and cross-attention SHARE weights (:148-164)."""
    s = _mha(sd, p + "self_att.att.", src_kv, src_kv, nhead)
    x = _mha(sd, p + "self_att.att.", src_q, src_kv, nhead)
    D = src_kv.shape[-1]
    y = src_kv + s * alpha + x * (1 - alpha)
    y = F.layer_norm(y, (D,), sd[p + "norm1.norm.weight"], sd[p + "norm1.norm.bias"], 1e-6)
    f = F.linear(y, sd[p + "pos_ffn.ffn.0.weight"], sd[p + "pos_ffn.ffn.0.bias"])
    f = F.linear(torch.relu(f), sd[p + "pos_ffn.ffn.3.weight"], sd[p + "pos_ffn.ffn.3.bias"])
    return F.layer_norm(y + f, (D,), sd[p + "norm2.norm.weight"], sd[p + "norm2.norm.bias"], 1e-6)


def fusion_forward(sd: Dict[str, torch.Tensor], audio, video, nhead=8, alpha=0.5):
    """fusion.py:192-210.  audio (B,T1,D), video (B,T2,D) -> (B,T1,D)."""
    B, Ta, D = audio.shape
    Tv = video.shape[1]
    diff = Ta - Tv
    if diff < 0:
        video = video[:, :diff]
    elif diff > 0:
        video = torch.cat([video, torch.zeros(B, diff, D)], dim=1)
    pe = positional_encoding(Ta, D)
    a = audio + pe
    v = video + pe
    o1 = rca_layer(sd, "fusion.layer1.", a, v, nhead, alpha)
    o2 = rca_layer(sd, "fusion.layer2.", v, a, nhead, alpha)
    return o1 + o2

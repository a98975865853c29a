"""Drop-in for `fusion.FusionRCA` (N20EMv2/audio_visual/fusion.py:186-210): same constructor, same
forward(audio_feats[B,T1,D], video_feats[B,T2,D]) -> [B,T1,D], same 25 state_dict keys (`fusion.positional_encoding.pe`,
`fusion.layer{1,2}.self_att.att.*`, `pos_ffn.ffn.{0,3}.*`, `norm{1,2}.norm.*`)."""
from __future__ import annotations

import math

import torch
from torch import nn

from .engine import FusionEngine


class _Att(nn.Module):
    def __init__(self, d_model, nhead):
        super().__init__()
        self.att = nn.MultiheadAttention(embed_dim=d_model, num_heads=nhead, dropout=0.0, bias=True)


class _FFN(nn.Module):
    def __init__(self, d_model, d_ffn):
        super().__init__()
        self.ffn = nn.Sequential(nn.Linear(d_model, d_ffn), nn.ReLU(), nn.Dropout(0.0), nn.Linear(d_ffn, d_model))


class _Norm(nn.Module):
    def __init__(self, d_model):
        super().__init__()
        self.norm = nn.LayerNorm(d_model, eps=1e-6)


class _RCALayer(nn.Module):
    def __init__(self, d_model, nhead, d_ffn):
        super().__init__()
        self.self_att = _Att(d_model, nhead)
        self.pos_ffn = _FFN(d_model, d_ffn)
        self.norm1 = _Norm(d_model)
        self.norm2 = _Norm(d_model)


class _PE(nn.Module):
    def __init__(self, d_model, max_len=2500):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        positions = torch.arange(0, max_len).unsqueeze(1).float()
        denominator = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(positions * denominator)
        pe[:, 1::2] = torch.cos(positions * denominator)
        self.register_buffer("pe", pe.unsqueeze(0))


class _RCANet(nn.Module):
    def __init__(self, d_model, nhead, d_ffn):
        super().__init__()
        self.positional_encoding = _PE(d_model)
        self.layer1 = _RCALayer(d_model, nhead, d_ffn)
        self.layer2 = _RCALayer(d_model, nhead, d_ffn)


class FusionRCA(nn.Module):
    def __init__(self, alpha=0.5, nhead=8, d_ffn=3072, d_model=1024):
        super().__init__()
        self.alpha, self.nhead, self.d_ffn, self.d_model = alpha, nhead, d_ffn, d_model
        self.fusion = _RCANet(d_model, nhead, d_ffn)
        self._engine = None
        self._key = None

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, device) -> FusionEngine:
        key = (str(device), self._weights_key())
        if self._engine is None or self._key != key:
            eng = FusionEngine(self.d_model, self.nhead, self.d_ffn, self.alpha, device)
            eng.load(self.state_dict())
            self._engine, self._key = eng, key
        return self._engine

    def forward(self, audio_feats, video_feats):
        if not audio_feats.is_cuda:
            raise RuntimeError("svt_speechbrain_b200.FusionRCA runs on CUDA (sm_100a) only; no CPU fallback")
        diff = audio_feats.shape[1] - video_feats.shape[1]
        if abs(diff) > 15:
            print("Alignment is wrong")  # reference fusion.py:204-205
        with torch.no_grad():
            return self.engine(audio_feats.device).forward(audio_feats, video_feats.to(audio_feats.device))

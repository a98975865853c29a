"""fp32 CPU restatement of the wav2vec2 AMT forward (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows, line by line:
  * MIR_ST500/huggingface_interface.py:263-298   (lobe: input LN, model(wav)[0], output LN)
  * transformers 5.5.0 models/wav2vec2/modeling_wav2vec2.py ("HF:")
      HF:254-323  conv layers (no-norm / layer-norm / group-norm variants)
      HF:326-379  positional conv embedding (+ weight-norm, same-pad)
      HF:382-434  feature encoder / feature projection
      HF:438-549  attention (eager softmax path)
      HF:552-573  feed forward
      HF:576-655  encoder layers (post-LN base / stable-LN large)
      HF:658-803  encoders
  * speechbrain/nnet/linear.py:61,74             (head, keys w.weight / w.bias)
  * MIR_ST500/train_audio_ssl.py:41-46,93-100    (head slicing, per-frame sigmoid/argmax)

The functions take a plain dict of fp32 tensors keyed by the reference's state_dict
names (prefix "model." as saved by the lobe) so the same weights feed the oracle,
the reference (when importable) and the CUDA path.

Pinned against outputs of the imported reference: oracle/make_golden.py (run in the authoring container,
where /root/reference exists) wrote the fixtures under tests/golden/, and tests/test_oracle_golden.py checks
this module against them on the CPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class W2V2Config:
    """Subset of HF Wav2Vec2Config that the forward depends on."""

    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    conv_dim: Tuple[int, ...] = (512,) * 7
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    conv_bias: bool = False
    feat_extract_norm: str = "group"  # "group" (base) | "layer" (large)
    do_stable_layer_norm: bool = False
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16
    layer_norm_eps: float = 1e-5
    # HuBERT (HF modeling_hubert.py, HubertFeatureProjection): base checkpoints project the conv features without the
    # LayerNorm that wav2vec2 and HuBERT-large apply first.  family picks the HF classes (reference :108-119).
    feat_proj_layer_norm: bool = True
    family: str = "wav2vec2"
    # data2vec-audio (HF modeling_data2vec_audio.py): the positional embedding is a stack of num_conv_pos_embeddings
    # (5) layers of [grouped conv k = conv_pos_kernel_size (19) -> LayerNorm without affine -> GELU], no weight norm.
    conv_pos_kernel_size: int = 19
    # WavLM (HF modeling_wavlm.py, WavLMAttention): bucketed relative position bias, gated per query row
    num_buckets: int = 320
    max_bucket_distance: int = 800
    # HuBERT variants with conv_pos_batch_norm (HF modeling_hubert.py, HubertPositionalConvEmbedding): BatchNorm1d before a
    # plain positional conv instead of weight norm
    conv_pos_batch_norm: bool = False

    @staticmethod
    def wavlm_base() -> "W2V2Config":
        return W2V2Config(family="wavlm")

    @staticmethod
    def wavlm_large() -> "W2V2Config":
        c = W2V2Config.large()
        c.family = "wavlm"
        return c

    @staticmethod
    def data2vec_base() -> "W2V2Config":
        return W2V2Config(family="data2vec", feat_extract_norm="layer", num_conv_pos_embeddings=5)

    @staticmethod
    def hubert_base() -> "W2V2Config":
        return W2V2Config(family="hubert", feat_proj_layer_norm=False)

    @staticmethod
    def hubert_large() -> "W2V2Config":
        c = W2V2Config.large()
        c.family = "hubert"
        return c

    @staticmethod
    def large() -> "W2V2Config":
        return W2V2Config(
            hidden_size=1024,
            num_hidden_layers=24,
            num_attention_heads=16,
            intermediate_size=4096,
            conv_bias=True,
            feat_extract_norm="layer",
            do_stable_layer_norm=True,
        )

    @staticmethod
    def base() -> "W2V2Config":
        return W2V2Config()

    @staticmethod
    def from_hf(cfg) -> "W2V2Config":
        return W2V2Config(
            hidden_size=cfg.hidden_size,
            num_hidden_layers=cfg.num_hidden_layers,
            num_attention_heads=cfg.num_attention_heads,
            intermediate_size=cfg.intermediate_size,
            conv_dim=tuple(cfg.conv_dim),
            conv_kernel=tuple(cfg.conv_kernel),
            conv_stride=tuple(cfg.conv_stride),
            conv_bias=bool(cfg.conv_bias),
            feat_extract_norm=cfg.feat_extract_norm,
            do_stable_layer_norm=bool(cfg.do_stable_layer_norm),
            num_conv_pos_embeddings=cfg.num_conv_pos_embeddings,
            num_conv_pos_embedding_groups=cfg.num_conv_pos_embedding_groups,
            layer_norm_eps=cfg.layer_norm_eps,
            feat_proj_layer_norm=bool(getattr(cfg, "feat_proj_layer_norm", True)),
            family={"Hub": "hubert", "Dat": "data2vec", "Wav": "wav2vec2" if type(cfg).__name__.startswith("Wav2") else "wavlm"}.get(
                type(cfg).__name__[:3], "wav2vec2"),
            num_buckets=int(getattr(cfg, "num_buckets", 320)),
            max_bucket_distance=int(getattr(cfg, "max_bucket_distance", 800)),
            conv_pos_batch_norm=bool(getattr(cfg, "conv_pos_batch_norm", False)),
            conv_pos_kernel_size=int(getattr(cfg, "conv_pos_kernel_size", 19)),
        )

    def hf_kwargs(self) -> dict:
        extra = ({"feat_proj_layer_norm": self.feat_proj_layer_norm, "conv_pos_batch_norm": self.conv_pos_batch_norm}
                 if self.family == "hubert" else {})
        if self.family == "data2vec":
            extra = {"conv_pos_kernel_size": self.conv_pos_kernel_size}
        if self.family == "wavlm":
            extra = {"num_buckets": self.num_buckets, "max_bucket_distance": self.max_bucket_distance}
        return dict(
            **extra,
            hidden_size=self.hidden_size,
            num_hidden_layers=self.num_hidden_layers,
            num_attention_heads=self.num_attention_heads,
            intermediate_size=self.intermediate_size,
            conv_dim=list(self.conv_dim),
            conv_kernel=list(self.conv_kernel),
            conv_stride=list(self.conv_stride),
            conv_bias=self.conv_bias,
            feat_extract_norm=self.feat_extract_norm,
            do_stable_layer_norm=self.do_stable_layer_norm,
            num_conv_pos_embeddings=self.num_conv_pos_embeddings,
            num_conv_pos_embedding_groups=self.num_conv_pos_embedding_groups,
            layer_norm_eps=self.layer_norm_eps,
        )

    def num_frames(self, n_samples: int) -> int:
        t = n_samples
        for k, s in zip(self.conv_kernel, self.conv_stride):
            t = (t - k) // s + 1
        return t


def whole_tensor_layer_norm(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """`F.layer_norm(x, x.shape)` (huggingface_interface.py:289,296): ONE mean / biased
    variance over every element of the call's tensor, no affine."""
    mean = x.mean()
    var = ((x - mean) ** 2).mean()
    return (x - mean) / torch.sqrt(var + eps)


def gelu(x):  # HF ACT2FN["gelu"] == exact erf GELU (HF activations.py)
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def pos_conv_weight(sd: Dict[str, torch.Tensor], prefix: str = "model.") -> torch.Tensor:
    """weight_norm(dim=2) recomposition, HF:343-355: w = g * v / ||v||, norm over dims (0,1)
    separately for every kernel tap.  Accepts both the parametrizations.* names (torch>=2.1)
    and the legacy weight_g / weight_v names."""
    base = prefix + "encoder.pos_conv_embed.conv."
    if base + "parametrizations.weight.original0" in sd:
        g = sd[base + "parametrizations.weight.original0"]
        v = sd[base + "parametrizations.weight.original1"]
    else:
        g = sd[base + "weight_g"]
        v = sd[base + "weight_v"]
    norm = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    return g * v / norm


def feature_encoder(cfg: W2V2Config, sd, x: torch.Tensor, prefix="model.", taps: Optional[dict] = None):
    """HF:382-419.  x: (B, L) -> (B, C, T)."""
    h = x[:, None, :]
    for i, (k, s) in enumerate(zip(cfg.conv_kernel, cfg.conv_stride)):
        p = f"{prefix}feature_extractor.conv_layers.{i}."
        bias = sd.get(p + "conv.bias") if cfg.conv_bias else None
        h = F.conv1d(h, sd[p + "conv.weight"], bias, stride=s)
        if cfg.feat_extract_norm == "layer":  # HF:275-299
            h = h.transpose(1, 2)
            h = F.layer_norm(h, (h.shape[-1],), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)
            h = h.transpose(1, 2)
        elif i == 0:  # HF:302-323 GroupNorm(num_groups=C): per (b, c) stats over time
            mean = h.mean(dim=2, keepdim=True)
            var = ((h - mean) ** 2).mean(dim=2, keepdim=True)
            h = (h - mean) / torch.sqrt(var + 1e-5)
            h = h * sd[p + "layer_norm.weight"][None, :, None] + sd[p + "layer_norm.bias"][None, :, None]
        h = gelu(h)
        if taps is not None:
            taps[f"conv{i}"] = h
    return h


def wavlm_relative_buckets(cfg: W2V2Config, T: int) -> torch.Tensor:
    """WavLMAttention._relative_positions_bucket on memory - context positions: (T, T) long, entry [i, j] for key j - query i."""
    rel = torch.arange(T)[None, :] - torch.arange(T)[:, None]
    nb = cfg.num_buckets // 2
    buckets = (rel > 0).long() * nb
    rel = rel.abs()
    max_exact = nb // 2
    large = torch.log(rel.float() / max_exact) / math.log(cfg.max_bucket_distance / max_exact) * (nb - max_exact)
    large = torch.min((max_exact + large).long(), torch.full_like(rel, nb - 1))
    return buckets + torch.where(rel < max_exact, rel, large)


def wavlm_position_bias(cfg: W2V2Config, sd, prefix: str, T: int) -> torch.Tensor:
    """compute_bias of layer 0 (the only layer with rel_attn_embed), reused by every layer: (H, T, T)."""
    emb = sd[prefix + "encoder.layers.0.attention.rel_attn_embed.weight"]  # (num_buckets, H)
    return emb[wavlm_relative_buckets(cfg, T)].permute(2, 0, 1)


def wavlm_gate(cfg: W2V2Config, sd, x, p):
    """gru_rel_pos gate of WavLMAttention.forward: (B, H, T) from the attention input rows x (B, T, D)."""
    B, T, D = x.shape
    H = cfg.num_attention_heads
    xh = x.view(B, T, H, D // H).permute(0, 2, 1, 3)
    proj = F.linear(xh, sd[p + "gru_rel_pos_linear.weight"], sd[p + "gru_rel_pos_linear.bias"])
    proj = proj.view(B, H, T, 2, 4).sum(-1)
    ga, gb = torch.sigmoid(proj).chunk(2, dim=-1)
    return (ga * (gb * sd[p + "gru_rel_pos_const"] - 1.0) + 2.0).squeeze(-1)


def attention(cfg: W2V2Config, sd, x, p, pos_bias: Optional[torch.Tensor] = None):
    """HF:466-549 with no attention mask (the lobe passes none, huggingface_interface.py:292).  WavLM: the scores get
    gate[b, h, i] * pos_bias[h, i, j] added before the softmax (additive attn_mask of F.multi_head_attention_forward)."""
    B, T, D = x.shape
    H = cfg.num_attention_heads
    dh = D // H
    q = F.linear(x, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    k = F.linear(x, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(x, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    q = q.view(B, T, H, dh).transpose(1, 2)
    k = k.view(B, T, H, dh).transpose(1, 2)
    v = v.view(B, T, H, dh).transpose(1, 2)
    w = torch.matmul(q, k.transpose(2, 3)) * (dh ** -0.5)
    if pos_bias is not None:
        w = w + wavlm_gate(cfg, sd, x, p)[:, :, :, None] * pos_bias[None]
    w = torch.softmax(w, dim=-1)
    o = torch.matmul(w, v).transpose(1, 2).reshape(B, T, D)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def feed_forward(sd, x, p):
    """HF:552-573."""
    h = F.linear(x, sd[p + "intermediate_dense.weight"], sd[p + "intermediate_dense.bias"])
    h = gelu(h)
    return F.linear(h, sd[p + "output_dense.weight"], sd[p + "output_dense.bias"])


def _ln(sd, x, p, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], eps)


def encoder(cfg: W2V2Config, sd, h, prefix="model.", taps: Optional[dict] = None):
    """HF:658-727 (base, post-LN) / HF:730-803 (large, stable-LN)."""
    eps = cfg.layer_norm_eps
    e = prefix + "encoder."
    if cfg.family == "data2vec":
        # Data2VecAudioPositionalConvEmbedding: 5 x [conv -> SamePad -> LN(no affine) -> GELU], then h + pos
        kpos = cfg.conv_pos_kernel_size
        pos = h.transpose(1, 2)
        for i in range(cfg.num_conv_pos_embeddings):
            lp = f"{e}pos_conv_embed.layers.{i}.conv."
            pos = F.conv1d(pos, sd[lp + "weight"], sd[lp + "bias"], padding=kpos // 2,
                           groups=cfg.num_conv_pos_embedding_groups)
            if kpos % 2 == 0:
                pos = pos[:, :, :-1]
            pos = F.layer_norm(pos.transpose(1, 2), (pos.shape[1],), None, None, 1e-5).transpose(1, 2)
            pos = gelu(pos)
        pos = pos.transpose(1, 2)
    else:
        x = h.transpose(1, 2)
        if cfg.conv_pos_batch_norm:  # eval-mode BatchNorm1d, then a plain conv
            bn = e + "pos_conv_embed.batch_norm."
            x = F.batch_norm(x, sd[bn + "running_mean"], sd[bn + "running_var"], sd[bn + "weight"], sd[bn + "bias"], False, 0.0, 1e-5)
            w = sd[e + "pos_conv_embed.conv.weight"]
        else:
            w = pos_conv_weight(sd, prefix)
        kpos = cfg.num_conv_pos_embeddings
        pos = F.conv1d(x, w, sd[e + "pos_conv_embed.conv.bias"], padding=kpos // 2,
                       groups=cfg.num_conv_pos_embedding_groups)
        if kpos % 2 == 0:
            pos = pos[:, :, :-1]  # HF:371-379 SamePad
        pos = gelu(pos).transpose(1, 2)
    h = h + pos
    if taps is not None:
        taps["pos"] = h
    if not cfg.do_stable_layer_norm:
        h = _ln(sd, h, e + "layer_norm.", eps)  # HF:692
    pb = wavlm_position_bias(cfg, sd, prefix, h.shape[1]) if cfg.family == "wavlm" else None
    for l in range(cfg.num_hidden_layers):
        p = f"{e}layers.{l}."
        if cfg.do_stable_layer_norm:  # HF:612-655
            h = h + attention(cfg, sd, _ln(sd, h, p + "layer_norm.", eps), p + "attention.", pb)
            h = h + feed_forward(sd, _ln(sd, h, p + "final_layer_norm.", eps), p + "feed_forward.")
        else:  # HF:576-609
            h = _ln(sd, h + attention(cfg, sd, h, p + "attention.", pb), p + "layer_norm.", eps)
            h = _ln(sd, h + feed_forward(sd, h, p + "feed_forward."), p + "final_layer_norm.", eps)
        if taps is not None:
            taps[f"layer{l}"] = h
    if cfg.do_stable_layer_norm:
        h = _ln(sd, h, e + "layer_norm.", eps)  # HF:792
    return h


def whole_tensor_stats(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(mean, biased variance) over every element, as `F.layer_norm(x, x.shape)` takes them."""
    mean = x.mean()
    return mean, ((x - mean) ** 2).mean()


def lobe_forward(cfg: W2V2Config, sd, wav: torch.Tensor, normalize_wav=True, output_norm=True,
                 prefix="model.", taps: Optional[dict] = None, in_stats=None, out_stats=None) -> torch.Tensor:
    """HuggingFaceWav2Vec2.extract_features (huggingface_interface.py:279-298). wav (B,L) -> (B,T,D).
    in_stats / out_stats = (mean, var): use these statistics in the two whole-tensor norms instead of the ones of `wav`
    itself -- that is how rows of a LARGER reference call are evaluated one clip at a time (amt_logits_of_clips)."""
    x = wav.float()
    if normalize_wav:
        if in_stats is None:
            x = whole_tensor_layer_norm(x)
        else:
            x = (x - in_stats[0]) / torch.sqrt(in_stats[1] + 1e-5)
    h = feature_encoder(cfg, sd, x, prefix, taps).transpose(1, 2)  # HF:1349
    if cfg.feat_proj_layer_norm:  # HF wav2vec2:1352 / hubert HubertFeatureProjection.forward
        h = _ln(sd, h, prefix + "feature_projection.layer_norm.", cfg.layer_norm_eps)
    h = F.linear(h, sd[prefix + "feature_projection.projection.weight"], sd[prefix + "feature_projection.projection.bias"])
    if taps is not None:
        taps["proj"] = h
    h = encoder(cfg, sd, h, prefix, taps)
    if taps is not None:
        taps["enc"] = h
    if output_norm:
        if out_stats is None:
            h = whole_tensor_layer_norm(h)
        else:
            h = (h - out_stats[0]) / torch.sqrt(out_stats[1] + 1e-5)
    return h


def amt_logits_of_clips(cfg, sd, head_sd, wav: torch.Tensor, clips, stat_clips=None) -> torch.Tensor:
    """Rows `clips` of amt_logits(cfg, sd, head_sd, wav) -- ONE reference call on the whole batch, whose two whole-tensor
    norms couple its clips -- evaluated clip by clip so that a 64 x 10 s batch never has to exist in fp32 on the host:
    the input statistics are taken over all of `wav` (exact), every clip of `stat_clips` (default: all clips = exact) is
    pushed through the encoder with them, the output statistics are pooled over those clips' features in float64, and the
    requested clips are normalised and sent through the head.  With stat_clips a SUBSET the output statistics are an
    estimate; it is exact whenever every row of the final features has the same mean and variance (e.g. gamma = 1, beta = 0
    in encoder.layer_norm, the HF initialisation)."""
    clips = list(clips)
    B = wav.shape[0]
    stat_clips = list(range(B)) if stat_clips is None else list(stat_clips)
    in_stats = whole_tensor_stats(wav.float())
    feats = {}
    for c in sorted(set(clips) | set(stat_clips)):
        feats[c] = lobe_forward(cfg, sd, wav[c:c + 1], in_stats=in_stats, output_norm=False)
    n = sum(feats[c].numel() for c in stat_clips)
    mean = sum(feats[c].double().sum() for c in stat_clips) / n
    var = sum(((feats[c].double() - mean) ** 2).sum() for c in stat_clips) / n
    out_stats = (mean.float(), var.float())
    return torch.cat([head_forward(head_sd, (feats[c] - out_stats[0]) / torch.sqrt(out_stats[1] + 1e-5)) for c in clips])


def head_forward(head_sd, feats: torch.Tensor) -> torch.Tensor:
    """speechbrain.nnet.linear.Linear (linear.py:61,74): keys w.weight (n,D), w.bias (n)."""
    return F.linear(feats, head_sd["w.weight"], head_sd.get("w.bias"))


def amt_logits(cfg, sd, head_sd, wav, **kw) -> torch.Tensor:
    """wav (B,L) -> logits (B,T,20): AMT.compute_forward, train_audio_ssl.py:28-48."""
    return head_forward(head_sd, lobe_forward(cfg, sd, wav, **kw))


def frame_info_from_logits(logits: torch.Tensor, pitch_octave_num=4, pitch_class_num=12):
    """train_audio_ssl.py:41-46,93-100 for ONE utterance: logits (T,20) ->
    (p_on fp32[T], p_off fp32[T], octave int64[T], pitch_class int64[T]).
    Columns: 0 onset, 1 offset, 2..2+O octave(+none), rest pitch class(+none)."""
    lo = logits.float().cpu()
    p_on = torch.sigmoid(lo[:, 0])
    p_off = torch.sigmoid(lo[:, 1])
    n_oct = pitch_octave_num + 1
    octv = lo[:, 2 : 2 + n_oct].argmax(dim=1)
    pc = lo[:, 2 + n_oct : 2 + n_oct + pitch_class_num + 1].argmax(dim=1)
    return p_on, p_off, octv, pc


def random_head(D: int, n: int = 20, seed: int = 0) -> Dict[str, torch.Tensor]:
    """torch.nn.Linear default init under a fixed seed (speechbrain Linear wraps nn.Linear)."""
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / math.sqrt(D)
    w = (torch.rand(n, D, generator=g) * 2 - 1) * bound
    b = (torch.rand(n, generator=g) * 2 - 1) * bound
    return {"w.weight": w, "w.bias": b}


def random_weights(cfg: W2V2Config, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init weights of the named architecture via HF's own `_init_weights`
    (transformers is a dependency of the reference and exists on the GPU box too).
    Returns the lobe-style state dict (prefix "model.")."""
    from transformers import (Data2VecAudioConfig, Data2VecAudioModel, HubertConfig, HubertModel, Wav2Vec2Config,
                              Wav2Vec2Model)

    torch.manual_seed(seed)
    if cfg.family == "hubert":
        m = HubertModel(HubertConfig(**cfg.hf_kwargs())).eval()
    elif cfg.family == "data2vec":
        m = Data2VecAudioModel(Data2VecAudioConfig(**cfg.hf_kwargs())).eval()
    elif cfg.family == "wavlm":
        from transformers import WavLMConfig, WavLMModel
        m = WavLMModel(WavLMConfig(**cfg.hf_kwargs())).eval()
    else:
        m = Wav2Vec2Model(Wav2Vec2Config(**cfg.hf_kwargs())).eval()
    return {"model." + k: v.detach().clone().float() for k, v in m.state_dict().items()}

"""fp32 CPU restatement of the AV-HuBERT video stream (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows:
  * N20EMv2/video_only/fairseq_interface.py:454-485  (FairseqAVHubertPretrain.forward / extract_features:
        optional whole-tensor input LN, model.extract_finetune({"video": x, "audio": None}), whole-tensor output LN)
  * N20EMv2/video_only/hubert.py:688-739             (AVHubertModel.extract_finetune: video features, zero audio
        stream, channel concat [audio, video], LayerNorm(2D), post_extract_proj 2D -> D, TransformerEncoder)
  * N20EMv2/video_only/hubert.py:311-326             (SubModel: ResEncoder -> Linear 512 -> D)
  * N20EMv2/video_only/resnet.py:37-171              (BasicBlock / ResNet-18 trunk with PReLU / ResEncoder:
        Conv3d(1->64,(5,7,7),s(1,2,2),p(2,3,3)) + BN3d + PReLU + MaxPool3d((1,3,3),s(1,2,2),p(0,1,1)))
  * fairseq `fairseq/models/wav2vec/wav2vec2.py` TransformerEncoder / TransformerSentenceEncoderLayer with
        layer_norm_first=True -- NOT in the reference tree (un-vendored av_hubert submodule, README.md:45-55, version
        unpinned).  Its published algorithm is restated here: x += GELU(SamePad(weight-normed grouped Conv1d(x)));
        N x [x += SelfAttn(LN(x)); x += fc2(GELU(fc1(LN(x))))]; LN.  That is structurally the HF stable-LN encoder,
        so the body reuses oracle/wav2vec2_oracle.encoder after renaming the fairseq keys.

Pinning: the ResNet front end is checked against the reference's own `resnet.ResEncoder` (importable: it needs only
torch) in tests/golden/avhubert_resnet_*.npz (oracle/make_golden.py).  The transformer body and the glue of hubert.py
cannot be imported here (they need fairseq): **parity unpinned** for that part beyond the HF-equivalence above.

State-dict keys are the reference module's (`model.` prefix as saved by FairseqAVHubertPretrain):
  model.feature_extractor_video.resnet.frontend3D.{0.weight, 1.{weight,bias,running_mean,running_var}, 2.weight}
  model.feature_extractor_video.resnet.trunk.layer{1..4}.{0,1}.{conv1,conv2}.weight / bn{1,2}.* / relu{1,2}.weight /
        downsample.{0.weight, 1.*}
  model.feature_extractor_video.proj.{weight,bias}, model.layer_norm.*, model.post_extract_proj.*,
  model.encoder.pos_conv.0.{bias,weight_g,weight_v}, model.encoder.layers.{l}.self_attn.{q,k,v,out}_proj.*,
  model.encoder.layers.{l}.{self_attn_layer_norm,fc1,fc2,final_layer_norm}.*, model.encoder.layer_norm.*
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import wav2vec2_oracle as wo

BN_EPS = 1e-5


@dataclass
class AVHubertConfig:
    """Fields of hubert.py's AVHubertConfig the video forward depends on (large_vox_iter5 values by default)."""

    encoder_embed_dim: int = 1024
    encoder_layers: int = 24
    encoder_attention_heads: int = 16
    encoder_ffn_embed_dim: int = 4096
    conv_pos: int = 128
    conv_pos_groups: int = 16
    layer_norm_eps: float = 1e-5

    def w2v2(self) -> wo.W2V2Config:
        return wo.W2V2Config(hidden_size=self.encoder_embed_dim, num_hidden_layers=self.encoder_layers,
                             num_attention_heads=self.encoder_attention_heads, intermediate_size=self.encoder_ffn_embed_dim,
                             do_stable_layer_norm=True, num_conv_pos_embeddings=self.conv_pos,
                             num_conv_pos_embedding_groups=self.conv_pos_groups, layer_norm_eps=self.layer_norm_eps)


def _bn(sd, x, p):
    """eval-mode BatchNorm (running statistics)"""
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, BN_EPS)


def basic_block(sd, x, p, stride):
    """resnet.py:37-76"""
    out = F.conv2d(x, sd[p + "conv1.weight"], None, stride=stride, padding=1)
    out = F.prelu(_bn(sd, out, p + "bn1."), sd[p + "relu1.weight"])
    out = _bn(sd, F.conv2d(out, sd[p + "conv2.weight"], None, stride=1, padding=1), p + "bn2.")
    if p + "downsample.0.weight" in sd:
        residual = _bn(sd, F.conv2d(x, sd[p + "downsample.0.weight"], None, stride=stride), p + "downsample.1.")
    else:
        residual = x
    return F.prelu(out + residual, sd[p + "relu2.weight"])


def res_encoder(sd, video: torch.Tensor, prefix: str, taps: Optional[dict] = None) -> torch.Tensor:
    """resnet.py:133-171.  video (B, 1, T, H, W) -> (B, 512, T)."""
    B, _, T, _, _ = video.shape
    f = prefix + "frontend3D."
    x = F.conv3d(video, sd[f + "0.weight"], None, stride=(1, 2, 2), padding=(2, 3, 3))
    x = F.prelu(_bn(sd, x, f + "1."), sd[f + "2.weight"])
    x = F.max_pool3d(x, kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
    if taps is not None:
        taps["frontend"] = x
    x = x.transpose(1, 2).reshape(B * T, x.shape[1], x.shape[3], x.shape[4])  # threeD_to_2D_tensor
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for bi in (0, 1):
            x = basic_block(sd, x, f"{prefix}trunk.layer{li}.{bi}.", stride if bi == 0 else 1)
        if taps is not None:
            taps[f"layer{li}"] = x
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)  # (B*T, 512)
    return x.view(B, T, -1).transpose(1, 2).contiguous()


def fairseq_to_hf_encoder_keys(sd: Dict[str, torch.Tensor], prefix: str = "model.") -> Dict[str, torch.Tensor]:
    """Rename fairseq TransformerEncoder keys to the HF Wav2Vec2Encoder names used by wav2vec2_oracle.encoder."""
    out = {}
    e = prefix + "encoder."
    for k, v in sd.items():
        if not k.startswith(e):
            continue
        n = k[len(e):]
        n = n.replace("pos_conv.0.", "pos_conv_embed.conv.")
        n = n.replace(".self_attn_layer_norm.", ".layer_norm.")
        n = n.replace(".self_attn.", ".attention.")
        n = n.replace(".fc1.", ".feed_forward.intermediate_dense.")
        n = n.replace(".fc2.", ".feed_forward.output_dense.")
        out[e + n] = v
    return out


def extract_finetune_video(cfg: AVHubertConfig, sd, video: torch.Tensor, prefix: str = "model.",
                           taps: Optional[dict] = None) -> torch.Tensor:
    """hubert.py:688-739 with source = {"video": video, "audio": None}.  video (B,1,T,88,88) -> (B, T, D)."""
    D = cfg.encoder_embed_dim
    fv = prefix + "feature_extractor_video."
    feats = res_encoder(sd, video.float(), fv + "resnet.", taps)                            # (B, 512, T)
    feats = F.linear(feats.transpose(1, 2), sd[fv + "proj.weight"], sd[fv + "proj.bias"])   # SubModel.proj -> (B, T, D)
    if taps is not None:
        taps["video_proj"] = feats
    audio = feats.new_zeros(feats.shape)                                                    # hubert.py:700-702
    x = torch.cat([audio, feats], dim=2)                                                    # channel concat, audio first (:708)
    x = F.layer_norm(x, (2 * D,), sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"], cfg.layer_norm_eps)
    x = F.linear(x, sd[prefix + "post_extract_proj.weight"], sd[prefix + "post_extract_proj.bias"])
    if taps is not None:
        taps["post_proj"] = x
    return wo.encoder(cfg.w2v2(), fairseq_to_hf_encoder_keys(sd, prefix), x, prefix, taps)


def lobe_forward(cfg: AVHubertConfig, sd, video: torch.Tensor, input_norm: bool = False, output_norm: bool = True,
                 prefix: str = "model.", taps: Optional[dict] = None) -> torch.Tensor:
    """FairseqAVHubertPretrain.extract_features (fairseq_interface.py:470-485)."""
    x = video.float()
    if input_norm:
        x = wo.whole_tensor_layer_norm(x)
    out = extract_finetune_video(cfg, sd, x, prefix, taps)
    if output_norm:
        out = wo.whole_tensor_layer_norm(out)
    return out


def random_weights(cfg: AVHubertConfig, seed: int = 0, prefix: str = "model.", hf_sd=None) -> Dict[str, torch.Tensor]:
    """Seeded random weights with the reference's shapes and init scales (resnet.py:93-100 conv init; BN with
    non-trivial running statistics and affines so that the folding is exercised; PReLU slopes around 0.25;
    transformer weights via HF's `_init_weights` renamed to fairseq keys)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    D = cfg.encoder_embed_dim

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def bn(p, c):
        sd[p + "weight"] = 1.0 + 0.1 * rn(c)
        sd[p + "bias"] = 0.1 * rn(c)
        sd[p + "running_mean"] = 0.1 * rn(c)
        sd[p + "running_var"] = 1.0 + 0.2 * torch.rand(c, generator=g)

    r = prefix + "feature_extractor_video.resnet."
    sd[r + "frontend3D.0.weight"] = rn(64, 1, 5, 7, 7, std=math.sqrt(2.0 / (5 * 7 * 7)))
    bn(r + "frontend3D.1.", 64)
    sd[r + "frontend3D.2.weight"] = 0.25 + 0.05 * rn(64)
    inpl = 64
    for li, planes in ((1, 64), (2, 128), (3, 256), (4, 512)):
        for bi in (0, 1):
            p = f"{r}trunk.layer{li}.{bi}."
            cin = inpl if bi == 0 else planes
            sd[p + "conv1.weight"] = rn(planes, cin, 3, 3, std=math.sqrt(2.0 / (9 * planes)))
            bn(p + "bn1.", planes)
            sd[p + "relu1.weight"] = 0.25 + 0.05 * rn(planes)
            sd[p + "conv2.weight"] = rn(planes, planes, 3, 3, std=math.sqrt(2.0 / (9 * planes)))
            bn(p + "bn2.", planes)
            sd[p + "relu2.weight"] = 0.25 + 0.05 * rn(planes)
            if bi == 0 and (li > 1):
                sd[p + "downsample.0.weight"] = rn(planes, cin, 1, 1, std=math.sqrt(2.0 / planes))
                bn(p + "downsample.1.", planes)
        inpl = planes
    fv = prefix + "feature_extractor_video."
    sd[fv + "proj.weight"] = rn(D, 512, std=1.0 / math.sqrt(512))
    sd[fv + "proj.bias"] = 0.05 * rn(D)
    sd[prefix + "layer_norm.weight"] = 1.0 + 0.1 * rn(2 * D)
    sd[prefix + "layer_norm.bias"] = 0.05 * rn(2 * D)
    sd[prefix + "post_extract_proj.weight"] = rn(D, 2 * D, std=1.0 / math.sqrt(2 * D))
    sd[prefix + "post_extract_proj.bias"] = 0.05 * rn(D)
    # transformer body: HF init of the equivalent stable-LN encoder, keys renamed to fairseq
    # (hf_sd: an already built wo.random_weights(cfg.w2v2(), seed) state dict, to skip building the HF model twice)
    hf = wo.random_weights(cfg.w2v2(), seed=seed) if hf_sd is None else hf_sd
    for k, v in hf.items():
        if not k.startswith("model.encoder."):
            continue
        n = k[len("model.encoder."):]
        n = n.replace("pos_conv_embed.conv.parametrizations.weight.original0", "pos_conv.0.weight_g")
        n = n.replace("pos_conv_embed.conv.parametrizations.weight.original1", "pos_conv.0.weight_v")
        n = n.replace("pos_conv_embed.conv.", "pos_conv.0.")
        n = n.replace(".feed_forward.intermediate_dense.", ".fc1.").replace(".feed_forward.output_dense.", ".fc2.")
        n = n.replace(".attention.", ".self_attn.")
        if ".final_layer_norm." not in n and n.startswith("layers.") and ".layer_norm." in n:
            n = n.replace(".layer_norm.", ".self_attn_layer_norm.")
        sd[prefix + "encoder." + n] = v
    return sd

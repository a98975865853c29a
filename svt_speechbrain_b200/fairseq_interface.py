"""Drop-in for the reference lobe `fairseq_interface.FairseqAVHubertPretrain`
(N20EMv2/video_only/fairseq_interface.py:350-499).

Same constructor kwargs (`pretrained_path, save_path, input_norm=None, output_norm=True, freeze=True, pretrain=True,
dropout=None`), same `forward({"video": FloatTensor[B, 1, T, 88, 88], "audio": None}) -> FloatTensor[B, T, D]` and the
same `state_dict()` keys (`model.feature_extractor_video.resnet.*`, `model.feature_extractor_{audio,video}.proj.*`,
`model.layer_norm.*`, `model.post_extract_proj.*`, `model.mask_emb`, `model.encoder.*` with fairseq names), so checkpoints
written by the reference (`Checkpointer`, `encoder.pt`) load unchanged.  In a recipe YAML
(N20EMv2/video_only/hparams/train_video_ssl.yaml:90-94) only the class path changes:

    encoder: !new:svt_speechbrain_b200.fairseq_interface.FairseqAVHubertPretrain

The reference builds the model through fairseq (`checkpoint_utils.load_model_ensemble_and_task`, :414-420).  fairseq is
not a dependency here: the checkpoint file is read with `torch.load` and only its `model` state dict and the few config
fields the video forward needs are used; this module is the parameter container, all arithmetic runs in libsvt_b200.so
(csrc/video.cu).  Inference only, CUDA (sm_100a) only, video modality only (the AMT recipes pass `audio=None`).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch import nn

from ._lib import VideoConfig
from .engine import VideoEngine

_LARGE = dict(encoder_embed_dim=1024, encoder_layers=24, encoder_attention_heads=16, encoder_ffn_embed_dim=4096,
              conv_pos=128, conv_pos_groups=16, audio_feat_dim=104)


def _block(inpl, planes, downsample):
    m = nn.Module()
    m.conv1 = nn.Conv2d(inpl, planes, 3, stride=2 if downsample else 1, padding=1, bias=False)
    m.bn1 = nn.BatchNorm2d(planes)
    m.relu1 = nn.PReLU(planes)
    m.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
    m.bn2 = nn.BatchNorm2d(planes)
    m.relu2 = nn.PReLU(planes)
    if downsample:
        m.downsample = nn.Sequential(nn.Conv2d(inpl, planes, 1, stride=2, bias=False), nn.BatchNorm2d(planes))
    return m


class _ResEncoder(nn.Module):  # parameter layout of resnet.py:133-171
    def __init__(self):
        super().__init__()
        self.frontend3D = nn.Sequential(nn.Conv3d(1, 64, (5, 7, 7), (1, 2, 2), (2, 3, 3), bias=False), nn.BatchNorm3d(64),
                                        nn.PReLU(64), nn.MaxPool3d((1, 3, 3), (1, 2, 2), (0, 1, 1)))
        self.trunk = nn.Module()
        inpl = 64
        for i, planes in enumerate((64, 128, 256, 512)):
            setattr(self.trunk, f"layer{i + 1}", nn.Sequential(_block(inpl, planes, i > 0), _block(planes, planes, False)))
            inpl = planes


class _SubModel(nn.Module):  # hubert.py:311-326
    def __init__(self, resnet, input_dim, D):
        super().__init__()
        if resnet is not None:
            self.resnet = resnet
        self.proj = nn.Linear(input_dim, D)


class _Layer(nn.Module):  # fairseq TransformerSentenceEncoderLayer parameter names
    def __init__(self, D, F):
        super().__init__()
        self.self_attn = nn.Module()
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            setattr(self.self_attn, n, nn.Linear(D, D))
        self.self_attn_layer_norm = nn.LayerNorm(D)
        self.fc1 = nn.Linear(D, F)
        self.fc2 = nn.Linear(F, D)
        self.final_layer_norm = nn.LayerNorm(D)


class _Encoder(nn.Module):  # fairseq wav2vec2 TransformerEncoder parameter names
    def __init__(self, c):
        super().__init__()
        D = c["encoder_embed_dim"]
        conv = nn.Conv1d(D, D, c["conv_pos"], padding=c["conv_pos"] // 2, groups=c["conv_pos_groups"])
        self.pos_conv = nn.Sequential(torch.nn.utils.weight_norm(conv, name="weight", dim=2))
        self.layers = nn.ModuleList([_Layer(D, c["encoder_ffn_embed_dim"]) for _ in range(c["encoder_layers"])])
        self.layer_norm = nn.LayerNorm(D)


class _AVHubert(nn.Module):  # hubert.py:328-400 (the tensors the checkpoint holds; pre-training heads are not kept)
    def __init__(self, c):
        super().__init__()
        D = c["encoder_embed_dim"]
        self.feature_extractor_audio = _SubModel(None, c["audio_feat_dim"], D)
        self.feature_extractor_video = _SubModel(_ResEncoder(), 512, D)
        self.post_extract_proj = nn.Linear(2 * D, D)
        self.mask_emb = nn.Parameter(torch.rand(D))
        self.encoder = _Encoder(c)
        self.layer_norm = nn.LayerNorm(2 * D)


def _cfg_from_checkpoint(ckpt) -> dict:
    """Pull the handful of model-config fields out of a fairseq checkpoint (`cfg.model` or legacy `args`)."""
    c = dict(_LARGE)
    src = None
    if isinstance(ckpt, dict):
        if ckpt.get("cfg") is not None:
            src = ckpt["cfg"]["model"] if "model" in ckpt["cfg"] else None
            if src is not None and "w2v_args" in src and src["w2v_args"] is not None:  # fine-tuned checkpoints nest it
                src = src["w2v_args"]["model"]
        elif ckpt.get("args") is not None:
            src = vars(ckpt["args"])
    if src is not None:
        for k in c:
            try:
                if src[k] is not None:
                    c[k] = int(src[k])
            except (KeyError, TypeError, AttributeError):
                pass
    return c


class FairseqAVHubertPretrain(nn.Module):
    def __init__(self, pretrained_path, save_path, input_norm=None, output_norm=True, freeze=True, pretrain=True, dropout=None,
                 model_config: Optional[dict] = None):
        super().__init__()
        cfg = dict(_LARGE)
        state = None
        # the reference first brings the checkpoint to `save_path` (speechbrain download_file: an existing destination is
        # kept, a local source is copied, a URL is fetched -- :399) and then loads `save_path` (:414-420)
        ckpt_path = None
        if save_path is not None and os.path.isfile(str(save_path)):
            ckpt_path = str(save_path)
        elif pretrained_path is not None and os.path.isfile(str(pretrained_path)):
            ckpt_path = str(pretrained_path)
            if save_path is not None:
                import shutil
                os.makedirs(os.path.dirname(os.path.abspath(str(save_path))), exist_ok=True)
                shutil.copyfile(ckpt_path, str(save_path))
                ckpt_path = str(save_path)
        if ckpt_path is not None:
            try:
                ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            except ModuleNotFoundError as e:  # fairseq checkpoints pickle an omegaconf config
                raise ImportError(f"reading {ckpt_path} needs the package its config was pickled with ({e.name}); "
                                  "re-save the checkpoint as {'model': state_dict} to load it without fairseq") from e
            cfg = _cfg_from_checkpoint(ckpt)
            state = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
        elif pretrain:
            raise FileNotFoundError(f"neither save_path={save_path!r} nor pretrained_path={pretrained_path!r} is a local file "
                                    "(fetching a URL needs network access: download the AV-HuBERT checkpoint first)")
        if model_config:
            cfg.update(model_config)
        self.model = _AVHubert(cfg)
        self.model_config = cfg
        if pretrain and state is not None:
            own = self.model.state_dict()
            picked = {k: v for k, v in state.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}
            missing = [k for k in own if k not in picked and not k.endswith("num_batches_tracked") and
                       not k.startswith("feature_extractor_audio.") and k != "mask_emb"]
            if missing:
                raise RuntimeError(f"checkpoint {ckpt_path} lacks {len(missing)} tensors, e.g. {missing[:3]}")
            self.model.load_state_dict(picked, strict=False)
        self.freeze = freeze
        self.normalize = bool(input_norm)  # the reference resolves input_norm=None from the task config (:423-429)
        self.output_norm = output_norm
        self.model.eval()
        if self.freeze:
            for p in self.model.parameters():
                p.requires_grad = False
        self._engine = None
        self._engine_key = None

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in self.model.buffers())

    def engine(self, device) -> VideoEngine:
        key = (str(device), self.normalize, self.output_norm, self._weights_key())
        if self._engine is None or self._engine_key != key:
            c = self.model_config
            vc = VideoConfig(c["encoder_embed_dim"], c["encoder_layers"], c["encoder_attention_heads"], c["encoder_ffn_embed_dim"],
                             c["conv_pos"], c["conv_pos_groups"], 1e-5, int(self.normalize), int(bool(self.output_norm)))
            eng = VideoEngine(vc, device)
            eng.load({"model." + k: v for k, v in self.model.state_dict().items()})
            self._engine, self._engine_key = eng, key
        return self._engine

    def forward(self, wav):
        """wav: {"video": (B, 1, T, 88, 88), "audio": None} -> (B, T, D); reference :454-468."""
        if self.freeze:
            with torch.no_grad():
                return self.extract_features(wav).detach()
        return self.extract_features(wav)

    def extract_features(self, wav):
        """Optional whole-tensor input LN, extract_finetune, optional whole-tensor output LN (reference :470-485)."""
        video = wav["video"] if isinstance(wav, dict) else wav
        if isinstance(wav, dict) and wav.get("audio") is not None:
            raise NotImplementedError("the B200 path builds the video-only stream used by the AMT recipes (audio=None)")
        if not video.is_cuda:
            raise RuntimeError("svt_speechbrain_b200.FairseqAVHubertPretrain runs on CUDA (sm_100a) only; no CPU fallback")
        return self.engine(video.device).forward(video)

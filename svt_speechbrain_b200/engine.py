"""Thin Python owners of the C-ABI handles (weights in, device pointers through, nothing computed here)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import EncoderConfig, FusionConfig, VideoConfig, check, current_stream_ptr, lib, ptr, require_cuda


def encoder_config_from_hf(cfg, normalize_wav: bool, output_norm: bool) -> EncoderConfig:
    """Map the HF Wav2Vec2Config / HubertConfig fields the forward depends on onto svt_encoder_config."""
    c = EncoderConfig()
    c.hidden_size = cfg.hidden_size
    c.num_layers = cfg.num_hidden_layers
    c.num_heads = cfg.num_attention_heads
    c.ffn_size = cfg.intermediate_size
    dims = list(cfg.conv_dim)
    if len(set(dims)) != 1:
        raise NotImplementedError("svt_speechbrain_b200: all conv feature-extractor layers must share one width")
    if len(dims) > _lib.MAX_CONV_LAYERS:
        raise NotImplementedError("too many conv layers")
    c.num_conv_layers = len(dims)
    c.conv_dim = dims[0]
    for i, (k, s) in enumerate(zip(cfg.conv_kernel, cfg.conv_stride)):
        c.conv_kernel[i] = k
        c.conv_stride[i] = s
    c.conv_bias = int(bool(cfg.conv_bias))
    # Data2VecAudioConfig defines neither field: its feature extractor is always the layer-norm variant and its encoder
    # post-LN (HF modeling_data2vec_audio.py: Data2VecAudioConvLayer, Data2VecAudioEncoder)
    data2vec = type(cfg).__name__.startswith("Data2Vec")
    feat_norm = getattr(cfg, "feat_extract_norm", "layer" if data2vec else "group")
    if feat_norm not in ("layer", "group"):
        raise ValueError(f"feat_extract_norm={feat_norm!r}")
    c.feat_norm_layer = int(feat_norm == "layer")
    c.stable_layer_norm = int(bool(getattr(cfg, "do_stable_layer_norm", False)))
    if type(cfg).__name__.startswith("Data2Vec"):  # stack of num_conv_pos_embeddings convs of conv_pos_kernel_size taps
        c.pos_conv_kernel = cfg.conv_pos_kernel_size
        c.pos_conv_layers = cfg.num_conv_pos_embeddings
    else:
        c.pos_conv_kernel = cfg.num_conv_pos_embeddings
        c.pos_conv_layers = 0
    c.pos_conv_groups = cfg.num_conv_pos_embedding_groups
    c.layer_norm_eps = float(cfg.layer_norm_eps)
    if getattr(cfg, "feat_extract_activation", "gelu") != "gelu" or getattr(cfg, "hidden_act", "gelu") != "gelu":
        raise NotImplementedError("only the exact-erf GELU activation of wav2vec2/HuBERT is built")
    c.normalize_wav = int(bool(normalize_wav))
    c.output_norm = int(bool(output_norm))
    c.pos_conv_batch_norm = int(bool(getattr(cfg, "conv_pos_batch_norm", False)))  # HubertConfig only
    if type(cfg).__name__.startswith("WavLM"):
        c.rel_pos_buckets = cfg.num_buckets
        c.rel_pos_max_distance = cfg.max_bucket_distance
    c.feat_proj_norm = int(bool(getattr(cfg, "feat_proj_layer_norm", True)))  # HubertConfig only; wav2vec2 always has it
    return c


def _set_tensors(setter, handle, sd: Dict[str, torch.Tensor]):
    for name, t in sd.items():
        if not torch.is_floating_point(t):
            continue
        h = t.detach().to("cpu", torch.float32).contiguous()
        shape = (C.c_int64 * max(h.dim(), 1))(*(list(h.shape) or [1]))
        check(setter(handle, name.encode(), C.c_void_p(h.data_ptr()), shape, h.dim(), 0))


class EncoderEngine:
    """svt_encoder handle + a cached device workspace (grown on demand, owned by torch's allocator)."""

    def __init__(self, cfg: EncoderConfig, device: torch.device):
        self._h = C.c_void_p()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("EncoderEngine needs a CUDA device (no CPU fallback)")
        with torch.cuda.device(self.device):
            check(lib().svt_encoder_create(C.byref(cfg), C.byref(self._h)))
        self._ws: Optional[torch.Tensor] = None
        self.n_out = 0

    def load(self, state_dict: Dict[str, torch.Tensor], head_w: Optional[torch.Tensor] = None,
             head_b: Optional[torch.Tensor] = None):
        with torch.cuda.device(self.device):
            _set_tensors(lib().svt_encoder_set_tensor, self._h, state_dict)
            check(lib().svt_encoder_finalize(self._h))
            if head_w is not None:
                self.set_head(head_w, head_b)

    def set_head(self, w: torch.Tensor, b: Optional[torch.Tensor]):
        w = w.detach().to("cpu", torch.float32).contiguous()
        b = None if b is None else b.detach().to("cpu", torch.float32).contiguous()
        with torch.cuda.device(self.device):
            check(lib().svt_encoder_set_head(self._h, ptr(w), ptr(b), w.shape[0]))
        self.n_out = w.shape[0]

    def set_norm_per_clip(self, per_clip: bool):
        """Whole-tensor norms per clip (= the reference's batch-size-1 evaluation loop) instead of per call."""
        check(lib().svt_encoder_set_norm_per_clip(self._h, int(bool(per_clip))))

    def num_frames(self, n_samples: int) -> int:
        return lib().svt_encoder_num_frames(self._h, n_samples)

    def workspace(self, B: int, L: int) -> torch.Tensor:
        need = lib().svt_encoder_workspace_bytes(self._h, B, L)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward(self, wav: torch.Tensor, want_feats: bool = True, want_logits: bool = False,
                logits_out: Optional[torch.Tensor] = None):
        """logits_out: optional pre-allocated (B, T, n_out) fp32 CUDA tensor to write the logits into (serving loops that
        hand the buffer to an asynchronous collective)."""
        require_cuda(wav, "EncoderEngine.forward")
        if wav.dim() != 2:
            raise ValueError(f"expected wav of shape (batch, samples), got {tuple(wav.shape)}")
        wav = wav.to(torch.float32).contiguous()
        B, L = wav.shape
        T = self.num_frames(L)
        if T <= 0:
            raise ValueError("input too short for the conv feature extractor")
        ws = self.workspace(B, L)
        feats = torch.empty(B, T, self.cfg.hidden_size, dtype=torch.float32, device=wav.device) if want_feats else None
        logits = None
        if want_logits:
            if self.n_out <= 0:
                raise RuntimeError("no head set")
            if logits_out is not None:
                if (tuple(logits_out.shape) != (B, T, self.n_out) or logits_out.dtype != torch.float32 or
                        not logits_out.is_contiguous() or logits_out.device != wav.device):
                    raise ValueError(f"logits_out must be a contiguous fp32 tensor of shape {(B, T, self.n_out)} on {wav.device}")
                logits = logits_out
            else:
                logits = torch.empty(B, T, self.n_out, dtype=torch.float32, device=wav.device)
        with torch.cuda.device(wav.device):
            check(lib().svt_encoder_forward(self._h, ptr(wav), B, L, ptr(ws), ws.numel(), ptr(feats), ptr(logits),
                                            current_stream_ptr()))
        return feats, logits

    def forward_host(self, wav_host: torch.Tensor, logits_host: torch.Tensor, wav_stage: torch.Tensor,
                     logits_stage: torch.Tensor):
        """HOST (pinned) wav -> H2D -> forward -> D2H logits, synchronous (svt_encoder_forward_host)."""
        B, L = wav_host.shape
        ws = self.workspace(B, L)
        with torch.cuda.device(self.device):
            check(lib().svt_encoder_forward_host(self._h, ptr(wav_host), B, L, ptr(ws), ws.numel(), ptr(wav_stage),
                                                 ptr(logits_stage), ptr(logits_host), current_stream_ptr()))

    def pipeline(self, batch: int, n_samples: int, depth: int = 2) -> "EncoderPipeline":
        """Serving loop for pinned HOST batches of one shape (svt_pipeline_*): copies overlap the previous batch's forward."""
        return EncoderPipeline(self, batch, n_samples, depth)

    def __del__(self):
        try:
            if self._h:
                lib().svt_encoder_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


class EncoderPipeline:
    """`depth` host batches in flight through svt_pipeline_submit / svt_pipeline_wait.

        pipe = engine.pipeline(B, L)
        t = pipe.submit(wav_pinned, logits_pinned)     # returns at once; (B, L) fp32 pinned -> (B, T, n_out) fp32 pinned
        ...                                            # submit the next batch before waiting: that is the overlap
        pipe.wait(t)                                   # logits_pinned is complete
    """

    def __init__(self, engine: EncoderEngine, batch: int, n_samples: int, depth: int = 2):
        self.engine, self.B, self.L, self.depth = engine, batch, n_samples, depth
        self.T = engine.num_frames(n_samples)
        if self.T <= 0:
            raise ValueError("input too short for the conv feature extractor")
        if engine.n_out <= 0:
            raise RuntimeError("no head set")
        dev = engine.device
        need = lib().svt_encoder_workspace_bytes(engine._h, batch, n_samples)
        self._ws = torch.empty(need, dtype=torch.uint8, device=dev)  # private: engine.forward may run concurrently
        self._wav = torch.empty(depth, batch, n_samples, dtype=torch.float32, device=dev)
        self._lg = torch.empty(depth, batch, self.T, engine.n_out, dtype=torch.float32, device=dev)
        self._h = C.c_void_p()
        with torch.cuda.device(dev):
            torch.cuda.synchronize(dev)  # the pipeline's own streams do not order against torch's allocator streams
            check(lib().svt_pipeline_create(engine._h, batch, n_samples, depth, ptr(self._ws), self._ws.numel(), ptr(self._wav),
                                            ptr(self._lg), C.byref(self._h)))

    def _check_host(self, t: torch.Tensor, shape, what: str):
        if t.is_cuda or not t.is_pinned() or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
            raise ValueError(f"{what}: expected a pinned contiguous fp32 host tensor of shape {tuple(shape)}")

    def submit(self, wav_host: torch.Tensor, logits_host: torch.Tensor) -> int:
        self._check_host(wav_host, (self.B, self.L), "wav_host")
        self._check_host(logits_host, (self.B, self.T, self.engine.n_out), "logits_host")
        ticket = C.c_longlong(-1)
        with torch.cuda.device(self.engine.device):
            check(lib().svt_pipeline_submit(self._h, ptr(wav_host), ptr(logits_host), C.byref(ticket)))
        return ticket.value

    def wait(self, ticket: int):
        check(lib().svt_pipeline_wait(self._h, ticket))

    def close(self):
        if self._h:
            lib().svt_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FusionEngine:
    def __init__(self, d_model: int, nhead: int, d_ffn: int, alpha: float, device):
        self.cfg = FusionConfig(d_model, nhead, d_ffn, alpha)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FusionEngine needs a CUDA device (no CPU fallback)")
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().svt_fusion_create(C.byref(self.cfg), C.byref(self._h)))
        self._ws = None

    def load(self, state_dict):
        with torch.cuda.device(self.device):
            _set_tensors(lib().svt_fusion_set_tensor, self._h, state_dict)
            check(lib().svt_fusion_finalize(self._h))

    def forward(self, audio: torch.Tensor, video: torch.Tensor) -> torch.Tensor:
        require_cuda(audio, "FusionEngine.forward")
        require_cuda(video, "FusionEngine.forward")
        audio = audio.to(torch.float32).contiguous()
        video = video.to(torch.float32).contiguous()
        B, Ta, D = audio.shape
        Bv, Tv, Dv = video.shape
        if B != Bv or D != Dv or D != self.cfg.d_model:
            raise ValueError(f"shape mismatch: audio {tuple(audio.shape)} video {tuple(video.shape)}")
        need = lib().svt_fusion_workspace_bytes(self._h, B, Ta)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, Ta, D, dtype=torch.float32, device=audio.device)
        with torch.cuda.device(audio.device):
            check(lib().svt_fusion_forward(self._h, ptr(audio), ptr(video), B, Ta, Tv, ptr(self._ws), self._ws.numel(),
                                           ptr(out), current_stream_ptr()))
        return out

    def __del__(self):
        try:
            if self._h:
                lib().svt_fusion_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


class VideoEngine:
    """svt_video handle (AV-HuBERT video stream) + cached workspace."""

    def __init__(self, cfg: VideoConfig, device):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VideoEngine needs a CUDA device (no CPU fallback)")
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().svt_video_create(C.byref(cfg), C.byref(self._h)))
        self._ws = None

    def load(self, state_dict):
        with torch.cuda.device(self.device):
            _set_tensors(lib().svt_video_set_tensor, self._h, state_dict)
            check(lib().svt_video_finalize(self._h))

    def forward(self, video: torch.Tensor) -> torch.Tensor:
        """video (B, 1, T, 88, 88) CUDA fp32 -> (B, T, D) fp32"""
        require_cuda(video, "VideoEngine.forward")
        if video.dim() != 5 or video.shape[1] != 1 or tuple(video.shape[3:]) != (88, 88):
            raise ValueError(f"expected video of shape (B, 1, T, 88, 88), got {tuple(video.shape)}")
        video = video.to(torch.float32).contiguous()
        B, _, T, _, _ = video.shape
        need = lib().svt_video_workspace_bytes(self._h, B, T)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, T, self.cfg.embed_dim, dtype=torch.float32, device=video.device)
        with torch.cuda.device(video.device):
            check(lib().svt_video_forward(self._h, ptr(video), B, T, ptr(self._ws), self._ws.numel(), ptr(out),
                                          current_stream_ptr()))
        return out

    def __del__(self):
        try:
            if self._h:
                lib().svt_video_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

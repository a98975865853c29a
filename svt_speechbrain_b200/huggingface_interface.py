"""Drop-in for the reference lobe `huggingface_interface.HuggingFaceWav2Vec2`
(MIR_ST500/huggingface_interface.py:47-298; byte-identical copy under N20EMv2/audio_only/).

Same constructor kwargs, same `forward(wav[B, L]) -> feats[B, T, D]` (fp32 in/out on wav.device), same
`state_dict()` keys (`model.` + HF names, so `wav2vec2.pt` files and SpeechBrain `Checkpointer` files load
unchanged), same error behaviour for unknown sources / missing checkpoints.  In a recipe YAML only the class
path changes:

    wav2vec2: !new:svt_speechbrain_b200.huggingface_interface.HuggingFaceWav2Vec2
        source: !ref <wav2vec2_hub>
        freeze: ...

The arithmetic runs in libsvt_b200.so (hand-written sm_100a kernels); the `transformers` module held in
`self.model` is only the parameter container the reference also uses.  Inference only: there is no backward
(`freeze=False` still works for evaluation, but gradients do not flow), and no CPU path.
"""
from __future__ import annotations

import os
import pathlib

import torch
from torch import nn

from .engine import EncoderEngine, encoder_config_from_hf

try:  # the reference raises the same way (huggingface_interface.py:25-38)
    from transformers import (Data2VecAudioConfig, Data2VecAudioModel, HubertConfig, HubertModel, Wav2Vec2Config,
                              Wav2Vec2FeatureExtractor, Wav2Vec2Model, WavLMConfig, WavLMModel)
except ImportError as e:  # pragma: no cover
    raise ImportError("Please install transformers to use the wav2vec2 / HuBERT lobes") from e

# families whose forward is exactly the wav2vec2 graph built in csrc/ (reference table :42-44)
_FAMILIES = {"wav2vec2": (Wav2Vec2Config, Wav2Vec2Model), "hubert": (HubertConfig, HubertModel),
             "data2vec": (Data2VecAudioConfig, Data2VecAudioModel), "wavlm": (WavLMConfig, WavLMModel)}


class HuggingFaceWav2Vec2(nn.Module):
    def __init__(self, source, save_path, pretrain=True, output_norm=True, freeze=True, freeze_feature_extractor=False,
                 apply_spec_augment=False):
        super().__init__()
        # feature extractor config decides input normalisation (reference :103-105,130)
        self.feature_extractor = Wav2Vec2FeatureExtractor.from_pretrained(source, cache_dir=save_path)
        family = None
        for key in ("hubert", "data2vec", "wavlm", "wav2vec2"):  # same substring order as the reference (:108-119)
            if key in source:
                family = key
                break
        if family is None:
            # the reference falls through to an UnboundLocalError here; keep "unknown source" loud but clearer
            raise UnboundLocalError(f"cannot pick a model family from source={source!r} (expected 'wav2vec2' or 'hubert')")
        if family not in _FAMILIES:
            raise NotImplementedError(f"{family}: not a family of the reference lobe (wav2vec2, hubert, data2vec, wavlm)")
        config_cls, model_cls = _FAMILIES[family]
        config = config_cls.from_pretrained(source, cache_dir=save_path)
        config.apply_spec_augment = apply_spec_augment  # inert at eval (HF:1292)
        if pretrain:
            self._from_pretrained(source, config, model_cls, save_path)
        else:
            self.model = model_cls(config)
        self.normalize_wav = self.feature_extractor.do_normalize
        self.freeze = freeze
        self.freeze_feature_extractor = freeze_feature_extractor
        self.output_norm = output_norm
        if self.freeze:
            self.model.eval()
            for p in self.model.parameters():
                p.requires_grad = False
        else:
            self.model.eval()  # inference-only implementation
            if self.freeze_feature_extractor:
                self.model.feature_extractor._freeze_parameters()
        self._engine = None
        self._engine_key = None

    def _from_pretrained(self, source, config, model_cls, save_path):
        """Reference :161-179: a SpeechBrain-pretrained checkpoint (`*.ckpt`, written by HuggingFaceWav2Vec2Pretrain) is
        loaded into a model built from the config; anything else goes through HF `from_pretrained`."""
        is_sb, ckpt_file = self._check_model_source(source)
        if is_sb:
            self.model = model_cls(config)
            self.model.gradient_checkpointing_disable()
            self._load_sb_pretrained_w2v2_parameters(ckpt_file)
        else:
            self.model = model_cls.from_pretrained(source, config=config, cache_dir=save_path)

    def _load_sb_pretrained_w2v2_parameters(self, path):
        """Reference :181-217: keys of a SpeechBrain pre-training checkpoint carry one extra level, `model.wav2vec2.<hf name>`;
        strip it, load non-strictly, warn about what did not transfer."""
        import logging

        log = logging.getLogger(__name__)
        orig = torch.load(path, map_location="cpu")
        modified = {k.replace("model.wav2vec2.", ""): v for k, v in orig.items() if "wav2vec2." in k}
        incompatible = self.model.load_state_dict(modified, strict=False)
        for k in incompatible.missing_keys:
            log.warning(f"During parameter transfer to {type(self.model).__name__} loading from {path}, the transferred "
                        f"parameters did not have parameters for the key: {k}")
        for k in incompatible.unexpected_keys:
            log.warning(f"The param with the key: {k} is discarded as it is useless for wav2vec 2.0 finetuning.")

    @staticmethod
    def _check_model_source(path):
        """Reference :219-261 -> (is_sb, checkpoint_filename).  A local directory holding a HF checkpoint (`*.bin`, or the
        `*.safetensors` newer transformers write) is HF; one holding a `*.ckpt` is SpeechBrain-pretrained; a local directory
        with neither raises FileNotFoundError.  A path that does not exist locally is a hub id: the reference asks the hub
        for its file list (`model_info`); there is no network here, so it is handed to HF `from_pretrained` as is."""
        source = pathlib.Path(path)
        if not source.exists():
            return False, ""
        files = sorted(os.listdir(path))
        if any(f.endswith((".bin", ".safetensors")) for f in files):
            return False, ""
        for f in files:
            if f.endswith(".ckpt"):
                return True, os.path.join(path, f)
        raise FileNotFoundError(f"{path} does not contain a .bin or .ckpt checkpoint !")

    # ------------------------------------------------------------------ engine management
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def engine(self, device) -> EncoderEngine:
        key = (str(device), self.normalize_wav, self.output_norm, self._weights_key())
        if self._engine is None or self._engine_key != key:
            cfg = encoder_config_from_hf(self.model.config, self.normalize_wav, self.output_norm)
            eng = EncoderEngine(cfg, device)
            eng.load({"model." + k: v for k, v in self.model.state_dict().items()})
            self._engine, self._engine_key = eng, key
        return self._engine

    # ------------------------------------------------------------------ reference API
    def forward(self, wav):
        """wav (B, L) -> (B, T, D); reference :263-277."""
        if self.freeze:
            with torch.no_grad():
                return self.extract_features(wav).detach()
        return self.extract_features(wav)

    def extract_features(self, wav):
        """Input LN over the whole (B, L) tensor, encoder, output LN over the whole (B, T, D) tensor
        (reference :279-298) -- all inside svt_encoder_forward."""
        if not wav.is_cuda:
            raise RuntimeError("svt_speechbrain_b200.HuggingFaceWav2Vec2 runs on CUDA (sm_100a) only; no CPU fallback")
        feats, _ = self.engine(wav.device).forward(wav, want_feats=True, want_logits=False)
        return feats

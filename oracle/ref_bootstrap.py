"""Import the REAL reference modules from /root/reference (authoring container only).

Used by oracle/make_golden.py and by tests that are skipped when /root/reference
is absent (it does not exist on the GPU box).  Recipe verified in SURVEY.md App. C:
 * stub `hyperpyyaml` / `ruamel.yaml` (imported at package import by
   speechbrain/core.py:33 and speechbrain/utils/train_logger.py:7),
 * put /root/reference and the recipe directories on sys.path,
 * build an offline HF model dir whose path contains "wav2vec2"
   (class picked by substring, MIR_ST500/huggingface_interface.py:108-119) and a
   dummy *.bin (`_check_model_source`, :230-261).
"""
import os
import sys
import types

REF_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "MIR_ST500"))


def _stub_modules():
    if "hyperpyyaml" not in sys.modules:
        m = types.ModuleType("hyperpyyaml")
        m.resolve_references = m.load_hyperpyyaml = lambda *a, **k: None
        sys.modules["hyperpyyaml"] = m
    if "ruamel" not in sys.modules:
        r = types.ModuleType("ruamel")
        ry = types.ModuleType("ruamel.yaml")
        r.yaml = ry
        sys.modules["ruamel"], sys.modules["ruamel.yaml"] = r, ry


def setup_paths():
    if not available():
        raise RuntimeError("/root/reference not present (GPU box?)")
    _stub_modules()
    for p in (
        os.path.join(REF_ROOT, "N20EMv2", "video_only"),
        os.path.join(REF_ROOT, "N20EMv2", "audio_visual"),
        os.path.join(REF_ROOT, "MIR_ST500"),
        REF_ROOT,
    ):
        if p not in sys.path:
            sys.path.insert(0, p)


def make_offline_model_dir(path, config_kwargs, seed=0, family="wav2vec2"):
    """Random-init HF Wav2Vec2Model / HubertModel + feature extractor saved under `path` (the reference picks the
    HF class by substring of the path, huggingface_interface.py:108-119)."""
    import torch
    from transformers import (Data2VecAudioConfig, Data2VecAudioModel, HubertConfig, HubertModel, Wav2Vec2Config,
                              Wav2Vec2FeatureExtractor, Wav2Vec2Model)

    assert family in path
    os.makedirs(path, exist_ok=True)
    torch.manual_seed(seed)
    if family == "hubert":
        HubertModel(HubertConfig(**config_kwargs)).save_pretrained(path)
    elif family == "data2vec":
        Data2VecAudioModel(Data2VecAudioConfig(**config_kwargs)).save_pretrained(path)
    elif family == "wavlm":
        from transformers import WavLMConfig, WavLMModel
        WavLMModel(WavLMConfig(**config_kwargs)).save_pretrained(path)
    else:
        Wav2Vec2Model(Wav2Vec2Config(**config_kwargs)).save_pretrained(path)
    Wav2Vec2FeatureExtractor(
        feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True, return_attention_mask=True
    ).save_pretrained(path)
    open(os.path.join(path, "dummy.bin"), "wb").close()
    return path


def reference_lobe(model_dir, output_norm=True):
    """The reference's own HuggingFaceWav2Vec2 (MIR_ST500/huggingface_interface.py:47)."""
    setup_paths()
    from huggingface_interface import HuggingFaceWav2Vec2  # reference module

    return HuggingFaceWav2Vec2(source=model_dir, save_path=model_dir, output_norm=output_norm, freeze=False).eval()


def reference_linear(input_size, n_neurons=20):
    setup_paths()
    import speechbrain as sb

    return sb.nnet.linear.Linear(input_size=input_size, n_neurons=n_neurons)


def reference_fusion(**kw):
    setup_paths()
    from fusion import FusionRCA  # N20EMv2/audio_visual/fusion.py:186

    return FusionRCA(**kw).eval()


def reference_res_encoder():
    """The reference's own lip front end, N20EMv2/video_only/resnet.py:133 (needs only torch)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_video_resnet", os.path.join(REF_ROOT, "N20EMv2", "video_only", "resnet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ResEncoder(relu_type="prelu", weights=None).eval()


def reference_frame2note():
    setup_paths()
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_mir_utils", os.path.join(REF_ROOT, "MIR_ST500", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.frame2note

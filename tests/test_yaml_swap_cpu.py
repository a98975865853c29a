"""The plugin boundary is the recipes' hparams YAML: `!new:huggingface_interface.HuggingFaceWav2Vec2`,
`!new:speechbrain.nnet.linear.Linear`, `!new:fusion.FusionRCA`, `!new:fairseq_interface.FairseqAVHubertPretrain`
(MIR_ST500/hparams/train_audio_ssl.yaml:95-99,113-119; audio_visual/hparams/train_rca_av.yaml:80-91;
video_only/hparams/train_video_ssl.yaml:90-103).  These tests load those files -- the committed excerpts, and the
reference's full files when /root/reference is present -- with ONLY the class strings replaced, instantiate the modules
the way `Brain.__init__` does (`torch.nn.ModuleDict(modules)`, speechbrain/core.py:508) and check the loader paths of the
lobes: `pretrain=True` from a local HF directory and from a SpeechBrain `.ckpt` (huggingface_interface.py:160-261)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from hyperyaml_subset import Hparams  # noqa: E402

import svt_speechbrain_b200 as svt  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
CLASS_MAP = {
    "huggingface_interface.HuggingFaceWav2Vec2": "svt_speechbrain_b200.huggingface_interface.HuggingFaceWav2Vec2",
    "speechbrain.nnet.linear.Linear": "svt_speechbrain_b200.linear.Linear",
    "fusion.FusionRCA": "svt_speechbrain_b200.fusion.FusionRCA",
    "fairseq_interface.FairseqAVHubertPretrain": "svt_speechbrain_b200.fairseq_interface.FairseqAVHubertPretrain",
}
TINY = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, feat_extract_norm="layer",
            conv_bias=True, do_stable_layer_norm=True, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=2)


def _yaml_sources(excerpt, ref_rel):
    out = [("excerpt", os.path.join(HERE, "golden", "hparams", excerpt))]
    if os.path.exists(os.path.join(REF, ref_rel)):
        out.append(("reference file", os.path.join(REF, ref_rel)))
    return out


def _hf_dir(tmp_path, name="wav2vec2-tiny", with_weights=True):
    from transformers import Wav2Vec2Config, Wav2Vec2FeatureExtractor, Wav2Vec2Model

    d = str(tmp_path / name)
    os.makedirs(d)
    torch.manual_seed(0)
    m = Wav2Vec2Model(Wav2Vec2Config(**TINY)).eval()
    if with_weights:
        m.save_pretrained(d)
    else:
        m.config.save_pretrained(d)
    Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True,
                             return_attention_mask=True).save_pretrained(d)
    return d, m


@pytest.mark.parametrize("which", [0, 1])
def test_audio_recipe_yaml_with_only_the_class_strings_changed(tmp_path, which):
    srcs = _yaml_sources("train_audio_ssl_excerpt.yaml", "MIR_ST500/hparams/train_audio_ssl.yaml")
    if which >= len(srcs):
        pytest.skip("/root/reference not present")
    d, hf = _hf_dir(tmp_path)
    hp = Hparams(open(srcs[which][1]).read(), class_map=CLASS_MAP,
                 overrides={"wav2vec2_hub": d, "wav2vec2_local": d, "feat_dim": 128, "data_folder": str(tmp_path)})
    modules = torch.nn.ModuleDict(hp["modules"])           # Brain.__init__, speechbrain/core.py:508
    lobe, head = modules["wav2vec2"], modules["model"]
    assert type(lobe) is svt.HuggingFaceWav2Vec2 and type(head) is svt.Linear
    assert lobe is hp["wav2vec2"] and head is hp["model"]  # `!ref <wav2vec2>` shares the instance
    assert lobe.output_norm is True and lobe.freeze is False and lobe.normalize_wav is True
    assert head.w.weight.shape == (hp["output_neurons"], 128)
    # pretrain=True (the YAML passes no `pretrain`): the weights are the saved ones, keys are `model.` + HF names
    want = {"model." + k: v for k, v in hf.state_dict().items()}
    got = lobe.state_dict()
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    # scalars the inference path reads
    assert hp["frame_rate"] == 49.8 and hp["onset_threshold"] == 0.4 and hp["offset_threshold"] == 0.5
    assert hp["sample_rate"] == 16000 and hp["dur_threshold"] == 5 and hp["test_batch_size"] == 1
    hpar = svt.AMTHparams(sample_rate=hp["sample_rate"], frame_rate=hp["frame_rate"], dur_threshold=hp["dur_threshold"],
                          onset_threshold=hp["onset_threshold"], offset_threshold=hp["offset_threshold"],
                          pitch_octave_num=hp["pitch_octave_num"], pitch_class_num=hp["pitch_class_num"])
    assert hpar.n_out == hp["output_neurons"]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):                  # no CPU fallback behind the drop-in
            lobe(torch.zeros(1, 16000))


@pytest.mark.parametrize("which", [0, 1])
def test_audio_visual_recipe_yaml(which):
    srcs = _yaml_sources("train_rca_av_excerpt.yaml", "N20EMv2/audio_visual/hparams/train_rca_av.yaml")
    if which >= len(srcs):
        pytest.skip("/root/reference not present")
    hp = Hparams(open(srcs[which][1]).read(), class_map=CLASS_MAP, overrides={"data_folder": "/nonexistent"})
    modules = torch.nn.ModuleDict(hp["modules"])
    assert type(modules["fusion"]) is svt.FusionRCA and type(modules["head"]) is svt.Linear
    model = hp["model"]                                    # !new:torch.nn.ModuleList [[fusion, head]]
    assert isinstance(model, torch.nn.ModuleList) and model[0] is modules["fusion"] and model[1] is modules["head"]
    keys = set(modules["fusion"].state_dict())
    assert len(keys) == 25 and "fusion.layer1.self_att.att.in_proj_weight" in keys and "fusion.positional_encoding.pe" in keys
    assert modules["head"].w.weight.shape == (20, 1024)


@pytest.mark.parametrize("which", [0, 1])
def test_video_recipe_yaml_and_checkpoint_staging(tmp_path, monkeypatch, which):
    """`encoder: !new:fairseq_interface.FairseqAVHubertPretrain` with `pretrained_path` pointing at a local checkpoint: the file
    is staged at the YAML's literal `save_path` (relative to the recipe's cwd, as the reference's download_file does,
    video_only/fairseq_interface.py:399) and its `model` tensors are loaded."""
    srcs = _yaml_sources("train_video_ssl_excerpt.yaml", "N20EMv2/video_only/hparams/train_video_ssl.yaml")
    if which >= len(srcs):
        pytest.skip("/root/reference not present")
    from svt_speechbrain_b200.fairseq_interface import FairseqAVHubertPretrain

    small = dict(encoder_embed_dim=128, encoder_layers=1, encoder_attention_heads=2, encoder_ffn_embed_dim=256, conv_pos=16,
                 conv_pos_groups=2)
    torch.manual_seed(1)
    donor = FairseqAVHubertPretrain(None, None, pretrain=False, model_config=small)
    ckpt = str(tmp_path / "large_vox_iter5_local.pt")
    torch.save({"cfg": {"model": dict(small)}, "model": donor.model.state_dict()}, ckpt)
    monkeypatch.chdir(tmp_path)
    hp = Hparams(open(srcs[which][1]).read(), class_map=CLASS_MAP,
                 overrides={"avhubert_url": ckpt, "feat_dim": 128, "data_folder": str(tmp_path)})
    modules = torch.nn.ModuleDict(hp["modules"])
    enc = modules["encoder"]
    assert type(enc) is FairseqAVHubertPretrain and enc.output_norm is True and enc.freeze is False
    assert os.path.isfile(tmp_path / "ssl_model" / "AVHuBERT" / "large_vox_iter5.pt")
    assert enc.model_config["encoder_embed_dim"] == 128 and enc.model_config["encoder_layers"] == 1
    for k, v in donor.state_dict().items():
        assert torch.equal(enc.state_dict()[k], v), k
    with pytest.raises(FileNotFoundError):                 # a URL cannot be fetched here: loud, not random weights
        FairseqAVHubertPretrain("https://dl.fbaipublicfiles.com/avhubert/model/x.pt", str(tmp_path / "nope.pt"))


def test_check_model_source_contract(tmp_path):
    """(is_sb, checkpoint_filename) exactly as the reference returns it (huggingface_interface.py:219-261)."""
    chk = svt.HuggingFaceWav2Vec2._check_model_source
    d, _ = _hf_dir(tmp_path, "wav2vec2-hf")
    assert chk(d) == (False, "")
    open(os.path.join(d, "pytorch_model.bin"), "wb").close()
    assert chk(d) == (False, "")
    sb = tmp_path / "wav2vec2-sb"
    sb.mkdir()
    (sb / "wav2vec2.ckpt").write_bytes(b"")
    assert chk(str(sb)) == (True, os.path.join(str(sb), "wav2vec2.ckpt"))
    empty = tmp_path / "wav2vec2-empty"
    empty.mkdir()
    with pytest.raises(FileNotFoundError, match="does not contain a .bin or .ckpt checkpoint"):
        chk(str(empty))
    assert chk("facebook/wav2vec2-large-lv60") == (False, "")   # not local: a hub id, handed to from_pretrained


def test_speechbrain_ckpt_source_is_loaded_with_the_wav2vec2_level_stripped(tmp_path, caplog):
    """A directory holding config + `*.ckpt` written by SpeechBrain's wav2vec2 pre-training: keys `model.wav2vec2.<hf name>`
    (huggingface_interface.py:181-217); extra keys are discarded with a warning, missing ones reported."""
    d, hf = _hf_dir(tmp_path, "wav2vec2-sbpretrained", with_weights=False)
    sd = {"model.wav2vec2." + k: v.clone() for k, v in hf.state_dict().items()}
    dropped = "model.wav2vec2.encoder.layer_norm.bias"
    del sd[dropped]
    sd["model.wav2vec2.quantizer.codevectors"] = torch.zeros(3)
    sd["model.project_q.weight"] = torch.zeros(3)           # no "wav2vec2." level: ignored silently, as in the reference
    torch.save(sd, os.path.join(d, "save.ckpt"))
    with caplog.at_level("WARNING"):
        lobe = svt.HuggingFaceWav2Vec2(source=d, save_path=d)   # pretrain=True
    got = lobe.model.state_dict()
    for k, v in hf.state_dict().items():
        if "model.wav2vec2." + k != dropped:
            assert torch.equal(got[k], v), k
    assert any("encoder.layer_norm.bias" in r.message for r in caplog.records)
    assert any("quantizer.codevectors" in r.message for r in caplog.records)


def test_unknown_family_and_missing_checkpoint_errors(tmp_path):
    d, _ = _hf_dir(tmp_path, "wav2vec2-nockpt", with_weights=False)
    with pytest.raises(FileNotFoundError):
        svt.HuggingFaceWav2Vec2(source=d, save_path=d)
    other = tmp_path / "someothermodel"
    os.rename(d, other)
    with pytest.raises(UnboundLocalError):                 # the reference's own failure mode for an unknown family (:108-119)
        svt.HuggingFaceWav2Vec2(source=str(other), save_path=str(other), pretrain=False)


def test_data2vec_config_without_wav2vec2_only_fields():
    """Data2VecAudioConfig defines neither `feat_extract_norm` nor `do_stable_layer_norm`: the config mapping must not
    depend on them (layer-norm feature extractor, post-LN encoder)."""
    from transformers import Data2VecAudioConfig
    from svt_speechbrain_b200.engine import encoder_config_from_hf

    cfg = Data2VecAudioConfig()
    for f in ("feat_extract_norm", "do_stable_layer_norm"):
        if hasattr(cfg, f):
            delattr(cfg, f)
    c = encoder_config_from_hf(cfg, True, True)
    assert c.feat_norm_layer == 1 and c.stable_layer_norm == 0 and c.pos_conv_layers == 5 and c.pos_conv_kernel == 19

"""Oracle vs the committed golden vectors (made from the real reference by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo
from oracle import make_golden as mg
from oracle import wav2vec2_oracle as wo

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name,cfg", [("w2v2_tiny_large", mg.TINY_LARGE), ("w2v2_tiny_base", mg.TINY_BASE)])
def test_tiny_stored_weights(name, cfg):
    g = _load(name)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    head = {k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("head/")}
    wav = torch.from_numpy(g["wav"])
    with torch.no_grad():
        feats = wo.lobe_forward(cfg, sd, wav)
        logits = wo.head_forward(head, feats)
    # fp32 CPU both sides; only op-ordering differences (fused HF sdpa vs explicit softmax)
    np.testing.assert_allclose(logits.numpy(), g["logits"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(feats[:, :, :8].numpy(), g["feats_head"], atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("name,cfg", [("w2v2_large_1s", wo.W2V2Config.large()), ("w2v2_base_1s", wo.W2V2Config.base()),
                                      ("hubert_base_1s", wo.W2V2Config.hubert_base()),
                                      ("hubert_large_1s", wo.W2V2Config.hubert_large()),
                                      ("data2vec_base_1s", wo.W2V2Config.data2vec_base()),
                                      ("hubert_base_posbn_1s", wo.W2V2Config(family="hubert", feat_proj_layer_norm=False,
                                                                             conv_pos_batch_norm=True)),
                                      ("wavlm_base_1s", wo.W2V2Config.wavlm_base()),
                                      ("wavlm_large_1s", wo.W2V2Config.wavlm_large()),
                                      ("wavlm_base_5s", wo.W2V2Config.wavlm_base())])
def test_full_arch_seeded_weights(name, cfg):
    g = _load(name)
    sd = mg.perturb_norm_affines(wo.random_weights(cfg, seed=int(g["weight_seed"])), seed=int(g["affine_seed"]))
    head = wo.random_head(cfg.hidden_size, 20, seed=int(g["head_seed"]))
    wav = mg.synth_wav(int(g["B"]), int(g["L"]), seed=int(g["wav_seed"]))
    with torch.no_grad():
        logits = wo.amt_logits(cfg, sd, head, wav)
    assert logits.shape == g["logits"].shape
    np.testing.assert_allclose(logits.numpy(), g["logits"], atol=1e-4, rtol=1e-4)


def test_num_frames():
    cfg = wo.W2V2Config.large()
    assert cfg.num_frames(160000) == 499 and cfg.num_frames(80000) == 249 and cfg.num_frames(16000) == 49


def test_whole_tensor_norm_couples_clips():
    x = torch.randn(3, 50)
    y = wo.whole_tensor_layer_norm(x)
    ref = torch.nn.functional.layer_norm(x, x.shape)
    np.testing.assert_allclose(y.numpy(), ref.numpy(), atol=1e-6)


def test_fusion_tiny():
    g = _load("fusion_tiny")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    out = fo.fusion_forward(sd, torch.from_numpy(g["a"]), torch.from_numpy(g["v"]), nhead=int(g["nhead"]))
    np.testing.assert_allclose(out.numpy(), g["out"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(fo.positional_encoding(64, int(g["D"]))[:, :16].numpy(), g["pe_head"], atol=1e-6)


@pytest.mark.parametrize("name", ["fusion_full", "fusion_full_pad"])
def test_fusion_full_seeded(name):
    g = _load(name)
    D = int(g["D"])
    sd = mg.random_fusion_weights(D, int(g["d_ffn"]), seed=int(g["w_seed"]))
    gen = torch.Generator().manual_seed(int(g["x_seed"]))
    a = torch.randn(int(g["B"]), int(g["Ta"]), D, generator=gen)
    v = torch.randn(int(g["B"]), int(g["Tv"]), D, generator=gen)
    out = fo.fusion_forward(sd, a, v, nhead=int(g["nhead"]))
    np.testing.assert_allclose(out.numpy(), g["out"], atol=5e-5, rtol=1e-4)


def test_avhubert_resnet_vs_reference_golden():
    """oracle/avhubert_oracle.res_encoder against the reference's own resnet.ResEncoder (N20EMv2/video_only/resnet.py)
    run by oracle/make_golden.py on seeded weights / input."""
    from oracle import avhubert_oracle as av

    d = np.load(os.path.join(GOLD, "avhubert_resnet_b2_t6.npz"))
    sd = av.random_weights(av.AVHubertConfig(encoder_layers=0), seed=int(d["weight_seed"]))
    g = torch.Generator().manual_seed(int(d["video_seed"]))
    video = torch.randn(int(d["B"]), 1, int(d["T"]), 88, 88, generator=g)
    with torch.no_grad():
        out = av.res_encoder(sd, video, "model.feature_extractor_video.resnet.")
    assert out.shape == (2, 512, 6)
    assert float((out - torch.from_numpy(d["out"])).abs().max()) < 1e-4


def test_avhubert_tiny_forward_shapes_and_key_translation():
    from oracle import avhubert_oracle as av

    cfg = av.AVHubertConfig(encoder_embed_dim=128, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_embed_dim=256,
                            conv_pos=16, conv_pos_groups=4)
    sd = av.random_weights(cfg, seed=1)
    assert "model.encoder.layers.1.self_attn_layer_norm.weight" in sd and "model.encoder.pos_conv.0.weight_g" in sd
    video = torch.randn(1, 1, 5, 88, 88, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        y = av.lobe_forward(cfg, sd, video)
    assert y.shape == (1, 5, 128) and abs(float(y.mean())) < 1e-4 and abs(float(y.var(unbiased=False)) - 1.0) < 1e-3

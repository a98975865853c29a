"""AV-HuBERT video stream (csrc/video.cu) vs the CPU oracle (oracle/avhubert_oracle.py, whose ResNet front end is
pinned to the reference's resnet.py) through the C ABI / the FairseqAVHubertPretrain drop-in."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _lobe(cfg, sd, output_norm=True, input_norm=None):
    from svt_speechbrain_b200.fairseq_interface import FairseqAVHubertPretrain
    mc = dict(encoder_embed_dim=cfg.encoder_embed_dim, encoder_layers=cfg.encoder_layers,
              encoder_attention_heads=cfg.encoder_attention_heads, encoder_ffn_embed_dim=cfg.encoder_ffn_embed_dim,
              conv_pos=cfg.conv_pos, conv_pos_groups=cfg.conv_pos_groups)
    lobe = FairseqAVHubertPretrain(None, None, input_norm=input_norm, output_norm=output_norm, pretrain=False, model_config=mc)
    own = lobe.state_dict()
    new = {k: sd[k] for k in own if k in sd}
    missing = [k for k in own if k not in new and "num_batches_tracked" not in k and "feature_extractor_audio" not in k
               and not k.endswith("mask_emb")]
    assert not missing, missing[:5]
    lobe.load_state_dict(new, strict=False)
    return lobe


def test_resnet_front_end_vs_reference_golden():
    """Zero transformer layers and identity glue: the stream reduces to ResEncoder -> proj; compare with the reference's own
    resnet.ResEncoder output (golden) pushed through the same proj / LN / post-proj on the CPU."""
    from oracle import avhubert_oracle as av
    d = np.load(os.path.join(GOLD, "avhubert_resnet_b2_t6.npz"))
    cfg = av.AVHubertConfig(encoder_embed_dim=256, encoder_layers=0, encoder_attention_heads=4, encoder_ffn_embed_dim=512,
                            conv_pos=16, conv_pos_groups=4)
    sd = av.random_weights(cfg, seed=int(d["weight_seed"]))
    g = torch.Generator().manual_seed(int(d["video_seed"]))
    video = torch.randn(int(d["B"]), 1, int(d["T"]), 88, 88, generator=g)
    taps = {}
    with torch.no_grad():
        ref = av.lobe_forward(cfg, sd, video, output_norm=False, taps=taps)
    # the oracle's front end is the golden (bit-exact on CPU): check that first, then the CUDA path against the oracle
    res = av.res_encoder(sd, video, "model.feature_extractor_video.resnet.")
    assert float((res - torch.from_numpy(d["out"])).abs().max()) < 1e-4
    got = _lobe(cfg, sd, output_norm=False)({"video": video.cuda(), "audio": None}).cpu()
    print(f"video stream (0 layers): max-abs {float((got - ref).abs().max()):.3e} rel-L2 {_rel(got, ref):.3e}")
    assert got.shape == ref.shape == (2, 6, 256)
    assert _rel(got, ref) < 2e-2


@pytest.mark.parametrize("T,B,input_norm", [(7, 2, False), (12, 1, True)])
def test_video_stream_small_transformer_vs_oracle(T, B, input_norm):
    from oracle import avhubert_oracle as av
    cfg = av.AVHubertConfig(encoder_embed_dim=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_embed_dim=512,
                            conv_pos=32, conv_pos_groups=4)
    sd = av.random_weights(cfg, seed=3)
    video = torch.randn(B, 1, T, 88, 88, generator=torch.Generator().manual_seed(T))
    with torch.no_grad():
        ref = av.lobe_forward(cfg, sd, video, input_norm=input_norm, output_norm=True)
    got = _lobe(cfg, sd, output_norm=True, input_norm=input_norm)({"video": video.cuda(), "audio": None}).cpu()
    print(f"video stream T={T} B={B}: max-abs {float((got - ref).abs().max()):.3e} rel-L2 {_rel(got, ref):.3e}")
    assert torch.isfinite(got).all()
    assert _rel(got, ref) < 2e-2 and float((got - ref).abs().max()) < 0.15


def test_video_stream_large_config_short_clip():
    """AV-HuBERT-large geometry (24 x 1024 / 4096 / 16 heads), 1 s of video."""
    from oracle import avhubert_oracle as av
    cfg = av.AVHubertConfig()
    sd = av.random_weights(cfg, seed=0)
    video = torch.randn(1, 1, 50, 88, 88, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        ref = av.lobe_forward(cfg, sd, video)
    got = _lobe(cfg, sd)({"video": video.cuda(), "audio": None}).cpu()
    print(f"video stream large T=50: max-abs {float((got - ref).abs().max()):.3e} rel-L2 {_rel(got, ref):.3e}")
    assert _rel(got, ref) < 2e-2


def test_state_dict_keys_follow_the_reference_names():
    from svt_speechbrain_b200.fairseq_interface import FairseqAVHubertPretrain
    lobe = FairseqAVHubertPretrain(None, None, pretrain=False,
                                   model_config=dict(encoder_embed_dim=256, encoder_layers=1, encoder_attention_heads=4,
                                                     encoder_ffn_embed_dim=512, conv_pos=16, conv_pos_groups=4))
    keys = set(lobe.state_dict())
    for k in ("model.feature_extractor_video.resnet.frontend3D.0.weight", "model.feature_extractor_video.resnet.frontend3D.1.running_var",
              "model.feature_extractor_video.resnet.trunk.layer2.0.downsample.0.weight",
              "model.feature_extractor_video.resnet.trunk.layer4.1.relu2.weight", "model.feature_extractor_video.proj.weight",
              "model.feature_extractor_audio.proj.bias", "model.layer_norm.weight", "model.post_extract_proj.weight",
              "model.mask_emb", "model.encoder.pos_conv.0.weight_g", "model.encoder.pos_conv.0.weight_v", "model.encoder.pos_conv.0.bias",
              "model.encoder.layers.0.self_attn.q_proj.weight", "model.encoder.layers.0.self_attn_layer_norm.bias",
              "model.encoder.layers.0.fc1.weight", "model.encoder.layers.0.final_layer_norm.weight", "model.encoder.layer_norm.weight"):
        assert k in keys, k
    with pytest.raises(RuntimeError):
        lobe({"video": torch.zeros(1, 1, 4, 88, 88), "audio": None})


def test_audio_visual_pipeline_vs_composed_oracle():
    """BASELINE config 4 in miniature: wav2vec2-large (1 s) + AV-HuBERT-large (50 frames) + FusionRCA + Linear head on the GPU
    against the same composition of the CPU oracles; notes decoded from the GPU logits equal the oracle decoder's."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import svt_speechbrain_b200 as svt
    from oracle import avhubert_oracle as av
    from oracle import fusion_oracle as fo
    from oracle import make_golden as mg
    from oracle import wav2vec2_oracle as wo
    from oracle.frame2note_oracle import frame2note as f2n_oracle
    from test_gpu_e2e import _build

    wcfg = wo.W2V2Config.large()
    alobe, lin, asd, head = _build(wcfg)
    vcfg = av.AVHubertConfig()
    vsd = av.random_weights(vcfg, seed=0)
    vlobe = _lobe(vcfg, vsd)
    fsd = mg.random_fusion_weights(1024, 3072, seed=3)
    fus = svt.FusionRCA()
    full = dict(fus.state_dict())
    full.update(fsd)
    fus.load_state_dict(full, strict=True)
    fus = fus.cuda()
    g = torch.Generator().manual_seed(21)
    wav = torch.randn(2, 16000, generator=g)
    video = torch.randn(2, 1, 50, 88, 88, generator=g)
    with torch.no_grad():
        a = wo.lobe_forward(wcfg, asd, wav)
        v = av.lobe_forward(vcfg, vsd, video)
        ref = wo.head_forward(head, fo.fusion_forward(fsd, a, v))
    tr = svt.AVTranscriber(alobe, vlobe, fus, lin)
    got = tr.logits(wav.cuda(), video.cuda())
    rel = _rel(got.cpu(), ref)
    print(f"AV pipeline logits: max-abs {float((got.cpu() - ref).abs().max()):.3e} rel-L2 {rel:.3e}")
    assert got.shape == ref.shape == (2, 49, 20)
    assert rel < 2e-2
    notes = tr.decode(got[0])
    p_on, p_off, octv, pc = wo.frame_info_from_logits(got[0].cpu())
    fi = [(p_on[i], p_off[i], int(octv[i]), int(pc[i])) for i in range(len(p_on))]
    want = np.array(f2n_oracle(fi, 0.4, 0.5, 1 / 49.8), dtype=np.float64).reshape(-1, 3)
    assert notes.shape == want.shape and np.array_equal(notes, want)


def test_eval_transform_matches_reference_compose():
    """svt_video_transform_u8 == Compose([Normalize(0, 255), CenterCrop(88), Normalize(0.421, 0.165)]) of the video
    recipe (video_only/train_video_ssl.py:454-457; numpy float64 arithmetic there, fp32 here)."""
    import numpy as np
    from svt_speechbrain_b200 import video_transforms as vt

    rng = np.random.default_rng(0)
    for (T, H, W) in [(7, 96, 96), (3, 88, 88), (5, 101, 97)]:
        frames = rng.integers(0, 256, (2, T, H, W), dtype=np.uint8)
        x = (frames[0].astype(np.float64) - 0.0) / 255.0                     # Normalize(0, 255)
        dw, dh = int(round((W - 88)) / 2.), int(round((H - 88)) / 2.)        # CenterCrop (utils.py:79-81)
        x = x[:, dh:dh + 88, dw:dw + 88]
        x = (x - 0.421) / 0.165                                              # Normalize(mean, std)
        got = vt.eval_transform(torch.from_numpy(frames).cuda())
        assert got.shape == (2, 1, T, 88, 88) and got.dtype == torch.float32
        assert np.abs(got[0, 0].cpu().numpy() - x).max() < 1e-6
    with pytest.raises(RuntimeError):
        vt.eval_transform(torch.zeros(2, 96, 96, dtype=torch.uint8))


def test_stage_c_from_cached_features(tmp_path):
    """feature_cache.transcribe_from_cache: cached per-song audio / video features -> utterance slicing of the reference
    recipe -> FusionRCA + head (batched) -> one decode == the same utterances pushed one by one (train_rca_av.py:28-51)."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import svt_speechbrain_b200 as svt
    from oracle import make_golden as mg
    from oracle import wav2vec2_oracle as wo
    from svt_speechbrain_b200 import feature_cache as fc

    fus = svt.FusionRCA()
    full = dict(fus.state_dict())
    full.update(mg.random_fusion_weights(1024, 3072, seed=3))
    fus.load_state_dict(full, strict=True)
    fus = fus.cuda()
    lin = svt.Linear(n_neurons=20, input_size=1024)
    lin.load_state_dict(wo.random_head(1024, 20, seed=0))
    lin = lin.cuda()
    g = torch.Generator().manual_seed(4)
    audio = torch.randn(611, 1024, generator=g)    # 12.3 s of 49.8 fps features -> utterances of 249 / 249 / 113 frames
    video = torch.randn(600, 1024, generator=g)    # shorter than the audio: the last utterance is zero-padded
    pa = fc.save_song_features(audio, fc.audio_feats_path(str(tmp_path / "s")))
    pv = fc.save_song_features(video, fc.video_feats_path(str(tmp_path / "s")))
    from svt_speechbrain_b200.amt import decode_logits
    notes = fc.transcribe_from_cache(fus, lin, svt.AMTHparams(), pa, pv, utter_num=3)
    pieces = []
    for u in (1, 2, 3):
        a, v = fc.load_av_utterance(pa, pv, u, 3)
        assert a.shape == v.shape
        pieces.append(lin(fus(a[None].cuda(), v[None].cuda()))[0])
    want = decode_logits(torch.cat(pieces), svt.AMTHparams())
    assert sum(p.shape[0] for p in pieces) == 611
    assert notes.shape == want.shape and np.allclose(notes, want)

"""Generate tests/golden/*.npz from the REAL reference (authoring container only).

Run:  python -m oracle.make_golden
Every fixture is produced by the reference's own modules imported from /root/reference
(oracle/ref_bootstrap.py): HuggingFaceWav2Vec2 (MIR_ST500/huggingface_interface.py:47),
speechbrain.nnet.linear.Linear, FusionRCA (N20EMv2/audio_visual/fusion.py:186) and
frame2note (MIR_ST500/utils.py:82).  Weights for the full-size architectures are NOT stored:
they are regenerated from seeds (HF `_init_weights` under torch.manual_seed, or the seeded
generators below), loaded into the reference with load_state_dict, and only inputs-by-seed +
reference outputs are committed.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np
import torch

from . import ref_bootstrap as rb
from .wav2vec2_oracle import W2V2Config, random_head, random_weights

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY_LARGE = W2V2Config(
    hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, conv_dim=(32,) * 7,
    conv_bias=True, feat_extract_norm="layer", do_stable_layer_norm=True, num_conv_pos_embeddings=16,
    num_conv_pos_embedding_groups=4,
)
TINY_BASE = W2V2Config(
    hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, conv_dim=(32,) * 7,
    conv_bias=False, feat_extract_norm="group", do_stable_layer_norm=False, num_conv_pos_embeddings=16,
    num_conv_pos_embedding_groups=4,
)


def synth_wav(B, L, seed=1986):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, L, generator=g)


def perturb_norm_affines(sd, seed=7):
    """HF init leaves every LayerNorm at weight=1,bias=0 and conv biases tiny; perturb them (seeded)
    so that affine terms and biases are actually exercised by the parity tests."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if "layer_norm" in k and k.endswith(".weight"):
            v = v + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("batch_norm.weight"):  # HuBERT conv_pos_batch_norm: exercise the running statistics too
            v = v + 0.2 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_mean"):
            v = v + 0.3 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            v = v * (0.5 + torch.rand(v.shape, generator=g))
        elif k.endswith("gru_rel_pos_const"):  # WavLM: ones at init
            v = v + 0.3 * torch.randn(v.shape, generator=g)
        elif k.endswith("rel_attn_embed.weight"):  # WavLM: make the position bias comparable to the scores
            v = v * 3.0
        elif k.endswith(".bias"):
            v = v + 0.05 * torch.randn(v.shape, generator=g)
        out[k] = v
    return out


def random_fusion_weights(D=1024, d_ffn=3072, seed=3):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def rn(*s, scale):
        return torch.randn(*s, generator=g) * scale

    for l in (1, 2):
        p = f"fusion.layer{l}."
        sd[p + "self_att.att.in_proj_weight"] = rn(3 * D, D, scale=D ** -0.5)
        sd[p + "self_att.att.in_proj_bias"] = rn(3 * D, scale=0.05)
        sd[p + "self_att.att.out_proj.weight"] = rn(D, D, scale=D ** -0.5)
        sd[p + "self_att.att.out_proj.bias"] = rn(D, scale=0.05)
        sd[p + "pos_ffn.ffn.0.weight"] = rn(d_ffn, D, scale=D ** -0.5)
        sd[p + "pos_ffn.ffn.0.bias"] = rn(d_ffn, scale=0.05)
        sd[p + "pos_ffn.ffn.3.weight"] = rn(D, d_ffn, scale=d_ffn ** -0.5)
        sd[p + "pos_ffn.ffn.3.bias"] = rn(D, scale=0.05)
        for nm in ("norm1", "norm2"):
            sd[p + nm + ".norm.weight"] = 1.0 + rn(D, scale=0.1)
            sd[p + nm + ".norm.bias"] = rn(D, scale=0.05)
    return sd


def random_frames(n, seed, p_hi=0.5):
    """Synthetic frame_info with many onsets/offsets and pitch ties (to exercise set-order tie-breaks)."""
    rng = np.random.default_rng(seed)
    lo_on = rng.normal(-1.0, 2.0, n).astype(np.float32)
    lo_off = rng.normal(-1.0, 2.0, n).astype(np.float32)
    # plateaus: repeat some onset logits so that `== max(window)` ties occur
    rep = rng.random(n) < 0.2
    lo_on[1:][rep[1:]] = lo_on[:-1][rep[1:]]
    octv = rng.integers(0, 5, n)
    pc = rng.integers(0, 13, n)
    few = rng.random() < p_hi  # narrow pitch alphabet => frequent ties
    if few:
        octv = rng.integers(0, 2, n)
        pc = rng.choice([0, 8, 1, 4, 12], n)
    return lo_on, lo_off, octv.astype(np.int64), pc.astype(np.int64)


def _ref_lobe_with(cfg: W2V2Config, sd, tmp):
    d = os.path.join(tmp, f"{cfg.family}-{abs(hash(str(cfg))) % 10**8}")
    rb.make_offline_model_dir(d, cfg.hf_kwargs(), seed=0, family=cfg.family)
    lobe = rb.reference_lobe(d, output_norm=True)
    missing = lobe.load_state_dict(sd, strict=True)
    return lobe


def gold_w2v2(name, cfg, B, L, tmp, store_weights, taps_keep=8):
    sd = perturb_norm_affines(random_weights(cfg, seed=0))
    head = random_head(cfg.hidden_size, 20, seed=0)
    wav = synth_wav(B, L)
    lobe = _ref_lobe_with(cfg, sd, tmp)
    lin = rb.reference_linear(cfg.hidden_size, 20)
    lin.load_state_dict(head)
    with torch.no_grad():
        feats = lobe(wav)
        logits = lin(feats)
    out = {
        "B": B, "L": L, "wav_seed": 1986, "weight_seed": 0, "affine_seed": 7, "head_seed": 0,
        "logits": logits.numpy(), "feats_head": feats[:, :, :taps_keep].numpy(),
        "feats_mean_abs": np.float64(feats.abs().mean().item()),
    }
    if store_weights:
        out["wav"] = wav.numpy()
        for k, v in sd.items():
            out["sd/" + k] = v.numpy()
        for k, v in head.items():
            out["head/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "logits", tuple(logits.shape), "absmax", float(logits.abs().max()))


def gold_hubert(tmp):
    """HuBERT through the same reference lobe (source path contains "hubert" -> HubertModel): base = group norm,
    post-LN layers, NO LayerNorm in the feature projection; large = wav2vec2-large's graph."""
    gold_w2v2("hubert_base_1s", W2V2Config.hubert_base(), B=2, L=16000, tmp=tmp, store_weights=False)
    gold_w2v2("hubert_large_1s", W2V2Config.hubert_large(), B=1, L=16000, tmp=tmp, store_weights=False)
    cfg = W2V2Config.hubert_base()
    cfg.conv_pos_batch_norm = True
    gold_w2v2("hubert_base_posbn_1s", cfg, B=2, L=16000, tmp=tmp, store_weights=False)


def gold_wavlm(tmp):
    """WavLM through the reference lobe (source path contains "wavlm" -> WavLMModel): gated relative position bias."""
    gold_w2v2("wavlm_base_1s", W2V2Config.wavlm_base(), B=2, L=16000, tmp=tmp, store_weights=False)
    gold_w2v2("wavlm_large_1s", W2V2Config.wavlm_large(), B=1, L=16000, tmp=tmp, store_weights=False)
    # 249 frames: relative distances beyond max_exact = 80 reach the logarithmic buckets
    gold_w2v2("wavlm_base_5s", W2V2Config.wavlm_base(), B=1, L=80000, tmp=tmp, store_weights=False)


def gold_data2vec(tmp):
    """data2vec-audio through the reference lobe (source path contains "data2vec" -> Data2VecAudioModel)."""
    gold_w2v2("data2vec_base_1s", W2V2Config.data2vec_base(), B=2, L=16000, tmp=tmp, store_weights=False)


def gold_fusion(name, D, d_ffn, nhead, B, Ta, Tv, store_weights):
    sd = random_fusion_weights(D, d_ffn, seed=3)
    g = torch.Generator().manual_seed(11)
    a = torch.randn(B, Ta, D, generator=g)
    v = torch.randn(B, Tv, D, generator=g)
    ref = rb.reference_fusion(alpha=0.5, nhead=nhead, d_ffn=d_ffn, d_model=D)
    full = dict(ref.state_dict())
    full.update(sd)  # keeps the reference's own `pe` buffer
    ref.load_state_dict(full, strict=True)
    with torch.no_grad():
        out = ref(a, v)
    rec = {"B": B, "Ta": Ta, "Tv": Tv, "D": D, "d_ffn": d_ffn, "nhead": nhead, "w_seed": 3, "x_seed": 11,
           "out": out.numpy(), "pe_head": ref.state_dict()["fusion.positional_encoding.pe"][0, :64, :16].numpy()}
    if store_weights:
        rec["a"], rec["v"] = a.numpy(), v.numpy()
        for k, t in sd.items():
            rec["sd/" + k] = t.numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(name, tuple(out.shape))


def gold_frame2note():
    f2n = rb.reference_frame2note()
    rec = {}
    cases = [(2, 0), (3, 1), (8, 2), (50, 3), (249, 4), (499, 5), (1500, 6), (14970, 7), (400, 8), (400, 9),
             (777, 10), (64, 11)]
    for n, seed in cases:
        lo_on, lo_off, octv, pc = random_frames(n, seed)
        p_on = torch.sigmoid(torch.from_numpy(lo_on))
        p_off = torch.sigmoid(torch.from_numpy(lo_off))
        # exactly the tuple the recipe builds (train_audio_ssl.py:95-100): 0-d fp32 tensors + python ints
        frame_info = [(p_on[i], p_off[i], int(octv[i]), int(pc[i])) for i in range(n)]
        for thr_name, (on_t, off_t) in {"a": (0.4, 0.5), "b": (0.7, 0.3)}.items():
            notes = np.array(f2n(frame_info, on_t, off_t, 1 / 49.8), dtype=np.float64).reshape(-1, 3)
            key = f"n{n}_s{seed}_{thr_name}"
            rec[key + "/notes"] = notes
            rec[key + "/thr"] = np.array([on_t, off_t])
        rec[f"n{n}_s{seed}/p_on"] = p_on.numpy()
        rec[f"n{n}_s{seed}/p_off"] = p_off.numpy()
        rec[f"n{n}_s{seed}/oct"] = octv
        rec[f"n{n}_s{seed}/pc"] = pc
    np.savez_compressed(os.path.join(GOLD, "frame2note_cases.npz"), **rec)
    print("frame2note cases", len(cases))


def gold_avhubert_resnet(name, B, T, seed=5, col_step=1):
    """Reference ResEncoder (N20EMv2/video_only/resnet.py:133-171) on seeded weights (avhubert_oracle.random_weights,
    BN with non-trivial running statistics) and a seeded normalised-video-like input."""
    from .avhubert_oracle import AVHubertConfig, random_weights as av_weights

    sd = av_weights(AVHubertConfig(encoder_layers=0), seed=seed)
    pre = "model.feature_extractor_video.resnet."
    ref = rb.reference_res_encoder()
    own = ref.state_dict()
    for k in own:
        if k.endswith("num_batches_tracked"):
            continue
        own[k] = sd[pre + k]
    ref.load_state_dict(own, strict=True)
    g = torch.Generator().manual_seed(seed + 100)
    video = torch.randn(B, 1, T, 88, 88, generator=g)
    with torch.no_grad():
        out = ref(video)  # (B, 512, T)
    extra = {} if col_step == 1 else {"col_step": col_step, "frame_sum": out.double().sum(1).numpy(),
                                      "frame_sumsq": (out.double() ** 2).sum(1).numpy()}
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), B=B, T=T, weight_seed=seed, video_seed=seed + 100,
                        out=out[:, ::col_step].numpy(), **extra)
    print(name, tuple(out.shape), "absmax", float(out.abs().max()))


def bench_wav(B, L, seed=1986):
    """Synthetic clips of different loudness (gain 0.25 .. 1.75 by clip index), so that the whole-tensor norms of a batched
    call really couple the clips."""
    gains = 0.25 + 1.5 * (torch.arange(B) % 7).float() / 6.0
    return synth_wav(B, L, seed) * gains[:, None]


def gold_w2v2_batch64(name, tmp, B=64, L=160000, keep=(0, 31, 63)):
    """BASELINE config 2 at its own shape through the REAL reference lobe: one call on (64, 160000) fp32 (whole-tensor
    norms over the batch), logits of three clips kept."""
    cfg = W2V2Config.large()
    sd = perturb_norm_affines(random_weights(cfg, seed=0))
    head = random_head(cfg.hidden_size, 20, seed=0)
    wav = bench_wav(B, L)
    lobe = _ref_lobe_with(cfg, sd, tmp)
    lin = rb.reference_linear(cfg.hidden_size, 20)
    lin.load_state_dict(head)
    with torch.no_grad():
        feats = lobe(wav)
        logits = lin(feats)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), B=B, L=L, wav_seed=1986, weight_seed=0, affine_seed=7, head_seed=0,
                        clips=np.array(keep), logits=logits[list(keep)].numpy(),
                        feats_sq_mean=np.float64((feats.double() ** 2).mean().item()))
    print(name, "logits", tuple(logits.shape), "absmax", float(logits.abs().max()))


def gold_fusion_10s(name, B=2, Ta=499, Tv=500, D=1024, d_ffn=3072, nhead=8, row_step=7, col_step=4):
    """FusionRCA at the frame counts of a 10-s utterance (499 audio / 500 video frames); a strided sample of the output plus
    a checksum of every row."""
    sd = random_fusion_weights(D, d_ffn, seed=3)
    g = torch.Generator().manual_seed(11)
    a = torch.randn(B, Ta, D, generator=g)
    v = torch.randn(B, Tv, D, generator=g)
    ref = rb.reference_fusion(alpha=0.5, nhead=nhead, d_ffn=d_ffn, d_model=D)
    full = dict(ref.state_dict())
    full.update(sd)
    ref.load_state_dict(full, strict=True)
    with torch.no_grad():
        out = ref(a, v)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), B=B, Ta=Ta, Tv=Tv, D=D, d_ffn=d_ffn, nhead=nhead, w_seed=3, x_seed=11,
                        row_step=row_step, col_step=col_step, out_sample=out[:, ::row_step, ::col_step].numpy(),
                        row_sum=out.double().sum(-1).numpy(), row_sumsq=(out.double() ** 2).sum(-1).numpy())
    print(name, tuple(out.shape))


def gold_bench_shapes():
    assert rb.available(), "needs /root/reference"
    torch.set_num_threads(os.cpu_count() or 1)
    with tempfile.TemporaryDirectory() as tmp:
        gold_w2v2_batch64("w2v2_large_10s_b64", tmp)
    gold_fusion_10s("fusion_10s")
    gold_avhubert_resnet("avhubert_resnet_b1_t500", B=1, T=500, col_step=5)


def main():
    assert rb.available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    with tempfile.TemporaryDirectory() as tmp:
        gold_w2v2("w2v2_tiny_large", TINY_LARGE, B=2, L=4000, tmp=tmp, store_weights=True)
        gold_w2v2("w2v2_tiny_base", TINY_BASE, B=2, L=4000, tmp=tmp, store_weights=True)
        gold_w2v2("w2v2_large_1s", W2V2Config.large(), B=2, L=16000, tmp=tmp, store_weights=False)
        gold_w2v2("w2v2_base_1s", W2V2Config.base(), B=2, L=16000, tmp=tmp, store_weights=False)
        gold_w2v2("w2v2_large_5s", W2V2Config.large(), B=1, L=80000, tmp=tmp, store_weights=False)
        gold_hubert(tmp)
        gold_data2vec(tmp)
        gold_wavlm(tmp)
    gold_fusion("fusion_tiny", D=64, d_ffn=96, nhead=4, B=2, Ta=13, Tv=15, store_weights=True)
    gold_fusion("fusion_full", D=1024, d_ffn=3072, nhead=8, B=1, Ta=49, Tv=50, store_weights=False)
    gold_fusion("fusion_full_pad", D=1024, d_ffn=3072, nhead=8, B=2, Ta=49, Tv=45, store_weights=False)
    gold_frame2note()
    gold_avhubert_resnet("avhubert_resnet_b2_t6", B=2, T=6)


if __name__ == "__main__":
    import sys
    if "bench_shapes" in sys.argv[1:]:  # the fixtures at the benchmarked shapes only (round 2); the rest stay as committed
        gold_bench_shapes()
    else:
        main()
        gold_bench_shapes()

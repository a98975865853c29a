"""Parity AT THE BENCHMARKED SHAPES (BASELINE.json configs 2-5), not only at 1-s miniatures:

  config 2   B = 64 x 160 000 samples, wav2vec2-large: a 4.0 GB workspace (byte offsets past 2^31) and a conv0 output of
             1.05 G elements.  Checked against (i) the REAL reference's logits for clips {0, 31, 63} of ONE batched call
             (tests/golden/w2v2_large_10s_b64.npz, made by oracle/make_golden.py from /root/reference), (ii) the oracle run
             clip by clip (per-clip normalisation scope = the reference's batch-size-1 evaluation), (iii) 64 batch-1 GPU calls.
  config 3   2 ranks over NCCL (skipped with fewer than 2 GPUs): gathered logits == the 1-GPU result of each shard.
  config 4   FusionRCA at 499 / 500 frames vs the reference's own output; the video lobe at 2 x 500 frames vs the oracle
             (whose ResNet is pinned at 500 frames by tests/golden/avhubert_resnet_b1_t500.npz).
  all        bf16-vs-fp32 NOTE agreement on 24 synthetic clips, reported in the log (SURVEY.md 8c asks for a report).

Tolerances as everywhere (bf16 storage / fp32 accumulate vs the fp32 reference): logits rel-L2 <= 2e-2, max-abs <= 0.1."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(__file__))
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LOGIT_REL_L2, LOGIT_MAX_ABS = 2e-2, 0.1


def _cmp(got, ref, tag):
    got = torch.as_tensor(got).detach().float().cpu().double()
    ref = torch.as_tensor(ref).double()
    err = float((got - ref).abs().max())
    rel = float((got - ref).norm() / ref.norm())
    print(f"{tag}: max-abs {err:.4e} rel-L2 {rel:.4e}")
    assert torch.isfinite(got).all()
    return rel, err


@pytest.fixture(scope="module")
def large():
    from oracle import wav2vec2_oracle as wo
    from test_gpu_e2e import _build

    cfg = wo.W2V2Config.large()
    lobe, lin, sd, head = _build(cfg)     # seeded HF init + perturbed LayerNorm affines / biases (the goldens' weights)
    import svt_speechbrain_b200 as svt
    return cfg, lobe, lin, sd, head, svt.AMTTranscriber(lobe, lin)


def test_config2_batch64_x_10s_vs_reference_golden_and_oracle(large):
    from oracle import make_golden as mg
    from oracle import wav2vec2_oracle as wo

    cfg, lobe, lin, sd, head, tr = large
    g = np.load(os.path.join(GOLD, "w2v2_large_10s_b64.npz"))
    B, L, keep = int(g["B"]), int(g["L"]), [int(c) for c in g["clips"]]
    assert (B, L) == (64, 160000)
    wav = mg.bench_wav(B, L, seed=int(g["wav_seed"]))
    dev = wav.cuda()
    eng = tr._engine()
    assert eng.workspace(B, L).numel() > (1 << 31)        # byte offsets really leave the 32-bit signed range at this shape
    # (i) one batched call, whole-tensor norms over the batch: the reference's own logits for three clips
    whole = tr.logits(dev)
    assert whole.shape == (B, 499, 20)
    rel, err = _cmp(whole[keep], g["logits"], "config 2, B=64 whole-batch norm vs REFERENCE golden (clips 0, 31, 63)")
    assert rel <= LOGIT_REL_L2 and err <= LOGIT_MAX_ABS
    # ... and the module-by-module route of AMT.compute_forward (feats out, separate head) on the same batch
    feats = lobe(dev)
    assert feats.shape == (B, 499, 1024)
    assert abs(float((feats.double() ** 2).mean()) - float(g["feats_sq_mean"])) < 2e-2
    assert torch.allclose(lin(feats[keep]), whole[keep], atol=1e-4)
    del feats
    # (ii) per-clip normalisation scope at the same shape vs the oracle, clip by clip
    per_clip = tr.logits(dev, per_clip_norm=True)
    with torch.no_grad():
        ref = torch.cat([wo.amt_logits(cfg, sd, head, wav[c:c + 1]) for c in keep])
    rel, err = _cmp(per_clip[keep], ref, "config 2, B=64 per-clip norm vs oracle at batch 1 (clips 0, 31, 63)")
    assert rel <= LOGIT_REL_L2 and err <= LOGIT_MAX_ABS
    assert float((per_clip[keep].cpu() - whole[keep].cpu()).abs().max()) > 1e-2   # the two scopes really differ here
    # (iii) ... and vs 64 separate GPU calls of batch size 1 (the reference's evaluation loop)
    worst = 0.0
    for c in range(B):
        one = tr.logits(dev[c:c + 1])
        worst = max(worst, float((one[0] - per_clip[c]).abs().max()))
    print(f"config 2: batched per-clip-norm call vs 64 batch-1 calls, worst max-abs {worst:.3e}")
    assert worst < 2e-3
    # misaligned views: rows of an odd-length batch and a sliced song are 4-byte, not 16-byte aligned.  The
    # fp32 partial sums of the input statistics group differently with the alignment (relative 1e-7), which flips a
    # handful of bf16 roundings in conv0's output; 24 layers later that is a logit difference at the bf16-noise level,
    # so the misaligned result is held to the same tolerance against the oracle, not to bit equality with the aligned one.
    odd = dev[:3, : 16001].contiguous()
    a = tr.logits(odd, per_clip_norm=True)
    b = torch.cat([tr.logits(odd[i:i + 1]) for i in range(3)])
    assert float((a - b).abs().max()) < 3e-2
    shifted = dev.reshape(-1)[1: 1 + 16000].unsqueeze(0)
    assert shifted.data_ptr() % 16 != 0
    got_shifted, got_aligned = tr.logits(shifted), tr.logits(shifted.clone())
    assert float((got_shifted - got_aligned).abs().max()) < 3e-2
    with torch.no_grad():
        ref1 = wo.amt_logits(cfg, sd, head, shifted.cpu())
    rel, err = _cmp(got_shifted, ref1, "4-byte-aligned (not 16-byte-aligned) wav view vs oracle")
    assert rel <= LOGIT_REL_L2 and err <= LOGIT_MAX_ABS
    rel, err = _cmp(a, torch.cat([wo.amt_logits(cfg, sd, head, odd[i:i + 1].cpu()) for i in range(3)]).detach(),
                    "odd-length (16001) batch, per-clip norm vs oracle")
    assert rel <= LOGIT_REL_L2 and err <= LOGIT_MAX_ABS


def test_whole_batch_norm_8_x_10s_vs_exact_oracle(large):
    """B = 8 x 10 s with the call-wide statistics, against the oracle's exact clip-by-clip restatement of ONE batched
    reference call (oracle.amt_logits_of_clips, pinned on the CPU against amt_logits)."""
    from oracle import make_golden as mg
    from oracle import wav2vec2_oracle as wo

    cfg, lobe, lin, sd, head, tr = large
    wav = mg.bench_wav(8, 160000, seed=77)
    with torch.no_grad():
        ref = wo.amt_logits_of_clips(cfg, sd, head, wav, clips=range(8))
    got = tr.logits(wav.cuda())
    rel, err = _cmp(got, ref, "B=8 x 10 s whole-batch norm vs exact oracle (all clips)")
    assert rel <= LOGIT_REL_L2 and err <= LOGIT_MAX_ABS


def test_bf16_vs_fp32_note_agreement_report(large):
    """24 synthetic 5-s clips (the recipes' dur_threshold): notes decoded from the GPU's bf16-storage logits vs notes
    decoded from the fp32 oracle's logits.  Random-init logits sit close to the decision thresholds, so agreement is
    REPORTED (and bounded loosely), not asserted to be exact; the decoder itself is bit-exact on identical logits."""
    import svt_speechbrain_b200 as svt
    from oracle import make_golden as mg
    from oracle import wav2vec2_oracle as wo
    from oracle.frame2note_oracle import frame2note as f2n_oracle
    from svt_speechbrain_b200.metrics import TranscriptionMeters

    cfg, lobe, lin, sd, head, tr = large
    hp = svt.AMTHparams()
    n_clips, L = 24, 80000
    wav = mg.bench_wav(n_clips, L, seed=2024)
    got = tr.logits(wav.cuda(), per_clip_norm=True)
    same_clips = same_notes = total_ref = total_est = 0
    flips = 0
    meters = TranscriptionMeters()
    for c in range(n_clips):
        with torch.no_grad():
            ref_lg = wo.amt_logits(cfg, sd, head, wav[c:c + 1])[0]
        p_on, p_off, octv, pc = wo.frame_info_from_logits(ref_lg)
        fi = [(p_on[i], p_off[i], int(octv[i]), int(pc[i])) for i in range(len(p_on))]
        want = np.array(f2n_oracle(fi, hp.onset_threshold, hp.offset_threshold, 1 / hp.frame_rate), dtype=np.float64).reshape(-1, 3)
        est = svt.decode_logits(got[c], hp)
        g_on, g_off, g_oct, g_pc = svt.frame_info(got[c], hp)
        flips += int(((g_on >= 0.4) != (p_on.numpy() >= 0.4)).sum() + ((g_off >= 0.5) != (p_off.numpy() >= 0.5)).sum()
                     + (g_oct != octv.numpy()).sum() + (g_pc != pc.numpy()).sum())
        same_clips += int(est.shape == want.shape and np.array_equal(est, want))
        ws, es = {tuple(r) for r in want.tolist()}, {tuple(r) for r in est.tolist()}
        same_notes += len(ws & es)
        total_ref += len(ws)
        total_est += len(es)
        if len(want) and len(est):
            meters.update(want, est)
    s = meters.summary()
    print(f"NOTE AGREEMENT bf16 GPU vs fp32 oracle, {n_clips} x 5-s clips, wav2vec2-large random init: "
          f"{same_clips}/{n_clips} clips with identical note lists; {same_notes} identical notes of {total_ref} reference / "
          f"{total_est} estimated; {flips} of {n_clips * 249 * 4} per-frame decisions differ; "
          f"COnPOff F1 {s['COnPOff_f1']:.4f} COnP F1 {s['COnP_f1']:.4f} COn F1 {s['COn_f1']:.4f}")
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "note_agreement.txt"), "w") as f:
            f.write(f"clips_identical {same_clips}/{n_clips}\nnotes_identical {same_notes} ref {total_ref} est {total_est}\n"
                    f"frame_decisions_differing {flips}/{n_clips * 249 * 4}\n{s}\n")
    assert total_ref > 100
    assert same_notes >= 0.8 * total_ref and s["COn_f1"] > 0.9


def test_fusion_at_10s_frame_counts_vs_reference_golden():
    import svt_speechbrain_b200 as svt
    from oracle import make_golden as mg

    g = np.load(os.path.join(GOLD, "fusion_10s.npz"))
    D, B, Ta, Tv = int(g["D"]), int(g["B"]), int(g["Ta"]), int(g["Tv"])
    assert (Ta, Tv) == (499, 500)
    fus = svt.FusionRCA(alpha=0.5, nhead=int(g["nhead"]), d_ffn=int(g["d_ffn"]), d_model=D)
    full = dict(fus.state_dict())
    full.update(mg.random_fusion_weights(D, int(g["d_ffn"]), seed=int(g["w_seed"])))
    fus.load_state_dict(full, strict=True)
    gen = torch.Generator().manual_seed(int(g["x_seed"]))
    a = torch.randn(B, Ta, D, generator=gen)
    v = torch.randn(B, Tv, D, generator=gen)
    out = fus.cuda()(a.cuda(), v.cuda()).cpu()
    assert out.shape == (B, Ta, D)
    rs, cs = int(g["row_step"]), int(g["col_step"])
    rel, err = _cmp(out[:, ::rs, ::cs], g["out_sample"], "FusionRCA 499 / 500 frames vs REFERENCE golden (strided sample)")
    assert rel <= 2e-2 and err <= 0.15
    # every row is covered by its checksum: sum of squares of a row (two LayerNorm outputs added: ~ 2 D) within 2 %,
    # row sums within bf16 noise of D elements
    sq = (out.double() ** 2).sum(-1).numpy()
    assert np.abs(sq / g["row_sumsq"] - 1).max() < 2e-2
    assert np.abs(out.double().sum(-1).numpy() - g["row_sum"]).max() < 0.05 * np.sqrt(D) * 4


def test_video_lobe_2_x_500_frames_vs_oracle():
    """AV-HuBERT-large geometry at config 4's length (500 lip frames), two clips.  The oracle's ResNet front end equals
    the reference's ResEncoder at 500 frames (golden); its transformer body is the fairseq encoder restated, parity
    unpinned (fairseq cannot be imported), as DESIGN.md states."""
    from oracle import avhubert_oracle as av
    from test_gpu_video import _lobe

    d = np.load(os.path.join(GOLD, "avhubert_resnet_b1_t500.npz"))
    cfg0 = av.AVHubertConfig(encoder_layers=0)
    sd0 = av.random_weights(cfg0, seed=int(d["weight_seed"]))
    video1 = torch.randn(int(d["B"]), 1, int(d["T"]), 88, 88, generator=torch.Generator().manual_seed(int(d["video_seed"])))
    with torch.no_grad():
        res = av.res_encoder(sd0, video1, "model.feature_extractor_video.resnet.")          # (B, 512, T)
    cs = int(d["col_step"])
    assert float((res[:, ::cs] - torch.from_numpy(d["out"])).abs().max()) < 2e-4              # oracle == reference at T = 500
    assert np.abs(res.double().sum(1).numpy() - d["frame_sum"]).max() < 1e-2

    cfg = av.AVHubertConfig()
    sd = av.random_weights(cfg, seed=0)
    video = torch.randn(2, 1, 500, 88, 88, generator=torch.Generator().manual_seed(4))
    video[1] *= 0.5
    with torch.no_grad():
        ref = av.lobe_forward(cfg, sd, video)
    got = _lobe(cfg, sd)({"video": video.cuda(), "audio": None}).cpu()
    assert got.shape == ref.shape == (2, 500, 1024)
    rel, err = _cmp(got, ref, "video lobe, AV-HuBERT-large, 2 x 500 frames vs oracle")
    assert rel <= 2e-2 and err <= 0.15


# ------------------------------------------------------------------------------------------------ config 3 (NCCL, 2 ranks)
def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import make_golden as mg
        from oracle import wav2vec2_oracle as wo
        from svt_speechbrain_b200.engine import EncoderEngine, encoder_config_from_hf
        from svt_speechbrain_b200.parallel import LogitsGatherer, gather_logits, shard_range
        from transformers import Wav2Vec2Config

        cfg = wo.W2V2Config.large()
        sd = mg.perturb_norm_affines(wo.random_weights(cfg, seed=0), seed=7)
        head = wo.random_head(cfg.hidden_size, 20, seed=0)
        eng = EncoderEngine(encoder_config_from_hf(Wav2Vec2Config(**cfg.hf_kwargs()), True, True), dev)
        eng.load(sd, head["w.weight"], head["w.bias"])
        n_total, L = 8, 160000
        wav = mg.bench_wav(n_total, L, seed=5)
        a, b = shard_range(n_total, rank, world)
        _, local = eng.forward(wav[a:b].to(dev), want_feats=False, want_logits=True)
        gathered = gather_logits(local, n_total)
        # the serving-loop gatherer: three steps through two slots, results consumed one step behind
        g = LogitsGatherer(tuple(local.shape), depth=2, device=dev)
        outs = []
        for k in range(3):
            eng.forward(wav[a:b].to(dev) * (1.0 + k), want_feats=False, want_logits=True, logits_out=g.local(k))
            g.submit(k)
            if k >= 1:
                outs.append(g.result(k - 1).clone())
        outs.append(g.result(2).clone())
        g.finish()
        ok = True
        msg = ""
        if rank == 0:
            # the same shards on ONE GPU: each rank's block of the gathered tensor must be that shard's 1-GPU result
            for r in range(world):
                ra, rb = shard_range(n_total, r, world)
                _, one = eng.forward(wav[ra:rb].to(dev), want_feats=False, want_logits=True)
                d = float((gathered[ra:rb] - one).abs().max())
                msg += f"shard {r}: max |gathered - 1-GPU| {d:.3e}; "
                ok = ok and d < 2e-3
                for k in range(3):
                    _, onek = eng.forward(wav[ra:rb].to(dev) * (1.0 + k), want_feats=False, want_logits=True)
                    ok = ok and float((outs[k][ra:rb] - onek).abs().max()) < 2e-3
            with torch.no_grad():
                ref = wo.amt_logits_of_clips(cfg, sd, head, wav[:b], clips=[0])
            rel = float((gathered[0:1].cpu() - ref).norm() / ref.norm())
            msg += f"clip 0 vs oracle (shard-wide statistics) rel-L2 {rel:.3e}"
            ok = ok and rel < 2e-2
        q.put((rank, ok, msg))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_config3_two_rank_nccl_gather_equals_one_gpu():
    import socket
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, ok, msg in res:
        print(f"rank {rank}: {msg}")
        assert ok, (rank, msg)

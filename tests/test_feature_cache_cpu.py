"""On-disk formats around the path: path rules, the stage-C utterance slicing / alignment rule
(N20EMv2/audio_visual/train_rca_av.py:398-441) and the save / load round trip.  CPU only."""
import os

import numpy as np
import torch

from svt_speechbrain_b200 import feature_cache as fc


def test_paths_follow_the_recipes():
    assert fc.audio_feats_path("/d/song1") == "/d/song1/noise_data/clean_feats.pt"
    assert fc.audio_feats_path("/d/song1", True, "musan", -5) == "/d/song1/noise_data/musan/SNR_-5dB_feats.pt"
    assert fc.video_feats_path("/d/song1") == "/d/song1/noise_data/video_feats.pt"
    assert fc.av_pred_path("/d/song1/noise_data") == "/d/song1/noise_data/clean_av_pred.npy"
    assert fc.av_pred_path("/d/song1/noise_data", True, "musan", 10) == "/d/song1/noise_data/musan/SNR_10dB_av_pred.npy"


def test_slice_av_utterance_rule():
    # a 23.4-s song: 1165 audio frames at 49.8 fps, 1170 video frames at 50 fps, 5 utterances of 5 s
    a = torch.arange(1165 * 4, dtype=torch.float32).reshape(1165, 4)
    v = torch.arange(1170 * 4, dtype=torch.float32).reshape(1170, 4) + 0.5
    total = 0
    for u in range(1, 6):
        s1, s2 = fc.slice_av_utterance(a, v, u, 5, 5, 49.8, 50, feat_dim=4)
        lo1, lo2 = round((u - 1) * 49.8 * 5), round((u - 1) * 50 * 5)
        hi1 = 1165 if u == 5 else round(u * 49.8 * 5)
        assert torch.equal(s1, a[lo1:hi1])
        assert s2.shape[0] == s1.shape[0]                       # video aligned to the audio frame count
        n = min(s1.shape[0], (1170 if u == 5 else round(u * 50 * 5)) - lo2)
        assert torch.equal(s2[:n], v[lo2:lo2 + n])
        total += s1.shape[0]
    assert total == 1165                                          # the utterances tile the song's audio frames
    # video shorter than audio -> zero padding (train_rca_av.py:438-439)
    s1, s2 = fc.slice_av_utterance(a, v[:1100], 5, 5, 5, 49.8, 50, feat_dim=4)
    assert s2.shape == s1.shape and float(s2[-1].abs().sum()) == 0.0 and float(s2[0].abs().sum()) > 0.0


def test_round_trip(tmp_path):
    feats = torch.randn(37, 1024)
    p = fc.save_song_features(feats, fc.audio_feats_path(str(tmp_path / "songA")))
    assert os.path.exists(p) and torch.equal(torch.load(p), feats)
    pv = fc.save_song_features(torch.randn(40, 1024), fc.video_feats_path(str(tmp_path / "songA")))
    s1, s2 = fc.load_av_utterance(p, pv, 1, 1)
    assert s1.shape == (37, 1024) and s2.shape == (37, 1024)
    notes = np.array([[0.1, 0.5, 60.0], [0.7, 1.2, 62.0]])
    pn = fc.save_notes(notes, fc.av_pred_path(str(tmp_path / "songA" / "noise_data")))
    assert np.array_equal(np.load(pn), notes)


def test_manifest_rows():
    rows = fc.iter_song_utterances(16000 * 23 + 6400)            # 23.4 s -> round(23.4 / 5) = 5 utterances
    assert [r[0] for r in rows] == [1, 2, 3, 4, 5] and all(r[1] == 5 for r in rows)
    assert rows[0][2:] == (0, 80000) and rows[-1][3] == 16000 * 23 + 6400

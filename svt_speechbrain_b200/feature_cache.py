"""On-disk formats either side of the hot path, so the accelerated path slots into the
authors' three-stage audio-visual workflow unchanged:

  stage A  N20EMv2/audio_only/extract_ssl_feats.py:102-116   per-song audio features  (frames, 1024) fp32, torch.save
             <song folder>/noise_data/clean_feats.pt   or   <song folder>/noise_data/<noise_type>/SNR_<snr>dB_feats.pt
  stage B  N20EMv2/video_only/extract_ssl_feats.py:102-111   per-song video features  (frames, 1024) fp32
             <song folder>/noise_data/video_feats.pt
  stage C  N20EMv2/audio_visual/train_rca_av.py:398-441      loads both, cuts utterance `utter_id` of `utter_num` by FRAME
             index (49.8 / 50 frames per second x dur_threshold), pads / trims the video frames to the audio frames
           N20EMv2/audio_visual/train_rca_av.py:113-123      note dumps  clean_av_pred.npy / SNR_<snr>dB_av_pred.npy

Nothing here computes: the features come from the lobes (HuggingFaceWav2Vec2 / FairseqAVHubertPretrain drop-ins), the
notes from AMTTranscriber.decode.  File names, tensor shapes, dtypes and slicing arithmetic follow the reference lines
cited above.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .amt import AMTHparams, split_song


# ----------------------------------------------------------------------------------------------- paths
def audio_feats_path(song_folder: str, add_noise: bool = False, noise_type: Optional[str] = None,
                     snr_db: Optional[int] = None) -> str:
    """audio_only/extract_ssl_feats.py:108-112."""
    if add_noise:
        if noise_type is None or snr_db is None:
            raise ValueError("add_noise=True needs noise_type and snr_db")
        return os.path.join(song_folder, "noise_data", noise_type, f"SNR_{snr_db}dB_feats.pt")
    return os.path.join(song_folder, "noise_data", "clean_feats.pt")


def video_feats_path(song_folder: str) -> str:
    """video_only/extract_ssl_feats.py:109."""
    return os.path.join(song_folder, "noise_data", "video_feats.pt")


def av_pred_path(noise_data_folder: str, add_noise: bool = False, noise_type: Optional[str] = None,
                 snr_db: Optional[int] = None) -> str:
    """audio_visual/train_rca_av.py:118-122 (the folder is the one that holds clean_feats.pt)."""
    if add_noise:
        if noise_type is None or snr_db is None:
            raise ValueError("add_noise=True needs noise_type and snr_db")
        return os.path.join(noise_data_folder, noise_type, f"SNR_{snr_db}dB_av_pred.npy")
    return os.path.join(noise_data_folder, "clean_av_pred.npy")


# ----------------------------------------------------------------------------------------------- stage A / B writers
@torch.no_grad()
def extract_song_features(lobe, wav: torch.Tensor, hparams: Optional[AMTHparams] = None, dur: Optional[float] = None,
                          batch_clips: int = 64, device="cuda") -> torch.Tensor:
    """Stage A for one song: utterances cut by the reference rule, each normalised on its own (the reference extracts
    with batch size 1), features concatenated along frames -> (frames, D) fp32 on the CPU, exactly the tensor
    extract_ssl_feats.py:107 saves.  Equal-length utterances run as one batched call with per-clip statistics."""
    hp = hparams or AMTHparams()
    dev = torch.device(device)
    wav = wav.to(dev, torch.float32).reshape(-1)
    spans = split_song(wav.numel(), hp, dur)
    eng = lobe.engine(dev)
    pieces: List[torch.Tensor] = []
    i = 0
    while i < len(spans):
        j = i
        L = spans[i][1] - spans[i][0]
        while j < len(spans) and j - i < batch_clips and spans[j][1] - spans[j][0] == L:
            j += 1
        clips = torch.stack([wav[a:b] for a, b in spans[i:j]])
        if j - i > 1 and L % 4 == 0:
            eng.set_norm_per_clip(True)
            try:
                feats, _ = eng.forward(clips, want_feats=True, want_logits=False)
            finally:
                eng.set_norm_per_clip(False)
            pieces.extend(feats[k] for k in range(j - i))
        else:
            for k in range(j - i):
                feats, _ = eng.forward(clips[k:k + 1], want_feats=True, want_logits=False)
                pieces.append(feats[0])
        i = j
    return torch.cat(pieces, dim=0).float().cpu()


def save_song_features(feats: torch.Tensor, path: str) -> str:
    """torch.save of a (frames, D) fp32 CPU tensor, creating the noise_data folders like the recipes' prepare step."""
    if feats.dim() != 2:
        raise ValueError(f"expected (frames, D), got {tuple(feats.shape)}")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(feats.detach().float().cpu().contiguous(), path)
    return path


def save_notes(notes: np.ndarray, path: str) -> str:
    """np.save of the (n_notes, 3) [onset s, offset s, MIDI pitch] array (train_rca_av.py:123)."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.save(path, np.asarray(notes))
    return path


# ----------------------------------------------------------------------------------------------- stage C reader
def load_av_utterance(audio_path: str, video_path: str, utter_id: int, utter_num: int, dur_threshold: float = 5,
                      audio_sample_rate: float = 49.8, video_sample_rate: float = 50,
                      feat_dim: int = 1024) -> Tuple[torch.Tensor, torch.Tensor]:
    """`audio_visual_pipeline` of train_rca_av.py:398-441: utterance `utter_id` (1-based) of a song's cached features.
    Audio frames [round((id-1) * 49.8 * dur), round(id * 49.8 * dur)), video likewise at 50 fps, the last utterance
    takes the rest; the video frames are trimmed or zero-padded to the audio frame count."""
    sig1 = torch.load(audio_path)
    sig2 = torch.load(video_path)
    return slice_av_utterance(sig1, sig2, utter_id, utter_num, dur_threshold, audio_sample_rate, video_sample_rate, feat_dim)


def slice_av_utterance(sig1: torch.Tensor, sig2: torch.Tensor, utter_id: int, utter_num: int, dur_threshold: float = 5,
                       audio_sample_rate: float = 49.8, video_sample_rate: float = 50,
                       feat_dim: int = 1024) -> Tuple[torch.Tensor, torch.Tensor]:
    utter_id, utter_num = int(utter_id), int(utter_num)
    if utter_id == utter_num:
        sig1 = sig1[round((utter_id - 1) * audio_sample_rate * dur_threshold):]
        sig2 = sig2[round((utter_id - 1) * video_sample_rate * dur_threshold):]
    else:
        sig1 = sig1[round((utter_id - 1) * audio_sample_rate * dur_threshold):round(utter_id * audio_sample_rate * dur_threshold)]
        sig2 = sig2[round((utter_id - 1) * video_sample_rate * dur_threshold):round(utter_id * video_sample_rate * dur_threshold)]
    frame1, frame2 = sig1.shape[0], sig2.shape[0]
    if frame1 < frame2:
        sig2 = sig2[:frame1]
    elif frame1 > frame2:
        sig2 = torch.cat([sig2, torch.zeros(frame1 - frame2, feat_dim, dtype=sig2.dtype)], dim=0)
    return sig1, sig2


@torch.no_grad()
def transcribe_from_cache(fusion, head, hparams: Optional[AMTHparams], audio_path: str, video_path: str, utter_num: int,
                          dur_threshold: float = 5, device="cuda", batch_utterances: int = 32) -> np.ndarray:
    """Stage C evaluation for one song from the cached features: every utterance through FusionRCA + head (the
    reference does it one utterance per step, train_rca_av.py:28-51,84-112), frames concatenated in utterance order,
    one decode.  Utterances of equal frame counts are batched (FusionRCA has no cross-clip coupling)."""
    from .amt import decode_logits

    dev = torch.device(device)
    sig1 = torch.load(audio_path)
    sig2 = torch.load(video_path)
    utts = [slice_av_utterance(sig1, sig2, u, utter_num, dur_threshold, feat_dim=sig1.shape[1]) for u in range(1, utter_num + 1)]
    out: List[Optional[torch.Tensor]] = [None] * utter_num
    order = sorted(range(utter_num), key=lambda u: (utts[u][0].shape[0], u))
    i = 0
    while i < utter_num:
        j = i
        n = utts[order[i]][0].shape[0]
        while j < utter_num and j - i < batch_utterances and utts[order[j]][0].shape[0] == n:
            j += 1
        a = torch.stack([utts[u][0] for u in order[i:j]]).to(dev, torch.float32)
        v = torch.stack([utts[u][1] for u in order[i:j]]).to(dev, torch.float32)
        lg = head(fusion(a, v))
        for k, u in enumerate(order[i:j]):
            out[u] = lg[k]
        i = j
    return decode_logits(torch.cat(out, dim=0), hparams or AMTHparams())


def iter_song_utterances(n_samples: int, hparams: Optional[AMTHparams] = None,
                         dur: Optional[float] = None) -> Sequence[Tuple[int, int, int, int]]:
    """(utter_id, utter_num, start sample, stop sample) rows of one song, the columns prepare_n20emv2.py writes into the
    recipes' CSV manifests (audio_visual/prepare_n20emv2.py:22) next to the wav / video paths."""
    spans = split_song(n_samples, hparams or AMTHparams(), dur)
    return [(i + 1, len(spans), a, b) for i, (a, b) in enumerate(spans)]

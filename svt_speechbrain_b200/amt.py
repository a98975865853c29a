"""AMT inference pipeline around the reference-shaped modules: wav -> lobe -> head -> per-frame
sigmoid / argmax -> frame2note, i.e. `AMT.compute_forward` + the evaluation half of `AMT.compute_objectives`
(MIR_ST500/train_audio_ssl.py:28-48, 85-108), with the reference's one-utterance-at-a-time loop replaced by
batched clips and with song chunking as in the recipes' dataio (prepare_benchmarks.py:117-130,
train_audio_ssl.py:373-390).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import check, current_stream_ptr, lib, ptr
from .utils import decode_arrays


@dataclass
class AMTHparams:
    """Scalars of MIR_ST500/hparams/train_audio_ssl.yaml that the inference path reads."""

    sample_rate: int = 16000
    frame_rate: float = 49.8
    dur_threshold: float = 5.0
    onset_threshold: float = 0.4
    offset_threshold: float = 0.5
    pitch_octave_num: int = 4
    pitch_class_num: int = 12

    @property
    def n_out(self) -> int:
        return 2 + (self.pitch_octave_num + 1) + (self.pitch_class_num + 1)


def split_song(n_samples: int, hp: AMTHparams, dur: Optional[float] = None) -> List[Tuple[int, int]]:
    """Sample ranges of a song's utterances, reference rule: utter_num = round(duration / dur); utterance i
    (1-based) covers [round((i-1)*sr*dur), round(i*sr*dur)), the last one takes the remainder."""
    dur = hp.dur_threshold if dur is None else dur
    duration = n_samples / hp.sample_rate
    utter_num = max(int(round(duration / dur)), 1)
    out = []
    for i in range(1, utter_num + 1):
        a = round((i - 1) * hp.sample_rate * dur)
        b = n_samples if i == utter_num else round(i * hp.sample_rate * dur)
        out.append((a, b))
    return out


def frame_info(logits: torch.Tensor, hp: "AMTHparams"):
    """(n_frames, 20) CUDA logits -> host arrays (p_on f32, p_off f32, octave i32, pitch_class i32)
    (train_audio_ssl.py:93-100).  argmax runs on the device (first max wins); the two sigmoids are taken on the HOST
    with torch so the probabilities are bit-identical to the CPU reference that frame2note's `==` / `>=` tests see."""
    return _unpack_frames(_pack_frames(logits, hp).cpu())  # ONE device->host transfer


def decode_logits(logits: torch.Tensor, hp: "AMTHparams") -> np.ndarray:
    """(n_frames, 20) logits of ONE song (utterances already concatenated in order) -> (n_notes, 3) float64
    [onset s, offset s, MIDI pitch] (frame2note, utils.py:82-149)."""
    p_on, p_off, octv, pc = frame_info(logits, hp)
    return decode_arrays(p_on, p_off, octv, pc, hp.onset_threshold, hp.offset_threshold, 1.0 / hp.frame_rate)


FRAME_HOP = 320      # samples between output frames of the conv stack (product of the strides 5 * 2^6)
FRAME_FIELD = 400    # receptive field of one output frame


def split_song_overlapped(n_samples: int, sample_rate: int, dur: float, overlap: float) -> List[Tuple[int, int]]:
    """Sliding windows of `dur` seconds every `dur - overlap` seconds (BASELINE config 5: long-form songs cut into
    overlapping windows).  Window starts are multiples of the frame hop so the windows' frames sit on one global frame
    grid; the last window ends at the end of the song."""
    if not 0 <= overlap < dur:
        raise ValueError("need 0 <= overlap < dur")
    win = int(round(dur * sample_rate))
    hop = max(FRAME_HOP, int(round((dur - overlap) * sample_rate)) // FRAME_HOP * FRAME_HOP)
    out, a = [], 0
    while True:
        b = min(a + win, n_samples)
        if out and b - a < FRAME_FIELD:  # a tail too short for one frame: the previous window takes it (as the reference's
            out[-1] = (out[-1][0], n_samples)  # last utterance takes the remainder, prepare_benchmarks.py:124-127)
            return out
        out.append((a, b))
        if b >= n_samples:
            return out
        a += hop


def stitch_plan(windows: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """Per window, the [lo, hi) range of its LOCAL frames kept in the stitched song: consecutive windows meet in the
    middle of their overlap (first / last window keep their outer ends), so every global frame comes from the window in
    which it is furthest from an edge.  Frame f of a window starting at sample a is global frame a / hop + f."""
    n = [max((b - a - FRAME_FIELD) // FRAME_HOP + 1, 0) for a, b in windows]
    g0 = [a // FRAME_HOP for a, _ in windows]
    plan = []
    for i in range(len(windows)):
        lo_g = g0[i] if i == 0 else (g0[i] + g0[i - 1] + n[i - 1] + 1) // 2
        hi_g = g0[i] + n[i] if i == len(windows) - 1 else (g0[i + 1] + g0[i] + n[i] + 1) // 2
        lo_g = max(lo_g, g0[i])
        hi_g = min(max(hi_g, lo_g), g0[i] + n[i])
        plan.append((lo_g - g0[i], hi_g - g0[i]))
    return plan


class AMTTranscriber:
    """lobe: svt_speechbrain_b200.HuggingFaceWav2Vec2; head: svt_speechbrain_b200.Linear (n_neurons = 20)."""

    def __init__(self, lobe, head, hparams: Optional[AMTHparams] = None, device="cuda"):
        self.lobe, self.head = lobe, head
        self.hp = hparams or AMTHparams()
        self.device = torch.device(device)
        if self.head.w.weight.shape[0] != self.hp.n_out:
            raise ValueError(f"head has {self.head.w.weight.shape[0]} outputs, hparams imply {self.hp.n_out}")
        self._head_key = None

    def _engine(self):
        eng = self.lobe.engine(self.device)
        key = (id(eng), self.head.w.weight.data_ptr(), self.head.w.weight._version,
               None if self.head.w.bias is None else self.head.w.bias._version)
        if self._head_key != key:
            eng.set_head(self.head.w.weight, self.head.w.bias)
            self._head_key = key
        return eng

    @torch.no_grad()
    def logits(self, wav: torch.Tensor, per_clip_norm: bool = False) -> torch.Tensor:
        """wav (B, L) CUDA fp32 -> frame logits (B, T, 20) CUDA fp32 (encoder + output norm + head fused).
        per_clip_norm: the two whole-tensor layer norms use one mean / variance per clip, i.e. the result of B separate
        reference calls of batch size 1 (the reference's evaluation loop) computed in one batched call."""
        eng = self._engine()
        eng.set_norm_per_clip(per_clip_norm)
        try:
            _, lg = eng.forward(wav, want_feats=False, want_logits=True)
        finally:
            eng.set_norm_per_clip(False)
        return lg

    @torch.no_grad()
    def _clip_logits(self, clips: Sequence[torch.Tensor], batch_clips: int, per_clip_norm: bool) -> List[torch.Tensor]:
        """Frame logits (T_i, 20) of 1-D clips, in order.  Runs of consecutive equal-length clips go through the encoder
        as one batch of at most `batch_clips`; with per_clip_norm every clip is normalised on its own (the reference's
        batch-size-1 evaluation).  Clips shorter than the conv stack's receptive field have no frame: they yield an empty
        (0, n_out) tensor instead of reaching the encoder (so a ragged tail never raises on one rank only)."""
        out: List[torch.Tensor] = []
        i = 0
        while i < len(clips):
            j = i
            L = clips[i].numel()
            while j < len(clips) and j - i < batch_clips and clips[j].numel() == L:
                j += 1
            if L < FRAME_FIELD:
                out.extend(torch.empty(0, self.hp.n_out, dtype=torch.float32, device=self.device) for _ in range(j - i))
                i = j
                continue
            batch = torch.stack(list(clips[i:j]))
            # a single clip is its own normalisation scope either way
            lg = self.logits(batch, per_clip_norm=per_clip_norm and j - i > 1)
            out.extend(lg[k] for k in range(lg.shape[0]))
            i = j
        return out

    def frame_info(self, logits: torch.Tensor):
        return frame_info(logits, self.hp)

    def decode(self, logits: torch.Tensor) -> np.ndarray:
        """(n_frames, 20) logits of ONE song (utterances already concatenated in order) -> (n_notes, 3) float64."""
        return decode_logits(logits, self.hp)

    @torch.no_grad()
    def transcribe_song(self, wav: torch.Tensor, dur: Optional[float] = None, batch_clips: int = 64,
                        per_clip_norm: bool = True) -> np.ndarray:
        """wav: 1-D waveform of a whole song.  Utterances are cut by the reference rule, run through the
        encoder, concatenated in order and decoded once.  Equal-length utterances are batched on the device.
        per_clip_norm=True reproduces the reference evaluation loop (batch size 1: the whole-tensor norms see one
        utterance at a time) with the per-clip normalisation scope of the encoder, still in batched calls."""
        wav = wav.to(self.device, torch.float32).reshape(-1)
        spans = split_song(wav.numel(), self.hp, dur)
        pieces = self._clip_logits([wav[a:b] for a, b in spans], batch_clips, per_clip_norm)
        return self.decode(torch.cat(pieces, dim=0))

    @torch.no_grad()
    def transcribe_long(self, wav: torch.Tensor, dur: float = 10.0, overlap: float = 1.0, batch_clips: int = 64,
                        per_clip_norm: bool = True, group=None, sharded: bool = True) -> np.ndarray:
        """Long-form song (BASELINE config 5): overlapping `dur`-second windows, sharded over the ranks of `group` when
        torch.distributed is initialised (contiguous blocks of windows per rank, no collective inside the forward), frame
        logits gathered in window order, overlaps resolved by `stitch_plan`, one decode of the stitched frames (identical
        on every rank).  overlap = 0 on a song whose length is a multiple of `dur` is `transcribe_song`."""
        return self.decode(self.long_form_logits(wav, dur, overlap, batch_clips, per_clip_norm, group, sharded))

    @torch.no_grad()
    def long_form_logits(self, wav: torch.Tensor, dur: float = 10.0, overlap: float = 1.0, batch_clips: int = 64,
                         per_clip_norm: bool = True, group=None, sharded: bool = True) -> torch.Tensor:
        """The stitched (n_frames, 20) frame logits behind `transcribe_long` (same on every rank).  sharded=False runs every
        window on this rank even when torch.distributed is initialised (the 1-GPU result, for comparison)."""
        import torch.distributed as dist
        from .parallel import gather_ragged, shard_range

        wav = wav.to(self.device, torch.float32).reshape(-1)
        windows = split_song_overlapped(wav.numel(), self.hp.sample_rate, dur, overlap)
        use_dist = sharded and dist.is_initialized()
        world = dist.get_world_size(group) if use_dist else 1
        rank = dist.get_rank(group) if use_dist else 0
        lo, hi = shard_range(len(windows), rank, world)
        local = self._clip_logits([wav[a:b] for a, b in windows[lo:hi]], batch_clips, per_clip_norm)
        pieces = gather_ragged(local, len(windows), group) if world > 1 else local
        plan = stitch_plan(windows)
        return torch.cat([p[a:b] for p, (a, b) in zip(pieces, plan)], dim=0)

    @torch.no_grad()
    def transcribe_songs(self, wavs: Sequence[torch.Tensor], dur: Optional[float] = None, batch_clips: int = 64,
                         per_clip_norm: bool = True) -> List[np.ndarray]:
        """Evaluation driver replacing the reference's one-utterance-at-a-time loop (speechbrain/core.py:1221-1226,
        train_audio_ssl.py:85-141): utterances of ALL songs are pooled, equal-length ones run in batches of `batch_clips`
        (per-clip normalisation = reference semantics), frames are put back in (song, utterance) order (:88,100) and every
        song is decoded once.

        Nothing on the device path synchronises: every batch's forward, per-frame argmax and device->host copy (into pinned
        memory, followed by an event) are queued up front; the host then walks the batches in order, waits for a batch's
        event and runs sigmoid + frame2note for the songs that batch completes -- while the GPU is already running the
        following batches."""
        hp = self.hp
        songs = [w.to(self.device, torch.float32, non_blocking=True).reshape(-1) for w in wavs]
        n_utt, plan = plan_song_batches([w.numel() for w in songs], hp, dur, batch_clips)
        stream = torch.cuda.current_stream(self.device)
        batches = []  # (pinned (frames, 4) tensor, event, [(song, utt, first row, rows)])
        for jobs in plan:
            lgs = self._clip_logits([songs[si][a:b] for si, _, a, b in jobs], batch_clips, per_clip_norm)
            rows, index = 0, []
            for (si, ui, _, _), lg in zip(jobs, lgs):
                index.append((si, ui, rows, int(lg.shape[0])))
                rows += int(lg.shape[0])
            packed = _pack_frames(torch.cat(lgs, dim=0) if len(lgs) > 1 else lgs[0], hp)
            host = torch.empty(packed.shape, dtype=torch.float32, pin_memory=True)
            host.copy_(packed, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            batches.append((host, ev, index))
        pieces = {}
        done = [0] * len(songs)
        results: List[Optional[np.ndarray]] = [None] * len(songs)
        for host, ev, index in batches:
            ev.synchronize()
            p_on, p_off, octv, pc = _unpack_frames(host)
            for si, ui, r0, n in index:
                pieces[(si, ui)] = (r0, n, p_on, p_off, octv, pc)
                done[si] += 1
                if done[si] == n_utt[si]:
                    parts = [pieces.pop((si, u)) for u in range(n_utt[si])]
                    cols = []
                    for c in range(4):
                        segs = [arrs[c][r:r + m] for r, m, *arrs in parts]
                        cols.append(segs[0] if len(segs) == 1 else np.concatenate(segs))
                    results[si] = decode_arrays(cols[0], cols[1], cols[2], cols[3], hp.onset_threshold, hp.offset_threshold,
                                                1.0 / hp.frame_rate)
        for si in range(len(songs)):
            if results[si] is None:  # a song without a single frame
                results[si] = np.zeros((0, 3), dtype=np.float64)
        return results


def _pack_frames(logits: torch.Tensor, hp: "AMTHparams") -> torch.Tensor:
    """(n_frames, 20) CUDA logits -> (n_frames, 4) fp32 on the device: onset / offset logits and the bits of the int32
    octave / pitch-class argmax (svt_frame_postproc, first max wins) -- everything the decoder needs, in one tensor."""
    lg = logits.reshape(-1, logits.shape[-1]).contiguous()
    n = lg.shape[0]
    octv = torch.empty(n, dtype=torch.int32, device=lg.device)
    pc = torch.empty(n, dtype=torch.int32, device=lg.device)
    with torch.cuda.device(lg.device):
        check(lib().svt_frame_postproc(ptr(lg), n, lg.shape[1], 2, hp.pitch_octave_num + 1,
                                       2 + hp.pitch_octave_num + 1, hp.pitch_class_num + 1, ptr(octv), ptr(pc),
                                       current_stream_ptr()))
    return torch.cat([lg[:, :2], octv.view(torch.float32).unsqueeze(1), pc.view(torch.float32).unsqueeze(1)], dim=1)


def _unpack_frames(packed_host: torch.Tensor):
    """Host half of frame_info: torch's CPU fp32 sigmoid (so frame2note's `==` / `>=` see the reference's bits) + the ids."""
    p = torch.sigmoid(packed_host[:, :2])
    ids = packed_host[:, 2:].contiguous().view(torch.int32)
    return (p[:, 0].contiguous().numpy(), p[:, 1].contiguous().numpy(), ids[:, 0].contiguous().numpy(),
            ids[:, 1].contiguous().numpy())


def plan_song_batches(n_samples: Sequence[int], hp: "AMTHparams", dur: Optional[float], batch_clips: int):
    """Host-side plan of the evaluation driver: every song is cut into utterances by the reference rule (split_song), the
    utterances of ALL songs are sorted by length so that equal lengths become neighbours, and runs of equal length are cut
    into batches of at most `batch_clips`.  Returns (utterances per song, [batch, ...]) with batch = [(song, utterance index,
    first sample, end sample), ...]; every utterance appears in exactly one batch."""
    jobs, n_utt = [], []
    for si, n in enumerate(n_samples):
        spans = split_song(n, hp, dur)
        n_utt.append(len(spans))
        for ui, (a, b) in enumerate(spans):
            jobs.append((b - a, si, ui, a, b))
    jobs.sort(key=lambda t: (t[0], t[1], t[2]))
    plan, i = [], 0
    while i < len(jobs):
        j = i
        while j < len(jobs) and j - i < batch_clips and jobs[j][0] == jobs[i][0]:
            j += 1
        plan.append([(si, ui, a, b) for _, si, ui, a, b in jobs[i:j]])
        i = j
    return n_utt, plan


def decode_logits_many(per_song: Sequence[torch.Tensor], hp: "AMTHparams") -> List[np.ndarray]:
    """Frame logits of several songs -> their note arrays with ONE device pass and ONE device->host transfer for all of
    them (the per-frame argmax of every song in one svt_frame_postproc launch), instead of three small synchronising
    copies per song; the host then runs sigmoid + frame2note song by song."""
    if not per_song:
        return []
    counts = [int(t.shape[0]) for t in per_song]
    p_on, p_off, octv, pc = frame_info(torch.cat(list(per_song), dim=0), hp)
    res, a = [], 0
    for n in counts:
        res.append(decode_arrays(p_on[a:a + n], p_off[a:a + n], octv[a:a + n], pc[a:a + n], hp.onset_threshold,
                                 hp.offset_threshold, 1.0 / hp.frame_rate))
        a += n
    return res


class AVTranscriber:
    """Online audio-visual AMT (BASELINE config 4): the reference does this in three offline stages with cached features
    (N20EMv2/audio_only/extract_ssl_feats.py, video_only/extract_ssl_feats.py, audio_visual/train_rca_av.py:28-51); here
    wav -> audio lobe, lip video -> video lobe, FusionRCA, Linear head and the decoder run back to back on one GPU.

    audio_lobe: HuggingFaceWav2Vec2, video_lobe: FairseqAVHubertPretrain, fusion: FusionRCA, head: Linear (20 outputs)."""

    def __init__(self, audio_lobe, video_lobe, fusion, head, hparams: Optional[AMTHparams] = None, device="cuda"):
        self.audio_lobe, self.video_lobe, self.fusion, self.head = audio_lobe, video_lobe, fusion, head
        self.hp = hparams or AMTHparams()
        self.device = torch.device(device)
        self.concurrent_streams = True  # run the two encoders on two CUDA streams (their small-batch grids leave SMs idle)
        self._side_stream = None

    @torch.no_grad()
    def logits(self, wav: torch.Tensor, video: torch.Tensor) -> torch.Tensor:
        """wav (B, L), video (B, 1, T, 88, 88) on CUDA -> frame logits (B, T_audio, 20)  (train_rca_av.py:38-49)."""
        if not self.concurrent_streams:
            a = self.audio_lobe(wav)
            v = self.video_lobe({"video": video, "audio": None})
            return self.head(self.fusion(a, v))
        # the two streams of the model are independent until the fusion: at the 4 clips per GPU of config 4 many of their
        # GEMMs have fewer tiles than the GPU has SMs, so the other encoder's kernels fill the idle ones
        main = torch.cuda.current_stream(wav.device)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=wav.device)
        side = self._side_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            a = self.audio_lobe(wav)
        v = self.video_lobe({"video": video, "audio": None})
        main.wait_stream(side)
        a.record_stream(main)
        return self.head(self.fusion(a, v))

    def decode(self, logits: torch.Tensor) -> np.ndarray:
        return decode_logits(logits, self.hp)

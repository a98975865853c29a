"""Note-level transcription metrics of the recipes' evaluation loop (MIR_ST500/train_audio_ssl.py:119-134,
N20EMv2/audio_visual/train_rca_av.py:119-140): COnPOff / COnP / COn (and COff) precision, recall and F-measure, averaged
over songs with `AverageMeter` (MIR_ST500/utils.py:222-238).

The reference calls `mir_eval.transcription.evaluate` (third-party, version unpinned, not installable here).  This module
restates its published algorithm on the host -- it is O(notes^2) integer / float64 work per song, nowhere near the hot path:

  * a reference note and an estimated note MATCH when |onset difference| <= onset_tolerance (differences rounded to 6
    decimals), pitches are within `pitch_tolerance` cents, and -- for COnPOff -- |offset difference| <=
    max(offset_ratio * reference duration, offset_min_tolerance), comparisons non-strict;
  * every note is used at most once: the score counts a MAXIMUM bipartite matching of the match graph (its size is
    unique, so precision / recall / F do not depend on which maximum matching is found);
  * precision = |matching| / |est|, recall = |matching| / |ref|, F = 2PR / (P + R); all three are 0 when either side is
    empty.

**Parity unpinned**: mir_eval cannot be imported in the authoring container, so these functions are pinned by hand-built cases and by a
brute-force matcher in tests/test_metrics_cpu.py, not by outputs of the reference's own dependency."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

N_DECIMALS = 6  # rounding of time differences before the tolerance test


def midi_to_hz(midi):
    """mir_eval.util.midi_to_hz: 440 * 2^((m - 69) / 12)."""
    return 440.0 * (2.0 ** ((np.asarray(midi, dtype=np.float64) - 69.0) / 12.0))


def _validate(intervals, pitches, what):
    intervals = np.asarray(intervals, dtype=np.float64).reshape(-1, 2)
    pitches = np.asarray(pitches, dtype=np.float64).reshape(-1)
    if intervals.shape[0] != pitches.shape[0]:
        raise ValueError(f"{what}: {intervals.shape[0]} intervals but {pitches.shape[0]} pitches")
    if intervals.size and (intervals[:, 1] - intervals[:, 0] <= 0).any():
        raise ValueError(f"{what}: all note durations must be strictly positive")
    if pitches.size and (pitches <= 0).any():
        raise ValueError(f"{what}: pitches must be positive frequencies in Hz")
    return intervals, pitches


def _max_bipartite_matching(adj: List[List[int]], n_right: int) -> int:
    """Size of a maximum matching; adj[u] lists the right vertices of left vertex u.  Augmenting paths (Kuhn) with an
    explicit stack, so long alternating chains in a dense song cannot hit Python's recursion limit."""
    match_r = [-1] * n_right   # right vertex -> left vertex
    size = 0
    for u0 in range(len(adj)):
        seen = [False] * n_right
        path = [(u0, iter(adj[u0]))]  # left vertices of the alternating path being grown, with their edge cursors
        entered = [-1]                # entered[i]: right vertex through which path[i] was reached
        while path:
            u, it = path[-1]
            v = next((x for x in it if not seen[x]), None)
            if v is None:
                path.pop()
                entered.pop()
                continue
            seen[v] = True
            if match_r[v] < 0:        # free right vertex: flip every edge of the path
                for i in range(len(path) - 1, -1, -1):
                    match_r[v] = path[i][0]
                    v = entered[i]
                size += 1
                break
            path.append((match_r[v], iter(adj[match_r[v]])))
            entered.append(v)
    return size


def _hit_matrix(ref_intervals, ref_pitches, est_intervals, est_pitches, onset_tolerance, pitch_tolerance, offset_ratio,
                offset_min_tolerance, use_onset=True, use_pitch=True):
    hit = np.ones((len(ref_pitches), len(est_pitches)), dtype=bool)
    if use_onset:
        d = np.around(np.abs(np.subtract.outer(ref_intervals[:, 0], est_intervals[:, 0])), decimals=N_DECIMALS)
        hit &= d <= onset_tolerance
    if use_pitch:
        d = np.abs(np.subtract.outer(1200.0 * np.log2(ref_pitches), 1200.0 * np.log2(est_pitches)))
        hit &= d <= pitch_tolerance
    if offset_ratio is not None:
        d = np.around(np.abs(np.subtract.outer(ref_intervals[:, 1], est_intervals[:, 1])), decimals=N_DECIMALS)
        tol = offset_ratio * (ref_intervals[:, 1] - ref_intervals[:, 0])
        tol = np.where(tol <= offset_min_tolerance, offset_min_tolerance, tol)
        hit &= d <= tol.reshape(-1, 1)
    return hit


def _prf(hit: np.ndarray):
    n_ref, n_est = hit.shape
    if n_ref == 0 or n_est == 0:
        return 0.0, 0.0, 0.0
    adj = [list(np.nonzero(hit[:, j])[0]) for j in range(n_est)]  # est -> refs, as mir_eval builds its graph
    m = _max_bipartite_matching(adj, n_ref)
    p, r = m / n_est, m / n_ref
    f = 0.0 if p == 0 and r == 0 else 2 * p * r / (p + r)
    return p, r, f


def evaluate(ref_intervals, ref_pitches, est_intervals, est_pitches, onset_tolerance=0.05, pitch_tolerance=50.0,
             offset_ratio=0.2, offset_min_tolerance=0.05) -> Dict[str, float]:
    """Same keys as mir_eval.transcription.evaluate (minus the average-overlap ratios the recipes never read).
    Pitches in Hz (the recipes convert MIDI with midi_to_hz first, train_audio_ssl.py:117-118)."""
    ri, rp = _validate(ref_intervals, ref_pitches, "reference")
    ei, ep = _validate(est_intervals, est_pitches, "estimate")
    a = (ri, rp, ei, ep, onset_tolerance, pitch_tolerance)
    out = {}
    out["Precision"], out["Recall"], out["F-measure"] = _prf(_hit_matrix(*a, offset_ratio, offset_min_tolerance))
    out["Precision_no_offset"], out["Recall_no_offset"], out["F-measure_no_offset"] = _prf(_hit_matrix(*a, None, offset_min_tolerance))
    out["Onset_Precision"], out["Onset_Recall"], out["Onset_F-measure"] = _prf(
        _hit_matrix(*a, None, offset_min_tolerance, use_pitch=False))
    out["Offset_Precision"], out["Offset_Recall"], out["Offset_F-measure"] = _prf(
        _hit_matrix(*a, offset_ratio, offset_min_tolerance, use_onset=False, use_pitch=False))
    return out


class AverageMeter:
    """MIR_ST500/utils.py:222-238."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class TranscriptionMeters:
    """The nine meters the evaluation stage keeps (train_audio_ssl.py:58-69,126-134): COnPOff / COnP / COn x P / R / F1."""

    _KEYS = {"COnPOff": ("Precision", "Recall", "F-measure"),
             "COnP": ("Precision_no_offset", "Recall_no_offset", "F-measure_no_offset"),
             "COn": ("Onset_Precision", "Onset_Recall", "Onset_F-measure")}

    def __init__(self, onset_tolerance=0.05, pitch_tolerance=50.0):
        self.onset_tolerance, self.pitch_tolerance = onset_tolerance, pitch_tolerance
        self.meters = {f"{m}_{s}": AverageMeter() for m in self._KEYS for s in ("precis", "recall", "f1")}

    def update(self, ref_notes: Sequence, est_notes: Sequence) -> Dict[str, float]:
        """ref_notes / est_notes: (n, 3) arrays [onset s, offset s, MIDI pitch] -- the decoder's output format."""
        ref = np.asarray(ref_notes, dtype=np.float64).reshape(-1, 3)
        est = np.asarray(est_notes, dtype=np.float64).reshape(-1, 3)
        raw = evaluate(ref[:, :2], midi_to_hz(ref[:, 2]), est[:, :2], midi_to_hz(est[:, 2]),
                       onset_tolerance=self.onset_tolerance, pitch_tolerance=self.pitch_tolerance)
        for m, keys in self._KEYS.items():
            for s, k in zip(("precis", "recall", "f1"), keys):
                self.meters[f"{m}_{s}"].update(raw[k])
        return raw

    def summary(self) -> Dict[str, float]:
        return {k: m.avg for k, m in self.meters.items()}

"""Note-level metrics (svt_speechbrain_b200/metrics.py) against hand-built cases and a brute-force maximum matching.
mir_eval is not installable here: parity with the reference's dependency is unpinned (stated in the module)."""
import itertools

import numpy as np
import pytest

from svt_speechbrain_b200 import metrics as M


def _notes(rows):
    a = np.asarray(rows, dtype=np.float64).reshape(-1, 3)
    return a[:, :2], M.midi_to_hz(a[:, 2])


def test_midi_to_hz():
    assert M.midi_to_hz(69) == 440.0 and abs(M.midi_to_hz(57) - 220.0) < 1e-12 and abs(M.midi_to_hz(81) - 880.0) < 1e-9


def test_perfect_and_empty():
    ri, rp = _notes([[0.0, 0.5, 60], [1.0, 1.4, 62]])
    r = M.evaluate(ri, rp, ri, rp)
    assert all(r[k] == 1.0 for k in ("Precision", "Recall", "F-measure", "F-measure_no_offset", "Onset_F-measure", "Offset_F-measure"))
    e = M.evaluate(ri, rp, np.zeros((0, 2)), np.zeros(0))
    assert all(v == 0.0 for v in e.values())


def test_tolerances_are_inclusive_and_rounded():
    ri, rp = _notes([[1.0, 2.0, 60]])
    # onset exactly 50 ms late (floating point makes 1.05 - 1.0 slightly > 0.05: the 6-decimal rounding admits it)
    ei, ep = _notes([[1.05, 2.0, 60]])
    assert M.evaluate(ri, rp, ei, ep)["Onset_F-measure"] == 1.0
    ei, ep = _notes([[1.0501, 2.0, 60]])
    assert M.evaluate(ri, rp, ei, ep)["Onset_F-measure"] == 0.0
    # pitch: 50 cents inclusive
    ei = ri.copy()
    assert M.evaluate(ri, rp, ei, rp * 2 ** (50 / 1200))["F-measure_no_offset"] == 1.0
    assert M.evaluate(ri, rp, ei, rp * 2 ** (50.01 / 1200))["F-measure_no_offset"] == 0.0
    # offset: max(0.2 * 1.0 s, 0.05) = 0.2 s; a 0.1-s note falls back to the 50-ms floor
    ei, ep = _notes([[1.0, 2.2, 60]])
    assert M.evaluate(ri, rp, ei, ep)["F-measure"] == 1.0
    ei, ep = _notes([[1.0, 2.21, 60]])
    r = M.evaluate(ri, rp, ei, ep)
    assert r["F-measure"] == 0.0 and r["F-measure_no_offset"] == 1.0 and r["Offset_F-measure"] == 0.0
    ri, rp = _notes([[1.0, 1.1, 60]])
    ei, ep = _notes([[1.0, 1.15, 60]])
    assert M.evaluate(ri, rp, ei, ep)["F-measure"] == 1.0
    ei, ep = _notes([[1.0, 1.16, 60]])
    assert M.evaluate(ri, rp, ei, ep)["F-measure"] == 0.0


def test_each_note_is_used_once_and_the_matching_is_maximum():
    # two estimates both within tolerance of ref 0, only one of them within tolerance of ref 1: a greedy matcher that gives
    # est 1 to ref 0 finds one pair, the maximum matching finds two
    ri, rp = _notes([[1.00, 1.5, 60], [1.08, 1.6, 60]])
    ei, ep = _notes([[1.04, 1.5, 60], [1.00, 1.5, 60]])
    r = M.evaluate(ri, rp, ei, ep)
    assert r["Onset_Precision"] == 1.0 and r["Onset_Recall"] == 1.0
    # three estimates crowding one reference note: one hit
    ei, ep = _notes([[1.0, 1.5, 60], [1.01, 1.5, 60], [1.02, 1.5, 60]])
    ri, rp = _notes([[1.0, 1.5, 60]])
    r = M.evaluate(ri, rp, ei, ep)
    assert r["Precision"] == pytest.approx(1 / 3) and r["Recall"] == 1.0 and r["F-measure"] == pytest.approx(0.5)


def _brute_force(hit):
    n_ref, n_est = hit.shape
    best = 0
    for k in range(min(n_ref, n_est), 0, -1):
        for refs in itertools.permutations(range(n_ref), k):
            for ests in itertools.combinations(range(n_est), k):
                if all(hit[r, e] for r, e in zip(refs, ests)):
                    return k
    return best


def test_matching_size_against_brute_force_on_random_graphs():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n_ref, n_est = rng.integers(1, 6, 2)
        hit = rng.random((n_ref, n_est)) < 0.4
        p, r, f = M._prf(hit)
        k = _brute_force(hit)
        assert p == pytest.approx(k / n_est) and r == pytest.approx(k / n_ref)


def test_meters_average_over_songs():
    tm = M.TranscriptionMeters()
    tm.update([[0, 1, 60]], [[0, 1, 60]])
    tm.update([[0, 1, 60]], [[0.5, 1.5, 61]])
    s = tm.summary()
    assert s["COnPOff_f1"] == 0.5 and s["COn_recall"] == 0.5 and tm.meters["COnP_f1"].count == 2
    with pytest.raises(ValueError):
        M.evaluate([[1.0, 1.0]], [440.0], [[0, 1]], [440.0])

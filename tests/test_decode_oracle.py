"""frame2note: Python + C restatements vs the reference's golden outputs; CPython set-order quirks."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle.frame2note_oracle import frame2note as f2n_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def clib():
    so = os.path.join(ROOT, "oracle", "_build", "libframe2note_oracle.so")
    if not os.path.exists(so):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "oracle", "frame2note_oracle.c")])
    return ctypes.CDLL(so)


def c_frame2note(lib, p_on, p_off, octv, pc, on_t, off_t, fs=1 / 49.8):
    n = len(p_on)
    out = np.zeros((n + 1, 3), dtype=np.float64)
    p_on = np.ascontiguousarray(p_on, np.float32)
    p_off = np.ascontiguousarray(p_off, np.float32)
    o = np.ascontiguousarray(octv, np.int32)
    c = np.ascontiguousarray(pc, np.int32)
    lib.frame2note_oracle.restype = ctypes.c_int
    k = lib.frame2note_oracle(
        p_on.ctypes.data_as(ctypes.c_void_p), p_off.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p),
        c.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), ctypes.c_double(on_t), ctypes.c_double(off_t),
        ctypes.c_double(fs), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n + 1))
    assert k >= 0
    return out[:k]


def _cases():
    g = np.load(os.path.join(GOLD, "frame2note_cases.npz"))
    keys = sorted({k.split("/")[0] for k in g.files if k.endswith("/notes")})
    for key in keys:
        base = key.rsplit("_", 1)[0]
        yield key, g[key + "/notes"], g[key + "/thr"], g[base + "/p_on"], g[base + "/p_off"], g[base + "/oct"], g[base + "/pc"]


def test_python_restatement_matches_reference_golden():
    for key, notes, thr, p_on, p_off, octv, pc in _cases():
        t_on, t_off = torch.from_numpy(p_on), torch.from_numpy(p_off)
        fi = [(t_on[i], t_off[i], int(octv[i]), int(pc[i])) for i in range(len(p_on))]
        got = np.array(f2n_py(fi, float(thr[0]), float(thr[1])), dtype=np.float64).reshape(-1, 3)
        assert got.shape == notes.shape, key
        assert np.array_equal(got, notes), key  # bit-exact float64 times and pitches


def test_c_restatement_matches_reference_golden(clib):
    total = 0
    for key, notes, thr, p_on, p_off, octv, pc in _cases():
        got = c_frame2note(clib, p_on, p_off, octv, pc, float(thr[0]), float(thr[1]))
        assert got.shape == notes.shape, key
        assert np.array_equal(got, notes), key
        total += len(notes)
    assert total > 500


def test_set_order_tie_breaks(clib):
    # [8,1,8,1]->8, [1,8,1,8]->8, [0,8,0,8]->0, [8,0,8,0]->8, [40,9,40,9]->40
    for seq, want in [([8, 1, 8, 1], 8), ([1, 8, 1, 8], 8), ([0, 8, 0, 8], 0), ([8, 0, 8, 0], 8), ([40, 9, 40, 9], 40)]:
        assert max(set(seq), key=seq.count) == want
        n = len(seq) + 2
        p_on = np.full(n, 0.1, np.float32); p_on[0] = 0.9
        p_off = np.full(n, 0.1, np.float32); p_off[-2] = 0.9
        octv = np.array([s // 12 for s in seq] + [4, 4]); pc = np.array([s % 12 for s in seq] + [12, 12])
        got = c_frame2note(clib, p_on, p_off, octv, pc, 0.4, 0.5)
        assert got.shape == (1, 3) and got[0, 2] == want + 36


def test_c_vs_python_random_mode_order(clib):
    rng = np.random.default_rng(0)
    for trial in range(300):
        k = int(rng.integers(1, 60))
        alphabet = rng.choice(48, size=int(rng.integers(1, 48)), replace=False)
        seq = [int(x) for x in rng.choice(alphabet, size=k)]
        want = max(set(seq), key=seq.count)
        n = k + 2
        p_on = np.full(n, 0.1, np.float32); p_on[0] = 0.9
        p_off = np.full(n, 0.1, np.float32); p_off[k] = 0.9
        octv = np.array([s // 12 for s in seq] + [4, 4]); pc = np.array([s % 12 for s in seq] + [12, 12])
        got = c_frame2note(clib, p_on, p_off, octv, pc, 0.4, 0.5)
        assert got[0, 2] == want + 36, (seq, want, got)

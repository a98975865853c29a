"""Drop-in for the recipes' `utils.frame2note` (MIR_ST500/utils.py:82-149): same signature, same return type
(python list of [float, float, int]), decoded by the C-ABI host function svt_frame2note."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib


def _as_f32(col):
    if len(col) and torch.is_tensor(col[0]):
        return torch.stack([c.detach().float().cpu().reshape(()) for c in col]).numpy()
    return np.asarray(col, dtype=np.float32)


def decode_arrays(p_on, p_off, octv, pc, onset_thres, offset_thres, frame_size=1 / 49.8) -> np.ndarray:
    """Array form: fp32 p_on/p_off, integer octave / pitch-class ids -> float64 (n_notes, 3)."""
    p_on = np.ascontiguousarray(p_on, dtype=np.float32)
    p_off = np.ascontiguousarray(p_off, dtype=np.float32)
    octv = np.ascontiguousarray(octv, dtype=np.int32)
    pc = np.ascontiguousarray(pc, dtype=np.int32)
    n = int(p_on.shape[0])
    out = np.empty((max(n, 1), 3), dtype=np.float64)
    count = C.c_int(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().svt_frame2note(vp(p_on), vp(p_off), vp(octv), vp(pc), n, float(onset_thres), float(offset_thres),
                               float(frame_size), vp(out), out.shape[0], C.byref(count)))
    return out[: count.value].copy()


def frame2note(frame_info, onset_thres, offset_thres, frame_size=1 / 49.8):
    """frame_info: sequence of (onset_prob, offset_prob, octave, pitch_class) exactly as built at
    MIR_ST500/train_audio_ssl.py:95-100 (0-d fp32 tensors / floats and ints)."""
    n = len(frame_info)
    if n == 0:
        return []
    p_on = _as_f32([f[0] for f in frame_info])
    p_off = _as_f32([f[1] for f in frame_info])
    octv = np.asarray([int(f[2]) for f in frame_info], dtype=np.int32)
    pc = np.asarray([int(f[3]) for f in frame_info], dtype=np.int32)
    try:
        notes = decode_arrays(p_on, p_off, octv, pc, onset_thres, offset_thres, frame_size)
    except Exception as e:  # the reference raises ValueError from np.amax on an empty window
        if "empty local-max window" in str(e):
            raise ValueError("zero-size array to reduction operation maximum which has no identity") from e
        raise
    return [[float(a), float(b), int(c)] for a, b, c in notes]

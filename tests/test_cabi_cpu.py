"""CPU-side checks of the boundary: the shared library loads, exports every symbol the header declares, and
fails loudly (no fallback) when asked to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import svt_speechbrain_b200 as svt
from svt_speechbrain_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "svt_b200.h")).read()
    declared = set(re.findall(r"\b(svt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = svt.lib()
    for name in declared:
        assert hasattr(L, name), name
    # every declared function has a ctypes signature in the binding (and vice versa)
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert L.svt_version() >= 100


def test_config_struct_layout_matches_header():
    # 6 ints + 2 * 8 ints + 5 ints + float + 7 ints
    assert C.sizeof(_lib.EncoderConfig) == 4 * (6 + 16 + 5 + 1 + 7)
    assert C.sizeof(_lib.FusionConfig) == 16


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_gpu_is_a_loud_error_not_a_fallback():
    L = svt.lib()
    assert L.svt_device_count() == 0
    cfg = _lib.EncoderConfig()
    cfg.hidden_size, cfg.num_layers, cfg.num_heads, cfg.ffn_size = 1024, 1, 16, 4096
    cfg.num_conv_layers, cfg.conv_dim = 7, 512
    for i, (k, s) in enumerate(zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2))):
        cfg.conv_kernel[i], cfg.conv_stride[i] = k, s
    cfg.pos_conv_kernel, cfg.pos_conv_groups, cfg.layer_norm_eps = 128, 16, 1e-5
    h = C.c_void_p()
    assert L.svt_encoder_create(C.byref(cfg), C.byref(h)) == 0
    assert L.svt_encoder_num_frames(h, 160000) == 499
    assert L.svt_encoder_workspace_bytes(h, 64, 160000) > 3 << 30
    x = np.zeros(4, np.float32)
    shape = (C.c_int64 * 1)(4)
    rc = L.svt_encoder_set_tensor(h, b"model.encoder.layer_norm.weight", x.ctypes.data_as(C.c_void_p), shape, 1, 0)
    assert rc == 7 and b"no CUDA device" in L.svt_last_error()
    assert L.svt_encoder_finalize(h) == 7
    L.svt_encoder_destroy(h)
    with pytest.raises(RuntimeError):
        svt.Linear(n_neurons=20, input_size=1024)(torch.zeros(1, 3, 1024))
    with pytest.raises(RuntimeError):
        svt.FusionRCA()(torch.zeros(1, 3, 1024), torch.zeros(1, 3, 1024))


def test_invalid_configs_are_rejected():
    L = svt.lib()
    cfg = _lib.EncoderConfig()
    h = C.c_void_p()
    assert L.svt_encoder_create(C.byref(cfg), C.byref(h)) != 0
    assert len(L.svt_last_error()) > 0
    f = _lib.FusionConfig(1000, 8, 3072, 0.5)
    assert L.svt_fusion_create(C.byref(f), C.byref(h)) != 0


def test_decoder_through_cabi_matches_reference_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "frame2note_cases.npz"))
    keys = sorted({k.split("/")[0] for k in g.files if k.endswith("/notes")})
    n_notes = 0
    for key in keys:
        base = key.rsplit("_", 1)[0]
        thr = g[key + "/thr"]
        got = svt.decode_arrays(g[base + "/p_on"], g[base + "/p_off"], g[base + "/oct"], g[base + "/pc"], thr[0], thr[1])
        assert got.shape == g[key + "/notes"].shape and np.array_equal(got, g[key + "/notes"]), key
        n_notes += len(got)
    assert n_notes > 500


def test_frame2note_dropin_signature_and_types():
    fi = [(torch.tensor(0.9), torch.tensor(0.1), 1, 2), (torch.tensor(0.1), torch.tensor(0.1), 1, 2),
          (torch.tensor(0.1), torch.tensor(0.9), 4, 12), (torch.tensor(0.2), torch.tensor(0.1), 4, 12)]
    notes = svt.frame2note(fi, onset_thres=0.4, offset_thres=0.5, frame_size=1 / 49.8)
    assert notes == [[0.0, 2 * (1 / 49.8), 50]] and isinstance(notes[0][2], int)
    assert svt.frame2note([], 0.4, 0.5) == []
    with pytest.raises(ValueError):  # reference: np.amax of an empty window
        svt.frame2note([(torch.tensor(0.9), torch.tensor(0.1), 1, 2)], 0.4, 0.5)


def test_split_song_reference_rule():
    hp = svt.AMTHparams()
    spans = svt.split_song(16000 * 300, hp, dur=10.0)
    assert len(spans) == 30 and spans[0] == (0, 160000) and spans[-1] == (29 * 160000, 300 * 16000)
    spans = svt.split_song(int(16000 * 17.4), hp)  # round(17.4/5) = 3 utterances, last takes the remainder (7.4 s)
    assert len(spans) == 3 and spans[-1] == (160000, int(16000 * 17.4))
    spans = svt.split_song(16000 * 2, hp)  # shorter than half a window still yields one utterance
    assert spans == [(0, 32000)]


def test_shard_ranges_cover_everything():
    from svt_speechbrain_b200.parallel import shard_range
    for n in (512, 13, 7, 1):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def test_wavlm_relative_bucket_matches_torch():
    """svt_wavlm_relative_bucket (host, fp32) == WavLMAttention._relative_positions_bucket as the oracle restates it
    (pinned to the reference by the wavlm_* goldens), for every relative position a 60-s clip can produce."""
    import torch  # noqa: F401
    from oracle import wav2vec2_oracle as wo
    cfg = wo.W2V2Config.wavlm_base()
    T = 3000
    b = wo.wavlm_relative_buckets(cfg, T)
    L = _lib.lib()
    for d in range(T):
        assert L.svt_wavlm_relative_bucket(d, cfg.num_buckets, cfg.max_bucket_distance) == int(b[0, d])
        assert L.svt_wavlm_relative_bucket(-d, cfg.num_buckets, cfg.max_bucket_distance) == int(b[d, 0])


def _header_struct_fields(name):
    """Field names of `typedef struct <name> { ... }` in include/svt_b200.h, in order (comments stripped)."""
    hdr = open(os.path.join(ROOT, "include", "svt_b200.h")).read()
    body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + r";", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return [re.match(r"\s*(?:const\s+)?\w+\s+(\w+)", d).group(1) for d in body.split(";") if d.strip()]


def test_ctypes_structs_follow_the_header_field_for_field():
    for cname, cls in (("svt_encoder_config", _lib.EncoderConfig), ("svt_fusion_config", _lib.FusionConfig),
                       ("svt_video_config", _lib.VideoConfig)):
        assert _header_struct_fields(cname) == [f[0] for f in cls._fields_], cname


def test_integration_md_ctypes_stub_runs():
    """INTEGRATION.md's ctypes stub is the reference-side binding a maintainer would add; everything above its
    'a B200 is needed' marker must execute as written (struct layout, config construction, host-only entry points)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(import ctypes as C, os\n.*?)```", text, re.S).group(1)
    head, marker, _ = block.partition("# ---- from here on a B200 is needed")
    assert marker
    env = {}
    os.environ["SVT_B200_LIB"] = svt.LIB_PATH
    try:
        exec(compile(head, "INTEGRATION.md", "exec"), env)
    finally:
        os.environ.pop("SVT_B200_LIB", None)
    assert env["T"] == 499
    assert C.sizeof(env["EncoderConfig"]) == C.sizeof(_lib.EncoderConfig)
    assert [f[0] for f in env["EncoderConfig"]._fields_] == [f[0] for f in _lib.EncoderConfig._fields_]
    env["L"].svt_encoder_destroy(env["h"])

"""tcgen05 / TMEM flash attention (csrc/attention_tc.cu) vs a torch fp32 reference and vs the mma.sync kernel.
Own file = own process in tools/run_gpu_checks.sh, so a device trap here cannot poison the other GPU tests."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _set_impl(v):
    from svt_speechbrain_b200._lib import check, lib
    check(lib().svt_set_option(b"attention_impl", v))


@pytest.fixture(autouse=True)
def _restore_impl():
    yield
    _set_impl(0)


def _ref(qkv, clips, Ta, T, heads, dh):
    D = heads * dh
    x = qkv.float().view(clips, Ta, 3, heads, dh)[:, :T]
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
    return ref.permute(0, 2, 1, 3).reshape(clips, T, D)


@pytest.mark.parametrize("heads,T,Ta,clips", [(16, 499, 500, 2), (12, 49, 50, 3), (4, 130, 130, 1), (2, 128, 128, 1),
                                               (3, 257, 260, 2), (16, 249, 250, 5), (1, 1, 4, 1), (2, 640, 640, 1)])
def test_attention_tc_vs_torch(heads, T, Ta, clips):
    from gpu_util import op_attention, rel_l2
    dh = 64
    D = heads * dh
    g = torch.Generator(device="cuda").manual_seed(T + heads)
    qkv = torch.randn(clips * Ta, 3 * D, device="cuda", generator=g).bfloat16()
    _set_impl(2)
    o = op_attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], T, T, clips, heads, dh, Ta, Ta)
    ref = _ref(qkv, clips, Ta, T, heads, dh)
    got = o.float().view(clips, Ta, D)[:, :T]
    err = (got - ref).abs().max().item()
    print(f"attention_tc heads={heads} T={T}: max abs err {err:.3e} rel_l2 {rel_l2(got, ref):.3e}")
    assert rel_l2(got, ref) < 1e-2 and err < 5e-2
    assert (o.view(clips, Ta, D)[:, T:] == 0).all()  # rows past T are never written
    _set_impl(1)
    o1 = op_attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], T, T, clips, heads, dh, Ta, Ta)
    assert rel_l2(o.float(), o1.float()) < 1e-2


def test_attention_tc_large_logits():
    """scores with a wide dynamic range (|s| up to ~60): the running-max rescaling must not lose rows."""
    from gpu_util import op_attention, rel_l2
    heads, T, Ta, clips, dh = 4, 300, 304, 2, 64
    D = heads * dh
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(clips * Ta, 3 * D, device="cuda", generator=g)
    qkv[:, :2 * D] *= 2.5
    qkv = qkv.bfloat16()
    _set_impl(2)
    o = op_attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], T, T, clips, heads, dh, Ta, Ta)
    ref = _ref(qkv, clips, Ta, T, heads, dh)
    got = o.float().view(clips, Ta, D)[:, :T]
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) < 1.5e-2


def test_attention_tc_cross_lengths():
    from gpu_util import op_attention, rel_l2
    dh, heads, Tq, Tk, clips = 64, 4, 77, 300, 2
    D = heads * dh
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(clips * Tq, D, device="cuda", generator=g).bfloat16()
    kv = torch.randn(clips * Tk, 3 * D, device="cuda", generator=g).bfloat16()
    _set_impl(2)
    o = op_attention(q, kv[:, D:2 * D], kv[:, 2 * D:], Tq, Tk, clips, heads, dh, Tq, Tk, ldo=2 * D)
    qq = q.float().view(clips, Tq, heads, dh).permute(0, 2, 1, 3)
    kk = kv[:, D:2 * D].float().reshape(clips, Tk, heads, dh).permute(0, 2, 1, 3)
    vv = kv[:, 2 * D:].float().reshape(clips, Tk, heads, dh).permute(0, 2, 1, 3)
    ref = (torch.softmax(qq @ kk.transpose(-1, -2), dim=-1) @ vv).permute(0, 2, 1, 3).reshape(clips * Tq, D)
    assert rel_l2(o[:, :D].float(), ref) < 1e-2

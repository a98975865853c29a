"""attention / layer norm / conv0 / head kernels vs torch fp32 references (C ABI hooks)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dh,heads,T,Ta,clips", [(64, 16, 499, 500, 2), (64, 12, 49, 50, 3), (64, 4, 130, 130, 1),
                                                  (128, 8, 499, 499, 2), (128, 8, 64, 64, 1)])
def test_attention_self(dh, heads, T, Ta, clips):
    from gpu_util import op_attention, rel_l2
    D = heads * dh
    g = torch.Generator(device="cuda").manual_seed(T + dh)
    qkv = torch.randn(clips * Ta, 3 * D, device="cuda", generator=g).bfloat16()
    o = op_attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], T, T, clips, heads, dh, Ta, Ta)
    x = qkv.float().view(clips, Ta, 3, heads, dh)[:, :T]
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
    ref = ref.permute(0, 2, 1, 3).reshape(clips, T, D)
    got = o.float().view(clips, Ta, D)[:, :T]
    err = (got - ref).abs().max().item()
    print(f"attention dh={dh} T={T}: max abs err {err:.3e} rel_l2 {rel_l2(got, ref):.3e}")
    assert rel_l2(got, ref) < 1e-2 and err < 5e-2
    assert (o.view(clips, Ta, D)[:, T:] == 0).all()


def test_attention_cross():
    from gpu_util import op_attention, rel_l2
    dh, heads, Tq, Tk, clips = 128, 8, 77, 130, 2
    D = heads * dh
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(clips * Tq, D, device="cuda", generator=g).bfloat16()
    kv = torch.randn(clips * Tk, 3 * D, device="cuda", generator=g).bfloat16()
    o = op_attention(q, kv[:, D:2 * D], kv[:, 2 * D:], Tq, Tk, clips, heads, dh, Tq, Tk, ldo=2 * D)
    qq = q.float().view(clips, Tq, heads, dh).permute(0, 2, 1, 3)
    kk = kv[:, D:2 * D].float().reshape(clips, Tk, heads, dh).permute(0, 2, 1, 3)
    vv = kv[:, 2 * D:].float().reshape(clips, Tk, heads, dh).permute(0, 2, 1, 3)
    ref = (torch.softmax(qq @ kk.transpose(-1, -2), dim=-1) @ vv).permute(0, 2, 1, 3).reshape(clips * Tq, D)
    assert rel_l2(o[:, :D].float(), ref) < 1e-2


@pytest.mark.parametrize("D", [512, 768, 1024])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("gelu", [0, 1])
def test_layer_norm(D, dtype, gelu):
    from gpu_util import op_layer_norm
    g = torch.Generator(device="cuda").manual_seed(D)
    x = (torch.randn(1003, D, device="cuda", generator=g) * 3 + 0.5).to(dtype)
    gamma = torch.randn(D, device="cuda", generator=g)
    beta = torch.randn(D, device="cuda", generator=g)
    yf, yb = op_layer_norm(x, gamma, beta, 1e-5, gelu)
    ref = torch.nn.functional.layer_norm(x.float(), (D,), gamma, beta, 1e-5)
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    assert (yf - ref).abs().max().item() < 2e-4
    assert ((yb.float() - ref).abs() / (1.0 + ref.abs())).max().item() < 6e-3  # bf16 rounding, 2^-9 relative


def test_conv0_layer_mode():
    from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr
    B, L = 3, 16000
    T = (L - 10) // 5 + 1
    Ta = (T + 63) // 64 * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    wav = torch.randn(B, L, device="cuda", generator=g) * 2 + 0.3
    w = torch.randn(512, 1, 10, device="cuda", generator=g) * 0.3
    bias, gamma, beta = (torch.randn(512, device="cuda", generator=g) for _ in range(3))
    out = torch.full((B, Ta, 512), float("nan"), device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(2, device="cuda", dtype=torch.float64)
    wkc = w[:, 0, :].t().contiguous()  # [k][C]
    check(lib().svt_op_conv0(ptr(wav), B, L, ptr(wkc), ptr(bias), ptr(gamma), ptr(beta), 1, ptr(out), Ta, ptr(stats),
                             current_stream_ptr()))
    torch.cuda.synchronize()
    xn = torch.nn.functional.layer_norm(wav, wav.shape)
    ref = torch.nn.functional.conv1d(xn[:, None], w, bias, stride=5).transpose(1, 2)
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(ref, (512,), gamma, beta, 1e-5))
    err = ((out[:, :T].float() - ref).abs() / (1.0 + ref.abs())).max().item()
    print("conv0 max scaled err", err)
    assert err < 6e-3  # bf16 output rounding (2^-9 relative) plus fp32 arithmetic differences
    assert (out[:, T:] == 0).all()


def test_linear_small_and_postproc():
    from svt_speechbrain_b200 import Linear
    from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr
    torch.manual_seed(0)
    lin = Linear(n_neurons=20, input_size=1024).cuda()
    x = torch.randn(2, 37, 1024, device="cuda")
    y = lin(x)
    ref = torch.nn.functional.linear(x, lin.w.weight, lin.w.bias)
    assert (y - ref).abs().max().item() < 1e-4
    lg = y.view(-1, 20).contiguous()
    lg[3, 2] = lg[3, 4] = 9.0  # tie -> first index wins
    octv = torch.empty(lg.shape[0], dtype=torch.int32, device="cuda")
    pc = torch.empty_like(octv)
    check(lib().svt_frame_postproc(ptr(lg), lg.shape[0], 20, 2, 5, 7, 13, ptr(octv), ptr(pc), current_stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(octv.long(), lg[:, 2:7].argmax(1)) and torch.equal(pc.long(), lg[:, 7:20].argmax(1))
    assert octv[3].item() == 0


def test_gelu_accuracy():
    """The fused epilogue GELU (exact-erf form, A&S 7.1.26 erfc) against float64 erf GELU, via LN(gelu=1) with
    identity affine on rows engineered to be already normalised."""
    import math
    from gpu_util import op_layer_norm
    D = 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(512, D, device="cuda", generator=g) * 2.5
    x = (x - x.mean(1, keepdim=True)) / x.var(1, unbiased=False, keepdim=True).add(1e-5).sqrt()
    gamma = torch.full((D,), 3.0, device="cuda")  # spread the arguments over (-12, 12)
    beta = torch.zeros(D, device="cuda")
    yf, _ = op_layer_norm(x, gamma, beta, 1e-5, 1)
    z = torch.nn.functional.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-5)
    ref = 0.5 * z * (1 + torch.erf(z / math.sqrt(2)))
    err = (yf.double() - ref).abs().max().item()
    print("gelu max abs err vs float64 erf-GELU", err)
    assert err < 5e-6

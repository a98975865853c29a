"""CTA-pair (cta_group::2) tcgen05 GEMM (csrc/gemm_tc2.cu) vs torch fp32 matmul and vs the one-CTA kernel.
Own file = own process in tools/run_gpu_checks.sh."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _set_impl(v):
    from svt_speechbrain_b200._lib import check, lib
    check(lib().svt_set_option(b"gemm_impl", v))


@pytest.fixture(autouse=True)
def _restore_impl():
    yield
    _set_impl(0)


def _ref(a, w, bias, resid, act):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.gelu(y)
    elif act == 2:
        y = torch.relu(y)
    if resid is not None:
        y = y + resid
    return y


@pytest.mark.parametrize("M,N,K", [(1024, 256, 64), (1024, 256, 256), (1280, 512, 1536), (2381, 1024, 512), (4096, 3072, 1024),
                                    (32000, 1024, 1024), (1025, 768, 3072), (40000, 512, 128)])
def test_pair_gemm_plain(M, N, K):
    from gpu_util import op_gemm, rel_l2
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    _set_impl(0)
    of, ob = op_gemm(a, w, out_f32=True, out_bf16=True)
    ref = _ref(a, w, None, None, 0)
    err = (of - ref).abs().max().item()
    print(f"pair gemm {M}x{N}x{K}: max abs err fp32 out {err:.3e}, rel_l2 {rel_l2(of, ref):.3e}")
    if not err < 2e-3:
        ok = ((of - ref).abs() < 2e-3)
        print("fraction ok", ok.float().mean().item(), "nan frac", torch.isnan(of).float().mean().item())
        print("ok by 128-row block:", ok.float().mean(1)[: 128 * 16].view(-1, 128).mean(1).tolist())
        print("ok by 64-col block:", ok.float().mean(0).view(-1, 64).mean(1).tolist())
    assert torch.isfinite(of).all()
    assert err < 2e-3, err
    assert rel_l2(ob.float(), ref) < 4e-3
    _set_impl(1)
    of1, _ = op_gemm(a, w, out_f32=True, out_bf16=False)
    assert (of - of1).abs().max().item() < 1e-3  # same products, same fp32 accumulation order per k-block


@pytest.mark.parametrize("act", [0, 1, 2])
def test_pair_gemm_epilogue(act):
    from gpu_util import op_gemm, rel_l2
    M, N, K = 1539, 512, 256
    g = torch.Generator(device="cuda").manual_seed(act)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    of, ob = op_gemm(a, w, bias=bias, resid=resid, out_f32=True, out_bf16=True, act=act)
    ref = _ref(a, w, bias, resid, act)
    assert (of - ref).abs().max().item() < 3e-3
    assert rel_l2(ob.float(), ref) < 4e-3
    of2, ob2 = op_gemm(a, w, bias=bias, out_f32=False, out_bf16=True, act=act)  # bf16-only store path
    assert rel_l2(ob2.float(), _ref(a, w, bias, None, act)) < 4e-3


@pytest.mark.parametrize("k,stride", [(3, 2), (2, 2)])
def test_pair_gemm_conv_view(k, stride):
    """strided conv1d as implicit GEMM over overlapping channel-last rows, through the pair kernel"""
    from gpu_util import op_gemm
    C, T_in = 512, 4001
    T_out = (T_in - k) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(k)
    x = torch.randn(T_in + 8, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(C, C, k, device="cuda", generator=g) / (C * k) ** 0.5).bfloat16()
    wp = w.permute(0, 2, 1).contiguous().view(C, k * C)
    of, _ = op_gemm(x, wp, out_f32=True, out_bf16=False, a_row_stride=stride * C, k_inner=C, M=T_out)
    ref = torch.nn.functional.conv1d(x[:T_in].float().t().unsqueeze(0), w.float(), stride=stride)[0].t()
    err = (of - ref).abs().max().item()
    print(f"pair conv view k={k} s={stride}: max abs err {err:.3e}")
    assert err < 2e-3


@pytest.mark.parametrize("M,K,conv", [(256, 1024, False), (1000, 1536, True), (128 * 2 * 74 * 3 + 77, 1024, True), (40000, 512, False)])
def test_gemm_with_fused_row_layer_norm_gelu(M, K, conv):
    """svt_op_gemm_rowln: conv-as-GEMM -> LayerNorm(512) -> GELU in one kernel (HF:275-299) vs fp32 torch.  Covers a single
    unit, a ragged tail, more than two units per CTA pair (the statistics slots are double-buffered by unit parity) and the
    overlapping-row conv view (k = 2 or 3, stride 2)."""
    import torch.nn.functional as F
    from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr

    g = torch.Generator(device="cuda").manual_seed(M + K)
    N = 512
    if conv:  # rows of C_in = 512 channels, output row m reads input rows 2m .. 2m + taps - 1 (contiguous taps * C_in elements)
        cin, taps = 512, K // 512
        x = (torch.randn(2 * M + taps, cin, device="cuda", generator=g) * 0.7).bfloat16()
        a_buf, stride, k_inner = x, 2 * cin, cin
        rows = torch.arange(M) if M <= 4096 else torch.cat([torch.arange(0, 300), torch.arange(M // 2, M // 2 + 300),
                                                            torch.arange(M - 200, M)])
        a_ref = torch.stack([x[2 * int(m): 2 * int(m) + taps].reshape(-1) for m in rows]).float()
    else:
        a_buf = (torch.randn(M, K, device="cuda", generator=g) * 0.7).bfloat16()
        a_ref, stride, k_inner, rows = a_buf.float(), K, K, torch.arange(M)
    rows = rows.cuda()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5 * 2.0).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.3
    gamma = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    beta = 0.3 * torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    check(lib().svt_op_gemm_rowln(ptr(a_buf), stride, k_inner, ptr(w), ptr(bias), ptr(gamma), ptr(beta), 1e-5, 1, ptr(out), M, N, K,
                                  current_stream_ptr()))
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    pre = a_ref @ w.float().t() + bias
    ref = F.gelu(F.layer_norm(pre, (N,), gamma, beta, 1e-5))
    got = out[rows].float()
    err = float((got - ref).abs().max())
    rel = float((got - ref).norm() / ref.norm())
    print(f"fused conv+LN+GELU M={M} K={K} conv={conv}: max-abs {err:.3e} rel-L2 {rel:.3e}")
    assert rel < 6e-3 and err < 6e-2
    # and the two-kernel route gives the same rows up to the rounding of the statistics
    two = torch.empty_like(out)
    check(lib().svt_op_gemm(ptr(a_buf), stride, k_inner, ptr(w), ptr(bias), None, None, ptr(two), M, N, K, N, 0, current_stream_ptr()))
    check(lib().svt_op_layer_norm(None, ptr(two), ptr(gamma), ptr(beta), ptr(two), None, M, N, 1e-5, 1, current_stream_ptr()))
    torch.cuda.synchronize()
    assert float((two.float() - out.float()).abs().max()) < 4e-2

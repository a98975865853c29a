"""tcgen05 GEMM vs torch fp32 matmul on bf16-rounded operands (C ABI hook svt_op_gemm)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


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


def _diagnose(a, w, got, ref):
    """Print which structured hypothesis explains a wrong GEMM (partial K, row/col permutation...)."""
    M, K = a.shape
    print("got[:4,:6]", got[:4, :6].tolist())
    print("ref[:4,:6]", ref[:4, :6].tolist())
    ok = ((got - ref).abs() < 2e-3)
    print("fraction ok", ok.float().mean().item(), "nan frac", torch.isnan(got).float().mean().item())
    print("ok by row%128 block of 8:", ok.float().view(M, -1).mean(1)[:128].view(-1, 8).mean(1).tolist())
    print("ok by col block of 8:", ok.float().mean(0).view(-1, 8).mean(1)[:32].tolist())
    af, wf = a.float(), w.float()
    for k0 in range(0, min(K, 64), 16):
        part = af[:, k0:k0 + 16] @ wf[:, k0:k0 + 16].t()
        print(f"match with only k[{k0}:{k0+16}]:", ((got - part).abs() < 2e-3).float().mean().item())
    for kb in range(0, K, 64):
        part = af[:, kb:kb + 64] @ wf[:, kb:kb + 64].t()
        print(f"match with only kblock {kb//64}:", ((got - part).abs() < 2e-3).float().mean().item())


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 256, 64), (256, 256, 128), (128, 128, 256), (1000, 512, 1536),
                                    (333, 1024, 512), (4096, 3072, 1024), (777, 768, 3072), (2000, 4096, 1024)])
def test_gemm_plain(M, N, K):
    from gpu_util import op_gemm, rel_l2
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    of, ob = op_gemm(a, w, out_f32=True, out_bf16=True)
    ref = _ref(a, w, None, None, 0)
    err = (of - ref).abs().max().item()
    print(f"gemm {M}x{N}x{K}: max abs err fp32 out {err:.3e}, rel_l2 {rel_l2(of, ref):.3e}")
    if not (err < 2e-3):
        _diagnose(a, w, of, ref)
    assert torch.isfinite(of).all()
    assert err < 2e-3, err  # fp32 accumulate of exact bf16 products: only summation order differs
    assert rel_l2(ob.float(), ref) < 4e-3  # bf16 output rounding


@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_epilogue(act):
    from gpu_util import op_gemm, rel_l2
    M, N, K = 515, 512, 256
    g = torch.Generator(device="cuda").manual_seed(act)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    of, ob = op_gemm(a, w, bias=bias, resid=resid, out_f32=True, out_bf16=True, act=act)
    ref = _ref(a, w, bias, resid, act)
    assert (of - ref).abs().max().item() < 2e-3
    assert rel_l2(ob.float(), ref) < 4e-3
    # in-place residual (out aliases resid), as the encoder uses it
    from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr
    r2 = resid.clone()
    check(lib().svt_op_gemm(ptr(a), K, K, ptr(w), ptr(bias), ptr(r2), ptr(r2), None, M, N, K, N, act, current_stream_ptr()))
    torch.cuda.synchronize()
    assert (r2 - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("k,stride", [(3, 2), (2, 2)])
def test_gemm_conv_view(k, stride):
    """strided conv1d over channel-last activations as an overlapping-row GEMM view == F.conv1d."""
    from gpu_util import op_gemm
    C, Tin, B = 512, 400, 2
    Tout = Tin // stride
    g = torch.Generator(device="cuda").manual_seed(k)
    x = torch.zeros(B * Tin * C + 4 * C, device="cuda").bfloat16()  # slack for the last overlapping rows
    x[: B * Tin * C] = torch.randn(B * Tin * C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(C, C, k, device="cuda", generator=g) / (C * k) ** 0.5).bfloat16()  # (C_out, C_in, k)
    wp = w.permute(0, 2, 1).contiguous().view(C, k * C)  # [C_out][tap][C_in]
    M = B * Tout
    of, _ = op_gemm(x, wp, out_f32=True, out_bf16=False, a_row_stride=stride * C, k_inner=C, M=M)
    xc = x[: B * Tin * C].view(B, Tin, C).float().permute(0, 2, 1)  # (B, C, T)
    ref = torch.nn.functional.conv1d(xc, w.float(), stride=stride).permute(0, 2, 1)  # (B, Tvalid, C)
    Tv = ref.shape[1]
    got = of.view(B, Tout, C)[:, :Tv]
    err = (got - ref).abs().max().item()
    print(f"conv view k={k} s={stride}: max abs err {err:.3e}")
    assert err < 2e-3


@pytest.mark.parametrize("D,groups,T,Ta,B", [(1024, 16, 499, 500, 2), (1024, 16, 49, 50, 3), (768, 16, 130, 132, 2)])
@pytest.mark.parametrize("impl,taps", [(0, 128), (1, 128), (0, 16), (0, 30)])
def test_posconv(D, groups, T, Ta, B, impl, taps):
    """impl 0: tap-paired kernel mode (four taps per k-block, N = 256) when taps % 4 == 0; impl 1 / other tap counts: one tap
    per k-block (N = 64).  Odd-length kernels keep every output frame, even ones drop the last (HF SamePad)."""
    from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr
    check(lib().svt_set_option(b"gemm_impl", impl))
    Dg = D // groups
    g = torch.Generator(device="cuda").manual_seed(D + T)
    x = torch.randn(B, Ta, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(D, Dg, taps, device="cuda", generator=g) / (Dg * taps) ** 0.5).bfloat16().float()
    bias = torch.randn(D, device="cuda", generator=g)
    resid = torch.randn(B, Ta, D, device="cuda", generator=g)
    packed = torch.zeros(groups * taps * 64 * 64, device="cuda", dtype=torch.bfloat16)
    if Dg == 64:
        check(lib().svt_op_pack_posconv(ptr(w.contiguous()), D, groups, taps, ptr(packed), current_stream_ptr()))
    else:
        p = torch.zeros(groups, taps, 64, 64, device="cuda")
        p[:, :, :Dg, :Dg] = w.view(groups, Dg, Dg, taps).permute(0, 3, 1, 2)
        packed = p.bfloat16().contiguous().view(-1)
    out = resid.clone()
    try:
        check(lib().svt_op_posconv(ptr(x), ptr(packed), ptr(bias), ptr(out), ptr(out), B, Ta, T, D, groups, taps,
                                   current_stream_ptr()))
        torch.cuda.synchronize()
    finally:
        check(lib().svt_set_option(b"gemm_impl", 0))
    xv = x[:, :T].float().permute(0, 2, 1)
    conv = torch.nn.functional.conv1d(xv, w, bias, padding=taps // 2, groups=groups)[:, :, :T]
    ref = resid[:, :T] + torch.nn.functional.gelu(conv).permute(0, 2, 1)
    err = (out[:, :T] - ref).abs().max().item()
    print(f"posconv D={D} T={T}: max abs err {err:.3e}")
    assert err < 3e-3
    assert torch.equal(out[:, T:], resid[:, T:])  # padding rows untouched


@pytest.mark.parametrize("M,D,N,act,offset", [(515, 1024, 512, 0, 0.0), (2048, 1024, 3072, 0, 0.5), (1500, 1024, 4096, 1, -1.0),
                                              (300, 512, 256, 1, 3.0)])
def test_gemm_folded_layer_norm_chain(M, D, N, act, offset):
    """producer GEMM (+residual, row statistics, bf16 copy) -> consumer GEMM with the LayerNorm folded in ==
    LayerNorm(h) @ W^T + b of the same fp32 rows (option "ln_fold", encoder.cu)."""
    from gpu_util import rel_l2
    from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr
    g = torch.Generator(device="cuda").manual_seed(M + N)
    K0 = 256
    a = torch.randn(M, K0, device="cuda", generator=g).bfloat16()
    w0 = (torch.randn(D, K0, device="cuda", generator=g) / K0 ** 0.5).bfloat16()
    b0 = torch.randn(D, device="cuda", generator=g)
    resid = torch.randn(M, D, device="cuda", generator=g) * (1 + torch.rand(M, 1, device="cuda", generator=g) * 4) + offset
    gamma = 1 + 0.3 * torch.randn(D, device="cuda", generator=g)
    beta = 0.2 * torch.randn(D, device="cuda", generator=g)
    w1 = torch.randn(N, D, device="cuda", generator=g) / D ** 0.5
    b1 = torch.randn(N, device="cuda", generator=g)
    eps = 1e-5
    # producer
    h = resid.clone()
    hb = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    stats = torch.full((M, D // 128, 2), float("nan"), device="cuda")  # every slot has exactly one writer
    check(lib().svt_op_gemm_ln(ptr(a), ptr(w0), ptr(b0), None, None, 0.0, ptr(stats), ptr(h), ptr(h), ptr(hb), M, D, K0, 0,
                               current_stream_ptr()))
    h_ref = a.float() @ w0.float().t() + b0 + resid
    assert (h - h_ref).abs().max().item() < 2e-3
    assert torch.equal(hb, h.bfloat16())
    hs = h.view(M, D // 128, 128)
    assert torch.allclose(stats[:, :, 0], hs.sum(2), rtol=1e-4, atol=2e-3)
    assert torch.allclose(stats[:, :, 1], (hs * hs).sum(2), rtol=1e-4, atol=2e-3)
    # row_stats_cast: the first link of the chain
    hb2 = torch.empty_like(hb)
    stats2 = torch.empty_like(stats)
    check(lib().svt_op_row_stats_cast(ptr(h), M, D, ptr(hb2), ptr(stats2), current_stream_ptr()))
    assert torch.equal(hb2, hb) and torch.allclose(stats2, stats, rtol=1e-4, atol=2e-3)
    # consumer: weights folded on the host exactly like encoder_finalize does on the device
    wf = (w1 * gamma[None, :]).bfloat16()
    colsum = wf.float().sum(1)
    d = b1 + w1 @ beta
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    check(lib().svt_op_gemm_ln(ptr(hb), ptr(wf), ptr(d), ptr(colsum), ptr(stats), eps, None, None, None, ptr(out), M, N, D, act,
                               current_stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(h, (D,), gamma, beta, eps) @ w1.t() + b1
    # the unfused path: normalised rows rounded to bf16, bf16 weights
    unf = torch.nn.functional.layer_norm(h, (D,), gamma, beta, eps).bfloat16().float() @ w1.bfloat16().float().t() + b1
    if act == 1:
        ref, unf = torch.nn.functional.gelu(ref), torch.nn.functional.gelu(unf)
    e_fold, e_unf = rel_l2(out.float(), ref), rel_l2(unf.bfloat16().float(), ref)
    print(f"folded LN M={M} D={D} N={N} offset={offset}: rel-L2 folded {e_fold:.3e} vs separate-LN {e_unf:.3e}")
    assert torch.isfinite(out.float()).all()
    assert e_fold < 1e-2 and e_fold < 3 * e_unf + 1e-3


@pytest.mark.parametrize("M,N,K,act,resid", [(2016, 1024, 4096, 0, True), (2016, 3072, 1024, 0, False), (500, 4096, 1024, 1, False),
                                            (1000, 1024, 1024, 0, True)])
def test_gemm_kernel_choices_agree(M, N, K, act, resid):
    """Option "gemm_impl": the CTA-pair kernel, the one-CTA kernel with 256-column tiles and with 128-column tiles (the
    small-problem path) compute the same thing; 0 = automatic choice."""
    from gpu_util import rel_l2
    from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    r0 = torch.randn(M, N, device="cuda", generator=g) if resid else None
    ref = _ref(a, w, bias, r0, act)
    outs = {}
    try:
        for impl in (0, 1, 2, 3):
            check(lib().svt_set_option(b"gemm_impl", impl))
            if resid:
                o = r0.clone()
                check(lib().svt_op_gemm(ptr(a), K, K, ptr(w), ptr(bias), ptr(o), ptr(o), None, M, N, K, N, act, current_stream_ptr()))
            else:
                o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
                check(lib().svt_op_gemm(ptr(a), K, K, ptr(w), ptr(bias), None, None, ptr(o), M, N, K, N, act, current_stream_ptr()))
            torch.cuda.synchronize()
            outs[impl] = o.float()
    finally:
        lib().svt_set_option(b"gemm_impl", 0)
    for impl, o in outs.items():
        assert rel_l2(o, ref) < 4e-3, impl
    # same k-order of the accumulation in every tiling: the variants agree exactly
    assert torch.equal(outs[1], outs[2]) and torch.equal(outs[1], outs[3]) and torch.equal(outs[0], outs[1])

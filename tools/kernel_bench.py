#!/usr/bin/env python
"""Per-kernel timings of the hot path's shapes at BASELINE config 2 (B=64, 10-s clips), CUDA events, L2 flushed
between launches.  Prints one line per kernel: us/launch, TFLOP/s or GB/s.  Development tool (not bench.py)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=8, warm=2):
    ts = []
    for i in range(warm + n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def gemm(name, M, N, K, act=0, resid=False, f32=False, k_inner=None, row_stride=None, out=None):
    a = torch.randn(M + 8, K if row_stride is None else row_stride, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.zeros(N, device=dev)
    h = torch.randn(M, N, device=dev) if (resid or f32) else None
    ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if not (resid or f32) else None
    s = current_stream_ptr()
    rs = K if row_stride is None else row_stride
    ki = K if k_inner is None else k_inner
    fn = lambda: check(lib().svt_op_gemm(ptr(a), rs, ki, ptr(w), ptr(bias), ptr(h) if resid else None, ptr(h), ptr(ob), M, N, K, N, act, s))
    us = timeit(fn)
    tf = 2.0 * M * N * K / us / 1e6
    print(f"{name:34s} M={M:8d} N={N:5d} K={K:5d}  {us:9.1f} us  {tf:7.1f} TFLOP/s", flush=True)
    if out is not None:
        out[name] = {"us": us, "tflops": tf}


def attention(name, impl, clips, heads, T, Ta, dh, out=None):
    D = heads * dh
    qkv = torch.randn(clips * Ta, 3 * D, device=dev)
    qkv[:, :D] *= dh ** -0.5  # the packed q-projection carries the 1/sqrt(d_h) scale
    qkv = qkv.bfloat16()
    o = torch.zeros(clips * Ta, D, device=dev, dtype=torch.bfloat16)
    check(lib().svt_set_option(b"attention_impl", impl))
    s = current_stream_ptr()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    fn = lambda: check(lib().svt_op_attention(ptr(q), ptr(k), ptr(v), ptr(o), 3 * D, 3 * D, 3 * D, D, T, T, Ta, Ta, clips, heads, dh, s))
    us = timeit(fn)
    check(lib().svt_set_option(b"attention_impl", 0))
    tf = 4.0 * clips * heads * T * T * dh / us / 1e6
    print(f"{name:34s} clips={clips} heads={heads} T={T} dh={dh}  {us:9.1f} us  {tf:7.1f} TFLOP/s", flush=True)
    if out is not None:
        out[name] = {"us": us, "tflops": tf}


def layer_norm(name, rows, D, src_f32, gelu, out=None):
    x = torch.randn(rows, D, device=dev)
    if not src_f32:
        x = x.bfloat16()
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    y = torch.empty(rows, D, device=dev, dtype=torch.bfloat16)
    s = current_stream_ptr()
    fn = lambda: check(lib().svt_op_layer_norm(ptr(x) if src_f32 else None, None if src_f32 else ptr(x), ptr(g), ptr(b), ptr(y), None, rows, D, 1e-5, gelu, s))
    us = timeit(fn)
    gb = rows * D * ((4 if src_f32 else 2) + 2) / us / 1e3
    print(f"{name:34s} rows={rows:8d} D={D}  {us:9.1f} us  {gb:7.1f} GB/s", flush=True)
    if out is not None:
        out[name] = {"us": us, "gbs": gb}


def main():
    out = {}
    B, T, Ta = 64, 499, 500
    M = B * Ta
    which = sys.argv[1:] or ["gemm", "attn", "ln", "conv"]
    if "gemm" in which:
        gemm("qkv (bf16 out)", M, 3072, 1024, out=out)
        gemm("out-proj (+resid fp32)", M, 1024, 1024, resid=True, out=out)
        for nm, (N_, K_) in {"out-proj (+resid fp32, +bf16 copy, +row stats)": (1024, 1024), "ffn2 (+resid fp32, +bf16 copy, +row stats)": (1024, 4096)}.items():
            a = torch.randn(M, K_, device=dev).bfloat16()
            w = (torch.randn(N_, K_, device=dev) * 0.03).bfloat16()
            bias = torch.zeros(N_, device=dev)
            h = torch.randn(M, N_, device=dev)
            hb = torch.empty(M, N_, device=dev, dtype=torch.bfloat16)
            st = torch.empty(M, N_ // 128, 2, device=dev)
            s_ = current_stream_ptr()
            fn = lambda: check(lib().svt_op_gemm_ln(ptr(a), ptr(w), ptr(bias), None, None, 1e-5, ptr(st), ptr(h), ptr(h), ptr(hb), M, N_, K_, 0, s_))
            us = timeit(fn)
            print(f"{nm:34s} M={M:8d} N={N_:5d} K={K_:5d}  {us:9.1f} us  {2.0 * M * N_ * K_ / us / 1e6:7.1f} TFLOP/s", flush=True)
            out[nm] = {"us": us}
        gemm("ffn1 (gelu, bf16 out)", M, 4096, 1024, act=1, out=out)
        gemm("ffn2 (+resid fp32)", M, 1024, 4096, resid=True, out=out)
        if "variants" in which:
            gemm("ffn1 shape, no act", M, 4096, 1024, act=0, out=out)
            gemm("ffn1 shape, relu", M, 4096, 1024, act=2, out=out)
            gemm("ffn1 shape, gelu, K=2048", M, 4096, 2048, act=1, out=out)
            gemm("ffn1 shape, no act, K=2048", M, 4096, 2048, act=0, out=out)
            gemm("qkv shape K=256", M, 3072, 256, act=0, out=out)
            gemm("qkv shape gelu K=256", M, 3072, 256, act=1, out=out)
        gemm("proj 512->1024 (fp32 out)", M, 1024, 512, f32=True, out=out)
    if "smallm" in which:
        # kernel / tile choice for small row counts: CTA-pair (3) vs one-CTA 256-column (1) vs one-CTA 128-column tiles (2)
        for clips in (1, 2, 4, 8, 12, 16, 24, 32):
            m = clips * Ta
            for nm, (N_, K_, act_, res_) in {"qkv": (3072, 1024, 0, False), "out": (1024, 1024, 0, True),
                                             "ffn1": (4096, 1024, 1, False), "ffn2": (1024, 4096, 0, True)}.items():
                for impl in (3, 1, 2):
                    check(lib().svt_set_option(b"gemm_impl", impl))
                    gemm(f"{nm} clips={clips} impl={impl}", m, N_, K_, act=act_, resid=res_, out=out)
        check(lib().svt_set_option(b"gemm_impl", 0))
    if "conv" in which:
        gemm("conv1 k3s2 (implicit)", 1024000, 512, 1536, k_inner=512, row_stride=1024, out=out)
        gemm("conv2 k3s2 (implicit)", 512000, 512, 1536, k_inner=512, row_stride=1024, out=out)
        gemm("conv5 k2s2 (implicit)", 64000, 512, 1024, k_inner=512, row_stride=1024, out=out)
    if "conv0" in which or "conv" in which:
        Bc, Lc = 64, 160000
        wav = torch.randn(Bc, Lc, device=dev)
        w_kc = (torch.randn(10, 512, device=dev) * 0.4).contiguous()
        bias, gam, bet = torch.randn(512, device=dev) * 0.1, torch.ones(512, device=dev), torch.zeros(512, device=dev)
        ta = 32000
        o = torch.empty(Bc, ta, 512, device=dev, dtype=torch.bfloat16)
        scratch = torch.zeros(4, dtype=torch.float64, device=dev)
        s_ = current_stream_ptr()
        for impl in (1, 0):
            check(lib().svt_set_option(b"conv0_impl", impl))
            fn = lambda: check(lib().svt_op_conv0(ptr(wav), Bc, Lc, ptr(w_kc), ptr(bias), ptr(gam), ptr(bet), 1, ptr(o), ta, ptr(scratch), s_))
            us = timeit(fn)
            nm = "conv0 SIMT (+stats pass)" if impl == 1 else "conv0 mma.sync (+stats, +tables)"
            print(f"{nm:34s} B={Bc} L={Lc}  {us:9.1f} us  {Bc * ta * 512 * 2 / us / 1e3:7.1f} GB/s written", flush=True)
            out[nm] = {"us": us}
        check(lib().svt_set_option(b"conv0_impl", 0))
    if "conv" in which:
        for nm, Mc, Kc in (("conv1", 1024000, 1536), ("conv2", 512000, 1536), ("conv5", 64000, 1024)):
            a = torch.randn(2 * Mc + 8, 512, device=dev).bfloat16()
            w = (torch.randn(512, Kc, device=dev) / Kc ** 0.5).bfloat16()
            bias, gam, bet = torch.zeros(512, device=dev), torch.ones(512, device=dev), torch.zeros(512, device=dev)
            o = torch.empty(Mc, 512, device=dev, dtype=torch.bfloat16)
            s_ = current_stream_ptr()
            fn = lambda: check(lib().svt_op_gemm_rowln(ptr(a), 1024, 512, ptr(w), ptr(bias), ptr(gam), ptr(bet), 1e-5, 1, ptr(o), Mc, 512, Kc, s_))
            us = timeit(fn)

            def two():
                check(lib().svt_op_gemm(ptr(a), 1024, 512, ptr(w), ptr(bias), None, None, ptr(o), Mc, 512, Kc, 512, 0, s_))
                check(lib().svt_op_layer_norm(None, ptr(o), ptr(gam), ptr(bet), ptr(o), None, Mc, 512, 1e-5, 1, s_))
            us2 = timeit(two)
            print(f"{nm + ' + LN(512) + GELU fused':34s} M={Mc:8d} N=  512 K={Kc:5d}  {us:9.1f} us  {2.0 * Mc * 512 * Kc / us / 1e6:7.1f} TFLOP/s   "
                  f"(GEMM + LayerNorm kernel: {us2:.1f} us)", flush=True)
            out[nm + " fused ln"] = {"us": us, "us_two_kernels": us2}
    if "posconv" in which or "conv" in which:
        D, G, taps = 1024, 16, 128
        x = torch.randn(B * Ta, D, device=dev).bfloat16()
        wpk = (torch.randn(G * taps * 64, 64, device=dev) * 0.01).bfloat16()
        bias = torch.zeros(D, device=dev)
        h = torch.randn(B * Ta, D, device=dev)
        s_ = current_stream_ptr()
        fn = lambda: check(lib().svt_op_posconv(ptr(x), ptr(wpk), ptr(bias), ptr(h), ptr(h), B, Ta, T, D, G, taps, s_))
        us = timeit(fn)
        tf = 2.0 * B * T * D * (D // G) * taps / us / 1e6
        print(f"{'posconv k=128 g=16 (+gelu +resid)':34s} clips={B} T={T} D={D}  {us:9.1f} us  {tf:7.1f} TFLOP/s", flush=True)
        out["posconv"] = {"us": us, "tflops": tf}
    if "attn" in which:
        attention("attention mma.sync dh64", 1, B, 16, T, Ta, 64, out=out)
        attention("attention tcgen05 dh64", 2, B, 16, T, Ta, 64, out=out)
        attention("attention mma.sync dh128 (fusion)", 1, 8, 8, T, Ta, 128, out=out)
    if "ln" in which:
        layer_norm("LN 1024 fp32->bf16", M, 1024, True, 0, out=out)
        layer_norm("LN 512 bf16->bf16 +gelu (conv1)", 1024000, 512, False, 1, out=out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

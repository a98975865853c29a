"""Helpers for the GPU parity tests: call the C-ABI op hooks with torch tensors."""
import ctypes as C

import torch

from svt_speechbrain_b200._lib import check, current_stream_ptr, lib, ptr


def bf16(x):
    return x.to(torch.bfloat16)


def op_gemm(a, w, bias=None, resid=None, out_f32=False, out_bf16=True, act=0, a_row_stride=None, k_inner=None, M=None):
    """a: bf16 CUDA buffer (flat or 2-D), w: (N, K) bf16."""
    N, K = w.shape
    if M is None:
        M = a.shape[0]
    if a_row_stride is None:
        a_row_stride = K
    if k_inner is None:
        k_inner = K
    of = torch.full((M, N), float("nan"), dtype=torch.float32, device=a.device) if out_f32 else None
    ob = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=a.device) if out_bf16 else None
    check(lib().svt_op_gemm(ptr(a), a_row_stride, k_inner, ptr(w), ptr(bias), ptr(resid), ptr(of), ptr(ob), M, N, K, N,
                            act, current_stream_ptr()))
    torch.cuda.synchronize()
    return of, ob


def op_layer_norm(x, gamma, beta, eps=1e-5, gelu=0, out_f32=True, out_bf16=True):
    rows, D = x.shape
    yf = torch.empty(rows, D, dtype=torch.float32, device=x.device) if out_f32 else None
    yb = torch.empty(rows, D, dtype=torch.bfloat16, device=x.device) if out_bf16 else None
    xf = x if x.dtype == torch.float32 else None
    xb = x if x.dtype == torch.bfloat16 else None
    check(lib().svt_op_layer_norm(ptr(xf), ptr(xb), ptr(gamma), ptr(beta), ptr(yb), ptr(yf), rows, D, eps, gelu,
                                  current_stream_ptr()))
    torch.cuda.synchronize()
    return yf, yb


def op_attention(q, k, v, Tq, Tk, clips, heads, dh, q_clip_rows, k_clip_rows, ldo=None):
    """q/k/v: bf16 2-D views (rows, ld) possibly column-sliced views of one buffer (stride(0) = ld)."""
    D = heads * dh
    ldo = D if ldo is None else ldo
    o = torch.zeros(clips * q_clip_rows, ldo, dtype=torch.bfloat16, device=q.device)
    check(lib().svt_op_attention(ptr(q), ptr(k), ptr(v), ptr(o), q.stride(0), k.stride(0), v.stride(0), ldo, Tq, Tk,
                                 q_clip_rows, k_clip_rows, clips, heads, dh, current_stream_ptr()))
    torch.cuda.synchronize()
    return o


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))

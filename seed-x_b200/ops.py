"""Thin Python wrappers over the C ABI (include/seedx.h).

Tensors are torch CUDA tensors used purely as device-memory handles: every wrapper passes raw pointers,
sizes and the current CUDA stream to libseedx.so; no torch compute op runs in here.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_SILU, F16, F32, GemmArgs, check, lib  # noqa: F401


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _dt(t):
    if t.dtype == torch.float16:
        return F16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.SeedxError("seedx ops need CUDA tensors: there is no CPU fallback path")


def gemm(a, w, out=None, *, bias=None, bias_m=None, bias_g=None, bias_g_rows=0, residual=None, res_row_mod=0,
         act=ACT_NONE, gated=False, alpha=1.0, out_dtype=torch.float16, tile_n=0):
    """out[b,m,n] = epi(alpha * a[b,m,:] . w[(b,)n,:]).

    a: fp16 [M,K] or [B,M,K] (last dim contiguous); w: fp16 [N,K] or [B,N,K]; returns/updates out [.., M, N_out].
    """
    _require_cuda(a, w, out, bias, residual)
    assert a.dtype == torch.float16 and w.dtype == torch.float16
    assert a.stride(-1) == 1 and w.stride(-1) == 1
    batched = a.dim() == 3
    B = a.shape[0] if batched else 1
    M, K = a.shape[-2], a.shape[-1]
    N = w.shape[-2]
    assert w.shape[-1] == K, (a.shape, w.shape)
    n_out = N // 2 if gated else N
    if out is None:
        shape = (B, M, n_out) if batched else (M, n_out)
        out = torch.empty(shape, device=a.device, dtype=out_dtype)
    assert out.stride(-1) == 1 and out.shape[-1] == n_out and out.shape[-2] == M
    g = GemmArgs()
    g.A, g.B, g.D = a.data_ptr(), w.data_ptr(), out.data_ptr()
    g.M, g.N, g.K, g.batch = M, N, K, B
    g.lda, g.ldb, g.ldd = a.stride(-2), w.stride(-2), out.stride(-2)
    g.strideA = a.stride(0) if batched else 0
    g.strideB = w.stride(0) if (w.dim() == 3) else 0
    g.strideD = out.stride(0) if batched else 0
    g.alpha = alpha
    for name, t in (("bias_n", bias), ("bias_m", bias_m), ("bias_g", bias_g)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous()
            setattr(g, name, t.data_ptr())
    g.bias_g_rows = bias_g_rows
    if residual is not None:
        assert residual.stride(-1) == 1
        g.residual = residual.data_ptr()
        g.residual_dtype = _dt(residual)
        g.ldr = residual.stride(-2)
        g.strideR = residual.stride(0) if (batched and residual.dim() == 3) else 0
        g.res_row_mod = res_row_mod
    g.act, g.gated, g.out_dtype = act, int(gated), _dt(out)
    g.tile_n = tile_n
    check(lib().seedx_gemm_f16(C.byref(g), _stream()), "seedx_gemm_f16")
    return out


def conv2d_nhwc(x, w, out=None, *, taps=3, bias=None, bias_g=None, residual=None, act=ACT_NONE,
                out_dtype=torch.float16, tile_n=0):
    """Stride-1 'same' convolution as an implicit GEMM.  x: fp16 NHWC [N,H,W,C]; w: fp16 [Cout, taps*taps*roundup(C,64)]
    (k = (kh*taps+kw)*Cpad + c); out: NHWC [N,H,W,Cout].  bias_g: fp32 [N, Cout] added per image."""
    _require_cuda(x, w, out, bias, residual)
    assert x.dtype == torch.float16 and x.is_contiguous() and w.is_contiguous()
    n, h, wd, c = x.shape
    cout, K = w.shape
    if out is None:
        out = torch.empty((n, h, wd, cout), device=x.device, dtype=out_dtype)
    assert out.is_contiguous()
    g = GemmArgs()
    g.A, g.B, g.D = x.data_ptr(), w.data_ptr(), out.data_ptr()
    g.M, g.N, g.K, g.batch = n * h * wd, cout, K, 1
    g.lda, g.ldb, g.ldd = c, K, cout
    g.alpha = 1.0
    if bias is not None:
        assert bias.dtype == torch.float32
        g.bias_n = bias.data_ptr()
    if bias_g is not None:
        assert bias_g.dtype == torch.float32 and bias_g.is_contiguous()
        g.bias_g = bias_g.data_ptr()
        g.bias_g_rows = h * wd
    if residual is not None:
        assert residual.is_contiguous()
        g.residual = residual.data_ptr()
        g.residual_dtype = _dt(residual)
        g.ldr = cout
    g.act, g.out_dtype = act, _dt(out)
    g.conv_taps_h = g.conv_taps_w = taps
    g.conv_n, g.conv_h, g.conv_w, g.conv_c = n, h, wd, c
    g.tile_n = tile_n
    check(lib().seedx_gemm_f16(C.byref(g), _stream()), "seedx_gemm_f16(conv)")
    return out

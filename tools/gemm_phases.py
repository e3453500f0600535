"""Per-CTA phase timeline of one GEMM launch (stamps from seedx_gemm_set_debug): where do the fixed ~10 us of a short-K GEMM go?"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
from seedx_b200._lib import lib
shapes = [(8192, 1280, 1280, "plain"), (8192, 1280, 1280, "res"), (32768, 640, 640, "res"), (8192, 10240, 1280, "geglu"), (8192, 1280, 5120, "res")]
names = ["entry", "prologue done", "pdl_wait done", "first operands landed", "first tile acc done", "last tile acc done", "epilogue drained", "exit"]
for (M, N, K, kind) in shapes:
    A = torch.randn(M, K, device="cuda").half(); W = torch.randn(N, K, device="cuda").half() * 0.03
    O = torch.randn(M, N // 2 if kind == "geglu" else N, device="cuda").half(); bias = torch.randn(N, device="cuda")
    def run():
        if kind == "plain": ops.gemm(A, W, out=O)
        elif kind == "res": ops.gemm(A, W, out=O, bias=bias, residual=O)
        else: ops.gemm(A, W, out=O, bias=bias, act=ops.ACT_GELU, gated=True)
    for _ in range(3): run()
    dbg = torch.zeros(148 * 8, device="cuda", dtype=torch.int64)
    lib().seedx_gemm_set_debug(C.c_void_p(dbg.data_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    lib().seedx_gemm_set_debug(C.c_void_p(0))
    d = dbg.view(148, 8).cpu().double()
    t0 = d[:, 0].min()
    d = (d - t0) / 1e3
    print(f"== M={M} N={N} K={K} {kind}: event time {e0.elapsed_time(e1)*1e3:.1f} us; first entry -> last exit {d[:, 7].max():.1f} us")
    for i, n in enumerate(names):
        col = d[:, i][d[:, i] > -1e6]
        print(f"   {n:24s} min {col.min():7.2f}  median {col.median():7.2f}  max {col.max():7.2f} us")

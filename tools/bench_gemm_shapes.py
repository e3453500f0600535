"""Per-launch time of the UNet's hot GEMM / conv shapes (8 samples) for the epilogue variants of round 2: plain, folded LayerNorm, output statistics,
weights prefetched before the PDL wait or not (dynamic_b), stream-K off / forced.  20 back-to-back launches inside one CUDA graph, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
from seedx_b200._lib import lib
torch.manual_seed(0)
dev = "cuda"


def timeit(f, n=20):
    f(); f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * n) * 1e3


M = 8192
h = torch.randn(M, 1280, device=dev).half()
st = ops.row_stats(h, 1e-5)
rows = []
for name, N, K, kind in [("o1/o2 (bias+residual)", 1280, 1280, "res"), ("q2", 1280, 1280, "ln"), ("qkv", 3840, 1280, "ln"), ("ff1 GEGLU", 10240, 1280, "lngate"),
                         ("ff2 (bias+residual)", 1280, 5120, "res")]:
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) * 0.03).half()
    bias, cs = torch.randn(N, device=dev), torch.randn(N, device=dev)
    out = torch.randn(M, N // 2 if kind == "lngate" else N, device=dev).half()
    rp = torch.empty(N // 32, M, 2, device=dev)
    cp = torch.empty(M // 32, N, 2, device=dev)
    gate = dict(act=ops.ACT_GELU, gated=True) if kind == "lngate" else {}
    res = dict(residual=out) if kind == "res" else {}
    for lab, kw in [("plain no bias", dict()), ("bias", dict(bias=bias)), ("bias dyn_b", dict(bias=bias, dynamic_b=True))] + \
                   ([("bias+LN", dict(bias=bias, ln=(st, cs)))] if kind.startswith("ln") else []) + \
                   ([("bias+rowpart", dict(bias=bias, row_part=rp)), ("bias+colpart", dict(bias=bias, col_part=cp))] if kind == "res" else []):
        for sk in (0, 2):
            lib().seedx_gemm_set_stream_k(sk)
            us = timeit(lambda: ops.gemm(a, w, out=out, **gate, **res, **kw))
            rows.append((name, N, K, lab, sk, us, 2.0 * M * N * K / us / 1e6))
    lib().seedx_gemm_set_stream_k(1)
x = torch.randn(8, 32, 32, 1280, device=dev).half(); wc = (torch.randn(1280, 9 * 1280, device=dev) * 0.01).half(); r = torch.randn(8, 32, 32, 1280, device=dev).half()
bias = torch.randn(1280, device=dev); cp = torch.empty(8 * 1024 // 32, 1280, 2, device=dev); o = torch.empty_like(r)
for lab, kw in [("conv3x3 1280 @32^2 bias+res", dict()), ("... +colpart", dict(col_part=cp))]:
    for sk in (0, 1, 2):
        lib().seedx_gemm_set_stream_k(sk)
        us = timeit(lambda: ops.conv2d_nhwc(x, wc, out=o, bias=bias, residual=r, **kw))
        rows.append((lab, 1280, 11520, "", sk, us, 2.0 * 8192 * 1280 * 11520 / us / 1e6))
lib().seedx_gemm_set_stream_k(1)
print(f"{'shape':30s} {'N':>6s} {'K':>6s} {'epilogue':16s} sk {'us':>8s} {'TF/s':>7s}")
for r_ in rows:
    print(f"{r_[0]:30s} {r_[1]:6d} {r_[2]:6d} {r_[3]:16s} {r_[4]:2d} {r_[5]:8.1f} {r_[6]:7.0f}")

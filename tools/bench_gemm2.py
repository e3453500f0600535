"""micro-benchmark of one GEMM shape across tile widths / CTA-pair mode / epilogue variants"""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
from seedx_b200._lib import lib
M, N, K = (int(v) for v in os.environ.get("SHAPE", "8192,1280,1280").split(","))
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda", dtype=torch.float32)
a = torch.randn(M, K, device="cuda").half(); w = torch.randn(N, K, device="cuda").half()
res16 = torch.randn(M, N, device="cuda").half(); res32 = torch.randn(M, N, device="cuda")
bias = torch.randn(N, device="cuda")
def t(fn, reps=8):
    fn(); fn(); ts = []
    for _ in range(reps):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for cl in (0, 2):
    lib().seedx_gemm_set_cluster(cl)
    for name, kw, od in (("plain f16", {}, torch.float16), ("bias+res16 f16", dict(bias=bias, residual=res16), torch.float16),
                         ("bias+res32 f32", dict(bias=bias, residual=res32), torch.float32)):
        out = torch.empty(M, N, device="cuda", dtype=od)
        row = []
        for bn in (0, 128, 160, 192, 256):
            ms = t(lambda: ops.gemm(a, w, out=out, tile_n=bn, **kw))
            row.append(f"bn={bn}: {ms*1e3:6.1f}us {2*M*N*K/ms/1e9:6.0f}TF")
        print(f"pair={cl} {name:16s} " + " | ".join(row), flush=True)

"""back-to-back launches (no host sync between them) of one GEMM shape: removes CPU launch gaps from the per-launch time"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
from seedx_b200._lib import lib
shapes = [(8192, 1280, 1280), (8192, 1280, 5120), (8192, 3840, 1280), (8192, 10240, 1280), (32768, 640, 640), (32768, 5120, 640), (8192, 8192, 8192)]
R = 8   # rotate over R operand sets so each launch streams its operands from HBM/L2 like in the model
for (M, N, K) in shapes:
    As = [torch.randn(M, K, device="cuda").half() for _ in range(R)]
    Ws = [torch.randn(N, K, device="cuda").half() for _ in range(R)]
    outs = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(R)]
    res = [torch.randn(M, N, device="cuda").half() for _ in range(R)]
    row = []
    for name, use_res in (("plain", False), ("res16", True)):
        for cl in (0, 2):
            lib().seedx_gemm_set_cluster(cl)
            def run(n):
                for i in range(n):
                    j = i % R
                    ops.gemm(As[j], Ws[j], out=outs[j], residual=res[j] if use_res else None)
            run(R); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run(32)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 64 * 1e3
            row.append(f"{name} pair={cl}: {us:7.1f}us {2*M*N*K/us/1e6:6.0f}TF")
    print(f"M={M} N={N} K={K}: " + " | ".join(row), flush=True)
lib().seedx_gemm_set_cluster(1)

"""UNet cross-attention shapes (64 context tokens): mma.sync kernel vs the tcgen05 kernels with a half-padded 128-key tile (SEEDX_FA_MIN_SK=64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
from seedx_b200._lib import lib
for (B, H, Sq, Sk, D) in [(8, 20, 1024, 64, 64), (8, 10, 4096, 64, 64), (2, 20, 1024, 64, 64)]:
    q = torch.randn(B, Sq, H, D, device="cuda").half(); k = torch.randn(B, Sk, H, D, device="cuda").half(); v = torch.randn(B, Sk, H, D, device="cuda").half()
    o = torch.empty(B, Sq, H, D, device="cuda", dtype=torch.float16)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float().permute(0, 2, 1, 3), k.float().permute(0, 2, 1, 3), v.float().permute(0, 2, 1, 3))
    res = []
    for impl in (0, 2, 1):
        lib().seedx_attention_set_impl(impl)
        f = lambda: ops.attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), o.permute(0, 2, 1, 3), scale=D ** -0.5)
        f(); f(); torch.cuda.synchronize()
        err = ((o.permute(0, 2, 1, 3).float() - ref).norm() / ref.norm()).item()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): f()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        res.append(f"{ {3:'pp',2:'tc',1:'mma'}[lib().seedx_attention_last_impl()] }: {us:7.1f} us rel {err:.1e}")
    lib().seedx_attention_set_impl(0)
    print(f"B={B} H={H} Sq={Sq} Sk={Sk} D={D}: " + " | ".join(res), flush=True)

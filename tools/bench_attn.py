"""attention kernel throughput: tcgen05 vs mma.sync on the UNet / ViT / LLM shapes"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
from seedx_b200._lib import lib
for (B, H, Sq, Sk, D, causal) in [(2, 10, 4096, 4096, 64, False), (8, 10, 4096, 4096, 64, False), (2, 20, 1024, 1024, 64, False), (8, 20, 1024, 1024, 64, False),
                                  (8, 20, 1024, 64, 64, False), (8, 16, 1024, 1024, 104, False), (1, 40, 175, 175, 128, True), (1, 40, 2048, 2048, 128, True)]:
    q = torch.randn(B, Sq, H, D, device="cuda").half(); k = torch.randn(B, Sk, H, D, device="cuda").half(); v = torch.randn(B, Sk, H, D, device="cuda").half()
    o = torch.empty(B, Sq, H, D, device="cuda", dtype=torch.float16)
    res = []
    for impl in (0, 2):
        lib().seedx_attention_set_impl(impl)
        f = lambda: ops.attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), o.permute(0, 2, 1, 3), scale=D ** -0.5, causal=causal)
        f(); f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 4.0 * B * H * Sq * Sk * D * (0.5 if causal else 1.0)
        res.append(f"{ {3:'pp',2:'tc',1:'mma'}[lib().seedx_attention_last_impl()] }: {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF/s")
    lib().seedx_attention_set_impl(0)
    print(f"B={B} H={H} Sq={Sq} Sk={Sk} D={D} causal={causal}: " + " | ".join(res), flush=True)

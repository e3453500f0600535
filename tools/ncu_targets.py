"""Launch a handful of representative kernels once each (after a warm-up) so `ncu --set full` can capture them.
usage: python tools/ncu_targets.py [conv|gemm|attn|gemv|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
Be = 8
if which in ("conv", "all"):
    x = torch.randn(Be, 64, 64, 640, device="cuda").half(); w = torch.randn(640, 9 * 640, device="cuda").half()
    out = torch.empty(Be, 64, 64, 640, device="cuda", dtype=torch.float16)
    for _ in range(3): ops.conv2d_nhwc(x, w, out=out)
if which in ("gemm", "all"):
    a = torch.randn(8192, 8192, device="cuda").half(); b = torch.randn(8192, 8192, device="cuda").half()
    o = torch.empty(8192, 8192, device="cuda", dtype=torch.float16)
    for _ in range(3): ops.gemm(a, b, out=o)
    a = torch.randn(Be * 1024, 1280, device="cuda").half(); b = torch.randn(10240, 1280, device="cuda").half()
    o = torch.empty(Be * 1024, 5120, device="cuda", dtype=torch.float16)
    for _ in range(3): ops.gemm(a, b, out=o, act=ops.ACT_GELU, gated=True)
if which in ("attn", "all"):
    q = torch.randn(Be, 4096, 10, 64, device="cuda").half(); k = torch.randn_like(q); v = torch.randn_like(q); o = torch.empty_like(q)
    for _ in range(3): ops.attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), o.permute(0, 2, 1, 3), scale=0.125)
if which in ("gemv", "all"):
    W = torch.randn(27648, 5120, device="cuda").half(); x = torch.randn(4, 5120, device="cuda"); o = torch.empty(4, 13824, device="cuda")
    rw = torch.ones(5120, device="cuda")
    for _ in range(3): ops.gemv(W, x, o, rms_w=rw, gated=True)
torch.cuda.synchronize()
print("done")

"""Steady-state HBM bandwidth of the decode GEMV per LLaMA-13B shape (distinct weight buffers cycled so the 126 MB L2 cannot hold them), and of one
layer's chain with and without the cache-attention kernel — separates what the GEMV loses inside a launch from what the 161 launch boundaries cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops

dev = "cuda"
NB = int(os.environ.get("NB", "1"))
H, FF = 5120, 13824


def timed(fn, reps=5):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


shapes = {"wqkv": (3 * H, H, True, False, False), "wo": (H, H, False, True, False), "wgu": (2 * FF, H, True, False, True), "wdown": (H, FF, False, True, False)}
COPIES = 6
Ws = {k: [torch.randn(n, kk, device=dev).half() * 0.02 for _ in range(COPIES)] for k, (n, kk, *_r) in shapes.items()}
ln = torch.ones(H, device=dev)
for name, (n, k, rms, res, gated) in shapes.items():
    x = torch.randn(NB, k, device=dev)
    o = torch.empty(NB, n // 2 if gated else n, device=dev)
    r = torch.randn(NB, n, device=dev) if res else None
    CALLS = 24

    def run():
        for i in range(CALLS):
            ops.gemv(Ws[name][i % COPIES], x, o, rms_w=ln if rms else None, eps=1e-5, residual=r, gated=gated)
    ms = timed(run) / CALLS
    mb = n * k * 2 / 1e6
    print(f"{name:6s} N={n:6d} K={k:6d}: {ms*1e3:7.1f} us  {mb/ms/1e3:6.2f} TB/s ({mb/ms/1e3/6.4852*100:.0f}% of measured peak)", flush=True)

# one layer's GEMV chain (no attention), 6 layers' worth of distinct weights
xa, xb = torch.randn(NB, H, device=dev), torch.randn(NB, H, device=dev)
qkv, att, g1 = torch.empty(NB, 3 * H, device=dev), torch.randn(NB, H, device=dev), torch.empty(NB, FF, device=dev)


def chain():
    for i in range(COPIES):
        ops.gemv(Ws["wqkv"][i], xa, qkv, rms_w=ln, eps=1e-5)
        ops.gemv(Ws["wo"][i], att, xb, residual=xa)
        ops.gemv(Ws["wgu"][i], xb, g1, rms_w=ln, eps=1e-5, gated=True)
        ops.gemv(Ws["wdown"][i], g1, xa, residual=xb)
ms = timed(chain) / COPIES
mb = sum(n * k * 2 for (n, k, *_r) in shapes.values()) / 1e6
print(f"layer chain (4 GEMVs, no attention): {ms*1e3:7.1f} us/layer  {mb/ms/1e3:6.2f} TB/s ({mb/ms/1e3/6.4852*100:.0f}%) -> x40 = {ms*40:.2f} ms/token", flush=True)

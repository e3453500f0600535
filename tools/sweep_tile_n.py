"""Does the wave/cycle model of choose_tile_n() (csrc/gemm_tc.cu) pick the fastest accumulator width?  Times every hot GEMM / conv shape of the UNet forward
(8 samples) with tile_n = auto and forced 128..256 (20 back-to-back launches in one CUDA graph, CUDA events) and prints the winner per shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
torch.manual_seed(0)
dev = "cuda"
TILES = [0, 128, 160, 192, 224, 256]


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


def report(name, count, fl, runs):
    best = min(runs, key=lambda kv: kv[1])
    auto = runs[0][1]
    line = " ".join(f"{t if t else 'auto':>4}:{us:7.1f}" for t, us in runs)
    flag = "" if auto <= best[1] * 1.02 else f"   <-- {best[0]} is {100 * (auto - best[1]) / auto:.1f} % faster: {count} launches x {auto - best[1]:.1f} us = {count * (auto - best[1]) / 1e3:.2f} ms/forward"
    print(f"{name:34s} x{count:3d} {fl / auto / 1e6:6.0f} TF/s | {line}{flag}", flush=True)


for (M, c, depth_blocks) in ((8192, 1280, 60), (32768, 640, 10)):
    h = torch.randn(M, c, device=dev).half()
    st = ops.row_stats(h, 1e-5)
    for name, N, K, kind, count in [("attn out / xattn out (res+stats)", c, c, "res", 2 * depth_blocks), ("xattn q (LN)", c, c, "ln", depth_blocks),
                                    ("qkv (LN)", 3 * c, c, "ln", depth_blocks), ("ff1 GEGLU (LN)", 8 * c, c, "lngate", depth_blocks),
                                    ("ff2 (res+stats)", c, 4 * c, "res", depth_blocks)]:
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) * 0.03).half()
        bias, cs = torch.randn(N, device=dev), torch.randn(N, device=dev)
        out = torch.randn(M, N // 2 if kind == "lngate" else N, device=dev).half()
        kw = dict(bias=bias)
        if kind.startswith("ln"):
            kw["ln"] = (st, cs)
        if kind == "lngate":
            kw.update(act=ops.ACT_GELU, gated=True)
        if kind == "res":
            rp = torch.empty(N // 32, M, 2, device=dev)
            kw.update(residual=out, row_part=rp, row_stats=(torch.empty(M, 2, device=dev), torch.zeros(M // 32, device=dev, dtype=torch.int32), 1e-5))
        runs = []
        for t in TILES:
            try:
                runs.append((t, timeit(lambda: ops.gemm(a, w, out=out, tile_n=t, **kw))))
            except Exception as e:  # a width the epilogue variant does not take
                runs.append((t, float("inf")))
        report(f"M={M} {name}", count, 2.0 * M * N * K, runs)

for (hw, cin, cout, count, res) in [(128, 320, 320, 8, True), (128, 960, 320, 1, False), (128, 640, 320, 4, False), (64, 320, 640, 1, False), (64, 640, 640, 7, True),
                                    (64, 1920, 640, 1, False), (64, 1280, 640, 1, False), (64, 960, 640, 1, False), (32, 640, 1280, 1, False),
                                    (32, 1280, 1280, 11, True), (32, 2560, 1280, 2, False), (32, 1920, 1280, 1, False)]:
    x = torch.randn(8, hw, hw, cin, device=dev).half()
    wc = (torch.randn(cout, 9 * cin, device=dev) * 0.01).half()
    r = torch.randn(8, hw, hw, cout, device=dev).half()
    o = torch.empty_like(r)
    bias = torch.randn(cout, device=dev)
    cp = torch.empty(8 * hw * hw // 32, cout, 2, device=dev) if hw * hw <= 65536 else None
    runs = []
    for t in TILES:
        try:
            runs.append((t, timeit(lambda: ops.conv2d_nhwc(x, wc, out=o, bias=bias, residual=r if res else None, col_part=cp, tile_n=t), n=10)))
        except Exception as e:
            runs.append((t, float("inf")))
    report(f"conv3x3 {cin}->{cout} @{hw}^2", count, 2.0 * 8 * hw * hw * cout * 9 * cin, runs)

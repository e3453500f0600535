"""Asymptotic stream rate of the decode GEMV: the same kernel on matrices of growing N (K = 5120), distinct buffers, in one graph.  time(N) = a + bytes / BW
separates the per-launch constant a from the steady-state bandwidth."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
K = 5120
res = []
for N in (5120, 10240, 20480, 40960, 81920, 163840):
    copies = max(2, min(8, int(400e6 // (N * K * 2)) + 1))
    Ws = [torch.randn(N, K, device="cuda").half() * 0.02 for _ in range(copies)]
    x = torch.randn(1, K, device="cuda"); o = torch.empty(1, N, device="cuda")
    calls = 16
    def run():
        for i in range(calls):
            ops.gemv(Ws[i % copies], x, o)
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): run()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s): run()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 / calls * 1e3
    mb = N * K * 2 / 1e6
    res.append((mb, us))
    print(f"N={N:7d}: {mb:8.1f} MB  {us:8.1f} us  {mb/us:6.2f} TB/s", flush=True)
    del Ws
(m0, t0), (m1, t1) = res[-3], res[-1]
bw = (m1 - m0) / (t1 - t0)
print(f"slope between N=40960 and N=163840: {bw:.2f} TB/s (MB/us); constant a = {t1 - m1 / bw:.1f} us per launch")

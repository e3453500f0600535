"""3x3 convs over image rows of >= 128 pixels: row mode (one 130-pixel A box per (kh, chunk) shared by the three kw taps) vs the nine-tile form.
The mode is read once per process (SEEDX_CONV_ROW=0/1): run twice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
torch.manual_seed(0)
dev = "cuda"


def timeit(f, n=10):
    f(); f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * n) * 1e3


print("SEEDX_CONV_ROW =", os.environ.get("SEEDX_CONV_ROW", "1 (default)"))
for (nimg, hw, cin, cout, res, label) in [(8, 128, 320, 320, True, "UNet 320->320 @128^2 +res"), (8, 128, 320, 320, False, "UNet 320->320 @128^2"),
                                          (8, 128, 640, 320, False, "UNet 640->320 @128^2"), (8, 128, 960, 320, False, "UNet 960->320 @128^2"),
                                          (1, 256, 512, 512, True, "VAE 512->512 @256^2 +res"), (1, 512, 512, 256, False, "VAE 512->256 @512^2"),
                                          (1, 512, 256, 256, True, "VAE 256->256 @512^2 +res"), (1, 1024, 256, 128, False, "VAE 256->128 @1024^2"),
                                          (1, 1024, 128, 128, True, "VAE 128->128 @1024^2 +res")]:
    x = torch.randn(nimg, hw, hw, cin, device=dev).half()
    wc = (torch.randn(cout, 9 * cin, device=dev) * 0.01).half()
    r = torch.randn(nimg, hw, hw, cout, device=dev).half()
    o = torch.empty_like(r)
    bias = torch.randn(cout, device=dev)
    us = timeit(lambda: ops.conv2d_nhwc(x, wc, out=o, bias=bias, residual=r if res else None))
    fl = 2.0 * nimg * hw * hw * cout * 9 * cin
    print(f"{label:30s} {us:8.1f} us  {fl / us / 1e6:6.0f} TF/s", flush=True)

"""Per-shape throughput of seedx_gemm_f16 (plain + conv) on the shapes of the three stages; CUDA events, L2 flushed between reps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops

flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda", dtype=torch.float32)


def timeit(fn, reps=6):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def run_gemm(name, M, N, K, tiles=(0,), **kw):
    a = torch.randn(M, K, device="cuda").half()
    w = torch.randn(N, K, device="cuda").half()
    out = torch.empty(M, N // 2 if kw.get("gated") else N, device="cuda", dtype=torch.float16)
    res = []
    for t in tiles:
        ms = timeit(lambda: ops.gemm(a, w, out=out, tile_n=t, **kw))
        res.append(f"bn={t}: {ms*1e3:8.1f} us {2*M*N*K/ms/1e9:7.1f} TF/s")
    print(f"{name:34s} M={M:6d} N={N:6d} K={K:6d}  " + " | ".join(res), flush=True)


def run_conv(name, n, h, w, c, cout, tiles=(0,)):
    x = torch.randn(n, h, w, c, device="cuda").half()
    cpad = (c + 63) // 64 * 64
    wt = torch.randn(cout, 9 * cpad, device="cuda").half()
    out = torch.empty(n, h, w, cout, device="cuda", dtype=torch.float16)
    res = []
    for t in tiles:
        ms = timeit(lambda: ops.conv2d_nhwc(x, wt, out=out, tile_n=t))
        res.append(f"bn={t}: {ms*1e3:8.1f} us {2*n*h*w*9*c*cout/ms/1e9:7.1f} TF/s")
    print(f"{name:34s} {n}x{h}x{w}x{c}->{cout}  " + " | ".join(res), flush=True)


T = tuple(int(t) for t in os.environ.get("TILES", "0,128,256").split(","))
from seedx_b200._lib import lib
lib().seedx_gemm_set_cluster(int(os.environ.get("CLUSTER", "1")))
run_gemm("square 8192", 8192, 8192, 8192, T)
run_gemm("square 4096", 4096, 4096, 4096, T)
Be = 2
for (lvl, S, C) in (("64^2", 4096, 640), ("32^2", 1024, 1280)):
    M = Be * S
    run_gemm(f"unet {lvl} qkv", M, 3 * C, C, T)
    run_gemm(f"unet {lvl} out/proj", M, C, C, T)
    run_gemm(f"unet {lvl} ff1 geglu", M, 8 * C, C, T, gated=True, act=ops.ACT_GELU)
    run_gemm(f"unet {lvl} ff2", M, C, 4 * C, T)
run_conv("unet conv 128^2 320->320", Be, 128, 128, 320, 320, T)
run_conv("unet conv 64^2 640->640", Be, 64, 64, 640, 640, T)
run_conv("unet conv 32^2 1280->1280", Be, 32, 32, 1280, 1280, T)
run_conv("unet conv 32^2 2560->1280", Be, 32, 32, 2560, 1280, T)
run_conv("unet conv 64^2 1920->640", Be, 64, 64, 1920, 640, T)
run_conv("unet conv 128^2 960->320", Be, 128, 128, 960, 320, T)
run_conv("vae conv 1024^2 128->128", 1, 1024, 1024, 128, 128, T)
run_conv("vae conv 512^2 256->256", 1, 512, 512, 256, 256, T)
for N_img in (2,):
    M = N_img * 1024
    run_gemm("vit qkv", M, 4992, 1664, T)
    run_gemm("vit out", M, 1664, 1664, T)
    run_gemm("vit fc gelu", M, 8192, 1664, T, act=ops.ACT_GELU)
    run_gemm("vit proj", M, 1664, 8192, T)
run_gemm("llm prefill qkv P=174", 174, 15360, 5120, T)
run_gemm("llm prefill gate/up P=174", 174, 27648, 5120, T, gated=True, act=ops.ACT_SILU)
run_gemm("llm prefill down P=174", 174, 5120, 13824, T)

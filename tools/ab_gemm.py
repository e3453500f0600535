"""A/B timing of the GEMM / conv shapes of all three stages against another build of the kernel library:
   SEEDX_LIB=seed-x_b200/lib/r01/libseedx_r01.so python tools/ab_gemm.py      (round-1 kernels)   vs   python tools/ab_gemm.py   (current build)
10 back-to-back launches in one CUDA graph, CUDA events; only features both builds have are used (bias / activation / gating / residual / conv)."""
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
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n) * 1e3


def gemm_case(name, M, N, K, bias=True, act=0, gated=False, res=None, out_dtype=torch.float16):
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
    b = torch.randn(N, device=dev) if bias else None
    no = N // 2 if gated else N
    out = torch.randn(M, no, device=dev).to(out_dtype)
    kw = dict(residual=out) if res else {}
    us = timeit(lambda: ops.gemm(a, w, out=out, bias=b, act=act, gated=gated, **kw))
    print(f"{name:44s} M={M:7d} N={N:6d} K={K:6d} {us:9.1f} us {2.0 * M * N * K / us / 1e6:7.0f} TF/s", flush=True)


def conv_case(name, n, h, w_, c, cout, res=False, temb=False):
    x = torch.randn(n, h, w_, c, device=dev).half()
    cp = (c + 63) // 64 * 64
    wt = (torch.randn(cout, 9 * cp, device=dev) * (9 * c) ** -0.5).half()
    b = torch.randn(cout, device=dev)
    out = torch.randn(n, h, w_, cout, device=dev).half()
    kw = dict(residual=out) if res else {}
    if temb:
        kw["bias_g"] = torch.randn(n, cout, device=dev)
    us = timeit(lambda: ops.conv2d_nhwc(x, wt, out=out, bias=b, **kw))
    print(f"{name:44s} M={n * h * w_:7d} N={cout:6d} K={9 * cp:6d} {us:9.1f} us {2.0 * n * h * w_ * cout * 9 * c / us / 1e6:7.0f} TF/s", flush=True)


print("library:", os.environ.get("SEEDX_LIB", "current build"))
gemm_case("ViT qkv (bias)", 8192, 4992, 1664)
gemm_case("ViT fc1 (bias + GELU)", 8192, 8192, 1664, act=ops.ACT_GELU)
gemm_case("ViT fc2 (bias + fp32 residual, fp32 out)", 8192, 1664, 8192, res=True, out_dtype=torch.float32)
gemm_case("LLM prefill qkv", 988, 15360, 5120, bias=False)
gemm_case("LLM prefill gate/up (SwiGLU)", 988, 27648, 5120, bias=False, act=ops.ACT_SILU, gated=True)
gemm_case("LLM prefill down (fp32 residual)", 988, 5120, 13824, bias=False, res=True, out_dtype=torch.float32)
gemm_case("UNet attn out (bias + residual)", 8192, 1280, 1280, res=True)
gemm_case("UNet q / plain", 8192, 1280, 1280, bias=False)
gemm_case("UNet qkv plain", 8192, 3840, 1280, bias=False)
gemm_case("UNet GEGLU (bias + GELU gate)", 8192, 10240, 1280, act=ops.ACT_GELU, gated=True)
gemm_case("UNet ff2 (bias + residual)", 8192, 1280, 5120, res=True)
gemm_case("UNet 64^2 attn out (bias + residual)", 32768, 640, 640, res=True)
gemm_case("UNet 64^2 GEGLU", 32768, 5120, 640, act=ops.ACT_GELU, gated=True)
conv_case("UNet conv 1280 @32^2 (bias + residual)", 8, 32, 32, 1280, 1280, res=True)
conv_case("UNet conv 1280 @32^2 (bias + temb)", 8, 32, 32, 1280, 1280, temb=True)
conv_case("UNet conv 640 @64^2 (bias + residual)", 8, 64, 64, 640, 640, res=True)
conv_case("UNet conv 320 @128^2 (bias + temb)", 8, 128, 128, 320, 320, temb=True)
conv_case("UNet conv 1920->640 @64^2", 8, 64, 64, 1920, 640, temb=True)
conv_case("VAE conv 128 @1024^2 (1 image)", 1, 1024, 1024, 128, 128)
conv_case("VAE conv 256 @512^2 (4 images)", 4, 512, 512, 256, 256, res=True)

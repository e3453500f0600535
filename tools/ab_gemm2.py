"""In-process A/B of several builds of the kernel library on the same GEMM / conv problems, measurements interleaved (build A, B, C, A, B, C ...) so
that clock / power drift hits all builds alike; min over rounds of the mean of 10 graph-captured launches.
usage: python tools/ab_gemm2.py name=path.so name=path.so ...   ("cur" = seed-x_b200/lib/libseedx.so is always included)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200._lib import GemmArgs, F16, F32, ACT_GELU, ACT_SILU
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = {"cur": os.path.join(ROOT, "seed-x_b200", "lib", "libseedx.so")}
for a in sys.argv[1:]:
    k, v = a.split("=")
    libs[k] = v
L = {k: C.CDLL(v) for k, v in libs.items()}
ws = torch.zeros(24 * 1024 * 1024 + 16384, device="cuda", dtype=torch.uint8)
for k, l in L.items():
    if hasattr(l, "seedx_gemm_set_workspace"):
        l.seedx_gemm_set_workspace(C.c_void_p(ws.data_ptr()), C.c_int64(ws.numel()))
        l.seedx_gemm_set_stream_k(int(os.environ.get("SK", "0")))
torch.manual_seed(0)
dev = "cuda"


def call(l, g):
    rc = l.seedx_gemm_f16(C.byref(g), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc


def graph_of(l, g, n=10):
    call(l, g); call(l, g); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            call(l, g)
    gr.replay(); torch.cuda.synchronize()
    return gr


def measure(graphs, n=10, rounds=4):
    best = {k: 1e9 for k in graphs}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(rounds):
        for k, gr in graphs.items():
            e0.record(); gr.replay(); gr.replay(); e1.record(); torch.cuda.synchronize()
            best[k] = min(best[k], e0.elapsed_time(e1) / (2 * n) * 1e3)
    return best


keep = []


def gemm_args(M, N, K, bias=True, act=0, gated=False, res=False, out32=False):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
    no = N // 2 if gated else N
    out = torch.randn(M, no, device=dev).to(torch.float32 if out32 else torch.float16)
    g = GemmArgs()
    g.A, g.B, g.D = a.data_ptr(), w.data_ptr(), out.data_ptr()
    g.M, g.N, g.K, g.batch, g.lda, g.ldb, g.ldd, g.alpha = M, N, K, 1, K, K, no, 1.0
    if bias:
        b = torch.randn(N, device=dev); g.bias_n = b.data_ptr(); keep.append(b)
    if res:
        g.residual, g.residual_dtype, g.ldr = out.data_ptr(), (F32 if out32 else F16), no
    g.act, g.gated, g.out_dtype = act, int(gated), (F32 if out32 else F16)
    keep.extend([a, w, out])
    return g, 2.0 * M * N * K


def conv_args(n, h, w_, c, cout, res=False, temb=False):
    x = torch.randn(n, h, w_, c, device=dev).half(); cp = (c + 63) // 64 * 64
    wt = (torch.randn(cout, 9 * cp, device=dev) * (9 * c) ** -0.5).half(); b = torch.randn(cout, device=dev)
    out = torch.randn(n, h, w_, cout, device=dev).half()
    g = GemmArgs()
    g.A, g.B, g.D = x.data_ptr(), wt.data_ptr(), out.data_ptr()
    g.M, g.N, g.K, g.batch, g.lda, g.ldb, g.ldd, g.alpha = n * h * w_, cout, 9 * cp, 1, c, 9 * cp, cout, 1.0
    g.bias_n = b.data_ptr()
    if res:
        g.residual, g.residual_dtype, g.ldr = out.data_ptr(), F16, cout
    if temb:
        t = torch.randn(n, cout, device=dev); g.bias_g, g.bias_g_rows = t.data_ptr(), h * w_; keep.append(t)
    g.act, g.out_dtype, g.conv_taps_h, g.conv_taps_w = 0, F16, 3, 3
    g.conv_n, g.conv_h, g.conv_w, g.conv_c = n, h, w_, c
    keep.extend([x, wt, b, out])
    return g, 2.0 * n * h * w_ * cout * 9 * c


cases = [("ViT qkv (bias)", gemm_args(8192, 4992, 1664)), ("ViT fc1 (bias+GELU)", gemm_args(8192, 8192, 1664, act=ACT_GELU)),
         ("ViT fc2 (bias+fp32 res)", gemm_args(8192, 1664, 8192, res=True, out32=True)),
         ("LLM qkv", gemm_args(988, 15360, 5120, bias=False)), ("LLM gate/up SwiGLU", gemm_args(988, 27648, 5120, bias=False, act=ACT_SILU, gated=True)),
         ("LLM down (fp32 res)", gemm_args(988, 5120, 13824, bias=False, res=True, out32=True)),
         ("UNet attn-out (bias+res)", gemm_args(8192, 1280, 1280, res=True)), ("UNet q plain", gemm_args(8192, 1280, 1280, bias=False)),
         ("UNet qkv plain", gemm_args(8192, 3840, 1280, bias=False)), ("UNet GEGLU", gemm_args(8192, 10240, 1280, act=ACT_GELU, gated=True)),
         ("UNet ff2 (bias+res)", gemm_args(8192, 1280, 5120, res=True)), ("UNet 64^2 attn-out", gemm_args(32768, 640, 640, res=True)),
         ("UNet 64^2 GEGLU", gemm_args(32768, 5120, 640, act=ACT_GELU, gated=True)),
         ("conv 1280@32^2 (bias+res)", conv_args(8, 32, 32, 1280, 1280, res=True)), ("conv 640@64^2 (bias+res)", conv_args(8, 64, 64, 640, 640, res=True)),
         ("conv 320@128^2 (bias+temb)", conv_args(8, 128, 128, 320, 320, temb=True)), ("conv 1920->640@64^2", conv_args(8, 64, 64, 1920, 640, temb=True)),
         ("VAE conv 128@1024^2 x1", conv_args(1, 1024, 1024, 128, 128)), ("VAE conv 256@512^2 x4", conv_args(4, 512, 512, 256, 256, res=True))]
print(f"{'case':28s} " + " ".join(f"{k:>14s}" for k in L) + "   (us per launch; TF/s of the first build)")
for name, (g, fl) in cases:
    graphs = {k: graph_of(l, g) for k, l in L.items()}
    best = measure(graphs)
    first = next(iter(best))
    print(f"{name:28s} " + " ".join(f"{best[k]:14.1f}" for k in L) + f"   {fl / best[first] / 1e6:6.0f}", flush=True)

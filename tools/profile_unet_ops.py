"""Per-op time breakdown of one full-size UNet forward (eager, CUDA events around every library call), grouped by op + shape."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import synth, ops
from seedx_b200.sdxl import UNet2DConditionModel, SDXL_UNET, EulerDiscreteScheduler
from seedx_b200.sampler import DenoiseLoop

B = int(os.environ.get("B", "4")); branches = 2
synth.set_device("cuda")
unet = UNet2DConditionModel(dict(SDXL_UNET)); unet.load_state_dict(synth.unet_state_dict(dict(SDXL_UNET)))
synth.set_device("cpu")
Be = B * branches
loop = DenoiseLoop(unet, EulerDiscreteScheduler(), B, (128, 128), branches, use_graph=False)
loop.set_condition(torch.randn(Be, 64, 2048, device="cuda"), torch.randn(Be, 1280, device="cuda"),
                   torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]], device="cuda").repeat(Be, 1))
loop.t_dev.fill_(981.0)
for _ in range(2): loop._forward()
torch.cuda.synchronize()

records = []
def wrap(name, fn, keyfn, flopfn):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        records.append((name, keyfn(*a, **k), flopfn(*a, **k), e0, e1))
        return r
    return w
def gemm_key(a, w, out=None, **k):
    M = a.shape[-2] * (a.shape[0] if a.dim() == 3 else 1)
    return f"M={M} N={w.shape[-2]} K={a.shape[-1]}" + (" gated" if k.get("gated") else "") + (" res" if k.get("residual") is not None else "") + (" f32out" if (out is not None and out.dtype == torch.float32) or k.get("out_dtype") == torch.float32 else "")
def gemm_fl(a, w, out=None, **k):
    M = a.shape[-2] * (a.shape[0] if a.dim() == 3 else 1); return 2.0 * M * w.shape[-2] * a.shape[-1]
def conv_key(x, w, out=None, **k): return f"{tuple(x.shape)}->{w.shape[0]}" + (" res" if k.get("residual") is not None else "")
def conv_fl(x, w, out=None, **k): n, h, wd, c = x.shape; return 2.0 * n * h * wd * 9 * c * w.shape[0]
def att_key(q, k_, v, out, **k): return f"B={out.shape[0]} H={out.shape[1]} Sq={out.shape[2]} Sk={k_.shape[2]} d={out.shape[3]}"
def att_fl(q, k_, v, out, **k): return 4.0 * out.shape[0] * out.shape[1] * out.shape[2] * k_.shape[2] * out.shape[3]
ops.gemm = wrap("gemm", ops.gemm, gemm_key, gemm_fl)
ops.conv2d_nhwc = wrap("conv", ops.conv2d_nhwc, conv_key, conv_fl)
ops.attention = wrap("attn", ops.attention, att_key, att_fl)
ops.layernorm = wrap("layernorm", ops.layernorm, lambda x, *a, **k: f"{tuple(x.shape)}", lambda *a, **k: 0.0)
ops.groupnorm_nhwc = wrap("groupnorm", ops.groupnorm_nhwc, lambda x, *a, **k: f"{tuple(x.shape)}+{0 if k.get('x2') is None else k['x2'].shape[3]}", lambda *a, **k: 0.0)
for n in ("cast", "upsample2x_nhwc", "im2col_nhwc", "unary_f16", "timestep_embedding"):
    setattr(ops, n, wrap(n, getattr(ops, n), lambda *a, **k: "", lambda *a, **k: 0.0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); loop._forward(); e1.record(); torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
agg = collections.OrderedDict()
for name, key, fl, a, b in records:
    k = (name, key); d = agg.setdefault(k, [0, 0.0, 0.0]); d[0] += 1; d[1] += a.elapsed_time(b); d[2] += fl
byname = collections.defaultdict(float)
for (name, key), (n, ms, fl) in agg.items(): byname[name] += ms
print(f"UNet forward Be={Be}: {tot:.2f} ms eager; sum of op times {sum(byname.values()):.2f} ms")
print("by op:", {k: round(v, 2) for k, v in sorted(byname.items(), key=lambda kv: -kv[1])})
for (name, key), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{ms:7.2f} ms {ms/tot*100:5.1f}%  n={n:3d}  {fl/ms/1e9 if ms > 0 else 0:7.1f} TF/s  {name:10s} {key}")

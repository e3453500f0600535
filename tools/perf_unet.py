"""Scratch perf probe: full-size SDXL UNet forward (random weights) — per-forward ms, with and without CUDA graph."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import synth, ops, _lib
from seedx_b200.sdxl import UNet2DConditionModel, SDXL_UNET, EulerDiscreteScheduler
from seedx_b200.sampler import DenoiseLoop

B = int(os.environ.get("B", "1"))
branches = int(os.environ.get("BR", "2"))
t0 = time.time()
cfg = dict(SDXL_UNET)
synth.set_device('cuda'); sd = synth.unet_state_dict(cfg); synth.set_device('cpu')
print("synth weights", time.time() - t0, "s", flush=True)
unet = UNet2DConditionModel(cfg)
unet.load_state_dict(sd)
del sd
print("loaded", time.time() - t0, "s", flush=True)
Be = B * branches
ctx = torch.randn(Be, 64, 2048, device="cuda")
te = torch.randn(Be, 1280, device="cuda")
tid = torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]], device="cuda").repeat(Be, 1)
loop = DenoiseLoop(unet, EulerDiscreteScheduler(), B, (128, 128), branches, use_graph=False)
loop.set_condition(ctx, te, tid)
loop.t_dev.fill_(981.0)
for _ in range(2):
    eps = loop._forward()
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eps = loop._forward()
e1.record(); torch.cuda.synchronize()
print(f"eager UNet forward Be={Be}: {e0.elapsed_time(e1)/3:.2f} ms, launches/forward {(_lib.launch_count()-n0)//3}", flush=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    eps = loop._forward()
for _ in range(2):
    g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    g.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"graph UNet forward Be={Be}: {ms:.2f} ms -> {6.75*Be/ms:.1f} TFLOP/s ({6.75*Be/ms/1422.7*100:.1f}% of sustained bf16 peak)", flush=True)
print("eps finite:", torch.isfinite(eps).all().item(), eps.float().std().item())

"""Scratch perf probe: full-size SDXL VAE decode (random weights): ms per image batch, eager."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import synth, _lib
from seedx_b200.sdxl import AutoencoderKL, SDXL_VAE
B = int(os.environ.get("B", "4"))
synth.set_device("cuda")
sd = {k: v for k, v in synth.vae_state_dict(SDXL_VAE).items() if not k.startswith(("encoder.", "quant_conv"))}
synth.set_device("cpu")
vae = AutoencoderKL(SDXL_VAE); vae.load_state_dict(sd)
z = torch.randn(B, 4, 128, 128, device="cuda")
for _ in range(2):
    img = vae.decode_nhwc(z)
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    img = vae.decode_nhwc(z)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"VAE decode B={B}: {ms:.1f} ms ({ms/B:.1f} ms/image, {10.47*B/ms:.0f} TFLOP/s), launches {(_lib.launch_count()-n0)//3}, finite={torch.isfinite(img).all().item()}")

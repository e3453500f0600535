"""One full-size VAE decode (B images, 128x128 latents -> 1024x1024) inside a cudaProfilerStart/Stop range for the per-kernel launch list (see
tools/round_gpu_check.sh `launches`).  env: B (default 4), SEEDX_EPI_STATS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import synth
from seedx_b200.sdxl import AutoencoderKL, SDXL_VAE
B = int(os.environ.get("B", "4"))
synth.set_device("cuda")
sd = {k: v for k, v in synth.vae_state_dict(SDXL_VAE).items() if not k.startswith(("encoder.", "quant_conv"))}
synth.set_device("cpu")
vae = AutoencoderKL(SDXL_VAE); vae.load_state_dict(sd)
z = torch.randn(B, 4, 128, 128, device="cuda")
vae.decode_nhwc(z); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
vae.decode_nhwc(z); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")

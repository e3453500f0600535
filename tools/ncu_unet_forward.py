"""One full-size UNet sample-forward (Be = BRANCHES x B samples, 128x128 latents) inside a cudaProfilerStart/Stop range, for the per-kernel launch
list with durations and DRAM bytes:
   ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
       --cache-control none --csv --log-file X python tools/ncu_unet_forward.py
env: B (requests, default 4), BRANCHES (2 = t2i, 3 = edit)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import synth
from seedx_b200.sdxl import UNet2DConditionModel, SDXL_UNET, EulerDiscreteScheduler
from seedx_b200.sampler import DenoiseLoop
B = int(os.environ.get("B", "4"))
BR = int(os.environ.get("BRANCHES", "2"))
cfg = dict(SDXL_UNET, in_channels=8 if BR == 3 else 4)
synth.set_device("cuda"); sd = synth.unet_state_dict(cfg); synth.set_device("cpu")
unet = UNet2DConditionModel(cfg); unet.load_state_dict(sd); del sd
Be = BR * B
loop = DenoiseLoop(unet, EulerDiscreteScheduler(), B, (128, 128), BR, use_graph=False)
loop.set_condition(torch.randn(Be, 64, 2048, device="cuda"), torch.randn(Be, 1280, device="cuda"),
                   torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]], device="cuda").repeat(Be, 1))
loop.t_dev.fill_(981.0)
loop._forward(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loop._forward(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")

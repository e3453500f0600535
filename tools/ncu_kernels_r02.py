"""One launch each (after warm-up, inside a cudaProfilerStart/Stop range) of the kernels whose `ncu --set full` summaries are committed under profiles/:
  ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/r02_kernels python tools/ncu_kernels_r02.py
Shapes are the ones of the bench step (8 UNet samples): see the comments."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
torch.manual_seed(0)
Be = 8
dev = "cuda"
jobs = []
# 1 self-attention of the 64x64 level: 8 samples x 10 heads x 4096 tokens, d = 64 (flash_attn_pp_kernel<64, *>)
qkv = torch.randn(Be, 4096, 3, 10, 64, device=dev).half()
o = torch.empty(Be, 4096, 10, 64, device=dev, dtype=torch.float16)
jobs.append(lambda: ops.attention(*(qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)), o.permute(0, 2, 1, 3), scale=0.125))
# 2 self-attention of the 32x32 level: 8 x 20 heads x 1024 tokens
qkv2 = torch.randn(Be, 1024, 3, 20, 64, device=dev).half()
o2 = torch.empty(Be, 1024, 20, 64, device=dev, dtype=torch.float16)
jobs.append(lambda: ops.attention(*(qkv2[:, :, i].permute(0, 2, 1, 3) for i in range(3)), o2.permute(0, 2, 1, 3), scale=0.125))
# 3 cross-attention: 1024 queries x 64 context keys, 20 heads
q3 = torch.randn(Be, 1024, 20, 64, device=dev).half()
kv3 = torch.randn(Be, 64, 2, 20, 64, device=dev).half()
jobs.append(lambda: ops.attention(q3.permute(0, 2, 1, 3), kv3[:, :, 0].permute(0, 2, 1, 3), kv3[:, :, 1].permute(0, 2, 1, 3), o2.permute(0, 2, 1, 3), scale=0.125))
# 4 GroupNorm + SiLU, 8 x 64x64 x 640 and 8 x 32x32 x 1280 (gn_stats_kernel + gn_apply_kernel)
x4 = torch.randn(Be, 64, 64, 640, device=dev).half()
g4, b4 = torch.ones(640, device=dev), torch.zeros(640, device=dev)
ws = ops.groupnorm_ws(Be, 32, dev)
jobs.append(lambda: ops.groupnorm_nhwc(x4, g4, b4, 1e-5, silu=True, stats_ws=ws))
# 5 LayerNorm of the transformer residual stream, 8192 x 1280 fp16 (layernorm_rows_kernel<__half, __half, 12>)
x5 = torch.randn(8192, 1280, device=dev).half()
g5, b5 = torch.ones(1280, device=dev), torch.zeros(1280, device=dev)
o5 = torch.empty_like(x5)
jobs.append(lambda: ops.layernorm(x5, g5, b5, 1e-5, out=o5))
# 6 GEGLU projection 8192 x 10240 x 1280 (gemm_tc_kernel<256,2>) and 7 the 1280-channel 3x3 conv at 32x32 (K = 11520) with residual
a6 = torch.randn(8192, 1280, device=dev).half(); w6 = (torch.randn(10240, 1280, device=dev) * 0.03).half(); o6 = torch.empty(8192, 5120, device=dev, dtype=torch.float16)
bias6 = torch.randn(10240, device=dev)
jobs.append(lambda: ops.gemm(a6, w6, out=o6, bias=bias6, act=ops.ACT_GELU, gated=True))
x7 = torch.randn(Be, 32, 32, 1280, device=dev).half(); w7 = (torch.randn(1280, 9 * 1280, device=dev) * 0.01).half(); r7 = torch.randn(Be, 32, 32, 1280, device=dev).half()
o7 = torch.empty_like(r7); bias7 = torch.randn(1280, device=dev)
jobs.append(lambda: ops.conv2d_nhwc(x7, w7, out=o7, bias=bias7, residual=r7))
# 8 attention-out projection with in-place residual, 8192 x 1280 x 1280 (the wave-quantised family)
a8 = torch.randn(8192, 1280, device=dev).half(); w8 = (torch.randn(1280, 1280, device=dev) * 0.03).half(); h8 = torch.randn(8192, 1280, device=dev).half()
jobs.append(lambda: ops.gemm(a8, w8, out=h8, bias=bias7, residual=h8))
# 9 decode GEMV: gate/up projection of LLaMA-13B for 4 lock-step sequences (gemv_mma_kernel) and 10 decode attention over a 512-token paged cache
W9 = torch.randn(27648, 5120, device=dev).half(); x9 = torch.randn(4, 5120, device=dev); o9 = torch.empty(4, 13824, device=dev); rw = torch.ones(5120, device=dev)
jobs.append(lambda: ops.gemv(W9, x9, o9, rms_w=rw, gated=True))
H, d, T, Bq = 40, 128, 512, 4
kc = torch.randn(Bq, 1024, H * d, device=dev).half(); vc = torch.randn_like(kc)
qkv10 = torch.randn(Bq, 3 * H * d, device=dev); st = torch.tensor([[T, 0, 0, 1]] * Bq, dtype=torch.int32, device=dev)
inv = (1.0 / (10000.0 ** (torch.arange(0, d, 2, dtype=torch.float32) / d))).to(dev); o10 = torch.empty(Bq, H * d, device=dev)
jobs.append(lambda: ops.decode_attention(qkv10, st, inv, kc, vc, o10, H, d))
if os.environ.get("JOBS"):                       # e.g. JOBS=0,1: only the two self-attention launches
    jobs = [jobs[int(i)] for i in os.environ["JOBS"].split(",")]
for _ in range(2):
    for j in jobs:
        j()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for j in jobs:
    j()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", len(jobs), "jobs")

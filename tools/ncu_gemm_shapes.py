"""The UNet's GEMM shapes (Be=8), one profiled launch each in a fixed order, for an ncu launch list:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file X python tools/ncu_gemm_shapes.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
SHAPES = [  # M, N, K, kind
    (8192, 1280, 1280, "plain"), (8192, 1280, 1280, "res"), (8192, 1280, 5120, "res"), (8192, 3840, 1280, "plain"), (8192, 10240, 1280, "geglu"),
    (32768, 640, 640, "plain"), (32768, 640, 640, "res"), (32768, 640, 2560, "res"), (32768, 1920, 640, "plain"), (32768, 5120, 640, "geglu"),
    (8192, 2560, 2048, "plain"), (8192, 8192, 8192, "plain"),
]
R = 3
bufs = []
for (M, N, K, kind) in SHAPES:
    A = [torch.randn(M, K, device="cuda").half() for _ in range(R)]
    W = [torch.randn(N, K, device="cuda").half() * 0.03 for _ in range(R)]
    No = N // 2 if kind == "geglu" else N
    O = [torch.randn(M, No, device="cuda").half() for _ in range(R)]
    bias = torch.randn(N, device="cuda")
    bufs.append((A, W, O, bias))
def run(i, j):
    M, N, K, kind = SHAPES[i]
    A, W, O, bias = bufs[i]
    if kind == "plain": ops.gemm(A[j], W[j], out=O[j])
    elif kind == "res": ops.gemm(A[j], W[j], out=O[j], bias=bias, residual=O[j])
    else: ops.gemm(A[j], W[j], out=O[j], bias=bias, act=ops.ACT_GELU, gated=True)
for i in range(len(SHAPES)):
    for j in range(R): run(i, j)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for rep in range(2):
    for i in range(len(SHAPES)):
        run(i, rep)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
for i, s in enumerate(SHAPES): print(i, s)

"""One profiled launch of one GEMM shape: python tools/ncu_gemm_one.py M N K plain|res|geglu   (use with ncu --profile-from-start off)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4]); kind = sys.argv[4]
R = 3
A = [torch.randn(M, K, device="cuda").half() for _ in range(R)]
W = [torch.randn(N, K, device="cuda").half() * 0.03 for _ in range(R)]
O = [torch.randn(M, N // 2 if kind == "geglu" else N, device="cuda").half() for _ in range(R)]
bias = torch.randn(N, device="cuda")
def run(j):
    if kind == "plain": ops.gemm(A[j], W[j], out=O[j])
    elif kind == "res": ops.gemm(A[j], W[j], out=O[j], bias=bias, residual=O[j])
    else: ops.gemm(A[j], W[j], out=O[j], bias=bias, act=ops.ACT_GELU, gated=True)
for j in range(R): run(j)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run(0)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()

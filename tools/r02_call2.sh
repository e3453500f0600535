#!/bin/bash
# GPU call 2 of round 2: correctness of the new kernels (folded LayerNorm, epilogue statistics, stream-K, 64-key tcgen05 cross-attention,
# unrolled decode attention), then A/B timings of the UNet forward.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/c2_pytest_kernels.log 2>&1
echo "[kernels] rc=$?"; tail -5 gpurun_out/c2_pytest_kernels.log
timeout 900 python -m pytest tests/test_sdxl_gpu.py tests/test_llm_gpu.py tests/test_fullsize_gpu.py tests/test_dropin_gpu.py tests/test_scripts_gpu.py tests/test_vit_gpu.py -m gpu -q -s --timeout=300 -p no:cacheprovider > gpurun_out/c2_pytest_models.log 2>&1
echo "[models] rc=$?"; grep -E "passed|failed|FAILED|rel |PSNR" gpurun_out/c2_pytest_models.log | tail -30
python tools/bench_xattn.py 2>&1 | tail -4
echo "== attention, H2 softmax (default)"; python tools/bench_attn.py 2>&1 | head -4
echo "== attention, round-1 softmax (SEEDX_PP_POLY_EVERY=4)"; SEEDX_PP_POLY_EVERY=4 python tools/bench_attn.py 2>&1 | head -4
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== UNet forward B=4: SEEDX_EPI_STATS=$1 SEEDX_GEMM_STREAM_K=$2"
  SEEDX_EPI_STATS=$1 SEEDX_GEMM_STREAM_K=$2 B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet|eps finite|Error|error" | tail -3
done

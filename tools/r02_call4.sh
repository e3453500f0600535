#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_sdxl_gpu.py tests/test_fullsize_gpu.py tests/test_llm_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/c4_pytest.log 2>&1
echo "[tests] rc=$?"; tail -6 gpurun_out/c4_pytest.log
python tools/bench_gemm_shapes.py 2>&1 | tail -60
for cfg in "1 1" "0 0"; do
  set -- $cfg
  echo "== UNet forward B=4: SEEDX_EPI_STATS=$1 SEEDX_GEMM_STREAM_K=$2"
  SEEDX_EPI_STATS=$1 SEEDX_GEMM_STREAM_K=$2 B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet|Error|error" | tail -3
done

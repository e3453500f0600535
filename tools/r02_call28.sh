#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sdxl_gpu.py tests/test_gemm_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/c28_pytest.log 2>&1
echo "[tests] rc=$?"; tail -3 gpurun_out/c28_pytest.log
bash tools/round_gpu_check.sh launches 2>&1 | grep -v "^  " | head -20
B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet" | tail -1

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_sdxl_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/c15_pytest.log 2>&1
echo "[tests] rc=$?"; tail -5 gpurun_out/c15_pytest.log
B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet|Error|error|launches" | tail -3
B=1 BR=3 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet|Error|error" | tail -2

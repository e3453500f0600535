#!/bin/bash
set -u
mkdir -p gpurun_out
R=seed-x_b200/lib/r01
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_sdxl_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/c14_pytest.log 2>&1
echo "[tests] rc=$?"; tail -3 gpurun_out/c14_pytest.log
SK=1 timeout 600 python tools/ab_gemm2.py r01=$R/libseedx_r01.so 2>&1 | tail -20
B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet|Error|error" | tail -3

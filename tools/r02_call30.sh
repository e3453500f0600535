#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_sdxl_gpu.py tests/test_fullsize_gpu.py tests/test_vit_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/c30_pytest.log 2>&1
echo "[tests] rc=$?"; tail -3 gpurun_out/c30_pytest.log
for t in 1 0 1 0; do echo "== SEEDX_GEMM_AUTOTUNE=$t"; SEEDX_GEMM_AUTOTUNE=$t B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet|tuned" | tail -2; done

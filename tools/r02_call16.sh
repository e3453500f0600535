#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider -k "epilogue_statistics" > gpurun_out/c16_pytest.log 2>&1
echo "[tests] rc=$?"; tail -3 gpurun_out/c16_pytest.log
for t in 1 0 1 0; do echo "== SEEDX_ROW_TICKETS=$t"; SEEDX_ROW_TICKETS=$t B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet" | tail -1; done

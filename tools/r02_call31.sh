#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider -k "conv" > gpurun_out/c31_pytest.log 2>&1
echo "[tests conv] rc=$?"; tail -6 gpurun_out/c31_pytest.log
SEEDX_CONV_ROW=1 timeout 300 python tools/bench_conv_row.py 2>&1 | tail -10
SEEDX_CONV_ROW=0 timeout 300 python tools/bench_conv_row.py 2>&1 | tail -10

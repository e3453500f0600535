#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 2 3; do SEEDX_GEMV_IMPL=$i timeout 600 python -m pytest tests/test_llm_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider -k "gemv or decode or greedy or batch" > gpurun_out/c21_pytest_$i.log 2>&1; echo "[tests impl=$i] rc=$?"; tail -3 gpurun_out/c21_pytest_$i.log; done
for i in 1 2 3; do echo "== SEEDX_GEMV_IMPL=$i"; SEEDX_GEMV_IMPL=$i timeout 300 python tools/perf_gemv.py 2>&1 | tail -5; done
for i in 1 2 3; do echo "== SEEDX_GEMV_IMPL=$i"; SEEDX_GEMV_IMPL=$i timeout 300 python tools/perf_llm.py 2>&1 | tail -2; done

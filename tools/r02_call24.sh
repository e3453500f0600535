#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider -k attention > gpurun_out/c24_pytest.log 2>&1
echo "[tests] rc=$?"; tail -3 gpurun_out/c24_pytest.log
timeout 300 python tools/bench_attn.py 2>&1 | head -6
for pe in 2 0; do echo "== SEEDX_PP_POLY_EVERY=$pe"; SEEDX_PP_POLY_EVERY=$pe timeout 300 python tools/bench_attn.py 2>&1 | sed -n 2p; done
for l in cur cur; do echo "== $l"; B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet" | tail -1; done

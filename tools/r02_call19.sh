#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_llm_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider -k "attention or gemv or decode or greedy or batch" > gpurun_out/c19_pytest.log 2>&1
echo "[tests] rc=$?"; tail -5 gpurun_out/c19_pytest.log
timeout 300 python tools/bench_attn.py 2>&1 | head -4
for kb in 0 160 320 80; do echo "== SEEDX_GEMV_PREFETCH_KB=$kb"; SEEDX_GEMV_PREFETCH_KB=$kb timeout 300 python tools/perf_gemv.py 2>&1 | tail -5; done
for kb in 0 160 320; do echo "== SEEDX_GEMV_PREFETCH_KB=$kb"; SEEDX_GEMV_PREFETCH_KB=$kb timeout 300 python tools/perf_llm.py 2>&1 | tail -2; done
for l in cur; do echo "== $l"; B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet" | tail -1; done

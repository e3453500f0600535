#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_llm_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider -k "attention or gemv or decode or greedy or batch" > gpurun_out/c18_pytest.log 2>&1
echo "[tests] rc=$?"; tail -5 gpurun_out/c18_pytest.log
timeout 300 python tools/bench_attn.py 2>&1 | head -4
SEEDX_LIB=seed-x_b200/lib/r02a/libseedx_prev.so timeout 300 python tools/bench_attn.py 2>&1 | head -4
for pe in 2 1 0; do echo "== SEEDX_PP_POLY_EVERY=$pe"; SEEDX_PP_POLY_EVERY=$pe timeout 300 python tools/bench_attn.py 2>&1 | head -4; done
for l in cur prev cur prev; do if [ $l = prev ]; then export SEEDX_LIB=seed-x_b200/lib/r02a/libseedx_prev.so; else unset SEEDX_LIB; fi; echo "== $l"; B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet" | tail -1; done
unset SEEDX_LIB
for i in 1 2; do echo "== SEEDX_GEMV_IMPL=$i"; SEEDX_GEMV_IMPL=$i timeout 300 python tools/perf_gemv.py 2>&1 | tail -5; done
for i in 1 2; do echo "== SEEDX_GEMV_IMPL=$i"; SEEDX_GEMV_IMPL=$i timeout 300 python tools/perf_llm.py 2>&1 | tail -2; done

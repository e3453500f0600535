#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider -k "attention" > gpurun_out/c17_pytest.log 2>&1
echo "[tests] rc=$?"; tail -5 gpurun_out/c17_pytest.log
timeout 300 python tools/bench_attn.py 2>&1 | tail -8
SEEDX_LIB=seed-x_b200/lib/r02a/libseedx_prev.so timeout 300 python tools/bench_attn.py 2>&1 | head -5
for l in cur prev cur prev; do if [ $l = prev ]; then export SEEDX_LIB=seed-x_b200/lib/r02a/libseedx_prev.so; else unset SEEDX_LIB; fi; echo "== $l"; B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet" | tail -1; done
unset SEEDX_LIB
timeout 300 python tools/perf_gemv.py 2>&1 | tail -8
timeout 300 python tools/perf_llm.py 2>&1 | tail -2

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider -k attention > gpurun_out/c26_pytest.log 2>&1
echo "[tests] rc=$?"; tail -3 gpurun_out/c26_pytest.log
for l in cur prev2 cur prev2; do if [ $l = prev2 ]; then export SEEDX_LIB=seed-x_b200/lib/r02a/libseedx_prev2.so; else unset SEEDX_LIB; fi; echo "== $l"; timeout 300 python tools/bench_attn.py 2>&1 | sed -n 2,4p; done

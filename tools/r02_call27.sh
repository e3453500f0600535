#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_llm_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/c27_pytest.log 2>&1
echo "[tests] rc=$?"; tail -3 gpurun_out/c27_pytest.log
timeout 300 python tools/perf_llm.py 2>&1 | tail -2
SEEDX_LIB=seed-x_b200/lib/r02a/libseedx_prev2.so timeout 300 python tools/perf_llm.py 2>&1 | tail -2
timeout 300 python tools/perf_llm.py 2>&1 | tail -2
for k in 80 81 80 81; do echo "== SK_MIN_KBLOCKS=$k"; SEEDX_GEMM_STREAM_K=1 SEEDX_SK_MIN_KBLOCKS=$k B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet" | tail -1; done

#!/bin/bash
set -u
mkdir -p gpurun_out
R=seed-x_b200/lib/r01
echo "== A/B stream-K auto (cur) vs r01"
SK=1 timeout 600 python tools/ab_gemm2.py r01=$R/libseedx_r01.so 2>&1 | tail -20
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_sdxl_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/c11_pytest.log 2>&1
echo "[tests] rc=$?"; tail -3 gpurun_out/c11_pytest.log
B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet|Error|error" | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; echo "[bench] rc=$?"; tail -2 gpurun_out/c11_bench.err
python -c "
import json; d=json.load(open('gpurun_out/c11_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'unet launch ms', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'], 'clocks', d['clocks']); print({k: round(v['frac'],3) for k,v in d['stage_roofline'].items()}); print(d['stage_detail_ms'])"

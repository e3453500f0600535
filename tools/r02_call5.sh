#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_sdxl_gpu.py tests/test_vit_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/c5_pytest.log 2>&1
echo "[tests] rc=$?"; tail -4 gpurun_out/c5_pytest.log
python tools/bench_gemm_shapes.py 2>&1 | grep -E "GEGLU|qkv|conv3x3" | grep -v "dyn_b"
for cfg in "1 1" "0 0"; do
  set -- $cfg
  echo "== UNet forward B=4: SEEDX_EPI_STATS=$1 SEEDX_GEMM_STREAM_K=$2"
  SEEDX_EPI_STATS=$1 SEEDX_GEMM_STREAM_K=$2 B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet|Error|error" | tail -3
done
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/c5_unet_launches.csv python tools/ncu_unet_forward.py > gpurun_out/c5_unet_launches.log 2>&1
echo "== launches (defaults) rc=$?"; python tools/summarize_launches.py gpurun_out/c5_unet_launches.csv 16
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "[bench] rc=$?"; tail -2 gpurun_out/c5_bench.err
python -c "
import json; d=json.load(open('gpurun_out/c5_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'unet launch ms', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'], 'clocks', d['clocks']); print({k: round(v['frac'],3) for k,v in d['stage_roofline'].items()})"

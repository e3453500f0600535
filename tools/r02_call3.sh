#!/bin/bash
# GPU call 3 of round 2: all kernel / model tests after the fixes, then launch lists of the UNet forward for three configurations to see where
# the forward moved relative to profiles/r02_unet_forward_launches_8samples.csv (start of round).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/c3_pytest_kernels.log 2>&1
echo "[kernels] rc=$?"; tail -8 gpurun_out/c3_pytest_kernels.log
timeout 900 python -m pytest tests/test_sdxl_gpu.py tests/test_llm_gpu.py tests/test_fullsize_gpu.py tests/test_dropin_gpu.py tests/test_scripts_gpu.py tests/test_vit_gpu.py -m gpu -q -s --timeout=300 -p no:cacheprovider > gpurun_out/c3_pytest_models.log 2>&1
echo "[models] rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/c3_pytest_models.log | tail -8
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for cfg in "0 0 off" "1 0 epi" "0 1 sk"; do
  set -- $cfg
  SEEDX_EPI_STATS=$1 SEEDX_GEMM_STREAM_K=$2 timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv \
     --log-file gpurun_out/c3_unet_launches_$3.csv python tools/ncu_unet_forward.py > gpurun_out/c3_unet_launches_$3.log 2>&1
  echo "== launches EPI_STATS=$1 STREAM_K=$2 rc=$?"; python tools/summarize_launches.py gpurun_out/c3_unet_launches_$3.csv 14
done

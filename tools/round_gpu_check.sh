#!/bin/bash
# One gpurun call that produces everything a round needs from the GPU, in order of value per second (a fresh box costs ~1 min of budget per call,
# so batch the steps).  Usage on the build container:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/round_gpu_check.sh tests bench launches ncu'
# Sections: tests (pytest -m gpu, ~1 min) | smoke | bench (N=1 default line, ~3 min) | cmp (comprehension workload, ~40 s) |
#           launches (ncu launch list of one UNet forward, ~2 min) | ncu (ncu --set full of the dominant GEMM shape, ~1 min)
# Everything lands in gpurun_out/ (merged back by gpurun); copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
for sec in "$@"; do
  case $sec in
    tests)    timeout 400 python -m pytest tests -m gpu -q -s --durations=10 --timeout=150 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
              echo "[tests] rc=$?"; grep -E "passed|failed|FAILED|rel " gpurun_out/pytest_gpu.log | tail -25 ;;
    smoke)    timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    bench)    timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench] rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json ;;
    cmp)      timeout 200 python bench.py --workload comprehension --no-cpu-baseline > gpurun_out/bench_cmp.json 2> gpurun_out/bench_cmp.err
              echo "[cmp] rc=$?"; cat gpurun_out/bench_cmp.json ;;
    launches) timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
                --log-file gpurun_out/unet_launches.csv python tools/ncu_unet_forward.py > gpurun_out/unet_launches.log 2>&1
              echo "[launches] rc=$?"; python tools/summarize_launches.py gpurun_out/unet_launches.csv | head -30 ;;
    ncu)      timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/gemm_geglu \
                python tools/ncu_gemm_one.py 8192 10240 1280 geglu > gpurun_out/ncu_gemm.log 2>&1
              echo "[ncu] rc=$?"; ls -la gpurun_out/gemm_geglu.ncu-rep ;;
    *)        echo "unknown section $sec" ;;
  esac
done

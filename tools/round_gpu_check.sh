#!/bin/bash
# One gpurun call that produces everything a round needs from the GPU, in order of value per second (a fresh box costs ~1 min of budget per call,
# so batch the steps).  Usage on the build container:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round_gpu_check.sh tests fullsize bench edit anyres cmp launches kernels'
# Sections: tests (pytest -m gpu without the full-size oracle file, ~1.5 min) | fullsize (tests/test_fullsize_oracle_gpu.py, ~6 min, mostly host CPU)
#           smoke | bench (N=1 default line, ~4 min) | edit | anyres | cmp (other workloads) | ref (CPU reference arm, i2i)
#           launches (ncu launch list + DRAM bytes of one UNet forward, t2i and edit, and of the token loop) | kernels (ncu --set full of 10 kernels)
# Everything lands in gpurun_out/ (merged back by gpurun); raw CSVs / text exports that should be judged are copied into profiles/.
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for sec in "$@"; do
  case $sec in
    tests)    timeout 600 python -m pytest tests -m gpu -q -s --durations=10 --timeout=300 -p no:cacheprovider --ignore=tests/test_fullsize_oracle_gpu.py > gpurun_out/pytest_gpu.log 2>&1
              echo "[tests] rc=$?"; grep -E "passed|failed|FAILED|rel |PSNR" gpurun_out/pytest_gpu.log | tail -40 ;;
    fullsize) timeout 1200 python -m pytest tests/test_fullsize_oracle_gpu.py -m gpu -q -s --durations=10 --timeout=600 -p no:cacheprovider > gpurun_out/pytest_fullsize.log 2>&1
              echo "[fullsize] rc=$?"; grep -E "passed|failed|FAILED|rel |PSNR|fullsize" gpurun_out/pytest_fullsize.log | tail -40 ;;
    smoke)    timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    bench)    timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench] rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json ;;
    edit)     timeout 600 python bench.py --workload edit --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_edit.json 2> gpurun_out/bench_edit.err
              echo "[edit] rc=$?"; tail -3 gpurun_out/bench_edit.err; cat gpurun_out/bench_edit.json ;;
    anyres)   timeout 600 python bench.py --workload anyres --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_anyres.json 2> gpurun_out/bench_anyres.err
              echo "[anyres] rc=$?"; tail -3 gpurun_out/bench_anyres.err; cat gpurun_out/bench_anyres.json ;;
    cmp)      timeout 300 python bench.py --workload comprehension --no-cpu-baseline > gpurun_out/bench_cmp.json 2> gpurun_out/bench_cmp.err
              echo "[cmp] rc=$?"; cat gpurun_out/bench_cmp.json ;;
    ref)      timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
              echo "[ref] rc=$?"; cat gpurun_out/bench_ref.json ;;
    launches) timeout 500 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv \
                --log-file gpurun_out/unet_launches.csv python tools/ncu_unet_forward.py > gpurun_out/unet_launches.log 2>&1
              echo "[launches t2i] rc=$?"; python tools/summarize_launches.py gpurun_out/unet_launches.csv | head -30
              BRANCHES=3 B=1 timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv \
                --log-file gpurun_out/unet_launches_edit_b1.csv python tools/ncu_unet_forward.py > gpurun_out/unet_launches_edit.log 2>&1
              echo "[launches edit] rc=$?"; python tools/summarize_launches.py gpurun_out/unet_launches_edit_b1.csv | head -12
              timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv \
                --log-file gpurun_out/decode_launches_b1.csv python tools/ncu_decode_step.py > gpurun_out/decode_launches.log 2>&1
              echo "[launches decode] rc=$?"; python tools/summarize_launches.py gpurun_out/decode_launches_b1.csv | head -14 ;;
    kernels)  timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/r02_kernels \
                python tools/ncu_kernels_r02.py > gpurun_out/ncu_kernels.log 2>&1
              echo "[kernels] rc=$?"; ls -la gpurun_out/r02_kernels.ncu-rep ;;
    *)        echo "unknown section $sec" ;;
  esac
done

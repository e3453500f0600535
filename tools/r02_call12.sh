#!/bin/bash
set -u
mkdir -p gpurun_out
R=seed-x_b200/lib/r01
SK=1 timeout 600 python tools/ab_gemm2.py r01=$R/libseedx_r01.so 2>&1 | tail -20
for e in 1 0; do echo "== VAE decode B=4 SEEDX_EPI_STATS=$e"; SEEDX_EPI_STATS=$e python tools/perf_vae.py 2>&1 | tail -1; done
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for e in 1 0; do
  SEEDX_EPI_STATS=$e timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/c12_vae_launches_epi$e.csv python tools/ncu_vae_decode.py > gpurun_out/c12_vae.log 2>&1
  echo "== VAE launches EPI_STATS=$e rc=$?"; python tools/summarize_launches.py gpurun_out/c12_vae_launches_epi$e.csv 14
done

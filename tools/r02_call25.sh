#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_attn.py 2>&1 | head -4
for l in cur prev cur prev; do if [ $l = prev ]; then export SEEDX_LIB=seed-x_b200/lib/r02a/libseedx_prev.so; else unset SEEDX_LIB; fi; echo "== $l"; B=4 timeout 300 python tools/perf_unet.py 2>&1 | grep -E "graph UNet" | tail -1; done
unset SEEDX_LIB
JOBS=0,1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/r02_attn_pp2 python tools/ncu_kernels_r02.py > gpurun_out/ncu_attn2.log 2>&1
echo rc=$?; ls -la gpurun_out/r02_attn_pp2.ncu-rep

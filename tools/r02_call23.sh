#!/bin/bash
mkdir -p gpurun_out
JOBS=0 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/r02_attn_pp python tools/ncu_kernels_r02.py > gpurun_out/ncu_attn.log 2>&1
echo rc=$?; ls -la gpurun_out/r02_attn_pp.ncu-rep

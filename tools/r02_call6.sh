#!/bin/bash
set -u
for lib in seed-x_b200/lib/r01/libseedx_r01.so "" seed-x_b200/lib/r01/libseedx_NO_BPRE.so seed-x_b200/lib/r01/libseedx_NO_VEC.so; do
  echo "=================== $lib"
  SEEDX_LIB=$lib SEEDX_GEMM_STREAM_K=0 timeout 300 python tools/ab_gemm.py 2>&1 | tail -22
done
echo "=================== current build, stream-K auto"
timeout 300 python tools/ab_gemm.py 2>&1 | grep -E "conv|ff2|fc2|down"

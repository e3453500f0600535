#!/bin/bash
set -u
R=seed-x_b200/lib/r01
echo "== stream-K off in all builds"
SK=0 timeout 600 python tools/ab_gemm2.py r01=$R/libseedx_r01.so nobpre=$R/libseedx_NO_BPRE.so novec=$R/libseedx_NO_VEC.so 2>&1 | tail -22
echo "== stream-K auto (cur / nobpre / novec), r01 for reference"
SK=1 timeout 600 python tools/ab_gemm2.py r01=$R/libseedx_r01.so 2>&1 | tail -22

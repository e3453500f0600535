#!/bin/bash
SEEDX_GEMV_IMPL=1 timeout 300 python tools/perf_gemv_asym.py 2>&1 | tail -8
SEEDX_GEMV_IMPL=2 timeout 300 python tools/perf_gemv_asym.py 2>&1 | tail -8

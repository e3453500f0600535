#!/bin/bash
bash tools/round_gpu_check.sh tests fullsize smoke bench edit anyres cmp ref launches kernels

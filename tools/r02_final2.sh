#!/bin/bash
bash tools/round_gpu_check.sh tests smoke bench cmp

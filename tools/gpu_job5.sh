#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=${1:-64}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_ --launch-skip 6 -c 1 -f -o gpurun_out/job5_attn_v$V python tools/kernel_bench.py --what attn1 --variants $V --iters 6 > gpurun_out/job5_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/job5_ncu.log

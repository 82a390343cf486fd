#!/bin/bash
# early-QK attention schedule (variant 128): parity, kernel bench against variant 0, hand-over trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/kernel_bench.jsonl
export LLB_WAIT_TIMEOUT_NS=200000000
if [ "$1" != "trace" ]; then
timeout 600 python -m pytest tests/test_attn_gpu.py -x -q -k "v128" > gpurun_out/b1_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/b1_tests.log
timeout 300 python tools/kernel_bench.py --what attn --variants 0,128 --iters 20 > gpurun_out/b1_kb.log 2>&1; echo "kb rc=$?"
grep -E "llb_attn|sdpa" gpurun_out/b1_kb.log | cut -c1-200
fi
LLB200_LIB=longlive_b200/libllb200_trace.so timeout 120 python tools/attn_trace.py --variant 128 --first 20 --n 3 > gpurun_out/trace_v128.txt 2>&1; tail -75 gpurun_out/trace_v128.txt

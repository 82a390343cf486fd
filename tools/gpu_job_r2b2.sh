#!/bin/bash
# staged TMA-store epilogue + 4 K/V stages: parity, kernel bench, cross-attention trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/kernel_bench.jsonl
export LLB_WAIT_TIMEOUT_NS=200000000
timeout 600 python -m pytest tests/test_attn_gpu.py -x -q > gpurun_out/b2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/b2_tests.log
timeout 300 python tools/kernel_bench.py --what attn --variants 0 --iters 20 > gpurun_out/b2_kb.log 2>&1; echo "kb rc=$?"
grep -E "llb_attn|sdpa" gpurun_out/b2_kb.log | cut -c1-200
LLB200_LIB=longlive_b200/libllb200_trace.so timeout 120 python tools/attn_trace.py --variant 0 --lk 512 > gpurun_out/trace_cross.txt 2>&1; grep -E "O complete|O drained|item start|first QK" gpurun_out/trace_cross.txt

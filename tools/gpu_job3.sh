#!/bin/bash
# Round-2 GPU job 3: attention with mask-free full tiles; polynomial period 4 / none / 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== attention tests"; timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_ulysses_gpu.py -q -m gpu -x > gpurun_out/job3_attn_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/job3_attn_tests.log
echo "== kernel bench"; timeout 300 python tools/kernel_bench.py --what attn --variants 0,1,2 --iters 20 > gpurun_out/job3_kernel_bench.log 2>&1; echo "rc=$?"; grep -E "llb_attn|sdpa" gpurun_out/job3_kernel_bench.log | cut -c1-160
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/job3_bench.json 2> gpurun_out/job3_bench.err; echo "rc=$?"; cut -c1-200 gpurun_out/job3_bench.json
echo "== ncu full"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_fwd --launch-skip 6 -c 1 -f -o gpurun_out/job3_attn_v0 python tools/kernel_bench.py --what attn1 --variants 0 --iters 6 > gpurun_out/job3_ncu.log 2>&1; echo "rc=$?"

#!/bin/bash
# Round-2 GPU job 2: cleaned attention kernel + optimistic-exponent variant: parity, speed, ncu source-level profile.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== attention tests"; timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_ulysses_gpu.py -q -m gpu -x > gpurun_out/job2_attn_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/job2_attn_tests.log
echo "== model tests (default variant, then 32)"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x > gpurun_out/job2_model_v0.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/job2_model_v0.log
LLB_ATTN_VARIANT=32 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_block_teacher_gpu.py -q -m gpu -x -s > gpurun_out/job2_model_v32.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/job2_model_v32.log
echo "== kernel bench"; rm -f gpurun_out/kernel_bench.jsonl; timeout 300 python tools/kernel_bench.py --what attn --variants 0,4,32 --iters 20 > gpurun_out/job2_kernel_bench.log 2>&1; echo "rc=$?"; grep llb_attn gpurun_out/job2_kernel_bench.log | cut -c1-160
echo "== bench v0 / v32"; timeout 600 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/job2_bench_v0.json 2> gpurun_out/job2_bench_v0.err; echo "rc=$?"; cut -c1-200 gpurun_out/job2_bench_v0.json
LLB_ATTN_VARIANT=32 timeout 600 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/job2_bench_v32.json 2> gpurun_out/job2_bench_v32.err; echo "rc=$?"; cut -c1-200 gpurun_out/job2_bench_v32.json
for v in 0 32; do
  echo "== ncu full variant $v"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_fwd --launch-skip 6 -c 1 -f -o gpurun_out/job2_attn_v$v python tools/kernel_bench.py --what attn1 --variants $v --iters 6 > gpurun_out/job2_ncu_v$v.log 2>&1; echo "rc=$?"
done
ls -la gpurun_out/*.ncu-rep

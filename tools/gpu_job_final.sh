#!/bin/bash
# Full GPU suite + smoke + bench (all keys) + the reference arm, as the driver runs them at round end.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full GPU suite"; ( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 ) > gpurun_out/final_suite.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/final_suite.log
echo "== kernel cases (sanitizer stand-in, plain)"; timeout 600 python tools/sanitizer_cases.py > gpurun_out/final_cases.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/final_cases.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/final_smoke.log
echo "== bench (driver flags)"; ( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?"; cut -c1-250 gpurun_out/final_bench.json; tail -4 gpurun_out/final_bench.err
echo "== bench --impl reference (driver flags)"; ( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "rc=$?"; cut -c1-1200 gpurun_out/final_bench_ref.json; tail -4 gpurun_out/final_bench_ref.err
echo "== bench with the CTA-pair attention kernel (informational)"; LLB_ATTN_VARIANT=64 timeout 600 python bench.py --no-cpu-baseline --no-reference-gpu --steps 5 --warmup 3 > gpurun_out/final_bench_v64.json 2> /dev/null; echo "rc=$?"; cut -c1-200 gpurun_out/final_bench_v64.json

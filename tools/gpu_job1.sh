#!/bin/bash
# Round-2 GPU job 1: new parity tests, full suite, bench, sanitizers, row-kernel ncu, kernel micro-bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/job1_gpu.txt 2>&1
echo "== new tests"; timeout 1500 python -m pytest tests/test_block_teacher_gpu.py tests/test_lora_gpu.py tests/test_reference_gpu.py tests/test_long_gpu.py -q -s -m gpu --durations=12 > gpurun_out/job1_new_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/job1_new_tests.log
echo "== rest of the suite"; timeout 1200 python -m pytest tests -q -m gpu --durations=10 --ignore=tests/test_block_teacher_gpu.py --ignore=tests/test_lora_gpu.py --ignore=tests/test_reference_gpu.py --ignore=tests/test_long_gpu.py > gpurun_out/job1_suite.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/job1_suite.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/job1_bench.json 2> gpurun_out/job1_bench.err; echo "rc=$?"; cut -c1-600 gpurun_out/job1_bench.json
echo "== kernel bench"; timeout 300 python tools/kernel_bench.py --what attn,gemm,row --variants 0 > gpurun_out/job1_kernel_bench.log 2>&1; echo "rc=$?"
echo "== row-kernel ncu"; timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct -k regex:"ln_modulate|rmsnorm" -c 90 --csv --log-file gpurun_out/job1_row_ncu.csv python tools/profile_forward.py --forwards 1 --layers 2 > gpurun_out/job1_row_ncu.log 2>&1; echo "rc=$?"
echo "== sanitizers"; LLB_SANITIZER_TIMEOUT=420 tools/run_sanitizers.sh memcheck synccheck racecheck; echo "rc=$?"

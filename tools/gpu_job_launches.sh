#!/bin/bash
# ncu launch lists of round 2: one steady-state forward (eager), and bench.py itself (a window inside the timed step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/profile_forward.py --forwards 2 > gpurun_out/r02_profile_forward_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/r02_profile_forward_plain.log | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"llb|kernel" -s 1670 -c 402 --csv --log-file gpurun_out/r02_launches_one_forward.csv python tools/profile_forward.py --forwards 2 > gpurun_out/r02_profile_forward_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r02_bench_plain_for_ncu.json 2>/dev/null; echo "bench plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 45000 -c 420 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "bench ncu rc=$?"

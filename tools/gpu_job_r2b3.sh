#!/bin/bash
# Full GPU suite + a short bench after a kernel change.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full GPU suite"; ( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 ) > gpurun_out/b3_suite.log 2>&1; echo "rc=$?"; tail -16 gpurun_out/b3_suite.log
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/b3_bench.json 2> gpurun_out/b3_bench.err; echo "rc=$?"; cut -c1-250 gpurun_out/b3_bench.json; tail -3 gpurun_out/b3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b3_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline']['frac'])
PY

#!/bin/bash
# two GPUs after the attention epilogue change: head-parallel single stream (peer-store output path) + bench under the driver's launch line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== P=2 graph + timeline"; timeout 600 $TR --master-port 29611 tools/ulysses_check.py --frames 21 --graph 1 --timeline 1 > gpurun_out/r02b_uly_P2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02b_uly_P2.log | cut -c1-900
echo "== bench --gpus 2 (driver launch line)"; timeout 900 $TR --master-port 29614 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02b_bench_n2.json") if l.startswith("{")][-1]); print("value", d["value"], d["clocks"], "ulysses", json.dumps(d.get("ulysses"))[:500])
except Exception as e: print("parse failed", e)
PY

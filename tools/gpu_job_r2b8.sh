#!/bin/bash
# N GPUs, final build: bench.py under the driver's launch line (independent streams + the `ulysses` key)
N=${1:-4}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 $TR --master-port 29614 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02b_bench_n${N}.json 2> gpurun_out/r02b_bench_n${N}.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02b_bench_n${N}.json") if l.startswith("{")][-1]); print("value", d["value"], d["clocks"], "ulysses", json.dumps(d.get("ulysses"))[:420])
except Exception as e: print("parse failed", e)
PY
tail -3 gpurun_out/r02b_bench_n${N}.err

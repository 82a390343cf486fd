#!/bin/bash
# Head-parallel single stream on N GPUs of one box: parity vs one GPU, timing, per-phase timeline, switch race, fp8, and
# bench.py under the driver's launch line (headline = N independent streams; `ulysses` key = one stream over N GPUs).
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  echo "== emulated-rank kernel tests"; timeout 300 python -m pytest tests/test_ulysses_gpu.py -q -m gpu > gpurun_out/uly_emul.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/uly_emul.log
fi
echo "== P=$N graph + timeline"; timeout 600 $TR --master-port 29611 tools/ulysses_check.py --frames 21 --graph 1 --timeline 1 > gpurun_out/uly_P${N}.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/uly_P${N}.log | cut -c1-1200
echo "== P=$N switch race"; timeout 600 $TR --master-port 29612 tools/ulysses_check.py --frames 12 --graph 1 --switch-race 1 > gpurun_out/uly_P${N}_race.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/uly_P${N}_race.log | cut -c1-600
echo "== P=$N fp8"; timeout 600 $TR --master-port 29613 tools/ulysses_check.py --frames 21 --graph 1 --fp8 1 > gpurun_out/uly_P${N}_fp8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/uly_P${N}_fp8.log | cut -c1-600
echo "== bench --gpus $N (driver launch line)"; timeout 900 $TR --master-port 29614 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_n${N}.json; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}.json") if l.startswith("{")][-1]); print("value", d["value"], "ulysses", json.dumps(d.get("ulysses"))[:700])
except Exception as e: print("parse failed", e)
PY

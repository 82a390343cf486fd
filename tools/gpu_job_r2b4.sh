#!/bin/bash
# A/B on one box under the power cap: share of exp2 evaluated on the FMA pipe in the attention kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/b4_$name.json 2> gpurun_out/b4_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/b4_{sys.argv[1]}.json'))
print(sys.argv[1], round(d['value'], 2), 'FPS', d['clocks']['sm_mhz'], 'MHz', round(d['value'] / d['clocks']['sm_mhz'] * 1000, 2), 'FPS/GHz', 'steady', round(d['config']['steady_state_video_fps'], 2), 'attn TF', round(d['roofline']['achieved']))
PY
}
run poly4 LLB_ATTN_POLY=4
run poly0 LLB_ATTN_POLY=0
run poly8 LLB_ATTN_POLY=8
run poly2 LLB_ATTN_POLY=2
run poly4b LLB_ATTN_POLY=4
run poly0b LLB_ATTN_POLY=0

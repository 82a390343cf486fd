#!/bin/bash
# A/B on one box under the power cap: tile mode of the N = 1536 GEMMs (o, cross-q, cross-o, ffn.2) inside the full pipeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/b4_$name.json 2> gpurun_out/b4_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/b4_{sys.argv[1]}.json'))
print(sys.argv[1], round(d['value'], 2), 'FPS', d['clocks']['sm_mhz'], 'MHz', round(d['value'] / d['clocks']['sm_mhz'] * 1000, 2), 'FPS/GHz', 'steady', round(d['config']['steady_state_video_fps'], 2))
PY
}
run base LLB_X=0
run pair256 LLB_GEMM_TILE_1536=1,256
run pair192 LLB_GEMM_TILE_1536=1,192
run single256 LLB_GEMM_TILE_1536=0,256
run single128 LLB_GEMM_TILE_1536=0,128
run base2 LLB_X=0

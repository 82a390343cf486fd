#!/bin/bash
# A/B: mbarrier waits as test_wait spin loops (LLB_MBAR_SPIN build) vs try_wait
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for L in longlive_b200/libllb200.so longlive_b200/libllb200_spin.so; do
  echo "== $L"; LLB200_LIB=$L timeout 300 python tools/kernel_bench.py --what attn --variants 0 --iters 20 2>/dev/null | grep llb_attn | cut -c1-160
done
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/b4_$name.json 2> gpurun_out/b4_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/b4_{sys.argv[1]}.json'))
print(sys.argv[1], round(d['value'], 2), 'FPS', d['clocks']['sm_mhz'], 'MHz', round(d['value'] / d['clocks']['sm_mhz'] * 1000, 2), 'FPS/GHz', 'steady', round(d['config']['steady_state_video_fps'], 2), 'attn TF', round(d['roofline']['achieved']))
PY
}
run base LLB_X=0
run spin LLB200_LIB=longlive_b200/libllb200_spin.so
run base2 LLB_X=0
run spin2 LLB200_LIB=longlive_b200/libllb200_spin.so

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_reference_gpu.py tests/test_attn_gpu.py tests/test_ulysses_gpu.py -q -m gpu -s > gpurun_out/job11_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/job11_tests.log; grep "rel-L2" gpurun_out/job11_tests.log | cut -c1-300
for ms in 16 4 2; do echo "== min split $ms"; LLB_ATTN_MIN_SPLIT=$ms timeout 300 python tools/kernel_bench.py --what attn --variants 0 --iters 30 2>&1 | grep llb_attn | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['shape'], round(d['ms'],4), round(d['tflops']))"; done

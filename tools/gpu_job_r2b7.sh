#!/bin/bash
# stream-K GEMM: parity tests, per-shape timing (classic vs stream-K per tile mode), then the pipeline A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LLB_WAIT_TIMEOUT_NS=300000000
timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q -k "stream_k" > gpurun_out/b7_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/b7_tests.log
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/b7_sk_bench.txt
import math, os, sys, time, torch
sys.path.insert(0, '.')
from longlive_b200 import ops
DEV='cuda'; bf=torch.bfloat16
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters
for (M,N,K,name) in [(4680,1536,1536,'o/cross'),(4680,1536,8960,'ffn2'),(4680,4608,1536,'qkv'),(4680,8960,1536,'ffn1'),(18720,1536,1536,'o recache'),(18720,1536,8960,'ffn2 recache')]:
    a=torch.randn(M,K,device=DEV,dtype=bf); w=torch.randn(N,K,device=DEV,dtype=bf)/math.sqrt(K); b=torch.randn(N,device=DEV,dtype=bf)
    x=torch.randn(M,N,device=DEV,dtype=bf); gate=torch.randn(3,N,device=DEV,dtype=bf); out=torch.empty(M,N,device=DEV,dtype=bf)
    fl=2.0*M*N*K
    kw=dict(epilogue=ops.EPI_BIAS_GATE_RES, gate=gate, rows_per_gate=M//3, res=x) if N==1536 else (dict(epilogue=ops.EPI_BIAS_GELU) if N==8960 else dict())
    res=[]
    for mode in ['auto','0,192','1,256','1,256,1','1,192,1','0,192,1','0,256,1','1,128,1']:
        if mode=='auto': os.environ.pop('LLB_GEMM_TILE',None)
        else: os.environ['LLB_GEMM_TILE']=mode
        ms=timeit(lambda: ops.gemm(a,w,b,out=out,**kw))
        res.append(f"{mode}: {ms*1e3:.1f} us {fl/ms/1e9:.0f} TF")
    os.environ.pop('LLB_GEMM_TILE',None)
    ms=timeit(lambda: torch.addmm(b,a,w.t(),out=out))
    print(name, (M,N,K), ' | '.join(res), f"| cublas {ms*1e3:.1f} us {fl/ms/1e9:.0f} TF", flush=True)
PY
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/b7_$name.json 2> gpurun_out/b7_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/b7_{sys.argv[1]}.json'))
print(sys.argv[1], round(d['value'], 2), 'FPS', d['clocks']['sm_mhz'], 'MHz', round(d['value'] / d['clocks']['sm_mhz'] * 1000, 2), 'FPS/GHz', 'steady', round(d['config']['steady_state_video_fps'], 2))
PY
}
run sk_on LLB_X=0
run sk_off LLB_GEMM_STREAMK=0
run sk_on2 LLB_X=0
run sk_off2 LLB_GEMM_STREAMK=0

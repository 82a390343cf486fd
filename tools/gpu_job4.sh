#!/bin/bash
# Round-2 GPU job 4: CTA-pair attention kernel, first contact
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smallest case"; timeout 120 python - > gpurun_out/job4_first.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from longlive_b200 import ops
from tests.test_attn_gpu import _attn_ref, rel_l2
torch.manual_seed(0)
for (Lq, H, rows, segs) in [(256, 1, 128, [(0, 128)]), (256, 1, 256, [(0, 256)]), (256, 1, 384, [(0, 384)]), (256, 1, 1024, [(0, 1024)]), (200, 2, 300, [(0, 300)]), (130, 1, 1000, [(5, 77), (300, 129)]), (4680, 12, 4680, [(0, 4680)])]:
    q = torch.randn(Lq, H * 128, device='cuda').bfloat16(); k = torch.randn(rows, H * 128, device='cuda').bfloat16(); v = torch.randn(rows, H * 128, device='cuda').bfloat16()
    sp = ops.step_params_tensor(ops.make_step_params(attn_segs=segs), 'cuda')
    out = ops.attention(q, k, v, sp, n_heads=H, variant=64)
    torch.cuda.synchronize()
    ref = _attn_ref(q, k, v, H, segs)
    base = ops.attention(q, k, v, sp, n_heads=H, variant=0)
    print(Lq, H, rows, segs, 'pair vs fp32', rel_l2(out, ref), 'single vs fp32', rel_l2(base, ref), 'finite', bool(torch.isfinite(out.float()).all()), flush=True)
    if rel_l2(out, ref) > 1e-2:
        d = (out.float() - ref).abs()
        print('  worst rows', d.amax(1).topk(5).indices.tolist(), 'worst cols', d.amax(0).topk(8).indices.tolist(), 'rowblock err', [round(float(d[i*128:(i+1)*128].mean()),4) for i in range((Lq+127)//128)][:6], 'colblock err', [round(float(d[:, i*32:(i+1)*32].mean()),4) for i in range(4)], flush=True)
PY
echo "rc=$?"; tail -25 gpurun_out/job4_first.log
echo "== pair tests"; timeout 600 python -m pytest tests/test_attn_gpu.py -q -m gpu -k v64 > gpurun_out/job4_pair_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/job4_pair_tests.log
echo "== kernel bench"; timeout 300 python tools/kernel_bench.py --what attn --variants 0,64 --iters 20 > gpurun_out/job4_kernel_bench.log 2>&1; echo "rc=$?"; grep -E "llb_attn|sdpa|rror" gpurun_out/job4_kernel_bench.log | cut -c1-170

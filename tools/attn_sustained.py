"""Sustained (power-capped) timing of the attention kernel with the SM clock sampled: ms, TFLOP/s, MHz, kilo-cycles per launch.
    [LLB200_LIB=<experiment build>] [V=<variant>] python tools/attn_sustained.py [Lq Lk] [--check]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from longlive_b200 import ops  # noqa: E402
import kernel_bench as kb  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
Lq, Lk = (int(args[0]), int(args[1])) if len(args) >= 2 else (4680, 18720)
H = 12
bf = torch.bfloat16
q = torch.randn(Lq, H * 128, device="cuda", dtype=bf)
k = torch.randn(Lk, H * 128, device="cuda", dtype=bf)
v = torch.randn(Lk, H * 128, device="cuda", dtype=bf)
out = torch.empty_like(q)
sp = ops.step_params_tensor(ops.make_step_params(attn_segs=[(0, Lk)]), "cuda")
fl = 4.0 * Lq * Lk * H * 128
variant = int(os.environ.get("V", "0"))
if "--check" in sys.argv:
    ops.attention(q, k, v, sp, n_heads=H, out=out, variant=variant)
    qs = q.view(1, Lq, H, 128).transpose(1, 2); ks = k.view(1, Lk, H, 128).transpose(1, 2); vs = v.view(1, Lk, H, 128).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(qs, ks, vs).transpose(1, 2).reshape(Lq, H * 128)
    print(json.dumps({"rel_l2_vs_sdpa": ((out.float() - ref.float()).norm() / ref.float().norm()).item()}), flush=True)
for rep in range(3):
    ms, mhz = kb.sustained(lambda: ops.attention(q, k, v, sp, n_heads=H, out=out, variant=variant), seconds=2.0)
    print(json.dumps({"lib": os.environ.get("LLB200_LIB", "default"), "shape": [Lq, Lk], "ms": round(ms, 4), "tflops": round(fl / ms / 1e9),
                      "sm_mhz": mhz, "kcycles": round(ms * mhz)}), flush=True)

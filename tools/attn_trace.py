"""Hand-over timeline of the attention kernel (debug build `make -C longlive_b200/csrc trace`): CTA 0's clock64() stamps
for the first key tiles of its first item, printed relative to tile `--from`.  Run with LLB200_LIB pointing at the trace build:
    LLB200_LIB=longlive_b200/libllb200_trace.so python tools/attn_trace.py --variant 128
"""
import argparse
import ctypes as C
import json
import os

import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("LLB200_LIB", os.path.join(os.path.dirname(__file__), "..", "longlive_b200", "libllb200_trace.so"))
from longlive_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", type=int, default=128)
    ap.add_argument("--lq", type=int, default=4680)
    ap.add_argument("--lk", type=int, default=18720)
    ap.add_argument("--first", type=int, default=20)
    ap.add_argument("--n", type=int, default=6)
    args = ap.parse_args()
    H = 12
    dev = "cuda"
    q = torch.randn(args.lq, H * 128, device=dev, dtype=torch.bfloat16)
    k = torch.randn(args.lk, H * 128, device=dev, dtype=torch.bfloat16)
    v = torch.randn(args.lk, H * 128, device=dev, dtype=torch.bfloat16)
    sp = ops.step_params_tensor(ops.make_step_params(attn_segs=[(0, args.lk)]), dev)
    for _ in range(3):
        ops.attention(q, k, v, sp, n_heads=H, variant=args.variant)
    torch.cuda.synchronize()
    n = 4 * 64 * 16
    buf = (C.c_ulonglong * n)()
    lib = _lib.lib()
    lib.llb_attn_trace_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    rc = lib.llb_attn_trace_read(buf, n)
    assert rc == 0, rc
    t = torch.tensor(list(buf), dtype=torch.int64).view(4, 64, 16)
    names0 = ["wait S", "S ready", None, None, "P full", None, "O complete", "O drained"]
    namesi = ["top", "kv ready", "P0 fired", "PV0+QK0 issued", "P1 fired", "PV1+QK1 issued", "item start", "first QK issued",
              "O free", "issuer enters its loop (kernel start + setup)"]
    short = (args.lk + 127) // 128 <= 8
    tiles = range(0, 16) if short else range(args.first, args.first + args.n)
    t0 = min(int(x) for x in t[:3, tiles.start].flatten() if int(x) > 0)
    rows = []
    for j in tiles:
        for role, nm in ((0, names0), (1, names0), (2, namesi)):
            for sidx, x in enumerate(nm):
                if x and int(t[role, j, sidx]) > 0:
                    lab = f"item {j // 8} tile {j % 8}" if short else f"tile {j:3d}"
                    rows.append((int(t[role, j, sidx]) - t0, lab, ["softmax0", "softmax1", "issuer"][role], x))
    for ts, lab, role, nm in sorted(rows):
        print(f"{ts:8d}  {lab}  {role:9s} {nm}")
    if not short:
        per = (int(t[0, args.first + args.n, 4]) - int(t[0, args.first, 4])) / args.n
        print(json.dumps({"variant": args.variant, "cycles_per_tile_chain0": per}))


if __name__ == "__main__":
    main()

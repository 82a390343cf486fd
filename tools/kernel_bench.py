"""Micro-benchmarks of the libllb200 kernels at the hot-path shapes (CUDA events, GPU box only).

Prints one JSON line per measurement and writes them to gpurun_out/kernel_bench.jsonl.  Library
kernels (cuBLAS via torch.matmul, flash-attn 2, torch SDPA) are timed beside ours as comparators
only; they are never on the product path.
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longlive_b200 import ops  # noqa: E402

DEV = "cuda"


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Event(enable_timing=True); en = torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(iters):
        fn()
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="gemm,attn,row")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--variants", default="0,1")
    args = ap.parse_args()
    os.makedirs("gpurun_out", exist_ok=True)
    outf = open("gpurun_out/kernel_bench.jsonl", "a")

    def emit(d):
        line = json.dumps(d)
        print(line, flush=True)
        outf.write(line + "\n"); outf.flush()

    bf = torch.bfloat16
    if "gemm" in args.what:
        for (M, N, K, name) in [(4680, 4608, 1536, "qkv"), (4680, 1536, 1536, "o/cross"),
                                (4680, 8960, 1536, "ffn1"), (4680, 1536, 8960, "ffn2"),
                                (18720, 4608, 1536, "qkv-recache"), (18720, 8960, 1536, "ffn1-recache")]:
            a = torch.randn(M, K, device=DEV, dtype=bf)
            w = torch.randn(N, K, device=DEV, dtype=bf) / math.sqrt(K)
            b = torch.randn(N, device=DEV, dtype=bf)
            out = torch.empty(M, N, device=DEV, dtype=bf)
            fl = 2.0 * M * N * K
            try:
                ms = timeit(lambda: ops.gemm(a, w, b, out=out), iters=args.iters)
                emit({"kernel": "llb_gemm_bf16", "shape": [M, N, K], "name": name, "ms": ms,
                      "tflops": fl / ms / 1e9})
            except Exception as e:  # keep going: this is a bring-up tool
                emit({"kernel": "llb_gemm_bf16", "shape": [M, N, K], "error": str(e)[:200]})
            try:
                a8, sa = ops.quant_rows_fp8(a)
                w8, sw = ops.quantize_weight_e4m3(w)
                ms = timeit(lambda: ops.gemm_fp8(a8, sa, w8, sw, b, out=out), iters=args.iters)
                emit({"kernel": "llb_gemm_fp8", "shape": [M, N, K], "name": name, "ms": ms,
                      "tflops": fl / ms / 1e9})
            except Exception as e:
                emit({"kernel": "llb_gemm_fp8", "shape": [M, N, K], "error": str(e)[:200]})
            ms = timeit(lambda: torch.addmm(b, a, w.t(), out=out), iters=args.iters)
            emit({"kernel": "cublas_addmm", "shape": [M, N, K], "name": name, "ms": ms,
                  "tflops": fl / ms / 1e9})
    if "attn" in args.what:
        H = 12
        for (Lq, Lk) in [(4680, 4680), (4680, 9360), (4680, 18720), (18720, 18720), (4680, 512)]:
            q = torch.randn(Lq, H * 128, device=DEV, dtype=bf)
            k = torch.randn(Lk, H * 128, device=DEV, dtype=bf)
            v = torch.randn(Lk, H * 128, device=DEV, dtype=bf)
            out = torch.empty_like(q)
            sp = ops.step_params_tensor(ops.make_step_params(attn_segs=[(0, Lk)]), DEV)
            fl = 4.0 * Lq * Lk * H * 128
            for variant in [int(x) for x in args.variants.split(",")]:
                try:
                    ms = timeit(lambda: ops.attention(q, k, v, sp, n_heads=H, out=out, variant=variant),
                                iters=args.iters)
                    emit({"kernel": "llb_attn_fwd", "variant": variant, "shape": [Lq, Lk, H], "ms": ms,
                          "tflops": fl / ms / 1e9})
                except Exception as e:
                    emit({"kernel": "llb_attn_fwd", "variant": variant, "shape": [Lq, Lk, H],
                          "error": str(e)[:200]})
            try:
                from flash_attn import flash_attn_func
                q4 = q.view(1, Lq, H, 128); k4 = k.view(1, Lk, H, 128); v4 = v.view(1, Lk, H, 128)
                ms = timeit(lambda: flash_attn_func(q4, k4, v4), iters=args.iters)
                emit({"kernel": "flash_attn2", "shape": [Lq, Lk, H], "ms": ms, "tflops": fl / ms / 1e9})
            except Exception as e:
                emit({"kernel": "flash_attn2", "error": str(e)[:200]})
            try:
                qs = q.view(1, Lq, H, 128).transpose(1, 2); ks = k.view(1, Lk, H, 128).transpose(1, 2)
                vs = v.view(1, Lk, H, 128).transpose(1, 2)
                ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qs, ks, vs),
                            iters=args.iters)
                emit({"kernel": "torch_sdpa", "shape": [Lq, Lk, H], "ms": ms, "tflops": fl / ms / 1e9})
            except Exception as e:
                emit({"kernel": "torch_sdpa", "error": str(e)[:200]})
    if "row" in args.what:
        rows, Cc, H = 4680, 1536, 12
        x = torch.randn(rows, Cc, device=DEV, dtype=bf)
        mod = torch.randn(3, 6 * Cc, device=DEV, dtype=bf)
        out = torch.empty_like(x)
        ms = timeit(lambda: ops.ln_modulate(x, shift=mod[:, :Cc], scale=mod[:, Cc:2 * Cc],
                                            rows_per_frame=1560, out=out), iters=args.iters)
        emit({"kernel": "ln_modulate", "ms": ms, "GBps": 2 * rows * Cc * 2 / ms / 1e6})
        qkv = torch.randn(rows, 3 * Cc, device=DEV, dtype=bf)
        wq = torch.ones(Cc, device=DEV, dtype=bf)
        kc = torch.zeros(18720, Cc, device=DEV, dtype=bf); vc = torch.zeros_like(kc)
        table = ops.build_rope_table().to(DEV)
        sp = ops.step_params_tensor(ops.make_step_params(0, writes=[(0, 14040, 4680)]), DEV)
        ms = timeit(lambda: ops.rmsnorm_rope_append(qkv, out, kc, vc, wq, wq, table, (30, 52), sp,
                                                    n_heads=H), iters=args.iters)
        emit({"kernel": "rmsnorm_rope_append", "ms": ms, "GBps": 6 * rows * Cc * 2 / ms / 1e6})


if __name__ == "__main__":
    main()

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


def sustained(fn, seconds=0.4):
    """Run fn back to back for `seconds`, sampling the SM clock from another thread: returns
    (ms per call, median SM MHz while running).  Power-capped kernels must be compared at the clock
    they actually ran at."""
    import threading
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    ms1 = timeit(fn, warmup=2, iters=5)
    iters = max(10, int(seconds * 1e3 / ms1))
    clocks, stop = [], threading.Event()

    def sample():
        while not stop.is_set():
            clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            stop.wait(0.02)
    th = threading.Thread(target=sample); th.start()
    ms = timeit(fn, warmup=iters // 4, iters=iters)
    stop.set(); th.join()
    clocks = sorted(clocks[len(clocks) // 3:]) or [0]
    return ms, clocks[len(clocks) // 2]


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
    if "tilesweep" in args.what:
        # every tile mode on the four block GEMM shapes with the epilogue the model uses there
        for (M, N, K, name, epi) in [(4680, 4608, 1536, "qkv", "bias"), (4680, 1536, 1536, "o", "gate_res"),
                                     (4680, 8960, 1536, "ffn1", "gelu"), (4680, 1536, 8960, "ffn2", "gate_res")]:
            a = torch.randn(M, K, device=DEV, dtype=bf)
            w = torch.randn(N, K, device=DEV, dtype=bf) / math.sqrt(K)
            b = torch.randn(N, device=DEV, dtype=bf)
            x = torch.randn(M, N, device=DEV, dtype=bf)
            gate = torch.randn(3, N, device=DEV, dtype=bf)
            out = torch.empty(M, N, device=DEV, dtype=bf)
            a8, sa = ops.quant_rows_fp8(a)
            w8, sw = ops.quantize_weight_e4m3(w)
            fl = 2.0 * M * N * K
            kw = {"bias": dict(), "gelu": dict(epilogue=ops.EPI_BIAS_GELU),
                  "gate_res": dict(epilogue=ops.EPI_BIAS_GATE_RES, gate=gate, rows_per_gate=M // 3, res=x)}[epi]
            for mode in ["auto", "0,128", "0,192", "0,256", "1,128", "1,192", "1,256"]:
                if mode == "auto":
                    os.environ.pop("LLB_GEMM_TILE", None)
                else:
                    os.environ["LLB_GEMM_TILE"] = mode
                for kind in ("bf16", "fp8"):
                    try:
                        if kind == "bf16":
                            ms, mhz = sustained(lambda: ops.gemm(a, w, b, out=out, **kw))
                        else:
                            ms, mhz = sustained(lambda: ops.gemm_fp8(a8, sa, w8, sw, b, out=out, **kw))
                        # tensor-pipe share: nominal 8192 dense bf16 flop/clk/SM (x2 for fp8) at the sampled clock
                        peak = 148 * 8192 * mhz * 1e6 * (2 if kind == "fp8" else 1) / 1e12
                        emit({"kernel": "llb_gemm_" + kind, "name": name, "epi": epi, "tile": mode,
                              "ms": round(ms, 5), "tflops": round(fl / ms / 1e9, 1), "sm_mhz": mhz,
                              "pipe_frac": round(fl / ms / 1e9 / peak, 3) if mhz else None})
                    except Exception as e:
                        emit({"kernel": "llb_gemm_" + kind, "name": name, "tile": mode, "error": str(e)[:200]})
            os.environ.pop("LLB_GEMM_TILE", None)
            ms, mhz = sustained(lambda: torch.addmm(b, a, w.t(), out=out))
            emit({"kernel": "cublas_addmm", "name": name, "epi": "bias", "tile": "cublas", "ms": round(ms, 5),
                  "tflops": round(fl / ms / 1e9, 1), "sm_mhz": mhz,
                  "pipe_frac": round(fl / ms / 1e9 / (148 * 8192 * mhz * 1e6 / 1e12), 3) if mhz else None})
    if "attn" in args.what:
        H = 12
        shapes = [(4680, 18720)] if "attn1" in args.what else [(4680, 4680), (4680, 9360), (4680, 18720), (18720, 18720), (4680, 512)]
        for (Lq, Lk) in shapes:
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
            if "attn1" in args.what:
                continue
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

"""Times the umT5 text encoder (umt5-xxl shape, random init, bf16) on one B200: the CUDA path of
longlive_b200.text_encoder against the reference's op sequence (oracle/t5_oracle.py, eager PyTorch) on the same GPU.

    python tools/t5_bench.py [--tokens 200] [--out gpurun_out/t5_bench.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters):
    fn(); torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(iters):
        fn()
    en.record(); torch.cuda.synchronize()
    return st.elapsed_time(en) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=200)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/t5_bench.json")
    ap.add_argument("--once", action="store_true",
                    help="one warm-up and one eager (no CUDA graph) encode of --tokens tokens, then exit: the run to "
                         "put under `ncu --metrics gpu__time_duration.sum` for a launch list")
    ap.add_argument("--trim", action="store_true", help="with --once: compute only round_up(tokens, 128) rows")
    args = ap.parse_args()
    from longlive_b200 import synth
    from longlive_b200.text_encoder import UMT5Encoder
    from oracle import t5_oracle as to
    dev = "cuda"
    enc = UMT5Encoder(device=dev, dtype=torch.bfloat16)
    synth.random_init_t5_(enc, seed=0, q_gain=32.0, pos_gain=32.0)
    cfg = to.T5Config()
    if args.once:
        enc.use_cuda_graph = False
        ids, mask = to.synth_token_ids(cfg, 7, args.tokens, 1)
        enc(ids, mask, trim_padding=args.trim)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        enc(ids, mask, trim_padding=args.trim)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    res = {"shape": "umt5-xxl encoder: 24 blocks, dim 4096, 64 heads x 64, FFN 10240, vocab 256384; bf16, random init",
           "gflop_per_512_tokens": 24 * 2 * 512 * (4 * 4096 * 4096 + 3 * 4096 * 10240) / 1e9 + 24 * 4 * 512 * 512 * 4096 / 1e9}
    orc = to.T5EncoderOracle(cfg, dict(enc.state_dict()))
    for n in sorted({args.tokens, 77, 512}):
        ids, mask = to.synth_token_ids(cfg, 7, n, 1)
        ids_d, mask_d = ids.to(dev), mask.to(dev)
        ours_trim = timed(lambda: enc(ids, mask), args.iters)
        ours_full = timed(lambda: enc(ids, mask, trim_padding=False), args.iters)
        ref = timed(lambda: orc.text_encoder_forward(ids_d, mask_d), max(2, args.iters // 3))
        a = enc(ids, mask)
        b = orc.text_encoder_forward(ids_d, mask_d)["prompt_embeds"]
        err = ((a[0, :n].float() - b[0, :n].float()).norm() / b[0, :n].float().norm()).item()
        res[f"tokens_{n}"] = {"ours_ms": ours_trim, "ours_all_512_rows_ms": ours_full, "reference_ops_eager_ms": ref,
                              "speedup_vs_eager": ref / ours_trim, "rel_l2_vs_oracle": err,
                              "tflops_all_512_rows": res["gflop_per_512_tokens"] / ours_full}
        print(n, res[f"tokens_{n}"], flush=True)
    res["launches_per_encode"] = 2 + 8 * 24
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

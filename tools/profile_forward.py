"""Profiling driver (run under ncu on the GPU box): full-size model, eager launches (no CUDA graph),
cache filled by 4 clean-context forwards, then `--forwards` steady-state forwards (roll + recompute).
Prints the number of libllb200 kernel launches before / during the steady-state part so that ncu's
--launch-skip / --launch-count can be set to exactly one forward."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longlive_b200 import ops, synth  # noqa: E402
from longlive_b200.model import CausalWanModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--forwards", type=int, default=2)
    ap.add_argument("--layers", type=int, default=30)
    a = ap.parse_args()
    dev = "cuda"
    model = CausalWanModel(local_attn_size=12, sink_size=3, num_layers=a.layers)
    synth.random_init_(model, 0)
    model = model.to(dev).to(torch.bfloat16)
    model.use_cuda_graph = False
    fs, L = 1560, 12 * 1560
    kv = [{"k": torch.zeros(1, L, 12, 128, dtype=torch.bfloat16, device=dev),
           "v": torch.zeros(1, L, 12, 128, dtype=torch.bfloat16, device=dev),
           "global_end_index": torch.zeros(1, dtype=torch.long, device=dev),
           "local_end_index": torch.zeros(1, dtype=torch.long, device=dev)} for _ in range(a.layers)]
    cc = [{"k": torch.zeros(1, 512, 12, 128, dtype=torch.bfloat16, device=dev),
           "v": torch.zeros(1, 512, 12, 128, dtype=torch.bfloat16, device=dev), "is_init": False}
          for _ in range(a.layers)]
    ctx = synth.prompt_embeds(100).to(dev)
    x = synth.latent_noise(0, 3).to(dev).permute(0, 2, 1, 3, 4).contiguous()
    t = torch.full((1, 3), 937.5, device=dev)
    for chunk in range(4):
        model(x, t=t, context=ctx, kv_cache=kv, crossattn_cache=cc, current_start=chunk * 3 * fs)
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    print(f"launches before steady state: {n0}", flush=True)
    for i in range(a.forwards):
        model(x, t=t, context=ctx, kv_cache=kv, crossattn_cache=cc, current_start=4 * 3 * fs)
    torch.cuda.synchronize()
    n1 = ops.launch_count()
    print(f"launches per steady-state forward: {(n1 - n0) // a.forwards}", flush=True)


if __name__ == "__main__":
    main()

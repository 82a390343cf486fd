"""BASELINE configs[2]/[3] at full length on the GPU box: 240 latent frames (80 chunks, rolling
eviction + sink retention), optionally with the 5 prompt switches of example/interactive_example
(switch_frame_indices 40..200) — CUDA pipelines vs the oracle pipeline, rel-L2 of the denoised latents
per chunk.  Writes gpurun_out/drift_240_<mode>.json (copied to profiles/ when committed)."""
import argparse
import json
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wan_oracle as wo  # noqa: E402
from oracle.make_golden import SeededNoise  # noqa: E402
from oracle.pipeline_oracle import run_pipeline  # noqa: E402
from longlive_b200.model import CausalWanModel  # noqa: E402
from longlive_b200.pipeline import InteractiveCausalInferencePipeline  # noqa: E402
from longlive_b200.wrapper import WanDiffusionWrapper  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--mode", default="single", choices=["single", "switch", "switch_global_sink"])
    a = ap.parse_args()
    dev = "cuda"
    cfg = wo.WanConfig()
    sd = wo.init_state_dict(cfg, seed=0)
    model = CausalWanModel(local_attn_size=12, sink_size=3)
    model.load_state_dict(sd)
    gen = WanDiffusionWrapper(model=model.to(dev).to(torch.bfloat16), timestep_shift=5.0)
    ogen = wo.OracleGenerator(wo.OracleModel(cfg, sd).to(dev), shift=5.0)
    switches = [] if a.mode == "single" else [s for s in (40, 80, 120, 160, 200) if s < a.frames]
    gs = a.mode == "switch_global_sink"
    prompts = [wo.synth_prompt_embeds(cfg, 100 + i, 90 + 35 * i).to(dev) for i in range(len(switches) + 1)]
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(1, a.frames, 16, 60, 104, generator=g).to(torch.bfloat16).to(dev)

    class MK(dict):
        __getattr__ = dict.get
    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=3, context_noise=0, global_sink=gs,
                                 model_kwargs=MK(local_attn_size=12, sink_size=3, timestep_shift=5.0))
    pipe = InteractiveCausalInferencePipeline(
        args, torch.device(dev), generator=gen,
        text_encoder=lambda text_prompts: {"prompt_embeds": prompts[int(text_prompts[0])]})
    sn = SeededNoise()
    pipe.renoise_fn = lambda like, b, s: sn(like)
    torch.cuda.synchronize(); t0 = time.time()
    _, lat = pipe.inference(noise, text_prompts_list=[[str(i)] for i in range(len(prompts))],
                            switch_frame_indices=switches, return_latents=True, profile=True)
    torch.cuda.synchronize(); t_ours = time.time() - t0
    sn2 = SeededNoise()
    t0 = time.time()
    olat, okv = run_pipeline(ogen, cfg, noise, prompts, switches, global_sink=gs,
                             renoise=lambda like, b, s: sn2(like))
    torch.cuda.synchronize(); t_oracle = time.time() - t0
    errs = [(((lat[:, c:c + 3].float() - olat[:, c:c + 3].float()).norm() /
              olat[:, c:c + 3].float().norm()).item()) for c in range(0, a.frames, 3)]
    n = len(errs)
    xs = torch.arange(n, dtype=torch.float64); ys = torch.tensor(errs, dtype=torch.float64)
    slope = (((xs - xs.mean()) * (ys - ys.mean())).sum() / ((xs - xs.mean()) ** 2).sum()).item()
    ring = pipe.kv_cache1[0]["_llb_ring"]
    res = {"mode": a.mode, "frames": a.frames, "switch_frames": [s["frame"] for s in pipe.switch_log],
           "rel_l2_per_chunk": errs, "max": max(errs), "mean": sum(errs) / n,
           "mean_first_10": sum(errs[:10]) / min(10, n), "mean_last_10": sum(errs[-10:]) / min(10, n),
           "slope_per_chunk": slope, "global_end": ring.global_end, "local_end": ring.local_end,
           "oracle_global_end": int(okv[0]["global_end_index"].item()),
           "oracle_local_end": int(okv[0]["local_end_index"].item()),
           "seconds_cuda_path_incl_capture": t_ours, "seconds_oracle_gpu": t_oracle,
           "profile": {k: v for k, v in (pipe.last_profile or {}).items() if k != "block_ms"}}
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/drift_{a.frames}_{a.mode}.json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "rel_l2_per_chunk"}))
    assert res["max"] <= 1e-2 and ring.global_end == res["oracle_global_end"]


if __name__ == "__main__":
    main()

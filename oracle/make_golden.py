"""ORACLE support — generates the committed fixtures under tests/golden/ by running the REAL
reference (imported from /root/reference through oracle/ref_shims.py) in the build container.

    python -m oracle.make_golden [--only traces|small|pipeline]

Fixtures (all inputs are regenerated from seeds by tests; only reference OUTPUTS are stored):
  index_trace_*.json     integer cache bookkeeping of CausalWanSelfAttention / _apply_cache_updates
                         for the pipelines' call pattern (spy on the reference model), incl. rolling
                         eviction, recompute calls and KV-recache with global_sink False / True
  small_model.pt         flow predictions + final caches of a 2-layer, 2-head, 24-token/frame model
                         over a call sequence with roll, recompute and recache (reference w/ SDPA)
  pipeline_small.pt      latents produced by the reference's own InteractiveCausalInferencePipeline +
                         WanDiffusionWrapper on a 1-layer 1536-dim model (2 chunks, 1 prompt switch)
                         plus the (current_start, timestep, sink_recache) call log
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims, wan_oracle as wo  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
TRACE_KEYS = ("action", "is_recompute", "current_end", "local_start_index",
              "local_end_index", "write_start_index", "write_end_index")


# ------------------------------------------------------------------------------------------------
def spy_model(model, log):
    """Records what the reference hands to _apply_cache_updates (layer 0) and the indices after."""
    orig = model._apply_cache_updates

    def wrapped(kv_cache, infos):
        _, (current_end, local_end, info) = infos[0]
        rec = {k: (info[k] if isinstance(info[k], (str, bool)) else int(info[k])) for k in TRACE_KEYS}
        rec["num_evicted"] = int(info["num_evicted_tokens"]) if "num_evicted_tokens" in info else None
        rec["num_rolled"] = int(info["num_rolled_tokens"]) if "num_rolled_tokens" in info else None
        rec["current_end"] = int(current_end)
        rec["new_tokens"] = int(info["new_k"].shape[1])
        orig(kv_cache, infos)
        rec["global_end_after"] = int(kv_cache[0]["global_end_index"].item())
        rec["local_end_after"] = int(kv_cache[0]["local_end_index"].item())
        log.append(rec)

    model._apply_cache_updates = wrapped


def call_pattern(num_frames, chunk, switches, local_attn, n_steps=4):
    """The pipelines' generator-call sequence: (current_start_frame, n_frames, sink_recache, kind)."""
    calls, seg, start = [], 0, 0
    while start < num_frames:
        if seg < len(switches) and start >= switches[seg]:
            seg += 1
            if start > 0:
                n = start if local_attn == -1 else min(local_attn, start)
                calls.append((start - n, n, "recache", seg))
        for s in range(n_steps):
            calls.append((start, chunk, f"denoise{s}", seg))
        calls.append((start, chunk, "clean", seg))
        start += chunk
    return calls


def make_traces():
    """Tiny-dim reference model, real cache logic; units are tokens with fs tokens per frame."""
    fs, H, W = 6, 4, 6
    scenarios = {
        "w12_s3_c3_T240_switch": dict(local=12, sink=3, chunk=3, T=240, switches=[40, 80, 120, 160, 200]),
        "w12_s3_c3_T21": dict(local=12, sink=3, chunk=3, T=21, switches=[]),
        "w9_s3_c3_T60": dict(local=9, sink=3, chunk=3, T=60, switches=[20]),
        "w6_s0_c1_T20": dict(local=6, sink=0, chunk=1, T=20, switches=[7]),
        "w10_s1_c1_T30": dict(local=10, sink=1, chunk=1, T=30, switches=[]),
        "w8_s2_c2_T40_early_switch": dict(local=8, sink=2, chunk=2, T=40, switches=[2, 30]),
        "global_c3_T12": dict(local=-1, sink=0, chunk=3, T=12, switches=[6]),
    }
    out = {}
    for name, sc in scenarios.items():
        for global_sink in ((False, True) if sc["switches"] else (False,)):
            cfg = wo.WanConfig(dim=16, ffn_dim=16, num_heads=2, num_layers=1, text_dim=8, text_len=4,
                               local_attn_size=sc["local"], sink_size=sc["sink"], frame_seqlen=fs)
            sd = wo.init_state_dict(cfg, seed=0)
            model = ref_shims.build_reference_model(cfg, sd, "sdpa")
            log = []
            spy_model(model, log)
            size = (sc["local"] if sc["local"] != -1 else sc["T"]) * fs
            kv = wo.new_kv_cache(cfg, 1, size, "cpu")
            cc = wo.new_crossattn_cache(cfg, 1, "cpu")
            ctx = wo.synth_prompt_embeds(cfg, 1, 3)
            calls = call_pattern(sc["T"], sc["chunk"], sc["switches"], sc["local"])
            for (start, n, kind, seg) in calls:
                if kind == "recache":
                    if not global_sink:
                        for c in kv:
                            c["k"].zero_(); c["v"].zero_()
                    for c in cc:
                        c["is_init"] = False
                x = torch.zeros(1, 16, n, H, W, dtype=torch.bfloat16)
                t = torch.zeros(1, n)
                model(x, t=t, context=ctx, seq_len=1 << 20, kv_cache=kv, crossattn_cache=cc,
                      current_start=start * fs,
                      sink_recache_after_switch=(kind == "recache" and not global_sink))
                log[-1]["kind"] = kind
                log[-1]["current_start"] = start * fs
            key = f"{name}_gs{int(global_sink)}"
            out[key] = {"config": {**sc, "frame_seqlen": fs, "cache_size": size, "global_sink": global_sink},
                        "calls": log}
            print(key, len(log), "calls; final G/Le =", log[-1]["global_end_after"], log[-1]["local_end_after"])
    with open(os.path.join(GOLDEN, "index_traces.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


# ------------------------------------------------------------------------------------------------
SMALL_CFG = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16,
                 local_attn_size=4, sink_size=1, frame_seqlen=24)
SMALL_HW = (8, 12)


def small_model_calls():
    """(start_frame, n_frames, t, kind, prompt_seed): fill, roll, recompute, recache (both flavours
    are exercised by running the sequence twice in the tests: sink_recache True here)."""
    calls = []
    for chunk in range(6):
        for t in (1000.0, 833.3333, 0.0):
            calls.append((chunk, 1, t, "gen", 100))
    calls.append((2, 4, 0.0, "recache", 101))
    for chunk in range(6, 8):
        for t in (937.5, 0.0):
            calls.append((chunk, 1, t, "gen", 101))
    return calls


def small_inputs(cfg, call_idx, n):
    g = torch.Generator().manual_seed(5000 + call_idx)
    return torch.randn(1, cfg.in_dim, n, *SMALL_HW, generator=g).to(torch.bfloat16)


def make_small():
    cfg = wo.WanConfig(**SMALL_CFG)
    sd = wo.init_state_dict(cfg, seed=0)
    model = ref_shims.build_reference_model(cfg, sd, "sdpa")
    size = cfg.local_attn_size * cfg.frame_seqlen
    kv = wo.new_kv_cache(cfg, 1, size, "cpu")
    cc = wo.new_crossattn_cache(cfg, 1, "cpu")
    flows = []
    for i, (start, n, t, kind, pseed) in enumerate(small_model_calls()):
        if kind == "recache":
            for c in kv:
                c["k"].zero_(); c["v"].zero_()
            for c in cc:
                c["is_init"] = False
        ctx = wo.synth_prompt_embeds(cfg, pseed, 9)
        x = small_inputs(cfg, i, n)
        f = model(x, t=torch.full((1, n), t), context=ctx, seq_len=1 << 20, kv_cache=kv, crossattn_cache=cc,
                  current_start=start * cfg.frame_seqlen, sink_recache_after_switch=(kind == "recache"))
        flows.append(f.clone())
    torch.save({"flows": flows, "k": [c["k"].clone() for c in kv], "v": [c["v"].clone() for c in kv],
                "global_end": int(kv[0]["global_end_index"].item()),
                "local_end": int(kv[0]["local_end_index"].item())},
               os.path.join(GOLDEN, "small_model.pt"))
    print("small_model.pt written:", len(flows), "forwards")


# ------------------------------------------------------------------------------------------------
PIPE_CFG = dict(dim=1536, ffn_dim=1024, num_heads=12, num_layers=1, text_dim=256, text_len=512,
                local_attn_size=12, sink_size=3, frame_seqlen=1560)


class SeededNoise:
    """Deterministic stand-in for torch.randn_like inside the pipelines: call k draws from seed
    7000 + k on the CPU, so the GPU run of longlive_b200 can consume identical noise."""

    def __init__(self):
        self.k = 0

    def __call__(self, like, *a, **kw):
        g = torch.Generator().manual_seed(7000 + self.k)
        self.k += 1
        return torch.randn(like.shape, generator=g).to(dtype=like.dtype, device=like.device)


def pipe_args(cfg, global_sink=False):
    class MK(dict):
        __getattr__ = dict.get
    return types.SimpleNamespace(
        denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True, num_frame_per_block=3,
        context_noise=0, global_sink=global_sink,
        model_kwargs=MK(local_attn_size=cfg.local_attn_size, sink_size=cfg.sink_size, timestep_shift=5.0))


def make_pipeline():
    cfg = wo.WanConfig(**PIPE_CFG)
    sd = wo.init_state_dict(cfg, seed=0)
    wrapper = ref_shims.build_reference_wrapper(cfg, sd, shift=5.0, attention_impl="sdpa")
    _, Interactive = ref_shims.reference_pipelines()
    prompts = {"a": wo.synth_prompt_embeds(cfg, 200, 77), "b": wo.synth_prompt_embeds(cfg, 201, 120)}

    def text_encoder(text_prompts):
        return {"prompt_embeds": prompts[text_prompts[0]]}

    vae = types.SimpleNamespace(decode_to_pixel=lambda latent, use_cache=False: latent.float())
    pipe = Interactive(pipe_args(cfg), torch.device("cpu"), generator=wrapper, text_encoder=text_encoder, vae=vae)
    pipe.num_transformer_blocks = cfg.num_layers
    log = []
    orig_fwd = wrapper.forward

    def spy(**kw):
        log.append({"current_start": int(kw["current_start"]), "frames": int(kw["noisy_image_or_video"].shape[1]),
                    "t": float(kw["timestep"].flatten()[0]), "t_dtype": str(kw["timestep"].dtype),
                    "sink_recache": bool(kw.get("sink_recache_after_switch", False))})
        return orig_fwd(**kw)

    wrapper.forward = spy
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(1, 6, 16, 60, 104, generator=g).to(torch.bfloat16)
    real_randn_like = torch.randn_like
    torch.randn_like = SeededNoise()
    try:
        _, latents = pipe.inference(noise, text_prompts_list=[["a"], ["b"]], switch_frame_indices=[3],
                                    return_latents=True)
    finally:
        torch.randn_like = real_randn_like
    torch.save({"latents": latents.clone(), "calls": log,
                "global_end": int(pipe.kv_cache1[0]["global_end_index"].item()),
                "local_end": int(pipe.kv_cache1[0]["local_end_index"].item())},
               os.path.join(GOLDEN, "pipeline_small.pt"))
    print("pipeline_small.pt written;", len(log), "generator calls")
    for c in log:
        print("  ", c)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="traces,small,pipeline")
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    if "traces" in a.only:
        make_traces()
    if "small" in a.only:
        make_small()
    if "pipeline" in a.only:
        make_pipeline()

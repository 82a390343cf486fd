"""ORACLE — test infrastructure only.  Functional restatement of the reference's two inference
loops over an OracleGenerator: pipeline/causal_inference.py:146-200 (chunk loop, 4-step DMD loop,
re-noising, clean-context pass) and pipeline/interactive_causal_inference.py:34-106, 237-331
(prompt switch + KV-recache).  Pinned against latents produced by the reference's own pipeline
classes (tests/golden/pipeline_small.pt).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from .wan_oracle import OracleGenerator, WanConfig, new_crossattn_cache, new_kv_cache


def run_pipeline(gen: OracleGenerator, cfg: WanConfig, noise: torch.Tensor, prompt_embeds: Sequence[torch.Tensor],
                 switch_frame_indices: Sequence[int] = (), *, denoising_step_list=(1000, 750, 500, 250),
                 warp: bool = True, frames_per_block: int = 3, context_noise: int = 0,
                 global_sink: bool = False, renoise: Optional[Callable] = None,
                 on_block: Optional[Callable] = None):
    """noise [B,T,16,H,W]; prompt_embeds[i] = [B,text_len,text_dim] for segment i.
    renoise(like, block, step) -> eps.  Returns (latents, kv_cache)."""
    B, T = noise.shape[:2]
    dev = noise.device
    fs = cfg.frame_seqlen
    steps = (gen.scheduler.warped_steps(list(denoising_step_list)) if warp
             else torch.tensor(list(denoising_step_list), dtype=torch.long))
    size = (cfg.local_attn_size if cfg.local_attn_size != -1 else T) * fs
    kv = new_kv_cache(cfg, B, size, dev)
    cc = new_crossattn_cache(cfg, B, dev)
    out = torch.zeros_like(noise)
    seg, start, block = 0, 0, 0
    while start < T:
        if seg < len(switch_frame_indices) and start >= switch_frame_indices[seg]:
            seg += 1
            # _recache_after_switch (interactive_causal_inference.py:34-106)
            if not global_sink:
                for c in kv:
                    c["k"].zero_(); c["v"].zero_()
            for c in cc:
                c["k"] = torch.zeros_like(c["k"]); c["v"] = torch.zeros_like(c["v"]); c["is_init"] = False
            if start > 0:
                n = start if cfg.local_attn_size == -1 else min(cfg.local_attn_size, start)
                frames = out[:, start - n:start]
                t0 = torch.ones([B, n], device=dev, dtype=torch.int64) * context_noise
                gen(frames, prompt_embeds[seg], t0, kv, cc, (start - n) * fs,
                    sink_recache_after_switch=not global_sink)
                for c in cc:
                    c["k"] = torch.zeros_like(c["k"]); c["v"] = torch.zeros_like(c["v"]); c["is_init"] = False
        F = frames_per_block
        x = noise[:, start:start + F]
        cond = prompt_embeds[seg]
        for i, ts in enumerate(steps):
            timestep = torch.ones([B, F], device=dev, dtype=torch.int64) * ts.to(dev)
            _, x0 = gen(x, cond, timestep, kv, cc, start * fs)
            if i < len(steps) - 1:
                flat = x0.flatten(0, 1)
                eps = renoise(flat, block, i) if renoise is not None else torch.randn_like(flat)
                nt = steps[i + 1].to(dev) * torch.ones([B * F], device=dev, dtype=torch.long)
                x = gen.scheduler.add_noise(flat, eps, nt).unflatten(0, x0.shape[:2])
        out[:, start:start + F] = x0
        tz = torch.ones_like(timestep) * context_noise
        gen(x0, cond, tz, kv, cc, start * fs)
        if on_block is not None:
            on_block(block, start, x0, kv)
        start += F
        block += 1
    return out, kv


class DeviceSeededNoise:
    """Deterministic stand-in for torch.randn_like in the pipelines' re-noising step, drawn ON the
    device (call k uses seed 7000 + k), so that long runs are not throttled by host RNG + H2D copies.
    The oracle pipeline and the CUDA pipeline consume identical draws when both run on the same GPU."""

    def __init__(self, device):
        self.k = 0
        self.g = torch.Generator(device=device)

    def __call__(self, like, *a, **kw):
        self.g.manual_seed(7000 + self.k)
        self.k += 1
        return torch.randn(like.shape, generator=self.g, device=like.device, dtype=torch.float32).to(like.dtype)

"""Synthetic weights / inputs for benchmarks and smoke runs (there are no checkpoints offline).

Random init of the named architecture following the distribution of the reference's
CausalWanModel.init_weights (wan/modules/causal_model.py:1265-1287): xavier-uniform Linear weights,
N(0, .02) text/time embedding MLPs, modulation ~ N(0,1)/sqrt(dim); biases get a small N(0, .02)
and the head weight N(0, .02) instead of zeros so every term of the computation is exercised.
"""
from __future__ import annotations

import math

import torch


@torch.no_grad()
def random_init_(model: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    g = torch.Generator(device="cpu").manual_seed(seed)
    for name, p in model.named_parameters():
        shape = p.shape
        if name.endswith("modulation"):
            v = torch.randn(shape, generator=g) / shape[-1] ** 0.5
        elif "norm" in name and name.endswith("weight"):
            v = 1 + 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            v = 0.02 * torch.randn(shape, generator=g)
        elif name.startswith(("text_embedding", "time_embedding", "head.head")):
            v = 0.02 * torch.randn(shape, generator=g)
        else:  # xavier-uniform on the flattened [out, in] view
            fan_out, fan_in = shape[0], math.prod(shape[1:])
            a = math.sqrt(6.0 / (fan_in + fan_out))
            v = (torch.rand(shape, generator=g) * 2 - 1) * a
        p.copy_(v.to(p.dtype))
    return model


def prompt_embeds(seed: int, valid_len: int = 200, text_len: int = 512, text_dim: int = 4096, batch: int = 1,
                  dtype=torch.bfloat16) -> torch.Tensor:
    """Synthetic umT5 embeddings with zeroed padding rows (reference: utils/wan_wrapper.py:52-53)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    e = torch.randn(batch, text_len, text_dim, generator=g)
    e[:, valid_len:] = 0
    return e.to(dtype)


def latent_noise(seed: int, frames: int, batch: int = 1, channels: int = 16, h: int = 60, w: int = 104,
                 dtype=torch.bfloat16) -> torch.Tensor:
    """inference.py:193 — randn([B, T, 16, 60, 104]) for 832x480."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(batch, frames, channels, h, w, generator=g).to(dtype)

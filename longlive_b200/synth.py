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


@torch.no_grad()
def random_init_t5_(encoder: torch.nn.Module, seed: int = 0, q_gain: float = 1.0, pos_gain: float = 1.0):
    """Random umT5 encoder weights with the reference's initialisation (wan/modules/t5.py:27-43): N(0, 1) token
    embedding, N(0, dim^-0.5) k / v / gate / fc1, N(0, (dim * dim_attn)^-0.5) q, N(0, (heads * dim_attn)^-0.5) o,
    N(0, dim_ffn^-0.5) fc2, N(0, (2 * buckets * heads)^-0.5) position tables, unit norm weights; generated on the
    parameters' own device (5.7 G values).  q_gain / pos_gain > 1 make the softmax peaked (stress setting)."""
    dim, dim_attn, dim_ffn = encoder.dim, encoder.dim_attn, encoder.dim_ffn
    std_of = {"token_embedding.weight": 1.0, "attn.q.weight": q_gain * (dim * dim_attn) ** -0.5,
              "attn.k.weight": dim ** -0.5, "attn.v.weight": dim ** -0.5,
              "attn.o.weight": (encoder.num_heads * dim_attn) ** -0.5, "ffn.gate.0.weight": dim ** -0.5,
              "ffn.fc1.weight": dim ** -0.5, "ffn.fc2.weight": dim_ffn ** -0.5,
              "pos_embedding.embedding.weight": pos_gain * (2 * encoder.num_buckets * encoder.num_heads) ** -0.5}
    gens = {}
    for name, p in encoder.named_parameters():
        if "norm" in name:
            p.fill_(1.0)
            continue
        std = next(v for k, v in std_of.items() if name.endswith(k))
        g = gens.setdefault(p.device, torch.Generator(device=p.device).manual_seed(seed))
        p.copy_((torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32) * std).to(p.dtype))
    return encoder

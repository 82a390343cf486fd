"""Checkpoint plumbing for the drop-in model (host-side only).

The reference loads `longlive_base.pt` into WanDiffusionWrapper (keys prefixed "model.", optionally
under "generator" / "generator_ema" / "model" and with FSDP wrapper prefixes, inference.py:72-93) and
then keeps a rank-256 peft LoRA UNMERGED on every nn.Linear inside the attention blocks
(utils/lora_utils.py:19-75, inference.py:100-130), i.e. it pays y = W x + (alpha/r) B (A x) on every
call (+24 % linear FLOPs).  Here LoRA is merged once into dense bf16 weights,
W' = W + (alpha / r) * B @ A, evaluated in fp32 and rounded once to the model dtype; the fused kernels
then see ordinary Linear weights.  Merging changes rounding (one rounding of W' instead of two bf16
matmul outputs added), which is within the per-chunk tolerance but is not bit-identical.
"""
from __future__ import annotations

import re
from typing import Dict, Mapping, Optional

import torch

_WRAPPER_PREFIXES = ("_fsdp_wrapped_module.", "_checkpoint_wrapped_module.", "_orig_mod.", "module.")


def clean_key(name: str) -> str:
    for p in _WRAPPER_PREFIXES:
        name = name.replace(p, "")
    return name


def extract_generator_state_dict(ckpt: Mapping, use_ema: bool = False) -> Dict[str, torch.Tensor]:
    """Returns a state dict keyed like CausalWanModel.state_dict() (no "model." prefix)."""
    if "generator" in ckpt or "generator_ema" in ckpt:
        raw = ckpt["generator_ema" if use_ema and "generator_ema" in ckpt else "generator"]
    elif "model" in ckpt and isinstance(ckpt["model"], Mapping):
        raw = ckpt["model"]
    else:
        raw = ckpt  # already a bare state dict
    out = {}
    for k, v in raw.items():
        k = clean_key(k)
        if k.startswith("model."):
            k = k[len("model."):]
        out[k] = v
    return out


_LORA_RE = re.compile(r"^(?:base_model\.model\.)?(?:model\.)?(?P<mod>.+?)\.lora_(?P<ab>[AB])(?:\.[^.]+)?\.weight$")


def merge_lora(state_dict: Dict[str, torch.Tensor], lora_state_dict: Mapping[str, torch.Tensor],
               alpha: float, rank: Optional[int] = None, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """W' = W + (alpha / r) * B @ A for every module that has a lora_A / lora_B pair.

    Accepts peft key styles ("base_model.model.<mod>.lora_A.weight", "...lora_A.default.weight") and a
    checkpoint wrapped as {"generator_lora": {...}}.  Returns a new dict; inputs are not modified."""
    if "generator_lora" in lora_state_dict:
        lora_state_dict = lora_state_dict["generator_lora"]
    pairs: Dict[str, Dict[str, torch.Tensor]] = {}
    for k, v in lora_state_dict.items():
        m = _LORA_RE.match(clean_key(k))
        if m:
            pairs.setdefault(m.group("mod"), {})[m.group("ab")] = v
    out = dict(state_dict)
    merged = 0
    for mod, ab in pairs.items():
        if "A" not in ab or "B" not in ab:
            raise KeyError(f"LoRA pair incomplete for {mod}")
        key = mod + ".weight"
        if key not in out:
            raise KeyError(f"LoRA targets {key}, which is not in the base state dict")
        A, B = ab["A"].float(), ab["B"].float()
        r = A.shape[0] if rank is None else rank
        out[key] = (out[key].float() + (alpha / r) * (B @ A)).to(dtype)
        merged += 1
    if merged == 0:
        raise ValueError("no lora_A / lora_B pairs found in the LoRA state dict")
    return out


def load_generator_weights(model: torch.nn.Module, ckpt: Mapping, lora: Optional[Mapping] = None,
                           lora_alpha: float = 256.0, lora_rank: Optional[int] = None,
                           use_ema: bool = False, strict: bool = True):
    """Loads a reference checkpoint (and optionally merges its LoRA) into CausalWanModel."""
    sd = extract_generator_state_dict(ckpt, use_ema)
    if lora is not None:
        sd = merge_lora(sd, lora, lora_alpha, lora_rank)
    return model.load_state_dict(sd, strict=strict)

"""Checkpoint key handling and LoRA merge (host-side)."""
import torch

from longlive_b200.checkpoint import extract_generator_state_dict, load_generator_weights, merge_lora
from longlive_b200.model import CausalWanModel
from oracle import wan_oracle as wo


def _small():
    cfg = wo.WanConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16,
                       local_attn_size=4, sink_size=1, frame_seqlen=24)
    return cfg, wo.init_state_dict(cfg, seed=0)


def test_reference_checkpoint_layouts_load_strictly():
    cfg, sd = _small()
    model = CausalWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16,
                           local_attn_size=4, sink_size=1, frame_seqlen=24)
    wrapped = {"generator": {"model._fsdp_wrapped_module." + k: v for k, v in sd.items()},
               "generator_ema": {"model." + k: v * 2 for k, v in sd.items()}}
    out = extract_generator_state_dict(wrapped)
    assert set(out) == set(sd) and torch.equal(out["head.head.weight"], sd["head.head.weight"])
    ema = extract_generator_state_dict(wrapped, use_ema=True)
    assert torch.equal(ema["head.head.bias"], sd["head.head.bias"] * 2)
    res = load_generator_weights(model, wrapped)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(model.blocks[1].ffn[2].weight.data.to(torch.bfloat16), sd["blocks.1.ffn.2.weight"])
    assert set(model.state_dict()) == set(sd), "parameter names must be the reference's"


def test_lora_merge_matches_unmerged_forward():
    cfg, sd = _small()
    g = torch.Generator().manual_seed(1)
    r, alpha = 8, 16.0
    lora = {}
    targets = ["blocks.0.self_attn.q", "blocks.0.ffn.0", "blocks.1.cross_attn.o", "blocks.1.ffn.2"]
    for i, t in enumerate(targets):
        out_f, in_f = sd[t + ".weight"].shape
        style = ("base_model.model." + t + ".lora_{}.weight", "base_model.model.model." + t + ".lora_{}.default.weight")[i % 2]
        lora[style.format("A")] = (torch.randn(r, in_f, generator=g) * 0.05).to(torch.bfloat16)
        lora[style.format("B")] = (torch.randn(out_f, r, generator=g) * 0.05).to(torch.bfloat16)
    merged = merge_lora(sd, {"generator_lora": lora}, alpha=alpha)
    for i, t in enumerate(targets):
        ks = [k for k in lora if t + ".lora_" in k]
        A = next(lora[k] for k in ks if "lora_A" in k).float()
        B = next(lora[k] for k in ks if "lora_B" in k).float()
        W = sd[t + ".weight"].float()
        x = torch.randn(5, W.shape[1], generator=g)
        y_unmerged = x @ W.t() + (alpha / r) * ((x @ A.t()) @ B.t())   # what peft computes per call
        y_merged = x @ merged[t + ".weight"].float().t()
        rel = ((y_merged - y_unmerged).norm() / y_unmerged.norm()).item()
        assert rel < 5e-3, (t, rel)  # one bf16 rounding of W'
    untouched = [k for k in sd if not any(k == t + ".weight" for t in targets)]
    assert all(merged[k] is sd[k] for k in untouched)

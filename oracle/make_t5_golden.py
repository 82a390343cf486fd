"""Generates tests/golden/t5_small.pt by running the REFERENCE umT5 encoder module itself (imported from
/root/reference, which only exists in the build container) on seeded token ids.  The fixture pins
oracle/t5_oracle.py - and through it the CUDA text encoder - to the reference.

    python -m oracle.make_t5_golden
"""
import os

import torch

from oracle import ref_shims
from oracle import t5_oracle as to

SMALL = dict(vocab=1000, dim=256, dim_attn=256, dim_ffn=640, num_heads=4, num_layers=3, num_buckets=32,
             text_len=96)
GAINS = dict(q_gain=8.0, pos_gain=8.0)   # logits of std 4 (peaked softmax), position bias of std 0.5
CASES = ((11, 37, 1), (12, 96, 1), (13, 50, 2))  # (seed, valid tokens, batch)


def load_reference_t5():
    """wan.modules.t5 through the package shells of ref_shims (its class T5EncoderModel evaluates
    torch.cuda.current_device() at definition time, t5.py:478)."""
    import importlib
    ref_shims.install()
    cur = torch.cuda.current_device
    try:
        torch.cuda.current_device = lambda: 0
        return importlib.import_module("wan.modules.t5")
    finally:
        torch.cuda.current_device = cur


def reference_encoder(mod, cfg: to.T5Config, sd, dtype):
    enc = mod.T5Encoder(vocab=cfg.vocab, dim=cfg.dim, dim_attn=cfg.dim_attn, dim_ffn=cfg.dim_ffn,
                        num_heads=cfg.num_heads, num_layers=cfg.num_layers, num_buckets=cfg.num_buckets,
                        shared_pos=False, dropout=0.1)
    enc.load_state_dict(sd, strict=True)
    return enc.to(dtype).eval().requires_grad_(False)


def reference_text_encoder_forward(enc, ids, mask):
    """The body of WanTextEncoder.forward after the tokenizer, executed on the reference encoder."""
    seq_lens = mask.gt(0).sum(dim=1).long()
    context = enc(ids, mask)
    for u, v in zip(context, seq_lens):
        u[v:] = 0.0
    return context


def main():
    mod = load_reference_t5()
    cfg = to.T5Config(**SMALL)
    out = {"cfg": SMALL, "gains": GAINS, "cases": CASES, "seed": 0}
    # integer contract: bucket of every (key - query) offset at the full text length
    emb = mod.T5RelativeEmbedding(32, 64, bidirectional=True)
    rel = torch.arange(512).unsqueeze(0) - torch.arange(512).unsqueeze(1)
    out["buckets_512"] = emb._relative_position_bucket(rel).to(torch.int8)
    for name, dtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        sd = to.init_state_dict(cfg, seed=0, dtype=dtype, **GAINS)
        enc = reference_encoder(mod, cfg, sd, dtype)
        res = []
        with torch.no_grad():
            for seed, n, b in CASES:
                ids, mask = to.synth_token_ids(cfg, seed, n, b)
                res.append(reference_text_encoder_forward(enc, ids, mask))
        out[name] = res
        print(name, [tuple(r.shape) for r in res], [float(r.float().abs().mean()) for r in res])
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "t5_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

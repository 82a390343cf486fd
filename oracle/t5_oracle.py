"""ORACLE — test infrastructure only (imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg; never by the product path).

Functional PyTorch restatement of the reference's umT5 text-encoder path (the step
right before the hot path):

    WanTextEncoder.forward          utils/wan_wrapper.py:43-57   tokenise -> encoder -> zero the padding rows
    T5Encoder.forward               wan/modules/t5.py:287-296    embedding -> 24 blocks -> final norm
    T5SelfAttention.forward         wan/modules/t5.py:163-168    x + attn(norm1(x)); x + ffn(norm2(x)); per-block
                                                                 relative-position table (shared_pos=False, :470)
    T5Attention.forward             wan/modules/t5.py:83-116     no 1/sqrt(d) scaling; bias = position bias, masked
                                                                 keys filled with finfo.min; softmax in fp32
    T5FeedForward.forward           wan/modules/t5.py:132-137    fc2(fc1(x) * gelu_tanh(gate(x)))
    T5LayerNorm.forward             wan/modules/t5.py:57-62      RMS norm, statistics in fp32
    T5RelativeEmbedding             wan/modules/t5.py:230-268    bidirectional log-spaced buckets (32, max_dist 128)
    umt5_xxl                        wan/modules/t5.py:459-472    vocab 256384, dim 4096, ffn 10240, 64 heads, 24 layers

At inference time the whole pipeline is cast to bf16 (inference.py:134, interactive_inference.py:138), so the
encoder runs in bf16 with every op rounding its result; this file keeps the reference's op order and rounding
points so that it is BIT-IDENTICAL to the reference module on the same weights (pinned by
tests/test_t5_oracle_cpu.py against tests/golden/t5_small.pt, which oracle/make_t5_golden.py generates by
running the reference module itself).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F


@dataclass
class T5Config:
    vocab: int = 256384
    dim: int = 4096
    dim_attn: int = 4096
    dim_ffn: int = 10240
    num_heads: int = 64
    num_layers: int = 24
    num_buckets: int = 32
    max_dist: int = 128
    text_len: int = 512
    eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.dim_attn // self.num_heads


def init_state_dict(cfg: T5Config, seed: int = 0, dtype=torch.bfloat16, q_gain: float = 1.0,
                    pos_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Random weights with the reference's initialisation (t5.py:27-43) and parameter names.
    q_gain / pos_gain scale the query projection and the position tables: the stock initialisation gives
    logits of std ~0.02 (a nearly uniform softmax); tests also use large gains to get peaked attention."""
    g = torch.Generator().manual_seed(seed)

    def normal(*shape, std):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    sd = {"token_embedding.weight": normal(cfg.vocab, cfg.dim, std=1.0)}
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = torch.ones(cfg.dim, dtype=dtype)
        sd[p + "attn.q.weight"] = normal(cfg.dim_attn, cfg.dim, std=q_gain * (cfg.dim * cfg.dim_attn) ** -0.5)
        sd[p + "attn.k.weight"] = normal(cfg.dim_attn, cfg.dim, std=cfg.dim ** -0.5)
        sd[p + "attn.v.weight"] = normal(cfg.dim_attn, cfg.dim, std=cfg.dim ** -0.5)
        sd[p + "attn.o.weight"] = normal(cfg.dim, cfg.dim_attn, std=(cfg.num_heads * cfg.dim_attn) ** -0.5)
        sd[p + "norm2.weight"] = torch.ones(cfg.dim, dtype=dtype)
        sd[p + "ffn.gate.0.weight"] = normal(cfg.dim_ffn, cfg.dim, std=cfg.dim ** -0.5)
        sd[p + "ffn.fc1.weight"] = normal(cfg.dim_ffn, cfg.dim, std=cfg.dim ** -0.5)
        sd[p + "ffn.fc2.weight"] = normal(cfg.dim, cfg.dim_ffn, std=cfg.dim_ffn ** -0.5)
        sd[p + "pos_embedding.embedding.weight"] = normal(
            cfg.num_buckets, cfg.num_heads, std=pos_gain * (2 * cfg.num_buckets * cfg.num_heads) ** -0.5)
    sd["norm.weight"] = torch.ones(cfg.dim, dtype=dtype)
    return sd


def synth_token_ids(cfg: T5Config, seed: int, valid_len: int, batch: int = 1):
    """Synthetic tokeniser output: `valid_len` random ids followed by the pad id 0, and the matching mask
    (the umT5 tokenizer pads with id 0 to 512 and returns a prefix mask, tokenizers.py:54-68)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(batch, cfg.text_len, dtype=torch.long)
    mask = torch.zeros(batch, cfg.text_len, dtype=torch.long)
    for b in range(batch):
        n = max(1, min(cfg.text_len, valid_len - 7 * b))
        ids[b, :n] = torch.randint(2, cfg.vocab, (n,), generator=g)
        ids[b, n - 1] = 1  # </s>
        mask[b, :n] = 1
    return ids, mask


# ------------------------------------------------------------------------------------------------
def relative_position_bucket(rel_pos: torch.Tensor, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """Bidirectional bucket of rel_pos = key - query (t5.py:249-268): half of the buckets per sign; inside a
    half, distances below max_exact map to themselves and larger ones logarithmically up to max_dist.  The
    logarithm is evaluated in float32 exactly as the reference does (the bucket is an integer contract)."""
    half = num_buckets // 2
    out = (rel_pos > 0).long() * half
    dist = rel_pos.abs()
    max_exact = half // 2
    large = max_exact + (torch.log(dist.float() / max_exact) / math.log(max_dist / max_exact)
                         * (half - max_exact)).long()
    large = torch.min(large, torch.full_like(large, half - 1))
    return out + torch.where(dist < max_exact, dist, large)


def bucket_table(length: int, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """bucket of (key - query) for every offset in [-(length-1), length-1]: int64 [2*length-1]."""
    d = torch.arange(-(length - 1), length)
    return relative_position_bucket(d, num_buckets, max_dist)


def position_bias(emb: torch.Tensor, lq: int, lk: int, num_buckets: int, max_dist: int) -> torch.Tensor:
    """T5RelativeEmbedding.forward (t5.py:240-247): [1, heads, lq, lk] in the table's dtype."""
    rel = torch.arange(lk, device=emb.device).unsqueeze(0) - torch.arange(lq, device=emb.device).unsqueeze(1)
    b = relative_position_bucket(rel, num_buckets, max_dist)
    return F.embedding(b, emb).permute(2, 0, 1).unsqueeze(0).contiguous()


def t5_layer_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """t5.py:57-62: x (bf16) times an fp32 rsqrt gives fp32; cast to the weight's dtype if that is 16-bit."""
    y = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    if w.dtype in (torch.float16, torch.bfloat16):
        y = y.type_as(w)
    return w * y


def gelu_tanh_chain(x: torch.Tensor) -> torch.Tensor:
    """t5.py:46-50: the tanh approximation spelled out op by op (each op rounds in x's dtype)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class T5EncoderOracle:
    def __init__(self, cfg: T5Config, sd: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.sd = dict(sd)

    def to(self, device):
        self.sd = {k: v.to(device) for k, v in self.sd.items()}
        return self

    def _attention(self, i: int, x: torch.Tensor, mask, bias: torch.Tensor) -> torch.Tensor:
        cfg, sd, p = self.cfg, self.sd, f"blocks.{i}.attn."
        b, n, c = x.size(0), cfg.num_heads, cfg.head_dim
        q = F.linear(x, sd[p + "q.weight"]).view(b, -1, n, c)
        k = F.linear(x, sd[p + "k.weight"]).view(b, -1, n, c)
        v = F.linear(x, sd[p + "v.weight"]).view(b, -1, n, c)
        attn_bias = x.new_zeros(b, n, q.size(1), k.size(1))
        attn_bias += bias
        if mask is not None:
            m = mask.view(b, 1, 1, -1) if mask.ndim == 2 else mask.unsqueeze(1)
            attn_bias.masked_fill_(m == 0, torch.finfo(x.dtype).min)
        logits = torch.einsum("binc,bjnc->bnij", q, k) + attn_bias  # no 1/sqrt(c): T5 folds it into the init
        prob = F.softmax(logits.float(), dim=-1).type_as(logits)
        y = torch.einsum("bnij,bjnc->binc", prob, v).reshape(b, -1, n * c)
        return F.linear(y, sd[p + "o.weight"])

    def _ffn(self, i: int, x: torch.Tensor) -> torch.Tensor:
        sd, p = self.sd, f"blocks.{i}.ffn."
        h = F.linear(x, sd[p + "fc1.weight"]) * gelu_tanh_chain(F.linear(x, sd[p + "gate.0.weight"]))
        return F.linear(h, sd[p + "fc2.weight"])

    @torch.no_grad()
    def encode(self, ids: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
        """T5Encoder.forward (t5.py:287-296), dropout = identity (eval)."""
        cfg, sd = self.cfg, self.sd
        x = F.embedding(ids, sd["token_embedding.weight"])
        L = x.size(1)
        for i in range(cfg.num_layers):
            p = f"blocks.{i}."
            e = position_bias(sd[p + "pos_embedding.embedding.weight"], L, L, cfg.num_buckets, cfg.max_dist)
            x = x + self._attention(i, t5_layer_norm(x, sd[p + "norm1.weight"], cfg.eps), mask, e)
            x = x + self._ffn(i, t5_layer_norm(x, sd[p + "norm2.weight"], cfg.eps))
        return t5_layer_norm(x, sd["norm.weight"], cfg.eps)

    @torch.no_grad()
    def text_encoder_forward(self, ids: torch.Tensor, mask: torch.Tensor) -> Dict[str, torch.Tensor]:
        """WanTextEncoder.forward after the tokenizer (utils/wan_wrapper.py:46-57)."""
        seq_lens = mask.gt(0).sum(dim=1).long()
        context = self.encode(ids, mask)
        for u, v in zip(context, seq_lens):
            u[v:] = 0.0
        return {"prompt_embeds": context}

"""umT5 text encoder on libllb200 — the step right BEFORE the hot path.

Host-side mirror of the reference interface:

    WanTextEncoder            utils/wan_wrapper.py:16-57      forward(text_prompts) -> {"prompt_embeds": [B, 512, 4096]}
    T5Encoder (umt5_xxl)      wan/modules/t5.py:271-296, 459-472   parameter names kept, so that
                              models_t5_umt5-xxl-enc-bf16.pth loads through load_state_dict unchanged
    HuggingfaceTokenizer      wan/modules/tokenizers.py:38-82

The reference runs the encoder in bf16 (the pipeline is cast with `.to(dtype=torch.bfloat16)`, inference.py:134)
as ~15 eager ops per block over all 512 padded positions.  Here one block is seven or eight launches:

    llb_rmsnorm                       norm1                                       (t5.py:57-62)
    llb_gemm_bf16  N = 3 * dim_attn   q | k | v fused, no bias                    (t5.py:92-94)
    llb_t5_attn                       bias + mask + softmax + PV per head, tcgen05 (t5.py:96-111)
    llb_gemm_bf16  + BIAS_RES         o projection, x + y (split-K for <= 256 rows) (t5.py:114, 166)
    llb_rmsnorm                       norm2
    llb_gemm_bf16  + GEGLU_BF16       fc1(x) * gelu(gate(x)) in ONE launch: gate | fc1 rows interleaved per 256-wide
                                      tile, gelu = the reference's bf16 op chain  (t5.py:46-50, 125, 133)
    llb_gemm_bf16  + BIAS_RES         fc2, x + y                                  (t5.py:135, 167)

(at <= 256 rows o and fc2 are split-K launches whose reduce launch also applies the norm that follows, so norm2 and the
next block's norm1 are not launches of their own; at 129-256 rows the gated FFN is two launches) plus the embedding gather
and the final norm fused with the zeroing of the padding rows: 170-195 launches for the 24-block encoder, replayed as one
CUDA graph per (batch, rows) shape.  Keys at or beyond a prompt's length have
probability exactly 0 (the reference fills them with finfo.min before an fp32 softmax), so padded rows never
influence valid rows, and WanTextEncoder zeroes them at the end: with `trim_padding` (default) only the first
round_up(max valid length, 128) rows are computed at all.

There is no CPU path: without libllb200.so / a CUDA device every call raises.
"""
from __future__ import annotations

import math
import os
import zlib
from typing import Dict, List, Optional

import torch
from torch import nn

from . import ops

_ROW_ALIGN = 128  # llb_t5_attn processes 128 query rows per CTA


def relative_position_buckets(max_len: int, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """int32 [2 * max_len - 1]: bucket of offset d = key - query for d = -(max_len-1) .. max_len-1
    (T5RelativeEmbedding._relative_position_bucket, t5.py:249-268, bidirectional).  Half of the buckets serve
    each sign; small distances keep their own bucket, larger ones are spaced logarithmically up to max_dist.
    The logarithm is taken in float32 as the reference does, so the table is bit-identical to it."""
    d = torch.arange(-(max_len - 1), max_len)
    half = num_buckets // 2
    exact = half // 2
    dist = d.abs()
    spaced = exact + (torch.log(dist.float() / exact) / math.log(max_dist / exact) * (half - exact)).long()
    spaced = spaced.clamp(max=half - 1)
    return (torch.where(dist < exact, dist, spaced) + (d > 0).long() * half).to(torch.int32)


class _Weight(nn.Module):
    """Parameter holder named like nn.Linear(bias=False) / nn.Embedding / T5LayerNorm."""

    def __init__(self, *shape, ones: bool = False, **kw):
        super().__init__()
        init = torch.ones(*shape, **kw) if ones else torch.empty(*shape, **kw)
        self.weight = nn.Parameter(init, requires_grad=False)


class _T5Attention(nn.Module):
    def __init__(self, dim, dim_attn, **kw):
        super().__init__()
        self.q, self.k, self.v = _Weight(dim_attn, dim, **kw), _Weight(dim_attn, dim, **kw), _Weight(dim_attn, dim, **kw)
        self.o = _Weight(dim, dim_attn, **kw)


class _T5FeedForward(nn.Module):
    def __init__(self, dim, dim_ffn, **kw):
        super().__init__()
        self.gate = nn.Sequential(_Weight(dim_ffn, dim, **kw))  # reference: Sequential(Linear, GELU) -> "gate.0.weight"
        self.fc1 = _Weight(dim_ffn, dim, **kw)
        self.fc2 = _Weight(dim, dim_ffn, **kw)


class _T5RelativeEmbedding(nn.Module):
    def __init__(self, num_buckets, num_heads, **kw):
        super().__init__()
        self.embedding = _Weight(num_buckets, num_heads, **kw)


class _T5Block(nn.Module):
    def __init__(self, dim, dim_attn, dim_ffn, num_heads, num_buckets, **kw):
        super().__init__()
        self.norm1 = _Weight(dim, ones=True, **kw)
        self.attn = _T5Attention(dim, dim_attn, **kw)
        self.norm2 = _Weight(dim, ones=True, **kw)
        self.ffn = _T5FeedForward(dim, dim_ffn, **kw)
        self.pos_embedding = _T5RelativeEmbedding(num_buckets, num_heads, **kw)  # shared_pos=False (t5.py:470)


class UMT5Encoder(nn.Module):
    """T5Encoder(shared_pos=False) with the umt5-xxl defaults (t5.py:459-472)."""

    def __init__(self, vocab: int = 256384, dim: int = 4096, dim_attn: int = 4096, dim_ffn: int = 10240,
                 num_heads: int = 64, num_layers: int = 24, num_buckets: int = 32, max_dist: int = 128,
                 text_len: int = 512, eps: float = 1e-6, device=None, dtype=None):
        super().__init__()
        kw = {"device": device, "dtype": dtype}  # allocate the 5.7 G parameters where / as they will be used
        if dim_attn // num_heads != 64 or dim_attn % num_heads:
            raise ValueError("llb_t5_attn is built for head_dim 64 (umT5)")
        self.vocab, self.dim, self.dim_attn, self.dim_ffn = vocab, dim, dim_attn, dim_ffn
        self.num_heads, self.num_layers, self.num_buckets, self.max_dist = num_heads, num_layers, num_buckets, max_dist
        self.text_len, self.eps = text_len, eps
        self.token_embedding = _Weight(vocab, dim, **kw)
        self.blocks = nn.ModuleList(
            [_T5Block(dim, dim_attn, dim_ffn, num_heads, num_buckets, **kw) for _ in range(num_layers)])
        self.norm = _Weight(dim, ones=True, **kw)
        self.use_cuda_graph = True
        self.trim_padding = True
        self.kernel_launches = 0
        self._packed, self._bufs, self._graphs = None, {}, {}

    # -- weights ---------------------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        self._packed, self._bufs, self._graphs = None, {}, {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed, self._graphs = None, {}
        return super().load_state_dict(*a, **k)

    def _pack(self):
        dev = self.token_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("longlive_b200.UMT5Encoder runs on CUDA only (no CPU fallback); call .to('cuda') first")
        c = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        P = {"emb": c(self.token_embedding.weight), "norm": c(self.norm.weight), "layers": []}
        for blk in self.blocks:
            P["layers"].append({
                "n1": c(blk.norm1.weight), "n2": c(blk.norm2.weight),
                "qkv": c(torch.cat([blk.attn.q.weight, blk.attn.k.weight, blk.attn.v.weight], 0)),
                "o": c(blk.attn.o.weight), "gate": c(blk.ffn.gate[0].weight), "fc1": c(blk.ffn.fc1.weight),
                "fc2": c(blk.ffn.fc2.weight), "pos": c(blk.pos_embedding.embedding.weight),
                # gate | fc1 as ONE weight in the tile layout of EPI_GEGLU_BF16 (needs dim_ffn % 128 == 0)
                "gf": ops.geglu_weight(c(blk.ffn.gate[0].weight), c(blk.ffn.fc1.weight))
                if self.dim_ffn % 128 == 0 else None,
            })
        # offsets up to the padded row count occur (padding rows / masked keys), so the table covers those too
        P["lut"] = relative_position_buckets(self._rows_max(), self.num_buckets, self.max_dist).to(dev)
        self._packed = P
        return P

    def _rows_max(self) -> int:
        return -(-self.text_len // _ROW_ALIGN) * _ROW_ALIGN

    def _workspace(self, B: int, Lp: int, dev):
        key = (B, Lp)
        b = self._bufs.get(key)
        if b is None:
            z = lambda *s, dt=torch.bfloat16: torch.zeros(*s, dtype=dt, device=dev)
            b = {"ids": z(B, self._rows_max(), dt=torch.int64), "lens": z(B, dt=torch.int32),
                 "x": z(B * Lp, self.dim), "xn": z(B * Lp, self.dim), "qkv": z(B * Lp, 3 * self.dim_attn),
                 "att": z(B * Lp, self.dim_attn), "g": z(B * Lp, self.dim_ffn), "h": z(B * Lp, self.dim_ffn),
                 "out": z(B, self.text_len, self.dim),
                 # short prompts: the two N = dim projections have too few tiles to pull their weights through the
                 # whole GPU, so their K range is split in two (fp32 partials, llb_gemm_bf16_splitk)
                 "ws": z(2 * B * Lp * self.dim, dt=torch.float32) if B * Lp <= 256 else None}
            self._bufs = {key: b} if len(self._bufs) > 4 else {**self._bufs, key: b}
        return b

    # -- one encoder pass on static buffers (graph-capturable) -------------------------------------
    def _run(self, b: dict, B: int, Lp: int, rows_out: int, zero_padding: bool):
        P = self._packed
        x, xn, qkv, att, g, h = b["x"], b["xn"], b["qkv"], b["att"], b["g"], b["h"]
        ops.embed_rows(P["emb"], b["ids"], Lp, out=x)
        ws = b["ws"]
        layers = P["layers"]
        normed = False  # xn already holds norm1(x) of the coming block (written by the previous block's fc2 reduce)
        for i, lw in enumerate(layers):
            if not normed:
                ops.rmsnorm(x, lw["n1"], self.eps, out=xn)
            ops.gemm(xn, lw["qkv"], out=qkv)
            ops.t5_attention(qkv, B, self.num_heads, b["lens"], lw["pos"], P["lut"], out=att)
            if ws is not None:  # split-K; its reduce launch also applies norm2
                ops.gemm_splitk(att, lw["o"], ws, 2, res=x, out=x, norm_w=lw["n2"], norm_out=xn, norm_eps=self.eps)
            else:
                ops.gemm(att, lw["o"], epilogue=ops.EPI_BIAS_RES, res=x, out=x)
                ops.rmsnorm(x, lw["n2"], self.eps, out=xn)
            # one launch for gate and fc1, except at 129-256 rows where its 80 pair tiles of 256 x 256 spill into a
            # second wave on 74 SM pairs while the two separate launches (54 tiles of 256 x 192 each) fit one wave each
            if lw["gf"] is not None and not 128 < B * Lp <= 256:
                ops.gemm(xn, lw["gf"], epilogue=ops.EPI_GEGLU_BF16, out=h)
            else:
                ops.gemm(xn, lw["gate"], epilogue=ops.EPI_BIAS_GELU_BF16, out=g)
                ops.gemm(xn, lw["fc1"], epilogue=ops.EPI_BIAS_MUL, res=g, out=h)
            if ws is not None:
                nxt = layers[i + 1]["n1"] if i + 1 < len(layers) else None  # the final norm has its own kernel
                ops.gemm_splitk(h, lw["fc2"], ws, 2, res=x, out=x, norm_w=nxt, norm_out=xn if nxt is not None else None,
                                norm_eps=self.eps)
                normed = nxt is not None
            else:
                ops.gemm(h, lw["fc2"], epilogue=ops.EPI_BIAS_RES, res=x, out=x)
        lens = b["lens"] if zero_padding else b["all_rows"]
        ops.t5_final_norm(x, P["norm"], B, rows_out, lens, self.eps, out=b["out"])

    @torch.no_grad()
    def forward(self, ids: torch.Tensor, mask: Optional[torch.Tensor] = None, *, zero_padding: bool = True,
                trim_padding: Optional[bool] = None) -> torch.Tensor:
        """ids int64 [B, L <= text_len], mask [B, L] (prefix of ones, as the tokenizer returns) -> bf16
        [B, L, dim].  zero_padding=True returns WanTextEncoder's result (rows at or beyond a prompt's length are
        0); zero_padding=False with trim_padding=False returns T5Encoder.forward's values for every row."""
        if self._packed is None:
            self._pack()
        dev = self._packed["emb"].device
        B, L = ids.shape
        if L > self.text_len:
            raise ValueError(f"{L} tokens > text_len {self.text_len}")
        if mask is None:
            mask = torch.ones_like(ids)
        mask_h = mask.detach().to("cpu")
        lens_h = mask_h.gt(0).sum(dim=1)
        if not torch.equal(mask_h.gt(0), torch.arange(L).unsqueeze(0) < lens_h.unsqueeze(1)):
            raise ValueError("UMT5Encoder needs prefix masks (valid tokens first), as HuggingfaceTokenizer returns")
        trim = self.trim_padding if trim_padding is None else trim_padding
        if trim and not zero_padding:
            raise ValueError("trim_padding needs zero_padding: trimmed rows are never computed")
        rows = int(lens_h.max()) if trim else L
        Lp = max(_ROW_ALIGN, -(-rows // _ROW_ALIGN) * _ROW_ALIGN)
        b = self._workspace(B, Lp, dev)
        b["ids"].zero_()
        b["ids"][:, :L].copy_(ids.to(torch.int64), non_blocking=True)
        b["lens"].copy_(lens_h.to(torch.int32), non_blocking=True)
        if not zero_padding and "all_rows" not in b:
            b["all_rows"] = torch.full((B,), L, dtype=torch.int32, device=dev)
        if self.use_cuda_graph:
            gkey = (B, Lp, L, zero_padding)
            g = self._graphs.get(gkey)
            if g is None:
                n0 = ops.launch_count()
                self._run(b, B, Lp, L, zero_padding)  # warm-up: sets kernel attributes outside the capture
                self.kernel_launches += ops.launch_count() - n0
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                n0 = ops.launch_count()
                with torch.cuda.graph(graph):
                    self._run(b, B, Lp, L, zero_padding)
                g = {"graph": graph, "launches": ops.launch_count() - n0}
                self._graphs = {gkey: g} if len(self._graphs) > 8 else {**self._graphs, gkey: g}
            g["graph"].replay()
            self.kernel_launches += g["launches"]
        else:
            n0 = ops.launch_count()
            self._run(b, B, Lp, L, zero_padding)
            self.kernel_launches += ops.launch_count() - n0
        return b["out"][:, :L].clone()


# ------------------------------------------------------------------------------------------------
class HuggingfaceTokenizer:
    """wan/modules/tokenizers.py:38-82 with clean='whitespace': unicode fix-up (ftfy, if installed), double HTML
    unescape, whitespace collapse, then the Hugging Face tokenizer padded / truncated to seq_len."""

    def __init__(self, name: str, seq_len: Optional[int] = None, clean: Optional[str] = None, **kwargs):
        assert clean in (None, "whitespace", "lower", "canonicalize")
        from transformers import AutoTokenizer
        self.name, self.seq_len, self.clean = name, seq_len, clean
        self.tokenizer = AutoTokenizer.from_pretrained(name, **kwargs)
        self.vocab_size = self.tokenizer.vocab_size

    def _clean(self, text: str) -> str:
        import html
        import re
        import string
        try:
            import ftfy
            text = ftfy.fix_text(text)
        except ImportError:  # not in this image; only matters for mojibake input
            pass
        text = html.unescape(html.unescape(text)).strip()
        if self.clean == "canonicalize":
            text = text.replace("_", " ").translate(str.maketrans("", "", string.punctuation)).lower()
        text = re.sub(r"\s+", " ", text).strip()
        return text.lower() if self.clean == "lower" else text

    def __call__(self, sequence, return_mask: bool = False, **kwargs):
        kw = {"return_tensors": "pt"}
        if self.seq_len is not None:
            kw.update(padding="max_length", truncation=True, max_length=self.seq_len)
        kw.update(kwargs)
        if isinstance(sequence, str):
            sequence = [sequence]
        if self.clean:
            sequence = [self._clean(s) for s in sequence]
        enc = self.tokenizer(sequence, **kw)
        return (enc.input_ids, enc.attention_mask) if return_mask else enc.input_ids


class HashTokenizer:
    """SYNTHETIC stand-in for the umT5 sentencepiece tokenizer (its vocabulary files are not available offline):
    one id per whitespace-separated word from a stable hash, then </s> (id 1), padded with id 0 to seq_len.
    Same output contract as HuggingfaceTokenizer(..., return_mask=True); for benchmarks and tests only."""

    def __init__(self, seq_len: int = 512, vocab_size: int = 256384):
        self.seq_len, self.vocab_size = seq_len, vocab_size

    def __call__(self, sequence, return_mask: bool = False, **kwargs):
        if isinstance(sequence, str):
            sequence = [sequence]
        ids = torch.zeros(len(sequence), self.seq_len, dtype=torch.long)
        mask = torch.zeros_like(ids)
        for i, s in enumerate(sequence):
            toks = [2 + zlib.crc32(w.encode("utf-8")) % (self.vocab_size - 2) for w in s.split()][: self.seq_len - 1]
            toks.append(1)
            ids[i, :len(toks)] = torch.tensor(toks)
            mask[i, :len(toks)] = 1
        return (ids, mask) if return_mask else ids


class WanTextEncoder(nn.Module):
    """Drop-in for utils/wan_wrapper.py:16-57.  With no arguments it looks for the reference's checkpoint layout
    (wan_models/Wan2.1-T2V-1.3B/models_t5_umt5-xxl-enc-bf16.pth and .../google/umt5-xxl/); both parts can be
    injected instead (random-init encoder, synthetic tokenizer) since no checkpoint exists offline."""

    CHECKPOINT = "wan_models/Wan2.1-T2V-1.3B/models_t5_umt5-xxl-enc-bf16.pth"
    TOKENIZER = "wan_models/Wan2.1-T2V-1.3B/google/umt5-xxl/"

    def __init__(self, text_encoder: Optional[UMT5Encoder] = None, tokenizer=None,
                 checkpoint_path: Optional[str] = None, tokenizer_path: Optional[str] = None) -> None:
        super().__init__()
        if text_encoder is None:
            path = checkpoint_path or self.CHECKPOINT
            if not os.path.exists(path):
                raise FileNotFoundError(f"umT5 checkpoint {path} not found; pass text_encoder=UMT5Encoder(...)")
            text_encoder = UMT5Encoder()
            text_encoder.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))
            text_encoder = text_encoder.to(torch.bfloat16)
            if torch.cuda.is_available():
                text_encoder = text_encoder.cuda()
        self.text_encoder = text_encoder.eval().requires_grad_(False)
        if tokenizer is None:
            tokenizer = HuggingfaceTokenizer(name=tokenizer_path or self.TOKENIZER, seq_len=text_encoder.text_len,
                                             clean="whitespace")
        self.tokenizer = tokenizer

    @property
    def device(self):
        return torch.cuda.current_device()

    @torch.no_grad()
    def forward(self, text_prompts: List[str]) -> Dict[str, torch.Tensor]:
        ids, mask = self.tokenizer(text_prompts, return_mask=True, add_special_tokens=True)
        return {"prompt_embeds": self.text_encoder(ids, mask, zero_padding=True)}

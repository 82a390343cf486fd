"""CausalWanModel — B200-native drop-in for the reference's wan/modules/causal_model.py::CausalWanModel
on the KV-cache inference path (``forward(kv_cache=..., crossattn_cache=..., current_start=...)``,
reference :907-1068, :1230-1238).

The module tree and parameter names are the reference's (patch_embedding, text_embedding.{0,2},
time_embedding.{0,2}, time_projection.1, blocks.{i}.{modulation, norm3, self_attn.{q,k,v,o,norm_q,
norm_k}, cross_attn.{...}, ffn.{0,2}}, head.{modulation, head}) so a reference state_dict loads
unchanged, and the attributes the pipelines poke (local_attn_size, num_frame_per_block, block_mask,
per-module max_attention_size) exist.  Only forward differs: it is a fixed sequence of libllb200
kernel launches (13 per block) over pre-allocated buffers, optionally replayed as one CUDA graph.

What the reference does per block and what runs here instead:
  norm1 + modulate (:445)                      -> llb_ln_modulate
  q/k/v Linear (:122-126)                      -> one fused-QKV llb_gemm_bf16
  RMSNorm q,k + RoPE + cache clone/roll/insert -> llb_rmsnorm_rope_append writing the KV ring in place
     (:123-124, :208-211, :251-311)               (index math: llb_kv_ring_plan, host integers only)
  cat(sink, window) + flash-attn (:331-360)    -> llb_attn_fwd over physical ring row ranges
  o Linear + gate + residual (:364, :456)      -> llb_gemm_bf16 epilogue GATE_RES (in place on x)
  norm3 (:460)                                 -> llb_ln_modulate (affine mode)
  cross q Linear + RMSNorm (model.py:172)      -> llb_gemm_bf16 + llb_rmsnorm
  cross attention (model.py:189)               -> llb_attn_fwd over the cached text K/V
  cross o Linear + residual (model.py:193,:460)-> llb_gemm_bf16 epilogue RES
  norm2 + modulate (:463-464)                  -> llb_ln_modulate
  ffn.0 + GELU (:406-407)                      -> llb_gemm_bf16 epilogue GELU
  ffn.2 + gate + residual (:408, :467-468)     -> llb_gemm_bf16 epilogue GATE_RES
  _apply_cache_updates (:849-905)              -> nothing to copy; host commits (G, Le, rot)
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import STEP_PARAMS_INT32
from .kv_ring import KvRing, RingPlan

_STATE_KEY = "_llb_ring"       # host ring state, stored in kv_cache[0]
_PARAM_SLOTS = 256             # pinned staging ring for llb_step_params uploads


class _Linear(nn.Module):
    """Parameter holder with nn.Linear's names; the math runs in llb_gemm_bf16."""

    def __init__(self, in_f: int, out_f: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_f, in_f), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(out_f), requires_grad=False)


class _Norm(nn.Module):
    def __init__(self, dim: int, bias: bool = False):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim), requires_grad=False)
        if bias:
            self.bias = nn.Parameter(torch.zeros(dim), requires_grad=False)


class _PatchEmbedding(nn.Module):
    def __init__(self, in_dim, dim, patch):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(dim, in_dim, *patch), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(dim), requires_grad=False)


class CausalWanSelfAttention(nn.Module):
    """Holds q/k/v/o + norm weights and the attributes the pipelines set
    (reference: wan/modules/causal_model.py:63-95)."""

    def __init__(self, dim, num_heads, local_attn_size=-1, sink_size=0, frame_seqlen=1560):
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.local_attn_size, self.sink_size = local_attn_size, sink_size
        self.max_attention_size = 32760 if local_attn_size == -1 else local_attn_size * frame_seqlen
        self.q, self.k, self.v, self.o = (_Linear(dim, dim) for _ in range(4))
        self.norm_q, self.norm_k = _Norm(dim), _Norm(dim)


class WanT2VCrossAttention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.q, self.k, self.v, self.o = (_Linear(dim, dim) for _ in range(4))
        self.norm_q, self.norm_k = _Norm(dim), _Norm(dim)


class CausalWanAttentionBlock(nn.Module):
    def __init__(self, dim, ffn_dim, num_heads, local_attn_size, sink_size, frame_seqlen):
        super().__init__()
        self.self_attn = CausalWanSelfAttention(dim, num_heads, local_attn_size, sink_size, frame_seqlen)
        self.norm3 = _Norm(dim, bias=True)
        self.cross_attn = WanT2VCrossAttention(dim, num_heads)
        self.ffn = nn.ModuleList([_Linear(dim, ffn_dim), nn.Identity(), _Linear(ffn_dim, dim)])
        self.modulation = nn.Parameter(torch.empty(1, 6, dim), requires_grad=False)


class CausalHead(nn.Module):
    def __init__(self, dim, out_dim, patch):
        super().__init__()
        self.head = _Linear(dim, out_dim * math.prod(patch))
        self.modulation = nn.Parameter(torch.empty(1, 2, dim), requires_grad=False)


class CausalWanModel(nn.Module):
    """See module docstring.  Constructor arguments follow the reference (causal_model.py:523-539)."""

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=1536,
                 ffn_dim=8960, freq_dim=256, text_dim=4096, out_dim=16, num_heads=12, num_layers=30,
                 local_attn_size=-1, sink_size=0, qk_norm=True, cross_attn_norm=True, eps=1e-6,
                 frame_seqlen=1560):
        super().__init__()
        if model_type != "t2v" or not qk_norm or not cross_attn_norm or tuple(patch_size) != (1, 2, 2):
            raise NotImplementedError("longlive_b200 covers the LongLive t2v configuration only")
        if dim % num_heads or dim // num_heads != 128:
            raise NotImplementedError("the sm_100a attention kernel is specialised for head_dim 128")
        self.model_type, self.patch_size, self.text_len = model_type, tuple(patch_size), text_len
        self.in_dim, self.dim, self.ffn_dim, self.freq_dim = in_dim, dim, ffn_dim, freq_dim
        self.text_dim, self.out_dim, self.num_heads, self.num_layers = text_dim, out_dim, num_heads, num_layers
        self.local_attn_size, self.sink_size, self.eps = local_attn_size, sink_size, eps
        self.frame_seqlen = frame_seqlen

        self.patch_embedding = _PatchEmbedding(in_dim, dim, self.patch_size)
        self.text_embedding = nn.ModuleList([_Linear(text_dim, dim), nn.Identity(), _Linear(dim, dim)])
        self.time_embedding = nn.ModuleList([_Linear(freq_dim, dim), nn.Identity(), _Linear(dim, dim)])
        self.time_projection = nn.ModuleList([nn.Identity(), _Linear(dim, dim * 6)])
        self.blocks = nn.ModuleList([
            CausalWanAttentionBlock(dim, ffn_dim, num_heads, local_attn_size, sink_size, frame_seqlen)
            for _ in range(num_layers)])
        self.head = CausalHead(dim, out_dim, self.patch_size)

        # attributes the reference pipelines read / write
        self.block_mask = None
        self.num_frame_per_block = 1
        self.independent_first_frame = False
        self.gradient_checkpointing = False

        self.use_cuda_graph = True
        self.attn_variant = int(os.environ.get("LLB_ATTN_VARIANT", "0"))  # llb_attn_fwd variant bits (A/B runs)
        # optional W8A8 (e4m3) linears inside the blocks (q/k/v, o, cross q/o, ffn); set before the
        # first forward or call refresh_weights() afterwards
        self.fp8_linears = False
        self._packed = None           # fused / packed weights
        self._bufs: Dict[tuple, dict] = {}
        self._graphs: Dict[tuple, dict] = {}
        self._param_ring = None
        self._param_slot = 0
        self.kernel_launches = 0      # libllb200 kernels executed on behalf of this model

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _prepare_blockwise_causal_attn_mask(device=None, num_frames=21, frame_seqlen=1560,
                                            num_frame_per_block=1, local_attn_size=-1):
        """The reference builds a flex-attention BlockMask here (causal_model.py:647-701) that the
        KV-cache path never reads; kept as a no-op for call compatibility."""
        return None

    def refresh_weights(self):
        """Drop packed weights, workspaces and captured graphs (after editing parameters / flags)."""
        self._packed, self._bufs, self._graphs = None, {}, {}

    def _apply(self, fn, *a, **k):  # weights moved / cast -> repack lazily
        self._packed, self._bufs, self._graphs = None, {}, {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed, self._graphs = None, {}
        return super().load_state_dict(*a, **k)

    # ------------------------------------------------------------------------------------------
    def _pack(self):
        """Fuse q|k|v (and cross k|v) weights, flatten the conv weight, stack modulations."""
        dev = self.patch_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("longlive_b200.CausalWanModel runs on CUDA only (no CPU fallback); "
                               "call .to('cuda') first")
        bf = torch.bfloat16
        c = lambda t: t.detach().to(device=dev, dtype=bf).contiguous()
        P = {"layers": []}
        P["patch_w"] = c(self.patch_embedding.weight.flatten(1))
        P["patch_b"] = c(self.patch_embedding.bias)
        for nm, ml, idx in (("text", self.text_embedding, (0, 2)), ("time", self.time_embedding, (0, 2))):
            for j, i in enumerate(idx):
                P[f"{nm}{j}_w"], P[f"{nm}{j}_b"] = c(ml[i].weight), c(ml[i].bias)
        P["tproj_w"], P["tproj_b"] = c(self.time_projection[1].weight), c(self.time_projection[1].bias)
        for blk in self.blocks:
            sa, ca = blk.self_attn, blk.cross_attn
            P["layers"].append({
                "qkv_w": c(torch.cat([sa.q.weight, sa.k.weight, sa.v.weight], 0)),
                "qkv_b": c(torch.cat([sa.q.bias, sa.k.bias, sa.v.bias], 0)),
                "nq": c(sa.norm_q.weight), "nk": c(sa.norm_k.weight),
                "o_w": c(sa.o.weight), "o_b": c(sa.o.bias),
                "n3_w": c(blk.norm3.weight), "n3_b": c(blk.norm3.bias),
                "cq_w": c(ca.q.weight), "cq_b": c(ca.q.bias), "cnq": c(ca.norm_q.weight),
                "ckv_w": c(torch.cat([ca.k.weight, ca.v.weight], 0)),
                "ckv_b": c(torch.cat([ca.k.bias, ca.v.bias], 0)), "cnk": c(ca.norm_k.weight),
                "co_w": c(ca.o.weight), "co_b": c(ca.o.bias),
                "f1_w": c(blk.ffn[0].weight), "f1_b": c(blk.ffn[0].bias),
                "f2_w": c(blk.ffn[2].weight), "f2_b": c(blk.ffn[2].bias),
            })
        if self.fp8_linears:
            for lw in P["layers"]:
                for nm in ("qkv", "o", "cq", "co", "f1", "f2"):
                    lw[nm + "_w8"], lw[nm + "_ws"] = ops.quantize_weight_e4m3(lw[nm + "_w"])
        P["mod_all"] = c(torch.stack([b.modulation.reshape(-1) for b in self.blocks], 0))  # [Lyr, 6C]
        P["head_mod"] = c(self.head.modulation.reshape(1, -1))                             # [1, 2C]
        P["head_w"], P["head_b"] = c(self.head.head.weight), c(self.head.head.bias)
        P["rope"] = ops.build_rope_table(self.dim // self.num_heads).to(dev)
        P["cross_segs"] = ops.step_params_tensor(
            ops.make_step_params(attn_segs=[(0, self.text_len)]), dev)
        self._packed = P
        return P

    def _workspace_for(self, B: int, F: int, H: int, W: int, dev):
        key = (B, F, H, W)
        if key not in self._bufs:
            bf = torch.bfloat16
            L = F * (H // 2) * (W // 2)
            R = B * L
            C_, Cf = self.dim, self.ffn_dim
            e = lambda *s: torch.empty(*s, dtype=bf, device=dev)
            self._bufs[key] = {
                "x_in": e(B, self.in_dim, F, H, W), "t_in": torch.empty(B * F, dtype=torch.float32, device=dev),
                "patches": e(R, self.in_dim * 4), "x": e(R, C_), "xm": e(R, C_), "qkv": e(R, 3 * C_),
                "q": e(R, C_), "attn": e(R, C_), "cq": e(R, C_), "h": e(R, Cf),
                "temb": e(B * F, self.freq_dim), "te1": e(B * F, C_), "e": e(B * F, C_), "es": e(B * F, C_),
                "e0": e(B * F, 6 * C_), "mod": e(self.num_layers, B * F, 6 * C_), "e2": e(B * F, 2 * C_),
                "hmod": e(1, B * F, 2 * C_), "y": e(R, self.out_dim * 4), "out": e(B, self.out_dim, F, H, W),
                "params": torch.zeros(STEP_PARAMS_INT32, dtype=torch.int32, device=dev),
            }
            if self.fp8_linears:
                self._bufs[key].update({
                    "a8": torch.empty(R, C_, dtype=torch.uint8, device=dev),
                    "h8": torch.empty(R, Cf, dtype=torch.uint8, device=dev),
                    "sa": torch.empty(R, dtype=torch.float32, device=dev),
                })
        return self._bufs[key]

    # ------------------------------------------------------------------------------------------
    def _ring_of(self, kv_cache: List[dict], frame_seqlen: int) -> KvRing:
        """Host ring state lives in kv_cache[0]; created lazily (one .item() sync) when the cache
        list was allocated by someone else, e.g. the reference's own pipeline."""
        c0 = kv_cache[0]
        ring = c0.get(_STATE_KEY)
        sa = self.blocks[0].self_attn
        if ring is None:
            ring = KvRing(c0["k"].shape[1], sa.sink_size * frame_seqlen, sa.max_attention_size,
                          sa.local_attn_size, int(c0["global_end_index"].item()),
                          int(c0["local_end_index"].item()))
            c0[_STATE_KEY] = ring
        elif "_llb_index_tensor" not in c0:
            # Cache allocated by someone else (the reference's own pipelines): its owner may reset the indices behind
            # our back - StreamingTrainingPipeline.clear_kv_cache zeroes them (pipeline/streaming_training.py:291-305).
            # Re-read them: one sync per forward, in a mode where the caller itself does ~150 (.item() per layer).
            g, l = int(c0["global_end_index"].item()), int(c0["local_end_index"].item())
            if (g, l) != (ring.global_end, ring.local_end):
                if (g, l) != (0, 0):
                    raise RuntimeError(f"kv_cache end indices were changed externally to ({g}, {l}); the ring can only adopt "
                                       "a reset to zero (it keeps the window rotated, not in logical order)")
                ring = KvRing(c0["k"].shape[1], sa.sink_size * frame_seqlen, sa.max_attention_size, sa.local_attn_size)
                c0[_STATE_KEY] = ring
                c0.pop("_llb_published", None)
        # the pipelines may change these between calls (_set_all_modules_max_attention_size)
        ring.cfg.max_attention_size = int(sa.max_attention_size)
        ring.cfg.local_attn_size = int(sa.local_attn_size)  # the module attribute gates the roll (:231)
        return ring

    def _upload_params(self, sp, dst: torch.Tensor):
        if self._param_ring is None:
            self._param_ring = torch.zeros(_PARAM_SLOTS, STEP_PARAMS_INT32, dtype=torch.int32).pin_memory()
            self._param_events = [None] * _PARAM_SLOTS
        i = self._param_slot
        self._param_slot = (i + 1) % _PARAM_SLOTS
        if self._param_events[i] is not None:
            self._param_events[i].synchronize()  # slot still in flight only if the host is >256 forwards ahead
        self._param_ring[i].copy_(torch.frombuffer(bytearray(bytes(sp)), dtype=torch.int32))
        dst.copy_(self._param_ring[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._param_events[i] = ev

    # ------------------------------------------------------------------------------------------
    def _init_cross_cache(self, context: torch.Tensor, crossattn_cache: List[dict]):
        """text_embedding MLP + per-layer cross K/V (reference: causal_model.py:984-989 and
        model.py:174-180).  Runs once per prompt; results are written IN PLACE into the cache tensors
        (the reference rebinds the dict entries instead)."""
        P = self._packed
        B = context.shape[0]
        C_ = self.dim
        ctx_in = context.to(torch.bfloat16)
        if ctx_in.shape[1] < self.text_len:  # reference pads to text_len with zeros (:985-988)
            pad = ctx_in.new_zeros(B, self.text_len - ctx_in.shape[1], ctx_in.shape[2])
            ctx_in = torch.cat([ctx_in, pad], 1)
        ctx_in = ctx_in.reshape(B * self.text_len, self.text_dim).contiguous()
        h = ops.gemm(ctx_in, P["text0_w"], P["text0_b"], epilogue=ops.EPI_BIAS_GELU)
        ctx = ops.gemm(h, P["text1_w"], P["text1_b"])
        for lw, cc in zip(P["layers"], crossattn_cache):
            if cc["is_init"]:
                continue
            kv = ops.gemm(ctx, lw["ckv_w"], lw["ckv_b"])  # [B*T, 2C]
            kdst = cc["k"].view(B * self.text_len, C_)
            vdst = cc["v"].view(B * self.text_len, C_)
            ops.rmsnorm(kv[:, :C_], lw["cnk"], self.eps, out=kdst)
            vdst.copy_(kv[:, C_:])
            cc["is_init"] = True

    # ------------------------------------------------------------------------------------------
    def _run_block(self, i: int, b: dict, kv_cache, crossattn_cache, B: int, F: int, H: int, W: int):
        """Block i (reference CausalWanAttentionBlock.forward, causal_model.py:413-477) as 13 launches,
        in place on the residual stream b["x"]; reads its adaLN rows from b["mod"][i]."""
        P = self._packed
        C_, nh, eps = self.dim, self.num_heads, self.eps
        gh, gw = H // 2, W // 2
        fs = gh * gw
        L = F * fs
        v = self.attn_variant
        x, xm = b["x"], b["xm"]
        f8 = self.fp8_linears
        lw = P["layers"][i]

        def lin(name, a, **kw):
            """One block Linear: bf16 tcgen05 GEMM, or W8A8 (a = (a8, scale)) when fp8_linears."""
            if f8:
                return ops.gemm_fp8(a[0], a[1], lw[name + "_w8"], lw[name + "_ws"], lw[name + "_b"], **kw)
            return ops.gemm(a, lw[name + "_w"], lw[name + "_b"], **kw)

        def ln_in(**kw):
            """LayerNorm(+modulate) producing the next Linear's input (bf16, or e4m3 + row scales)."""
            if f8:
                return ops.ln_modulate_fp8(x, b["a8"], b["sa"], eps=eps, **kw)
            return ops.ln_modulate(x, eps=eps, out=xm, **kw)

        def act_in(t, buf="a8"):
            return ops.quant_rows_fp8(t, b[buf], b["sa"]) if f8 else t

        m = b["mod"][i]  # [B*F, 6C]
        e = [m[:, k * C_:(k + 1) * C_] for k in range(6)]
        lin("qkv", ln_in(shift=e[0], scale=e[1], rows_per_frame=fs), out=b["qkv"])
        kc, vc = kv_cache[i]["k"], kv_cache[i]["v"]
        for bi in range(B):
            rows = slice(bi * L, (bi + 1) * L)
            k2, v2 = kc[bi].view(-1, C_), vc[bi].view(-1, C_)
            ops.rmsnorm_rope_append(b["qkv"][rows], b["q"][rows], k2, v2, lw["nq"], lw["nk"], P["rope"],
                                    (gh, gw), b["params"], n_heads=nh, eps=eps)
            ops.attention(b["q"][rows], k2, v2, b["params"], n_heads=nh, out=b["attn"][rows], variant=v)
        lin("o", act_in(b["attn"]), epilogue=ops.EPI_BIAS_GATE_RES, gate=e[2], rows_per_gate=fs,
            res=x, out=x)
        lin("cq", ln_in(ln_w=lw["n3_w"], ln_b=lw["n3_b"]), out=b["cq"])
        ops.rmsnorm(b["cq"], lw["cnq"], eps, out=b["q"])
        ck, cv = crossattn_cache[i]["k"], crossattn_cache[i]["v"]
        for bi in range(B):
            rows = slice(bi * L, (bi + 1) * L)
            ops.attention(b["q"][rows], ck[bi].view(-1, C_), cv[bi].view(-1, C_), P["cross_segs"],
                          n_heads=nh, out=b["attn"][rows], variant=v)
        lin("co", act_in(b["attn"]), epilogue=ops.EPI_BIAS_RES, res=x, out=x)
        lin("f1", ln_in(shift=e[3], scale=e[4], rows_per_frame=fs), epilogue=ops.EPI_BIAS_GELU, out=b["h"])
        lin("f2", act_in(b["h"], "h8"), epilogue=ops.EPI_BIAS_GATE_RES, gate=e[5], rows_per_gate=fs,
            res=x, out=x)

    def _run_blocks(self, b: dict, kv_cache, crossattn_cache, B: int, F: int, H: int, W: int):
        """The captured part: everything between the static input buffers and the static output."""
        P = self._packed
        C_, nh, eps = self.dim, self.num_heads, self.eps
        gh, gw = H // 2, W // 2
        fs = gh * gw
        L = F * fs
        v = self.attn_variant
        # embeddings (causal_model.py:959-979)
        for bi in range(B):
            ops.patchify(b["x_in"][bi], out=b["patches"][bi * L:(bi + 1) * L])
        ops.gemm(b["patches"], P["patch_w"], P["patch_b"], out=b["x"])
        ops.sinusoidal(b["t_in"], self.freq_dim, out=b["temb"])
        ops.gemm(b["temb"], P["time0_w"], P["time0_b"], epilogue=ops.EPI_BIAS_SILU, out=b["te1"])
        ops.gemm(b["te1"], P["time1_w"], P["time1_b"], out=b["e"])
        ops.silu(b["e"], out=b["es"])
        ops.gemm(b["es"], P["tproj_w"], P["tproj_b"], out=b["e0"])
        ops.modulation_table(P["mod_all"], b["e0"], out=b["mod"])
        for i in range(len(P["layers"])):
            self._run_block(i, b, kv_cache, crossattn_cache, B, F, H, W)
        x, xm = b["x"], b["xm"]
        # head (causal_model.py:497-508): modulation [1,2,C] + e [B,F,1,C]
        b["e2"][:, :C_].copy_(b["e"]); b["e2"][:, C_:].copy_(b["e"])
        ops.modulation_table(P["head_mod"], b["e2"], out=b["hmod"])
        hm = b["hmod"][0]
        ops.ln_modulate(x, shift=hm[:, :C_], scale=hm[:, C_:], rows_per_frame=fs, eps=eps, out=xm)
        ops.gemm(xm, P["head_w"], P["head_b"], out=b["y"])
        for bi in range(B):
            ops.unpatchify(b["y"][bi * L:(bi + 1) * L], self.out_dim, F, H, W, out=b["out"][bi])

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, t=None, context=None, seq_len=None, kv_cache=None, crossattn_cache=None,
                current_start: int = 0, cache_start=None, sink_recache_after_switch=False, **unused):
        """x [B, C_in, F, H, W] (tensor or list of [C_in, F, H, W]); t [B, F]; context
        [B, <=text_len, text_dim] -> flow prediction [B, C_out, F, H, W] (reference :907-1068).
        `cache_start` is accepted and ignored like in the reference (:118-119)."""
        if kv_cache is None:
            raise NotImplementedError("only the KV-cache inference path exists in this build "
                                      "(the reference's _forward_train raises as well, :1102-1103)")
        if isinstance(x, (list, tuple)):
            x = torch.stack(list(x))
        if isinstance(context, (list, tuple)):
            context = torch.stack(list(context))
        if not x.is_cuda:
            raise RuntimeError("longlive_b200 needs CUDA tensors (no CPU fallback)")
        if self._packed is None:
            self._pack()
        B, _, F, H, W = x.shape
        fs = (H // 2) * (W // 2)
        L = F * fs
        assert seq_len is None or L <= seq_len
        assert current_start % fs == 0, "current_start must be frame aligned"
        if current_start // fs + F > ops.ROPE_MAX_POS or H // 2 > ops.ROPE_MAX_POS or W // 2 > ops.ROPE_MAX_POS:
            # the reference's freqs table has 1024 positions per axis (causal_model.py:622-629) and its
            # causal_rope_apply fails on a short slice beyond it (:46-52)
            raise ValueError(f"RoPE position {current_start // fs + F} frames / grid {H // 2}x{W // 2} exceeds the "
                             f"{ops.ROPE_MAX_POS}-position table")
        b = self._workspace_for(B, F, H, W, x.device)

        # --- integer bookkeeping (host only; replaces the reference's .item() round trips)
        ring = self._ring_of(kv_cache, fs)
        plan: RingPlan = ring.plan(current_start, L, sink_recache_after_switch)
        self._upload_params(ring.step_params(plan, current_start // fs), b["params"])
        self.last_plan = plan

        # --- per-prompt work
        if not all(cc["is_init"] for cc in crossattn_cache):
            self._init_cross_cache(context, crossattn_cache)

        b["x_in"].copy_(x.to(torch.bfloat16))
        b["t_in"].copy_(t.reshape(-1).to(torch.float32))

        if self.use_cuda_graph:
            gkey = (B, F, H, W, self.attn_variant,
                    tuple(c["k"].data_ptr() for c in kv_cache), tuple(c["v"].data_ptr() for c in kv_cache),
                    tuple(c["k"].data_ptr() for c in crossattn_cache),
                    tuple(c["v"].data_ptr() for c in crossattn_cache))
            g = self._graphs.get(gkey)
            if g is None:
                # warm-up run (also sets kernel attributes), then capture
                n0 = ops.launch_count()
                self._run_blocks(b, kv_cache, crossattn_cache, B, F, H, W)
                self.kernel_launches += ops.launch_count() - n0
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                n0 = ops.launch_count()
                with torch.cuda.graph(graph):
                    self._run_blocks(b, kv_cache, crossattn_cache, B, F, H, W)
                g = {"graph": graph, "launches": ops.launch_count() - n0}
                self._graphs = {gkey: g} if len(self._graphs) > 8 else {**self._graphs, gkey: g}
            g["graph"].replay()
            self.kernel_launches += g["launches"]
        else:
            n0 = ops.launch_count()
            self._run_blocks(b, kv_cache, crossattn_cache, B, F, H, W)
            self.kernel_launches += ops.launch_count() - n0

        # --- commit indices (reference: _apply_cache_updates :900-904)
        ring.commit(plan)
        self._publish_indices(kv_cache, ring)
        return b["out"].to(x.dtype).clone()

    @torch.no_grad()
    def forward_block(self, i: int, x_tokens: torch.Tensor, e0: torch.Tensor, kv_cache, crossattn_cache,
                      current_start: int, grid, sink_recache_after_switch: bool = False,
                      commit: bool = False) -> torch.Tensor:
        """Block i alone (reference CausalWanAttentionBlock.forward, causal_model.py:413-477) on a
        caller-supplied residual stream: x_tokens [B, L, C], e0 [B, F, 6, C] (time_projection output,
        :979), grid = (F, h, w) in patches.  Cross-attention caches must already be initialised.
        The ring is planned like in forward(); it is committed only when `commit` (one block is not
        a whole forward).  Used for teacher-forced per-block parity (tests/test_block_teacher_gpu.py)."""
        if self._packed is None:
            self._pack()
        B, L, C_ = x_tokens.shape
        F, gh, gw = grid
        assert L == F * gh * gw and C_ == self.dim and all(cc["is_init"] for cc in crossattn_cache)
        b = self._workspace_for(B, F, 2 * gh, 2 * gw, x_tokens.device)
        ring = self._ring_of(kv_cache, gh * gw)
        plan = ring.plan(current_start, L, sink_recache_after_switch)
        self._upload_params(ring.step_params(plan, current_start // (gh * gw)), b["params"])
        self.last_plan = plan
        b["x"].copy_(x_tokens.reshape(B * L, C_).to(torch.bfloat16))
        b["e0"].copy_(e0.reshape(B * F, 6 * C_).to(torch.bfloat16))
        n0 = ops.launch_count()
        ops.modulation_table(self._packed["mod_all"], b["e0"], out=b["mod"])
        self._run_block(i, b, kv_cache, crossattn_cache, B, F, 2 * gh, 2 * gw)
        self.kernel_launches += ops.launch_count() - n0
        if commit:
            ring.commit(plan)
            self._publish_indices(kv_cache, ring)
        return b["x"].view(B, L, C_).clone()

    def _publish_indices(self, kv_cache, ring: KvRing):
        """Keep kv_cache[*]['global_end_index'/'local_end_index'] observable with reference values.
        Pipelines of this package allocate them as views of one [layers, 2] tensor (2 tiny fills);
        foreign caches get per-layer fills."""
        state = (ring.global_end, ring.local_end)
        if kv_cache[0].get("_llb_published") == state:
            return
        kv_cache[0]["_llb_published"] = state
        shared = kv_cache[0].get("_llb_index_tensor")
        if shared is not None:
            shared[:, 0].fill_(ring.global_end)
            shared[:, 1].fill_(ring.local_end)
        else:
            for c in kv_cache:
                c["global_end_index"].fill_(ring.global_end)
                c["local_end_index"].fill_(ring.local_end)

"""Tensor-facing wrappers around the C ABI (include/llb200.h).

PyTorch is used for device memory and streams only: every function takes CUDA bf16 tensors,
passes raw pointers + strides + the current stream to libllb200.so and returns the output tensor.
Nothing here computes on the host or falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (EPI_BIAS, EPI_BIAS_F32, EPI_BIAS_GATE_RES, EPI_BIAS_GELU, EPI_BIAS_GELU_BF16, EPI_BIAS_MUL,
                   EPI_BIAS_RES, EPI_BIAS_SILU, EPI_GEGLU_BF16, STEP_PARAMS_INT32, StepParams)

__all__ = [
    "gemm", "gemm_splitk", "gemm_fp8", "quant_rows_fp8", "ln_modulate_fp8", "quantize_weight_e4m3", "attention", "ln_modulate", "rmsnorm", "rmsnorm_rope_append", "patchify", "unpatchify",
    "sinusoidal", "modulation_table", "silu", "make_step_params", "step_params_tensor",
    "build_rope_table", "launch_count",
    "EPI_BIAS", "EPI_BIAS_GELU", "EPI_BIAS_SILU", "EPI_BIAS_GATE_RES", "EPI_BIAS_RES", "EPI_BIAS_F32",
    "EPI_BIAS_MUL", "EPI_BIAS_GELU_BF16", "EPI_GEGLU_BF16", "geglu_weight", "embed_rows", "t5_attention", "t5_final_norm",
]


ROPE_MAX_POS = 1024  # LLB_ROPE_MAX_POS (include/llb200.h): rows of the (cos, sin) table per axis


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, name: str, dtype=torch.bfloat16) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: longlive_b200 kernels need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise RuntimeError(f"{name}: innermost dimension must be contiguous")


def launch_count() -> int:
    return int(_lib.lib().llb_launch_count())


# ------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
         epilogue: int = EPI_BIAS, out: Optional[torch.Tensor] = None,
         gate: Optional[torch.Tensor] = None, rows_per_gate: int = 0, gate_row0: int = 0,
         res: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T + bias).  a/w/out 2-D bf16 (row stride arbitrary);
    with EPI_BIAS_F32 the output is float32."""
    _req(a, "a"); _req(w, "w")
    M, K = a.shape
    N, K2 = w.shape
    assert K == K2, (a.shape, w.shape)
    out_dtype = torch.float32 if epilogue == EPI_BIAS_F32 else torch.bfloat16
    if out is None:
        out = torch.empty((M, N // 2 if epilogue == EPI_GEGLU_BF16 else N), dtype=out_dtype, device=a.device)
    _req(out, "out", out_dtype)
    for t, n in ((bias, "bias"), (gate, "gate"), (res, "res")):
        if t is not None:
            _req(t, n)
    rc = _lib.lib().llb_gemm_bf16(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0),
        M, N, K, epilogue, _ptr(bias), _ptr(gate), gate.stride(0) if gate is not None else 0,
        rows_per_gate, gate_row0, _ptr(res), res.stride(0) if res is not None else 0, _stream())
    _lib.check(rc, "llb_gemm_bf16")
    return out


def geglu_weight(gate_w: torch.Tensor, fc1_w: torch.Tensor) -> torch.Tensor:
    """Weight layout of EPI_GEGLU_BF16: 256-row tiles of [128 gate rows | the 128 fc1 rows of the same output columns]."""
    F, K = gate_w.shape
    assert fc1_w.shape == (F, K) and F % 128 == 0
    return torch.stack([gate_w.view(F // 128, 128, K), fc1_w.view(F // 128, 128, K)], dim=1).reshape(2 * F, K).contiguous()


def gemm_splitk(a: torch.Tensor, w: torch.Tensor, workspace: torch.Tensor, k_splits: int,
                bias: Optional[torch.Tensor] = None, *, res: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, norm_w: Optional[torch.Tensor] = None,
                norm_out: Optional[torch.Tensor] = None, norm_eps: float = 1e-6) -> torch.Tensor:
    """Split-K GEMM for skinny problems: out = (res +) a @ w^T (+ bias); workspace: float32, >= k_splits*M*N elements.
    With norm_w / norm_out the reduce launch also writes the RMS norm of the result rows into norm_out."""
    _req(a, "a"); _req(w, "w"); _req(workspace, "workspace", torch.float32)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _req(out, "out")
    for t, n in ((bias, "bias"), (res, "res")):
        if t is not None:
            _req(t, n)
    rc = _lib.lib().llb_gemm_bf16_splitk(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), M, N, K,
        EPI_BIAS_RES if res is not None else EPI_BIAS, _ptr(bias), _ptr(res), res.stride(0) if res is not None else 0,
        k_splits, workspace.data_ptr(), workspace.numel() * 4, _ptr(norm_w) if norm_out is not None else None,
        _ptr(norm_out), norm_out.stride(0) if norm_out is not None else 0, C.c_float(norm_eps), _stream())
    _lib.check(rc, "llb_gemm_bf16_splitk")
    return out


def gemm_fp8(a8: torch.Tensor, a_scale: torch.Tensor, w8: torch.Tensor, w_scale: torch.Tensor,
             bias: Optional[torch.Tensor] = None, *, epilogue: int = EPI_BIAS, out: Optional[torch.Tensor] = None,
             gate: Optional[torch.Tensor] = None, rows_per_gate: int = 0, gate_row0: int = 0,
             res: Optional[torch.Tensor] = None) -> torch.Tensor:
    """W8A8 linear: a8 [M,K] / w8 [N,K] e4m3 (torch.float8_e4m3fn or uint8 storage), a_scale [M] and
    w_scale [N] float32; out bf16 [M,N] = epilogue(acc * a_scale[:,None] * w_scale[None,:] + bias)."""
    for t, n in ((a8, "a8"), (w8, "w8")):
        if not t.is_cuda or t.element_size() != 1 or t.stride(-1) != 1:
            raise RuntimeError(f"{n}: expected a 1-byte CUDA tensor with contiguous rows")
    _req(a_scale, "a_scale", torch.float32); _req(w_scale, "w_scale", torch.float32)
    M, K = a8.shape
    N, K2 = w8.shape
    assert K == K2 and a_scale.numel() == M and w_scale.numel() == N
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a8.device)
    rc = _lib.lib().llb_gemm_fp8(
        a8.data_ptr(), a8.stride(0), a_scale.data_ptr(), w8.data_ptr(), w8.stride(0), w_scale.data_ptr(),
        out.data_ptr(), out.stride(0), M, N, K, epilogue, _ptr(bias), _ptr(gate),
        gate.stride(0) if gate is not None else 0, rows_per_gate, gate_row0, _ptr(res),
        res.stride(0) if res is not None else 0, _stream())
    _lib.check(rc, "llb_gemm_fp8")
    return out


def quant_rows_fp8(x: torch.Tensor, out8: Optional[torch.Tensor] = None,
                   out_scale: Optional[torch.Tensor] = None):
    """bf16 [rows, C] -> (e4m3 bytes [rows, C] as uint8, float32 scale [rows])."""
    _req(x, "x")
    rows, Cc = x.shape
    if out8 is None:
        out8 = torch.empty((rows, Cc), dtype=torch.uint8, device=x.device)
    if out_scale is None:
        out_scale = torch.empty(rows, dtype=torch.float32, device=x.device)
    rc = _lib.lib().llb_quant_rows_fp8(x.data_ptr(), x.stride(0), out8.data_ptr(), out8.stride(0),
                                       out_scale.data_ptr(), rows, Cc, _stream())
    _lib.check(rc, "llb_quant_rows_fp8")
    return out8, out_scale


def ln_modulate_fp8(x: torch.Tensor, out8: torch.Tensor, out_scale: torch.Tensor, *,
                    shift: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None,
                    rows_per_frame: int = 0, row0: int = 0, ln_w: Optional[torch.Tensor] = None,
                    ln_b: Optional[torch.Tensor] = None, eps: float = 1e-6):
    _req(x, "x")
    rows, Cc = x.shape
    ld_mod = shift.stride(0) if shift is not None else 0
    rc = _lib.lib().llb_ln_modulate_fp8(
        x.data_ptr(), x.stride(0), out8.data_ptr(), out8.stride(0), out_scale.data_ptr(), rows, Cc,
        _ptr(shift), _ptr(scale), ld_mod, rows_per_frame, row0, _ptr(ln_w), _ptr(ln_b), C.c_float(eps), _stream())
    _lib.check(rc, "llb_ln_modulate_fp8")
    return out8, out_scale


def quantize_weight_e4m3(w: torch.Tensor):
    """Static per-output-channel e4m3 quantisation of a Linear weight [N, K] (done once at load time)."""
    wf = w.float()
    sc = (wf.abs().amax(dim=1).clamp_min(1e-12) / 448.0)
    w8 = (wf / sc[:, None]).to(torch.float8_e4m3fn).view(torch.uint8).contiguous()
    return w8, sc.to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------
def make_step_params(rope_start_frame: int = 0,
                     writes: Sequence[Tuple[int, int, int]] = (),
                     attn_segs: Sequence[Tuple[int, int]] = ()) -> StepParams:
    """writes: (src_row, dst_row, n) triples; attn_segs: (start_row, len) pairs."""
    sp = StepParams()
    sp.rope_start_frame = rope_start_frame
    assert len(writes) <= _lib.LLB_MAX_SEGS and len(attn_segs) <= _lib.LLB_MAX_SEGS
    sp.n_write_segs = len(writes)
    for i, (s, d, n) in enumerate(writes):
        sp.write_src[i], sp.write_dst[i], sp.write_n[i] = s, d, n
    sp.n_attn_segs = len(attn_segs)
    for i, (s, n) in enumerate(attn_segs):
        sp.attn_start[i], sp.attn_len[i] = s, n
    return sp


def step_params_tensor(sp: StepParams, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Uploads a StepParams struct into an int32 device tensor (stream-ordered copy)."""
    host = torch.frombuffer(bytearray(bytes(sp)), dtype=torch.int32).clone()
    if out is None:
        return host.to(device)
    out.copy_(host, non_blocking=False)
    return out


_ATTN_WS = {}


def attention_workspace(device) -> torch.Tensor:
    """Scratch of llb_attn_fwd's split-remainder scheduling (zeroed once; flags are self-resetting).
    One buffer per (device, stream): launches on one stream are serialised, launches on different
    streams must not share partials.  A graph-capture stream gets (and keeps) its own buffer."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    ws = _ATTN_WS.get(key)
    if ws is None:
        n = int(_lib.lib().llb_attn_workspace_bytes())
        ws = torch.zeros(n, dtype=torch.uint8, device=dev)
        _ATTN_WS[key] = ws
    return ws


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, segs_dev: torch.Tensor, *,
              n_heads: int, scale: Optional[float] = None, out: Optional[torch.Tensor] = None,
              variant: int = 0, shard=None) -> torch.Tensor:
    """q [Lq, H*128], k/v [rows, H*128] (any row stride), segs_dev: int32 StepParams tensor.
    shard (head-parallel mode): _lib.OutShard; output rows then go to the peers' buffers."""
    _req(q, "q"); _req(k, "k"); _req(v, "v")
    assert segs_dev.dtype == torch.int32 and segs_dev.numel() >= STEP_PARAMS_INT32 and segs_dev.is_cuda
    Lq = q.shape[0]
    # (head-parallel buffers may be wider than the heads this launch covers: only the first n_heads * 128 columns are read)
    assert q.shape[1] >= n_heads * 128 and k.shape[1] >= n_heads * 128 and v.shape == k.shape
    if out is None:
        out = torch.empty((Lq, n_heads * 128), dtype=torch.bfloat16, device=q.device)
    if scale is None:
        scale = 128 ** -0.5
    ws = attention_workspace(q.device)
    rc = _lib.lib().llb_attn_fwd(
        q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
        out.data_ptr(), out.stride(0), Lq, n_heads, k.shape[0], segs_dev.data_ptr(),
        C.c_float(scale), variant, ws.data_ptr(), ws.numel(),
        C.byref(shard) if shard is not None else None, _stream())
    _lib.check(rc, "llb_attn_fwd")
    return out


# ------------------------------------------------------------------------------------------------
def ln_modulate(x: torch.Tensor, *, shift: Optional[torch.Tensor] = None,
                scale: Optional[torch.Tensor] = None, rows_per_frame: int = 0, row0: int = 0,
                ln_w: Optional[torch.Tensor] = None, ln_b: Optional[torch.Tensor] = None,
                eps: float = 1e-6, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, "x")
    rows, Cc = x.shape
    if out is None:
        out = torch.empty((rows, Cc), dtype=torch.bfloat16, device=x.device)
    ld_mod = 0
    if shift is not None:
        _req(shift, "shift"); _req(scale, "scale")
        assert shift.stride(0) == scale.stride(0)
        ld_mod = shift.stride(0)
    rc = _lib.lib().llb_ln_modulate(
        x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, Cc, _ptr(shift), _ptr(scale),
        ld_mod, rows_per_frame, row0, _ptr(ln_w), _ptr(ln_b), C.c_float(eps), _stream())
    _lib.check(rc, "llb_ln_modulate")
    return out


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, "x"); _req(w, "w")
    rows, Cc = x.shape
    if out is None:
        out = torch.empty((rows, Cc), dtype=torch.bfloat16, device=x.device)
    rc = _lib.lib().llb_rmsnorm(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, Cc,
                                w.data_ptr(), C.c_float(eps), _stream())
    _lib.check(rc, "llb_rmsnorm")
    return out


def rmsnorm_rope_append(qkv: torch.Tensor, q_out: Optional[torch.Tensor], k_cache: Optional[torch.Tensor],
                        v_cache: Optional[torch.Tensor], wq: torch.Tensor, wk: torch.Tensor,
                        rope_cs: torch.Tensor, grid_hw: Tuple[int, int], params_dev: torch.Tensor, *,
                        n_heads: int, eps: float = 1e-6, shard=None) -> Optional[torch.Tensor]:
    """qkv [rows, 3*H*128]; q_out [rows, H*128]; k_cache/v_cache [cache_rows, H*128].
    shard (head-parallel mode): _lib.QkvShard with the peers' Q / K / V base pointers."""
    _req(qkv, "qkv")
    if q_out is not None:
        _req(q_out, "q_out")
    assert rope_cs.dtype == torch.float32 and rope_cs.is_cuda and rope_cs.is_contiguous()
    rows = qkv.shape[0]
    ld_cache = 0
    if k_cache is not None:
        _req(k_cache, "k_cache"); _req(v_cache, "v_cache")
        assert k_cache.stride(0) == v_cache.stride(0)
        ld_cache = k_cache.stride(0)
    rc = _lib.lib().llb_rmsnorm_rope_append(
        qkv.data_ptr(), qkv.stride(0), _ptr(q_out), q_out.stride(0) if q_out is not None else 0,
        _ptr(k_cache), _ptr(v_cache), ld_cache, rows, n_heads, wq.data_ptr(), wk.data_ptr(), C.c_float(eps),
        rope_cs.data_ptr(), grid_hw[0], grid_hw[1], params_dev.data_ptr(),
        C.byref(shard) if shard is not None else None, _stream())
    _lib.check(rc, "llb_rmsnorm_rope_append")
    return q_out


def patchify(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [C_in, F, H, W] contiguous -> [F*(H/2)*(W/2), 4*C_in]."""
    _req(x, "x")
    assert x.is_contiguous()
    c, f, h, w = x.shape
    if out is None:
        out = torch.empty((f * (h // 2) * (w // 2), 4 * c), dtype=torch.bfloat16, device=x.device)
    rc = _lib.lib().llb_patchify(x.data_ptr(), out.data_ptr(), c, f, h, w, _stream())
    _lib.check(rc, "llb_patchify")
    return out


def unpatchify(y: torch.Tensor, c_out: int, frames: int, H: int, W: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(y, "y")
    assert y.is_contiguous() and y.shape == (frames * (H // 2) * (W // 2), 4 * c_out)
    if out is None:
        out = torch.empty((c_out, frames, H, W), dtype=torch.bfloat16, device=y.device)
    rc = _lib.lib().llb_unpatchify(y.data_ptr(), out.data_ptr(), c_out, frames, H, W, _stream())
    _lib.check(rc, "llb_unpatchify")
    return out


def sinusoidal(t: torch.Tensor, dim: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(t, "t", torch.float32)
    t = t.contiguous().view(-1)
    if out is None:
        out = torch.empty((t.numel(), dim), dtype=torch.bfloat16, device=t.device)
    rc = _lib.lib().llb_sinusoidal(t.data_ptr(), out.data_ptr(), t.numel(), dim, _stream())
    _lib.check(rc, "llb_sinusoidal")
    return out


def modulation_table(modulation: torch.Tensor, e0: torch.Tensor,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """modulation [n_layers, width], e0 [n_frames, width] -> [n_layers, n_frames, width]."""
    _req(modulation, "modulation"); _req(e0, "e0")
    assert modulation.is_contiguous() and e0.is_contiguous()
    nl, width = modulation.shape
    nf = e0.shape[0]
    assert e0.shape[1] == width
    if out is None:
        out = torch.empty((nl, nf, width), dtype=torch.bfloat16, device=e0.device)
    rc = _lib.lib().llb_modulation_table(modulation.data_ptr(), e0.data_ptr(), out.data_ptr(), nl, nf,
                                         width, _stream())
    _lib.check(rc, "llb_modulation_table")
    return out


def silu(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, "x")
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    rc = _lib.lib().llb_silu(x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    _lib.check(rc, "llb_silu")
    return out


def peer_barrier(flags_peers_dev: torch.Tensor, rank: int, n_ranks: int, epoch: torch.Tensor) -> None:
    """flags_peers_dev: int64 device tensor with every rank's flag-array address; epoch: uint32[1]."""
    rc = _lib.lib().llb_peer_barrier(flags_peers_dev.data_ptr(), rank, n_ranks, epoch.data_ptr(), _stream())
    _lib.check(rc, "llb_peer_barrier")


# ------------------------------------------------------------------------------------------------
def build_rope_table(head_dim: int = 128, max_pos: int = 1024, theta: float = 10000.0) -> torch.Tensor:
    """(cos, sin) table [max_pos, head_dim/2, 2] float32 for the 3-D RoPE of CausalWanModel.

    Same construction as the reference's ``freqs`` (wan/modules/causal_model.py:622-629 with
    rope_params at wan/modules/model.py:29-36): three frequency groups of head_dim/2 complex pairs,
    [frame | h | w] = [c - 2*(c//3), c//3, c//3], each group using theta^(-2i/dim_group) with
    dim_group = d - 4*(d//6), 2*(d//6), 2*(d//6).  Angles are formed in float64.
    """
    d = head_dim
    dims = [d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)]
    cols = []
    pos = torch.arange(max_pos, dtype=torch.float64)
    for dg in dims:
        inv = 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float64),
                              torch.arange(0, dg, 2, dtype=torch.float64) / dg)
        cols.append(torch.outer(pos, inv))
    ang = torch.cat(cols, dim=1)  # [max_pos, d/2]
    assert ang.shape[1] == d // 2
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------
# umT5 text encoder kernels (include/llb200.h, last section)
def embed_rows(table: torch.Tensor, ids: torch.Tensor, rows_per_seq: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b * rows_per_seq + r] = table[ids[b, r]] for r < rows_per_seq; ids int64 [B, >= rows_per_seq]."""
    _req(table, "table"); _req(ids, "ids", torch.int64)
    B = ids.shape[0]
    assert ids.shape[1] >= rows_per_seq
    Cc = table.shape[1]
    if out is None:
        out = torch.empty((B * rows_per_seq, Cc), dtype=torch.bfloat16, device=table.device)
    _req(out, "out")
    rc = _lib.lib().llb_embed_rows(table.data_ptr(), table.shape[0], ids.data_ptr(), ids.stride(0), out.data_ptr(),
                                   out.stride(0), B, rows_per_seq, Cc, _stream())
    _lib.check(rc, "llb_embed_rows")
    return out


def t5_attention(qkv: torch.Tensor, batch: int, n_heads: int, seq_lens: torch.Tensor, pos_emb: torch.Tensor,
                 bucket_lut: torch.Tensor, out: Optional[torch.Tensor] = None, max_seq_len: int = 0) -> torch.Tensor:
    """T5Attention core on the fused projection output qkv [batch * rows_per_seq, 3 * n_heads * 64];
    seq_lens int32 [batch] (device), pos_emb bf16 [num_buckets, n_heads], bucket_lut int32 [2 * center + 1]."""
    _req(qkv, "qkv"); _req(pos_emb, "pos_emb"); _req(seq_lens, "seq_lens", torch.int32)
    _req(bucket_lut, "bucket_lut", torch.int32)
    rows = qkv.shape[0]
    assert rows % batch == 0 and pos_emb.shape[1] == n_heads and pos_emb.is_contiguous()
    if out is None:
        out = torch.empty((rows, n_heads * 64), dtype=torch.bfloat16, device=qkv.device)
    _req(out, "out")
    rc = _lib.lib().llb_t5_attn(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), batch, rows // batch,
                                n_heads, max_seq_len, seq_lens.data_ptr(), pos_emb.data_ptr(), bucket_lut.data_ptr(),
                                (bucket_lut.numel() - 1) // 2, _stream())
    _lib.check(rc, "llb_t5_attn")
    return out


def t5_final_norm(x: torch.Tensor, w: torch.Tensor, batch: int, rows_out: int, seq_lens: torch.Tensor,
                  eps: float = 1e-6, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[batch, rows_out, C]: final T5LayerNorm of the valid rows, zeros elsewhere."""
    _req(x, "x"); _req(w, "w"); _req(seq_lens, "seq_lens", torch.int32)
    rows, Cc = x.shape
    assert rows % batch == 0
    if out is None:
        out = torch.empty((batch, rows_out, Cc), dtype=torch.bfloat16, device=x.device)
    _req(out, "out")
    assert out.is_contiguous()
    rc = _lib.lib().llb_t5_final_norm(x.data_ptr(), x.stride(0), out.data_ptr(), Cc, batch, rows // batch, rows_out,
                                      Cc, w.data_ptr(), C.c_float(eps), seq_lens.data_ptr(), _stream())
    _lib.check(rc, "llb_t5_final_norm")
    return out

"""ctypes binding of libllb200.so (the C ABI declared in include/llb200.h).

The library is built in-tree (longlive_b200/libllb200.so) by ``__graft_entry__.build()`` or
``make -C longlive_b200/csrc``.  There is no fallback: if the library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# LLB200_LIB: alternative build of the same library (tools/attn_debug_sweep.py times instrumented builds)
LIB_PATH = os.environ.get("LLB200_LIB") or os.path.join(_HERE, "libllb200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

LLB_MAX_SEGS = 4

# epilogues (include/llb200.h)
(EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_SILU, EPI_BIAS_GATE_RES, EPI_BIAS_RES, EPI_BIAS_F32, EPI_BIAS_MUL,
 EPI_BIAS_GELU_BF16, EPI_GEGLU_BF16) = range(9)


class KvState(C.Structure):
    _fields_ = [("global_end", C.c_int64), ("local_end", C.c_int64), ("rot", C.c_int64)]


class KvConfig(C.Structure):
    _fields_ = [
        ("cache_size", C.c_int64),
        ("sink_tokens", C.c_int64),
        ("max_attention_size", C.c_int64),
        ("local_attn_size", C.c_int32),
    ]


class KvPlan(C.Structure):
    _fields_ = [
        ("action", C.c_int32),
        ("is_recompute", C.c_int32),
        ("current_end", C.c_int64),
        ("num_evicted", C.c_int64),
        ("num_rolled", C.c_int64),
        ("local_start", C.c_int64),
        ("local_end", C.c_int64),
        ("write_start", C.c_int64),
        ("write_end", C.c_int64),
        ("roped_offset", C.c_int64),
        ("write_len", C.c_int64),
        ("attn_sink_len", C.c_int64),
        ("attn_window_start", C.c_int64),
        ("rot_after", C.c_int64),
        ("n_write_segs", C.c_int32),
        ("write_src", C.c_int64 * LLB_MAX_SEGS),
        ("write_dst", C.c_int64 * LLB_MAX_SEGS),
        ("write_n", C.c_int64 * LLB_MAX_SEGS),
        ("n_attn_segs", C.c_int32),
        ("attn_start", C.c_int64 * LLB_MAX_SEGS),
        ("attn_len", C.c_int64 * LLB_MAX_SEGS),
        ("attn_total", C.c_int64),
    ]


class StepParams(C.Structure):
    """Mirror of llb_step_params (device-resident per-forward parameters), 18 x int32."""

    _fields_ = [
        ("rope_start_frame", C.c_int32),
        ("n_write_segs", C.c_int32),
        ("write_src", C.c_int32 * LLB_MAX_SEGS),
        ("write_dst", C.c_int32 * LLB_MAX_SEGS),
        ("write_n", C.c_int32 * LLB_MAX_SEGS),
        ("n_attn_segs", C.c_int32),
        ("attn_start", C.c_int32 * LLB_MAX_SEGS),
        ("attn_len", C.c_int32 * LLB_MAX_SEGS),
        ("reserved", C.c_int32 * 1),
    ]


STEP_PARAMS_INT32 = C.sizeof(StepParams) // 4
LLB_MAX_RANKS = 8


class QkvShard(C.Structure):
    """llb_qkv_shard: head-parallel destination table of llb_rmsnorm_rope_append."""

    _fields_ = [("n_ranks", C.c_int32), ("heads_per_rank", C.c_int32), ("row0", C.c_int32),
                ("round_robin", C.c_int32), ("q_peers", C.c_void_p * LLB_MAX_RANKS),
                ("k_peers", C.c_void_p * LLB_MAX_RANKS), ("v_peers", C.c_void_p * LLB_MAX_RANKS)]


class OutShard(C.Structure):
    """llb_out_shard: head-parallel destination table of llb_attn_fwd."""

    _fields_ = [("n_ranks", C.c_int32), ("rows_per_rank", C.c_int32), ("head_col0", C.c_int32),
                ("head_col_stride", C.c_int32), ("ld_out", C.c_int64), ("out_peers", C.c_void_p * LLB_MAX_RANKS)]

class Conv3dDesc(C.Structure):
    """llb_conv3d_desc (include/llb200.h): one causal convolution over channels-last frame rings."""

    _fields_ = [("inp", C.c_void_p), ("in_frames", C.c_int), ("in_t0", C.c_int),
                ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("ld_in", C.c_int),
                ("Cout", C.c_int), ("ld_out", C.c_int),
                ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("kt", C.c_int), ("kh", C.c_int), ("kw", C.c_int),
                ("out", C.c_void_p), ("out_frames", C.c_int), ("out_t0", C.c_int), ("out_t_step", C.c_int),
                ("res", C.c_void_p), ("res_frames", C.c_int), ("res_t0", C.c_int),
                ("T", C.c_int),
                ("norm_out", C.c_void_p), ("norm_frames", C.c_int), ("norm_t0", C.c_int),
                ("norm_gamma", C.c_void_p), ("norm_channels", C.c_int), ("norm_silu", C.c_int)]


_PROTOS = {
    # name: (restype, argtypes)
    "llb_version": (C.c_int, []),
    "llb_last_error": (C.c_char_p, []),
    "llb_launch_count": (C.c_int64, []),
    "llb_kv_ring_plan": (
        C.c_int,
        [C.POINTER(KvConfig), C.POINTER(KvState), C.c_int64, C.c_int64, C.c_int32, C.POINTER(KvPlan)],
    ),
    "llb_kv_ring_commit": (C.c_int, [C.POINTER(KvPlan), C.POINTER(KvState)]),
    "llb_kv_ring_phys": (C.c_int64, [C.POINTER(KvConfig), C.c_int64, C.c_int64]),
    "llb_gemm_bf16": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
         C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p],
    ),
    "llb_gemm_bf16_splitk": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
         C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
         C.c_float, C.c_void_p],
    ),
    "llb_gemm_fp8": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
         C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64,
         C.c_void_p],
    ),
    "llb_ln_modulate_fp8": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
         C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p],
    ),
    "llb_quant_rows_fp8": (
        C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "llb_attn_fwd": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
         C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_int64,
         C.POINTER(OutShard), C.c_void_p],
    ),
    "llb_attn_workspace_bytes": (C.c_int64, []),
    "llb_ln_modulate": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
         C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p],
    ),
    "llb_rmsnorm_rope_append": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
         C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
         C.POINTER(QkvShard), C.c_void_p],
    ),
    "llb_peer_barrier": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "llb_rmsnorm": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p],
    ),
    "llb_patchify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "llb_unpatchify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "llb_sinusoidal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "llb_modulation_table": (
        C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "llb_silu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "llb_conv3d": (C.c_int, [C.POINTER(Conv3dDesc), C.c_void_p]),
    "llb_vae_norm": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int,
         C.c_void_p, C.c_int, C.c_void_p]),
    "llb_vae_upsample2x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "llb_transpose_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "llb_softmax_rows": (
        C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "llb_vae_latent_in": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
         C.c_int64, C.c_int, C.c_void_p]),
    "llb_vae_pixel_out": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "llb_embed_rows": (
        C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                  C.c_void_p]),
    "llb_t5_attn": (
        C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "llb_t5_final_norm": (
        C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                  C.c_float, C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libllb200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j4"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libllb200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(longlive_b200 has no CPU or PyTorch fallback)")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().llb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")

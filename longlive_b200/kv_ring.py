"""Host-side KV ring-buffer bookkeeping (integer only) — thin Python face of the C planner
``llb_kv_ring_plan`` / ``llb_kv_ring_commit`` in libllb200.so.

Replaces the `.item()`-driven index math, `clone()`, roll `memmove` and `torch.cat` of
CausalWanSelfAttention.forward (wan/modules/causal_model.py:206-360) and the second roll+insert of
CausalWanModel._apply_cache_updates (:849-905).  The reference keeps the local window in
chronological order by moving data; here the rolling region [sink, size) is a ring and eviction is
an index rotation, so no cache byte is ever copied.  `global_end_index` / `local_end_index` keep
the reference's values exactly; `logical_view` returns K/V in the reference's logical order.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Tuple

import torch

from . import _lib
from ._lib import KvConfig, KvPlan, KvState, StepParams


@dataclass
class RingPlan:
    """One forward's cache plan.  The first block mirrors the reference's cache_update_info."""
    action: str
    is_recompute: bool
    current_end: int
    num_evicted: int
    num_rolled: int
    local_start: int
    local_end: int
    write_start: int
    write_end: int
    roped_offset: int
    write_len: int
    attn_sink_len: int
    attn_window_start: int
    rot_after: int
    writes: List[Tuple[int, int, int]]     # (src row in new tokens, physical dst row, n)
    attn_segs: List[Tuple[int, int]]       # physical (start row, len)
    raw: KvPlan

    @property
    def attn_total(self) -> int:
        return sum(n for _, n in self.attn_segs)


class KvRing:
    """Ring state shared by all layers of one kv_cache list (the reference keeps identical
    global/local end indices in every layer, SURVEY.md 8a note vii)."""

    def __init__(self, cache_size: int, sink_tokens: int, max_attention_size: int,
                 local_attn_size: int, global_end: int = 0, local_end: int = 0):
        self.cfg = KvConfig(cache_size, sink_tokens, max_attention_size, local_attn_size)
        self.state = KvState(global_end, local_end, 0)

    # -- reference-visible scalars
    @property
    def global_end(self) -> int:
        return int(self.state.global_end)

    @property
    def local_end(self) -> int:
        return int(self.state.local_end)

    @property
    def rot(self) -> int:
        return int(self.state.rot)

    def plan(self, current_start: int, num_new: int, sink_recache_after_switch: bool = False) -> RingPlan:
        p = KvPlan()
        rc = _lib.lib().llb_kv_ring_plan(C.byref(self.cfg), C.byref(self.state), int(current_start),
                                         int(num_new), int(bool(sink_recache_after_switch)), C.byref(p))
        _lib.check(rc, "llb_kv_ring_plan")
        return RingPlan(
            action="roll_and_insert" if p.action == 1 else "direct_insert",
            is_recompute=bool(p.is_recompute), current_end=p.current_end, num_evicted=p.num_evicted,
            num_rolled=p.num_rolled, local_start=p.local_start, local_end=p.local_end,
            write_start=p.write_start, write_end=p.write_end, roped_offset=p.roped_offset,
            write_len=p.write_len, attn_sink_len=p.attn_sink_len, attn_window_start=p.attn_window_start,
            rot_after=p.rot_after,
            writes=[(p.write_src[i], p.write_dst[i], p.write_n[i]) for i in range(p.n_write_segs)],
            attn_segs=[(p.attn_start[i], p.attn_len[i]) for i in range(p.n_attn_segs)],
            raw=p)

    def commit(self, plan: RingPlan) -> None:
        rc = _lib.lib().llb_kv_ring_commit(C.byref(plan.raw), C.byref(self.state))
        _lib.check(rc, "llb_kv_ring_commit")

    def phys(self, logical: int, rot: int | None = None) -> int:
        return int(_lib.lib().llb_kv_ring_phys(C.byref(self.cfg), self.rot if rot is None else rot,
                                               int(logical)))

    def logical_index(self, device=None) -> torch.Tensor:
        """int64 [cache_size]: physical row of every logical position (for logical_view)."""
        size, S = int(self.cfg.cache_size), int(self.cfg.sink_tokens)
        idx = torch.arange(size, dtype=torch.long)
        ring = size - S
        if ring > 0 and self.rot:
            idx[S:] = S + (torch.arange(ring, dtype=torch.long) + self.rot) % ring
        return idx.to(device) if device is not None else idx

    def step_params(self, plan: RingPlan, rope_start_frame: int) -> StepParams:
        sp = StepParams()
        sp.rope_start_frame = rope_start_frame
        sp.n_write_segs = len(plan.writes)
        for i, (s, d, n) in enumerate(plan.writes):
            sp.write_src[i], sp.write_dst[i], sp.write_n[i] = s, d, n
        sp.n_attn_segs = len(plan.attn_segs)
        for i, (s, n) in enumerate(plan.attn_segs):
            sp.attn_start[i], sp.attn_len[i] = s, n
        return sp


def logical_view(kv_cache_layer: dict, ring: KvRing):
    """K, V of one layer re-ordered into the reference's logical layout [B, size, H, D]."""
    idx = ring.logical_index(kv_cache_layer["k"].device)
    return kv_cache_layer["k"][:, idx], kv_cache_layer["v"][:, idx]

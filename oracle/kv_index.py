"""ORACLE — test infrastructure only.  Pure-Python restatement of the reference's KV-cache index
arithmetic (all quantities in tokens), independent of the C planner it checks.

Follows wan/modules/causal_model.py:213-246 (roll branch), :291-306 (direct branch), :331-360
(attended range) and :849-905 (commit).  Pinned by tests/golden/index_traces.json, which was
recorded from the reference model itself (oracle/make_golden.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class RefCacheSim:
    """Simulates ONE layer's cache content symbolically: every slot holds a label (or None = zero)."""
    size: int
    sink_tokens: int
    max_attention_size: int
    local_attn_size: int
    G: int = 0
    Le: int = 0
    slots: List[Optional[tuple]] = field(default_factory=list)

    def __post_init__(self):
        if not self.slots:
            self.slots = [None] * self.size

    def zero(self):
        self.slots = [None] * self.size

    def step(self, current_start: int, n: int, labels: List[tuple], sink_recache: bool = False) -> dict:
        S, size, M = self.sink_tokens, self.size, self.max_attention_size
        G, Le = self.G, self.Le
        current_end = current_start + n                                               # :213
        is_recompute = current_end <= G and current_start > 0                         # :230
        temp = list(self.slots)                                                       # clone :251/:298
        rec = {"current_start": current_start, "current_end": current_end, "is_recompute": is_recompute,
               "num_evicted": None, "num_rolled": None}
        if self.local_attn_size != -1 and current_end > G and n + Le > size:          # :231-232
            evicted = n + Le - size                                                   # :235
            rolled = Le - evicted - S                                                 # :236
            Le2 = Le + current_end - G - evicted                                      # :244-245
            Ls2 = Le2 - n                                                             # :246
            temp[S:S + rolled] = temp[S + evicted:S + evicted + rolled]               # :257-260
            ws = max(Ls2, S) if is_recompute else Ls2                                 # :264
            rec.update(action="roll_and_insert", num_evicted=evicted, num_rolled=rolled)
        else:
            Le2 = Le + current_end - G                                                # :293
            Ls2 = Le2 - n                                                             # :294
            ws = max(Ls2, S) if is_recompute else Ls2                                 # :302
            if sink_recache:
                ws = Ls2                                                              # :303-304
            rec.update(action="direct_insert")
        off = max(0, ws - Ls2)                                                        # :265 / :305
        wl = max(0, Le2 - ws)                                                         # :266 / :306
        if wl > 0:
            temp[ws:Le2] = labels[off:off + wl]                                       # :268 / :310
        if S > 0:                                                                     # :331-353
            budget = M - S
            w0 = max(S, Le2 - budget) if budget > 0 else Le2
            attended = temp[:S] + (temp[w0:Le2] if budget > 0 else [])
        else:                                                                         # :354-360
            w0 = max(0, Le2 - M)
            attended = temp[w0:Le2]
        rec.update(local_start_index=Ls2, local_end_index=Le2, write_start_index=ws, write_end_index=Le2,
                   roped_offset=off, write_len=wl, attn_window_start=w0, attended=attended,
                   new_tokens=wl)
        # commit (:861-904): same roll + insert on the real cache, indices unless recompute
        self.slots = temp
        if not is_recompute:
            self.G, self.Le = current_end, Le2
        rec.update(global_end_after=self.G, local_end_after=self.Le)
        return rec

"""KV ring planner (C, libllb200.so) vs the reference's cache bookkeeping — bit-exact integers.

Three independent parties must agree on every call of every recorded scenario:
  golden   tests/golden/index_traces.json, recorded from the REAL reference model
  oracle   oracle/kv_index.py, a pure-Python restatement (also simulates slot contents)
  product  llb_kv_ring_plan / llb_kv_ring_commit through the C ABI
and the ring layout (no data movement) must expose exactly the reference's logical cache content
and attended key set.
"""
import json
import os

import pytest

from longlive_b200.kv_ring import KvRing
from oracle.kv_index import RefCacheSim

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "index_traces.json")
with open(GOLDEN) as f:
    TRACES = json.load(f)


def _mk(cfg):
    fs = cfg["frame_seqlen"]
    size = cfg["cache_size"]
    S = cfg["sink"] * fs
    M = 32760 if cfg["local"] == -1 else cfg["local"] * fs
    return fs, size, S, M


@pytest.mark.parametrize("name", sorted(TRACES))
def test_planner_matches_reference_trace(name):
    sc = TRACES[name]
    cfg = sc["config"]
    fs, size, S, M = _mk(cfg)
    ring = KvRing(size, S, M, cfg["local"])
    sim = RefCacheSim(size, S, M, cfg["local"])
    phys = [None] * size  # what the ring buffer holds, by PHYSICAL row
    for ci, g in enumerate(sc["calls"]):
        cur, n = g["current_start"], g["current_end"] - g["current_start"]
        recache = g["kind"] == "recache"
        sink_recache = recache and not cfg["global_sink"]
        if recache and not cfg["global_sink"]:
            sim.zero()
            phys = [None] * size
        labels = [(ci, i) for i in range(n)]
        o = sim.step(cur, n, labels, sink_recache)
        p = ring.plan(cur, n, sink_recache)
        # ---- golden (real reference) vs oracle restatement
        for k in ("action", "is_recompute", "current_end", "local_start_index", "local_end_index",
                  "write_start_index", "write_end_index", "new_tokens", "global_end_after", "local_end_after"):
            assert o[k] == g[k], f"{name} call {ci} {k}: oracle {o[k]} != golden {g[k]}"
        if g["action"] == "roll_and_insert":
            assert (o["num_evicted"], o["num_rolled"]) == (g["num_evicted"], g["num_rolled"])
        # ---- product planner vs golden
        assert p.action == g["action"], f"{name} call {ci}"
        assert p.is_recompute == g["is_recompute"]
        assert p.current_end == g["current_end"]
        assert (p.local_start, p.local_end) == (g["local_start_index"], g["local_end_index"])
        assert (p.write_start, p.write_end) == (g["write_start_index"], g["write_end_index"])
        assert p.write_len == g["new_tokens"]
        if g["action"] == "roll_and_insert":
            assert (p.num_evicted, p.num_rolled) == (g["num_evicted"], g["num_rolled"])
        assert p.attn_window_start == o["attn_window_start"]
        # ---- execute the physical plan symbolically
        assert sum(nw for _, _, nw in p.writes) == p.write_len
        for (src, dst, nw) in p.writes:
            assert 0 <= dst and dst + nw <= size
            phys[dst:dst + nw] = labels[src:src + nw]
        attended = []
        for (s, ln) in p.attn_segs:
            assert 0 <= s and s + ln <= size and ln > 0
            attended += phys[s:s + ln]
        key = lambda x: (-1, -1) if x is None else x
        assert sorted(attended, key=key) == sorted(o["attended"], key=key), f"{name} call {ci}: attended set"
        ring.commit(p)
        assert (ring.global_end, ring.local_end) == (g["global_end_after"], g["local_end_after"])
        # ---- logical view of the ring == reference cache content, wherever the reference holds data
        idx = ring.logical_index().tolist()
        le = ring.local_end
        for pos in range(le):
            assert phys[idx[pos]] == sim.slots[pos], f"{name} call {ci}: logical pos {pos}"


def test_steady_state_is_one_segment_and_zero_copy():
    """North-star config: after the cache fills, attention reads ONE contiguous range [0, size) and
    every chunk writes exactly n rows (the reference moves ~0.5 GB per layer-forward here)."""
    fs = 1560
    ring = KvRing(12 * fs, 3 * fs, 12 * fs, 12)
    for chunk in range(40):
        for call in range(5):
            p = ring.plan(chunk * 3 * fs, 3 * fs)
            if chunk >= 4:
                assert p.attn_segs == [(0, 12 * fs)]
                assert (p.action == "roll_and_insert") == (call == 0)
            assert p.write_len == 3 * fs and len(p.writes) == 1
            ring.commit(p)
    assert ring.global_end == 120 * fs and ring.local_end == 12 * fs


def test_planner_rejects_bad_arguments():
    ring = KvRing(100, 10, 100, 10)
    with pytest.raises(RuntimeError):
        ring.plan(0, 0)
    with pytest.raises(RuntimeError):
        ring.plan(-5, 10)

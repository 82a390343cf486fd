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


def _pipeline_calls(rng, T, chunk, local, switches):
    """The call pattern of the two pipelines (causal_inference.py:139-200, interactive...:283-330): per chunk four
    denoising calls + one clean-context call at the same position; at a prompt switch one recache call over the
    last min(local, start) frames first."""
    calls, start = [], 0
    while start < T:
        if start in switches and start > 0:
            n = start if local == -1 else min(local, start)
            calls.append(("recache", start - n, n))
        for _ in range(5):
            calls.append(("denoise", start, chunk))
        start += chunk
    return calls


@pytest.mark.parametrize("seed", range(40))
def test_planner_matches_oracle_on_random_pipelines(seed):
    """Beyond the 12 recorded reference traces: random geometry (tokens per frame, sink, window, chunk, video
    length, switch points, both global_sink settings) driven with the pipelines' call pattern; the C planner must
    agree with the pure-Python restatement of the reference on every integer, on the attended key SET and on the
    logical cache content."""
    import random
    rng = random.Random(seed)
    fs = rng.choice([1, 3, 4, 7])
    chunk = rng.choice([1, 2, 3])
    sink = rng.choice([0, 1, 2, 3])
    T = chunk * rng.randint(4, 24)
    if rng.random() < 0.2:
        local, size_frames = -1, T  # global attention: the cache holds the whole video, never rolls
    else:
        local = max(sink + chunk, rng.randint(sink + chunk, sink + 4 * chunk + 3))
        size_frames = local
    global_sink = rng.random() < 0.5
    size, S = size_frames * fs, sink * fs
    M = 32760 if local == -1 else local * fs
    switches = set(rng.sample(range(chunk, T, chunk), k=min(3, len(range(chunk, T, chunk))))) if rng.random() < 0.7 else set()
    ring = KvRing(size, S, M, local)
    sim = RefCacheSim(size, S, M, local)
    phys = [None] * size
    for ci, (kind, f0, nf) in enumerate(_pipeline_calls(rng, T, chunk, local, switches)):
        cur, n = f0 * fs, nf * fs
        recache = kind == "recache"
        sink_recache = recache and not global_sink
        if sink_recache:
            sim.zero()
            phys = [None] * size
        labels = [(ci, i) for i in range(n)]
        o = sim.step(cur, n, labels, sink_recache)
        p = ring.plan(cur, n, sink_recache)
        ctx = f"seed {seed} call {ci} {kind} fs={fs} chunk={chunk} sink={sink} local={local} gs={global_sink}"
        assert p.action == o["action"] and p.is_recompute == o["is_recompute"], ctx
        assert p.current_end == o["current_end"], ctx
        assert (p.local_start, p.local_end) == (o["local_start_index"], o["local_end_index"]), ctx
        assert (p.write_start, p.write_end, p.write_len) == (o["write_start_index"], o["write_end_index"], o["write_len"]), ctx
        assert p.roped_offset == o["roped_offset"], ctx
        if o["action"] == "roll_and_insert":
            assert (p.num_evicted, p.num_rolled) == (o["num_evicted"], o["num_rolled"]), ctx
        assert sum(nw for _, _, nw in p.writes) == p.write_len, ctx
        for (src, dst, nw) in p.writes:
            assert 0 <= dst and dst + nw <= size, ctx
            phys[dst:dst + nw] = labels[src:src + nw]
        attended = []
        for (s, ln) in p.attn_segs:
            assert 0 <= s and s + ln <= size and ln > 0, ctx
            attended += phys[s:s + ln]
        key = lambda x: (-1, -1) if x is None else x
        assert sorted(attended, key=key) == sorted(o["attended"], key=key), ctx
        ring.commit(p)
        assert (ring.global_end, ring.local_end) == (o["global_end_after"], o["local_end_after"]), ctx
        idx = ring.logical_index().tolist()
        for pos in range(ring.local_end):
            assert phys[idx[pos]] == sim.slots[pos], ctx


@pytest.mark.skipif(not os.path.isdir("/root/reference/wan/modules"), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_live_reference_on_random_geometry(seed):
    """Pins oracle/kv_index.py beyond the recorded traces: a tiny REAL reference model (its own cache logic, spied at
    _apply_cache_updates) is driven with the pipelines' call pattern on random geometry; every integer must match."""
    import random
    import torch
    from oracle import make_golden as mg
    from oracle import ref_shims
    from oracle import wan_oracle as wo
    rng = random.Random(100 + seed)
    H, W = rng.choice([(2, 2), (2, 6), (4, 4), (4, 6)])
    fs = (H // 2) * (W // 2)
    chunk = rng.choice([1, 2, 3])
    sink = rng.choice([0, 1, 2, 3])
    T = chunk * rng.randint(4, 10)
    local = -1 if seed == 5 else rng.randint(sink + chunk, sink + 3 * chunk + 2)
    global_sink = rng.random() < 0.5
    switches = sorted(rng.sample(range(chunk, T, chunk), k=min(2, len(range(chunk, T, chunk)))))
    cfg = wo.WanConfig(dim=16, ffn_dim=16, num_heads=2, num_layers=1, text_dim=8, text_len=4,
                       local_attn_size=local, sink_size=sink, frame_seqlen=fs)
    model = ref_shims.build_reference_model(cfg, wo.init_state_dict(cfg, seed=0), "sdpa")
    log = []
    mg.spy_model(model, log)
    size = (local if local != -1 else T) * fs
    kv = wo.new_kv_cache(cfg, 1, size, "cpu")
    cc = wo.new_crossattn_cache(cfg, 1, "cpu")
    ctx = wo.synth_prompt_embeds(cfg, 1, 3)
    S, M = sink * fs, (32760 if local == -1 else local * fs)
    sim = RefCacheSim(size, S, M, local)
    for ci, (start, n, kind, seg) in enumerate(mg.call_pattern(T, chunk, switches, local)):
        recache = kind == "recache"
        if recache:
            if not global_sink:
                for c in kv:
                    c["k"].zero_(); c["v"].zero_()
                sim.zero()
            for c in cc:
                c["is_init"] = False
        x = torch.zeros(1, 16, n, H, W, dtype=torch.bfloat16)
        with torch.no_grad():
            model(x, t=torch.zeros(1, n), context=ctx, seq_len=1 << 20, kv_cache=kv, crossattn_cache=cc,
                  current_start=start * fs, sink_recache_after_switch=(recache and not global_sink))
        g = log[-1]
        o = sim.step(start * fs, n * fs, [(ci, i) for i in range(n * fs)], recache and not global_sink)
        ctxs = f"seed {seed} call {ci} {kind}: fs={fs} chunk={chunk} sink={sink} local={local} gs={global_sink}"
        for k in ("action", "is_recompute", "current_end", "local_start_index", "local_end_index",
                  "write_start_index", "write_end_index", "global_end_after", "local_end_after"):
            assert o[k] == g[k], f"{ctxs}: {k} oracle {o[k]} != reference {g[k]}"
        assert o["new_tokens"] == g["new_tokens"], ctxs
        if g["action"] == "roll_and_insert":
            assert (o["num_evicted"], o["num_rolled"]) == (g["num_evicted"], g["num_rolled"]), ctxs

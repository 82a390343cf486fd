"""Longer-horizon parity on the B200: rolling eviction over many chunks (no drift growth) and the
interactive prompt-switch path with KV-recache (both global_sink settings), CUDA pipelines vs the
oracle pipeline on the same GPU with identical weights, latents, embeddings and re-noise draws.
The full 240-frame versions (BASELINE configs[2], [3]) are run by tools/drift_240.py; their
results are committed under profiles/."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _setup(global_sink=False):
    from oracle import wan_oracle as wo
    from longlive_b200.model import CausalWanModel
    from longlive_b200.wrapper import WanDiffusionWrapper
    cfg = wo.WanConfig()
    sd = wo.init_state_dict(cfg, seed=0)
    model = CausalWanModel(local_attn_size=12, sink_size=3)
    model.load_state_dict(sd)
    model = model.to(DEV).to(torch.bfloat16)
    gen = WanDiffusionWrapper(model=model, timestep_shift=5.0)
    ogen = wo.OracleGenerator(wo.OracleModel(cfg, sd).to(DEV), shift=5.0)

    class MK(dict):
        __getattr__ = dict.get
    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=3, context_noise=0, global_sink=global_sink,
                                 model_kwargs=MK(local_attn_size=12, sink_size=3, timestep_shift=5.0))
    return cfg, gen, ogen, args


def test_rolling_window_no_drift_36_frames():
    from oracle import wan_oracle as wo
    from oracle.make_golden import SeededNoise
    from oracle.pipeline_oracle import run_pipeline
    from longlive_b200.pipeline import CausalInferencePipeline
    cfg, gen, ogen, args = _setup()
    T = 36  # 12 chunks: 4 fill the cache, 8 roll
    noise = wo.synth_prompt_embeds  # placeholder to keep flake quiet
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(1, T, 16, 60, 104, generator=g).to(torch.bfloat16).to(DEV)
    prompt = wo.synth_prompt_embeds(cfg, 100, 150).to(DEV)
    sn = SeededNoise()
    olat, okv = run_pipeline(ogen, cfg, noise, [prompt], renoise=lambda like, b, s: sn(like))
    pipe = CausalInferencePipeline(args, torch.device(DEV), generator=gen,
                                   text_encoder=lambda text_prompts: {"prompt_embeds": prompt})
    sn2 = SeededNoise()
    pipe.renoise_fn = lambda like, b, s: sn2(like)
    _, lat = pipe.inference(noise, ["p"], return_latents=True)
    errs = [rel_l2(lat[:, c:c + 3], olat[:, c:c + 3]) for c in range(0, T, 3)]
    print("rel-L2 per chunk:", [f"{e:.2e}" for e in errs])
    assert max(errs) <= 1e-2, errs
    # no drift growth: the rolling chunks stay at the level of the first ones
    assert sum(errs[-4:]) / 4 <= 1.25 * (sum(errs[:4]) / 4) + 5e-4, errs
    ring = pipe.kv_cache1[0]["_llb_ring"]
    assert (ring.global_end, ring.local_end) == (T * 1560, 12 * 1560)
    assert ring.rot == (8 * 3 * 1560) % (9 * 1560)


@pytest.mark.parametrize("global_sink", [False, True], ids=["local_sink", "global_sink"])
def test_interactive_switches_vs_oracle(global_sink):
    from oracle import wan_oracle as wo
    from oracle.make_golden import SeededNoise
    from oracle.pipeline_oracle import run_pipeline
    from longlive_b200.kv_ring import logical_view
    from longlive_b200.pipeline import InteractiveCausalInferencePipeline
    cfg, gen, ogen, args = _setup(global_sink)
    T, switches = 27, [4, 19]  # fire at frames 6 (recache 6 frames) and 21 (recache 12 frames, ring rotated)
    g = torch.Generator().manual_seed(4)
    noise = torch.randn(1, T, 16, 60, 104, generator=g).to(torch.bfloat16).to(DEV)
    prompts = [wo.synth_prompt_embeds(cfg, 100 + i, 100 + 40 * i).to(DEV) for i in range(3)]
    sn = SeededNoise()
    olat, okv = run_pipeline(ogen, cfg, noise, prompts, switches, global_sink=global_sink,
                             renoise=lambda like, b, s: sn(like))
    pipe = InteractiveCausalInferencePipeline(
        args, torch.device(DEV), generator=gen,
        text_encoder=lambda text_prompts: {"prompt_embeds": prompts[int(text_prompts[0])]})
    sn2 = SeededNoise()
    pipe.renoise_fn = lambda like, b, s: sn2(like)
    _, lat = pipe.inference(noise, text_prompts_list=[["0"], ["1"], ["2"]], switch_frame_indices=switches,
                            return_latents=True)
    assert [s["frame"] for s in pipe.switch_log] == [6, 21]
    assert [s["recached_frames"] for s in pipe.switch_log] == [6, 12]
    errs = [rel_l2(lat[:, c:c + 3], olat[:, c:c + 3]) for c in range(0, T, 3)]
    print(f"global_sink={global_sink} rel-L2 per chunk:", [f"{e:.2e}" for e in errs])
    assert max(errs) <= 1e-2, errs
    ring = pipe.kv_cache1[0]["_llb_ring"]
    assert ring.global_end == int(okv[0]["global_end_index"].item())
    assert ring.local_end == int(okv[0]["local_end_index"].item())
    for l in (0, 29):
        k, v = logical_view(pipe.kv_cache1[l], ring)
        assert rel_l2(k, okv[l]["k"]) < 3e-2 and rel_l2(v, okv[l]["v"]) < 3e-2


# ------------------------------------------------------------------------------------------------
# BASELINE configs[2] / configs[3] at FULL length (240 latent frames = 80 chunks = 400 denoising
# forwards + 5 recaches), in the driver-run suite.  The oracle runs bf16 fused SDPA here - the
# function the reference itself evaluates through flash-attn (wan/modules/attention.py:131-145) - so
# that one mode costs ~40 s instead of ~200 s; the shorter tests above keep the exact-fp32 checker.
# ------------------------------------------------------------------------------------------------
_SHARED = {}


def _shared_model():
    """One 30-block model (and its captured graphs) for the three 240-frame runs."""
    if "model" not in _SHARED:
        from oracle import wan_oracle as wo
        from longlive_b200.model import CausalWanModel
        sd = wo.init_state_dict(wo.WanConfig(), seed=0)
        model = CausalWanModel(local_attn_size=12, sink_size=3)
        model.load_state_dict(sd)
        _SHARED["model"], _SHARED["sd"] = model.to(DEV).to(torch.bfloat16), sd
    return _SHARED["model"], _SHARED["sd"]


def _check_240(res, switched):
    errs = res["rel_l2_per_chunk"]
    print({k: v for k, v in res.items() if k not in ("rel_l2_per_chunk", "profile")})
    assert len(errs) == 80
    assert max(errs) <= 1e-2, (max(errs), errs)                      # north star: rel-L2 <= 1e-2 per chunk
    # no drift GROWTH over 240 frames: last ten chunks not above the level of the first ten, no upward trend
    # (with prompt switches the level steps with the prompt - 7.7e-3 down to 6.2e-3 for the longer prompts -
    # so the fitted slope may be negative; measured: +6e-8 per chunk single prompt, -2.2e-5 with switches)
    assert res["mean_last_10"] <= 1.15 * res["mean_first_10"] + 3e-4, res
    assert res["slope_per_chunk"] <= 5e-6, res["slope_per_chunk"]
    # integer contract at the end of the run
    assert res["global_end"] == res["oracle_global_end"] == 240 * 1560 == 374400
    assert res["local_end"] == res["oracle_local_end"] == 12 * 1560 == 18720
    if switched:
        # switch_frame_indices 40..200 fire at the first chunk start >= index (interactive...:237-264)
        assert res["switch_frames"] == [42, 81, 120, 162, 201]
        assert res["recached_frames"] == [12] * 5
    else:
        assert res["switch_frames"] == []


def test_config2_240_frames_rolling_no_drift():
    """configs[2]: 60 s video, 240 latent frames, rolling eviction from chunk 4 on, sink retained."""
    from tools import drift_240
    model, sd = _shared_model()
    res = drift_240.run("single", 240, "sdpa", DEV, model=model, state_dict=sd, timed_second_pass=True)
    drift_240.save(res)
    _check_240(res, switched=False)


@pytest.mark.parametrize("mode", ["switch", "switch_global_sink"])
def test_config3_240_frames_six_prompts(mode):
    """configs[3]: 240 frames, 6 prompts / 5 switches with KV-recache, global_sink False and True."""
    from tools import drift_240
    model, sd = _shared_model()
    res = drift_240.run(mode, 240, "sdpa", DEV, model=model, state_dict=sd)
    drift_240.save(res)
    _check_240(res, switched=True)

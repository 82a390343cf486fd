"""The UNMODIFIED reference on the same B200, beside the CUDA path.

`__graft_entry__.build()` mirrors the reference tree to baseline/_ref/ (git-ignored, shipped to the
GPU box); here it runs with its own GPU kernels - flash-attn 2 (wan/modules/attention.py:131-145),
cuBLAS, ATen - through its own classes:

  * INTEGRATION.md option A: the reference's InteractiveCausalInferencePipeline drives OUR generator
    (constructor injection) and must produce exactly what our own pipeline produces;
  * latents of our pipeline vs latents of the reference's pipeline + WanDiffusionWrapper +
    CausalWanModel on identical weights / noise / embeddings / re-noise draws: a 1-layer 1536-wide
    model with two prompt switches, and the full 30-block shape over 5 chunks (north-star gate:
    rel-L2 <= 1e-2 per chunk, indices equal).
"""
import types

import pytest
import torch

from oracle import ref_shims

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shims.shipped_available(),
                                 reason="baseline/_ref missing: run __graft_entry__.build() in the build container")]
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _args(cfg, global_sink=False):
    class MK(dict):
        __getattr__ = dict.get
    return types.SimpleNamespace(
        denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True, num_frame_per_block=3,
        context_noise=0, global_sink=global_sink,
        model_kwargs=MK(local_attn_size=cfg.local_attn_size, sink_size=cfg.sink_size, timestep_shift=5.0))


def _our_generator(cfg, sd):
    from longlive_b200.model import CausalWanModel
    from longlive_b200.wrapper import WanDiffusionWrapper
    m = CausalWanModel(dim=cfg.dim, ffn_dim=cfg.ffn_dim, num_heads=cfg.num_heads, num_layers=cfg.num_layers,
                       text_dim=cfg.text_dim, text_len=cfg.text_len, local_attn_size=cfg.local_attn_size,
                       sink_size=cfg.sink_size, frame_seqlen=cfg.frame_seqlen)
    m.load_state_dict(sd)
    return WanDiffusionWrapper(model=m.to(DEV).to(torch.bfloat16), timestep_shift=5.0)


class _SeededRandnLike:
    """Patches torch.randn_like (the re-noise draw of both pipelines) with on-device seeded draws."""

    def __enter__(self):
        from oracle.pipeline_oracle import DeviceSeededNoise
        self.real = torch.randn_like
        torch.randn_like = DeviceSeededNoise(DEV)
        return self

    def __exit__(self, *a):
        torch.randn_like = self.real


_VAE = types.SimpleNamespace(decode_to_pixel=lambda latent, use_cache=False: latent.float())


def _run_interactive(PipeCls, cfg, gen, prompts, noise, switches, global_sink=False):
    pipe = PipeCls(_args(cfg, global_sink), torch.device(DEV), generator=gen,
                   text_encoder=lambda text_prompts: {"prompt_embeds": prompts[int(text_prompts[0])]}, vae=_VAE)
    pipe.num_transformer_blocks = cfg.num_layers
    with _SeededRandnLike(), torch.no_grad():  # inference.py wraps the call in no_grad as well
        _, lat = pipe.inference(noise, text_prompts_list=[[str(i)] for i in range(len(prompts))],
                                switch_frame_indices=switches, return_latents=True)
    return lat, pipe


@pytest.mark.parametrize("global_sink", [False, True], ids=["local_sink", "global_sink"])
def test_reference_pipeline_drives_our_generator(global_sink):
    """INTEGRATION.md option A, executed: reference pipeline class + our WanDiffusionWrapper."""
    from oracle import wan_oracle as wo
    from oracle.make_golden import PIPE_CFG
    from longlive_b200.pipeline import InteractiveCausalInferencePipeline as Ours
    ref_shims.use_shipped_copy()
    ref_shims.install("flash")
    _, RefInteractive = ref_shims.reference_pipelines()
    cfg = wo.WanConfig(**PIPE_CFG)
    sd = wo.init_state_dict(cfg, seed=0)
    gen = _our_generator(cfg, sd)
    T, switches = 21, [4, 13]
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(1, T, 16, 60, 104, generator=g).to(torch.bfloat16).to(DEV)
    prompts = [wo.synth_prompt_embeds(cfg, 300 + i, 80 + 50 * i).to(DEV) for i in range(3)]
    lat_ref_pipe, rp = _run_interactive(RefInteractive, cfg, gen, prompts, noise, switches, global_sink)
    g_end = int(rp.kv_cache1[0]["global_end_index"].item())
    l_end = int(rp.kv_cache1[-1]["local_end_index"].item())
    lat_our_pipe, op = _run_interactive(Ours, cfg, gen, prompts, noise, switches, global_sink)
    assert [s["frame"] for s in op.switch_log] == [6, 15]
    # same generator, same call sequence: identical results, and the reference-allocated cache dicts
    # carry the reference-visible indices
    assert torch.equal(lat_ref_pipe, lat_our_pipe), rel_l2(lat_ref_pipe, lat_our_pipe)
    assert (g_end, l_end) == (T * 1560, 12 * 1560)
    assert int(op.kv_cache1[0]["global_end_index"].item()) == g_end


def test_one_layer_pipeline_vs_reference_flash_attn():
    """Our pipeline vs the reference's pipeline / wrapper / model with flash-attn, 1 x 1536-wide block,
    21 frames, two prompt switches (the second one on a rotated ring)."""
    from oracle import wan_oracle as wo
    from oracle.make_golden import PIPE_CFG
    from longlive_b200.pipeline import InteractiveCausalInferencePipeline as Ours
    ref_shims.use_shipped_copy()
    cfg = wo.WanConfig(**PIPE_CFG)
    sd = wo.init_state_dict(cfg, seed=0)
    wrapper = ref_shims.build_reference_wrapper(cfg, sd, shift=5.0, attention_impl="flash", device=DEV)
    _, RefInteractive = ref_shims.reference_pipelines()
    T, switches = 21, [4, 16]
    g = torch.Generator().manual_seed(6)
    noise = torch.randn(1, T, 16, 60, 104, generator=g).to(torch.bfloat16).to(DEV)
    prompts = [wo.synth_prompt_embeds(cfg, 400 + i, 77 + 60 * i).to(DEV) for i in range(3)]
    rlat, rp = _run_interactive(RefInteractive, cfg, wrapper, prompts, noise, switches)
    olat, op = _run_interactive(Ours, cfg, _our_generator(cfg, sd), prompts, noise, switches)
    errs = [rel_l2(olat[:, c:c + 3], rlat[:, c:c + 3]) for c in range(0, T, 3)]
    print("1-layer latents rel-L2 per chunk vs the reference (flash-attn):", [f"{e:.2e}" for e in errs])
    assert max(errs) <= 1e-2, errs
    assert int(op.kv_cache1[0]["global_end_index"].item()) == int(rp.kv_cache1[0]["global_end_index"].item())
    assert int(op.kv_cache1[0]["local_end_index"].item()) == int(rp.kv_cache1[0]["local_end_index"].item())
    from longlive_b200.kv_ring import logical_view
    k, v = logical_view(op.kv_cache1[0], op.kv_cache1[0]["_llb_ring"])
    assert rel_l2(k, rp.kv_cache1[0]["k"]) < 1e-2 and rel_l2(v, rp.kv_cache1[0]["v"]) < 1e-2


def test_full_size_pipeline_vs_reference_flash_attn():
    """Wan2.1-T2V-1.3B shape (30 blocks), 5 chunks (cache fill + first eviction), 4-step DMD + clean pass:
    our CausalInferencePipeline vs the reference's own GPU path on the same B200."""
    from oracle import wan_oracle as wo
    from longlive_b200.kv_ring import logical_view
    from longlive_b200.pipeline import CausalInferencePipeline as Ours
    ref_shims.use_shipped_copy()
    cfg = wo.WanConfig()
    sd = wo.init_state_dict(cfg, seed=0)
    wrapper = ref_shims.build_reference_wrapper(cfg, sd, shift=5.0, attention_impl="flash", device=DEV)
    RefPipe, _ = ref_shims.reference_pipelines()
    T = 15
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(1, T, 16, 60, 104, generator=g).to(torch.bfloat16).to(DEV)
    prompt = wo.synth_prompt_embeds(cfg, 100, 200).to(DEV)
    te = lambda text_prompts: {"prompt_embeds": prompt}
    rp = RefPipe(_args(cfg), torch.device(DEV), generator=wrapper, text_encoder=te, vae=_VAE)
    with _SeededRandnLike(), torch.no_grad():
        _, rlat = rp.inference(noise, ["p"], return_latents=True)
    op = Ours(_args(cfg), torch.device(DEV), generator=_our_generator(cfg, sd), text_encoder=te)
    with _SeededRandnLike(), torch.no_grad():
        _, olat = op.inference(noise, ["p"], return_latents=True)
    errs = [rel_l2(olat[:, c:c + 3], rlat[:, c:c + 3]) for c in range(0, T, 3)]
    print("full-size latents rel-L2 per chunk vs the reference (flash-attn):", [f"{e:.2e}" for e in errs])
    assert max(errs) <= 1e-2, errs
    ring = op.kv_cache1[0]["_llb_ring"]
    assert ring.global_end == int(rp.kv_cache1[0]["global_end_index"].item()) == T * 1560
    assert ring.local_end == int(rp.kv_cache1[0]["local_end_index"].item()) == 12 * 1560
    kerr = []
    for l in (0, 29):
        k, v = logical_view(op.kv_cache1[l], ring)
        kerr.append((rel_l2(k, rp.kv_cache1[l]["k"]), rel_l2(v, rp.kv_cache1[l]["v"])))
    print("cache K/V rel-L2 vs the reference's cache (layers 0, 29):", kerr)
    assert max(max(p) for p in kerr) < 3e-2


def _streaming_rollout(cls, cfg, gen, noise, prompts):
    """Three 6-frame chunks continuing one cache (18 frames: the 12-frame window becomes a sub-range of the 33-frame
    cache), then clear_kv_cache() and a fresh sequence under another prompt."""
    sch = gen.get_scheduler()
    table = torch.cat((sch.timesteps.cpu(), torch.tensor([0.0])))
    steps = table[1000 - torch.tensor([1000, 750, 500, 250])]
    pipe = cls(denoising_step_list=steps, scheduler=sch, generator=gen, num_frame_per_block=3,
               same_step_across_blocks=False, last_step_only=False, context_noise=0,
               local_attn_size=cfg.local_attn_size, slice_last_frames=21)
    pipe.num_transformer_blocks = cfg.num_layers
    pipe._initialize_kv_cache(1, torch.bfloat16, torch.device(DEV))
    pipe._initialize_crossattn_cache(1, torch.bfloat16, torch.device(DEV))
    outs, ends = [], []
    with _SeededRandnLike(), torch.no_grad():
        for call, (start, prompt) in enumerate([(0, 0), (6, 0), (12, 0), (0, 1)]):
            if call == 3:
                pipe.clear_kv_cache()
            torch.manual_seed(100 + call)   # the exit steps are drawn with torch.randint on the device
            out = pipe.generate_chunk_with_cache(noise[:, 6 * call:6 * call + 6], {"prompt_embeds": prompts[prompt]},
                                                 current_start_frame=start, requires_grad=False)
            outs.append(out[0])
            ends.append((int(pipe.kv_cache1[0]["global_end_index"].item()), int(pipe.kv_cache1[0]["local_end_index"].item())))
    return torch.cat(outs, 1), ends


def test_streaming_training_rollout_vs_reference():
    """SURVEY 8f rank 4 (gradient-free part): our StreamingTrainingPipeline + CUDA generator vs the reference's class +
    wrapper + model with flash-attn; and the reference's class driving OUR generator (incl. its clear_kv_cache, which
    zeroes the end indices of a cache our model has adopted)."""
    import importlib
    from oracle import wan_oracle as wo
    from oracle.make_golden import PIPE_CFG
    from longlive_b200.pipeline import StreamingTrainingPipeline as Ours
    ref_shims.use_shipped_copy()
    cfg = wo.WanConfig(**PIPE_CFG)
    sd = wo.init_state_dict(cfg, seed=0)
    wrapper = ref_shims.build_reference_wrapper(cfg, sd, shift=5.0, attention_impl="flash", device=DEV)
    RefCls = importlib.import_module("pipeline.streaming_training").StreamingTrainingPipeline
    g = torch.Generator().manual_seed(8)
    noise = torch.randn(1, 24, 16, 60, 104, generator=g).to(torch.bfloat16).to(DEV)
    prompts = [wo.synth_prompt_embeds(cfg, 500 + i, 90 + 70 * i).to(DEV) for i in range(2)]
    r_out, r_ends = _streaming_rollout(RefCls, cfg, wrapper, noise, prompts)
    ours = _our_generator(cfg, sd)
    o_out, o_ends = _streaming_rollout(Ours, cfg, ours, noise, prompts)
    assert o_ends == r_ends == [(6 * 1560, 6 * 1560), (12 * 1560, 12 * 1560), (18 * 1560, 18 * 1560), (6 * 1560, 6 * 1560)]
    errs = [rel_l2(o_out[:, c:c + 3], r_out[:, c:c + 3]) for c in range(0, 24, 3)]
    print("streaming-training rollout latents rel-L2 per block vs the reference (flash-attn):", [f"{e:.2e}" for e in errs])
    assert max(errs) <= 1e-2, errs
    # the reference's pipeline class around our generator: same results as our class, bit for bit
    a_out, a_ends = _streaming_rollout(RefCls, cfg, ours, noise, prompts)
    assert a_ends == r_ends and torch.equal(a_out, o_out), rel_l2(a_out, o_out)

"""Host logic of longlive_b200.pipeline.StreamingTrainingPipeline (gradient-free rollout) against
the REAL reference class (pipeline/streaming_training.py) driven by the same recording fake generator, same RNG
seeds: identical generator-call sequence (current_start, timesteps, shapes), identical outputs and returned
timestep levels.  Needs the reference tree (build container: /root/reference, GPU box: baseline/_ref)."""
import types

import pytest
import torch

from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")


class _FakeModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.attn = torch.nn.Module()
        self.attn.max_attention_size = 32760
        self.local_attn_size = -1
        self.num_layers, self.num_heads, self.text_len, self.frame_seqlen = 2, 12, 512, 1560


class _FakeGenerator:
    """Deterministic stand-in: x0 = 0.5 * x + 0.01 * t / 1000, and a log of every call."""

    def __init__(self):
        self.model = _FakeModel()
        self.log = []

    def __call__(self, noisy_image_or_video, conditional_dict, timestep, kv_cache, crossattn_cache, current_start, **kw):
        assert not torch.is_grad_enabled()
        self.log.append((int(current_start), tuple(noisy_image_or_video.shape), float(timestep.flatten()[0]),
                         float(noisy_image_or_video.float().sum())))
        x0 = 0.5 * noisy_image_or_video + 0.01 * timestep.float().view(*timestep.shape, 1, 1, 1) / 1000
        return x0, x0


def _run(cls, same_step, last_step_only):
    from longlive_b200.scheduler import FlowMatchScheduler
    sch = FlowMatchScheduler(shift=5.0, sigma_min=0.0, extra_one_step=True)
    sch.set_timesteps(1000, training=True)
    table = torch.cat((sch.timesteps.cpu(), torch.tensor([0.0])))
    steps = table[1000 - torch.tensor([1000, 750, 500, 250])]
    gen = _FakeGenerator()
    pipe = cls(denoising_step_list=steps, scheduler=sch, generator=gen, num_frame_per_block=3,
               same_step_across_blocks=same_step, last_step_only=last_step_only, context_noise=0,
               local_attn_size=12, slice_last_frames=21)
    pipe.num_transformer_blocks = 2
    assert pipe.kv_cache_size == (12 + 21) * 1560
    pipe._initialize_kv_cache(1, torch.bfloat16, "cpu")
    pipe._initialize_crossattn_cache(1, torch.bfloat16, "cpu")
    assert pipe.kv_cache1[0]["k"].shape == (1, 33 * 1560, 12, 128)
    torch.manual_seed(7)
    outs = []
    g = torch.Generator().manual_seed(1)
    for start in (0, 6):
        noise = torch.randn(1, 6, 4, 4, 6, generator=g)
        outs.append(pipe.generate_chunk_with_cache(noise, {"prompt_embeds": None}, current_start_frame=start,
                                                   requires_grad=False, return_sim_step=True))
    pipe.kv_cache1[0]["global_end_index"].fill_(5)
    pipe.clear_kv_cache()
    assert int(pipe.kv_cache1[0]["global_end_index"].item()) == 0 and pipe.crossattn_cache[0]["is_init"] is False
    assert gen.model.attn.max_attention_size == 12 * 1560 and gen.model.local_attn_size == 12
    return gen.log, outs


@pytest.mark.parametrize("same_step,last_step_only", [(False, False), (True, False), (True, True)])
def test_rollout_matches_reference_class(same_step, last_step_only):
    import importlib
    ref_shims.install("sdpa")
    cur = torch.cuda.current_device
    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: 0
    try:
        ref_cls = importlib.import_module("pipeline.streaming_training").StreamingTrainingPipeline
    finally:
        torch.cuda.current_device = cur
    from longlive_b200.pipeline import StreamingTrainingPipeline
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self   # the reference calls .cuda() on the schedule tables (:244-252)
    try:
        with torch.no_grad():
            ref_log, ref_outs = _run(ref_cls, same_step, last_step_only)
    finally:
        torch.Tensor.cuda = real_cuda
    our_log, our_outs = _run(StreamingTrainingPipeline, same_step, last_step_only)
    assert our_log == ref_log and len(our_log) >= 2 * 2 * 2
    for a, b in zip(our_outs, ref_outs):
        assert torch.equal(a[0], b[0]) and a[1:] == b[1:]


def test_requires_grad_is_refused_loudly():
    from longlive_b200.pipeline import StreamingTrainingPipeline
    pipe = StreamingTrainingPipeline([1000, 500], scheduler=None, generator=_FakeGenerator(), local_attn_size=12)
    with pytest.raises(NotImplementedError):
        pipe.generate_chunk_with_cache(torch.zeros(1, 3, 4, 4, 6), {}, requires_grad=True)

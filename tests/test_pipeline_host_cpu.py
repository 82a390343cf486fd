"""Host logic of the drop-in pipelines, without a GPU: with a recording fake generator, the call
sequence (current_start, frames, timestep value + dtype, sink_recache flag) must equal what the
reference's own InteractiveCausalInferencePipeline issued (tests/golden/pipeline_small.pt), and the
cache objects must have the reference's keys, shapes and dtypes."""
import os
import types

import torch

from longlive_b200.pipeline import CausalInferencePipeline, InteractiveCausalInferencePipeline
from longlive_b200.scheduler import FlowMatchScheduler

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class FakeModel(torch.nn.Module):
    num_layers, num_heads, text_len, frame_seqlen = 2, 12, 512, 1560

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.attn = torch.nn.Module()
        self.attn.max_attention_size = 32760
        self.local_attn_size = -1
        self.num_frame_per_block = 1
        self.block_mask = None

    @staticmethod
    def _prepare_blockwise_causal_attn_mask(**kw):
        return None


class FakeGenerator(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.model = FakeModel()
        self.scheduler = FlowMatchScheduler(shift=5.0)
        self.calls = []

    def get_scheduler(self):
        return self.scheduler

    def forward(self, noisy_image_or_video, conditional_dict, timestep, kv_cache, crossattn_cache,
                current_start, cache_start=None, sink_recache_after_switch=False):
        self.calls.append({"current_start": int(current_start), "frames": noisy_image_or_video.shape[1],
                           "t": float(timestep.flatten()[0]), "t_dtype": str(timestep.dtype),
                           "sink_recache": bool(sink_recache_after_switch),
                           "prompt": conditional_dict["id"],
                           "cross_init": [c["is_init"] for c in crossattn_cache]})
        for c in crossattn_cache:
            c["is_init"] = True
        return noisy_image_or_video * 0.5, noisy_image_or_video * 0.25


def _args(global_sink=False):
    class MK(dict):
        __getattr__ = dict.get
    return types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=3, context_noise=0, global_sink=global_sink,
                                 model_kwargs=MK(local_attn_size=12, sink_size=3, timestep_shift=5.0))


def test_interactive_call_pattern_matches_reference_pipeline():
    gold = torch.load(os.path.join(GOLDEN, "pipeline_small.pt"))
    gen = FakeGenerator()
    pipe = InteractiveCausalInferencePipeline(_args(), torch.device("cpu"), generator=gen,
                                              text_encoder=lambda text_prompts: {"id": text_prompts[0]})
    noise = torch.zeros(1, 6, 16, 4, 4, dtype=torch.bfloat16)
    video, lat = pipe.inference(noise, text_prompts_list=[["a"], ["b"]], switch_frame_indices=[3],
                                return_latents=True)
    assert len(gen.calls) == len(gold["calls"])
    for mine, ref in zip(gen.calls, gold["calls"]):
        for k in ("current_start", "frames", "t_dtype", "sink_recache"):
            assert mine[k] == ref[k], (mine, ref)
        assert abs(mine["t"] - ref["t"]) < 1e-3
    # prompt of segment 1 is used from the recache call on, and the cross cache is re-initialised
    # both for the recache forward and for the first denoise step after it
    assert [c["prompt"] for c in gen.calls] == ["a"] * 5 + ["b"] * 6
    assert gen.calls[5]["cross_init"] == [False, False] and gen.calls[6]["cross_init"] == [False, False]
    assert gen.calls[7]["cross_init"] == [True, True]
    # reference cache contract (pipeline/causal_inference.py:255-293)
    kv, cc = pipe.kv_cache1, pipe.crossattn_cache
    assert len(kv) == 2 and kv[0]["k"].shape == (1, 12 * 1560, 12, 128) and kv[0]["k"].dtype == torch.bfloat16
    assert kv[1]["global_end_index"].dtype == torch.long and kv[1]["global_end_index"].shape == (1,)
    assert cc[0]["v"].shape == (1, 512, 12, 128)
    assert gen.model.attn.max_attention_size == 12 * 1560 and gen.model.local_attn_size == 12
    assert lat.shape == noise.shape and video.shape == noise.shape


def test_single_prompt_pipeline_call_pattern():
    gen = FakeGenerator()
    pipe = CausalInferencePipeline(_args(), torch.device("cpu"), generator=gen,
                                   text_encoder=lambda text_prompts: {"id": text_prompts[0]})
    seen = []
    pipe.renoise_fn = lambda like, block, step: (seen.append((block, step)) or torch.zeros_like(like))
    pipe.inference(torch.zeros(1, 9, 16, 4, 4, dtype=torch.bfloat16), ["a"])
    assert [c["current_start"] for c in gen.calls] == [f * 1560 for f in (0, 3, 6) for _ in range(5)]
    ts = [round(c["t"], 2) for c in gen.calls[:5]]
    assert ts == [1000.0, 937.5, 833.33, 625.0, 0.0]
    assert seen == [(b, s) for b in range(3) for s in range(3)]


def test_switch_positions_follow_reference_rule():
    """switch_frame_indices 40..200 with 3-frame chunks fire at frames 42, 81, 120, 162, 201 and
    each recaches the previous 12 frames."""
    gen = FakeGenerator()
    pipe = InteractiveCausalInferencePipeline(_args(), torch.device("cpu"), generator=gen,
                                              text_encoder=lambda text_prompts: {"id": text_prompts[0]})
    pipe.inference(torch.zeros(1, 240, 16, 2, 2, dtype=torch.bfloat16),
                   text_prompts_list=[[str(i)] for i in range(6)],
                   switch_frame_indices=[40, 80, 120, 160, 200])
    assert [s["frame"] for s in pipe.switch_log] == [42, 81, 120, 162, 201]
    assert all(s["recached_frames"] == 12 for s in pipe.switch_log)
    rec = [c for c in gen.calls if c["frames"] == 12]
    assert [c["current_start"] // 1560 for c in rec] == [30, 69, 108, 150, 189]
    assert len(gen.calls) == 80 * 5 + 5

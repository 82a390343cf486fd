"""InteractiveCausalInferencePipeline — prompt switching with KV-recache, same class name and
`inference()` signature as the reference's pipeline/interactive_causal_inference.py:20-432.

At the first chunk whose start frame is >= the next switch index (reference :237-264) the
pipeline calls `_recache_after_switch` (reference :34-106):
  * unless `global_sink`, zero the K/V of every layer (end indices are NOT reset);
  * zero the cross-attention caches and mark them uninitialised;
  * run ONE batched forward at t = context_noise over the last min(local_attn_size, cur) output
    frames under the new prompt, with sink_recache_after_switch = not global_sink, so K/V of the
    window (and, without global sink, the new sink) are recomputed in a single pass;
  * reset the cross-attention caches again (the next denoise step re-projects the same text K/V).
The flex-attention BlockMask the reference builds at :73-84 is never read by the KV path, so it is
not built here (model._prepare_blockwise_causal_attn_mask is a no-op kept for compatibility).
"""
from __future__ import annotations

from typing import List

import torch

from .causal_inference import CausalInferencePipeline, _Profiler


class InteractiveCausalInferencePipeline(CausalInferencePipeline):
    def __init__(self, args, device, *, generator=None, text_encoder=None, vae=None):
        super().__init__(args, device, generator=generator, text_encoder=text_encoder, vae=vae)
        self.global_sink = getattr(args, "global_sink", False)
        self.switch_log: List[dict] = []

    def _reset_crossattn_cache(self):
        for blk in self.crossattn_cache:
            blk["k"].zero_()
            blk["v"].zero_()
            blk["is_init"] = False

    def _recache_after_switch(self, output, current_start_frame, new_conditional_dict):
        if not self.global_sink:
            for cache in self.kv_cache1:
                cache["k"].zero_()
                cache["v"].zero_()
        self._reset_crossattn_cache()
        if current_start_frame == 0:
            return
        n = (current_start_frame if self.local_attn_size == -1
             else min(self.local_attn_size, current_start_frame))
        first = current_start_frame - n
        frames = output[:, first:current_start_frame]
        dev = next(self.generator.parameters()).device
        if frames.device != dev:
            frames = frames.to(dev)
        B = frames.shape[0]
        self.generator.model.block_mask = self.generator.model._prepare_blockwise_causal_attn_mask(
            device=dev, num_frames=n, frame_seqlen=self.frame_seq_length,
            num_frame_per_block=self.num_frame_per_block, local_attn_size=self.local_attn_size)
        context_timestep = torch.ones([B, n], device=dev, dtype=torch.int64) * self.args.context_noise
        with torch.no_grad():
            self.generator(
                noisy_image_or_video=frames, conditional_dict=new_conditional_dict,
                timestep=context_timestep, kv_cache=self.kv_cache1, crossattn_cache=self.crossattn_cache,
                current_start=first * self.frame_seq_length,
                sink_recache_after_switch=not self.global_sink)
        self._reset_crossattn_cache()
        self.switch_log.append({"frame": current_start_frame, "recached_frames": n, "first": first})

    def inference(self, noise: torch.Tensor, *, text_prompts_list: List[List[str]],
                  switch_frame_indices: List[int], return_latents: bool = False,
                  low_memory: bool = False, profile: bool = False):
        """noise [B, T, 16, H, W]; text_prompts_list[i] = prompts of segment i; segment i+1 starts at
        the first chunk whose start frame is >= switch_frame_indices[i]."""
        B, T = noise.shape[:2]
        assert len(text_prompts_list) >= 1, "text_prompts_list must not be empty"
        assert len(switch_frame_indices) == len(text_prompts_list) - 1, (
            "length of switch_frame_indices should be one less than text_prompts_list")
        assert T % self.num_frame_per_block == 0
        num_blocks = T // self.num_frame_per_block
        prof = _Profiler(profile)
        prof.start("init")
        cond_list = [self.text_encoder(text_prompts=p) for p in text_prompts_list]
        output = torch.zeros_like(noise)
        self._prepare(noise)
        self.switch_log = []
        prof.stop("init")
        prof.start("diffusion")
        segment, start = 0, 0
        switch_blocks, recache_ms = [], []
        for bi in range(num_blocks):
            prof.block_start()
            if segment < len(switch_frame_indices) and start >= switch_frame_indices[segment]:
                segment += 1
                if profile:
                    s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
                    s0.record()
                self._recache_after_switch(output, start, cond_list[segment])
                if profile:
                    s1.record(); torch.cuda.synchronize(); recache_ms.append(s0.elapsed_time(s1))
                switch_blocks.append(bi)
            F = self.num_frame_per_block
            output[:, start:start + F] = self._denoise_block(noise[:, start:start + F], cond_list[segment],
                                                             start, bi)
            prof.block_stop()
            start += F
        prof.stop("diffusion")
        prof.start("vae")
        video = self.vae.decode_to_pixel(output, use_cache=False)
        video = (video * 0.5 + 0.5).clamp(0, 1)
        prof.stop("vae")
        rep = prof.report(self.num_frame_per_block, switch_blocks=switch_blocks)
        if rep is not None:
            rep["switch_blocks"] = switch_blocks
            rep["recache_ms"] = recache_ms
            if recache_ms:
                print(f"  - Recache overhead (mean of {len(recache_ms)} switches): "
                      f"{sum(recache_ms) / len(recache_ms):.2f} ms")
        self.last_profile = rep
        return (video, output) if return_latents else video

"""StreamingTrainingPipeline — the training-side caller of the same generator contract,
same class name, constructor and `generate_chunk_with_cache()` signature as the reference's
pipeline/streaming_training.py:18-343 (used by model/streaming_training.py:254-259).

What it exercises that the inference pipelines do not:
  * a KV cache LARGER than the attention window: kv_cache_size = (local_attn_size + slice_last_frames) * 1560
    tokens (reference :48-49) while max_attention_size stays local_attn_size * 1560, so the attended set is
    "sink ++ last window" of a cache that only rolls after 33 frames (two physical ranges in our ring);
  * a chunk is generated in several calls that continue one cache (`current_start_frame`), and the cache is
    reset between sequences by ZEROING the end indices (`clear_kv_cache`, reference :291-305);
  * every block stops denoising at a randomly drawn exit step (`generate_and_sync_list`, :51-71; synchronised
    over ranks), then writes the context K/V with a re-noised clean pass (:215-235).

Scope: the rollout WITHOUT gradients (`requires_grad=False`: what the critic / data-generation passes and
every non-exit step run).  The exit step under `torch.enable_grad()` needs backward kernels for the whole
transformer step, which libllb200 does not have: `requires_grad=True` raises instead of silently returning
tensors without a graph.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


class StreamingTrainingPipeline:
    def __init__(self, denoising_step_list: List[int], scheduler, generator, num_frame_per_block: int = 3,
                 same_step_across_blocks: bool = False, last_step_only: bool = False, context_noise: int = 0,
                 **kwargs):
        self.scheduler = scheduler
        self.generator = generator
        steps = denoising_step_list
        if steps[-1] == 0:          # the trailing zero timestep is not denoised from (reference :35-36)
            steps = steps[:-1]
        self.denoising_step_list = steps
        model = generator.model
        self.num_transformer_blocks = getattr(model, "num_layers", 30)
        self.frame_seq_length = getattr(model, "frame_seqlen", 1560)
        self.num_frame_per_block = num_frame_per_block
        self.context_noise = context_noise
        self.kv_cache1 = None
        self.crossattn_cache = None
        self.same_step_across_blocks = same_step_across_blocks
        self.last_step_only = last_step_only
        self.local_attn_size = kwargs.get("local_attn_size", -1)
        slice_last_frames = int(kwargs.get("slice_last_frames", 21))
        self.kv_cache_size = (self.local_attn_size + slice_last_frames) * self.frame_seq_length

    # ------------------------------------------------------------------------------------------
    def generate_and_sync_list(self, num_blocks: int, num_denoising_steps: int, device) -> List[int]:
        """One exit-step index per block, drawn on rank 0 and broadcast (reference :51-71)."""
        rank = dist.get_rank() if dist.is_initialized() else 0
        if rank == 0:
            indices = torch.randint(low=0, high=num_denoising_steps, size=(num_blocks,), device=device)
            if self.last_step_only:
                indices = torch.full_like(indices, num_denoising_steps - 1)
        else:
            indices = torch.empty(num_blocks, dtype=torch.long, device=device)
        if dist.is_initialized():
            dist.broadcast(indices, src=0)
        return indices.tolist()

    def generate_chunk_with_cache(self, noise: torch.Tensor, conditional_dict: dict, *, current_start_frame: int = 0,
                                  requires_grad: bool = True, return_sim_step: bool = False
                                  ) -> Tuple[torch.Tensor, Optional[int], Optional[int]]:
        """noise [B, chunk_frames, C, H, W] -> (output, denoised_timestep_from, denoised_timestep_to[, steps]);
        continues self.kv_cache1 / self.crossattn_cache at `current_start_frame` (reference :73-257)."""
        if requires_grad:
            raise NotImplementedError(
                "longlive_b200 has no backward kernels: only the gradient-free rollout of StreamingTrainingPipeline is "
                "implemented (pass requires_grad=False); the exit step under enable_grad stays with the reference")
        if self.kv_cache1 is None or self.crossattn_cache is None:
            raise RuntimeError("call _initialize_kv_cache / _initialize_crossattn_cache first (the reference's trainer "
                               "does, model/streaming_training.py)")
        B, chunk_frames = noise.shape[:2]
        assert chunk_frames % self.num_frame_per_block == 0
        num_blocks = chunk_frames // self.num_frame_per_block
        dev = noise.device
        output = torch.zeros_like(noise)
        n_steps = len(self.denoising_step_list)
        exit_flags = self.generate_and_sync_list(num_blocks, n_steps, device=dev)
        self.generator.model.local_attn_size = int(self.local_attn_size)
        self._set_all_modules_max_attention_size(int(self.local_attn_size))
        F = self.num_frame_per_block
        with torch.no_grad():
            for block in range(num_blocks):
                local = block * F
                cur = (current_start_frame + local) * self.frame_seq_length
                x = noise[:, local:local + F]
                exit_at = exit_flags[0] if self.same_step_across_blocks else exit_flags[block]
                for step, t in enumerate(self.denoising_step_list):
                    timestep = torch.ones([B, F], device=dev, dtype=torch.int64) * t
                    _, denoised = self.generator(noisy_image_or_video=x, conditional_dict=conditional_dict,
                                                 timestep=timestep, kv_cache=self.kv_cache1,
                                                 crossattn_cache=self.crossattn_cache, current_start=cur)
                    if step == exit_at:
                        break
                    if step < n_steps - 1:   # re-noise to the next step's level (reference :181-190)
                        flat = denoised.flatten(0, 1)
                        nt = self.denoising_step_list[step + 1] * torch.ones([B * F], device=dev, dtype=torch.long)
                        x = self.scheduler.add_noise(flat, torch.randn_like(flat), nt).unflatten(0, denoised.shape[:2])
                output[:, local:local + F] = denoised
                # context pass: clean (context_noise-level) K / V of this block into the cache (reference :215-235)
                ctx_t = torch.ones_like(timestep) * self.context_noise
                flat = denoised.flatten(0, 1)
                ctx_x = self.scheduler.add_noise(flat, torch.randn_like(flat), ctx_t.flatten(0, 1)
                                                 ).unflatten(0, denoised.shape[:2])
                self.generator(noisy_image_or_video=ctx_x, conditional_dict=conditional_dict, timestep=ctx_t,
                               kv_cache=self.kv_cache1, crossattn_cache=self.crossattn_cache, current_start=cur)
        # which noise levels the (shared) exit step went from / to (reference :241-253)
        if not self.same_step_across_blocks:
            t_from, t_to = None, None
        else:
            sched_t = self.scheduler.timesteps.to(dev)

            def level(step):
                t = torch.as_tensor(self.denoising_step_list[step]).to(dev)
                return 1000 - torch.argmin((sched_t - t).abs(), dim=0).item()
            t_from = level(exit_flags[0])
            t_to = 0 if exit_flags[0] == n_steps - 1 else level(exit_flags[0] + 1)
        if return_sim_step:
            return output, t_from, t_to, exit_flags[0] + 1
        return output, t_from, t_to

    # ------------------------------------------------------------------------------------------
    def _initialize_kv_cache(self, batch_size, dtype, device):
        """Per layer {"k","v": zeros[B, kv_cache_size, heads, 128], "global_end_index", "local_end_index"}
        (reference :259-274).  The index tensors are views of one [layers, 2] tensor (two fills per forward)."""
        model = self.generator.model
        heads = getattr(model, "num_heads", 12)
        n = self.num_transformer_blocks
        index = torch.zeros(n, 2, dtype=torch.long, device=device)
        self.kv_cache1 = [{
            "k": torch.zeros([batch_size, self.kv_cache_size, heads, 128], dtype=dtype, device=device),
            "v": torch.zeros([batch_size, self.kv_cache_size, heads, 128], dtype=dtype, device=device),
            "global_end_index": index[i, 0:1], "local_end_index": index[i, 1:2]} for i in range(n)]
        self.kv_cache1[0]["_llb_index_tensor"] = index

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        model = self.generator.model
        heads, tl = getattr(model, "num_heads", 12), getattr(model, "text_len", 512)
        self.crossattn_cache = [{
            "k": torch.zeros([batch_size, tl, heads, 128], dtype=dtype, device=device),
            "v": torch.zeros([batch_size, tl, heads, 128], dtype=dtype, device=device),
            "is_init": False} for _ in range(self.num_transformer_blocks)]

    def clear_kv_cache(self):
        """Start a new sequence in the SAME allocation: zero K / V and the end indices, mark the text K / V stale
        (reference :276-305); the host-side ring state goes with the indices."""
        if self.kv_cache1 is not None:
            for blk in self.kv_cache1:
                blk["k"].zero_(); blk["v"].zero_()
            c0 = self.kv_cache1[0]
            c0["_llb_index_tensor"].zero_()
            c0.pop("_llb_ring", None); c0.pop("_llb_published", None)
        if self.crossattn_cache is not None:
            for blk in self.crossattn_cache:
                blk["k"].zero_(); blk["v"].zero_(); blk["is_init"] = False

    def _set_all_modules_max_attention_size(self, local_attn_size_value: int):
        if isinstance(local_attn_size_value, (list, tuple)):
            raise ValueError("_set_all_modules_max_attention_size expects an int, got list/tuple.")
        target = 32760 if int(local_attn_size_value) == -1 else int(local_attn_size_value) * self.frame_seq_length
        for _, module in self.generator.model.named_modules():
            if hasattr(module, "max_attention_size"):
                module.max_attention_size = target

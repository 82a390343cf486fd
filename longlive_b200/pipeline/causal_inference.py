"""CausalInferencePipeline — same class name, constructor and `inference()` signature as the
reference's pipeline/causal_inference.py:13-253, driving the B200-native generator.

Per 3-latent-frame chunk: the 4-step DMD denoise loop (reference :154-188), each step one model
forward at the same `current_start`, re-noising between steps with the scheduler's add_noise and
fresh N(0,1) noise from the global torch RNG (:173-178), then one "clean context" forward at
t = context_noise whose output is discarded and which leaves clean K/V in the cache (:193-200).

Differences that do not change results: caches are allocated with the reference's keys/shapes but
the 30 `global_end_index` / `local_end_index` tensors are views of one [layers, 2] tensor (so
publishing them costs two fills, not sixty), and the model keeps the ring rotation on the host.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from ..wrapper import WanDiffusionWrapper


class _NullTextEncoder:
    def __call__(self, text_prompts):
        raise RuntimeError("no text encoder was injected: pass text_encoder=<callable returning "
                           "{'prompt_embeds': [B,512,4096]}> (umT5 is outside this package's scope)")


class _IdentityVAE:
    """Stand-in for WanVAEWrapper: the VAE decode is outside the denoising hot path."""

    def decode_to_pixel(self, latent, use_cache=False):
        return latent


class CausalInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None):
        super().__init__()
        mk = getattr(args, "model_kwargs", {})
        self.generator = generator if generator is not None else WanDiffusionWrapper(**dict(mk), is_causal=True)
        self.text_encoder = text_encoder if text_encoder is not None else _NullTextEncoder()
        self.vae = vae if vae is not None else _IdentityVAE()

        self.scheduler = self.generator.get_scheduler()
        steps = torch.tensor(list(args.denoising_step_list), dtype=torch.long)
        if args.warp_denoising_step:
            # warp the nominal steps through the shifted schedule (reference :33-37)
            table = torch.cat((self.scheduler.timesteps.cpu(), torch.tensor([0.0], dtype=torch.float32)))
            steps = table[1000 - steps]
        self.denoising_step_list = steps

        model = self.generator.model
        self.num_transformer_blocks = getattr(model, "num_layers", 30)
        self.frame_seq_length = getattr(model, "frame_seqlen", 1560)
        self.kv_cache1 = None
        self.crossattn_cache = None
        self.args = args
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 1)
        self.local_attn_size = _get(mk, "local_attn_size", -1)
        if self.num_frame_per_block > 1:
            model.num_frame_per_block = self.num_frame_per_block
        # test hook: replaces torch.randn_like in the re-noising step when set
        self.renoise_fn: Optional[Callable[[torch.Tensor, int, int], torch.Tensor]] = None
        self.last_profile = None

    # ------------------------------------------------------------------------------------------
    def _kv_cache_size(self, num_output_frames: int) -> int:
        local = _get(getattr(self.args, "model_kwargs", {}), "local_attn_size", -1)
        frames = local if local != -1 else num_output_frames
        return frames * self.frame_seq_length

    def _prepare(self, noise: torch.Tensor):
        batch_size, num_output_frames = noise.shape[:2]
        self._initialize_kv_cache(batch_size, noise.dtype, noise.device,
                                  kv_cache_size_override=self._kv_cache_size(num_output_frames))
        self._initialize_crossattn_cache(batch_size, noise.dtype, noise.device)
        self.generator.model.local_attn_size = self.local_attn_size
        self._set_all_modules_max_attention_size(self.local_attn_size)

    def _denoise_block(self, noisy_input, cond, start_frame: int, block_index: int):
        """4-step DMD loop + clean-context pass for one chunk; returns the denoised latents."""
        B, F = noisy_input.shape[:2]
        dev = noisy_input.device
        cur = start_frame * self.frame_seq_length
        n_steps = len(self.denoising_step_list)
        for index, current_timestep in enumerate(self.denoising_step_list):
            timestep = torch.ones([B, F], device=dev, dtype=torch.int64) * current_timestep
            _, denoised = self.generator(
                noisy_image_or_video=noisy_input, conditional_dict=cond, timestep=timestep,
                kv_cache=self.kv_cache1, crossattn_cache=self.crossattn_cache, current_start=cur)
            if index < n_steps - 1:
                flat = denoised.flatten(0, 1)
                eps = (self.renoise_fn(flat, block_index, index) if self.renoise_fn is not None
                       else torch.randn_like(flat))
                next_t = self.denoising_step_list[index + 1] * torch.ones([B * F], device=dev, dtype=torch.long)
                noisy_input = self.scheduler.add_noise(flat, eps, next_t).unflatten(0, denoised.shape[:2])
        context_timestep = torch.ones_like(timestep) * self.args.context_noise
        self.generator(
            noisy_image_or_video=denoised, conditional_dict=cond, timestep=context_timestep,
            kv_cache=self.kv_cache1, crossattn_cache=self.crossattn_cache, current_start=cur)
        return denoised

    # ------------------------------------------------------------------------------------------
    def inference(self, noise: torch.Tensor, text_prompts: List[str], return_latents: bool = False,
                  profile: bool = False, low_memory: bool = False):
        """noise [B, T, 16, H, W] -> video (and latents [B, T, 16, H, W] if return_latents)."""
        B, T = noise.shape[:2]
        assert T % self.num_frame_per_block == 0
        num_blocks = T // self.num_frame_per_block
        cond = self.text_encoder(text_prompts=text_prompts)
        output = torch.zeros_like(noise)
        prof = _Profiler(profile)
        prof.start("init")
        self._prepare(noise)
        prof.stop("init")
        prof.start("diffusion")
        start = 0
        for bi in range(num_blocks):
            prof.block_start()
            F = self.num_frame_per_block
            output[:, start:start + F] = self._denoise_block(noise[:, start:start + F], cond, start, bi)
            prof.block_stop()
            start += F
        prof.stop("diffusion")
        prof.start("vae")
        video = self.vae.decode_to_pixel(output, use_cache=False)
        video = (video * 0.5 + 0.5).clamp(0, 1)
        prof.stop("vae")
        self.last_profile = prof.report(self.num_frame_per_block, switch_blocks=())
        return (video, output) if return_latents else video

    # ------------------------------------------------------------------------------------------
    def _initialize_kv_cache(self, batch_size, dtype, device, kv_cache_size_override: int | None = None):
        """Reference :255-279: per layer {"k","v": zeros[B, size, heads, 128], "global_end_index",
        "local_end_index": int64[1]}."""
        if kv_cache_size_override is not None:
            size = kv_cache_size_override
        else:
            size = self.local_attn_size * self.frame_seq_length if self.local_attn_size != -1 else 32760
        model = self.generator.model
        heads, hd = getattr(model, "num_heads", 12), 128
        n = self.num_transformer_blocks
        old = self.kv_cache1
        if (old is not None and not hasattr(model, "allocate_kv_cache") and len(old) == n
                and old[0]["k"].shape == (batch_size, size, heads, hd) and old[0]["k"].dtype == dtype and old[0]["k"].device == torch.device(device)
                and "_llb_index_tensor" in old[0]):
            # same geometry as the previous video: re-zero in place (keeps device pointers, so the
            # model's captured CUDA graphs stay valid) instead of re-allocating 3.5 GB
            for c in old:
                c["k"].zero_(); c["v"].zero_()
            old[0]["_llb_index_tensor"].zero_()
            old[0].pop("_llb_ring", None); old[0].pop("_llb_published", None)
            return
        if hasattr(model, "allocate_kv_cache"):  # head-parallel model: head-sharded symmetric ring
            self.kv_cache1 = model.allocate_kv_cache(batch_size, size, dtype, device)
            return
        index = torch.zeros(n, 2, dtype=torch.long, device=device)
        cache = []
        for i in range(n):
            cache.append({
                "k": torch.zeros([batch_size, size, heads, hd], dtype=dtype, device=device),
                "v": torch.zeros([batch_size, size, heads, hd], dtype=dtype, device=device),
                "global_end_index": index[i, 0:1],
                "local_end_index": index[i, 1:2],
            })
        cache[0]["_llb_index_tensor"] = index
        self.kv_cache1 = cache

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        """Reference :281-293."""
        model = self.generator.model
        heads, hd, tl = getattr(model, "num_heads", 12), 128, getattr(model, "text_len", 512)
        old = self.crossattn_cache
        if (old is not None and len(old) == self.num_transformer_blocks
                and old[0]["k"].shape == (batch_size, tl, heads, hd) and old[0]["k"].dtype == dtype
                and old[0]["k"].device == torch.device(device)):
            for c in old:
                c["k"].zero_(); c["v"].zero_(); c["is_init"] = False
            return
        self.crossattn_cache = [{
            "k": torch.zeros([batch_size, tl, heads, hd], dtype=dtype, device=device),
            "v": torch.zeros([batch_size, tl, heads, hd], dtype=dtype, device=device),
            "is_init": False,
        } for _ in range(self.num_transformer_blocks)]

    def _set_all_modules_max_attention_size(self, local_attn_size_value: int):
        """Reference :295-329: every module that carries `max_attention_size` gets
        local_attn_size * frame_seq_length (or 32760 for global attention)."""
        target = 32760 if local_attn_size_value == -1 else int(local_attn_size_value) * self.frame_seq_length
        for _, module in self.generator.model.named_modules():
            if hasattr(module, "max_attention_size"):
                module.max_attention_size = target


def _get(mk, name, default):
    if isinstance(mk, dict):
        return mk.get(name, default)
    return getattr(mk, name, default)


class _Profiler:
    """CUDA-event timers with the reference's metric definitions (reference :97-107, :202-248):
    steady-state inter-frame latency = mean block time excluding block 0 (and prompt-switch
    blocks) / num_frame_per_block."""

    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.ev = {}
        self.blocks = []
        self._bs = None

    def start(self, name):
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True); e.record(); self.ev[name] = [e, None]

    def stop(self, name):
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True); e.record(); self.ev[name][1] = e

    def block_start(self):
        if self.enabled:
            self._bs = torch.cuda.Event(enable_timing=True); self._bs.record()

    def block_stop(self):
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True); e.record(); self.blocks.append((self._bs, e))

    def report(self, frames_per_block: int, switch_blocks=()):
        if not self.enabled:
            return None
        torch.cuda.synchronize()
        times = {k: a.elapsed_time(b) for k, (a, b) in self.ev.items() if b is not None}
        bt = [a.elapsed_time(b) for a, b in self.blocks]
        steady = [t for i, t in enumerate(bt) if i > 0 and i not in set(switch_blocks)] or bt[1:] or bt
        avg = sum(steady) / max(1, len(steady))
        rep = {"init_ms": times.get("init"), "diffusion_ms": times.get("diffusion"), "vae_ms": times.get("vae"),
               "block_ms": bt, "steady_block_ms": avg, "inter_frame_latency_ms": avg / frames_per_block,
               "video_fps_steady": 4000.0 * frames_per_block / avg if avg else None}
        print("Profiling results:")
        print(f"  - Initialization/caching time: {rep['init_ms']:.2f} ms")
        print(f"  - Diffusion generation time: {rep['diffusion_ms']:.2f} ms")
        print(f"  - Steady-state inter-frame latency: {rep['inter_frame_latency_ms']:.2f} ms/frame "
              f"(avg block time {avg:.2f} ms for {frames_per_block} frames)")
        return rep

"""WanDiffusionWrapper — the `generator` object the LongLive pipelines call
(reference: utils/wan_wrapper.py:120-322, KV branch of forward :247-257, :291-300).

    flow_pred, pred_x0 = generator(noisy_image_or_video [B,F,16,H,W], conditional_dict,
                                   timestep [B,F], kv_cache, crossattn_cache, current_start, ...)

The transformer step runs in longlive_b200.model.CausalWanModel (libllb200 kernels); the
flow -> x0 conversion (x0 = x_t - sigma_t * flow, in float64 like the reference :175-199) is a
three-op elementwise expression on a [B*F,16,H,W] tensor and stays in PyTorch.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .model import CausalWanModel
from .scheduler import FlowMatchScheduler


class WanDiffusionWrapper(torch.nn.Module):
    def __init__(self, model: Optional[CausalWanModel] = None, timestep_shift: float = 8.0,
                 local_attn_size: int = -1, sink_size: int = 0, is_causal: bool = True,
                 model_name: str = "Wan2.1-T2V-1.3B", **model_kwargs):
        super().__init__()
        if not is_causal:
            raise NotImplementedError("only the causal generator is on the LongLive inference path")
        if model is None:
            # random-init weights of the named architecture; load a checkpoint with
            # generator.model.load_state_dict(reference_state_dict)
            model = CausalWanModel(local_attn_size=local_attn_size, sink_size=sink_size, **model_kwargs)
        self.model = model
        self.model.eval()
        self.uniform_timestep = False
        self.scheduler = FlowMatchScheduler(shift=timestep_shift, sigma_min=0.0, extra_one_step=True)
        self.scheduler.set_timesteps(1000, training=True)
        las = local_attn_size if isinstance(local_attn_size, int) else -1
        self.seq_len = 1560 * las if las > 21 else 32760

    def get_scheduler(self) -> FlowMatchScheduler:
        return self.scheduler

    def _convert_flow_pred_to_x0(self, flow_pred, xt, timestep):
        sch = self.scheduler
        idx = sch.sigma_index(timestep.to(torch.float64) if timestep.dtype != torch.float64 else timestep)
        sigma_t = sch.sigmas.double()[idx].reshape(-1, 1, 1, 1)
        return (xt.double() - sigma_t * flow_pred.double()).to(flow_pred.dtype)

    @torch.no_grad()
    def forward(self, noisy_image_or_video: torch.Tensor, conditional_dict: dict, timestep: torch.Tensor,
                kv_cache: Optional[List[dict]] = None, crossattn_cache: Optional[List[dict]] = None,
                current_start: Optional[int] = None, classify_mode: bool = False,
                concat_time_embeddings: bool = False, clean_x=None, aug_t=None,
                cache_start: Optional[int] = None, sink_recache_after_switch: bool = False):
        if kv_cache is None or classify_mode or clean_x is not None:
            raise NotImplementedError("longlive_b200 implements the KV-cache inference call only")
        prompt_embeds = conditional_dict["prompt_embeds"]
        flow_pred = self.model(
            noisy_image_or_video.permute(0, 2, 1, 3, 4), t=timestep, context=prompt_embeds,
            seq_len=self.seq_len, kv_cache=kv_cache, crossattn_cache=crossattn_cache,
            current_start=current_start, cache_start=cache_start,
            sink_recache_after_switch=sink_recache_after_switch).permute(0, 2, 1, 3, 4)
        pred_x0 = self._convert_flow_pred_to_x0(
            flow_pred.flatten(0, 1), noisy_image_or_video.flatten(0, 1), timestep.flatten(0, 1)
        ).unflatten(0, flow_pred.shape[:2])
        return flow_pred, pred_x0

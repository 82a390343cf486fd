"""Flow-matching schedule used by the LongLive pipelines (host-side, tiny tensors).

Behaviour follows the reference's utils/scheduler.py::FlowMatchScheduler as configured by
WanDiffusionWrapper (utils/wan_wrapper.py:142-145): shift = timestep_shift, sigma_min = 0,
extra_one_step = True, 1000 training timesteps.  Only what the inference pipelines touch is
provided: `timesteps`, `sigmas`, `add_noise` (scheduler.py:159-176) and sigma lookup by nearest
timestep (wan_wrapper.py:191-194).
"""
from __future__ import annotations

import torch


class FlowMatchScheduler:
    def __init__(self, shift: float = 5.0, sigma_min: float = 0.0, extra_one_step: bool = True,
                 num_train_timesteps: int = 1000):
        self.shift, self.sigma_min, self.extra_one_step = shift, sigma_min, extra_one_step
        self.num_train_timesteps = num_train_timesteps
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, num_inference_steps: int = 1000, denoising_strength: float = 1.0,
                      training: bool = False):
        start = self.sigma_min + (1.0 - self.sigma_min) * denoising_strength
        if self.extra_one_step:
            sig = torch.linspace(start, self.sigma_min, num_inference_steps + 1)[:-1]
        else:
            sig = torch.linspace(start, self.sigma_min, num_inference_steps)
        self.sigmas = self.shift * sig / (1 + (self.shift - 1) * sig)
        self.timesteps = self.sigmas * self.num_train_timesteps

    def _to(self, device):
        if self.sigmas.device != device:
            self.sigmas = self.sigmas.to(device)
            self.timesteps = self.timesteps.to(device)

    def sigma_index(self, timestep: torch.Tensor) -> torch.Tensor:
        self._to(timestep.device)
        return torch.argmin((self.timesteps.unsqueeze(0) - timestep.unsqueeze(1)).abs(), dim=1)

    def add_noise(self, original_samples, noise, timestep):
        """x_t = (1 - sigma) x0 + sigma eps for the nearest tabulated timestep."""
        if timestep.ndim == 2:
            timestep = timestep.flatten(0, 1)
        self._to(noise.device)
        sigma = self.sigmas[self.sigma_index(timestep)].reshape(-1, 1, 1, 1)
        return ((1 - sigma) * original_samples + sigma * noise).type_as(noise)

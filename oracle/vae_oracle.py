"""TEST INFRASTRUCTURE - CPU restatement of the reference's streaming VAE decoder (the step right after the
denoising path).  Only tests/, __graft_entry__.smoke() and bench.py's
baseline legs may import this module; the product path (longlive_b200/vae.py -> libllb200.so) never does.

What is restated, with the reference lines it follows (all under /root/reference):
  * WanVAE_.cached_decode / decode          wan/modules/vae.py:545-593   un-scale, conv2 (1x1x1), then the
                                                                         decoder ONE latent frame per call
  * Decoder3d.forward                       wan/modules/vae.py:423-472   conv1, middle, upsamples, head
  * CausalConv3d.forward + the feat_cache   wan/modules/vae.py:28-36,    every 3-tap causal conv sees the last two
    protocol of its callers                 202-220, 426-438, 455-470    input frames of the previous calls
  * ResidualBlock / AttentionBlock          wan/modules/vae.py:186-262
  * Resample (upsample2d / upsample3d)      wan/modules/vae.py:101-138   incl. the first-call 'Rep' behaviour:
        call 0 skips time_conv altogether, call 1 runs it on zero history and the frame of call 0 never
        enters the history
  * RMS_norm, Upsample                      wan/modules/vae.py:39-63
  * WanVAEWrapper.decode_to_pixel           utils/wan_wrapper.py:96-117  scale, .float().clamp_(-1, 1), permute

The per-conv cache handling is written once: every causal conv keeps `hist`, the last two frames of its
input stream (zeros before the stream starts).  That is equivalent to the reference's cache_x bookkeeping
(cache = x[:, :, -2:], topped up with the previous cache's last frame when the call brought one frame;
missing frames are zero padding), which tests/test_vae_oracle_cpu.py pins bit-for-bit against the reference
module itself (imported from /root/reference in this container; committed fixture tests/golden/vae_small.pt
for the GPU box).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class VaeConfig:
    """Wan2.1 VAE shape (wan/modules/vae.py:612-626): dim 96, z 16, multipliers (1, 2, 4, 4), two residual
    blocks per stage (+1 in the decoder), temporal upsampling in the first two decoder stages."""
    dim: int = 96
    z_dim: int = 16
    dim_mult: tuple = (1, 2, 4, 4)
    num_res_blocks: int = 2
    temporal_upsample: tuple = (True, True, False)  # decoder order = reversed encoder downsample flags

    @property
    def decoder_dims(self) -> List[int]:
        return [self.dim * u for u in (self.dim_mult[-1],) + tuple(reversed(self.dim_mult))]


LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
               0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
              3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


def decoder_plan(cfg: VaeConfig) -> List[tuple]:
    """The decoder as a flat list of steps, in execution order (Decoder3d.__init__, vae.py:389-421):
    ("conv", prefix, cin, cout) | ("res", prefix, cin, cout) | ("attn", prefix, c) |
    ("up", prefix, c, temporal) | ("head", prefix, c)."""
    dims = cfg.decoder_dims
    plan: List[tuple] = [("conv", "decoder.conv1", cfg.z_dim, dims[0]),
                         ("res", "decoder.middle.0", dims[0], dims[0]),
                         ("attn", "decoder.middle.1", dims[0]),
                         ("res", "decoder.middle.2", dims[0], dims[0])]
    idx = 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(cfg.num_res_blocks + 1):
            plan.append(("res", f"decoder.upsamples.{idx}", cin, cout))
            idx += 1
            cin = cout
        if i != len(cfg.dim_mult) - 1:
            plan.append(("up", f"decoder.upsamples.{idx}", cout, bool(cfg.temporal_upsample[i])))
            idx += 1
    plan.append(("head", "decoder.head", dims[-1]))
    return plan


def init_state_dict(cfg: VaeConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random decoder weights under the reference's parameter names (only the keys decode touches).
    Scaled so activations stay O(1) through the stack; the attention output projection is NOT zero (the
    reference zero-initialises it, vae.py:236, which would hide the attention block from parity tests)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, *k):
        fan_in = cin
        for d in k:
            fan_in *= d
        sd[name + ".weight"] = (torch.randn(cout, cin, *k, generator=g) / fan_in ** 0.5).to(dtype)
        sd[name + ".bias"] = (0.1 * torch.randn(cout, generator=g)).to(dtype)

    def gamma(name, c, nd):
        sd[name + ".gamma"] = (1.0 + 0.1 * torch.randn(c, *([1] * nd), generator=g)).to(dtype)

    conv("conv2", cfg.z_dim, cfg.z_dim, 1, 1, 1)
    for step in decoder_plan(cfg):
        kind, p = step[0], step[1]
        if kind == "conv":
            conv(p, step[3], step[2], 3, 3, 3)
        elif kind == "res":
            cin, cout = step[2], step[3]
            gamma(p + ".residual.0", cin, 3)
            conv(p + ".residual.2", cout, cin, 3, 3, 3)
            gamma(p + ".residual.3", cout, 3)
            conv(p + ".residual.6", cout, cout, 3, 3, 3)
            if cin != cout:
                conv(p + ".shortcut", cout, cin, 1, 1, 1)
        elif kind == "attn":
            c = step[2]
            gamma(p + ".norm", c, 2)
            conv(p + ".to_qkv", 3 * c, c, 1, 1)
            conv(p + ".proj", c, c, 1, 1)
        elif kind == "up":
            c, temporal = step[2], step[3]
            conv(p + ".resample.1", c // 2, c, 3, 3)
            if temporal:
                conv(p + ".time_conv", 2 * c, c, 3, 1, 1)
        elif kind == "head":
            gamma(p + ".0", step[2], 3)
            conv(p + ".2", 3, step[2], 3, 3, 3)
    return sd


def rms_norm(x: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    """RMS_norm.forward (vae.py:51-54), channel-first, bias == 0."""
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma + 0.


@dataclass
class VaeDecoderOracle:
    cfg: VaeConfig
    sd: Dict[str, torch.Tensor]
    hist: Dict[str, Optional[torch.Tensor]] = field(default_factory=dict)   # conv name -> last 2 input frames
    up_calls: Dict[str, int] = field(default_factory=dict)                   # temporal upsampler -> calls so far

    def to(self, device):
        self.sd = {k: v.to(device) for k, v in self.sd.items()}
        return self

    def clear_cache(self):
        """WanVAE_.clear_cache (vae.py:602-609)."""
        self.hist.clear()
        self.up_calls.clear()

    # --- CausalConv3d.forward with the 2-frame history its callers maintain
    def _causal_conv(self, name: str, x: torch.Tensor) -> torch.Tensor:
        w, b = self.sd[name + ".weight"], self.sd[name + ".bias"]
        kt, kh, kw = w.shape[2:]
        if kt == 1:
            return F.conv3d(F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, 0, 0)), w, b)
        h = self.hist.get(name)
        if h is None:
            h = torch.zeros_like(x[:, :, :1]).repeat(1, 1, 2, 1, 1)
        xin = torch.cat([h, x], dim=2)                       # zero frames == the causal zero padding
        self.hist[name] = xin[:, :, -2:].clone()
        return F.conv3d(F.pad(xin, (kw // 2, kw // 2, kh // 2, kh // 2, 0, 0)), w, b)

    def _res(self, p: str, x: torch.Tensor, cin: int, cout: int) -> torch.Tensor:
        """ResidualBlock.forward (vae.py:202-220)."""
        h = x if cin == cout else F.conv3d(x, self.sd[p + ".shortcut.weight"], self.sd[p + ".shortcut.bias"])
        y = F.silu(rms_norm(x, self.sd[p + ".residual.0.gamma"]))
        y = self._causal_conv(p + ".residual.2", y)
        y = F.silu(rms_norm(y, self.sd[p + ".residual.3.gamma"]))
        y = self._causal_conv(p + ".residual.6", y)
        return y + h

    def _attn(self, p: str, x: torch.Tensor) -> torch.Tensor:
        """AttentionBlock.forward (vae.py:240-262): one head of width C over the H*W positions of each frame."""
        b, c, t, hh, ww = x.shape
        y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww)
        y = rms_norm(y, self.sd[p + ".norm.gamma"])
        qkv = F.conv2d(y, self.sd[p + ".to_qkv.weight"], self.sd[p + ".to_qkv.bias"])
        q, k, v = qkv.reshape(b * t, 1, 3 * c, hh * ww).permute(0, 1, 3, 2).contiguous().chunk(3, dim=-1)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.squeeze(1).permute(0, 2, 1).reshape(b * t, c, hh, ww)
        o = F.conv2d(o, self.sd[p + ".proj.weight"], self.sd[p + ".proj.bias"])
        o = o.reshape(b, t, c, hh, ww).permute(0, 2, 1, 3, 4)
        return o + x

    def _up(self, p: str, x: torch.Tensor, c: int, temporal: bool) -> torch.Tensor:
        """Resample.forward, modes upsample2d / upsample3d (vae.py:101-138)."""
        b, _, t, hh, ww = x.shape
        if temporal:
            n = self.up_calls.get(p, 0)
            self.up_calls[p] = n + 1
            if n > 0:  # the very first call passes through (feat_cache[idx] = 'Rep')
                y = self._causal_conv(p + ".time_conv", x)          # [b, 2c, t, h, w]
                y = y.reshape(b, 2, c, t, hh, ww)
                x = torch.stack((y[:, 0], y[:, 1]), dim=3).reshape(b, c, 2 * t, hh, ww)
                t = 2 * t
        y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww)
        y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest").type_as(y)
        y = F.conv2d(y, self.sd[p + ".resample.1.weight"], self.sd[p + ".resample.1.bias"], padding=1)
        return y.reshape(b, t, c // 2, 2 * hh, 2 * ww).permute(0, 2, 1, 3, 4)

    def decode_one(self, x: torch.Tensor) -> torch.Tensor:
        """Decoder3d.forward on ONE latent frame [1, z, 1, h, w] (already through conv2)."""
        for step in decoder_plan(self.cfg):
            kind, p = step[0], step[1]
            if kind == "conv":
                x = self._causal_conv(p, x)
            elif kind == "res":
                x = self._res(p, x, step[2], step[3])
            elif kind == "attn":
                x = self._attn(p, x)
            elif kind == "up":
                x = self._up(p, x, step[2], step[3])
            else:
                x = F.silu(rms_norm(x, self.sd[p + ".0.gamma"]))
                x = self._causal_conv(p + ".2", x)
        return x

    def cached_decode(self, z: torch.Tensor, scale) -> torch.Tensor:
        """WanVAE_.cached_decode (vae.py:571-593): z [1, z_dim, T, h, w] -> [1, 3, T', 8h, 8w]; the caches
        persist across calls (T' = 4T, minus 3 on the stream's first call)."""
        zd = self.cfg.z_dim
        z = z / scale[1].view(1, zd, 1, 1, 1) + scale[0].view(1, zd, 1, 1, 1)
        x = F.conv3d(z, self.sd["conv2.weight"], self.sd["conv2.bias"])
        outs = [self.decode_one(x[:, :, i:i + 1]) for i in range(x.shape[2])]
        return torch.cat(outs, dim=2)

    def decode(self, z: torch.Tensor, scale) -> torch.Tensor:
        """WanVAE_.decode (vae.py:545-569): same, on a fresh cache, cleared again afterwards."""
        self.clear_cache()
        out = self.cached_decode(z, scale)
        self.clear_cache()
        return out

    def decode_to_pixel(self, latent: torch.Tensor, use_cache: bool = False) -> torch.Tensor:
        """WanVAEWrapper.decode_to_pixel (utils/wan_wrapper.py:96-117): latent [B, T, z, h, w] ->
        video [B, T', 3, 8h, 8w] float32 in [-1, 1]."""
        zs = latent.permute(0, 2, 1, 3, 4)
        mean = torch.tensor(LATENT_MEAN[:self.cfg.z_dim], dtype=torch.float32)
        std = torch.tensor(LATENT_STD[:self.cfg.z_dim], dtype=torch.float32)
        scale = [mean.to(device=latent.device, dtype=latent.dtype),
                 1.0 / std.to(device=latent.device, dtype=latent.dtype)]
        fn = self.cached_decode if use_cache else self.decode
        out = [fn(u.unsqueeze(0), scale).float().clamp_(-1, 1).squeeze(0) for u in zs]
        return torch.stack(out, dim=0).permute(0, 2, 1, 3, 4)

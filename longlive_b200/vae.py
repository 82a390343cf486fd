"""Streaming VAE decoder on libllb200.so - the step right after the denoising path.

Host-side mirror of the reference's decode interface:
  * ``WanVAEWrapper.decode_to_pixel(latent [B, T, 16, h, w], use_cache) -> [B, T', 3, 8h, 8w] float32``
    (utils/wan_wrapper.py:96-117), backed by
  * ``WanVAEDecoder.cached_decode / decode / clear_cache`` (wan/modules/vae.py:545-609),
with the reference's parameter names (``conv2.*``, ``decoder.conv1.*``, ``decoder.middle.*``,
``decoder.upsamples.*``, ``decoder.head.*``), so ``Wan2.1_VAE.pth`` loads with ``load_state_dict``.

B200 design:
  * activations are channels-last bf16 ``[frames, H, W, Cp]`` (Cp = channels padded to a multiple of 64 with
    zeros), so a 3-D convolution is an implicit GEMM whose A operand is a TMA box of pixels x channels
    (``llb_conv3d``, tcgen05, taps = K loop) and RMS_norm is a per-pixel row kernel;
  * the reference's per-convolution feature cache (the last two input frames of every causal conv,
    vae.py:202-220) is not a set of tensors that get cloned and concatenated every call: the producer of a
    conv's input writes into a ring of frames and the temporal taps read the ring slots before the new ones;
  * the single-head attention block runs as GEMM (fp32 logits) -> row softmax -> GEMM on the same tcgen05 GEMM
    as the denoiser.
Every tensor op below is a libllb200 kernel; torch only allocates memory (and copies frames between buffers).
There is no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import Conv3dDesc

LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
               0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
              3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def _pad32(c: int) -> int:
    return (c + 31) // 32 * 32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------
# thin wrappers over the C ABI
# ------------------------------------------------------------------------------------------------
class FrameRing:
    """[frames, H, W, Cp] bf16, zero-initialised; ``reserve(T)`` hands out the slots of the next T frames."""

    def __init__(self, frames: int, H: int, W: int, Cp: int, device):
        self.buf = torch.zeros(frames, H, W, Cp, dtype=torch.bfloat16, device=device)
        self.frames, self.pos = frames, 0

    def reserve(self, T: int) -> int:
        t0 = self.pos
        self.pos = (self.pos + T) % self.frames
        return t0

    def reset(self):
        self.buf.zero_()
        self.pos = 0


def conv3d(inp: torch.Tensor, in_t0: int, weight: torch.Tensor, bias: Optional[torch.Tensor], k,
           out: Optional[torch.Tensor], T: int, *, out_t0: int = 0, out_t_step: int = 1,
           res: Optional[torch.Tensor] = None, res_t0: int = 0, norm: Optional[dict] = None):
    """llb_conv3d on channels-last rings: inp [Fi, H, W, ld_in], out [Fo, H, W, ld_out], weight [Cout, taps*Cin]
    with Cin <= ld_in, Cout <= ld_out multiples of 32 (channels beyond them are padding and stay untouched).
    norm = {"ring": FrameRing, "t0": int, "gamma": [Cout] bf16, "C": real channels, "silu": bool} additionally
    writes RMS_norm(result) (* gamma, SiLU) into that ring; `out` may then be None."""
    d = Conv3dDesc()
    taps = k[0] * k[1] * k[2]
    d.inp, d.in_frames, d.in_t0 = inp.data_ptr(), inp.shape[0], in_t0
    dst = out if out is not None else norm["ring"].buf
    d.H, d.W, d.ld_in, d.ld_out = inp.shape[1], inp.shape[2], inp.shape[3], dst.shape[3]
    d.Cin, d.Cout = weight.shape[1] // taps, weight.shape[0]
    d.weight, d.bias = weight.data_ptr(), (bias.data_ptr() if bias is not None else None)
    d.kt, d.kh, d.kw = k
    d.out, d.out_frames = (out.data_ptr(), out.shape[0]) if out is not None else (None, 0)
    d.out_t0, d.out_t_step = out_t0, out_t_step
    if norm is not None:
        nb = norm["ring"].buf
        assert nb.shape[1:] == dst.shape[1:] and nb.is_contiguous() and norm["gamma"].numel() == weight.shape[0]
        d.norm_out, d.norm_frames, d.norm_t0 = nb.data_ptr(), nb.shape[0], norm["t0"]
        d.norm_gamma, d.norm_channels, d.norm_silu = norm["gamma"].data_ptr(), norm["C"], int(norm["silu"])
    d.res, d.res_frames, d.res_t0 = (res.data_ptr() if res is not None else None), (res.shape[0] if res is not None else 0), res_t0
    d.T = T
    assert inp.is_contiguous() and dst.is_contiguous() and weight.is_contiguous()
    assert weight.shape[1] == taps * d.Cin and d.Cin <= d.ld_in and d.Cout <= d.ld_out, (weight.shape, k, inp.shape, out.shape)
    assert bias is None or bias.numel() == d.Cout
    _lib.check(_lib.lib().llb_conv3d(C.byref(d), _stream()), "llb_conv3d")
    return out


FUSED_NORM_MAX_CHANNELS = 192  # llb_conv3d's fused norm needs a pixel's channels in one tile


def vae_norm(inp: torch.Tensor, in_t0: int, out: torch.Tensor, out_t0: int, T: int, C_real: int, gamma: torch.Tensor,
             silu: bool):
    pixels = inp.shape[1] * inp.shape[2]
    _lib.check(_lib.lib().llb_vae_norm(inp.data_ptr(), inp.shape[0], in_t0, out.data_ptr(), out.shape[0], out_t0, T,
                                       pixels, inp.shape[3], C_real, gamma.data_ptr(), int(silu), _stream()), "llb_vae_norm")
    return out


def upsample2x(inp: torch.Tensor, out: torch.Tensor, T: int):
    _lib.check(_lib.lib().llb_vae_upsample2x(inp.data_ptr(), out.data_ptr(), T, inp.shape[1], inp.shape[2], inp.shape[3],
                                             _stream()), "llb_vae_upsample2x")
    return out


def transpose(inp: torch.Tensor, out: torch.Tensor):
    _lib.check(_lib.lib().llb_transpose_bf16(inp.data_ptr(), inp.stride(0), out.data_ptr(), out.stride(0), inp.shape[0],
                                             inp.shape[1], _stream()), "llb_transpose_bf16")
    return out


def softmax_rows(logits: torch.Tensor, out: torch.Tensor, cols_valid: int, scale: float):
    _lib.check(_lib.lib().llb_softmax_rows(logits.data_ptr(), logits.stride(0), out.data_ptr(), out.stride(0),
                                           logits.shape[0], cols_valid, out.shape[1], scale, _stream()), "llb_softmax_rows")
    return out


# ------------------------------------------------------------------------------------------------
# parameters: reference names, reference shapes
# ------------------------------------------------------------------------------------------------
def _decoder_plan(dim, z_dim, dim_mult, num_res_blocks, temporal_upsample):
    """Same flat step list as Decoder3d.__init__ builds (wan/modules/vae.py:389-421)."""
    dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
    plan = [("conv", "decoder.conv1", z_dim, dims[0]), ("res", "decoder.middle.0", dims[0], dims[0]),
            ("attn", "decoder.middle.1", dims[0]), ("res", "decoder.middle.2", dims[0], dims[0])]
    idx = 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin //= 2
        for _ in range(num_res_blocks + 1):
            plan.append(("res", f"decoder.upsamples.{idx}", cin, cout))
            idx += 1
            cin = cout
        if i != len(dim_mult) - 1:
            plan.append(("up", f"decoder.upsamples.{idx}", cout, bool(temporal_upsample[i])))
            idx += 1
    plan.append(("head", "decoder.head", dims[-1]))
    return plan


class WanVAEDecoder(nn.Module):
    """Decoder half of WanVAE_ (wan/modules/vae.py:483-609) on libllb200.  Parameters keep the reference's
    names and shapes; ``refresh_weights()`` re-packs them for the kernels after loading."""

    def __init__(self, dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_upsample=(True, True, False)):
        super().__init__()
        self.dim, self.z_dim = dim, z_dim
        self.plan = _decoder_plan(dim, z_dim, list(dim_mult), num_res_blocks, temporal_upsample)
        self._names: Dict[str, str] = {}

        def conv(name, cout, cin, *k):
            self._reg(name + ".weight", torch.zeros(cout, cin, *k))
            self._reg(name + ".bias", torch.zeros(cout))

        def gamma(name, c, nd):
            self._reg(name + ".gamma", torch.ones(c, *([1] * nd)))

        conv("conv2", z_dim, z_dim, 1, 1, 1)
        for step in self.plan:
            kind, p = step[0], step[1]
            if kind == "conv":
                conv(p, step[3], step[2], 3, 3, 3)
            elif kind == "res":
                cin, cout = step[2], step[3]
                gamma(p + ".residual.0", cin, 3); conv(p + ".residual.2", cout, cin, 3, 3, 3)
                gamma(p + ".residual.3", cout, 3); conv(p + ".residual.6", cout, cout, 3, 3, 3)
                if cin != cout:
                    conv(p + ".shortcut", cout, cin, 1, 1, 1)
            elif kind == "attn":
                gamma(p + ".norm", step[2], 2); conv(p + ".to_qkv", 3 * step[2], step[2], 1, 1); conv(p + ".proj", step[2], step[2], 1, 1)
            elif kind == "up":
                conv(p + ".resample.1", step[2] // 2, step[2], 3, 3)
                if step[3]:
                    conv(p + ".time_conv", 2 * step[2], step[2], 3, 1, 1)
            else:
                gamma(p + ".0", step[2], 3); conv(p + ".2", 3, step[2], 3, 3, 3)
        self.fuse_norm = True  # RMS_norm + SiLU inside the producing convolution's epilogue where the tile allows
        # Optional: steady-state frames replay one of six captured launch sequences.  Measured neutral (17.9 vs
        # 17.4 ms per latent frame: the decode is GPU-bound, launches already run ahead), so off by default.
        self.use_cuda_graph = False
        self._graphs: Dict[tuple, dict] = {}
        self._static: Dict[tuple, dict] = {}
        self._frames_done = 0
        self._packed: Dict[str, torch.Tensor] = {}
        self._rings: Dict[str, FrameRing] = {}
        self._scratch: Dict[str, torch.Tensor] = {}
        self._up_calls: Dict[str, int] = {}

    # nn.Module cannot hold dotted parameter names directly: keep a flat name map and present the reference's
    # names through state_dict() / load_state_dict().
    def _reg(self, name: str, t: torch.Tensor):
        flat = name.replace(".", "__")
        self._names[name] = flat
        self.register_parameter(flat, nn.Parameter(t, requires_grad=False))

    def state_dict(self, *a, **k):
        sd = super().state_dict(*a, **k)
        back = {v: n for n, v in self._names.items()}
        return {back.get(key, key): val for key, val in sd.items()}

    def load_state_dict(self, sd, strict: bool = True, **k):
        own = {n for n in self._names}
        flat = {self._names[n]: v for n, v in sd.items() if n in own}
        missing = [n for n in own if n not in sd]
        unexpected = [n for n in sd if n not in own and not n.startswith(("encoder.", "conv1."))]
        if strict and (missing or unexpected):
            raise RuntimeError(f"WanVAEDecoder.load_state_dict: missing {missing[:4]} unexpected {unexpected[:4]}")
        res = super().load_state_dict(flat, strict=False)
        self._packed.clear()
        self._graphs.clear(); self._static.clear()  # captured graphs hold pointers to the packed weights
        return res

    def _p(self, name: str) -> torch.Tensor:
        return getattr(self, self._names[name])

    def _apply(self, fn, *a, **k):
        self._packed.clear()  # packed copies follow the parameters' device / dtype
        self._graphs.clear(); self._static.clear(); self._rings.clear(); self._scratch.clear()
        return super()._apply(fn, *a, **k)

    # ---------------------------------------------------------------------------------- weight packing
    def _conv_w(self, name: str):
        """[Cout, Cin, (kt,) kh, kw] -> ([Cout32, taps * Cin32] tap-major bf16, bias [Cout32], (kt, kh, kw)) with the
        channel counts rounded up to multiples of 32 (zero rows / columns)."""
        if name not in self._packed:
            w, b = self._p(name + ".weight"), self._p(name + ".bias")
            if w.dim() == 4:
                w = w.unsqueeze(2)
            cout, cin, kt, kh, kw = w.shape
            wp = torch.zeros(_pad32(cout), kt, kh, kw, _pad32(cin), dtype=torch.bfloat16, device=w.device)
            wp[:cout, :, :, :, :cin] = w.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
            bp = torch.zeros(_pad32(cout), dtype=torch.bfloat16, device=w.device)
            bp[:cout] = b.to(torch.bfloat16)
            self._packed[name] = (wp.reshape(_pad32(cout), -1).contiguous(), bp, (kt, kh, kw))
        return self._packed[name]

    def _gamma(self, name: str):
        if name not in self._packed:
            g = self._p(name + ".gamma").reshape(-1)
            gp = torch.zeros(_pad64(g.numel()), dtype=torch.bfloat16, device=g.device)
            gp[:g.numel()] = g.to(torch.bfloat16)
            self._packed[name] = gp
        return self._packed[name]

    def _gamma32(self, name: str):
        """gamma padded to the conv kernels' 32-channel granularity (fused-norm argument)."""
        key = name + "#32"
        if key not in self._packed:
            g = self._p(name + ".gamma").reshape(-1)
            gp = torch.zeros(_pad32(g.numel()), dtype=torch.bfloat16, device=g.device)
            gp[:g.numel()] = g.to(torch.bfloat16)
            self._packed[key] = gp
        return self._packed[key]

    def _attn_w(self, p: str, c: int):
        """to_qkv as one [3*Cp, Cp] GEMM weight with q | k | v row blocks each padded to Cp, and proj [Cp, Cp]."""
        key = p + "#attn"
        if key not in self._packed:
            cp = _pad64(c)
            wq, bq = self._p(p + ".to_qkv.weight").reshape(3 * c, c), self._p(p + ".to_qkv.bias")
            w = torch.zeros(3 * cp, cp, dtype=torch.bfloat16, device=wq.device)
            b = torch.zeros(3 * cp, dtype=torch.bfloat16, device=wq.device)
            for i in range(3):
                w[i * cp:i * cp + c, :c] = wq[i * c:(i + 1) * c].to(torch.bfloat16)
                b[i * cp:i * cp + c] = bq[i * c:(i + 1) * c].to(torch.bfloat16)
            wo = torch.zeros(cp, cp, dtype=torch.bfloat16, device=wq.device)
            wo[:c, :c] = self._p(p + ".proj.weight").reshape(c, c).to(torch.bfloat16)
            bo = torch.zeros(cp, dtype=torch.bfloat16, device=wq.device)
            bo[:c] = self._p(p + ".proj.bias").to(torch.bfloat16)
            self._packed[key] = (w, b, wo, bo)
        return self._packed[key]

    def refresh_weights(self):
        self._packed.clear()
        self._graphs.clear(); self._static.clear()

    # ---------------------------------------------------------------------------------- buffers
    def _ring(self, key: str, frames: int, H: int, W: int, Cp: int, device) -> FrameRing:
        r = self._rings.get(key)
        if r is None or r.buf.shape != (frames, H, W, Cp) or r.buf.device != torch.device(device):
            if r is not None:
                # a captured graph bakes in the old buffer's address: decoding at another resolution drops them
                self._graphs.clear()
            r = FrameRing(frames, H, W, Cp, device)
            self._rings[key] = r
        return r

    def _buf(self, key: str, shape, device, dtype=torch.bfloat16) -> torch.Tensor:
        """Persistent scratch, zero-filled at creation: kernels only ever write the real channels / rows of a
        buffer, so its padding stays zero for the life of the decoder."""
        b = self._scratch.get(key)
        if b is None or tuple(b.shape) != tuple(shape) or b.dtype != dtype or b.device != torch.device(device):
            if b is not None:
                self._graphs.clear()  # same reason as in _ring
            b = torch.zeros(shape, dtype=dtype, device=device)
            self._scratch[key] = b
        return b

    def clear_cache(self):
        """WanVAE_.clear_cache (vae.py:602-609): forget the stream (all conv histories back to zero).  Buffers
        and captured graphs stay: a new stream walks through the same ring positions."""
        for r in self._rings.values():
            r.reset()
        self._up_calls.clear()
        self._frames_done = 0

    # ---------------------------------------------------------------------------------- the decoder
    def _causal_conv(self, name: str, ring: FrameRing, t0: int, T: int, out, res=None, norm=None):
        w, b, k = self._conv_w(name)
        return conv3d(ring.buf, t0, w, b, k, out, T, res=res, norm=norm)

    def _norm_target(self, nxt, H: int, W: int, tmax: int, T: int, c: int, dev) -> Optional[dict]:
        """If the step after the current one starts with RMS_norm + SiLU of the current result (a residual
        block or the head) and the channel count fits, reserve the slots of its input ring so that the current
        step's last convolution writes the normalised frames there itself."""
        if nxt is None or nxt[0] not in ("res", "head") or _pad32(c) > FUSED_NORM_MAX_CHANNELS or not self.fuse_norm:
            return None
        if nxt[0] == "res":
            ring, gname = self._ring(nxt[1] + "#a", tmax + 2, H, W, _pad64(c), dev), nxt[1] + ".residual.0"
        else:
            ring, gname = self._ring(nxt[1] + "#h", tmax + 2, H, W, _pad64(c), dev), nxt[1] + ".0"
        return {"ring": ring, "t0": ring.reserve(T), "gamma": self._gamma32(gname), "C": c, "silu": True}

    def _res_block(self, p: str, x: torch.Tensor, T: int, cin: int, cout: int, tmax: int, pre=None, nxt=None) -> torch.Tensor:
        """ResidualBlock.forward (vae.py:202-220): x [T, H, W, Cin_p] -> [T, H, W, Cout_p].  `pre`: the previous
        convolution already left SiLU(RMS_norm(x)) in this block's input ring; `nxt`: where to leave
        SiLU(RMS_norm(result)) for the next step."""
        dev = x.device
        _, H, W, cinp = x.shape
        coutp = _pad64(cout)
        ra = self._ring(p + "#a", tmax + 2, H, W, cinp, dev)
        rb = self._ring(p + "#b", tmax + 2, H, W, coutp, dev)
        if pre is None:
            ta = ra.reserve(T)
            vae_norm(x, 0, ra.buf, ta, T, cin, self._gamma(p + ".residual.0"), True)
        else:
            assert pre["ring"] is ra
            ta = pre["t0"]
        tb = rb.reserve(T)
        if _pad32(cout) <= FUSED_NORM_MAX_CHANNELS and self.fuse_norm:
            # conv1 -> RMS_norm -> SiLU in one kernel; the un-normalised conv1 output has no other reader
            self._causal_conv(p + ".residual.2", ra, ta, T, None,
                              norm={"ring": rb, "t0": tb, "gamma": self._gamma32(p + ".residual.3"), "C": cout, "silu": True})
        else:
            y = self._buf(f"y{H}x{W}x{coutp}c{cout}", (tmax, H, W, coutp), dev)
            self._causal_conv(p + ".residual.2", ra, ta, T, y)
            vae_norm(y, 0, rb.buf, tb, T, cout, self._gamma(p + ".residual.3"), True)
        if cin != cout:
            w, b, k = self._conv_w(p + ".shortcut")
            h = self._buf(f"h{H}x{W}x{coutp}c{cout}", (tmax, H, W, coutp), dev)
            conv3d(x, 0, w, b, k, h, T)
        else:
            h = x
        # x may be overwritten in place when the shapes agree: each output element reads its own residual first
        out = x if cin == cout else self._buf(f"x{H}x{W}x{coutp}#{p}", (tmax, H, W, coutp), dev)
        self._causal_conv(p + ".residual.6", rb, tb, T, out, res=h, norm=nxt)
        return out

    def _attn_block(self, p: str, x: torch.Tensor, T: int, c: int) -> torch.Tensor:
        """AttentionBlock.forward (vae.py:240-262), one frame at a time."""
        dev = x.device
        _, H, W, cp = x.shape
        n = H * W
        npad = (n + 7) // 8 * 8
        wqkv, bqkv, wo, bo = self._attn_w(p, c)
        xn = self._buf(f"an{n}x{cp}", (1, H, W, cp), dev)
        qkv = self._buf(f"aqkv{npad}x{cp}", (npad, 3 * cp), dev)   # rows >= n stay zero
        logits = self._buf(f"al{n}x{npad}", (n, npad), dev, torch.float32)
        prob = self._buf(f"ap{n}x{npad}", (n, npad), dev)
        vt = self._buf(f"avt{cp}x{npad}", (cp, npad), dev)
        o = self._buf(f"ao{n}x{cp}", (n, cp), dev)
        for t in range(T):
            xt = x[t:t + 1]
            vae_norm(xt, 0, xn, 0, 1, c, self._gamma(p + ".norm"), False)
            ops.gemm(xn.view(n, cp), wqkv, bqkv, out=qkv[:n])
            ops.gemm(qkv[:n, :cp], qkv[:, cp:2 * cp], None, epilogue=ops.EPI_BIAS_F32, out=logits)
            softmax_rows(logits, prob, n, float(c) ** -0.5)
            transpose(qkv[:, 2 * cp:], vt)
            ops.gemm(prob, vt, None, out=o)
            xv = xt.view(n, cp)
            ops.gemm(o, wo, bo, epilogue=ops.EPI_BIAS_RES, res=xv, out=xv)
        return x

    def _upsample(self, p: str, x: torch.Tensor, T: int, c: int, temporal: bool, tmax_in: int, nxt_step=None):
        """Resample.forward, upsample2d / upsample3d (vae.py:101-138) -> (x', T', fused-norm target or None)."""
        dev = x.device
        _, H, W, cp = x.shape
        tmax_out = 2 * tmax_in if temporal else tmax_in
        if temporal:
            n = self._up_calls.get(p, 0)
            self._up_calls[p] = n + 1
            if n > 0:  # the stream's first call passes through untouched ('Rep')
                ring = self._ring(p + "#t", tmax_in + 2, H, W, cp, dev)
                t0 = ring.reserve(T)
                for i in range(T):  # plain device copies: x is also the residual stream, the ring is the conv's input
                    ring.buf[(t0 + i) % ring.frames].copy_(x[i])
                w, b, k = self._conv_w(p + ".time_conv")       # [2c, 3c]: rows [0, c) -> even frames, [c, 2c) -> odd
                y = self._buf(f"tc{H}x{W}x{cp}", (tmax_out, H, W, cp), dev)
                conv3d(ring.buf, t0, w[:c], b[:c], k, y, T, out_t0=0, out_t_step=2)
                conv3d(ring.buf, t0, w[c:], b[c:], k, y, T, out_t0=1, out_t_step=2)
                x, T = y, 2 * T
        up = self._buf(f"up{2 * H}x{2 * W}x{cp}", (tmax_out, 2 * H, 2 * W, cp), dev)
        upsample2x(x, up, T)
        w, b, k = self._conv_w(p + ".resample.1")
        out = self._buf(f"x{2 * H}x{2 * W}#{p}", (tmax_out, 2 * H, 2 * W, _pad64(c // 2)), dev)
        nxt = self._norm_target(nxt_step, 2 * H, 2 * W, tmax_out, T, c // 2, dev)
        conv3d(up, 0, w, b, k, out, T, norm=nxt)
        return out, T, nxt

    def decode_one(self, ring0: FrameRing, t0: int, out_pixels: torch.Tensor) -> int:
        """Decoder3d.forward for ONE latent frame that llb_vae_latent_in has placed at ring0[t0].
        Writes [T', 3, 8h, 8w] float32 into out_pixels and returns T' (1 for the stream's first frame, else 4)."""
        dev = ring0.buf.device
        _, H, W, _ = ring0.buf.shape
        T, tmax = 1, 1
        x = None
        pre = None  # fused-norm hand-over from the previous step's last convolution to this step
        for i, step in enumerate(self.plan):
            kind, p = step[0], step[1]
            nxt_step = self.plan[i + 1] if i + 1 < len(self.plan) else None
            if kind == "conv":
                x = self._buf(f"x{H}x{W}x{_pad64(step[3])}#{p}", (tmax, H, W, _pad64(step[3])), dev)
                nxt = self._norm_target(nxt_step, H, W, tmax, T, step[3], dev)
                self._causal_conv(p, ring0, t0, T, x, norm=nxt)
                pre = nxt
            elif kind == "res":
                nxt = self._norm_target(nxt_step, H, W, tmax, T, step[3], dev)
                x = self._res_block(p, x, T, step[2], step[3], tmax, pre=pre, nxt=nxt)
                pre = nxt
            elif kind == "attn":
                x = self._attn_block(p, x, T, step[2])
                pre = None
            elif kind == "up":
                if step[3] and step[2] % 32 != 0:
                    raise RuntimeError("temporal upsampling needs a channel count that is a multiple of 32")
                x, T, pre = self._upsample(p, x, T, step[2], step[3], tmax, nxt_step)
                if step[3]:
                    tmax *= 2
                H, W = 2 * H, 2 * W
            else:
                cp = x.shape[3]
                rh = self._ring(p + "#h", tmax + 2, H, W, cp, dev)
                if pre is None:
                    th = rh.reserve(T)
                    vae_norm(x, 0, rh.buf, th, T, step[2], self._gamma(p + ".0"), True)
                else:
                    assert pre["ring"] is rh
                    th = pre["t0"]
                y = self._buf(f"head{H}x{W}", (tmax, H, W, 64), dev)
                self._causal_conv(p + ".2", rh, th, T, y)
                _lib.check(_lib.lib().llb_vae_pixel_out(y.data_ptr(), out_pixels.data_ptr(), T, H * W, 64, _stream()),
                           "llb_vae_pixel_out")
        return T

    @torch.no_grad()
    def cached_decode(self, z: torch.Tensor, scale) -> torch.Tensor:
        """WanVAE_.cached_decode (vae.py:571-593): z [1, z_dim, T, h, w] bf16 -> [1, 3, T', 8h, 8w] float32 in
        [-1, 1] (the wrapper's .float().clamp_ is folded into the last kernel); caches persist across calls."""
        if not z.is_cuda:
            raise RuntimeError("WanVAEDecoder needs CUDA tensors (no CPU fallback)")
        assert z.dim() == 5 and z.shape[0] == 1 and z.shape[1] == self.z_dim, z.shape
        dev = z.device
        _, zc, Tl, h, w = z.shape
        zb = z.to(torch.bfloat16)
        # launch arguments must be pointer-stable across calls (CUDA-graph replay): static staging buffers
        st = self._static.get((h, w, str(dev)))
        if st is None:
            st = {"z": torch.zeros(zc, 1, h, w, dtype=torch.bfloat16, device=dev),
                  "px": torch.zeros(4, 3, 8 * h, 8 * w, dtype=torch.float32, device=dev),
                  "mean": torch.zeros(zc, dtype=torch.bfloat16, device=dev),
                  "inv_std": torch.zeros(zc, dtype=torch.bfloat16, device=dev),
                  "w2": self._p("conv2.weight").reshape(zc, zc).to(dev, torch.bfloat16).contiguous(),
                  "b2": self._p("conv2.bias").to(dev, torch.bfloat16).contiguous()}
            self._static[(h, w, str(dev))] = st
        st["mean"].copy_(scale[0].to(dev, torch.bfloat16))
        st["inv_std"].copy_(scale[1].to(dev, torch.bfloat16))
        ring0 = self._ring("latent", 3, h, w, _pad64(zc), dev)

        def one_frame():
            t0 = ring0.reserve(1)
            _lib.check(_lib.lib().llb_vae_latent_in(st["z"].data_ptr(), st["mean"].data_ptr(), st["inv_std"].data_ptr(),
                                                    st["w2"].data_ptr(), st["b2"].data_ptr(), ring0.buf.data_ptr(),
                                                    ring0.frames, t0, 1, zc, h * w, ring0.buf.shape[3], _stream()),
                       "llb_vae_latent_in")
            return self.decode_one(ring0, t0, st["px"])

        outs: List[torch.Tensor] = []
        for i in range(Tl):
            st["z"].copy_(zb[0, :, i:i + 1])
            n = self._frames_done
            self._frames_done += 1
            # Frame 0 of a stream is structurally different (no temporal upsampling) and frame 1 is the first to
            # touch every buffer: both run eagerly.  From then on the launch sequence of a frame depends only on
            # the ring positions, which repeat with period 6 (rings of 3 / 4 / 6 frames advancing by 1 / 2 / 4).
            if not self.use_cuda_graph or n < 2:
                Tn = one_frame()
            else:
                key = ((n - 1) % 6, h, w, str(dev))
                ent = self._graphs.get(key)
                if ent is None:
                    before = {k: r.pos for k, r in self._rings.items()}
                    calls = dict(self._up_calls)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        Tn = one_frame()
                    ent = {"graph": g, "T": Tn, "before": before, "after": {k: r.pos for k, r in self._rings.items()},
                           "n_rings": len(self._rings)}
                    self._graphs[key] = ent
                    self._up_calls = calls
                else:
                    assert len(self._rings) == ent["n_rings"] and all(self._rings[k].pos == v for k, v in ent["before"].items()), \
                        "VAE ring positions diverged from the captured schedule"
                for k, v in ent["after"].items():
                    self._rings[k].pos = v
                for k in [k for k in self.plan if k[0] == "up" and k[3]]:
                    self._up_calls[k[1]] = self._up_calls.get(k[1], 0) + 1
                ent["graph"].replay()
                Tn = ent["T"]
            outs.append(st["px"][:Tn].clone())
        return torch.cat(outs, 0).permute(1, 0, 2, 3).unsqueeze(0)

    def decode(self, z: torch.Tensor, scale) -> torch.Tensor:
        """WanVAE_.decode (vae.py:545-569): fresh cache before, cleared after."""
        self.clear_cache()
        out = self.cached_decode(z, scale)
        self.clear_cache()
        return out


class WanVAEWrapper(nn.Module):
    """Decode half of the reference's WanVAEWrapper (utils/wan_wrapper.py:60-117)."""

    def __init__(self, decoder: Optional[WanVAEDecoder] = None):
        super().__init__()
        self.model = decoder if decoder is not None else WanVAEDecoder()
        self.mean = torch.tensor(LATENT_MEAN, dtype=torch.float32)
        self.std = torch.tensor(LATENT_STD, dtype=torch.float32)

    def decode_to_pixel(self, latent: torch.Tensor, use_cache: bool = False) -> torch.Tensor:
        """latent [B, T, 16, h, w] -> video [B, T', 3, 8h, 8w] float32 in [-1, 1]."""
        zs = latent.permute(0, 2, 1, 3, 4)
        if use_cache:
            assert latent.shape[0] == 1, "Batch size must be 1 when using cache"
        device, dtype = latent.device, latent.dtype
        zd = self.model.z_dim
        scale = [self.mean[:zd].to(device=device, dtype=dtype), 1.0 / self.std[:zd].to(device=device, dtype=dtype)]
        fn = self.model.cached_decode if use_cache else self.model.decode
        out = [fn(u.unsqueeze(0), scale).squeeze(0) for u in zs]
        return torch.stack(out, dim=0).permute(0, 2, 1, 3, 4)

"""Streaming VAE decoder on libllb200 (longlive_b200/vae.py) against the oracle restatement of the reference
decoder (oracle/vae_oracle.py, itself pinned bit-for-bit to the reference module) and against the committed
outputs of the reference module (tests/golden/vae_small.pt)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "vae_small.pt")


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _cl(x):  # [C, T, H, W] -> channels-last [T, H, W, C]
    return x.permute(1, 2, 3, 0).contiguous()


@pytest.mark.parametrize("cin,cout,k,H,W", [(64, 64, (3, 3, 3), 5, 7), (128, 192, (3, 3, 3), 24, 40), (64, 128, (1, 3, 3), 17, 33),
                                            (192, 64, (3, 1, 1), 9, 16), (64, 64, (1, 1, 1), 8, 16), (384, 384, (3, 3, 3), 12, 20),
                                            (96, 96, (3, 3, 3), 20, 36), (32, 96, (3, 3, 3), 9, 17), (192, 96, (1, 3, 3), 16, 32),
                                            (96, 32, (3, 3, 3), 11, 19)])
@pytest.mark.parametrize("mt", ["1", "2"])
def test_conv3d_matches_torch_with_ring_history(monkeypatch, mt, cin, cout, k, H, W):
    """Two consecutive calls (T = 2 then T = 3) on a 5-frame ring: the second call's temporal taps must see the
    first call's last frames through the ring (wrap-around included), exactly like CausalConv3d with its cache."""
    from longlive_b200 import vae
    monkeypatch.setenv("LLB_CONV_MT", mt)   # patches per CTA tile (2 is only honoured for <= 128 output channels)
    g = torch.Generator().manual_seed(cin + cout + H)
    kt, kh, kw = k
    w = (torch.randn(cout, cin, kt, kh, kw, generator=g) / (cin * kt * kh * kw) ** 0.5).to(torch.bfloat16)
    b = (0.1 * torch.randn(cout, generator=g)).to(torch.bfloat16)
    wp = w.permute(0, 2, 3, 4, 1).reshape(cout, -1).contiguous().to(DEV)
    frames = [torch.randn(cin, 1, H, W, generator=g).to(torch.bfloat16) for _ in range(5)]
    cinp, coutp = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64   # channel strides (padding must stay untouched)
    ring = vae.FrameRing(5, H, W, cinp, DEV)
    stream = torch.zeros(cin, 2, H, W, dtype=torch.bfloat16)  # two zero frames = causal padding
    done = 0
    for T in (2, 3):
        new = torch.cat(frames[done:done + T], 1)
        t0 = ring.reserve(T)
        for i in range(T):
            ring.buf[(t0 + i) % 5][..., :cin].copy_(_cl(new[:, i:i + 1])[0])
        res = torch.randn(T, H, W, coutp, generator=g).to(torch.bfloat16).to(DEV)
        out = torch.full((T, H, W, coutp), 3.0, dtype=torch.bfloat16, device=DEV)
        vae.conv3d(ring.buf, t0, wp, b.to(DEV), k, out, T, res=res)
        assert (out[..., cout:] == 3.0).all()
        out, res = out[..., :cout], res[..., :cout]
        stream = torch.cat([stream, new], 1)
        xin = stream[:, -(T + kt - 1):] if kt == 3 else new
        ref = F.conv3d(F.pad(xin.float().unsqueeze(0), (kw // 2, kw // 2, kh // 2, kh // 2, 0, 0)), w.float(), b.float())[0]
        ref = _cl(ref.to(torch.bfloat16)).float() + res.cpu().float()
        err = rel_l2(out.cpu(), ref)
        assert err < 4e-3, f"T={T}: rel-L2 {err}"
        done += T


@pytest.mark.parametrize("C,silu", [(16, True), (96, True), (192, False), (384, True)])
def test_vae_norm_matches_reference_op_chain(C, silu):
    from longlive_b200 import vae
    from oracle import vae_oracle as vo
    g = torch.Generator().manual_seed(C)
    Cp = (C + 63) // 64 * 64
    x = torch.randn(1, C, 2, 9, 11, generator=g).to(torch.bfloat16)
    gamma = (1 + 0.1 * torch.randn(C, 1, 1, 1, generator=g)).to(torch.bfloat16)
    ref = vo.rms_norm(x, gamma)
    if silu:
        ref = F.silu(ref)
    xin = torch.zeros(2, 9, 11, Cp, dtype=torch.bfloat16)
    xin[..., :C] = _cl(x[0])
    gp = torch.zeros(Cp, dtype=torch.bfloat16); gp[:C] = gamma.reshape(-1)
    out = torch.empty(2, 9, 11, Cp, dtype=torch.bfloat16, device=DEV)
    vae.vae_norm(xin.to(DEV), 0, out, 0, 2, C, gp.to(DEV), silu)
    got = out.cpu()
    assert Cp == C or got[..., C:].abs().max().item() == 0
    diff = (got[..., :C].float() - _cl(ref[0]).float()).abs()
    # same rounding chain: identical up to 1 bf16 ulp where the fp32 norm / exp differ in the last bit
    assert (diff <= _cl(ref[0]).float().abs() * 2 ** -7 + 1e-6).all()
    assert (diff == 0).float().mean().item() > 0.98


def _decoder(cfg_kwargs, sd):
    from longlive_b200.vae import WanVAEDecoder
    dec = WanVAEDecoder(dim=cfg_kwargs["dim"], z_dim=cfg_kwargs["z_dim"], dim_mult=cfg_kwargs["dim_mult"],
                        num_res_blocks=cfg_kwargs["num_res_blocks"], temporal_upsample=cfg_kwargs["temporal_upsample"])
    dec.load_state_dict(sd)
    return dec.to(DEV)


def test_small_decoder_streaming_vs_reference_golden_and_oracle():
    from oracle import vae_oracle as vo
    from oracle.make_vae_golden import latents, scale_of
    gold = torch.load(GOLDEN)
    cfg = vo.VaeConfig(**gold["cfg"])
    sd = vo.init_state_dict(cfg, seed=gold["seed"], dtype=torch.bfloat16)
    dec = _decoder(gold["cfg"], sd)
    scale = [s.to(DEV) for s in scale_of(cfg, torch.bfloat16)]
    for i, t in enumerate(gold["chunks"]):
        z = latents(cfg, 10 + i, t).to(torch.bfloat16).to(DEV)
        out = dec.cached_decode(z, scale).cpu()
        ref = gold["bf16"]["stream"][i].float().clamp(-1, 1)        # the reference module's own output
        exact = gold["f32"]["stream"][i].float().clamp(-1, 1)
        assert out.shape == ref.shape
        e_ref, e_exact, floor = rel_l2(out, ref), rel_l2(out, exact), rel_l2(ref, exact)
        print(f"call {i}: vs reference bf16 {e_ref:.3e}, vs fp32 {e_exact:.3e} (reference bf16 vs fp32 {floor:.3e})")
        assert e_ref < 2e-2 and e_exact < 1.5 * floor + 5e-3
    whole = dec.decode(latents(cfg, 99, 4).to(torch.bfloat16).to(DEV), scale).cpu()
    assert rel_l2(whole, gold["bf16"]["whole"].float().clamp(-1, 1)) < 2e-2


def test_decode_to_pixel_wrapper_and_cache_reset():
    from oracle import vae_oracle as vo
    from oracle.make_vae_golden import SMALL, latents
    from longlive_b200.vae import WanVAEWrapper
    cfg = vo.VaeConfig(**SMALL)
    sd = vo.init_state_dict(cfg, seed=3, dtype=torch.bfloat16)
    wrap = WanVAEWrapper(_decoder(SMALL, sd))
    oracle = vo.VaeDecoderOracle(cfg, sd)
    lat = latents(cfg, 21, 3).permute(0, 2, 1, 3, 4).to(torch.bfloat16)          # [B, T, z, h, w]
    a = wrap.decode_to_pixel(lat.to(DEV), use_cache=False)
    b = wrap.decode_to_pixel(lat.to(DEV), use_cache=False)                      # decode() must leave no state behind
    assert torch.equal(a, b)
    with torch.no_grad():
        ref = oracle.decode_to_pixel(lat, use_cache=False)
    assert a.shape == ref.shape == (1, 9, 3, 40, 56) and a.dtype == torch.float32
    assert rel_l2(a.cpu(), ref) < 2e-2
    # streaming: two cached calls == one call over the concatenation
    wrap.model.clear_cache()
    s1 = wrap.decode_to_pixel(lat[:, :1].to(DEV), use_cache=True)
    s2 = wrap.decode_to_pixel(lat[:, 1:].to(DEV), use_cache=True)
    assert torch.equal(torch.cat([s1, s2], 1), a)


def test_pipeline_decodes_with_native_vae_full_resolution():
    """The reference seam: CausalInferencePipeline(args, device, generator=, text_encoder=, vae=) with the
    libllb200 VAE wrapper as `vae` - 6 latent frames at 60 x 104 -> 21 frames of 480 x 832 video in [0, 1],
    compared with the oracle decoding the same latents with the same (small-width) decoder weights."""
    from oracle import vae_oracle as vo
    from oracle import wan_oracle as wo
    from oracle.make_golden import PIPE_CFG
    from longlive_b200.pipeline import CausalInferencePipeline
    from longlive_b200.vae import WanVAEWrapper
    from longlive_b200.wrapper import WanDiffusionWrapper
    from tests.test_model_gpu import _model_from, _pipe_args
    cfg = wo.WanConfig(**PIPE_CFG)
    gen = WanDiffusionWrapper(model=_model_from(cfg, wo.init_state_dict(cfg, seed=0)), timestep_shift=5.0)
    vcfg = vo.VaeConfig(dim=32, z_dim=16)
    vsd = vo.init_state_dict(vcfg, seed=4, dtype=torch.bfloat16)
    vae = WanVAEWrapper(_decoder(dict(dim=32, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2,
                                      temporal_upsample=(True, True, False)), vsd))
    ctx = wo.synth_prompt_embeds(cfg, 200, 77).to(DEV)
    pipe = CausalInferencePipeline(_pipe_args(cfg), torch.device(DEV), generator=gen,
                                   text_encoder=lambda text_prompts: {"prompt_embeds": ctx}, vae=vae)
    noise = torch.randn(1, 6, 16, 60, 104, generator=torch.Generator().manual_seed(0)).to(torch.bfloat16).to(DEV)
    video, lat = pipe.inference(noise, text_prompts=["a"], return_latents=True)
    assert video.shape == (1, 21, 3, 480, 832) and video.dtype == torch.float32
    assert float(video.min()) >= 0.0 and float(video.max()) <= 1.0
    oracle = vo.VaeDecoderOracle(vcfg, vsd).to(DEV)
    with torch.no_grad():
        ref = (oracle.decode_to_pixel(lat, use_cache=False) * 0.5 + 0.5).clamp(0, 1)
    err = rel_l2(video, ref)
    print(f"pipeline video vs oracle decode of the same latents: rel-L2 {err:.3e}")
    assert err < 2e-2


@pytest.mark.parametrize("cin,cout,silu", [(96, 96, True), (192, 192, True), (64, 32, False), (128, 128, True)])
@pytest.mark.parametrize("mt", ["1", "2"])
def test_conv3d_fused_norm_matches_separate_norm(monkeypatch, mt, cin, cout, silu):
    """llb_conv3d with the RMS_norm (+SiLU) of its result fused into the epilogue == llb_conv3d followed by
    llb_vae_norm, and both == torch (conv -> + residual -> normalize * sqrt(C) * gamma -> SiLU)."""
    from longlive_b200 import vae
    from oracle import vae_oracle as vo
    monkeypatch.setenv("LLB_CONV_MT", mt)
    g = torch.Generator().manual_seed(cin * 3 + cout)
    H, W, T = 37, 21, 2
    cinp, coutp = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
    w = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5).to(torch.bfloat16)
    b = (0.1 * torch.randn(cout, generator=g)).to(torch.bfloat16)
    gamma = (1 + 0.1 * torch.randn(cout, generator=g)).to(torch.bfloat16)
    x = torch.randn(cin, T, H, W, generator=g).to(torch.bfloat16)
    res = torch.randn(T, H, W, coutp, generator=g).to(torch.bfloat16).to(DEV)
    ring = vae.FrameRing(T + 2, H, W, cinp, DEV)
    t0 = ring.reserve(T)
    for i in range(T):
        ring.buf[t0 + i][..., :cin].copy_(_cl(x[:, i:i + 1])[0])
    wp = w.permute(0, 2, 3, 4, 1).reshape(cout, -1).contiguous().to(DEV)
    # separate
    out_a = torch.zeros(T, H, W, coutp, dtype=torch.bfloat16, device=DEV)
    vae.conv3d(ring.buf, t0, wp, b.to(DEV), (3, 3, 3), out_a, T, res=res)
    gp = torch.zeros(coutp, dtype=torch.bfloat16); gp[:cout] = gamma
    n_a = torch.zeros(T, H, W, coutp, dtype=torch.bfloat16, device=DEV)
    vae.vae_norm(out_a, 0, n_a, 0, T, cout, gp.to(DEV), silu)
    # fused, into ring slots 1.. of a 4-frame ring, raw output optional
    nring = vae.FrameRing(T + 2, H, W, coutp, DEV)
    nring.reserve(1)
    tn = nring.reserve(T)
    out_b = torch.zeros_like(out_a)
    vae.conv3d(ring.buf, t0, wp, b.to(DEV), (3, 3, 3), out_b, T, res=res,
               norm={"ring": nring, "t0": tn, "gamma": gamma.to(DEV), "C": cout, "silu": silu})
    assert torch.equal(out_a, out_b)
    n_b = torch.stack([nring.buf[(tn + i) % nring.frames] for i in range(T)])
    assert coutp == cout or n_b[..., cout:].abs().max().item() == 0
    mism = (n_a != n_b).float().mean().item()
    assert mism < 2e-3, f"fused vs separate norm differ in {mism:.2%} of the elements"
    assert rel_l2(n_b, n_a) < 1e-3
    # vs torch
    y = F.conv3d(F.pad(torch.cat([torch.zeros(cin, 2, H, W), x.float()], 1).unsqueeze(0), (1, 1, 1, 1, 0, 0)), w.float(), b.float())
    y = (y.to(torch.bfloat16) + res.cpu()[..., :cout].permute(3, 0, 1, 2).unsqueeze(0)).to(torch.bfloat16)
    ref = vo.rms_norm(y, gamma.view(-1, 1, 1, 1))
    ref = F.silu(ref) if silu else ref
    assert rel_l2(n_b.cpu()[..., :cout], _cl(ref[0])) < 6e-3
    # raw output not requested at all
    nring2 = vae.FrameRing(T + 2, H, W, coutp, DEV)
    vae.conv3d(ring.buf, t0, wp, b.to(DEV), (3, 3, 3), None, T, res=None,
               norm={"ring": nring2, "t0": 0, "gamma": gamma.to(DEV), "C": cout, "silu": silu})
    assert nring2.buf[:T].abs().sum().item() > 0


def test_decoder_cuda_graph_replay_matches_eager_launches():
    from oracle import vae_oracle as vo
    from oracle.make_vae_golden import SMALL, latents, scale_of
    cfg = vo.VaeConfig(**SMALL)
    sd = vo.init_state_dict(cfg, seed=6, dtype=torch.bfloat16)
    scale = [s.to(DEV) for s in scale_of(cfg, torch.bfloat16)]
    z = latents(cfg, 50, 16).to(torch.bfloat16).to(DEV)          # 16 frames: every graph phase replayed at least once
    outs = []
    for graph in (False, True):
        dec = _decoder(SMALL, sd)
        dec.use_cuda_graph = graph
        a = dec.cached_decode(z[:, :, :9], scale)
        b = dec.cached_decode(z[:, :, 9:], scale)
        dec.clear_cache()
        c = dec.cached_decode(z, scale)                          # second stream reuses the captured graphs
        assert torch.equal(torch.cat([a, b], 2), c)
        outs.append(c)
    assert torch.equal(outs[0], outs[1])


def test_decoder_fused_and_unfused_norm_agree():
    from oracle import vae_oracle as vo
    from oracle.make_vae_golden import SMALL, latents, scale_of
    cfg = vo.VaeConfig(**SMALL)
    sd = vo.init_state_dict(cfg, seed=5, dtype=torch.bfloat16)
    scale = [s.to(DEV) for s in scale_of(cfg, torch.bfloat16)]
    outs = []
    for fuse in (True, False):
        dec = _decoder(SMALL, sd)
        dec.fuse_norm = fuse
        outs.append(torch.cat([dec.cached_decode(latents(cfg, 40 + i, t).to(torch.bfloat16).to(DEV), scale) for i, t in enumerate((1, 2, 2))], 2))
    assert rel_l2(outs[0], outs[1]) < 5e-3  # rare 1-ulp differences in the norm, amplified by the following layers; the bf16 floor is 1.7e-2


def test_full_size_decoder_first_frames_vs_oracle():
    """Wan2.1 VAE decoder shape (dim 96, 384/192/96 channels) at 60 x 104 latents = 832 x 480 video: the stream's
    first latent frame (1 video frame) and the next two (8 video frames) against the oracle on the same GPU."""
    from oracle import vae_oracle as vo
    cfg = vo.VaeConfig()
    sd = vo.init_state_dict(cfg, seed=0, dtype=torch.bfloat16)
    dec = _decoder(dict(dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_upsample=(True, True, False)), sd)
    oracle = vo.VaeDecoderOracle(cfg, sd).to(DEV)
    scale = [torch.tensor(vo.LATENT_MEAN).to(torch.bfloat16).to(DEV), (1.0 / torch.tensor(vo.LATENT_STD)).to(torch.bfloat16).to(DEV)]
    z = torch.randn(1, 16, 3, 60, 104, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(DEV)
    with torch.no_grad():
        for sl in (slice(0, 1), slice(1, 3)):
            a = dec.cached_decode(z[:, :, sl], scale)
            b = oracle.cached_decode(z[:, :, sl], scale).float().clamp_(-1, 1)
            assert a.shape == b.shape and a.shape[-2:] == (480, 832)
            err = rel_l2(a, b)
            print(f"full-size VAE frames {sl}: rel-L2 vs oracle (torch / cuDNN bf16) {err:.3e}")
            assert err < 3e-2   # two bf16 implementations of a 31-convolution stack; the reference's bf16-vs-fp32 is 1.7e-2


def test_long_stream_has_no_ring_drift():
    """40 latent frames streamed in uneven calls: every ring wraps many times; the error against the oracle
    (same stream, same weights, on the GPU through torch ops) must stay at the two-bf16-implementations floor."""
    from oracle import vae_oracle as vo
    from oracle.make_vae_golden import SMALL, latents, scale_of
    cfg = vo.VaeConfig(**SMALL)
    sd = vo.init_state_dict(cfg, seed=8, dtype=torch.bfloat16)
    dec = _decoder(SMALL, sd)
    oracle = vo.VaeDecoderOracle(cfg, sd).to(DEV)
    scale = [s.to(DEV) for s in scale_of(cfg, torch.bfloat16)]
    z = latents(cfg, 60, 40).to(torch.bfloat16).to(DEV)
    errs, pos = [], 0
    with torch.no_grad():
        for n in (1, 3, 3, 2, 5, 1, 1, 4, 3, 3, 7, 3, 4):
            a = dec.cached_decode(z[:, :, pos:pos + n], scale)
            b = oracle.cached_decode(z[:, :, pos:pos + n], scale).float().clamp_(-1, 1)
            errs.append(rel_l2(a, b))
            pos += n
    assert pos == 40
    print("per-call rel-L2:", " ".join(f"{e:.2e}" for e in errs))
    assert max(errs) < 3e-2 and errs[-1] < 1.5 * (sum(errs[:4]) / 4)

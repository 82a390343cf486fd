"""Pins oracle/vae_oracle.py to the reference VAE decoder: bit-exact against the committed golden outputs
(produced by the reference module itself, oracle/make_vae_golden.py) and, when /root/reference is present
(the build container), bit-exact against the live reference module on fresh inputs."""
import os

import pytest
import torch

from oracle import vae_oracle as vo
from oracle.make_vae_golden import CHUNKS, SMALL, latents, scale_of

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "vae_small.pt")


@pytest.mark.parametrize("name,dtype", [("f32", torch.float32), ("bf16", torch.bfloat16)])
def test_oracle_matches_reference_golden_bit_exact(name, dtype):
    gold = torch.load(GOLDEN)
    cfg = vo.VaeConfig(**gold["cfg"])
    dec = vo.VaeDecoderOracle(cfg, vo.init_state_dict(cfg, seed=gold["seed"], dtype=dtype))
    scale = scale_of(cfg, dtype)
    with torch.no_grad():
        for i, t in enumerate(gold["chunks"]):
            out = dec.cached_decode(latents(cfg, 10 + i, t).to(dtype), scale)
            assert out.shape == gold[name]["stream"][i].shape
            assert torch.equal(out, gold[name]["stream"][i]), f"{name}: streaming call {i} differs"
        whole = dec.decode(latents(cfg, 99, 4).to(dtype), scale)
    assert torch.equal(whole, gold[name]["whole"])
    assert not dec.hist and not dec.up_calls  # decode() leaves a clean cache, like the reference


def test_stream_shapes_and_first_frame_rule():
    cfg = vo.VaeConfig(**SMALL)
    dec = vo.VaeDecoderOracle(cfg, vo.init_state_dict(cfg, seed=1))
    scale = scale_of(cfg, torch.float32)
    with torch.no_grad():
        a = dec.cached_decode(latents(cfg, 1, 1), scale)   # the stream's first latent frame -> 1 video frame
        b = dec.cached_decode(latents(cfg, 2, 1), scale)   # every later one -> 4
    assert a.shape[2] == 1 and b.shape[2] == 4 and a.shape[-2:] == (40, 56)


def test_decode_to_pixel_layout_and_clamp():
    cfg = vo.VaeConfig(**SMALL)
    dec = vo.VaeDecoderOracle(cfg, vo.init_state_dict(cfg, seed=2))
    lat = 3.0 * latents(cfg, 3, 2).permute(0, 2, 1, 3, 4)  # [B, T, z, h, w]
    with torch.no_grad():
        vid = dec.decode_to_pixel(lat, use_cache=False)
    assert vid.shape == (1, 5, 3, 40, 56) and vid.dtype == torch.float32
    assert float(vid.max()) <= 1.0 and float(vid.min()) >= -1.0


@pytest.mark.skipif(not os.path.exists("/root/reference/wan/modules/vae.py"), reason="reference tree not present")
def test_oracle_matches_live_reference_module():
    from oracle.make_vae_golden import load_reference_vae, reference_model
    cfg = vo.VaeConfig(dim=8, z_dim=4, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_upsample=(True, True, False))
    sd = vo.init_state_dict(cfg, seed=7)
    ref = reference_model(load_reference_vae(), cfg, sd, torch.float32)
    ref.clear_cache()
    dec = vo.VaeDecoderOracle(cfg, sd)
    scale = [torch.zeros(4), torch.ones(4)]
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for t in (1, 1, 2, 3):
            z = torch.randn(1, 4, t, 4, 6, generator=g)
            assert torch.equal(dec.cached_decode(z, scale), ref.cached_decode(z, scale))

"""Generates tests/golden/vae_small.pt by running the REFERENCE VAE module itself (imported from
/root/reference, which only exists in the build container) on seeded inputs.  The fixture pins
oracle/vae_oracle.py - and through it the CUDA decoder - to the reference's streaming decode.

    python -m oracle.make_vae_golden
"""
import importlib.util
import os

import torch

from oracle import vae_oracle as vo

SMALL = dict(dim=16, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_upsample=(True, True, False))
H, W = 5, 7          # latent size: 40 x 56 pixels, not a multiple of the CUDA conv tile (8 x 16)
CHUNKS = (3, 2, 1)   # latent frames per streaming call


def load_reference_vae():
    spec = importlib.util.spec_from_file_location("ref_vae", "/root/reference/wan/modules/vae.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_model(mod, cfg: vo.VaeConfig, sd, dtype):
    vae = mod.WanVAE_(dim=cfg.dim, z_dim=cfg.z_dim, dim_mult=list(cfg.dim_mult), num_res_blocks=cfg.num_res_blocks,
                      attn_scales=[], temperal_downsample=list(reversed(cfg.temporal_upsample)))
    missing, unexpected = vae.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("encoder.", "conv1.")) for k in missing), (missing, unexpected)
    return vae.to(dtype).eval().requires_grad_(False)


def latents(cfg, seed, t):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, cfg.z_dim, t, H, W, generator=g)


def scale_of(cfg, dtype):
    return [torch.tensor(vo.LATENT_MEAN[:cfg.z_dim]).to(dtype), (1.0 / torch.tensor(vo.LATENT_STD[:cfg.z_dim])).to(dtype)]


def main():
    mod = load_reference_vae()
    cfg = vo.VaeConfig(**SMALL)
    out = {"cfg": SMALL, "H": H, "W": W, "chunks": CHUNKS, "seed": 0}
    for name, dtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        sd = vo.init_state_dict(cfg, seed=0, dtype=dtype)
        ref = reference_model(mod, cfg, sd, dtype)
        ref.clear_cache()
        scale = scale_of(cfg, dtype)
        with torch.no_grad():
            stream = [ref.cached_decode(latents(cfg, 10 + i, t).to(dtype), scale) for i, t in enumerate(CHUNKS)]
            whole = ref.decode(latents(cfg, 99, 4).to(dtype), scale)
        out[name] = {"stream": stream, "whole": whole}
        print(name, [tuple(s.shape) for s in stream], tuple(whole.shape), float(whole.float().abs().mean()))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "vae_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""Full-size streaming VAE decode on the B200: libllb200 (longlive_b200/vae.py) vs the oracle's torch ops
(cuDNN convolutions) on the same GPU, same random-init bf16 weights, same latents.
Prints one JSON object; run on the GPU box:  python tools/vae_bench.py [--chunks 4]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longlive_b200.vae import WanVAEDecoder, _decoder_plan  # noqa: E402

DEV = "cuda"


def decoder_flops_per_latent_frame(dim=96, z=16, mult=(1, 2, 4, 4), h=60, w=104):
    """Algorithmic FLOPs (real channel counts, 2 per multiply-add) of one steady-state latent frame
    (1 / 2 / 4 / 4 frames at the four resolutions)."""
    plan = _decoder_plan(dim, z, list(mult), 2, (True, True, False))
    fl, T, H, W = 0.0, 1, h, w
    for s in plan:
        px = T * H * W
        if s[0] == "conv":
            fl += 2.0 * px * s[2] * s[3] * 27
        elif s[0] == "res":
            fl += 2.0 * px * (s[2] * s[3] * 27 + s[3] * s[3] * 27 + (s[2] * s[3] if s[2] != s[3] else 0))
        elif s[0] == "attn":
            n = H * W
            fl += T * (2.0 * n * s[2] * 4 * s[2] + 4.0 * n * n * s[2])
        elif s[0] == "up":
            if s[3]:
                fl += 2.0 * px * s[2] * 2 * s[2] * 3
                T *= 2
            H, W = 2 * H, 2 * W
            fl += 2.0 * T * H * W * s[2] * (s[2] // 2) * 9
        else:
            fl += 2.0 * px * s[2] * 3 * 27
    return fl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--no-eager", action="store_true")
    a = ap.parse_args()
    from oracle import vae_oracle as vo  # baseline leg only
    cfg = vo.VaeConfig()
    sd = vo.init_state_dict(cfg, seed=0, dtype=torch.bfloat16)
    dec = WanVAEDecoder()
    dec.load_state_dict(sd)
    dec = dec.to(DEV)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 16, 1 + 3 * (a.chunks + 1), 60, 104, generator=g).to(torch.bfloat16).to(DEV)
    scale = [torch.tensor(vo.LATENT_MEAN).to(torch.bfloat16).to(DEV), (1.0 / torch.tensor(vo.LATENT_STD)).to(torch.bfloat16).to(DEV)]
    fl = decoder_flops_per_latent_frame()

    def run(fn, clear):
        clear()
        outs, times = [], []
        outs.append(fn(lat[:, :, :1]))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for c in range(a.chunks + 1):
            torch.cuda.synchronize()
            ev0.record()
            o = fn(lat[:, :, 1 + 3 * c:4 + 3 * c])
            ev1.record()
            torch.cuda.synchronize()
            times.append(ev0.elapsed_time(ev1))
            if c < 2:
                outs.append(o)
        return outs, sorted(times[1:])[len(times[1:]) // 2]

    torch.cuda.reset_peak_memory_stats()
    outs, ms = run(lambda z: dec.cached_decode(z, scale), dec.clear_cache)
    res = {"what": "streaming decode of 3 latent frames (12 video frames at 832x480) per call, steady state, bf16, Wan2.1 VAE shape, random init",
           "native_ms_per_chunk": ms, "native_ms_per_latent_frame": ms / 3, "native_video_fps": 12e3 / ms,
           "algorithmic_tflop_per_latent_frame": fl / 1e12, "native_tflops": 3 * fl / ms / 1e9,
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    if not a.no_eager:
        oracle = vo.VaeDecoderOracle(cfg, sd).to(DEV)
        with torch.no_grad():
            eouts, ems = run(lambda z: oracle.cached_decode(z, scale).float().clamp_(-1, 1), oracle.clear_cache)
        rel = [float(((x.float() - y.float()).norm() / y.float().norm())) for x, y in zip(outs, eouts)]
        res.update({"eager_torch_ms_per_chunk": ems, "speedup_vs_eager_torch": ems / ms, "rel_l2_vs_eager_per_call": rel})
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/vae_bench.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

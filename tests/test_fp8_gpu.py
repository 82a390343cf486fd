"""Optional FP8 (W8A8 e4m3) linear path: kernels vs exact fp32 references built from the SAME
quantised operands, and the end-to-end effect on the model (stated separately from the bf16 gate:
FP8 is a lossy option the reference only advertises, README.md:50 / reports.md:24)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _deq(t8, scale):
    return t8.view(torch.float8_e4m3fn).float() * scale[:, None]


@pytest.mark.parametrize("M,N,K", [(4680, 1536, 1536), (4680, 8960, 1536), (4680, 1536, 8960), (300, 4608, 1536),
                                   (77, 136, 272)])
def test_gemm_fp8_matches_dequantised_reference(M, N, K):
    from longlive_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * torch.rand(M, 1, generator=g) * 3).to(torch.bfloat16).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).to(DEV)
    b = torch.randn(N, generator=g).to(torch.bfloat16).to(DEV)
    a8, sa = ops.quant_rows_fp8(a)
    w8, sw = ops.quantize_weight_e4m3(w)
    out = ops.gemm_fp8(a8, sa, w8, sw, b)
    ref = _deq(a8, sa) @ _deq(w8, sw).t() + b.float()
    assert rel_l2(out, ref) < 4e-3  # only the bf16 output rounding
    # and the quantisation itself is within e4m3 precision of the bf16 GEMM
    full = a.float() @ w.float().t() + b.float()
    assert rel_l2(out, full) < 6e-2


def test_gemm_fp8_gate_residual_epilogue():
    from longlive_b200 import ops
    M, N, K, F = 4680, 1536, 8960, 3
    g = torch.Generator(device="cpu").manual_seed(3)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).to(DEV)
    b = (torch.randn(N, generator=g) * 0.1).to(torch.bfloat16).to(DEV)
    x = torch.randn(M, N, generator=g).to(torch.bfloat16).to(DEV)
    gate = torch.randn(F, N, generator=g).to(torch.bfloat16).to(DEV)
    a8, sa = ops.quant_rows_fp8(a)
    w8, sw = ops.quantize_weight_e4m3(w)
    xin = x.clone()
    out = ops.gemm_fp8(a8, sa, w8, sw, b, epilogue=ops.EPI_BIAS_GATE_RES, gate=gate, rows_per_gate=M // F,
                       res=xin, out=xin)
    y = (_deq(a8, sa) @ _deq(w8, sw).t() + b.float()).to(torch.bfloat16)
    ref = x.float() + (y.float() * gate.float().repeat_interleave(M // F, 0)).to(torch.bfloat16).float()
    assert rel_l2(out, ref) < 5e-3


def test_row_quantisation_kernels():
    from longlive_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn(1000, 8960, generator=g) * 2).to(torch.bfloat16).to(DEV)
    q8, sc = ops.quant_rows_fp8(x)
    amax = x.float().abs().amax(1)
    assert torch.allclose(sc, amax / 448.0, rtol=1e-6)
    ref8 = (x.float() / sc[:, None]).to(torch.float8_e4m3fn).view(torch.uint8)
    assert (q8 != ref8).float().mean().item() < 5e-3  # x * (1/s) vs x / s may differ on exact ties
    assert rel_l2(_deq(q8, sc), x) < 4e-2
    # fused LN + modulate + quantise == bf16 kernel followed by the standalone quantiser
    rows, Cc, F = 4680, 1536, 3
    xx = (torch.randn(rows, Cc, generator=g) * 2 + 0.3).to(torch.bfloat16).to(DEV)
    mod = (torch.randn(F, 6 * Cc, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    y = ops.ln_modulate(xx, shift=mod[:, :Cc], scale=mod[:, Cc:2 * Cc], rows_per_frame=rows // F)
    y8_ref, s_ref = ops.quant_rows_fp8(y)
    y8 = torch.empty(rows, Cc, dtype=torch.uint8, device=DEV); s8 = torch.empty(rows, dtype=torch.float32, device=DEV)
    ops.ln_modulate_fp8(xx, y8, s8, shift=mod[:, :Cc], scale=mod[:, Cc:2 * Cc], rows_per_frame=rows // F)
    assert torch.equal(y8, y8_ref) and torch.equal(s8, s_ref)


def test_model_fp8_linears_vs_bf16():
    """Full-size model, cache fill + one rolling chunk: FP8 linears vs the bf16 CUDA path."""
    import types
    from oracle import wan_oracle as wo
    from oracle.make_golden import SeededNoise
    from longlive_b200.model import CausalWanModel
    from longlive_b200.pipeline import CausalInferencePipeline
    from longlive_b200.wrapper import WanDiffusionWrapper
    cfg = wo.WanConfig()
    sd = wo.init_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(1, 15, 16, 60, 104, generator=g).to(torch.bfloat16).to(DEV)
    prompt = wo.synth_prompt_embeds(cfg, 100, 200).to(DEV)

    class MK(dict):
        __getattr__ = dict.get
    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=3, context_noise=0, global_sink=False,
                                 model_kwargs=MK(local_attn_size=12, sink_size=3, timestep_shift=5.0))
    lats = []
    for fp8 in (False, True):
        model = CausalWanModel(local_attn_size=12, sink_size=3)
        model.load_state_dict(sd)
        model.fp8_linears = fp8
        model = model.to(DEV).to(torch.bfloat16)
        pipe = CausalInferencePipeline(args, torch.device(DEV), generator=WanDiffusionWrapper(model=model, timestep_shift=5.0),
                                       text_encoder=lambda text_prompts: {"prompt_embeds": prompt})
        sn = SeededNoise()
        pipe.renoise_fn = lambda like, b, s: sn(like)
        lats.append(pipe.inference(noise, ["p"], return_latents=True)[1])
        del pipe, model
        torch.cuda.empty_cache()
    errs = [rel_l2(lats[1][:, c:c + 3], lats[0][:, c:c + 3]) for c in range(0, 15, 3)]
    print("fp8-linear latents rel-L2 per chunk vs bf16 path:", [f"{e:.2e}" for e in errs])
    assert max(errs) < 0.12, errs
    assert torch.isfinite(lats[1].float()).all()

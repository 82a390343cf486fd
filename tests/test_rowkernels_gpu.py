"""Per-kernel parity of libllb200.so (through the C ABI) against plain fp32 PyTorch references.

These run on the B200 box (`pytest -m gpu`).  Tolerances are stated per test: the kernels round to
bf16 at the same points as the reference ops, so most comparisons are at bf16 round-off level.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from longlive_b200 import ops
    return ops


def rel_l2(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------ row kernels
def test_ln_modulate_and_affine():
    ops = _ops()
    rows, Cc, F = 4680, 1536, 3
    g = torch.Generator(device="cpu").manual_seed(3)
    x = bf(torch.randn(rows, Cc, generator=g) * 2 + 0.3).to(DEV)
    mod = bf(torch.randn(F, 6 * Cc, generator=g) * 0.5).to(DEV)
    shift, scale = mod[:, 0:Cc], mod[:, Cc:2 * Cc]
    out = ops.ln_modulate(x, shift=shift, scale=scale, rows_per_frame=rows // F)
    ln = torch.nn.functional.layer_norm(x.float(), (Cc,), eps=1e-6)
    ref = bf(bf(bf(ln).float() * bf(1 + scale.float()).float().repeat_interleave(rows // F, 0)).float()
             + shift.float().repeat_interleave(rows // F, 0))
    assert rel_l2(out, ref) < 3e-3
    mism = (out != ref).float().mean().item()
    assert mism < 0.02, f"{mism:.4f} of elements differ from the bf16-stepped reference"
    w = bf(torch.randn(Cc, generator=g)).to(DEV); b = bf(torch.randn(Cc, generator=g)).to(DEV)
    out2 = ops.ln_modulate(x, ln_w=w, ln_b=b)
    ref2 = torch.nn.functional.layer_norm(x.float(), (Cc,), w.float(), b.float(), eps=1e-6)
    assert rel_l2(out2, ref2) < 3e-3


def test_rmsnorm():
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(4)
    x = bf(torch.randn(512, 1536, generator=g) * 3).to(DEV)
    w = bf(torch.randn(1536, generator=g)).to(DEV)
    out = ops.rmsnorm(x, w)
    xf = x.float()
    ref = bf(bf(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).float() * w.float())
    assert rel_l2(out, ref) < 2e-3
    assert (out != ref).float().mean().item() < 0.01


def _rope_ref(x, H, start_frame, F, gh, gw, table):
    """x [L, H*128] bf16 -> roped (fp64 complex math like causal_rope_apply), bf16."""
    L = x.shape[0]
    xc = torch.view_as_complex(x.to(torch.float64).view(L, H, 64, 2))
    t = torch.arange(L, device=x.device)
    f = t // (gh * gw); h = (t % (gh * gw)) // gw; w = t % gw
    tab = torch.view_as_complex(table.to(torch.float64).to(x.device))  # [1024, 64]
    i = torch.arange(64, device=x.device)
    pos = torch.where(i[None, :] < 22, (start_frame + f)[:, None],
                      torch.where(i[None, :] < 43, h[:, None], w[:, None]))
    fr = tab[pos, i[None, :].expand(L, 64)]  # [L, 64]
    out = torch.view_as_real(xc * fr[:, None, :]).flatten(1)
    return out.to(torch.bfloat16)


def test_rmsnorm_rope_append():
    ops = _ops()
    H, F, gh, gw = 12, 3, 30, 52
    L, Cc = F * gh * gw, H * 128
    g = torch.Generator(device="cpu").manual_seed(6)
    qkv = bf(torch.randn(L, 3 * Cc, generator=g)).to(DEV)
    wq = bf(1 + 0.1 * torch.randn(Cc, generator=g)).to(DEV)
    wk = bf(1 + 0.1 * torch.randn(Cc, generator=g)).to(DEV)
    table = ops.build_rope_table().to(DEV)
    kc = torch.zeros(18720, Cc, dtype=torch.bfloat16, device=DEV)
    vc = torch.zeros_like(kc)
    qo = torch.empty(L, Cc, dtype=torch.bfloat16, device=DEV)
    # new tokens [1560, 4680) go to ring rows: first 2000 -> 16720.., remaining 1120 -> 4680..
    writes = [(1560, 16720, 2000), (3560, 4680, 1120)]
    sp = ops.step_params_tensor(ops.make_step_params(rope_start_frame=7, writes=writes), DEV)
    ops.rmsnorm_rope_append(qkv, qo, kc, vc, wq, wk, table, (gh, gw), sp, n_heads=H)

    def norm(x, w):
        xf = x.float()
        return bf(bf(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).float() * w.float())

    q_ref = _rope_ref(norm(qkv[:, :Cc], wq), H, 7, F, gh, gw, table)
    k_ref = _rope_ref(norm(qkv[:, Cc:2 * Cc], wk), H, 7, F, gh, gw, table)
    assert rel_l2(qo, q_ref) < 2e-3
    assert (qo != q_ref).float().mean().item() < 0.01
    k_exp = torch.zeros_like(kc); v_exp = torch.zeros_like(vc)
    for s, d, n in writes:
        k_exp[d:d + n] = k_ref[s:s + n]
        v_exp[d:d + n] = qkv[s:s + n, 2 * Cc:]
    assert torch.equal(vc, v_exp), "V rows must be copied bit-exactly to the planned ring rows"
    assert (kc != k_exp).float().mean().item() < 0.01
    written = torch.zeros(18720, dtype=torch.bool, device=DEV)
    for s, d, n in writes:
        written[d:d + n] = True
    assert kc[~written].abs().max().item() == 0, "rows outside the plan must stay untouched"


def test_glue_kernels():
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(8)
    x = bf(torch.randn(16, 3, 60, 104, generator=g)).to(DEV)
    p = ops.patchify(x)
    ref = x.view(16, 3, 30, 2, 52, 2).permute(1, 2, 4, 0, 3, 5).reshape(3 * 30 * 52, 64)
    assert torch.equal(p, ref)
    y = bf(torch.randn(3 * 30 * 52, 64, generator=g)).to(DEV)
    u = ops.unpatchify(y, 16, 3, 60, 104)
    ref_u = torch.einsum("fhwpqrc->cfphqwr", y.view(3, 30, 52, 1, 2, 2, 16)).reshape(16, 3, 60, 104)
    assert torch.equal(u, ref_u)
    t = torch.tensor([1000.0, 937.5, 0.0, 625.0], device=DEV)
    s = ops.sinusoidal(t, 256)
    half = 128
    sinus = torch.outer(t.double(), torch.pow(10000, -torch.arange(half, device=DEV).double() / half))
    ref_s = torch.cat([torch.cos(sinus), torch.sin(sinus)], dim=1).to(torch.bfloat16)
    assert (s.float() - ref_s.float()).abs().max().item() <= 2 ** -7
    mod = bf(torch.randn(30, 6 * 1536, generator=g)).to(DEV)
    e0 = bf(torch.randn(3, 6 * 1536, generator=g)).to(DEV)
    tab = ops.modulation_table(mod, e0)
    assert torch.equal(tab, (mod[:, None, :] + e0[None]))
    z = bf(torch.randn(3, 1536, generator=g)).to(DEV)
    assert rel_l2(ops.silu(z), torch.nn.functional.silu(z.float())) < 3e-3

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


# ------------------------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [
    (128, 128, 64), (128, 128, 128), (256, 256, 512), (4680, 1536, 1536), (300, 1536, 1536),
    (4680, 4608, 1536), (4680, 8960, 1536), (4680, 1536, 8960), (3, 1536, 256), (3, 9216, 1536),
    (4680, 64, 1536), (512, 1536, 4096), (4680, 1536, 64), (72, 136, 200),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bias(M, N, K):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    b = bf(torch.randn(N, generator=g)).to(DEV)
    out = ops.gemm(a, w, b)
    ref = a.float() @ w.float().t() + b.float()
    torch.cuda.synchronize()
    assert out.shape == (M, N)
    err = rel_l2(out, ref)
    assert err < 4e-3, f"rel-L2 {err}"  # bf16 output rounding is ~2e-3 rel-L2
    # elementwise: within 1 bf16 ulp of the fp32 result (plus fp32 accumulation-order noise)
    diff = (out.float() - ref).abs()
    tol = ref.abs() * 2 ** -7 + 1e-2
    assert (diff <= tol).all(), f"max abs diff {diff.max().item()}"


def test_gemm_no_bias_strided():
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(1)
    big = bf(torch.randn(200, 3 * 256, generator=g)).to(DEV)
    a = big[:, 256:512]  # row stride 768
    w = bf(torch.randn(128, 256, generator=g) / 16).to(DEV)
    outbuf = torch.zeros(200, 512, dtype=torch.bfloat16, device=DEV)
    out = ops.gemm(a, w, None, out=outbuf[:, 128:256])
    ref = a.float() @ w.float().t()
    assert rel_l2(out, ref) < 4e-3
    assert outbuf[:, :128].abs().max().item() == 0 and outbuf[:, 256:].abs().max().item() == 0


@pytest.mark.parametrize("epi", ["gelu", "silu", "gate_res", "res"])
def test_gemm_epilogues(epi):
    ops = _ops()
    M, N, K, F = 4680, 1536, 1536, 3
    g = torch.Generator(device="cpu").manual_seed(11)
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    b = bf(torch.randn(N, generator=g) * 0.1).to(DEV)
    y = bf(a.float() @ w.float().t() + b.float())  # the reference's materialised Linear output
    if epi == "gelu":
        out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GELU)
        ref = torch.nn.functional.gelu(y.float(), approximate="tanh")
    elif epi == "silu":
        out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_SILU)
        ref = torch.nn.functional.silu(y.float())
    elif epi == "gate_res":
        x = bf(torch.randn(M, N, generator=g)).to(DEV)
        gate = bf(torch.randn(F, 6 * N, generator=g)).to(DEV)[:, 2 * N:3 * N]  # strided view
        xin = x.clone()
        out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GATE_RES, gate=gate, rows_per_gate=M // F,
                       res=xin, out=xin)  # in place, like x = x + y * e[2]
        gfull = gate.float().repeat_interleave(M // F, dim=0)
        ref = x.float() + bf(y.float() * gfull).float()
    else:
        x = bf(torch.randn(M, N, generator=g)).to(DEV)
        out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_RES, res=x)
        ref = x.float() + y.float()
    err = rel_l2(out, ref)
    assert err < 5e-3, f"{epi}: rel-L2 {err}"




# Every tile mode the heuristic can pick (single CTA or CTA pair x tile width), forced through the
# library's LLB_GEMM_TILE override, on ragged shapes: M not a multiple of 256 (the second CTA of the
# last pair is partly / entirely out of bounds), N not a multiple of the tile width, K tail.
TILE_MODES = [(0, 128), (0, 192), (0, 256), (1, 128), (1, 192), (1, 256)]


@pytest.fixture
def force_tile(monkeypatch):
    def _set(pair, bn):
        monkeypatch.setenv("LLB_GEMM_TILE", f"{pair},{bn}")
    yield _set
    monkeypatch.delenv("LLB_GEMM_TILE", raising=False)


@pytest.mark.parametrize("pair,bn", TILE_MODES)
@pytest.mark.parametrize("M,N,K", [(4680, 1536, 1536), (1560, 4608, 512), (72, 136, 200), (385, 8960, 320),
                                   (129, 200, 64), (300, 1536, 8960)])
def test_gemm_tile_modes(force_tile, pair, bn, M, N, K):
    ops = _ops()
    force_tile(pair, bn)
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    b = bf(torch.randn(N, generator=g)).to(DEV)
    x = bf(torch.randn(M, N, generator=g)).to(DEV)
    guard = torch.full((M + 4, N), 7.0, dtype=torch.bfloat16, device=DEV)  # rows past M must stay untouched
    out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_RES, res=x, out=guard[:M])
    ref = x.float() + bf(a.float() @ w.float().t() + b.float()).float()
    torch.cuda.synchronize()
    err = rel_l2(out, ref)
    assert err < 5e-3, f"pair={pair} bn={bn}: rel-L2 {err}"
    assert (guard[M:] == 7.0).all()


@pytest.mark.parametrize("pair,bn", TILE_MODES)
def test_gemm_tile_modes_bitwise_equal(force_tile, pair, bn):
    """All tile modes accumulate each output in the same k order, so they agree bit for bit."""
    ops = _ops()
    M, N, K = 1000, 1536, 1536
    g = torch.Generator(device="cpu").manual_seed(5)
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    b = bf(torch.randn(N, generator=g)).to(DEV)
    force_tile(0, 128)
    base = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GELU)
    force_tile(pair, bn)
    out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GELU)
    assert torch.equal(out, base)

"""Small-shape invocations of every hand-written kernel family, meant to run under
`compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck` (tools/run_sanitizers.sh): GEMM single tile
and CTA pair (cluster, cta_group::2, multicast commit) with every epilogue class, fp8, split-K; attention incl.
the stream-K split / merge path (cross-CTA flag spins) and multi-range key sets; row kernels; conv3d with ring
history; t5_attn; the peer barrier.  Each case also checks its result against PyTorch, so a sanitizer pass
means "clean AND correct".  Shapes are tiny on purpose: racecheck serialises shared-memory accesses."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longlive_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
bf = lambda x: x.to(torch.bfloat16)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def case_gemm():
    g = torch.Generator().manual_seed(0)
    for pair, bn in ((0, 128), (0, 192), (0, 256), (1, 128), (1, 256)):
        os.environ["LLB_GEMM_TILE"] = f"{pair},{bn}"
        for (M, N, K) in ((72, 136, 200), (300, 520, 256)):
            a = bf(torch.randn(M, K, generator=g)).to(DEV)
            w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
            b = bf(torch.randn(N, generator=g)).to(DEV)
            x = bf(torch.randn(M, N, generator=g)).to(DEV)
            gate = bf(torch.randn(2, N, generator=g)).to(DEV)
            base = a.float() @ w.float().t() + b.float()
            assert rel(ops.gemm(a, w, b), base) < 4e-3
            assert rel(ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GELU),
                       torch.nn.functional.gelu(bf(base).float(), approximate="tanh")) < 6e-3
            out = ops.gemm(a, w, b, epilogue=ops.EPI_BIAS_GATE_RES, gate=gate, rows_per_gate=(M + 1) // 2, res=x,
                           out=x.clone())
            rows = torch.arange(M, device=DEV) // ((M + 1) // 2)
            assert rel(out, x.float() + bf(bf(base).float() * gate.float()[rows]).float()) < 5e-3
    os.environ.pop("LLB_GEMM_TILE", None)
    # split-K with the fused residual + RMS norm reduce
    M, N, K = 128, 256, 1024
    a = bf(torch.randn(M, K, generator=g)).to(DEV); w = bf(torch.randn(N, K, generator=g) / 32).to(DEV)
    r = bf(torch.randn(M, N, generator=g)).to(DEV)
    ws = torch.empty(2 * M * N, dtype=torch.float32, device=DEV)
    out = ops.gemm_splitk(a, w, ws, 2, None, res=r)
    assert rel(out, r.float() + a.float() @ w.float().t()) < 4e-3
    # fp8
    a8, sa = ops.quant_rows_fp8(bf(torch.randn(96, 256, generator=g)).to(DEV))
    w8, sw = ops.quantize_weight_e4m3(bf(torch.randn(160, 256, generator=g) / 16).to(DEV))
    ops.gemm_fp8(a8, sa, w8, sw, None)


def _attn_ref(q, k, v, H, segs):
    Lq = q.shape[0]
    idx = torch.cat([torch.arange(s, s + n, device=q.device) for s, n in segs])
    qh = q.float().view(Lq, H, 128).transpose(0, 1)
    kh = k.float()[idx].view(-1, H, 128).transpose(0, 1)
    vh = v.float()[idx].view(-1, H, 128).transpose(0, 1)
    p = torch.softmax(torch.einsum("hqd,hkd->hqk", qh, kh) * 128 ** -0.5, dim=-1)
    return torch.einsum("hqk,hkd->hqd", p, vh).transpose(0, 1).reshape(Lq, H * 128)


def case_attention():
    g = torch.Generator().manual_seed(1)
    # (Lq, H, rows, segs): one tile; ragged; multi-range; a shape whose items are all split over CTAs
    for (Lq, H, rows, segs) in ((128, 1, 128, [(0, 128)]), (200, 2, 300, [(0, 300)]),
                                (130, 1, 1000, [(5, 77), (300, 129)]), (520, 2, 1500, [(0, 1500)]),
                                (300, 3, 512, [(0, 512)])):
        q = bf(torch.randn(Lq, H * 128, generator=g)).to(DEV)
        k = bf(torch.randn(rows, H * 128, generator=g)).to(DEV)
        v = bf(torch.randn(rows, H * 128, generator=g)).to(DEV)
        sp = ops.step_params_tensor(ops.make_step_params(attn_segs=segs), DEV)
        out = ops.attention(q, k, v, sp, n_heads=H)
        assert rel(out, _attn_ref(q, k, v, H, segs)) < 8e-3
        assert torch.equal(out, ops.attention(q, k, v, sp, n_heads=H))  # workspace flags consumed


def case_rows():
    g = torch.Generator().manual_seed(2)
    H, F, gh, gw = 2, 2, 4, 6
    L, Cc = F * gh * gw, H * 128
    qkv = bf(torch.randn(L, 3 * Cc, generator=g)).to(DEV)
    wq = bf(1 + 0.1 * torch.randn(Cc, generator=g)).to(DEV); wk = bf(1 + 0.1 * torch.randn(Cc, generator=g)).to(DEV)
    table = ops.build_rope_table().to(DEV)
    kc = torch.zeros(96, Cc, dtype=torch.bfloat16, device=DEV); vc = torch.zeros_like(kc)
    qo = torch.empty(L, Cc, dtype=torch.bfloat16, device=DEV)
    sp = ops.step_params_tensor(ops.make_step_params(3, writes=[(0, 80, 16), (16, 24, L - 16)]), DEV)
    ops.rmsnorm_rope_append(qkv, qo, kc, vc, wq, wk, table, (gh, gw), sp, n_heads=H)
    assert torch.equal(vc[80:96], qkv[:16, 2 * Cc:]) and torch.equal(vc[24:24 + L - 16], qkv[16:, 2 * Cc:])
    x = bf(torch.randn(50, 1536, generator=g)).to(DEV)
    sh = bf(torch.randn(2, 1536, generator=g)).to(DEV); sc = bf(torch.randn(2, 1536, generator=g)).to(DEV)
    y = ops.ln_modulate(x, shift=sh, scale=sc, rows_per_frame=25)
    ln = torch.nn.functional.layer_norm(x.float(), (1536,), eps=1e-6)
    rows = torch.arange(50, device=DEV) // 25
    assert rel(y, bf(bf(bf(ln).float() * (1 + sc.float()[rows])).float() + sh.float()[rows])) < 4e-3
    w = bf(1 + 0.1 * torch.randn(1536, generator=g)).to(DEV)
    n = ops.rmsnorm(x, w)
    xf = x.float()
    assert rel(n, bf(bf(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).float() * w.float())) < 3e-3
    ops.rmsnorm(bf(torch.randn(9, 4096, generator=g)).to(DEV), bf(torch.ones(4096)).to(DEV))


def case_conv3d():
    import torch.nn.functional as Fn
    from longlive_b200 import vae
    g = torch.Generator().manual_seed(3)
    for mt in ("1", "2"):
        os.environ["LLB_CONV_MT"] = mt
        for (cin, cout, k, H, W) in ((64, 64, (3, 3, 3), 5, 7), (32, 96, (3, 3, 3), 9, 17), (96, 32, (1, 3, 3), 8, 9)):
            kt, kh, kw = k
            w = bf(torch.randn(cout, cin, kt, kh, kw, generator=g) / (cin * kt * kh * kw) ** 0.5)
            b = bf(0.1 * torch.randn(cout, generator=g))
            wp = w.permute(0, 2, 3, 4, 1).reshape(cout, -1).contiguous().to(DEV)
            cinp, coutp = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
            ring = vae.FrameRing(5, H, W, cinp, DEV)
            new = bf(torch.randn(cin, 2, H, W, generator=g))
            t0 = ring.reserve(2)
            for i in range(2):
                ring.buf[(t0 + i) % 5][..., :cin].copy_(new[:, i].permute(1, 2, 0))
            out = torch.zeros((2, H, W, coutp), dtype=torch.bfloat16, device=DEV)
            vae.conv3d(ring.buf, t0, wp, b.to(DEV), k, out, 2)
            xin = torch.cat([torch.zeros(cin, kt - 1, H, W), new.float()], 1) if kt == 3 else new.float()
            ref = Fn.conv3d(Fn.pad(xin.unsqueeze(0), (kw // 2, kw // 2, kh // 2, kh // 2, 0, 0)), w.float(), b.float())[0]
            assert rel(out[..., :cout].cpu(), ref.permute(1, 2, 3, 0)) < 4e-3
    os.environ.pop("LLB_CONV_MT", None)


def case_t5():
    from longlive_b200.text_encoder import relative_position_buckets
    g = torch.Generator().manual_seed(4)
    for (Lp, lens, H) in ((128, (1, 37), 2), (256, (256, 130), 2)):
        B = len(lens)
        qkv = bf(torch.randn(B * Lp, 3 * H * 64, generator=g))
        qkv[:, :H * 64] *= 0.4
        qkv = qkv.to(DEV)
        pos = bf(torch.randn(32, H, generator=g) * 0.7).to(DEV)
        lut = relative_position_buckets(512).to(DEV)
        out = ops.t5_attention(qkv, B, H, torch.tensor(lens, dtype=torch.int32, device=DEV), pos, lut)
        assert torch.isfinite(out.float()).all()


def case_barrier():
    flags = torch.zeros(8, dtype=torch.int32, device=DEV)
    ptrs = torch.tensor([flags.data_ptr()], dtype=torch.int64, device=DEV)
    epoch = torch.zeros(1, dtype=torch.int32, device=DEV)
    for _ in range(3):
        ops.peer_barrier(ptrs, 0, 1, epoch)
    torch.cuda.synchronize()
    assert int(epoch.item()) == 3


CASES = {"gemm": case_gemm, "attention": case_attention, "rows": case_rows, "conv3d": case_conv3d, "t5": case_t5,
         "barrier": case_barrier}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        CASES[n]()
        torch.cuda.synchronize()
        print(f"case {n}: ok ({ops.launch_count()} launches so far)", flush=True)
    print("SANITIZER_CASES_OK")

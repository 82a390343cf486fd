"""Head-parallel (Ulysses) kernels on ONE GPU: the ranks are emulated by giving every "peer" its own
buffer on the same device, so the fused exchange stores of llb_rmsnorm_rope_append / llb_attn_fwd and
the device-side barrier are exercised without a multi-GPU box.  The multi-GPU run itself
(tools/ulysses_check.py under torchrun on 2 and 4 B200s) is recorded in profiles/r01_ulysses_*.json."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("P", [2, 4, 8])
def test_sharded_append_and_attention_match_unsharded(P):
    """P = 2, 4: contiguous head blocks; P = 8: 12 heads dealt round-robin (ranks 0-3 own two heads, 4-7 one)."""
    from longlive_b200 import _lib, ops
    H, F, gh, gw = 12, 3, 30, 52
    L, Cc, size = F * gh * gw, H * 128, 12 * 1560
    rr = H % P != 0
    hp, Lp = -(-H // P), L // P
    hw = hp * 128
    heads_of = [list(range(r, H, P)) if rr else list(range(r * hp, (r + 1) * hp)) for r in range(P)]
    cols_of = [torch.cat([torch.arange(h * 128, (h + 1) * 128) for h in hs]).to(DEV) for hs in heads_of]
    g = torch.Generator(device="cpu").manual_seed(9)
    qkv = (torch.randn(L, 3 * Cc, generator=g)).to(torch.bfloat16).to(DEV)
    wq = (1 + 0.1 * torch.randn(Cc, generator=g)).to(torch.bfloat16).to(DEV)
    wk = (1 + 0.1 * torch.randn(Cc, generator=g)).to(torch.bfloat16).to(DEV)
    table = ops.build_rope_table().to(DEV)
    old_k = torch.randn(size, Cc, generator=g).to(torch.bfloat16).to(DEV)
    old_v = torch.randn(size, Cc, generator=g).to(torch.bfloat16).to(DEV)
    # ring plan: new tokens wrap around the end of the rolling region; attend everything
    writes = [(0, size - 3000, 3000), (3000, 4680, L - 3000)]
    sp = ops.step_params_tensor(ops.make_step_params(5, writes=writes, attn_segs=[(0, size)]), DEV)

    # ---- single-GPU reference
    kc, vc = old_k.clone(), old_v.clone()
    q_ref = torch.empty(L, Cc, dtype=torch.bfloat16, device=DEV)
    ops.rmsnorm_rope_append(qkv, q_ref, kc, vc, wq, wk, table, (gh, gw), sp, n_heads=H)
    o_ref = ops.attention(q_ref, kc, vc, sp, n_heads=H)

    # ---- P emulated ranks: head-sharded Q / K / V, token-sharded attention output
    q_sh = [torch.zeros(L, hw, dtype=torch.bfloat16, device=DEV) for _ in range(P)]
    def shard_of(t, r):   # rank r's [rows, hp * 128] buffer holding its heads of t (unused columns zero)
        buf = torch.zeros(t.shape[0], hw, dtype=t.dtype, device=DEV)
        buf[:, :cols_of[r].numel()] = t[:, cols_of[r]]
        return buf
    k_sh = [shard_of(old_k, r) for r in range(P)]
    v_sh = [shard_of(old_v, r) for r in range(P)]
    o_sh = [torch.zeros(Lp, Cc, dtype=torch.bfloat16, device=DEV) for _ in range(P)]
    for r in range(P):  # token shard r sends its head slices to every "rank"
        sh = _lib.QkvShard()
        sh.n_ranks, sh.heads_per_rank, sh.row0, sh.round_robin = P, hp, r * Lp, int(rr)
        for j in range(P):
            sh.q_peers[j], sh.k_peers[j], sh.v_peers[j] = q_sh[j].data_ptr(), k_sh[j].data_ptr(), v_sh[j].data_ptr()
        ops.rmsnorm_rope_append(qkv[r * Lp:(r + 1) * Lp], None, None, None, wq, wk, table, (gh, gw), sp,
                                n_heads=H, shard=sh)
    for r in range(P):
        n = cols_of[r].numel()
        assert torch.equal(q_sh[r][:, :n], q_ref[:, cols_of[r]])
        assert torch.equal(k_sh[r][:, :n], kc[:, cols_of[r]])
        assert torch.equal(v_sh[r][:, :n], vc[:, cols_of[r]])
        assert q_sh[r][:, n:].abs().max().item() == 0 if n < hw else True
    for r in range(P):  # head shard r returns its output columns to the owners of the token rows
        osd = _lib.OutShard()
        osd.n_ranks, osd.rows_per_rank, osd.ld_out = P, Lp, Cc
        osd.head_col0, osd.head_col_stride = (r * 128, P * 128) if rr else (r * hw, 128)
        for j in range(P):
            osd.out_peers[j] = o_sh[j].data_ptr()
        ops.attention(q_sh[r], k_sh[r], v_sh[r], sp, n_heads=len(heads_of[r]), shard=osd)
    o_par = torch.cat(o_sh, 0)
    err = rel_l2(o_par, o_ref)
    print(f"P={P}: sharded vs unsharded attention rel-L2 {err:.3e}")
    assert err < 4e-3, f"P={P}: rel-L2 {err}"  # kv-split / merge order and the lazy-max P rounding differ


def test_peer_barrier_single_rank_and_sequence():
    from longlive_b200 import ops
    flags = torch.zeros(8, dtype=torch.int32, device=DEV)
    ptrs = torch.tensor([flags.data_ptr()], dtype=torch.int64, device=DEV)
    epoch = torch.zeros(1, dtype=torch.int32, device=DEV)
    for i in range(5):
        ops.peer_barrier(ptrs, 0, 1, epoch)
    torch.cuda.synchronize()
    assert int(epoch.item()) == 5 and int(flags[0].item()) == 5

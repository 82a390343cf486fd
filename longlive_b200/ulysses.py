"""Optional single-stream head parallelism over P GPUs of one NVLink/NVSwitch box.

Model after the reference's only sequence-parallel code, the Ulysses scheme of
wan/distributed/xdit_context_parallel.py:131-192 (never wired to the causal model there): tokens
are sharded L/P per rank for every GEMM and row kernel (weights replicated), heads are sharded
12/P per rank for self-attention and for the KV ring (cache memory / P).

B200-native twist: there is no all-to-all on the data path.  The two exchanges per block are fused
into the producing kernels as direct stores into the peers' symmetric (NVLink-mapped) buffers:
  llb_rmsnorm_rope_append  sends head h of the roped Q and of K / V to rank h // (12/P): Q into
                           that rank's [L, 12/P*128] query buffer, K / V straight into its ring rows;
  llb_attn_fwd             returns output rows to the rank that owns the token rows,
each followed by llb_peer_barrier (a one-CTA flag exchange over the same mapping).  NCCL
(torch.distributed) is only used to set up the symmetric allocation and to all-gather the final
[L, 64] head output.  Cross-attention needs no exchange (text K/V are replicated).

Every rank runs the same pipeline code on the same inputs (same seeds), so host-side ring plans
are identical; P must divide the tokens per chunk.  P in {2, 3, 4, 6} divides the 12 heads of Wan-1.3B
(contiguous head blocks); P = 8 deals the heads round-robin (ranks 0-3 own two heads, ranks 4-7 one: the
attention step then scales like P = 6 while every GEMM / row kernel scales by 8).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.distributed as dist

from . import _lib, ops
from .model import CausalWanModel


class SymmetricArena:
    """One symmetric allocation per rank, bump-allocated; gives local tensor views and peer addresses."""

    def __init__(self, nbytes: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert self.ptrs[self.rank] == self.buf.data_ptr()
        self.off = 0
        self.nbytes = nbytes

    def alloc(self, shape, dtype) -> "SymTensor":
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        off = (self.off + 255) // 256 * 256
        assert off + nbytes <= self.nbytes, "symmetric arena too small"
        self.off = off + nbytes
        local = self.buf[off:off + nbytes].view(dtype).view(*shape)
        return SymTensor(local, [p + off for p in self.ptrs])


class SymTensor:
    def __init__(self, local: torch.Tensor, peer_ptrs: List[int]):
        self.local, self.peer_ptrs = local, peer_ptrs


class UlyssesCausalWanModel(CausalWanModel):
    """CausalWanModel whose forward is split over the ranks of `group` (see module docstring)."""

    def setup_parallel(self, group=None, max_tokens: int = 12 * 1560, cache_tokens: int = 12 * 1560,
                       use_cuda_graph: bool = False):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.P = dist.get_rank(self.group), dist.get_world_size(self.group)
        assert self.P <= min(_lib.LLB_MAX_RANKS, self.num_heads)
        # P divides the head count: contiguous head blocks.  Otherwise (12 heads over 8 ranks) heads are dealt
        # round-robin: rank r owns heads r, r + P, ... - ranks 0-3 two heads, ranks 4-7 one; every rank's
        # buffers are `hp` = 2 heads wide and the attention launch of a rank covers the heads it owns.
        self.round_robin = self.num_heads % self.P != 0
        self.hp = -(-self.num_heads // self.P)
        self.my_heads = len(range(self.rank, self.num_heads, self.P)) if self.round_robin else self.hp
        dev = self.patch_embedding.weight.device
        hw = self.hp * 128
        kv_bytes = self.num_layers * 2 * cache_tokens * hw * 2
        need = kv_bytes + max_tokens * hw * 2 + (max_tokens // self.P + 8) * self.dim * 2 + (1 << 20)
        self.arena = SymmetricArena(need, dev, self.group)
        self.q_sym = self.arena.alloc((max_tokens, hw), torch.bfloat16)
        self.attn_sym = self.arena.alloc((max_tokens // self.P + 8, self.dim), torch.bfloat16)
        self.k_sym = [self.arena.alloc((cache_tokens, hw), torch.bfloat16) for _ in range(self.num_layers)]
        self.v_sym = [self.arena.alloc((cache_tokens, hw), torch.bfloat16) for _ in range(self.num_layers)]
        self.flags = self.arena.alloc((_lib.LLB_MAX_RANKS,), torch.int32)
        self.flag_ptrs_dev = torch.tensor(self.flags.peer_ptrs, dtype=torch.int64, device=dev)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cache_tokens = cache_tokens
        # graph capture also records the NCCL all-gather of the head output and the peer barriers
        self.use_cuda_graph = use_cuda_graph
        dist.barrier(self.group)
        torch.cuda.synchronize()
        return self

    # ---- cache allocation hook used by the pipelines: head-sharded ring in symmetric memory
    def allocate_kv_cache(self, batch_size: int, size: int, dtype, device):
        assert batch_size == 1 and size <= self.cache_tokens and dtype == torch.bfloat16
        index = torch.zeros(self.num_layers, 2, dtype=torch.long, device=device)
        cache = []
        for i in range(self.num_layers):
            k = self.k_sym[i].local[:size].view(1, size, self.hp, 128)
            v = self.v_sym[i].local[:size].view(1, size, self.hp, 128)
            k.zero_(); v.zero_()
            cache.append({"k": k, "v": v, "global_end_index": index[i, 0:1], "local_end_index": index[i, 1:2]})
        cache[0]["_llb_index_tensor"] = index
        return cache

    def _barrier(self):
        ops.peer_barrier(self.flag_ptrs_dev, self.rank, self.P, self.epoch)

    # ---- optional per-phase timeline (eager mode only): CUDA events around the five phases of every block
    PHASES = ("dense", "append_send", "barrier_qkv", "attn_send", "barrier_out")

    def start_timeline(self):
        """Record events for the next forwards (use_cuda_graph must be False); read with phase_ms()."""
        assert not self.use_cuda_graph, "the timeline records events between launches: eager mode only"
        self._tl = []

    def _mark(self, phase: str):
        tl = getattr(self, "_tl", None)
        if tl is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            tl.append((phase, ev))

    def phase_ms(self) -> dict:
        """Milliseconds per phase summed over everything recorded since start_timeline(): each mark closes the
        phase named by the PREVIOUS mark.  'barrier_*' = wait for the slowest rank + latency of the remote stores
        that are still in flight; '*_send' = the producing kernel including its NVLink stores."""
        torch.cuda.synchronize()
        tl, self._tl = self._tl, None
        out = {k: 0.0 for k in self.PHASES}
        for (ph, a), (_, b) in zip(tl[:-1], tl[1:]):
            if ph != "end":
                out[ph] += a.elapsed_time(b)
        return out

    def _run_blocks(self, b: dict, kv_cache, crossattn_cache, B: int, F: int, H: int, W: int):
        assert B == 1, "head-parallel mode handles one stream"
        Pk = self._packed
        C_, eps = self.dim, self.eps
        gh, gw = H // 2, W // 2
        fs = gh * gw
        L = F * fs
        P, r = self.P, self.rank
        assert L % P == 0, "tokens per chunk must divide evenly over the ranks"
        Lp = L // P
        r0 = r * Lp
        rows = slice(0, Lp)
        hw = self.hp * 128
        v = self.attn_variant
        # Layer 0 stores K / V straight into the PEERS' rings before the first in-forward barrier.  A peer may still
        # be zeroing its ring slices (allocate_kv_cache, the recache reset of a prompt switch) when a faster rank
        # gets here, so every forward starts with a barrier: no remote store can precede a rank's local zeroing.
        self._barrier()
        # embeddings: every rank patchifies the (replicated) input, then works on its own rows
        ops.patchify(b["x_in"][0], out=b["patches"][:L])
        x, xm = b["x"][rows], b["xm"][rows]
        ops.gemm(b["patches"][r0:r0 + Lp], Pk["patch_w"], Pk["patch_b"], out=x)
        ops.sinusoidal(b["t_in"], self.freq_dim, out=b["temb"])
        ops.gemm(b["temb"], Pk["time0_w"], Pk["time0_b"], epilogue=ops.EPI_BIAS_SILU, out=b["te1"])
        ops.gemm(b["te1"], Pk["time1_w"], Pk["time1_b"], out=b["e"])
        ops.silu(b["e"], out=b["es"])
        ops.gemm(b["es"], Pk["tproj_w"], Pk["tproj_b"], out=b["e0"])
        ops.modulation_table(Pk["mod_all"], b["e0"], out=b["mod"])
        qkv, qbuf, cq, h = b["qkv"][rows], b["q"][rows], b["cq"][rows], b["h"][rows]
        attn_local = self.attn_sym.local[:Lp]
        q_full = self.q_sym.local[:L]
        out_sh = _lib.OutShard()
        out_sh.n_ranks, out_sh.rows_per_rank, out_sh.ld_out = P, Lp, C_
        if self.round_robin:   # local head i is global head r + i * P
            out_sh.head_col0, out_sh.head_col_stride = r * 128, P * 128
        else:
            out_sh.head_col0, out_sh.head_col_stride = r * hw, 128
        for j in range(P):
            out_sh.out_peers[j] = self.attn_sym.peer_ptrs[j]
        f8 = self.fp8_linears

        def lin(name, lw, a, **kw):
            """One block Linear on this rank's token rows: bf16 GEMM, or W8A8 (a = (a8, scale)) when fp8_linears."""
            if f8:
                return ops.gemm_fp8(a[0], a[1], lw[name + "_w8"], lw[name + "_ws"], lw[name + "_b"], **kw)
            return ops.gemm(a, lw[name + "_w"], lw[name + "_b"], **kw)

        def ln_in(**kw):
            if f8:
                return ops.ln_modulate_fp8(x, b["a8"][rows], b["sa"][rows], eps=eps, **kw)
            return ops.ln_modulate(x, eps=eps, out=xm, **kw)

        def act_in(t, buf="a8"):
            return ops.quant_rows_fp8(t, b[buf][rows], b["sa"][rows]) if f8 else t

        self._mark("dense")
        for i, lw in enumerate(Pk["layers"]):
            m = b["mod"][i]
            e = [m[:, k * C_:(k + 1) * C_] for k in range(6)]
            lin("qkv", lw, ln_in(shift=e[0], scale=e[1], rows_per_frame=fs, row0=r0), out=qkv)
            self._mark("append_send")
            sh = _lib.QkvShard()
            sh.n_ranks, sh.heads_per_rank, sh.row0, sh.round_robin = P, self.hp, r0, int(self.round_robin)
            for j in range(P):
                sh.q_peers[j] = self.q_sym.peer_ptrs[j]
                sh.k_peers[j] = self.k_sym[i].peer_ptrs[j]
                sh.v_peers[j] = self.v_sym[i].peer_ptrs[j]
            ops.rmsnorm_rope_append(qkv, None, None, None, lw["nq"], lw["nk"], Pk["rope"], (gh, gw), b["params"],
                                    n_heads=self.num_heads, eps=eps, shard=sh)
            self._mark("barrier_qkv")
            self._barrier()  # every rank's head slices (Q, K, V) have landed
            self._mark("attn_send")
            k2 = kv_cache[i]["k"][0].view(-1, hw)
            v2 = kv_cache[i]["v"][0].view(-1, hw)
            ops.attention(q_full, k2, v2, b["params"], n_heads=self.my_heads, out=q_full, variant=v, shard=out_sh)
            self._mark("barrier_out")
            self._barrier()  # every rank's token rows of the attention output have landed
            self._mark("dense")
            lin("o", lw, act_in(attn_local), epilogue=ops.EPI_BIAS_GATE_RES, gate=e[2], rows_per_gate=fs,
                gate_row0=r0, res=x, out=x)
            lin("cq", lw, ln_in(ln_w=lw["n3_w"], ln_b=lw["n3_b"]), out=cq)
            ops.rmsnorm(cq, lw["cnq"], eps, out=qbuf)
            ck, cv = crossattn_cache[i]["k"], crossattn_cache[i]["v"]
            ops.attention(qbuf, ck[0].view(-1, C_), cv[0].view(-1, C_), Pk["cross_segs"], n_heads=self.num_heads,
                          out=b["attn"][rows], variant=v)
            lin("co", lw, act_in(b["attn"][rows]), epilogue=ops.EPI_BIAS_RES, res=x, out=x)
            lin("f1", lw, ln_in(shift=e[3], scale=e[4], rows_per_frame=fs, row0=r0), epilogue=ops.EPI_BIAS_GELU, out=h)
            lin("f2", lw, act_in(h, "h8"), epilogue=ops.EPI_BIAS_GATE_RES, gate=e[5], rows_per_gate=fs,
                gate_row0=r0, res=x, out=x)
        b["e2"][:, :C_].copy_(b["e"]); b["e2"][:, C_:].copy_(b["e"])
        ops.modulation_table(Pk["head_mod"], b["e2"], out=b["hmod"])
        hm = b["hmod"][0]
        ops.ln_modulate(x, shift=hm[:, :C_], scale=hm[:, C_:], rows_per_frame=fs, row0=r0, eps=eps, out=xm)
        ops.gemm(xm, Pk["head_w"], Pk["head_b"], out=b["y"][r0:r0 + Lp])
        self._mark("end")
        # gather the [L, 64] head output (tiny) so every rank can unpatchify the full chunk
        dist.all_gather_into_tensor(b["y"][:L], b["y"][r0:r0 + Lp].clone(), group=self.group)
        ops.unpatchify(b["y"][:L], self.out_dim, F, H, W, out=b["out"][0])

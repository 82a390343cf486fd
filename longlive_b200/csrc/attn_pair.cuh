// llb_attn_fwd, CTA-pair kernel (tcgen05 cta_group::2) — included by attn_fwd.cu (shares KvTileIter, SegIter, the
// exp2 helpers and the split-remainder workspace layout with the single-CTA kernel above).
//
// Why a second design.  ncu on the single-CTA kernel (profiles/r02_attn_ncu_source.md): each of its two Q tiles runs the
// serial chain  S ready -> softmax (B ~ 1400-1850 cycles, issue bound) -> P ready -> PV + QK (1024 tensor cycles) -> S
// ready, with 750 cycles of hand-over latency, so the tensor pipe is busy 2 x 1024 of ~3500 cycles (55-59 %).  A third
// chain does not fit the 512 TMEM columns, and with P aliased onto S the next QK cannot be issued before PV has
// consumed P.  This kernel changes the decomposition instead of the constants:
//
//   * a CLUSTER OF TWO CTAs (one TPC) owns 256 query rows of one head, 128 rows per CTA; every MMA is a
//     cta_group::2 instruction with M = 256 issued by the leader CTA.  The B operand of such an MMA is split along N
//     between the two CTAs, so each SM loads and reads only HALF of every K tile (64 of the 128 keys) and half of
//     every V tile (64 of the 128 head-dim columns): per-SM shared-memory and L2 traffic per flop stay what they are in
//     the single-CTA kernel although every SM now works on ONE Q tile;
//   * with one Q tile per SM the 512 TMEM columns hold  S_A | S_B | O | P_A | P_B  (128 + 128 + 128 + 64 + 64): two S
//     buffers, and P in columns of its own.  Key tiles alternate between chain A and chain B.  QK(j+2) is issued as
//     soon as the softmax warpgroup has READ S(j) into registers - it no longer waits for PV(j) - so S(j+1) and
//     S(j+2) are always ready and the two softmax warpgroups (chain A: warps 0-3, chain B: warps 4-7, both on the SAME
//     128 rows, different key tiles) run back to back, two warps per scheduler instead of one;
//   * both chains accumulate into the one O.  The exponent offset (running row maximum, lazily updated) is therefore
//     shared: tile decisions are taken in tile order through a per-quadrant sequence number in shared memory; the
//     (rare) O rescale waits for the previous tile's PV, and each warpgroup keeps its own partial row sum, re-based
//     when it sees that the shared maximum moved.
//
// Scheduling is the single-CTA kernel's (persistent whole-item rounds + split remainder merged through the
// workspace), with the cluster as the worker.
//
// STATUS (round 2, measured on B200, profiles/r02_attn_pair.md): bit-for-bit the same error against fp32 as the
// single-CTA kernel on every test shape (tests/test_attn_gpu.py, variant 64), but NOT faster yet - 1170-1200 TFLOP/s
// against 1230-1240 at Lq 4680 x Lk 18720 - so it is selectable (variant bit 6) and not the default.  Its own ncu
// profile shows the decomposition working as intended (the softmax warpgroups wait for S 4 % of the time instead
// of 49 %) and what it costs: two softmax warps per scheduler issue at 30 % each instead of 41 % alone (MIO / XU
// queue throttling appears), and the ordering protocol adds ~330 cycles per tile.
#pragma once

namespace llb {

constexpr int kPairStages = 10;             // 16 KB half tiles (K: 64 keys x 128 d, V: 128 keys x 64 d) in flight
constexpr int kHalfTileBytes = 64 * 128 * 2;
constexpr int kPairBarBytes = 512;
constexpr int kPairXchgBytes = 128 * 4 + 2 * 128 * 4 + 64;  // m_ref[128], l_x[2][128], dec[4]
constexpr int kPairSmemBytes = 1024 + kTileBytes + kPairStages * kHalfTileBytes + kPairBarBytes + kPairXchgBytes;
constexpr int kPairPoly = 4;
constexpr int kQkAhead = 4;                 // QK(j + 4) is issued right behind PV(j): see the MMA warp

__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_shared_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared_u32(uint32_t saddr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
// bounded spin until the shared-memory word equals `want` (a protocol bug traps instead of hanging the box)
__device__ __forceinline__ void spin_shared_eq(uint32_t saddr, uint32_t want) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (ld_acquire_shared_u32(saddr) != want) {
    if ((++spins & 0xfffu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > LLB_WAIT_TIMEOUT_NS) __trap();
    }
  }
}

__global__ void __maxnreg__(168)
attn_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  constexpr int kMmaWarp = 8, kTmaWarp = 9;
  constexpr int kS = kPairStages;
  extern __shared__ uint8_t smem_raw[];
  // identical offsets in both CTAs of the pair (cta_group::2 MMAs and multicast commits rely on that)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_base = smem_base;                       // this CTA's 128 query rows: two [128 x 64] boxes
  const uint32_t kv_base = q_base + kTileBytes;            // kS half tiles
  const uint32_t bar_base = kv_base + kS * kHalfTileBytes;
  uint8_t* bar_gen = smem_gen + (bar_base - smem_base);
  const uint32_t qfull_bar = bar_base, qempty_bar = bar_base + 8;
  auto sfull_bar = [&](int c) { return bar_base + 8u * (2 + c); };   // S_c(j) complete            (commit, both CTAs)
  auto sfree_bar = [&](int c) { return bar_base + 8u * (4 + c); };   // S_c(j) is in registers     (leader, 8 arrivals)
  auto pfull_bar = [&](int c) { return bar_base + 8u * (6 + c); };   // P_c(j) is in TMEM          (leader, 8 arrivals)
  auto pfree_bar = [&](int c) { return bar_base + 8u * (8 + c); };   // PV(j) of chain c complete  (commit, both CTAs)
  const uint32_t ofree_bar = bar_base + 8u * 10;                     // O drained                  (leader, 16 arrivals)
  auto kvfull_bar = [&](int s) { return bar_base + 8u * (11 + s); };        // leader
  auto kvempty_bar = [&](int s) { return bar_base + 8u * (11 + kS + s); };  // commit, both CTAs
  const uint32_t tmem_slot = bar_base + 8u * (11 + 2 * kS);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(bar_gen + 8 * (11 + 2 * kS));
  static_assert(8 * (12 + 2 * kS) <= kPairBarBytes, "barrier block");
  float* m_ref = reinterpret_cast<float*>(bar_gen + kPairBarBytes);  // shared exponent offset per row
  float* l_x = m_ref + 128;                                           // [2][128] partial row sums at item end
  const uint32_t dec_base = bar_base + kPairBarBytes + 3 * 128 * 4;   // dec[4]: tile decisions taken, per quadrant

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int worker = blockIdx.x >> 1;
  const int n_workers = gridDim.x >> 1;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(qfull_bar, 1);
    mbar_init(qempty_bar, 1);
    for (int c = 0; c < 2; ++c) {
      mbar_init(sfull_bar(c), 1);
      mbar_init(sfree_bar(c), 8);   // 4 softmax warps of chain c in each CTA
      mbar_init(pfull_bar(c), 8);
      mbar_init(pfree_bar(c), 1);
    }
    mbar_init(ofree_bar, 16);       // 8 softmax warps in each CTA
    for (int s = 0; s < kS; ++s) {
      mbar_init(kvfull_bar(s), 1);
      mbar_init(kvempty_bar(s), 1);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 128) m_ref[threadIdx.x] = -INFINITY;
  if (threadIdx.x < 4) reinterpret_cast<volatile uint32_t*>(bar_gen + kPairBarBytes + 3 * 128 * 4)[threadIdx.x] = 0u;
  if (warp == kMmaWarp) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers must exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  griddep_wait();
  griddep_launch_dependents();

#define LLB_PAIR_INIT_WORK()                                             \
  KvTileIter kv_it;                                                      \
  kv_it.init(p.segs);                                                    \
  SegIter sg;                                                            \
  sg.init(kv_it.total_tiles(), p.n_heads * p.n_pairs, n_workers, worker, p.min_split_tiles)

  if (warp >= kMmaWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == kTmaWarp) {
      // ------------------------------------------------------------------ TMA producer (both CTAs)
      if (lane == 0) {
        LLB_PAIR_INIT_WORK();
        const uint32_t lead_qfull = mapa_shared(qfull_bar, 0);
        int stage = 0;
        uint32_t phase = 0, qph = 0;
        for (; sg.ok; sg.next()) {
          const int head = sg.item / p.n_pairs;
          const int q_row0 = (sg.item - head * p.n_pairs) * 256 + static_cast<int>(rank) * 128;
          const int col = head * 128;
          // the Q slot is free once the MMA warp has issued the previous item's last QK (rows beyond Lq are
          // zero-filled by the TMA unit: the second CTA of a ragged last pair computes on zeros and stores nothing)
          mbar_wait(qempty_bar, qph ^ 1);
          qph ^= 1;
          if (rank == 0) mbar_arrive_expect_tx(qfull_bar, 2 * kTileBytes);  // both CTAs' bytes land on the leader
          tma_load_2d_pair(q_base, &tmap_q, lead_qfull, col, q_row0);
          tma_load_2d_pair(q_base + kBoxBytes, &tmap_q, lead_qfull, col + 64, q_row0);
          auto begin_stage = [&]() {
            mbar_wait(kvempty_bar(stage), phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(kvfull_bar(stage), 2 * kHalfTileBytes);
          };
          auto end_stage = [&]() { if (++stage == kS) { stage = 0; phase ^= 1; } };
          // K tile: this CTA supplies keys [rank*64, rank*64+64) (B operand of QK is split along N = keys)
          auto load_k = [&](int row0) {
            begin_stage();
            const uint32_t dst = kv_base + stage * kHalfTileBytes;
            const uint32_t lead = mapa_shared(kvfull_bar(stage), 0);
            tma_load_2d_pair(dst, &tmap_k, lead, col, row0 + static_cast<int>(rank) * 64);
            tma_load_2d_pair(dst + kHalfTileBytes / 2, &tmap_k, lead, col + 64, row0 + static_cast<int>(rank) * 64);
            end_stage();
          };
          // V tile: this CTA supplies head-dim columns [rank*64, rank*64+64) (B operand of PV is split along N = d)
          auto load_v = [&](int row0) {
            begin_stage();
            const uint32_t dst = kv_base + stage * kHalfTileBytes;
            tma_load_2d_pair(dst, &tmap_v, mapa_shared(kvfull_bar(stage), 0), col + static_cast<int>(rank) * 64, row0);
            end_stage();
          };
          // consumption order of the MMA warp: K_0 .. K_3, then V_j, K_{j+4} for j = 0 ..
          const int nt = sg.t1 - sg.t0;
          kv_it.seek(sg.t0);
          KvTileIter k_it = kv_it;
          int row0, valid;
          for (int i = 0; i < kQkAhead && i < nt; ++i) {
            k_it.get(row0, valid);
            load_k(row0);
            k_it.next();
          }
          for (int j = 0; j < nt; ++j) {
            kv_it.get(row0, valid);
            load_v(row0);
            kv_it.next();
            if (j + kQkAhead < nt) {
              k_it.get(row0, valid);
              load_k(row0);
              k_it.next();
            }
          }
        }
      }
    } else if (warp == kMmaWarp && rank == 0) {
      // ------------------------------------------------------------------ MMA issuer (leader CTA only)
      constexpr uint32_t idesc_qk = umma_idesc_bf16(256, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(256, 128, 0, 1);
      // S_c = Q K^T: A = each CTA's own 128 query rows (two 64-wide d boxes), B = each CTA's 64 keys (two 64-wide d
      // boxes of 64 rows); 16 d per instruction
      auto issue_qk = [&](int c, uint32_t kst) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t oa = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
          const uint32_t ob = (kk >> 2) * (kHalfTileBytes / 2) + (kk & 3) * 32;
          umma_ss_pair(tmem_base + c * 128, umma_desc_kmajor(q_base + oa), umma_desc_kmajor(kst + ob), idesc_qk,
                       kk != 0);
        }
      };
      // O += P_c V: A = P_c (bf16, 8 TMEM columns per 16 keys, each CTA its own 128 rows), B = each CTA's 64 head-dim
      // columns of V (one [128 keys x 64 d] box, MN-major: 16 keys per instruction are two 8-row groups)
      auto issue_pv = [&](int c, uint32_t vst, bool first) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t bdesc = umma_desc_mnmajor(vst + kk * 2048, kHalfTileBytes);
          umma_ts_pair(tmem_base + 256, tmem_base + 384 + c * 64 + kk * 8, bdesc, idesc_pv,
                       (first && kk == 0) ? 0u : 1u);
        }
      };
      LLB_PAIR_INIT_WORK();
      int stage = 0;
      uint32_t phase = 0, qph = 0, oseg = 0;
      uint32_t qk_cnt[2] = {0, 0};  // QKs issued per chain (s_free phase)
      uint32_t pv_cnt[2] = {0, 0};  // PVs issued per chain (p_full phase)
      auto advance = [&]() { if (++stage == kS) { stage = 0; phase ^= 1; } };
      for (; sg.ok; sg.next()) {
        const int nt = sg.t1 - sg.t0;
        mbar_wait(qfull_bar, qph);
        qph ^= 1;
        // S_c(jq) = Q K_jq^T once S_c's previous content has been read by its softmax warpgroup
        auto qk_step = [&](int jq) {
          const int c = jq & 1;
          const int kstage = stage;
          mbar_wait(kvfull_bar(stage), phase);
          const uint32_t kst = kv_base + stage * kHalfTileBytes;
          advance();
          if (qk_cnt[c] > 0) mbar_wait(sfree_bar(c), (qk_cnt[c] - 1) & 1);
          qk_cnt[c]++;
          tc_fence_after();
          if (elect_one()) {
            issue_qk(c, kst);
            umma_commit_pair(sfull_bar(c), 3);
            umma_commit_pair(kvempty_bar(kstage), 3);
            if (jq == nt - 1) umma_commit_pair(qempty_bar, 3);  // the item's last QK: the Q slot can be refilled
          }
          __syncwarp();
        };
        // Issue order = the order in which things become ready.  When a softmax warpgroup finishes tile j and starts
        // tile j+2, two MMA groups become issuable together: PV(j) (P_c(j) is complete) and QK(j+4) (S_c(j+2) has just
        // been read into registers, so its buffer is free).  Hence QKs run kQkAhead = 4 tiles ahead of the PVs:
        // QK(0..3) up front (QK(2), QK(3) wait for the first reads of S_A, S_B), then PV(j), QK(j+4).  With QK(j+2)
        // queued behind PV(j-1) instead, S(j+2) only became ready half a period late (first ncu profile of this kernel).
        for (int i = 0; i < kQkAhead && i < nt; ++i) qk_step(i);
        // O of the previous item must have been drained before this item's first PV overwrites it
        mbar_wait(ofree_bar, (oseg & 1) ^ 1);
        oseg++;
        for (int j = 0; j < nt; ++j) {
          const int c = j & 1;
          const int vstage = stage;
          mbar_wait(kvfull_bar(stage), phase);
          const uint32_t vst = kv_base + stage * kHalfTileBytes;
          advance();
          mbar_wait(pfull_bar(c), pv_cnt[c] & 1);
          pv_cnt[c]++;
          tc_fence_after();
          if (elect_one()) {
            issue_pv(c, vst, j == 0);
            umma_commit_pair(pfree_bar(c), 3);
            umma_commit_pair(kvempty_bar(vstage), 3);
          }
          __syncwarp();
          if (j + kQkAhead < nt) qk_step(j + kQkAhead);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // -------------------------------------------------------------------- softmax warps
    const int ch = warp >> 2;  // chain: key tiles jj = ch, ch + 2, ... of every item (jj counted from the item's first tile)
    const int q = warp & 3;    // TMEM lane quadrant; warps (0,q) and (1,q) own the same 32 rows
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_s = tmem_base + lane_off + ch * 128;
    const uint32_t t_p = tmem_base + lane_off + 384 + ch * 64;
    const uint32_t t_o = tmem_base + lane_off + 256;
    const uint32_t dec_addr = dec_base + 4u * q;
    const uint32_t lead_sfree = mapa_shared(sfree_bar(ch), 0);
    const uint32_t lead_pfull = mapa_shared(pfull_bar(ch), 0);
    const uint32_t lead_ofree = mapa_shared(ofree_bar, 0);
    const float c = p.scale_log2;
    uint32_t n_mine = 0, n_other = 0;  // key tiles handled so far by this chain / the other chain (barrier phases)
    LLB_PAIR_INIT_WORK();

    // wait_p: parity to wait for on p_free[ch] before the FIRST store of a tile (-1 = nothing to wait for)
    auto exp_chunk = [&](const uint32_t (&s)[32], int cc, float2 c2, float2 neg2, float2& la, float2& lb, int wait_p) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 tt = __ffma2_rn(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), c2, neg2);
        float2 pp;
        if ((i % kPairPoly) == kPairPoly - 1) {
          pp = exp2_poly2(tt);
        } else {
          pp.x = ex2_approx(tt.x);
          pp.y = ex2_approx(tt.y);
        }
        if (i & 1) lb = __fadd2_rn(lb, pp);
        else la = __fadd2_rn(la, pp);
        pk[i] = pack_bf16x2(pp.x, pp.y);
      }
      if (cc == 0 && wait_p >= 0) {
        // P_c still feeds PV of this chain's previous tile until that MMA group completes; the first 32 keys of
        // this tile are exponentiated before the wait, which hides it
        mbar_wait(pfree_bar(ch), static_cast<uint32_t>(wait_p));
        tc_fence_after();
      }
      tmem_st16(t_p + cc * 16, pk);
    };
    auto rescale_o = [&](float f) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t ov[32];
        tmem_ld32(t_o + cc * 32, ov);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
        tmem_st32(t_o + cc * 32, ov);
      }
      tmem_wait_st();
    };

    for (; sg.ok; sg.next()) {
      const int head = sg.item / p.n_pairs;
      const int q_row0 = (sg.item - head * p.n_pairs) * 256 + static_cast<int>(rank) * 128;
      const int grow = q_row0 + row;
      const int nt = sg.t1 - sg.t0;
      const int n_mine_item = (nt - ch + 1) >> 1;   // tiles of this chain in this item
      const int n_other_item = nt - n_mine_item;
      float m_seen = -INFINITY;  // the shared offset as this warpgroup last saw it (its l is relative to it)
      float l = 0.f;             // sum of this chain's probabilities
      kv_it.seek(sg.t0 + ch);
      for (int jj = ch; jj < nt; jj += 2) {
        int row0, valid;
        kv_it.get(row0, valid);
        kv_it.next();
        kv_it.next();
        mbar_wait(sfull_bar(ch), n_mine & 1);
        tc_fence_after();
        auto tile = [&](auto masked_tag) {
          constexpr bool kMasked = decltype(masked_tag)::value;
          uint32_t sv[4][32];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) tmem_ld32(t_s + cc * 32, sv[cc]);
          tmem_wait_ld();
          // S_c is in registers: the MMA warp may overwrite it with S_c(jj + 2) right away
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lead_sfree);
          if constexpr (kMasked) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (cc * 32 + i >= valid) sv[cc][i] = 0xff800000u;  // -inf
          }
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            mx0 = fmaxf(mx0, __uint_as_float(sv[0][i]));
            mx1 = fmaxf(mx1, __uint_as_float(sv[1][i]));
            mx2 = fmaxf(mx2, __uint_as_float(sv[2][i]));
            mx3 = fmaxf(mx3, __uint_as_float(sv[3][i]));
          }
          const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
          // ---- ordered decision: tile jj decides after tile jj-1 (taken by the other chain's warp of this quadrant)
          spin_shared_eq(dec_addr, static_cast<uint32_t>(jj));
          const float m_cur = m_ref[row];
          if (m_cur != m_seen) {  // the other chain moved the offset since this chain's last tile: re-base l
            l *= ex2_approx((m_seen - m_cur) * c);  // m_seen = -inf only while l == 0
            m_seen = m_cur;
          }
          const float m_new = fmaxf(m_cur, mx);
          const bool need = (m_new - m_cur) * c > 8.0f;  // lazy: only when the maximum grew by more than 2^8
          if (__any_sync(0xffffffffu, need)) {
            const float f = ex2_approx((m_cur - m_new) * c);  // 0 on the item's first tile (m_cur = -inf)
            if (jj > 0) {
              // O must be stable: PV(jj-1) (other chain) complete, and PV(jj) cannot be issued before this P exists
              mbar_wait(pfree_bar(ch ^ 1), (n_other + ((jj - 1) >> 1)) & 1);
              tc_fence_after();
              rescale_o(f);
              tc_fence_before();
            }
            l *= f;
            m_ref[row] = m_new;
            m_seen = m_new;
          }
          __syncwarp();
          if (lane == 0) st_release_shared_u32(dec_addr, static_cast<uint32_t>(jj + 1));
          const float neg = -m_seen * c;
          const float2 c2 = make_float2(c, c), neg2 = make_float2(neg, neg);
          float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
          const int wait_p = n_mine > 0 ? static_cast<int>((n_mine - 1) & 1) : -1;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) exp_chunk(sv[cc], cc, c2, neg2, la, lb, wait_p);
          la = __fadd2_rn(la, lb);
          l += la.x + la.y;
        };
        if (valid < 128) tile(std::true_type{});
        else tile(std::false_type{});
        n_mine++;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_pfull);
      }
      n_other += n_other_item;

      // ---- item epilogue: every decision of the item has been taken -> final offset; last PV complete -> O final
      spin_shared_eq(dec_addr, static_cast<uint32_t>(nt));
      float m_used = m_ref[row];
      if (m_used != m_seen) l *= ex2_approx((m_seen - m_used) * c);
      if (n_mine_item == 0) l = 0.f;  // a one-tile item: chain 1 saw nothing (m_seen = -inf: keep NaN out)
      {
        const int cl = (nt - 1) & 1;  // chain of the item's last tile; its PV completes last (in-order commits)
        const uint32_t done = (cl == ch) ? n_mine : n_other;
        mbar_wait(pfree_bar(cl), (done - 1) & 1);
        tc_fence_after();
      }
      l_x[ch * 128 + row] = l;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      l += l_x[(ch ^ 1) * 128 + row];
      asm volatile("bar.sync 1, 256;" ::: "memory");  // both chains have read m_ref / dec / l_x of this item
      if (ch == 0) {  // reset the protocol state for the next item (chain 1 cannot pass the barrier below before this)
        m_ref[row] = -INFINITY;
        if (lane == 0) st_release_shared_u32(dec_addr, 0u);
      }
      const bool tail_part = sg.in_remainder() && sg.t0 > 0;     // earlier kv tiles live in the previous cluster
      const bool head_part = sg.in_remainder() && sg.t1 < sg.T;  // later kv tiles live in the next cluster
      const bool row_ok = grow < p.Lq;
      float a_own = 1.0f, a_oth = 0.0f;
      const float4* wo_in = nullptr;
      uint32_t* flag_in = nullptr;
      if (head_part && row_ok) {
        uint8_t* wsb = p.workspace + static_cast<int64_t>(blockIdx.x + 2) * kWsPerCta;  // same rank, next cluster
        flag_in = reinterpret_cast<uint32_t*>(wsb + kWsOBytes + kWsMlBytes) + row;
        uint32_t spins = 0;
        uint64_t t_start = 0;
        while (ld_acquire_u32(flag_in) == 0u) {
          if ((++spins & 0xfffu) == 0) {
            const uint64_t now = global_timer_ns();
            if (t_start == 0) t_start = now;
            else if (now - t_start > LLB_WAIT_TIMEOUT_NS) __trap();
          }
        }
        const volatile float* mlp = reinterpret_cast<const volatile float*>(wsb + kWsOBytes) + 2 * row;
        const float m_oth = mlp[0], l_oth = mlp[1];
        const float m = fmaxf(m_used, m_oth);
        a_own = ex2_approx((m_used - m) * c);
        a_oth = ex2_approx((m_oth - m) * c);
        l = l * a_own + l_oth * a_oth;
        m_used = m;
        wo_in = reinterpret_cast<const float4*>(wsb) + row;
      }
      uint8_t* wsb_out = p.workspace + static_cast<int64_t>(blockIdx.x) * kWsPerCta;
      float4* wo_out = reinterpret_cast<float4*>(wsb_out) + row;
      if (!tail_part) {
        const float inv = 1.0f / l;
        a_own *= inv;
        a_oth *= inv;
      }
      __nv_bfloat16* orow = p.out + static_cast<int64_t>(grow) * p.ldo + head * 128;
      if (p.shard.n_ranks > 1 && row_ok) {
        const int r = grow / p.shard.rows_per_rank;
        orow = static_cast<__nv_bfloat16*>(p.shard.out_peers[r]) +
               static_cast<int64_t>(grow - r * p.shard.rows_per_rank) * p.shard.ld_out + p.shard.head_col0 +
               head * p.shard.head_col_stride;
      }
      // chain c drains output columns [c*64, c*64 + 64)
#pragma unroll
      for (int c2i = 0; c2i < 2; ++c2i) {
        const int cc = ch * 2 + c2i;
        uint32_t ov[32];
        tmem_ld32(t_o + cc * 32, ov);
        tmem_wait_ld();
        if (row_ok) {
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(ov[i]) * a_own;
          if (wo_in != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 x = ld_cg_f4(wo_in + (cc * 8 + i) * 128);
              o[4 * i] += x.x * a_oth;
              o[4 * i + 1] += x.y * a_oth;
              o[4 * i + 2] += x.z * a_oth;
              o[4 * i + 3] += x.w * a_oth;
            }
          }
          if (tail_part) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              wo_out[(cc * 8 + i) * 128] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 w;
              w.x = pack_bf16x2(o[8 * i + 0], o[8 * i + 1]);
              w.y = pack_bf16x2(o[8 * i + 2], o[8 * i + 3]);
              w.z = pack_bf16x2(o[8 * i + 4], o[8 * i + 5]);
              w.w = pack_bf16x2(o[8 * i + 6], o[8 * i + 7]);
              *reinterpret_cast<uint4*>(orow + cc * 32 + i * 8) = w;
            }
          }
        }
      }
      // the merged partial is consumed and our own published by the chain-0 thread of the row, after BOTH chains are
      // done with the workspace rows (their writes fenced, then the barrier)
      if (tail_part) __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ch == 0) {
        if (flag_in != nullptr) st_release_u32(flag_in, 0u);
        if (tail_part && row_ok) {
          reinterpret_cast<float2*>(wsb_out + kWsOBytes)[row] = make_float2(m_used, l);
          __threadfence();
          st_release_u32(reinterpret_cast<uint32_t*>(wsb_out + kWsOBytes + kWsMlBytes) + row, 1u);
        }
      }
      // O drained: the leader's MMA warp may start the next item's accumulation
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_ofree);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

static int launch_attn_pair(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                            int sms, cudaStream_t stream) {
  LLB_SET_MAX_SMEM(attn_pair_kernel, kPairSmemBytes);
  const int grid = (sms / 2) * 2;
  LLB_CUDA(launch_ex(attn_pair_kernel, dim3(grid), dim3(kAttnThreads), kPairSmemBytes, stream, 2, false, tq, tk, tv, p));
  LLB_LAUNCH_CHECK("attn_pair_kernel");
  return LLB_OK;
}

}  // namespace llb

// llb_attn_fwd — dense flash-style attention for sm_100a (tcgen05 + TMEM + TMA), head_dim 128.
//
//   out[Lq, H*128] = softmax(Q K^T * scale) V        per head, bf16 in/out, fp32 S/O in TMEM
//
// Replaces flash_attn_varlen_func as reached from wan/modules/attention.py:116-145 for
//   * self-attention over the rolling KV cache (wan/modules/causal_model.py:331-360): the keys are
//     "sink ++ local window" of the cache.  The reference clones the cache, memmoves the window
//     and torch.cat's sink+window before every call; here K/V tiles are TMA-loaded IN PLACE from
//     the ring buffer: the attended set is a short list of physical row ranges (softmax attention
//     does not depend on key order), given in device memory so a captured CUDA graph can be
//     replayed for every chunk;
//   * cross-attention over the cached 512 text keys (wan/modules/model.py:189).
//
// CTA = one head x 256 query rows = two 128-row Q tiles that ping-pong on the tensor core:
//   warps 0-3   softmax for Q tile 0   (thread == S/O row; TMEM lane quadrant = warp % 4)
//   warps 4-7   softmax for Q tile 1
//   warp  8     MMA issuer (one thread): S_t = Q_t K_j^T, O_t += P_t V_j via tcgen05.mma
//   warp  9     TMA producer: Q tiles once, then K_0,V_0,K_1,V_1,... through a shared-memory ring
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each.  P (bf16) overwrites the first 64
// columns of its S tile and is consumed straight from TMEM as the A operand of the PV MMA.
// (Round-1 experiments that measured slower - P staged through shared memory with an early-QK schedule,
// a 64-key half-tile pipeline, two softmax warpgroups per Q tile, MUFU-only exp2 - and the compile-time
// timing instrumentation were removed from the shipped kernel; DESIGN.md 4.1 keeps their numbers, the
// history keeps their code.)
// While the softmax warps of one Q tile work on S_t(j), the tensor core runs the other tile's
// PV(j-1) and QK(j), so exp2/max/sum overlap with MMA issue.
// Online softmax uses lazy rescaling: O is only rescaled (in TMEM, by the softmax warps) when the
// running row max grows by more than 2^8, so the common path never touches O.
//
// Scheduling (persistent; whole-item rounds + split remainder): 12 heads x 19 Q pairs = 228 items
// do not divide over 148 SMs (1.54 waves).  The grid is one CTA per SM.  Every CTA first processes
// n_items / G whole items in lockstep rounds (CTAs sharing a head stream the same K/V tiles out of
// L2 at the same time); the n_items % G left-over items are then cut into equal contiguous ranges
// of the flattened (item, kv-tile) space, one per CTA.  A part that does not contain kv tile 0
// writes its un-normalised O (fp32) and (m, l) per row to a workspace and raises a per-row flag;
// the CTA owning the preceding kv range of the same item (always the previous CTA) waits for it,
// merges, and either passes the merged partial on or - if it owns kv tile 0 - normalises and
// writes the output.  The launch is cooperative (all CTAs co-resident: grid <= #SMs, 1 CTA/SM), so the flag wait cannot
// deadlock; flags are consumed (reset to 0) by their reader, which keeps the kernel replayable
// inside a CUDA graph.
#include "llb_common.cuh"
#include "llb_host.h"

#include <stdlib.h>
#include <string.h>

#include <type_traits>


namespace llb {

// LLB_ATTN_TRACE (debug builds only, tools/attn_trace.py): CTA 0 records clock64() at the hand-over points of the first
// 64 key tiles of its first item: g_attn_trace[role][tile][slot], role 0 / 1 = softmax chain 0 / 1, 2 = the MMA issuer.
#ifdef LLB_ATTN_TRACE
__device__ unsigned long long g_attn_trace[4 * 64 * 16];
#define LLB_TRACE(role, tile, slot)                                                                              \
  do {                                                                                                           \
    if (blockIdx.x == 0 && lane == 0 && trace_on && (tile) < 64) g_attn_trace[((role) * 64 + (tile)) * 16 + (slot)] = clock64(); \
  } while (0)
#else
#define LLB_TRACE(role, tile, slot) do {} while (0)
#endif

constexpr int kAttnThreads = 384;  // 3 warpgroups: softmax0, softmax1, {MMA, TMA, 2 idle warps}
constexpr int kTileBytes = 128 * 128 * 2;  // one [128 x 128] bf16 operand tile (two SW128 boxes)
constexpr int kBoxBytes = 128 * 64 * 2;
constexpr int kStages = 4;          // K/V tiles in flight (the fifth tile of shared memory stages the output, below)
constexpr int kOutStageBytes = 128 * 64 * 2;  // per Q tile: 128 rows x 64 output columns, SWIZZLE_128B, for the TMA store
constexpr int kAttnSmemBytes = 1024 + 2 * kTileBytes + kStages * kTileBytes + 2 * kOutStageBytes + 256;

// workspace per CTA: partial O [2 tiles][32 col4][128 rows] float4, (m,l) [2][128] float2, flags [2][128]
constexpr int64_t kWsOBytes = 2ll * 32 * 128 * 16;
constexpr int64_t kWsMlBytes = 2ll * 128 * 8;
constexpr int64_t kWsFlagBytes = 2ll * 128 * 4;
constexpr int64_t kWsPerCta = kWsOBytes + kWsMlBytes + kWsFlagBytes;

struct AttnParams {
  __nv_bfloat16* out;
  int64_t ldo;
  int Lq;
  int n_heads;
  int n_pairs;  // ceil(Lq / 256)
  float scale_log2;
  const llb_step_params* segs;
  uint8_t* workspace;  // gridDim.x * kWsPerCta bytes, flags zero-initialised once
  int min_split_tiles;  // the remainder items are split over CTAs only when an item has at least this many kv tiles
  llb_out_shard shard; // n_ranks == 1: single GPU
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA/ALU pipes for a pair of values (FA4-style MUFU offload): t = j + r with j = round(t)
// taken from the low mantissa bits of t + 1.5*2^23, 2^r by a degree-3 polynomial on [-0.5, 0.5]
// (max relative error 1.6e-4, far below the bf16 rounding of P), and 2^j added into the exponent.
__device__ __forceinline__ float2 exp2_poly2(float2 t) {
  t.x = fmaxf(t.x, -125.0f);
  t.y = fmaxf(t.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 u = __fadd2_rn(t, magic);
  const float2 j = __fadd2_rn(u, make_float2(-12582912.0f, -12582912.0f));
  const float2 r = __ffma2_rn(j, make_float2(-1.0f, -1.0f), t);
  float2 q = __ffma2_rn(r, make_float2(0.05676588788628578f, 0.05676588788628578f),
                        make_float2(0.24273726344108582f, 0.24273726344108582f));
  q = __ffma2_rn(q, r, make_float2(0.6929193139076233f, 0.6929193139076233f));
  q = __ffma2_rn(q, r, make_float2(0.9999317526817322f, 0.9999317526817322f));
  q.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(u.x) << 23));
  q.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(u.y) << 23));
  return q;
}

__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// Walks the attended key tiles: segments in order, 128-row tiles inside each segment.  Only
// scalars are kept; the (tiny) segment table is re-read from global memory on a segment switch.
struct KvTileIter {
  const llb_step_params* sp;
  int nseg, seg, off, cur_start, cur_len;
  __device__ __forceinline__ void load_seg() {
    cur_len = 0;
    while (seg < nseg) {
      cur_start = sp->attn_start[seg];
      cur_len = sp->attn_len[seg];
      if (cur_len > 0) break;
      ++seg;
    }
  }
  __device__ __forceinline__ void init(const llb_step_params* sp_) {
    sp = sp_;
    nseg = min(sp->n_attn_segs, LLB_MAX_SEGS);
    seg = 0;
    off = 0;
    cur_start = 0;
    load_seg();
  }
  __device__ __forceinline__ int total_tiles() const {
    int n = 0;
    for (int i = 0; i < nseg; ++i) {
      const int l = sp->attn_len[i];
      if (l > 0) n += (l + 127) >> 7;
    }
    return n;
  }
  // position on kv tile index t (0-based over the whole attended set)
  __device__ __forceinline__ void seek(int t) {
    seg = 0;
    off = 0;
    load_seg();
    while (seg < nseg) {
      const int nt = (cur_len + 127) >> 7;
      if (t < nt) { off = t << 7; return; }
      t -= nt;
      ++seg;
      load_seg();
    }
  }
  // current tile: first key row and number of valid keys (1..128)
  __device__ __forceinline__ void get(int& row0, int& valid) const {
    row0 = cur_start + off;
    valid = min(128, cur_len - off);
  }
  __device__ __forceinline__ void next() {
    off += 128;
    if (off >= cur_len) {
      off = 0;
      ++seg;
      load_seg();
    }
  }
};

// Per-CTA list of work segments (item, kv tiles [t0, t1)).
//   rounds:    items k*G + c for k < n_items / G are processed whole; all CTAs walk the kv tiles in
//              lockstep, so CTAs that share a head hit the same K/V tiles in L2 at the same time
//   remainder: the last n_items % G items are cut into equal contiguous ranges of the flattened
//              (item, kv tile) space, one range per CTA (at most kMaxParts parts per item); parts of
//              one item live in consecutive CTAs and are merged head <- ... <- tail through the workspace
constexpr int kMaxParts = 4;
struct SegIter {
  int T, G, c, rounds, rem_items, k;
  bool split;
  int ru, ru_end;
  int item, t0, t1;
  bool ok;
  __device__ __forceinline__ void init(int T_, int n_items, int G_, int c_, int min_split) {
    T = T_; G = G_; c = c_;
    rounds = n_items / G;
    rem_items = n_items - rounds * G;
    k = 0;
    split = (T >= min_split) && rem_items > 0;
    ru = ru_end = 0;
    if (split) {
      const int gp = min(G, rem_items * kMaxParts);  // CTAs that take part in the remainder
      if (c < gp) {
        const long long R = static_cast<long long>(rem_items) * T;
        ru = static_cast<int>(R * c / gp);
        ru_end = static_cast<int>(R * (c + 1) / gp);
      }
    }
    load();
  }
  __device__ __forceinline__ void load() {
    ok = false;
    if (T <= 0) return;
    if (k < rounds) {
      item = k * G + c; t0 = 0; t1 = T; ok = true;
    } else if (split) {
      if (ru < ru_end) {
        const int ri = ru / T;
        item = rounds * G + ri;
        t0 = ru - ri * T;
        const int rem = ru_end - ru;
        t1 = (rem < T - t0) ? t0 + rem : T;
        ok = true;
      }
    } else if (k == rounds && c < rem_items) {
      item = rounds * G + c; t0 = 0; t1 = T; ok = true;
    }
  }
  __device__ __forceinline__ bool in_remainder() const { return k >= rounds && split; }
  __device__ __forceinline__ void next() {
    if (k < rounds || !split) ++k;
    else ru += t1 - t0;
    load();
  }
};

// kPoly: every kPoly-th pair of probabilities is exponentiated on the FMA pipe (exp2_poly2) instead of MUFU
// (0 = MUFU only; shipped: 4 - MUFU only and every 8th pair measured the same within noise, profiles/r02_attn_pair.md).  ncu (profiles/r02_attn_ncu_source.md): the softmax warps are ISSUE bound - one warp per
// scheduler, 40 % of its cycles issuing and 37 % in fixed-latency waits, XU (MUFU) pipe 44 % busy - so the
// polynomial (12 instructions per pair against 2) is only worth what MUFU time it removes from the chain.
// (An "early-QK" schedule - every key tile as two 64-key halves so that QK_b(j+1) can be issued while P(j) is still being
// written, one MMA issuer warp per chain - was built in round 2 and measured equal: N = 64 MMAs cost 48 cycles instead of
// 32 because the Q operand is re-read from shared memory; profiles/r02_attn_early_qk.md, commit 49c72de.)
template <int kPoly>
__global__ void __maxnreg__(168)  // = 65536 / 384 threads, rounded down to the allocation unit
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                const AttnParams p) {
  constexpr int kMmaWarp = 8, kTmaWarp = 9;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_base = smem_base;                                   // 2 tiles
  const uint32_t kv_base = q_base + 2 * kTileBytes;                    // kStages tiles
  const uint32_t ostage_base = kv_base + kStages * kTileBytes;         // 2 x kOutStageBytes
  const uint32_t bar_base = ostage_base + 2 * kOutStageBytes;
  uint8_t* bar_gen = smem_gen + (bar_base - smem_base);
  auto qfull_bar = [&](int s) { return bar_base + 8u * s; };
  auto qempty_bar = [&](int s) { return bar_base + 8u * (2 + s); };
  auto sfull_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto pfull_bar = [&](int s) { return bar_base + 8u * (6 + s); };
  auto odone_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto ofree_bar = [&](int s) { return bar_base + 8u * (10 + s); };
  auto kvfull_bar = [&](int s) { return bar_base + 8u * (12 + s); };
  auto kvempty_bar = [&](int s) { return bar_base + 8u * (12 + kStages + s); };
  const uint32_t tmem_slot = bar_base + 8u * (14 + 2 * kStages);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(bar_gen + 8 * (14 + 2 * kStages));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
    for (int s = 0; s < 2; ++s) {
      mbar_init(qfull_bar(s), 1);
      mbar_init(qempty_bar(s), 1);
      mbar_init(sfull_bar(s), 1);
      mbar_init(pfull_bar(s), 4);  // one arrive per softmax warp
      mbar_init(odone_bar(s), 1);
      mbar_init(ofree_bar(s), 4);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(kvfull_bar(s), 1);
      mbar_init(kvempty_bar(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // programmatic dependent launch (see llb_common.cuh): the setup above may overlap the previous
  // kernel's tail; the kernel after this one may start its own setup as soon as our CTAs drain
  griddep_wait();
  griddep_launch_dependents();

  // work decomposition: identical in every role, evaluated inside each role branch so that only
  // the state a role needs stays live under its register budget
#define LLB_ATTN_INIT_WORK()                                            \
  KvTileIter kv_it;                                                     \
  kv_it.init(p.segs);                                                   \
  SegIter sg;                                                           \
  sg.init(kv_it.total_tiles(), p.n_heads * p.n_pairs, gridDim.x, blockIdx.x, p.min_split_tiles)

  // Register re-distribution: the kernel launches at 168 regs/thread (65536 / 384); the two softmax
  // warpgroups hold a full 128-column S row per thread and take 208, the MMA/TMA warpgroup keeps 88
  // ((168-88)*128 registers released >= (208-168)*256 requested, so the inc never blocks).
  if (warp >= kMmaWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == kTmaWarp) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        LLB_ATTN_INIT_WORK();
        int stage = 0;
        uint32_t phase = 0;
        uint32_t qph0 = 0, qph1 = 0;
        for (; sg.ok; sg.next()) {
          const int head = sg.item / p.n_pairs;
          const int q_row0 = (sg.item - head * p.n_pairs) * 256;
          const bool has1 = q_row0 + 128 < p.Lq;
          const int col = head * 128;
          // Q tiles: wait until the MMA warp has issued the previous segment's last QK on this slot
          mbar_wait(qempty_bar(0), qph0 ^ 1);
          qph0 ^= 1;
          mbar_arrive_expect_tx(qfull_bar(0), kTileBytes);
          tma_load_2d(q_base, &tmap_q, qfull_bar(0), col, q_row0);
          tma_load_2d(q_base + kBoxBytes, &tmap_q, qfull_bar(0), col + 64, q_row0);
          if (has1) {
            mbar_wait(qempty_bar(1), qph1 ^ 1);
            qph1 ^= 1;
            mbar_arrive_expect_tx(qfull_bar(1), kTileBytes);
            tma_load_2d(q_base + kTileBytes, &tmap_q, qfull_bar(1), col, q_row0 + 128);
            tma_load_2d(q_base + kTileBytes + kBoxBytes, &tmap_q, qfull_bar(1), col + 64, q_row0 + 128);
          }
          auto load_tile = [&](const CUtensorMap* tm, int row0) {
            mbar_wait(kvempty_bar(stage), phase ^ 1);
            const uint32_t dst = kv_base + stage * kTileBytes;
            mbar_arrive_expect_tx(kvfull_bar(stage), kTileBytes);
            tma_load_2d(dst, tm, kvfull_bar(stage), col, row0);
            tma_load_2d(dst + kBoxBytes, tm, kvfull_bar(stage), col + 64, row0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          };
          kv_it.seek(sg.t0);
          int row0, valid;
          // consumption order K_j, V_j, K_j+1, V_j+1, ...
          for (int j = sg.t0; j < sg.t1; ++j) {
            kv_it.get(row0, valid);
            load_tile(&tmap_k, row0);
            load_tile(&tmap_v, row0);
            kv_it.next();
          }
        }
      }
    } else if (warp == kMmaWarp) {
      // ------------------------------------------------------------------ MMA issuer
      // The whole warp runs this loop (waits included); one elected lane issues the tcgen05 ops.
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, 0, 1);
      auto issue_qk = [&](int t, uint32_t kst) {
        const uint32_t qa = q_base + t * kTileBytes;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t o = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
          umma_ss(tmem_base + t * 128, umma_desc_kmajor(qa + o), umma_desc_kmajor(kst + o),
                  idesc_qk, kk != 0);
        }
      };
      auto issue_pv = [&](int t, uint32_t vst, bool first) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          // V tile: rows = keys (K dim), two 64-wide d boxes 16 KB apart (MN dim); 16 keys per MMA;
          // A = P_t (bf16, 8 TMEM columns per 16 keys)
          const uint64_t bdesc = umma_desc_mnmajor(vst + kk * 2048, kBoxBytes);
          umma_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + kk * 8, bdesc, idesc_pv,
                  (first && kk == 0) ? 0u : 1u);
        }
      };
      LLB_ATTN_INIT_WORK();
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() { if (++stage == kStages) { stage = 0; phase ^= 1; } };
      uint32_t qph0 = 0, qph1 = 0;    // q_full consumer phases
      uint32_t pcnt0 = 0, pcnt1 = 0;  // kv tiles processed per Q tile (p_full phase)
      uint32_t oseg0 = 0, oseg1 = 0;  // segments started per Q tile (o_free phase)
      {
        [[maybe_unused]] const bool trace_on = true;
        LLB_TRACE(2, 0, 9);
      }
      for (; sg.ok; sg.next()) {
        const int head = sg.item / p.n_pairs;
        const int q_row0 = (sg.item - head * p.n_pairs) * 256;
        const bool has1 = q_row0 + 128 < p.Lq;
        const int nt = sg.t1 - sg.t0;
        [[maybe_unused]] const bool trace_on = sg.T <= 8 ? sg.k < 8 : sg.k == 0;
        [[maybe_unused]] const int trace_base = sg.T <= 8 ? sg.k * 8 : 0;
        LLB_TRACE(2, trace_base, 6);
        // prologue: S_t(first) = Q_t K^T
        mbar_wait(qfull_bar(0), qph0);
        qph0 ^= 1;
        if (has1) {
          mbar_wait(qfull_bar(1), qph1);
          qph1 ^= 1;
        }
        mbar_wait(kvfull_bar(stage), phase);
        tc_fence_after();
        uint32_t kst = kv_base + stage * kTileBytes;
        if (elect_one()) {
          issue_qk(0, kst);
          umma_commit(sfull_bar(0));
          if (nt == 1) umma_commit(qempty_bar(0));
          if (has1) {
            issue_qk(1, kst);
            umma_commit(sfull_bar(1));
            if (nt == 1) umma_commit(qempty_bar(1));
          }
          umma_commit(kvempty_bar(stage));
        }
        __syncwarp();
        advance();
        LLB_TRACE(2, trace_base, 7);
        // O_t of the previous segment must have been drained by its softmax warps before the
        // first PV of this segment overwrites it
        mbar_wait(ofree_bar(0), (oseg0 & 1) ^ 1);
        oseg0++;
        if (has1) {
          mbar_wait(ofree_bar(1), (oseg1 & 1) ^ 1);
          oseg1++;
        }
        LLB_TRACE(2, trace_base, 8);
        for (int j = 0; j < nt; ++j) {
          const bool more = j + 1 < nt;
          const bool last_qk = j + 2 == nt;  // the QK issued in this iteration is the segment's last
          // V_j
          const int vstage = stage;
          LLB_TRACE(2, trace_base + j, 0);
          mbar_wait(kvfull_bar(stage), phase);
          const uint32_t vst = kv_base + stage * kTileBytes;
          advance();
          // K_{j+1}
          int kstage = 0;
          if (more) {
            kstage = stage;
            mbar_wait(kvfull_bar(stage), phase);
            kst = kv_base + stage * kTileBytes;
            advance();
          }
          LLB_TRACE(2, trace_base + j, 1);
          // tile 0: O_0 += P_0(j) V_j ; S_0(j+1) = Q_0 K_{j+1}^T
          mbar_wait(pfull_bar(0), pcnt0 & 1);
          pcnt0++;
          tc_fence_after();
          LLB_TRACE(2, trace_base + j, 2);
          if (elect_one()) {
            issue_pv(0, vst, j == 0);
            umma_commit(odone_bar(0));
            if (more) {
              issue_qk(0, kst);
              umma_commit(sfull_bar(0));
              if (last_qk) umma_commit(qempty_bar(0));
            }
            if (!has1) {
              umma_commit(kvempty_bar(vstage));
              if (more) umma_commit(kvempty_bar(kstage));
            }
          }
          __syncwarp();
          LLB_TRACE(2, trace_base + j, 3);
          if (has1) {
            mbar_wait(pfull_bar(1), pcnt1 & 1);
            pcnt1++;
            tc_fence_after();
            LLB_TRACE(2, trace_base + j, 4);
            if (elect_one()) {
              issue_pv(1, vst, j == 0);
              umma_commit(odone_bar(1));
              umma_commit(kvempty_bar(vstage));
              if (more) {
                issue_qk(1, kst);
                umma_commit(sfull_bar(1));
                if (last_qk) umma_commit(qempty_bar(1));
                umma_commit(kvempty_bar(kstage));
              }
            }
            __syncwarp();
            LLB_TRACE(2, trace_base + j, 5);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // -------------------------------------------------------------------- softmax warps
    const int t = warp >> 2;  // Q tile handled by this warpgroup
    const int q = warp & 3;   // TMEM lane quadrant
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_s = tmem_base + lane_off + t * 128;
    const uint32_t t_o = tmem_base + lane_off + 256 + t * 128;
    const float c = p.scale_log2;
    uint32_t cnt = 0;  // kv tiles processed by this warpgroup (s_full / o_done phase)
    const int64_t ws_row = static_cast<int64_t>(t) * 128 + row_in_tile;
    LLB_ATTN_INIT_WORK();

    // P chunk cc (keys cc*32 .. +31) = exp2(S * c + neg), packed to bf16 into TMEM columns [cc*16, cc*16+16) of
    // the S tile; row sums of the fp32 probabilities accumulate in la / lb
    auto exp_chunk = [&](const uint32_t (&s)[32], uint32_t dst, float2 c2, float2 neg2, float2& la, float2& lb) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 tt = __ffma2_rn(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), c2, neg2);
        float2 pp;
        if (kPoly > 0 && (i % (kPoly > 0 ? kPoly : 1)) == (kPoly > 0 ? kPoly : 1) - 1) {
          pp = exp2_poly2(tt);
        } else {
          pp.x = ex2_approx(tt.x);
          pp.y = ex2_approx(tt.y);
        }
        if (i & 1) lb = __fadd2_rn(lb, pp);
        else la = __fadd2_rn(la, pp);
        pk[i] = pack_bf16x2(pp.x, pp.y);
      }
      tmem_st16(dst, pk);
    };
    // O_t *= f (TMEM read-modify-write by the row's owner; O must be stable, see the call sites)
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
      const int q_row0 = (sg.item - head * p.n_pairs) * 256;
      const bool has1 = q_row0 + 128 < p.Lq;
      if (t == 1 && !has1) continue;
      const int grow = q_row0 + t * 128 + row_in_tile;
      float m_used = -INFINITY;
      float l = 0.f;
      [[maybe_unused]] const bool trace_on = (sg.T <= 8 ? sg.k < 8 : sg.k == 0) && q == 0;
      [[maybe_unused]] const int trace_base = sg.T <= 8 ? sg.k * 8 : 0;
      kv_it.seek(sg.t0);
      for (int j = sg.t0; j < sg.t1; ++j, kv_it.next()) {
        int row0, valid;
        kv_it.get(row0, valid);
        LLB_TRACE(t, trace_base + j - sg.t0, 0);
        mbar_wait(sfull_bar(t), cnt & 1);
        cnt++;
        tc_fence_after();
        LLB_TRACE(t, trace_base + j - sg.t0, 1);
        // The tile body exists twice: full tiles (the common case) carry no masking code at all - as one body with
        // a run-time `valid < 128` test ptxas if-converted the masking into 128 ISETP + 128 SEL executed on EVERY
        // tile, 30 % of the softmax instructions (profiles/r02_attn_ncu_source.md)
        auto tile = [&](auto masked_tag) {
          constexpr bool kMasked = decltype(masked_tag)::value;
          uint32_t sv[4][32];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) tmem_ld32(t_s + cc * 32, sv[cc]);
          tmem_wait_ld();
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
          const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
          // lazy rescale: only when the max moved by more than 2^8 in the exp2 domain
          const bool need = (m_new - m_used) * c > 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            const float f = ex2_approx((m_used - m_new) * c);  // 0 on the first tile (m_used=-inf)
            // O must be stable: PV_t(j-1) complete, which S_t(j) being ready implies (issued before QK_t(j))
            if (j > sg.t0) rescale_o(f);
            l *= f;
            m_used = m_new;
          }
          const float neg = -m_used * c;
          const float2 c2 = make_float2(c, c), neg2 = make_float2(neg, neg);
          float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) exp_chunk(sv[cc], t_s + cc * 16, c2, neg2, la, lb);
          la = __fadd2_rn(la, lb);
          l += la.x + la.y;
        };
        if (valid < 128) tile(std::true_type{});
        else tile(std::false_type{});
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pfull_bar(t));
        LLB_TRACE(t, trace_base + j - sg.t0, 4);
      }
      // ---- segment epilogue
      mbar_wait(odone_bar(t), (cnt - 1) & 1);
      tc_fence_after();
      LLB_TRACE(t, trace_base + sg.t1 - sg.t0 - 1, 6);
      const bool tail_part = sg.in_remainder() && sg.t0 > 0;     // earlier kv tiles live in CTA blockIdx.x - 1
      const bool head_part = sg.in_remainder() && sg.t1 < sg.T;  // later kv tiles live in CTA blockIdx.x + 1
      const bool row_ok = grow < p.Lq;
      float a_own = 1.0f, a_oth = 0.0f;
      const float4* wo_in = nullptr;
      uint32_t* flag_in = nullptr;
      if (head_part && row_ok) {
        // wait for the partial (O, m, l) of the later kv tiles of this item
        uint8_t* wsb = p.workspace + static_cast<int64_t>(blockIdx.x + 1) * kWsPerCta;
        flag_in = reinterpret_cast<uint32_t*>(wsb + kWsOBytes + kWsMlBytes) + ws_row;
        uint32_t spins = 0;
        uint64_t t_start = 0;
        while (ld_acquire_u32(flag_in) == 0u) {
          if ((++spins & 0xfffu) == 0) {
            const uint64_t now = global_timer_ns();
            if (t_start == 0) t_start = now;
            else if (now - t_start > LLB_WAIT_TIMEOUT_NS) __trap();
          }
        }
        const volatile float* mlp = reinterpret_cast<const volatile float*>(wsb + kWsOBytes) + 2 * ws_row;
        const float m_oth = mlp[0], l_oth = mlp[1];
        const float m = fmaxf(m_used, m_oth);
        a_own = ex2_approx((m_used - m) * c);
        a_oth = ex2_approx((m_oth - m) * c);
        l = l * a_own + l_oth * a_oth;
        m_used = m;
        wo_in = reinterpret_cast<const float4*>(wsb) + static_cast<int64_t>(t) * (32 * 128) + row_in_tile;
      }
      uint8_t* wsb_out = p.workspace + static_cast<int64_t>(blockIdx.x) * kWsPerCta;
      float4* wo_out = reinterpret_cast<float4*>(wsb_out) + static_cast<int64_t>(t) * (32 * 128) + row_in_tile;
      if (!tail_part) {  // this CTA owns kv tile 0 of the item: normalise and write the output
        const float inv = 1.0f / l;
        a_own *= inv;
        a_oth *= inv;
      }
      __nv_bfloat16* orow = p.out + static_cast<int64_t>(grow) * p.ldo + head * 128;
      if (p.shard.n_ranks > 1 && row_ok) {
        // head-parallel mode: the return exchange is fused into the store - this token row belongs
        // to rank grow / rows_per_rank; write our heads' columns into its (peer-mapped) buffer
        const int r = grow / p.shard.rows_per_rank;
        orow = static_cast<__nv_bfloat16*>(p.shard.out_peers[r]) +
               static_cast<int64_t>(grow - r * p.shard.rows_per_rank) * p.shard.ld_out + p.shard.head_col0 +
               head * p.shard.head_col_stride;
      }
      // Final output of a single-GPU launch: through shared memory and a TMA store.  Written straight from registers a
      // thread owns a ROW, so a warp's 16-byte stores hit 32 different sectors: draining one O tile took 5600 - 7400 cycles
      // of LSU time (tools/attn_trace.py), a quarter of a cross-attention launch.  Staged as two [128 x 64] SWIZZLE_128B
      // boxes per Q tile, the store is two bulk tensor copies; rows beyond Lq are clipped by the tensor map.
      const bool staged = !tail_part && p.shard.n_ranks == 1;
      if (staged) {
        const uint32_t obuf = ostage_base + t * kOutStageBytes;
        const uint32_t orow_s = obuf + row_in_tile * 128;
        const bool leader = (warp & 3) == 0 && lane == 0;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t pk[32];
#pragma unroll
          for (int c2i = 0; c2i < 2; ++c2i) {
            const int cc = hh * 2 + c2i;
            uint32_t ov[32];
            tmem_ld32(t_o + cc * 32, ov);
            tmem_wait_ld();
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
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[c2i * 16 + i] = pack_bf16x2(o[2 * i], o[2 * i + 1]);
          }
          // the previous bulk store out of this buffer (first half, or the previous item) must have read it
          if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint32_t a = orow_s + static_cast<uint32_t>((ch ^ (row_in_tile & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * ch]), "r"(pk[4 * ch + 1]),
                         "r"(pk[4 * ch + 2]), "r"(pk[4 * ch + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
          if (leader) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmap_o)),
                         "r"(obuf), "r"(head * 128 + hh * 64), "r"(q_row0 + t * 128)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
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
      }
      if (flag_in != nullptr) st_release_u32(flag_in, 0u);  // consume: ready for the next launch
      if (tail_part && row_ok) {
        reinterpret_cast<float2*>(wsb_out + kWsOBytes)[ws_row] = make_float2(m_used, l);
        __threadfence();
        st_release_u32(reinterpret_cast<uint32_t*>(wsb_out + kWsOBytes + kWsMlBytes) + ws_row, 1u);
      }
      // O_t drained: the MMA warp may start the next segment's accumulation
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ofree_bar(t));
      LLB_TRACE(t, trace_base + sg.t1 - sg.t0 - 1, 7);
    }
    // the bulk stores of this thread are complete (and have released the staging buffer) before the CTA exits
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int kPoly>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                       const AttnParams& p, int grid, cudaStream_t stream) {
  LLB_SET_MAX_SMEM((attn_fwd_kernel<kPoly>), kAttnSmemBytes);
  // cooperative launch: the runtime guarantees (or refuses) co-residency of all CTAs, which the
  // partial-merge flag wait relies on
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kAttnThreads);
  cfg.dynamicSmemBytes = kAttnSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n_attr = 0;
  static const bool coop = getenv("LLB_ATTN_COOP") == nullptr || atoi(getenv("LLB_ATTN_COOP")) != 0;
  if (coop) {
    attr[n_attr].id = cudaLaunchAttributeCooperative;
    attr[n_attr].val.cooperative = 1;
    ++n_attr;
  }
  // LLB_ATTN_PDL=1 also launches the attention kernel itself programmatically (off by default: it is
  // the dependents of this long kernel, not the kernel, that gain from the overlap)
  static const bool pdl = getenv("LLB_ATTN_PDL") != nullptr && atoi(getenv("LLB_ATTN_PDL")) != 0;
  if (pdl && pdl_enabled()) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  LLB_CUDA(cudaLaunchKernelEx(&cfg, attn_fwd_kernel<kPoly>, tq, tk, tv, to, p));
  LLB_LAUNCH_CHECK("attn_fwd_kernel");
  return LLB_OK;
}

}  // namespace llb

#include "attn_pair.cuh"

#ifdef LLB_ATTN_TRACE
extern "C" int llb_attn_trace_read(unsigned long long* host, int n) {
  cudaDeviceSynchronize();
  return static_cast<int>(cudaMemcpyFromSymbol(host, llb::g_attn_trace, sizeof(unsigned long long) * n));
}
#endif

extern "C" int64_t llb_attn_workspace_bytes(void) {
  const int sms = llb::device_sm_count();
  return static_cast<int64_t>(sms > 0 ? sms : 148) * llb::kWsPerCta;
}

extern "C" int llb_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                            int64_t ldv, void* out, int64_t ldo, int Lq, int n_heads, int kv_rows,
                            const llb_step_params* seg_dev, float scale, int variant, void* workspace,
                            int64_t workspace_bytes, const llb_out_shard* shard, void* stream) {
  using namespace llb;
  LLB_CHECK_ARG(q && k && v && out && seg_dev, "attn: null tensor");
  LLB_CHECK_ARG(Lq > 0 && n_heads > 0 && kv_rows > 0, "attn: bad shape");
  LLB_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
                "attn: leading dims must be multiples of 8");
  LLB_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "attn: out must be 16-byte aligned");
  const int sms = device_sm_count();
  LLB_CHECK_ARG(sms > 0, "attn: no CUDA device");
  const int n_pairs = (Lq + 255) / 256;
  const int n_items = n_pairs * n_heads;
  // one CTA per SM; when there are fewer items than SMs the scheduler splits every item's kv range
  // (rounds = 0, all items are "remainder"), CTAs without work exit immediately
  int grid = sms;
  // debug knob (needs LLB_ATTN_COOP=0): one CTA per item instead of a persistent grid
  if (getenv("LLB_ATTN_GRID_ITEMS") != nullptr && atoi(getenv("LLB_ATTN_GRID_ITEMS")) != 0) grid = n_items;
  LLB_CHECK_ARG(workspace != nullptr && workspace_bytes >= static_cast<int64_t>(sms) * kWsPerCta &&
                    (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                "attn: needs a 16-byte aligned workspace of llb_attn_workspace_bytes() bytes, zeroed once");
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_2d_bf16(&tq, q, Lq, static_cast<uint64_t>(n_heads) * 128, ldq, 128, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tk, k, kv_rows, static_cast<uint64_t>(n_heads) * 128, ldk, 128, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tv, v, kv_rows, static_cast<uint64_t>(n_heads) * 128, ldv, 128, 64);
  if (rc) return rc;
  AttnParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.Lq = Lq;
  p.n_heads = n_heads;
  p.n_pairs = n_pairs;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.segs = seg_dev;
  p.workspace = static_cast<uint8_t*>(workspace);
  // measured (profiles/r02_attn_pair.md): splitting the 4-tile items of cross-attention costs more in merges than the
  // idle half round it removes (371 -> 308 TFLOP/s at 4 / 296 at 2)
  p.min_split_tiles = 16;
  memset(&p.shard, 0, sizeof(p.shard));
  p.shard.n_ranks = 1;
  if (shard != nullptr && shard->n_ranks > 1) {
    LLB_CHECK_ARG(shard->n_ranks <= LLB_MAX_RANKS && shard->rows_per_rank > 0 && shard->ld_out % 8 == 0 &&
                      shard->head_col0 % 8 == 0 && shard->head_col_stride % 8 == 0 && shard->head_col_stride >= 0,
                  "attn: bad output shard description");
    for (int r = 0; r < shard->n_ranks; ++r)
      LLB_CHECK_ARG(shard->out_peers[r] != nullptr, "attn: null peer pointer");
    p.shard = *shard;
  }
  if (p.shard.head_col_stride == 0) p.shard.head_col_stride = 128;  // contiguous local heads
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // variant bit 6 (64): the CTA-pair kernel (attn_pair.cuh)
  if (variant & 64) {
    LLB_CHECK_ARG(sms >= 2, "attn: the CTA-pair kernel needs at least two SMs");
    CUtensorMap tk64;
    rc = make_tmap_2d_bf16(&tk64, k, kv_rows, static_cast<uint64_t>(n_heads) * 128, ldk, 64, 64);
    if (rc) return rc;
    return launch_attn_pair(tq, tk64, tv, p, sms, s);
  }
  // output map for the staged TMA store (single-GPU launches; the head-parallel path stores to peer memory directly)
  CUtensorMap to = tq;
  if (p.shard.n_ranks == 1) {
    rc = make_tmap_2d_bf16(&to, out, Lq, static_cast<uint64_t>(n_heads) * 128, ldo, 128, 64);
    if (rc) return rc;
  }
  // kPoly = 4: A/B in the full pipeline under the power cap (profiles/r02_poly_ab.txt): MUFU only runs at a higher clock
  // (1650 vs 1620 MHz) but slower per clock, 95.2 vs 95.9 FPS; every 8th pair 95.2, every 2nd 93.2
  return launch_attn<4>(tq, tk, tv, to, p, grid, s);
}

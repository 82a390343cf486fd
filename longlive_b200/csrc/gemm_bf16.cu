// llb_gemm_bf16 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] = epilogue(A[M,K] @ W[N,K]^T + bias)       A, W, out bf16; fp32 accumulation in TMEM
//
// Replaces the nn.Linear calls of the LongLive block (wan/modules/causal_model.py:122-126, 364,
// 406-408, 492; wan/modules/model.py:172-178, 193) and folds the elementwise tail that follows
// each of them in the reference into the epilogue:
//   GELU-tanh (causal_model.py:407), gated residual x + y*e[k] (:456, :467-468), plain residual
//   (:460), SiLU (:606).  Rounding points mirror the reference (bf16 after the Linear, after the
//   gate multiply, after the residual add) so the result is bit-comparable to the PyTorch path.
//
// Structure (one CTA per SM, 320 threads):
//   warp 0      TMA producer: A tile [128 x 64] and W tile per k-block into a kStages-deep
//               shared-memory ring (SWIZZLE_128B), full/empty mbarriers
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma x4 per k-block into one of two
//               TMEM accumulator stages; tcgen05.commit frees smem slots / signals the epilogue
//   warps 2..9  epilogue: two warps per TMEM lane quadrant, each taking every other 32-column
//               slab: tcgen05.ld (thread == row), bias, activation, round to bf16, transpose
//               through a swizzled shared-memory staging buffer so that global stores (and
//               residual / gate loads) are coalesced 64-byte row segments
// The two TMEM stages let tile i's epilogue overlap tile i+1's main loop.
//
// Two tile modes.  The main loop is bound by L2 -> SM bytes, not by the tensor pipe (ncu: ~14.5 TB/s
// of xbar reads at 60 % tensor-pipe activity with 128 x 192 tiles), so the mode that moves fewer bytes
// per flop wins whenever the tile count still fills the machine:
//   single  one CTA computes a 128 x BN tile            (16 KB A + BN*128 B W per k-block)
//   pair    a cluster of two CTAs on one TPC computes a 256 x BN tile with cta_group::2 MMAs: each
//           CTA loads its own 128 rows of A and HALF of the W tile (16 KB + BN*64 B per k-block),
//           the leader CTA issues the MMAs for both, each CTA's TMEM receives its 128 rows and
//           its own epilogue warps drain them.
#include <stdlib.h>

#include "llb_common.cuh"
#include "llb_host.h"

namespace llb {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr int kEpiStageBytesPerWarp = 32 * 64;  // 32 rows x 32 bf16 columns, 16-byte chunks xor-swizzled

template <int BN, bool kPair>
struct GemmCfg {
  static constexpr int kStageA = kBM * kBK * 2;
  static constexpr int kRowsB = kPair ? BN / 2 : BN;  // W rows this CTA loads per k-block
  static constexpr int kStageB = kRowsB * kBK * 2;
  static constexpr int kStageBytes = kStageA + kStageB;
  // TMEM allocations are powers of two >= 32 columns; two accumulator stages of BN columns each
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  // per epilogue warp: fp32 bias and w_scale for the BN/2 columns it owns
  static constexpr int kEpiVecBytesPerWarp = 2 * (BN / 2) * 4;
  static constexpr int kEpiBytes = kEpiWarps * (kEpiStageBytesPerWarp + kEpiVecBytesPerWarp);
  static constexpr int kFixedBytes = 1024 /*align slack*/ + kEpiBytes + 256 /*barriers*/;
  static constexpr int kStagesFit = (232448 - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kSmemBytes = kFixedBytes + kStages * kStageBytes;
  static_assert(kStages >= 3 && kSmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static_assert(2 * kStages + 5 <= 32, "barrier block is 256 bytes");
};

struct GemmParams {
  int M, N, K;
  int epilogue;
  __nv_bfloat16* out;
  int64_t ldo;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* gate;
  int64_t ld_gate;
  int rows_per_gate;
  int gate_row0;
  const __nv_bfloat16* res;
  int64_t ld_res;
  int num_m_tiles, num_n_tiles;  // pair mode: num_m_tiles counts 256-row tiles
  int k_splits, kb_per_split;    // split-K (llb_gemm_bf16_splitk): tile t also selects k-blocks [ks * kb_per_split, ...)
                                 // and its fp32 partial goes to slice ks of the workspace `out` points at
  const float* a_scale;  // fp8 path: per-row activation scale [M]
  const float* w_scale;  // fp8 path: per-output-channel weight scale [N]
};

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // torch.nn.GELU(approximate='tanh'): 0.5*x*(1+tanh(u)), u = sqrt(2/pi)*(x+0.044715*x^3).
  // 0.5*(1+tanh(u)) == sigmoid(2u): same function, no cancellation in the negative tail, and it
  // needs one ex2 + one rcp instead of a full-precision tanhf.
  const float kBeta = 0.7978845608028654f, kKappa = 0.044715f;
  const float u = kBeta * (x + kKappa * x * x * x);
  return __fdividef(x, 1.0f + __expf(-2.0f * u));
}
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// The umT5 encoder's GELU module (wan/modules/t5.py:46-50) spells the tanh approximation out as seven bf16
// tensor ops, each rounding its result; in particular 1 + tanh(u) cancels in bf16 for negative inputs.  Parity
// with the reference means reproducing that chain, not the exact function.  x is already bf16-valued.
__device__ __forceinline__ float gelu_bf16_chain_f(float x) {
  const float p3 = bf16_round(x * x * x);                         // torch.pow(x, 3.0)
  const float a = bf16_round(0.044715f * p3);
  const float b = bf16_round(x + a);
  const float c = bf16_round(0.7978845608028654f * b);            // math.sqrt(2 / pi) * (...)
  const float t = bf16_round(1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * c)));  // tanh(c), error << bf16 ulp
  const float u = bf16_round(1.0f + t);
  return bf16_round(0.5f * x) * u;
}

// kFp8: A and W are e4m3 bytes (K-block = 128 elements = the same 128-byte swizzle row), the MMA is
// kind::f8f6f4 (K = 32 per instruction, twice the bf16 rate) and the epilogue applies
// a_scale[row] * w_scale[col] before the bias.
template <int BN, bool kFp8, bool kPair>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using Cfg = GemmCfg<BN, kPair>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment (the dynamic smem base is the same in both CTAs of a
  // pair, so every offset below is too - cta_group::2 MMAs and multicast commits rely on that)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stage_base = smem_base;
  const uint32_t epi_base = smem_base + kStages * Cfg::kStageBytes;
  uint8_t* epi_gen = smem_gen + kStages * Cfg::kStageBytes;
  const uint32_t bar_base = epi_base + Cfg::kEpiBytes;
  // barrier layout: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(epi_gen + Cfg::kEpiBytes + 8 * (2 * kStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;  // 0 = leader of the pair
  const int worker = kPair ? (blockIdx.x >> 1) : blockIdx.x;
  const int num_workers = kPair ? (gridDim.x >> 1) : gridDim.x;
  const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
  const int num_tiles = tiles_mn * p.k_splits;
  constexpr int kKElems = kFp8 ? 128 : kBK;  // elements per k-block (always 128 bytes)
  const int num_kb_total = (p.K + kKElems - 1) / kKElems;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      // one arrive per epilogue warp; in pair mode the leader's barrier also collects the peer's
      mbar_init(tempty_bar(s), kPair ? 2 * kEpiWarps : kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (kPair) {
      tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();  // the peer's barriers must exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous
  // kernel's tail when launched programmatically; from here on its output is complete and visible.
  griddep_wait();
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        const int ks = tile / tiles_mn, tmn = tile - ks * tiles_mn;
        const int m_idx = tmn % p.num_m_tiles;
        const int n_idx = tmn / p.num_m_tiles;
        const int a_row = kPair ? (m_idx * 2 + static_cast<int>(cta_rank)) * kBM : m_idx * kBM;
        const int b_row = n_idx * BN + (kPair ? static_cast<int>(cta_rank) * (BN / 2) : 0);
        const int kb0 = ks * p.kb_per_split;
        const int num_kb = min(p.kb_per_split, num_kb_total - kb0);
        for (int kb = kb0; kb < kb0 + num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = stage_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kStageA;
          if constexpr (kPair) {
            // both CTAs' bytes are counted on the leader's barrier, which the leader arms once
            const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
            tma_load_2d_pair(sa, &tmap_a, lead_full, kb * kKElems, a_row);
            tma_load_2d_pair(sb, &tmap_b, lead_full, kb * kKElems, b_row);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kKElems, a_row);
            tma_load_2d(sb, &tmap_b, full_bar(stage), kb * kKElems, b_row);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Whole warp in the loop (so descriptors stay in uniform registers); one elected lane issues.
    // Pair mode: only the leader CTA issues; its commits are multicast to both CTAs' barriers.
    if (cta_rank == 0) {
      constexpr int kMmaM = kPair ? 2 * kBM : kBM;
      constexpr uint32_t idesc = kFp8 ? umma_idesc_e4m3(kMmaM, BN) : umma_idesc_bf16(kMmaM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int num_kb = min(p.kb_per_split, num_kb_total - (tile / tiles_mn) * p.kb_per_split);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = stage_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kStageA;
          const uint64_t da = umma_desc_kmajor(sa);
          const uint64_t db = umma_desc_kmajor(sb);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the
              // (addr >> 4) field of the descriptor
              const uint32_t accum = (kb | k) != 0;
              if constexpr (kPair) {
                if constexpr (kFp8) umma_ss_f8_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                else umma_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
              } else {
                if constexpr (kFp8) umma_ss_f8(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                else umma_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
              }
            }
            if constexpr (kPair) {
              umma_commit_pair(empty_bar(stage), 3);
              if (kb == num_kb - 1) umma_commit_pair(tfull_bar(acc), 3);
            } else {
              umma_commit(empty_bar(stage));
              if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    const int h = (warp - 2) >> 2;   // which 32-column slab of every 64 columns it owns
    uint8_t* my_stage = epi_gen + (warp - 2) * kEpiStageBytesPerWarp;
    float* my_bias = reinterpret_cast<float*>(epi_gen + kEpiWarps * kEpiStageBytesPerWarp +
                                              (warp - 2) * Cfg::kEpiVecBytesPerWarp);
    float* my_ws = my_bias + BN / 2;
    const uint32_t tempty_lead0 = kPair ? mapa_shared(tempty_bar(0), 0) : tempty_bar(0);
    const int epi = p.epilogue;
    const bool has_mul = epi == LLB_EPI_BIAS_MUL;
    const bool has_res = (epi == LLB_EPI_BIAS_GATE_RES || epi == LLB_EPI_BIAS_RES || has_mul);
    const bool has_gate = epi == LLB_EPI_BIAS_GATE_RES;
    int it = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
      const int ks = tile / tiles_mn, tmn = tile - ks * tiles_mn;
      const int m_idx = tmn % p.num_m_tiles;
      const int n_idx = tmn / p.num_m_tiles;
      const int row_base = (kPair ? (m_idx * 2 + static_cast<int>(cta_rank)) * kBM : m_idx * kBM) + q * 32;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // Stage this warp's bias (and fp8 weight scales) as fp32 in shared memory while the main loop
      // is still running: the per-column values are then broadcast LDS reads instead of global loads
      // whose latency the epilogue warps cannot hide.
#pragma unroll
      for (int i = 0; i < BN / 64; ++i) {
        const int cg = n_idx * BN + (2 * i + h) * 32 + lane;
        my_bias[i * 32 + lane] = (p.bias != nullptr && cg < p.N) ? __bfloat162float(p.bias[cg]) : 0.f;
        if constexpr (kFp8) my_ws[i * 32 + lane] = cg < p.N ? __ldg(p.w_scale + cg) : 0.f;
      }
      float row_scale = 1.0f;
      if constexpr (kFp8) row_scale = row_base + lane < p.M ? __ldg(p.a_scale + row_base + lane) : 0.f;
      __syncwarp();
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + h * 32;
      if constexpr (BN == 256 && !kFp8) {
        if (epi == LLB_EPI_GEGLU_BF16) {
          // Gated FFN in one launch (umT5, wan/modules/t5.py:133): the weight rows of a 256-wide tile are 128 gate rows
          // followed by the 128 fc1 rows of the same output columns, so this warp finds the gate slab (2c + h) and the
          // fc1 slab 128 columns further in the same accumulator: out = bf16(fc1) * gelu_bf16_chain(bf16(gate)).
          // thread == row writes 32 consecutive bf16 (64 bytes) per slab; the output has N / 2 columns.
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t vg[32], vf[32];
            tmem_ld32(t_row + c * 64, vg);
            tmem_ld32(t_row + (c + 2) * 64, vf);
            tmem_wait_ld();
            if (c == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (kPair) mbar_arrive_cluster(tempty_lead0 + 8u * acc);
                else mbar_arrive(tempty_bar(acc));
              }
            }
            const int grow = row_base + lane;
            const int ocol = n_idx * 128 + (2 * c + h) * 32;
            if (grow < p.M && ocol < p.N / 2) {
              uint32_t o[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float g0 = gelu_bf16_chain_f(bf16_round(__uint_as_float(vg[2 * i])));
                const float g1 = gelu_bf16_chain_f(bf16_round(__uint_as_float(vg[2 * i + 1])));
                o[i] = pack_bf16x2(bf16_round(__uint_as_float(vf[2 * i])) * bf16_round(g0),
                                   bf16_round(__uint_as_float(vf[2 * i + 1])) * bf16_round(g1));
              }
              uint4* orow = reinterpret_cast<uint4*>(p.out + static_cast<int64_t>(grow) * p.ldo + ocol);
#pragma unroll
              for (int i = 0; i < 4; ++i) orow[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            }
          }
          continue;
        }
      }
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int col0 = n_idx * BN + (2 * c + h) * 32;
        uint32_t v[32];
        tmem_ld32(t_row + c * 64, v);
        tmem_wait_ld();
        if (c == BN / 64 - 1) {
          // accumulator slabs fully drained into registers: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kPair) mbar_arrive_cluster(tempty_lead0 + 8u * acc);
            else mbar_arrive(tempty_bar(acc));
          }
        }
        if (epi == LLB_EPI_BIAS_F32) {
          // fp32 output: thread == row writes its 32 consecutive columns (one full 128-byte line)
          const int grow = row_base + lane;
          if (grow < p.M) {
            float* orow = reinterpret_cast<float*>(p.out) + (static_cast<int64_t>(ks) * p.M + grow) * p.ldo + col0;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (col0 + g * 4 < p.N) {
                const float4 b4 = *reinterpret_cast<const float4*>(my_bias + c * 32 + g * 4);
                float4 o4 = make_float4(__uint_as_float(v[g * 4]) + b4.x, __uint_as_float(v[g * 4 + 1]) + b4.y,
                                        __uint_as_float(v[g * 4 + 2]) + b4.z, __uint_as_float(v[g * 4 + 3]) + b4.w);
                if constexpr (kFp8) {
                  const float4 w4 = *reinterpret_cast<const float4*>(my_ws + c * 32 + g * 4);
                  o4 = make_float4(__uint_as_float(v[g * 4]) * w4.x * row_scale + b4.x,
                                   __uint_as_float(v[g * 4 + 1]) * w4.y * row_scale + b4.y,
                                   __uint_as_float(v[g * 4 + 2]) * w4.z * row_scale + b4.z,
                                   __uint_as_float(v[g * 4 + 3]) * w4.w * row_scale + b4.w);
                }
                *reinterpret_cast<float4*>(orow + g * 4) = o4;
              }
            }
          }
          continue;
        }
        // phase 1: thread == row.  bias (+activation) -> bf16 -> staging row (64 bytes, 4 chunks)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t packed[4];
          // 8 bias values (and fp8 scales) for columns col0 + 8g .. +7: broadcast LDS
          const float4 b0 = *reinterpret_cast<const float4*>(my_bias + c * 32 + g * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(my_bias + c * 32 + g * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float ws[8];
          if constexpr (kFp8) {
            const float4 w0 = *reinterpret_cast<const float4*>(my_ws + c * 32 + g * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(my_ws + c * 32 + g * 8 + 4);
            ws[0] = w0.x * row_scale; ws[1] = w0.y * row_scale; ws[2] = w0.z * row_scale; ws[3] = w0.w * row_scale;
            ws[4] = w1.x * row_scale; ws[5] = w1.y * row_scale; ws[6] = w1.z * row_scale; ws[7] = w1.w * row_scale;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = g * 8 + e * 2;  // column inside the 32-wide slab
            float a0 = __uint_as_float(v[j]);
            float a1 = __uint_as_float(v[j + 1]);
            if constexpr (kFp8) {
              a0 *= ws[2 * e];
              a1 *= ws[2 * e + 1];
            }
            float y0 = a0 + bb[2 * e], y1 = a1 + bb[2 * e + 1];
            if (epi == LLB_EPI_BIAS_GELU) {
              y0 = gelu_tanh_f(bf16_round(y0));
              y1 = gelu_tanh_f(bf16_round(y1));
            } else if (epi == LLB_EPI_BIAS_SILU) {
              y0 = silu_f(bf16_round(y0));
              y1 = silu_f(bf16_round(y1));
            } else if (epi == LLB_EPI_BIAS_GELU_BF16) {
              y0 = gelu_bf16_chain_f(bf16_round(y0));
              y1 = gelu_bf16_chain_f(bf16_round(y1));
            }
            packed[e] = pack_bf16x2(y0, y1);
          }
          *reinterpret_cast<uint4*>(my_stage + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
              make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        __syncwarp();
        // phase 2: 4 lanes cover one 64-byte row segment, 8 rows per pass -> coalesced global traffic.
        // The residual / gate loads of all four passes are issued together before any arithmetic so
        // their latencies overlap.
        const int seg = lane & 3;
        const int gcol = col0 + seg * 8;
        const bool col_ok = gcol < p.N;
        uint4 yv[4], xv[4], gv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2);
          const int grow = row_base + r;
          yv[i] = *reinterpret_cast<const uint4*>(my_stage + r * 64 + ((seg ^ ((r >> 1) & 3)) << 4));
          xv[i] = make_uint4(0, 0, 0, 0);
          gv[i] = make_uint4(0, 0, 0, 0);
          if (grow < p.M && col_ok) {
            if (has_res)
              xv[i] = *reinterpret_cast<const uint4*>(p.res + static_cast<int64_t>(grow) * p.ld_res + gcol);
            if (has_gate)
              gv[i] = __ldg(reinterpret_cast<const uint4*>(
                  p.gate + static_cast<int64_t>((grow + p.gate_row0) / p.rows_per_gate) * p.ld_gate + gcol));
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2);
          const int grow = row_base + r;
          if (grow < p.M && col_ok) {
            uint4 y = yv[i];
            if (has_res) {
              const uint32_t* yy = reinterpret_cast<const uint32_t*>(&yv[i]);
              const uint32_t* xx = reinterpret_cast<const uint32_t*>(&xv[i]);
              const uint32_t* gg = reinterpret_cast<const uint32_t*>(&gv[i]);
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float y0 = bf16_lo(yy[e]), y1 = bf16_hi(yy[e]);
                if (has_gate) {
                  y0 = bf16_round(y0 * bf16_lo(gg[e]));
                  y1 = bf16_round(y1 * bf16_hi(gg[e]));
                }
                o[e] = has_mul ? pack_bf16x2(bf16_lo(xx[e]) * y0, bf16_hi(xx[e]) * y1)
                               : pack_bf16x2(bf16_lo(xx[e]) + y0, bf16_hi(xx[e]) + y1);
              }
              y = make_uint4(o[0], o[1], o[2], o[3]);
            }
            *reinterpret_cast<uint4*>(p.out + static_cast<int64_t>(grow) * p.ldo + gcol) = y;
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  // pair mode: neither CTA may exit (or free its TMEM) while the other can still signal its barriers
  if constexpr (kPair) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, bool kFp8, bool kPair>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, kPair>;
  LLB_SET_MAX_SMEM((gemm_bf16_kernel<BN, kFp8, kPair>), Cfg::kSmemBytes);
  const int sms = device_sm_count();
  LLB_CHECK_ARG(sms > 0, "no CUDA device");
  const int tiles = p.num_m_tiles * p.num_n_tiles * p.k_splits;
  const int workers = kPair ? sms / 2 : sms;
  const int grid = (tiles < workers ? tiles : workers) * (kPair ? 2 : 1);
  LLB_CUDA(launch_ex(gemm_bf16_kernel<BN, kFp8, kPair>, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream,
                     kPair ? 2 : 1, true, ta, tb, p));
  LLB_LAUNCH_CHECK("gemm_bf16_kernel");
  return LLB_OK;
}

template <bool kFp8, bool kPair>
static int dispatch_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       cudaStream_t s) {
  switch (bn) {
    case 64:
      if constexpr (kPair) break;  // a pair tile narrower than 128 columns is never chosen
      else return launch_gemm<64, kFp8, false>(ta, tb, p, s);
    case 128: return launch_gemm<128, kFp8, kPair>(ta, tb, p, s);
    case 192: return launch_gemm<192, kFp8, kPair>(ta, tb, p, s);
    case 256: return launch_gemm<256, kFp8, kPair>(ta, tb, p, s);
    default: break;
  }
  set_error("gemm: unsupported tile width %d (pair=%d)", bn, static_cast<int>(kPair));
  return LLB_E_INVALID;
}

}  // namespace llb

static int gemm_impl(bool fp8, const void* A, int64_t lda, const void* W, int64_t ldw, void* out,
                     int64_t ldo, int M, int N, int K, int epilogue, const void* bias, const void* gate,
                     int64_t ld_gate, int rows_per_gate, int gate_row0, const void* res, int64_t ld_res,
                     const float* a_scale, const float* w_scale, void* stream, int k_splits = 1) {
  using namespace llb;
  LLB_CHECK_ARG(A && W && out, "gemm: null tensor");
  LLB_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  LLB_CHECK_ARG(K % (fp8 ? 16 : 8) == 0 && N % 8 == 0, "gemm: K / N alignment (K=%d N=%d)", K, N);
  LLB_CHECK_ARG(lda % (fp8 ? 16 : 8) == 0 && ldw % (fp8 ? 16 : 8) == 0 && (ldo % 8 == 0 || epilogue == LLB_EPI_BIAS_F32),
                "gemm: leading dims must be 16-byte multiples");
  LLB_CHECK_ARG(!fp8 || (a_scale && w_scale && (reinterpret_cast<uintptr_t>(w_scale) & 15) == 0),
                "gemm_fp8: needs a_scale[M] and a 16-byte aligned w_scale[N]");
  LLB_CHECK_ARG(epilogue >= 0 && epilogue <= LLB_EPI_GEGLU_BF16, "gemm: unknown epilogue %d", epilogue);
  LLB_CHECK_ARG(epilogue != LLB_EPI_BIAS_F32 || ldo % 4 == 0, "gemm: fp32 output needs ldo %% 4 == 0");
  {
    const int nkb = (K + (fp8 ? 128 : kBK) - 1) / (fp8 ? 128 : kBK);
    LLB_CHECK_ARG(k_splits >= 1 && k_splits <= nkb && (k_splits - 1) * ((nkb + k_splits - 1) / k_splits) < nkb,
                  "gemm: k_splits=%d leaves an empty split for K=%d", k_splits, K);
    LLB_CHECK_ARG(k_splits == 1 || (epilogue == LLB_EPI_BIAS_F32 && bias == nullptr),
                  "gemm: split-K writes raw fp32 partials");
  }
  if (epilogue == LLB_EPI_BIAS_GATE_RES) {
    LLB_CHECK_ARG(gate && res && rows_per_gate > 0 && ld_gate % 8 == 0 && ld_res % 8 == 0,
                  "gemm: gate/residual epilogue needs gate, res, rows_per_gate");
  }
  if (epilogue == LLB_EPI_BIAS_RES || epilogue == LLB_EPI_BIAS_MUL)
    LLB_CHECK_ARG(res && ld_res % 8 == 0, "gemm: residual / multiplier epilogue needs res");
  LLB_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "gemm: out must be 16-byte aligned");
  LLB_CHECK_ARG((reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(gate) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(res) & 15) == 0, "gemm: bias/gate/res must be 16-byte aligned");

  // Tile shape: minimise waves x per-tile time.  The per-tile costs are the measured steady-state time
  // per k-block (ns, B200, every SM streaming; tools/kernel_bench.py --what tilesweep, see DESIGN.md
  // section 4.2): the main loop runs at the L2 -> SM transfer rate, so wider tiles and CTA pairs (which
  // move 1.5-1.7x fewer bytes per flop) are cheaper per flop, but they quantise the tile count coarser.
  // With e4m3 operands a pair tile is never faster than a single one.
  const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
  int bn = 64;
  bool pair = false;
  if (N > 64) {
    static const double kCostBf16[2][3] = {{330, 388, 446}, {275, 325, 420}};  // [pair][bn 128/192/256]
    static const double kCostFp8[2][3] = {{383, 475, 567}, {375, 475, 600}};
    double best = 1e30;
    for (int pr = 0; pr < 2; ++pr) {
      const int workers = pr ? sms / 2 : sms;
      const int m_tiles = pr ? (M + 2 * kBM - 1) / (2 * kBM) : (M + kBM - 1) / kBM;
      for (int i = 0; i < 3; ++i) {
        const int cand = 128 + 64 * i;
        const int tiles = m_tiles * ((N + cand - 1) / cand) * k_splits;
        const int waves = (tiles + workers - 1) / workers;
        const double cost = waves * (fp8 ? kCostFp8[pr][i] : kCostBf16[pr][i]);
        if (cost < best) { best = cost; bn = cand; pair = pr != 0; }
      }
    }
  }
  if (epilogue == LLB_EPI_GEGLU_BF16) {
    // gate | fc1 interleaved in 256-row weight tiles: only 256-wide tiles see both halves of an output column
    LLB_CHECK_ARG(!fp8 && N % 256 == 0 && bias == nullptr && k_splits == 1, "gemm: GEGLU needs N %% 256 == 0, no bias");
    const int pair_waves = (((M + 2 * kBM - 1) / (2 * kBM)) * (N / 256) + sms / 2 - 1) / (sms / 2);
    const int single_waves = (((M + kBM - 1) / kBM) * (N / 256) + sms - 1) / sms;
    bn = 256;
    pair = pair_waves * 420.0 <= single_waves * 446.0;
  }
  // debugging / benchmarking override: LLB_GEMM_TILE="<pair 0|1>,<bn>"
  const char* force = getenv("LLB_GEMM_TILE");
  if (force != nullptr && N > 64 && epilogue != LLB_EPI_GEGLU_BF16) {
    int fp = 0, fb = 0;
    if (sscanf(force, "%d,%d", &fp, &fb) == 2 && (fb == 128 || fb == 192 || fb == 256)) {
      pair = fp != 0;
      bn = fb;
    }
  }

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.epilogue = epilogue;
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.gate = static_cast<const __nv_bfloat16*>(gate); p.ld_gate = ld_gate;
  p.rows_per_gate = rows_per_gate > 0 ? rows_per_gate : 1;
  p.gate_row0 = gate_row0;
  p.res = static_cast<const __nv_bfloat16*>(res); p.ld_res = ld_res;
  p.num_m_tiles = pair ? (M + 2 * kBM - 1) / (2 * kBM) : (M + kBM - 1) / kBM;
  p.num_n_tiles = (N + bn - 1) / bn;
  p.a_scale = a_scale;
  p.w_scale = w_scale;
  p.k_splits = k_splits;
  p.kb_per_split = ((K + (fp8 ? 128 : kBK) - 1) / (fp8 ? 128 : kBK) + k_splits - 1) / k_splits;

  CUtensorMap ta, tb;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int b_box_rows = pair ? bn / 2 : bn;  // each CTA of a pair loads half of the W tile
  int rc;
  if (fp8) {
    rc = make_tmap_2d_u8(&ta, A, M, K, lda, kBM, 128);
    if (rc) return rc;
    rc = make_tmap_2d_u8(&tb, W, N, K, ldw, b_box_rows, 128);
    if (rc) return rc;
    return pair ? dispatch_bn<true, true>(bn, ta, tb, p, s) : dispatch_bn<true, false>(bn, ta, tb, p, s);
  }
  rc = make_tmap_2d_bf16(&ta, A, M, K, lda, kBM, kBK);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tb, W, N, K, ldw, b_box_rows, kBK);
  if (rc) return rc;
  return pair ? dispatch_bn<false, true>(bn, ta, tb, p, s) : dispatch_bn<false, false>(bn, ta, tb, p, s);
}

namespace llb {
// out[m, n..n+7] = epilogue(sum over splits of ws[s][m][n..] + bias): the second half of a split-K GEMM.  Same rounding
// points as the fused epilogues: y = bf16(acc + bias), then bf16(res + y).
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int k_splits, int M, int N, __nv_bfloat16* __restrict__ out,
                     int64_t ldo, const __nv_bfloat16* __restrict__ bias, const __nv_bfloat16* res, int64_t ld_res) {
  griddep_wait();
  griddep_launch_dependents();
  const int nvec = N / 8;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(M) * nvec) return;
  const int m = static_cast<int>(idx / nvec), n = static_cast<int>(idx % nvec) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < k_splits; ++s) {
    const float4* src = reinterpret_cast<const float4*>(ws + (static_cast<int64_t>(s) * M + m) * N + n);
    const float4 a = src[0], b = src[1];
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
  }
  if (bias != nullptr) {
    const uint4 b4 = *reinterpret_cast<const uint4*>(bias + n);
    const uint32_t* bw = reinterpret_cast<const uint32_t*>(&b4);
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc[2 * e] += bf16_lo(bw[e]); acc[2 * e + 1] += bf16_hi(bw[e]); }
  }
  uint32_t o[4];
  if (res != nullptr) {
    const uint4 r4 = *reinterpret_cast<const uint4*>(res + static_cast<int64_t>(m) * ld_res + n);
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(&r4);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o[e] = pack_bf16x2(bf16_lo(rw[e]) + bf16_round(acc[2 * e]), bf16_hi(rw[e]) + bf16_round(acc[2 * e + 1]));
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = pack_bf16x2(acc[2 * e], acc[2 * e + 1]);
  }
  *reinterpret_cast<uint4*>(out + static_cast<int64_t>(m) * ldo + n) = make_uint4(o[0], o[1], o[2], o[3]);
}
// Row-per-CTA form of the reduction that also applies the RMS norm that follows the projection (T5LayerNorm /
// WanRMSNorm arithmetic: bf16(bf16(x * rstd) * w), statistics over the bf16-rounded x): x goes to `out`, the
// normalised row to `norm_out`, and the separate norm launch disappears.  N <= 256 * 4 * 8 = 8192.
__global__ void __launch_bounds__(256)
splitk_reduce_norm_kernel(const float* __restrict__ ws, int k_splits, int M, int N, __nv_bfloat16* __restrict__ out,
                          int64_t ldo, const __nv_bfloat16* __restrict__ bias, const __nv_bfloat16* res, int64_t ld_res,
                          const __nv_bfloat16* __restrict__ norm_w, __nv_bfloat16* __restrict__ norm_out,
                          int64_t ld_norm, float eps) {
  __shared__ float red[8];
  griddep_wait();
  griddep_launch_dependents();
  const int m = blockIdx.x;
  const int nvec = N / 8;
  uint4 xv[4];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      const int n = vi * 8;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int s = 0; s < k_splits; ++s) {
        const float4* src = reinterpret_cast<const float4*>(ws + (static_cast<int64_t>(s) * M + m) * N + n);
        const float4 a = src[0], b = src[1];
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
        acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
      }
      if (bias != nullptr) {
        const uint4 b4 = *reinterpret_cast<const uint4*>(bias + n);
        const uint32_t* bw = reinterpret_cast<const uint32_t*>(&b4);
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc[2 * e] += bf16_lo(bw[e]); acc[2 * e + 1] += bf16_hi(bw[e]); }
      }
      uint32_t o[4];
      if (res != nullptr) {
        const uint4 r4 = *reinterpret_cast<const uint4*>(res + static_cast<int64_t>(m) * ld_res + n);
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(&r4);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          o[e] = pack_bf16x2(bf16_lo(rw[e]) + bf16_round(acc[2 * e]), bf16_hi(rw[e]) + bf16_round(acc[2 * e + 1]));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = pack_bf16x2(acc[2 * e], acc[2 * e + 1]);
      }
      xv[i] = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(out + static_cast<int64_t>(m) * ldo + n) = xv[i];
#pragma unroll
      for (int e = 0; e < 4; ++e) ss += bf16_lo(o[e]) * bf16_lo(o[e]) + bf16_hi(o[e]) * bf16_hi(o[e]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float rstd = rsqrtf(tot / static_cast<float>(N) + eps);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&xv[i]);
      const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(norm_w) + vi);
      const uint32_t* g = reinterpret_cast<const uint32_t*>(&g4);
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        o[e] = pack_bf16x2(bf16_round(bf16_lo(w[e]) * rstd) * bf16_lo(g[e]),
                           bf16_round(bf16_hi(w[e]) * rstd) * bf16_hi(g[e]));
      *reinterpret_cast<uint4*>(norm_out + static_cast<int64_t>(m) * ld_norm + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}
}  // namespace llb

extern "C" int llb_gemm_bf16_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                                    int M, int N, int K, int epilogue, const void* bias, const void* res,
                                    int64_t ld_res, int k_splits, void* workspace, int64_t workspace_bytes,
                                    const void* norm_w, void* norm_out, int64_t ld_norm, float norm_eps,
                                    void* stream) {
  using namespace llb;
  LLB_CHECK_ARG(epilogue == LLB_EPI_BIAS || epilogue == LLB_EPI_BIAS_RES, "gemm_splitk: epilogue %d unsupported", epilogue);
  LLB_CHECK_ARG(out && workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && N % 8 == 0 && ldo % 8 == 0 &&
                    workspace_bytes >= static_cast<int64_t>(k_splits) * M * N * 4,
                "gemm_splitk: needs a 16-byte aligned workspace of k_splits * M * N floats");
  LLB_CHECK_ARG((epilogue == LLB_EPI_BIAS_RES) == (res != nullptr) && ld_res % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(res) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "gemm_splitk: residual / alignment");
  int rc = gemm_impl(false, A, lda, W, ldw, workspace, N, M, N, K, LLB_EPI_BIAS_F32, nullptr, nullptr, 0, 0, 0, nullptr,
                     0, nullptr, nullptr, stream, k_splits);
  if (rc) return rc;
  if (norm_out != nullptr) {
    LLB_CHECK_ARG(norm_w != nullptr && N <= 8192 && ld_norm % 8 == 0 && (reinterpret_cast<uintptr_t>(norm_out) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(norm_w) & 15) == 0,
                  "gemm_splitk: fused norm needs norm_w, N <= 8192 and 16-byte aligned rows");
    LLB_CUDA(launch_ex(splitk_reduce_norm_kernel, dim3(static_cast<unsigned>(M)), dim3(256), 0,
                       static_cast<cudaStream_t>(stream), 1, true, static_cast<const float*>(workspace), k_splits, M, N,
                       static_cast<__nv_bfloat16*>(out), ldo, static_cast<const __nv_bfloat16*>(bias),
                       static_cast<const __nv_bfloat16*>(res), ld_res, static_cast<const __nv_bfloat16*>(norm_w),
                       static_cast<__nv_bfloat16*>(norm_out), ld_norm, norm_eps));
    LLB_LAUNCH_CHECK("splitk_reduce_norm_kernel");
    return LLB_OK;
  }
  const int64_t total = static_cast<int64_t>(M) * (N / 8);
  LLB_CUDA(launch_ex(splitk_reduce_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
                     static_cast<cudaStream_t>(stream), 1, true, static_cast<const float*>(workspace), k_splits, M, N,
                     static_cast<__nv_bfloat16*>(out), ldo, static_cast<const __nv_bfloat16*>(bias),
                     static_cast<const __nv_bfloat16*>(res), ld_res));
  LLB_LAUNCH_CHECK("splitk_reduce_kernel");
  return LLB_OK;
}

extern "C" int llb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out,
                             int64_t ldo, int M, int N, int K, int epilogue, const void* bias,
                             const void* gate, int64_t ld_gate, int rows_per_gate, int gate_row0,
                             const void* res, int64_t ld_res, void* stream) {
  return gemm_impl(false, A, lda, W, ldw, out, ldo, M, N, K, epilogue, bias, gate, ld_gate, rows_per_gate,
                   gate_row0, res, ld_res, nullptr, nullptr, stream);
}

extern "C" int llb_gemm_fp8(const void* A8, int64_t lda, const float* a_scale, const void* W8, int64_t ldw,
                            const float* w_scale, void* out, int64_t ldo, int M, int N, int K, int epilogue,
                            const void* bias, const void* gate, int64_t ld_gate, int rows_per_gate,
                            int gate_row0, const void* res, int64_t ld_res, void* stream) {
  return gemm_impl(true, A8, lda, W8, ldw, out, ldo, M, N, K, epilogue, bias, gate, ld_gate, rows_per_gate,
                   gate_row0, res, ld_res, a_scale, w_scale, stream);
}

// Kernels of the umT5 text encoder (reference wan/modules/t5.py, called from
// WanTextEncoder.forward, utils/wan_wrapper.py:43-57).  The encoder's FLOPs are its seven Linears per block
// (98 %), which run on llb_gemm_bf16 (tcgen05); this file holds what is left:
//   llb_embed_rows     token-embedding gather                                       (t5.py:288)
//   llb_t5_attn        per-head attention with relative-position bias + key mask    (t5.py:96-111)
//   llb_t5_final_norm  final T5LayerNorm fused with the zeroing of the padding rows (t5.py:294, wan_wrapper.py:52-53)
// The wide-row RMS norm used by the blocks lives in row_kernels.cu (llb_rmsnorm, C > 2048 path).
#include <cuda_bf16.h>
#include <math_constants.h>
#include <stdlib.h>

#include "llb_common.cuh"
#include "llb_host.h"

namespace llb {

// ------------------------------------------------------------------------------------------------
// Embedding gather: out[b * rows_per_seq + r, :] = table[ids[b * ld_ids + r], :]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_rows_kernel(const __nv_bfloat16* __restrict__ table, int64_t vocab, const int64_t* __restrict__ ids,
                  int64_t ld_ids, __nv_bfloat16* __restrict__ out, int64_t ldo, int rows_per_seq, int C) {
  const int r = blockIdx.x, b = blockIdx.y;
  int64_t id = ids[static_cast<int64_t>(b) * ld_ids + r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // nn.Embedding would raise; clamp instead of faulting
  const uint4* src = reinterpret_cast<const uint4*>(table + id * C);
  uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(b) * rows_per_seq + r) * ldo);
  for (int i = threadIdx.x; i < C / 8; i += blockDim.x) dst[i] = __ldg(src + i);
}

// ------------------------------------------------------------------------------------------------
// Final norm + padding: out[b, r, :] = r < seq_lens[b] ? T5LayerNorm(x[b * rows_per_seq + r]) : 0
// for r < rows_out (the wrapper's `u[v:] = 0.0`; rows the encoder never computed are zero as well).
// ------------------------------------------------------------------------------------------------
constexpr int kWideThreads = 256;
constexpr int kWideVec = 4;  // C <= 256 * 4 * 8 = 8192

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = lane < kWideThreads / 32 ? red[lane] : 0.f;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return __shfl_sync(0xffffffffu, t, 0);
}

__global__ void __launch_bounds__(kWideThreads)
t5_final_norm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
                     int64_t ldo, int rows_per_seq, int rows_out, int C, const __nv_bfloat16* __restrict__ wgt,
                     float eps, const int32_t* __restrict__ seq_lens) {
  __shared__ float red[kWideThreads / 32];
  const int r = blockIdx.x, b = blockIdx.y;
  const int nvec = C / 8;
  uint4* orow = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(b) * rows_out + r) * ldo);
  const int len = seq_lens[b];
  if (r >= len || r >= rows_per_seq) {  // uniform per block
    for (int i = threadIdx.x; i < nvec; i += kWideThreads) orow[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const uint4* xr = reinterpret_cast<const uint4*>(x + (static_cast<int64_t>(b) * rows_per_seq + r) * ldx);
  uint4 v[kWideVec];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kWideVec; ++i) {
    const int vi = threadIdx.x + i * kWideThreads;
    if (vi < nvec) {
      v[i] = xr[vi];
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = bf16_lo(w[e]), c = bf16_hi(w[e]);
        ss += a * a + c * c;
      }
    }
  }
  const float rstd = rsqrtf(block_sum_256(ss, red) / static_cast<float>(C) + eps);
#pragma unroll
  for (int i = 0; i < kWideVec; ++i) {
    const int vi = threadIdx.x + i * kWideThreads;
    if (vi < nvec) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[i]);
      const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(wgt) + vi);
      const uint32_t* g = reinterpret_cast<const uint32_t*>(&g4);
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        o[e] = pack_bf16x2(bf16_round(bf16_lo(w[e]) * rstd) * bf16_lo(g[e]),
                           bf16_round(bf16_hi(w[e]) * rstd) * bf16_hi(g[e]));
      orow[vi] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// T5Attention core (t5.py:96-111) for head_dim 64 on the 5th-generation tensor cores:
//   logits = bf16(q . k)                      (einsum in bf16, fp32 accumulate; T5 does not scale)
//   logits = bf16(logits + bias[h][key - query])    bias = per-block bucket table, buckets from a host LUT
//   masked keys (key >= seq_len)  ->  probability exactly 0  (reference: finfo.min before the fp32 softmax)
//   out = bf16(bf16(softmax_fp32(logits)) V)
// (A register-level mma.sync.m16n8k16 version of this kernel was written first: 32 us per launch at 512 rows; the
// tcgen05 version below started at the same time - both bound by the per-element softmax work of 128-512 threads
// per SM -, reached 23 us with the packed bias / rounding pass and replaced it.)
// ------------------------------------------------------------------------------------------------
constexpr int kT5D = 64;

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bf16(a) + bf16(b) -> bf16, two lanes at once (add.rn.bf16x2: one rounding of the exact sum, which is what the
// reference's bf16 tensor add yields)
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// ------------------------------------------------------------------------------------------------
// 512 keys x fp32 is exactly the 512 TMEM columns of an SM,
// so one CTA (128 query rows of one head) gets the WHOLE logit matrix of its rows in one go:
//   TMA:   Q [128 x 64], K and V [kv x 64] of the head as 128-row SWIZZLE_128B boxes straight out of the fused
//          q|k|v projection output (one tensor map, three column offsets)
//   MMA 1: S = Q K^T                tcgen05.mma SS, 4 x (kv / 128) instructions, N = 128 each -> TMEM columns [0, kv)
//   softmax warps (thread = row), three passes over TMEM, no recomputation:
//          1  x = bf16(bf16(s) + bias), masked -> -inf, row max; x written back over s
//          2  e = exp2((x - max) log2e), row sum;               e written back over x
//          3  P = bf16(e / sum), packed two per column, written from column 0 upwards (chunk c of P lands in the
//             columns of chunk c / 2 of e, which is already consumed)
//   MMA 2: O = P V                  tcgen05.mma TS (P from TMEM, V as MN-major operand), kv / 16 instructions, N = 64,
//          accumulator in columns [256, 320)
// Rows of V at or beyond the prompt's length are zeroed in shared memory before MMA 2 (P = 0 there, but 0 x NaN
// is NaN); masked K rows only produce logits that are replaced by -inf, never used in arithmetic.
// ------------------------------------------------------------------------------------------------
constexpr int kTcThreads = 160;  // 4 softmax warps (TMEM lane quadrants) + 1 TMA / MMA warp

struct T5TcParams {
  __nv_bfloat16* out;
  int64_t ldo;
  int rows_per_seq, n_heads, kv_cap;  // kv_cap: multiple of 128
  const int32_t* seq_lens;
  const __nv_bfloat16* pos_emb;
  const int32_t* bucket_lut;
  int lut_center;
};

__global__ void __launch_bounds__(kTcThreads, 1)
t5_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap, const T5TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // [0, 1024): barriers + TMEM slot; then Q (16 KB), K, V (kv_cap x 128 B each), bias table
  const uint32_t bar_load = base, bar_s = base + 8, bar_p = base + 16, bar_o = base + 24, bar_v = base + 32;
  const uint32_t tmem_slot = base + 64;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 64);
  const uint32_t q_s = base + 1024;
  const uint32_t k_s = q_s + 16384;
  const uint32_t v_s = k_s + static_cast<uint32_t>(p.kv_cap) * 128u;
  uint8_t* v_gen = gen + 1024 + 16384 + static_cast<size_t>(p.kv_cap) * 128;
  // bias2[i] = (bias[i], bias[i + 1]) as bf16x2 for offset index i = key - query + Lp - 1: one aligned 32-bit read
  // serves two adjacent keys whatever the parity of i
  uint32_t* bias2 = reinterpret_cast<uint32_t*>(v_gen + static_cast<size_t>(p.kv_cap) * 128);  // [2 * Lp - 1]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Lp = p.rows_per_seq;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int64_t seq_row0 = static_cast<int64_t>(b) * Lp;
  int kv_len = p.seq_lens[b];
  kv_len = kv_len < 0 ? 0 : (kv_len > p.kv_cap ? p.kv_cap : kv_len);
  if (kv_len == 0) {  // nothing to attend (rows are zeroed downstream anyway)
    for (int i = tid; i < 128 * 8; i += kTcThreads)
      *reinterpret_cast<uint4*>(p.out + (seq_row0 + qb * 128 + (i >> 3)) * p.ldo + h * kT5D + (i & 7) * 8) =
          make_uint4(0, 0, 0, 0);
    return;
  }
  const int kv_tiles = (kv_len + 127) >> 7;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap);
      mbar_init(bar_load, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 4);
      mbar_init(bar_o, 1);
      mbar_init(bar_v, 4);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA + MMA issue (whole warp, one lane acts)
    const int width = p.n_heads * kT5D;
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_load, static_cast<uint32_t>(1 + 2 * kv_tiles) * 16384u);
      tma_load_2d(q_s, &tmap, bar_load, h * kT5D, static_cast<int>(seq_row0) + qb * 128);
      for (int i = 0; i < kv_tiles; ++i) {
        tma_load_2d(k_s + i * 16384, &tmap, bar_load, width + h * kT5D, static_cast<int>(seq_row0) + i * 128);
        tma_load_2d(v_s + i * 16384, &tmap, bar_load, 2 * width + h * kT5D, static_cast<int>(seq_row0) + i * 128);
      }
    }
    __syncwarp();
    mbar_wait(bar_load, 0);
    tc_fence_after();
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kT5D, 0, 1);
    if (elect_one()) {
      for (int i = 0; i < kv_tiles; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k)  // head dim 64 = 4 x K16; +32 bytes inside the 128-byte swizzle row
          umma_ss(tmem_base + i * 128, umma_desc_kmajor(q_s + k * 32), umma_desc_kmajor(k_s + i * 16384 + k * 32),
                  idesc_qk, k != 0);
      }
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_v, 0);  // V tail rows zeroed (generic-proxy writes fenced by the writers)
    mbar_wait(bar_p, 0);  // P complete in TMEM
    tc_fence_after();
    if (elect_one()) {
      for (int kk = 0; kk < kv_tiles * 8; ++kk)  // 16 keys per instruction
        umma_ts(tmem_base + 256, tmem_base + kk * 8, umma_desc_mnmajor(v_s + kk * 2048, 16384), idesc_pv, kk != 0);
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps, thread = query row
    const int row = qb * 128 + warp * 32 + lane;  // within the sequence
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    {
      const unsigned short* emb = reinterpret_cast<const unsigned short*>(p.pos_emb);
      for (int i = tid; i < 2 * Lp - 1; i += 128) {
        const int d = i - (Lp - 1) + p.lut_center;
        const uint32_t lo = emb[p.bucket_lut[d] * p.n_heads + h];
        const uint32_t hi = i + 1 < 2 * Lp - 1 ? emb[p.bucket_lut[d + 1] * p.n_heads + h] : 0u;
        bias2[i] = lo | (hi << 16);
      }
    }
    mbar_wait(bar_load, 0);
    {
      const int r = kv_len + tid;  // one row per thread covers the at most 127 tail rows of the last box
      if (r < kv_tiles * 128) {
        uint4* vr = reinterpret_cast<uint4*>(v_gen + static_cast<size_t>(r) * 128);
#pragma unroll
        for (int i = 0; i < 8; ++i) vr[i] = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async_smem();
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // bias table complete
    if (lane == 0) mbar_arrive(bar_v);
    const uint32_t* brow = bias2 + (Lp - 1 - row);  // + (even) key column
    constexpr float kLog2e = 1.4426950408889634f;
    const int nchunks = kv_tiles * 4;

    mbar_wait(bar_s, 0);
    tc_fence_after();
    // Every pass streams the row through two register buffers: the tcgen05.ld of the next 32-column chunk is in
    // flight while the current one is processed (one warp per scheduler, nothing else hides that latency).
    uint32_t va[32], vb[32];
    // pass 1: logits with bias and mask, the reference's two bf16 roundings, row max
    float m = -CUDART_INF_F;
    auto pass1 = [&](uint32_t (&v)[32], int c) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {  // two keys at a time: cvt.rn.bf16x2, add.rn.bf16x2 against the paired bias
        const uint32_t x2 = add_bf16x2(pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])),
                                       brow[c * 32 + 2 * i]);
        v[2 * i] = x2 << 16;
        v[2 * i + 1] = x2 & 0xffff0000u;
      }
      if (c * 32 + 32 > kv_len) {  // only the chunk that straddles the prompt length (and those beyond it)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i >= kv_len) v[i] = 0xff800000u;  // -inf
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) m = fmaxf(m, fmaxf(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])));
      tmem_st32(t_row + c * 32, v);
    };
    tmem_ld32(t_row, va);
    tmem_wait_ld();
#pragma unroll 1
    for (int c = 0; c < nchunks; c += 2) {  // nchunks is a multiple of 4
      tmem_ld32(t_row + (c + 1) * 32, vb);
      pass1(va, c);
      tmem_wait_ld();
      if (c + 2 < nchunks) tmem_ld32(t_row + (c + 2) * 32, va);
      pass1(vb, c + 1);
      tmem_wait_ld();
    }
    tmem_wait_st();
    // pass 2: exponentials (fp32, like F.softmax(attn.float())) and the row sum
    const float mneg = -m * kLog2e;
    float l = 0.f;
    auto pass2 = [&](uint32_t (&v)[32], int c) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = ex2_ftz(fmaf(__uint_as_float(v[i]), kLog2e, mneg));
        l += e;
        v[i] = __float_as_uint(e);
      }
      tmem_st32(t_row + c * 32, v);
    };
    tmem_ld32(t_row, va);
    tmem_wait_ld();
#pragma unroll 1
    for (int c = 0; c < nchunks; c += 2) {
      tmem_ld32(t_row + (c + 1) * 32, vb);
      pass2(va, c);
      tmem_wait_ld();
      if (c + 2 < nchunks) tmem_ld32(t_row + (c + 2) * 32, va);
      pass2(vb, c + 1);
      tmem_wait_ld();
    }
    tmem_wait_st();
    // pass 3: P = bf16(softmax), two keys per column, compacted to columns [0, kv / 2): chunk c lands in the columns
    // of chunk c / 2, which is consumed (the chunk in flight is c + 1)
    const float inv = 1.f / l;
    auto pass3 = [&](const uint32_t (&v)[32], int c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * inv, __uint_as_float(v[2 * i + 1]) * inv);
      tmem_st16(t_row + c * 16, pk);
    };
    tmem_ld32(t_row, va);
    tmem_wait_ld();
#pragma unroll 1
    for (int c = 0; c < nchunks; c += 2) {
      tmem_ld32(t_row + (c + 1) * 32, vb);
      pass3(va, c);
      tmem_wait_ld();
      if (c + 2 < nchunks) tmem_ld32(t_row + (c + 2) * 32, va);
      pass3(vb, c + 1);
      tmem_wait_ld();
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);

    // epilogue: O [128 x 64] fp32 -> bf16 -> out
    mbar_wait(bar_o, 0);
    tc_fence_after();
    __nv_bfloat16* orow = p.out + (seq_row0 + row) * p.ldo + h * kT5D;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(t_row + 256 + c * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]));
        w.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
        w.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
        w.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
        *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace llb

extern "C" int llb_embed_rows(const void* table, int64_t vocab, const void* ids, int64_t ld_ids, void* out,
                              int64_t ldo, int batch, int rows_per_seq, int C, void* stream) {
  using namespace llb;
  LLB_CHECK_ARG(table && ids && out && vocab > 0 && batch > 0 && rows_per_seq > 0, "embed_rows: bad arguments");
  LLB_CHECK_ARG(C > 0 && C % 8 == 0 && ldo % 8 == 0, "embed_rows: C / ldo must be multiples of 8");
  embed_rows_kernel<<<dim3(rows_per_seq, batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(table), vocab, static_cast<const int64_t*>(ids), ld_ids,
      static_cast<__nv_bfloat16*>(out), ldo, rows_per_seq, C);
  LLB_LAUNCH_CHECK("embed_rows_kernel");
  return LLB_OK;
}

extern "C" int llb_t5_final_norm(const void* x, int64_t ldx, void* out, int64_t ldo, int batch, int rows_per_seq,
                                 int rows_out, int C, const void* w, float eps, const int32_t* seq_lens_dev,
                                 void* stream) {
  using namespace llb;
  LLB_CHECK_ARG(x && out && w && seq_lens_dev && batch > 0 && rows_per_seq > 0 && rows_out > 0,
                "t5_final_norm: bad arguments");
  LLB_CHECK_ARG(C % 8 == 0 && C <= kWideThreads * kWideVec * 8 && ldx % 8 == 0 && ldo % 8 == 0,
                "t5_final_norm: C=%d unsupported", C);
  t5_final_norm_kernel<<<dim3(rows_out, batch), kWideThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, rows_per_seq, rows_out, C,
      static_cast<const __nv_bfloat16*>(w), eps, seq_lens_dev);
  LLB_LAUNCH_CHECK("t5_final_norm_kernel");
  return LLB_OK;
}

extern "C" int llb_t5_attn(const void* qkv, int64_t ld_qkv, void* out, int64_t ldo, int batch, int rows_per_seq,
                           int n_heads, int max_seq_len, const int32_t* seq_lens_dev, const void* pos_emb,
                           const int32_t* bucket_lut_dev, int lut_center, void* stream) {
  using namespace llb;
  LLB_CHECK_ARG(qkv && out && seq_lens_dev && pos_emb && bucket_lut_dev && batch > 0 && n_heads > 0,
                "t5_attn: null tensor / bad shape");
  LLB_CHECK_ARG(rows_per_seq > 0 && rows_per_seq % 128 == 0 && rows_per_seq <= 512,
                "t5_attn: rows_per_seq=%d must be a multiple of 128 (<= 512: the logits of a row fill TMEM)", rows_per_seq);
  LLB_CHECK_ARG(lut_center >= rows_per_seq - 1, "t5_attn: bucket LUT too short for %d rows", rows_per_seq);
  LLB_CHECK_ARG(ld_qkv % 8 == 0 && ldo % 8 == 0 && ld_qkv >= 3 * n_heads * kT5D && ldo >= n_heads * kT5D,
                "t5_attn: leading dimensions");
  LLB_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "t5_attn: 16-byte alignment");
  // the logits of 128 query rows against up to 512 keys fill the 512 TMEM columns exactly; shared memory holds the
  // valid keys only (max_seq_len: host-known upper bound of seq_lens, <= 0 = rows_per_seq)
  {
    int cap = max_seq_len > 0 && max_seq_len < rows_per_seq ? max_seq_len : rows_per_seq;
    cap = (cap + 127) & ~127;
    const size_t smem_tc = 1024 + 1024 + 16384 + static_cast<size_t>(cap) * 256 + (2 * rows_per_seq - 1) * sizeof(float);
    static std::atomic<size_t> attr_tc[64];  // largest opt-in so far, per device ordinal (the attribute is per device)
    int dev = 0;
    LLB_CUDA(cudaGetDevice(&dev));
    if (smem_tc > attr_tc[dev & 63].load(std::memory_order_acquire)) {
      LLB_CUDA(cudaFuncSetAttribute(t5_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_tc)));
      attr_tc[dev & 63].store(smem_tc, std::memory_order_release);
    }
    CUtensorMap tm;
    int rc = make_tmap_2d_bf16(&tm, qkv, static_cast<uint64_t>(batch) * rows_per_seq, static_cast<uint64_t>(3) * n_heads * kT5D,
                               ld_qkv, 128, 64);
    if (rc) return rc;
    T5TcParams tp;
    tp.out = static_cast<__nv_bfloat16*>(out); tp.ldo = ldo;
    tp.rows_per_seq = rows_per_seq; tp.n_heads = n_heads; tp.kv_cap = cap;
    tp.seq_lens = seq_lens_dev;
    tp.pos_emb = static_cast<const __nv_bfloat16*>(pos_emb);
    tp.bucket_lut = bucket_lut_dev; tp.lut_center = lut_center;
    t5_attn_tc_kernel<<<dim3(rows_per_seq / 128, n_heads, batch), kTcThreads, smem_tc, static_cast<cudaStream_t>(stream)>>>(tm, tp);
    LLB_LAUNCH_CHECK("t5_attn_tc_kernel");
    return LLB_OK;
  }
}


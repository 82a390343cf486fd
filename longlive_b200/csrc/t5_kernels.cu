// Kernels of the umT5 text encoder (reference wan/modules/t5.py, called from
// WanTextEncoder.forward, utils/wan_wrapper.py:43-57).  The encoder's FLOPs are its seven Linears per block
// (98 %), which run on llb_gemm_bf16 (tcgen05); this file holds what is left:
//   llb_embed_rows     token-embedding gather                                       (t5.py:288)
//   llb_t5_attn        per-head attention with relative-position bias + key mask    (t5.py:96-111)
//   llb_t5_final_norm  final T5LayerNorm fused with the zeroing of the padding rows (t5.py:294, wan_wrapper.py:52-53)
// The wide-row RMS norm used by the blocks lives in row_kernels.cu (llb_rmsnorm, C > 2048 path).
#include <cuda_bf16.h>
#include <math_constants.h>

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
// T5Attention core (t5.py:96-111) for head_dim 64:
//   logits = bf16(q . k)                      (einsum in bf16, fp32 accumulate; T5 does not scale)
//   logits = bf16(logits + bias[h][key - query])    bias = per-block bucket table, buckets from a host LUT
//   masked keys (key >= seq_len)  ->  probability exactly 0  (reference: finfo.min before the fp32 softmax)
//   out = bf16(softmax_fp32(logits) V)
// This is 2 % of the encoder's FLOPs (4.3 of 202 GFLOP per block at 512 tokens), sequence length <= 512 and
// needs a per-element bias, so it is a register-level flash kernel on mma.sync.m16n8k16 rather than a TMEM
// pipeline: one CTA = 128 query rows of one (batch, head), 8 warps x 16 rows; all keys / values of the head
// are staged once in shared memory (row stride 144 B: conflict-free for both the 32-bit K fragment loads
// and ldmatrix.trans on V).  Two passes over the keys, S recomputed in the second: pass 1 yields the fp32 row
// maximum and denominator, pass 2 forms P = bf16(softmax) - the tensor the reference materialises - and O += P V,
// so the rounding points are the reference's (an online softmax would round un-normalised probabilities).
// ------------------------------------------------------------------------------------------------
constexpr int kT5D = 64;
constexpr int kT5Stride = 72;

struct T5AttnParams {
  const __nv_bfloat16* qkv;  // [batch * rows_per_seq, ld]: q | k | v column blocks of width n_heads * 64
  int64_t ld;
  __nv_bfloat16* out;        // [batch * rows_per_seq, ldo]
  int64_t ldo;
  int rows_per_seq, n_heads;
  int kv_cap;                     // shared-memory capacity in keys (multiple of 64, >= every valid length)
  const int32_t* seq_lens;        // [batch] valid keys per sequence
  const __nv_bfloat16* pos_emb;   // [num_buckets, n_heads]
  const int32_t* bucket_lut;      // [2 * lut_center + 1]: bucket of (key - query) + lut_center
  int lut_center;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t saddr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(saddr));
}
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

// kWarps x 16 query rows per CTA (8 warps = 128 rows, or 16 warps = 256 rows when that still fills the GPU:
// the keys / values of the head are then staged once per 256 rows and a 512-token prompt is a single wave).
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
t5_attn_kernel(const T5AttnParams p) {
  constexpr int kThreads = kWarps * 32, kRows = kWarps * 16;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int Lp = p.rows_per_seq;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Vs = Ks + static_cast<size_t>(p.kv_cap) * kT5Stride;
  // bias2[i] = (bias[i], bias[i + 1]) as bf16x2 for offset index i = key - query + Lp - 1: one aligned 32-bit
  // read serves the two adjacent keys of an accumulator pair whatever the parity of i
  uint32_t* bias2 = reinterpret_cast<uint32_t*>(Vs + static_cast<size_t>(p.kv_cap) * kT5Stride);  // [2 * Lp - 1]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int width = p.n_heads * kT5D;
  const int64_t seq_row0 = static_cast<int64_t>(b) * Lp;
  const int row_cta = blockIdx.x * kRows;
  const int row_w = row_cta + warp * 16;  // first query row of this warp (within the sequence)

  int kv_len = p.seq_lens[b];
  kv_len = kv_len < 0 ? 0 : (kv_len > p.kv_cap ? p.kv_cap : kv_len);
  const int kv_pad = (kv_len + 63) & ~63;  // <= kv_cap (a multiple of 64)

  if (kv_len == 0) {  // nothing to attend: the reference would produce a uniform average; rows are zeroed later anyway
    for (int i = tid; i < kRows * (kT5D / 8); i += kThreads) {
      const int r = i / (kT5D / 8), c = i % (kT5D / 8);
      *reinterpret_cast<uint4*>(p.out + (seq_row0 + row_cta + r) * p.ldo + h * kT5D + c * 8) = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  // ---- stage K, V (rows of masked keys zeroed: 0 * garbage must stay 0) and this head's bias pairs
  {
    const __nv_bfloat16* kbase = p.qkv + seq_row0 * p.ld + width + h * kT5D;
    const __nv_bfloat16* vbase = kbase + width;
    for (int i = tid; i < kv_pad * (kT5D / 8); i += kThreads) {
      const int key = i >> 3, c = i & 7;
      uint4 kk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (key < kv_len) {
        kk = *reinterpret_cast<const uint4*>(kbase + static_cast<int64_t>(key) * p.ld + c * 8);
        vv = *reinterpret_cast<const uint4*>(vbase + static_cast<int64_t>(key) * p.ld + c * 8);
      }
      *reinterpret_cast<uint4*>(Ks + key * kT5Stride + c * 8) = kk;
      *reinterpret_cast<uint4*>(Vs + key * kT5Stride + c * 8) = vv;
    }
    const unsigned short* emb = reinterpret_cast<const unsigned short*>(p.pos_emb);
    for (int i = tid; i < 2 * Lp - 1; i += kThreads) {
      const int d = i - (Lp - 1) + p.lut_center;
      const uint32_t lo = emb[p.bucket_lut[d] * p.n_heads + h];
      const uint32_t hi = i + 1 < 2 * Lp - 1 ? emb[p.bucket_lut[d + 1] * p.n_heads + h] : 0u;
      bias2[i] = lo | (hi << 16);
    }
  }

  // ---- Q fragments straight from global memory (A operand, row-major 16 x 16 per k-chunk)
  uint32_t qa[4][4];
  {
    const __nv_bfloat16* q0 = p.qkv + (seq_row0 + row_w + g) * p.ld + h * kT5D + 2 * t;
    const __nv_bfloat16* q1 = q0 + 8 * p.ld;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      qa[kc][0] = *reinterpret_cast<const uint32_t*>(q0 + kc * 16);
      qa[kc][1] = *reinterpret_cast<const uint32_t*>(q1 + kc * 16);
      qa[kc][2] = *reinterpret_cast<const uint32_t*>(q0 + kc * 16 + 8);
      qa[kc][3] = *reinterpret_cast<const uint32_t*>(q1 + kc * 16 + 8);
    }
  }
  __syncthreads();

  constexpr float kLog2e = 1.4426950408889634f;
  // S block (16 query rows x 64 keys per warp): bf16(bf16(q . k) + bias) as two packed roundings, keys beyond the
  // prompt's length set to -inf (only the block that straddles kv_len needs the comparison)
  const uint32_t* bias_r0 = bias2 + (2 * t - (row_w + g) + Lp - 1);  // + key column base; row g
  const uint32_t* bias_r1 = bias_r0 - 8;                              // row g + 8
  auto logits_block = [&](int kb, float (&s)[8][4]) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const __nv_bfloat16* krow = Ks + (kb + nt * 8 + g) * kT5Stride + 2 * t;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + kc * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + kc * 16 + 8);
        mma_bf16_16816(s[nt], qa[kc], b0, b1);
      }
      const uint32_t x0 = add_bf16x2(pack_bf16x2(s[nt][0], s[nt][1]), bias_r0[kb + nt * 8]);
      const uint32_t x1 = add_bf16x2(pack_bf16x2(s[nt][2], s[nt][3]), bias_r1[kb + nt * 8]);
      s[nt][0] = bf16_lo(x0); s[nt][1] = bf16_hi(x0);
      s[nt][2] = bf16_lo(x1); s[nt][3] = bf16_hi(x1);
    }
    if (kb + 64 > kv_len) {  // warp-uniform
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = kb + nt * 8 + 2 * t;
        if (col >= kv_len) s[nt][0] = s[nt][2] = -CUDART_INF_F;
        if (col + 1 >= kv_len) s[nt][1] = s[nt][3] = -CUDART_INF_F;
      }
    }
  };

  // ---- pass 1: row maximum and softmax denominator (fp32, like F.softmax(attn.float()))
  float m_run[2] = {-CUDART_INF_F, -CUDART_INF_F};
  float l_run[2] = {0.f, 0.f};
  for (int kb = 0; kb < kv_pad; kb += 64) {
    float s[8][4];
    logits_block(kb, s);
    float mx[2] = {-CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float mneg[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);  // finite from the first block on (key 0 is never masked)
      l_run[r] *= ex2_ftz((m_run[r] - m_new) * kLog2e);
      m_run[r] = m_new;
      mneg[r] = -m_new * kLog2e;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      l_run[0] += ex2_ftz(fmaf(s[nt][0], kLog2e, mneg[0])) + ex2_ftz(fmaf(s[nt][1], kLog2e, mneg[0]));
      l_run[1] += ex2_ftz(fmaf(s[nt][2], kLog2e, mneg[1])) + ex2_ftz(fmaf(s[nt][3], kLog2e, mneg[1]));
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const float mneg0 = -m_run[0] * kLog2e, mneg1 = -m_run[1] * kLog2e;

  // ---- pass 2: P = bf16(softmax) exactly as the reference materialises it (.type_as(attn)), O += P V
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  // ldmatrix.x4.trans lane address inside a 16-key x 16-dim V block: matrices (keys 0-7 | 8-15) x (dims 0-7 | 8-15)
  const uint32_t v_lane = smem_u32(Vs) + static_cast<uint32_t>(((lane & 7) + ((lane >> 3) & 1) * 8) * kT5Stride +
                                                               (lane >> 4) * 8) * 2u;
  for (int kb = 0; kb < kv_pad; kb += 64) {
    float s[8][4];
    logits_block(kb, s);
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = ex2_ftz(fmaf(s[nt][0], kLog2e, mneg0)) * inv0, p1 = ex2_ftz(fmaf(s[nt][1], kLog2e, mneg0)) * inv0;
      const float p2 = ex2_ftz(fmaf(s[nt][2], kLog2e, mneg1)) * inv1, p3 = ex2_ftz(fmaf(s[nt][3], kLog2e, mneg1)) * inv1;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    // k index = key (16 per chunk), n index = head dim (8 per tile, two tiles per ldmatrix.x4)
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t r0, r1, r2, r3;
        ldmatrix_x4_trans(v_lane + static_cast<uint32_t>((kb + kc * 16) * kT5Stride + dp * 16) * 2u, r0, r1, r2, r3);
        mma_bf16_16816(o[2 * dp], pa[kc], r0, r1);
        mma_bf16_16816(o[2 * dp + 1], pa[kc], r2, r3);
      }
    }
  }

  // ---- store (P was normalised before the product, as in the reference)
  __nv_bfloat16* o0 = p.out + (seq_row0 + row_w + g) * p.ldo + h * kT5D + 2 * t;
  __nv_bfloat16* o1 = o0 + 8 * p.ldo;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    *reinterpret_cast<uint32_t*>(o0 + dt * 8) = pack_bf16x2(o[dt][0], o[dt][1]);
    *reinterpret_cast<uint32_t*>(o1 + dt * 8) = pack_bf16x2(o[dt][2], o[dt][3]);
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
  LLB_CHECK_ARG(rows_per_seq > 0 && rows_per_seq % 128 == 0 && rows_per_seq <= 1024,
                "t5_attn: rows_per_seq=%d must be a multiple of 128 (<= 1024)", rows_per_seq);
  LLB_CHECK_ARG(lut_center >= rows_per_seq - 1, "t5_attn: bucket LUT too short for %d rows", rows_per_seq);
  LLB_CHECK_ARG(ld_qkv % 8 == 0 && ldo % 8 == 0 && ld_qkv >= 3 * n_heads * kT5D && ldo >= n_heads * kT5D,
                "t5_attn: leading dimensions");
  LLB_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "t5_attn: 16-byte alignment");
  // shared memory holds the valid keys only (max_seq_len: host-known upper bound of seq_lens, <= 0 = rows_per_seq)
  int kv_cap = max_seq_len > 0 && max_seq_len < rows_per_seq ? max_seq_len : rows_per_seq;
  kv_cap = (kv_cap + 63) & ~63;
  const size_t smem = static_cast<size_t>(kv_cap) * kT5Stride * 2 * 2 + (2 * rows_per_seq - 1) * sizeof(uint32_t);
  LLB_CHECK_ARG(smem <= 227 * 1024, "t5_attn: %d keys do not fit shared memory", kv_cap);
  // 256-row CTAs when they still cover most of the GPU in one wave, else 128-row CTAs
  const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
  const bool wide = rows_per_seq % 256 == 0 && (rows_per_seq / 256) * n_heads * batch >= (sms * 3) / 4;
  static size_t attr_smem[2] = {0, 0};
  if (smem > attr_smem[wide]) {
    if (wide) LLB_CUDA(cudaFuncSetAttribute(t5_attn_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    else LLB_CUDA(cudaFuncSetAttribute(t5_attn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_smem[wide] = smem;
  }
  T5AttnParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv); p.ld = ld_qkv;
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo;
  p.rows_per_seq = rows_per_seq; p.n_heads = n_heads; p.kv_cap = kv_cap;
  p.seq_lens = seq_lens_dev;
  p.pos_emb = static_cast<const __nv_bfloat16*>(pos_emb);
  p.bucket_lut = bucket_lut_dev; p.lut_center = lut_center;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (wide) t5_attn_kernel<16><<<dim3(rows_per_seq / 256, n_heads, batch), 512, smem, st>>>(p);
  else t5_attn_kernel<8><<<dim3(rows_per_seq / 128, n_heads, batch), 256, smem, st>>>(p);
  LLB_LAUNCH_CHECK("t5_attn_kernel");
  return LLB_OK;
}

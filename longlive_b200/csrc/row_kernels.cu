// HBM-bound row kernels of the LongLive block and the small glue kernels of
// CausalWanModel._forward_inference.  One warp owns one token row (C = 1536 -> six 16-byte vectors
// per lane), so every global access is a fully coalesced 128-bit transaction and the row statistics
// are warp-shuffle reductions.  Rounding points follow the reference's bf16 materialisation points
// (see each kernel) so the outputs are bit-comparable to the PyTorch path.
#include "llb_common.cuh"
#include "llb_host.h"

#include <cuda_fp8.h>
#include <string.h>

namespace llb {

constexpr int kRowWarps = 8;  // rows per CTA
constexpr int kMaxVec = 8;    // supports C <= 32 lanes * 8 vecs * 8 elems = 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (+ adaLN modulate)           wan/modules/model.py:89-99, causal_model.py:445,463-464,507
//   modulate mode:  t = bf16(LN(x)); out = bf16(bf16(t * bf16(1 + scale)) + shift)
//   affine mode:    out = bf16(LN(x) * w + b)                       (norm3, elementwise_affine)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// 8 bf16 (one uint4) * inv_scale -> 8 e4m3 bytes (one uint2), round-to-nearest, saturating
__device__ __forceinline__ uint2 quant8_e4m3(const uint4& v, float inv_scale) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
  uint32_t r[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const unsigned short lo = __nv_cvt_float2_to_fp8x2(
        make_float2(bf16_lo(w[2 * h]) * inv_scale, bf16_hi(w[2 * h]) * inv_scale), __NV_SATFINITE, __NV_E4M3);
    const unsigned short hi = __nv_cvt_float2_to_fp8x2(
        make_float2(bf16_lo(w[2 * h + 1]) * inv_scale, bf16_hi(w[2 * h + 1]) * inv_scale), __NV_SATFINITE, __NV_E4M3);
    r[h] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
  }
  return make_uint2(r[0], r[1]);
}
__device__ __forceinline__ float absmax8(const uint4& v) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
  float m = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) m = fmaxf(m, fmaxf(fabsf(bf16_lo(w[e])), fabsf(bf16_hi(w[e]))));
  return m;
}

// kFp8Out: the modulated row is additionally quantised for the FP8 linear that consumes it:
// per-row scale = amax / 448 (e4m3 max), out8 = round(y / scale); the bf16 output is skipped.
template <bool kFp8Out, int NV>
__global__ void __launch_bounds__(kRowWarps * 32, NV <= 6 ? 4 : 2)
ln_modulate_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
                   int64_t ldo, int rows, int C, const __nv_bfloat16* __restrict__ shift,
                   const __nv_bfloat16* __restrict__ scale, int64_t ld_mod, int rows_per_frame,
                   int row0, const __nv_bfloat16* __restrict__ ln_w,
                   const __nv_bfloat16* __restrict__ ln_b, float eps, uint8_t* __restrict__ out8,
                   int64_t ld8, float* __restrict__ out_scale) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's output is complete past here
  griddep_launch_dependents();
  const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = C / 8;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
  uint4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      v[i] = xr[vi];
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) s += bf16_lo(w[e]) + bf16_hi(w[e]);
    }
  }
  // Occupancy beats prefetching here: at 64 registers four CTAs fit an SM and all 4680 rows of a chunk are
  // in flight at once (one wave); the per-frame shift / scale vectors are L1 / L2 hits when they are needed.
  const bool affine = ln_w != nullptr;
  const int64_t mrow = static_cast<int64_t>((row0 + row) / rows_per_frame) * ld_mod;
  const float mean = warp_sum(s) / static_cast<float>(C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = bf16_lo(w[e]) - mean, b = bf16_hi(w[e]) - mean;
        ss += a * a + b * b;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(C) + eps);
  uint4* orow = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo);
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[i]);
      uint4 a4, b4;
      if (affine) {
        a4 = __ldg(reinterpret_cast<const uint4*>(ln_w) + vi);
        b4 = __ldg(reinterpret_cast<const uint4*>(ln_b) + vi);
      } else {
        a4 = __ldg(reinterpret_cast<const uint4*>(scale + mrow) + vi);
        b4 = __ldg(reinterpret_cast<const uint4*>(shift + mrow) + vi);
      }
      const uint32_t* aw = reinterpret_cast<const uint32_t*>(&a4);
      const uint32_t* bw = reinterpret_cast<const uint32_t*>(&b4);
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float n0 = (bf16_lo(w[e]) - mean) * rstd, n1 = (bf16_hi(w[e]) - mean) * rstd;
        float y0, y1;
        if (affine) {
          y0 = n0 * bf16_lo(aw[e]) + bf16_lo(bw[e]);
          y1 = n1 * bf16_hi(aw[e]) + bf16_hi(bw[e]);
        } else {
          const float s0 = bf16_round(1.0f + bf16_lo(aw[e])), s1 = bf16_round(1.0f + bf16_hi(aw[e]));
          y0 = bf16_round(bf16_round(n0) * s0) + bf16_lo(bw[e]);
          y1 = bf16_round(bf16_round(n1) * s1) + bf16_hi(bw[e]);
        }
        o[e] = pack_bf16x2(y0, y1);
      }
      if constexpr (kFp8Out) {
        v[i] = make_uint4(o[0], o[1], o[2], o[3]);  // keep the bf16-rounded result for the quant pass
        amax = fmaxf(amax, absmax8(v[i]));
      } else {
        orow[vi] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  if constexpr (kFp8Out) {
    amax = warp_max(amax);
    const float sc = amax > 0.f ? amax / 448.0f : 1.0f;
    const float inv = 1.0f / sc;
    uint2* o8 = reinterpret_cast<uint2*>(out8 + static_cast<int64_t>(row) * ld8);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) o8[vi] = quant8_e4m3(v[i], inv);
    }
    if (lane == 0) out_scale[row] = sc;
  }
}

// Row-wise dynamic e4m3 quantisation of a bf16 matrix (any C % 8 == 0): two passes over the row.
__global__ void __launch_bounds__(kRowWarps * 32)
quant_rows_fp8_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, uint8_t* __restrict__ out8,
                      int64_t ld8, float* __restrict__ out_scale, int rows, int C) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's output is complete past here
  griddep_launch_dependents();
  const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = C / 8;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
  float amax = 0.f;
  for (int vi = lane; vi < nvec; vi += 32) amax = fmaxf(amax, absmax8(xr[vi]));
  amax = warp_max(amax);
  const float sc = amax > 0.f ? amax / 448.0f : 1.0f;
  const float inv = 1.0f / sc;
  uint2* o8 = reinterpret_cast<uint2*>(out8 + static_cast<int64_t>(row) * ld8);
  for (int vi = lane; vi < nvec; vi += 32) o8[vi] = quant8_e4m3(xr[vi], inv);
  if (lane == 0) out_scale[row] = sc;
}

// ------------------------------------------------------------------------------------------------
// WanRMSNorm row helper (model.py:78-86): y = bf16(bf16(x * rsqrt(mean(x^2) + eps)) * w)
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ float row_sumsq(const uint4 (&v)[NV], int lane, int nvec) {
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + i * 32 < nvec) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = bf16_lo(w[e]), b = bf16_hi(w[e]);
        ss += a * a + b * b;
      }
    }
  }
  return warp_sum(ss);
}

__global__ void __launch_bounds__(kRowWarps * 32)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
               int64_t ldo, int rows, int C, const __nv_bfloat16* __restrict__ wgt, float eps) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's output is complete past here
  griddep_launch_dependents();
  const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = C / 8;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
  uint4 v[kMaxVec];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (lane + i * 32 < nvec) v[i] = xr[lane + i * 32];
  const float rstd = rsqrtf(row_sumsq(v, lane, nvec) / static_cast<float>(C) + eps);
  uint4* orow = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int vi = lane + i * 32;
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

// Rows wider than one warp's register budget (C > 2048: T5LayerNorm of the umT5 text encoder, t5.py:57-62,
// the same arithmetic as WanRMSNorm): one CTA of 256 threads per row, up to four 16-byte vectors per thread.
constexpr int kWideThreads = 256;
constexpr int kWideVec = 4;  // C <= 256 * 4 * 8 = 8192
__global__ void __launch_bounds__(kWideThreads)
rmsnorm_wide_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
                    int64_t ldo, int C, const __nv_bfloat16* __restrict__ wgt, float eps) {
  __shared__ float red[kWideThreads / 32];
  griddep_wait();
  griddep_launch_dependents();
  const int row = blockIdx.x;
  const int nvec = C / 8;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
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
        const float a = bf16_lo(w[e]), b = bf16_hi(w[e]);
        ss += a * a + b * b;
      }
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kWideThreads / 32; ++i) tot += red[i];
  const float rstd = rsqrtf(tot / static_cast<float>(C) + eps);
  uint4* orow = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo);
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
// Fused q/k RMSNorm + 3-D RoPE + KV-ring append.
//   q  = rope(rmsnorm(qkv[:, 0:C]))        -> q_out                    (causal_model.py:122-128, 208-211)
//   k  = rope(rmsnorm(qkv[:, C:2C]))       -> k_cache[phys(row)]       (:268 / :310)
//   v  = qkv[:, 2C:3C]                     -> v_cache[phys(row)]       (:269 / :311)
// RoPE (causal_model.py:32-60): within each 128-wide head, complex pair i = (x[2i], x[2i+1]) is
// multiplied by exp(i * pos * theta_i) with pos = frame (i < 22), h (22 <= i < 43), w (i >= 43);
// rope_cs[pos][i] = (cos, sin) as float2, computed in fp64 on the host like the reference's table.
// The product is evaluated in fp32 (reference: fp64) and rounded to bf16.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ring_dst_row(const llb_step_params* sp, int row) {
  // physical cache row for new-token index `row`, or -1 if this token is not stored
  const int n = sp->n_write_segs;
#pragma unroll
  for (int i = 0; i < LLB_MAX_SEGS; ++i) {
    if (i < n) {
      const int src = sp->write_src[i], len = sp->write_n[i];
      if (row >= src && row < src + len) return sp->write_dst[i] + (row - src);
    }
  }
  return -1;
}

// Head-parallel destination of the 16-byte vector vi of a [*, C] row (head = vi / 16): owner rank and vector
// index inside that rank's [*, heads_per_rank * 128] buffer.
__device__ __forceinline__ void shard_slot(const llb_qkv_shard& sh, int vi, int vec_per_rank, int& r, int& lv) {
  if (sh.round_robin) {
    const int head = vi >> 4;
    r = head % sh.n_ranks;
    lv = (head / sh.n_ranks) * 16 + (vi & 15);
  } else {
    r = vi / vec_per_rank;
    lv = vi - r * vec_per_rank;
  }
}

// One warp per (token row, part) with part = blockIdx.y: 0 = q, 1 = k, 2 = v.  Splitting the row three ways
// keeps the per-thread state at one 3 KB vector set (70 registers, three CTAs per SM) instead of five, which
// is what this latency-bound pass needs: more rows in flight, not more loads per thread.
template <int NV>
__global__ void __launch_bounds__(kRowWarps * 32, 3)
rmsnorm_rope_append_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld_qkv,
                           __nv_bfloat16* __restrict__ q_out, int64_t ldq,
                           __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache,
                           int64_t ld_cache, int rows, int C,
                           const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                           float eps, const float2* __restrict__ rope_cs, int grid_h, int grid_w,
                           const llb_step_params* __restrict__ sp, const llb_qkv_shard sh) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's output is complete past here
  griddep_launch_dependents();
  const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int part = blockIdx.y;
  if (row >= rows) return;
  const int nvec = C / 8;
  const int grow = sh.row0 + row;  // row in the whole chunk (token-sharded callers pass their offset)
  const bool sharded = sh.n_ranks > 1;
  const int dst = (part != 0 && (k_cache != nullptr || sharded)) ? ring_dst_row(sp, grow) : -1;
  if (part != 0 && dst < 0) return;  // this token's K / V are not stored (sink re-cache rule)
  const int vec_per_rank = sh.heads_per_rank * 16;  // 16-byte vectors per rank's head slice
  const int64_t ldp = static_cast<int64_t>(vec_per_rank);  // peer leading dim in uint4 units
  const uint4* src = reinterpret_cast<const uint4*>(qkv + static_cast<int64_t>(row) * ld_qkv + part * C);

  if (part == 2) {
    // V: plain copy into the ring slot (or the owner rank's ring)
    uint4* vo = sharded ? nullptr : reinterpret_cast<uint4*>(v_cache + static_cast<int64_t>(dst) * ld_cache);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const uint4 v = src[vi];
        if (!sharded) {
          vo[vi] = v;
        } else {
          int r, lv;
          shard_slot(sh, vi, vec_per_rank, r, lv);
          reinterpret_cast<uint4*>(sh.v_peers[r])[static_cast<int64_t>(dst) * ldp + lv] = v;
        }
      }
    }
    return;
  }

  const __nv_bfloat16* wgt = part == 0 ? wq : wk;
  uint4 xv[NV], wv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + i * 32 < nvec) {
      xv[i] = src[lane + i * 32];
      wv[i] = __ldg(reinterpret_cast<const uint4*>(wgt) + lane + i * 32);
    }
  }
  const float rstd = rsqrtf(row_sumsq(xv, lane, nvec) / static_cast<float>(C) + eps);

  // token -> (frame, h, w), row-major over (frames, grid_h, grid_w)
  const int hw = grid_h * grid_w;
  const int f = grow / hw, rem = grow - f * hw;
  const int ph = rem / grid_w, pw = rem - ph * grid_w;
  const int pf = sp->rope_start_frame + f;
  uint4* orow = part == 0 ? reinterpret_cast<uint4*>(q_out + static_cast<int64_t>(row) * ldq)
                          : reinterpret_cast<uint4*>(k_cache + static_cast<int64_t>(dst) * ld_cache);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      // this vector covers channels [8*vi, 8*vi+8) = complex pairs 4*(vi%16) .. +3 of head vi/16
      const int pair0 = (vi & 15) * 4;
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xv[i]);
      const uint32_t* gw = reinterpret_cast<const uint32_t*>(&wv[i]);
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int pi = pair0 + e;
        const int pos = pi < 22 ? pf : (pi < 43 ? ph : pw);
        const float2 cs = __ldg(rope_cs + pos * 64 + pi);
        // RMSNorm: bf16(x * rstd) then * weight, rounded to bf16 (model.py:83)
        const float a = bf16_round(bf16_round(bf16_lo(xw[e]) * rstd) * bf16_lo(gw[e]));
        const float b = bf16_round(bf16_round(bf16_hi(xw[e]) * rstd) * bf16_hi(gw[e]));
        o[e] = pack_bf16x2(a * cs.x - b * cs.y, a * cs.y + b * cs.x);
      }
      const uint4 ov = make_uint4(o[0], o[1], o[2], o[3]);
      if (!sharded) {
        orow[vi] = ov;
      } else {
        // head exchange fused into the store: this 16-byte vector belongs to head vi/16, owned by
        // rank (vi/16)/heads_per_rank (or (vi/16) % n_ranks with the round-robin map); write it straight into
        // that rank's (peer-mapped) buffers
        int r, lv;
        shard_slot(sh, vi, vec_per_rank, r, lv);
        if (part == 0) reinterpret_cast<uint4*>(sh.q_peers[r])[static_cast<int64_t>(grow) * ldp + lv] = ov;
        else reinterpret_cast<uint4*>(sh.k_peers[r])[static_cast<int64_t>(dst) * ldp + lv] = ov;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// glue kernels
// ------------------------------------------------------------------------------------------------
// x [C_in, F, H, W] -> out [F*(H/2)*(W/2), C_in*4]; column = c*4 + ph*2 + pw  (Conv3d weight
// [1536, C_in, 1, 2, 2].flatten(1) order), row = f*(H/2)*(W/2) + (h/2)*(W/2) + (w/2).
__global__ void patchify_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                int c_in, int frames, int H, int W) {
  const int64_t total = static_cast<int64_t>(c_in) * frames * H * W;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int w = idx % W;
  const int h = (idx / W) % H;
  const int f = (idx / (static_cast<int64_t>(W) * H)) % frames;
  const int c = idx / (static_cast<int64_t>(W) * H * frames);
  const int64_t row = (static_cast<int64_t>(f) * (H / 2) + h / 2) * (W / 2) + w / 2;
  out[row * (c_in * 4) + c * 4 + (h & 1) * 2 + (w & 1)] = x[idx];
}

// y [F*(H/2)*(W/2), 4*C_out] with column = (ph*2 + pw)*C_out + c  ->  out [C_out, F, H, W]
// (einsum 'fhwpqrc->cfphqwr', causal_model.py:1259-1261)
__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out,
                                  int c_out, int frames, int H, int W) {
  const int64_t total = static_cast<int64_t>(c_out) * frames * H * W;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int w = idx % W;
  const int h = (idx / W) % H;
  const int f = (idx / (static_cast<int64_t>(W) * H)) % frames;
  const int c = idx / (static_cast<int64_t>(W) * H * frames);
  const int64_t row = (static_cast<int64_t>(f) * (H / 2) + h / 2) * (W / 2) + w / 2;
  out[idx] = y[row * (c_out * 4) + ((h & 1) * 2 + (w & 1)) * c_out + c];
}

// sinusoidal_embedding_1d (model.py:15-25): fp64, cos half first, result cast to bf16.
__global__ void sinusoidal_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int n,
                                  int dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * dim) return;
  const int r = idx / dim, c = idx % dim;
  const int half = dim / 2;
  const int i = c < half ? c : c - half;
  const double freq = pow(10000.0, -static_cast<double>(i) / static_cast<double>(half));
  const double a = static_cast<double>(t[r]) * freq;
  const double v = c < half ? cos(a) : sin(a);
  out[idx] = __float2bfloat16_rn(static_cast<float>(v));
}

// out[l, r, :] = bf16(table[l, r % table_rows, :] + e[r, :])
__global__ void modulation_table_kernel(const __nv_bfloat16* __restrict__ table,
                                        const __nv_bfloat16* __restrict__ e,
                                        __nv_bfloat16* __restrict__ out, int n_layers, int n_frames,
                                        int width) {
  const int64_t total = static_cast<int64_t>(n_layers) * n_frames * width;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % width;
  const int f = (idx / width) % n_frames;
  const int l = idx / (static_cast<int64_t>(width) * n_frames);
  out[idx] = __float2bfloat16_rn(__bfloat162float(table[static_cast<int64_t>(l) * width + c]) +
                                 __bfloat162float(e[static_cast<int64_t>(f) * width + c]));
}

__global__ void silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                            int64_t n) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  const float v = __bfloat162float(x[idx]);
  out[idx] = __float2bfloat16_rn(v / (1.0f + expf(-v)));
}

// Cross-rank barrier over peer-mapped flags (one CTA, thread j talks to rank j).
__global__ void peer_barrier_kernel(uint32_t* const* __restrict__ flags_peers, int rank, int n_ranks,
                                    uint32_t* __restrict__ epoch_local) {
  __shared__ uint32_t e_sh;
  if (threadIdx.x == 0) e_sh = *epoch_local + 1;
  __syncthreads();
  const uint32_t e = e_sh;
  const int j = threadIdx.x;
  if (j < n_ranks) {
    __threadfence_system();  // order this rank's earlier peer stores before the signal
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags_peers[j] + rank), "r"(e) : "memory");
    const uint32_t* mine = flags_peers[rank] + j;
    uint32_t v, spins = 0;
    uint64_t t0 = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((++spins & 0xfffu) == 0) {
        const uint64_t now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 5 * LLB_WAIT_TIMEOUT_NS) __trap();
      }
    } while (static_cast<int32_t>(v - e) < 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch_local = e;
}

}  // namespace llb

using namespace llb;

extern "C" int llb_peer_barrier(void* const* flags_peers_dev, int rank, int n_ranks, void* epoch_local,
                                void* stream) {
  LLB_CHECK_ARG(flags_peers_dev && epoch_local && n_ranks >= 1 && n_ranks <= LLB_MAX_RANKS && rank >= 0 &&
                    rank < n_ranks, "peer_barrier: bad arguments");
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint32_t* const*>(flags_peers_dev), rank, n_ranks, static_cast<uint32_t*>(epoch_local));
  LLB_LAUNCH_CHECK("peer_barrier_kernel");
  return LLB_OK;
}

extern "C" int llb_ln_modulate(const void* x, int64_t ldx, void* out, int64_t ldo, int rows, int C,
                               const void* shift, const void* scale, int64_t ld_mod,
                               int rows_per_frame, int row0, const void* ln_w, const void* ln_b,
                               float eps, void* stream) {
  LLB_CHECK_ARG(x && out && rows > 0, "ln_modulate: null tensor / no rows");
  LLB_CHECK_ARG(C % 8 == 0 && C <= 32 * kMaxVec * 8, "ln_modulate: C=%d unsupported", C);
  LLB_CHECK_ARG(ldx % 8 == 0 && ldo % 8 == 0 && ld_mod % 8 == 0, "ln_modulate: leading dims % 8");
  const bool affine = ln_w != nullptr;
  LLB_CHECK_ARG(affine ? (ln_b != nullptr) : (shift && scale && rows_per_frame > 0),
                "ln_modulate: need (ln_w, ln_b) or (shift, scale, rows_per_frame)");
  const int grid = (rows + kRowWarps - 1) / kRowWarps;
  auto kern = (C + 255) / 256 <= 6 ? ln_modulate_kernel<false, 6> : ln_modulate_kernel<false, kMaxVec>;
  LLB_CUDA(launch_ex(kern, dim3(grid), dim3(kRowWarps * 32), 0, static_cast<cudaStream_t>(stream), 1, true,
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, rows, C,
      static_cast<const __nv_bfloat16*>(shift), static_cast<const __nv_bfloat16*>(scale), ld_mod,
      rows_per_frame > 0 ? rows_per_frame : 1, row0, static_cast<const __nv_bfloat16*>(ln_w),
      static_cast<const __nv_bfloat16*>(ln_b), eps, nullptr, 0, nullptr));
  LLB_LAUNCH_CHECK("ln_modulate_kernel");
  return LLB_OK;
}

extern "C" int llb_ln_modulate_fp8(const void* x, int64_t ldx, void* out8, int64_t ld8, float* out_scale,
                                   int rows, int C, const void* shift, const void* scale, int64_t ld_mod,
                                   int rows_per_frame, int row0, const void* ln_w, const void* ln_b,
                                   float eps, void* stream) {
  LLB_CHECK_ARG(x && out8 && out_scale && rows > 0, "ln_modulate_fp8: null tensor / no rows");
  LLB_CHECK_ARG(C % 8 == 0 && C <= 32 * kMaxVec * 8, "ln_modulate_fp8: C=%d unsupported", C);
  LLB_CHECK_ARG(ldx % 8 == 0 && ld8 % 16 == 0 && ld_mod % 8 == 0, "ln_modulate_fp8: leading dims");
  const bool affine = ln_w != nullptr;
  LLB_CHECK_ARG(affine ? (ln_b != nullptr) : (shift && scale && rows_per_frame > 0),
                "ln_modulate_fp8: need (ln_w, ln_b) or (shift, scale, rows_per_frame)");
  const int grid = (rows + kRowWarps - 1) / kRowWarps;
  auto kern = (C + 255) / 256 <= 6 ? ln_modulate_kernel<true, 6> : ln_modulate_kernel<true, kMaxVec>;
  LLB_CUDA(launch_ex(kern, dim3(grid), dim3(kRowWarps * 32), 0, static_cast<cudaStream_t>(stream), 1, true,
      static_cast<const __nv_bfloat16*>(x), ldx, nullptr, 0, rows, C,
      static_cast<const __nv_bfloat16*>(shift), static_cast<const __nv_bfloat16*>(scale), ld_mod,
      rows_per_frame > 0 ? rows_per_frame : 1, row0, static_cast<const __nv_bfloat16*>(ln_w),
      static_cast<const __nv_bfloat16*>(ln_b), eps, static_cast<uint8_t*>(out8), ld8, out_scale));
  LLB_LAUNCH_CHECK("ln_modulate_kernel<fp8>");
  return LLB_OK;
}

extern "C" int llb_quant_rows_fp8(const void* x, int64_t ldx, void* out8, int64_t ld8, float* out_scale,
                                  int rows, int C, void* stream) {
  LLB_CHECK_ARG(x && out8 && out_scale && rows > 0 && C > 0 && C % 8 == 0, "quant_rows_fp8: bad arguments");
  LLB_CHECK_ARG(ldx % 8 == 0 && ld8 % 16 == 0, "quant_rows_fp8: leading dims");
  const int grid = (rows + kRowWarps - 1) / kRowWarps;
  LLB_CUDA(launch_ex(quant_rows_fp8_kernel, dim3(grid), dim3(kRowWarps * 32), 0, static_cast<cudaStream_t>(stream), 1, true,
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<uint8_t*>(out8), ld8, out_scale, rows, C));
  LLB_LAUNCH_CHECK("quant_rows_fp8_kernel");
  return LLB_OK;
}

extern "C" int llb_rmsnorm(const void* x, int64_t ldx, void* out, int64_t ldo, int rows, int C,
                           const void* w, float eps, void* stream) {
  LLB_CHECK_ARG(x && out && w && rows > 0, "rmsnorm: null tensor / no rows");
  LLB_CHECK_ARG(C % 8 == 0 && C <= kWideThreads * kWideVec * 8, "rmsnorm: C=%d unsupported", C);
  LLB_CHECK_ARG(ldx % 8 == 0 && ldo % 8 == 0, "rmsnorm: leading dims % 8");
  if (C > 32 * kMaxVec * 8) {
    LLB_CUDA(launch_ex(rmsnorm_wide_kernel, dim3(rows), dim3(kWideThreads), 0, static_cast<cudaStream_t>(stream), 1,
        true, static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, C,
        static_cast<const __nv_bfloat16*>(w), eps));
    LLB_LAUNCH_CHECK("rmsnorm_wide_kernel");
    return LLB_OK;
  }
  const int grid = (rows + kRowWarps - 1) / kRowWarps;
  LLB_CUDA(launch_ex(rmsnorm_kernel, dim3(grid), dim3(kRowWarps * 32), 0, static_cast<cudaStream_t>(stream), 1, true,
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, rows, C,
      static_cast<const __nv_bfloat16*>(w), eps));
  LLB_LAUNCH_CHECK("rmsnorm_kernel");
  return LLB_OK;
}

extern "C" int llb_rmsnorm_rope_append(const void* qkv, int64_t ld_qkv, void* q_out, int64_t ldq,
                                       void* k_cache, void* v_cache, int64_t ld_cache, int rows,
                                       int n_heads, const void* wq, const void* wk, float eps,
                                       const void* rope_cs, int grid_h, int grid_w,
                                       const llb_step_params* p_dev, const llb_qkv_shard* shard,
                                       void* stream) {
  LLB_CHECK_ARG(qkv && wq && wk && rope_cs && p_dev && rows > 0 && (q_out || shard),
                "rmsnorm_rope_append: null tensor / no rows");
  llb_qkv_shard sh;
  memset(&sh, 0, sizeof(sh));
  sh.n_ranks = 1;
  sh.heads_per_rank = n_heads;
  if (shard != nullptr) {
    sh = *shard;
    LLB_CHECK_ARG(sh.n_ranks >= 1 && sh.n_ranks <= LLB_MAX_RANKS &&
                      (sh.round_robin ? (sh.n_ranks <= n_heads && sh.heads_per_rank == (n_heads + sh.n_ranks - 1) / sh.n_ranks)
                                      : sh.heads_per_rank * sh.n_ranks == n_heads),
                  "rmsnorm_rope_append: bad shard description");
    for (int r = 0; r < sh.n_ranks && sh.n_ranks > 1; ++r)
      LLB_CHECK_ARG(sh.q_peers[r] && sh.k_peers[r] && sh.v_peers[r], "rmsnorm_rope_append: null peer pointer");
  }
  const int C = n_heads * 128;
  LLB_CHECK_ARG(C <= 32 * kMaxVec * 8, "rmsnorm_rope_append: n_heads=%d unsupported", n_heads);
  LLB_CHECK_ARG((k_cache == nullptr) == (v_cache == nullptr), "rmsnorm_rope_append: k/v cache mismatch");
  LLB_CHECK_ARG(ld_qkv % 8 == 0 && ldq % 8 == 0 && ld_cache % 8 == 0 && grid_h > 0 && grid_w > 0,
                "rmsnorm_rope_append: bad strides / grid");
  // the (cos, sin) table has LLB_ROPE_MAX_POS rows per axis (the reference's freqs[1024], causal_model.py:622-629);
  // the frame offset lives in device memory (p_dev->rope_start_frame), so the caller bounds start_frame + frames
  LLB_CHECK_ARG(grid_h <= LLB_ROPE_MAX_POS && grid_w <= LLB_ROPE_MAX_POS &&
                    (rows + grid_h * grid_w - 1) / (grid_h * grid_w) <= LLB_ROPE_MAX_POS,
                "rmsnorm_rope_append: grid %d x %d / %d rows exceed the %d-position RoPE table", grid_h, grid_w, rows,
                LLB_ROPE_MAX_POS);
  const int grid = (rows + kRowWarps - 1) / kRowWarps;
  auto kern = (C + 255) / 256 <= 6 ? rmsnorm_rope_append_kernel<6> : rmsnorm_rope_append_kernel<kMaxVec>;
  // grid.y: q / k / v parts; without a cache (norm + rope only) and unsharded there is nothing to do for v
  const int parts = (k_cache != nullptr || sh.n_ranks > 1) ? 3 : 1;
  LLB_CUDA(launch_ex(kern, dim3(grid, parts), dim3(kRowWarps * 32), 0, static_cast<cudaStream_t>(stream), 1, true,
      static_cast<const __nv_bfloat16*>(qkv), ld_qkv, static_cast<__nv_bfloat16*>(q_out), ldq,
      static_cast<__nv_bfloat16*>(k_cache), static_cast<__nv_bfloat16*>(v_cache), ld_cache, rows, C,
      static_cast<const __nv_bfloat16*>(wq), static_cast<const __nv_bfloat16*>(wk), eps,
      static_cast<const float2*>(rope_cs), grid_h, grid_w, p_dev, sh));
  LLB_LAUNCH_CHECK("rmsnorm_rope_append_kernel");
  return LLB_OK;
}

extern "C" int llb_patchify(const void* x, void* out, int c_in, int frames, int H, int W,
                            void* stream) {
  LLB_CHECK_ARG(x && out && c_in > 0 && frames > 0 && H % 2 == 0 && W % 2 == 0, "patchify: bad args");
  const int64_t total = static_cast<int64_t>(c_in) * frames * H * W;
  patchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), c_in, frames, H, W);
  LLB_LAUNCH_CHECK("patchify_kernel");
  return LLB_OK;
}

extern "C" int llb_unpatchify(const void* y, void* out, int c_out, int frames, int H, int W,
                              void* stream) {
  LLB_CHECK_ARG(y && out && c_out > 0 && frames > 0 && H % 2 == 0 && W % 2 == 0, "unpatchify: bad args");
  const int64_t total = static_cast<int64_t>(c_out) * frames * H * W;
  unpatchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(out), c_out, frames, H, W);
  LLB_LAUNCH_CHECK("unpatchify_kernel");
  return LLB_OK;
}

extern "C" int llb_sinusoidal(const float* t, void* out, int n, int dim, void* stream) {
  LLB_CHECK_ARG(t && out && n > 0 && dim > 0 && dim % 2 == 0, "sinusoidal: bad args");
  sinusoidal_kernel<<<(n * dim + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      t, static_cast<__nv_bfloat16*>(out), n, dim);
  LLB_LAUNCH_CHECK("sinusoidal_kernel");
  return LLB_OK;
}

extern "C" int llb_modulation_table(const void* modulation, const void* e0, void* out, int n_layers,
                                    int n_frames, int width, void* stream) {
  LLB_CHECK_ARG(modulation && e0 && out && n_layers > 0 && n_frames > 0 && width > 0,
                "modulation_table: bad args");
  const int64_t total = static_cast<int64_t>(n_layers) * n_frames * width;
  modulation_table_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                            static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(modulation), static_cast<const __nv_bfloat16*>(e0),
      static_cast<__nv_bfloat16*>(out), n_layers, n_frames, width);
  LLB_LAUNCH_CHECK("modulation_table_kernel");
  return LLB_OK;
}

extern "C" int llb_silu(const void* x, void* out, int64_t n, void* stream) {
  LLB_CHECK_ARG(x && out && n > 0, "silu: bad args");
  silu_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), n);
  LLB_LAUNCH_CHECK("silu_kernel");
  return LLB_OK;
}

// Element-wise / row kernels of the streaming VAE decoder (reference: wan/modules/vae.py), all on
// channels-last bf16 activations [frames, H, W, Cp] with Cp = channels padded to a multiple of 64 (pad
// channels are kept at exactly zero by every kernel).  These are HBM-bound; each pass reads and writes its
// tensor once with 16-byte accesses.  bf16 rounding points follow the reference's op sequence so that the
// results track the PyTorch bf16 path element for element.
#include "llb_common.cuh"
#include "llb_host.h"

namespace llb {

__device__ __forceinline__ float silu_ref(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// RMS_norm.forward (vae.py:51-54) = F.normalize(x, dim=channels) * sqrt(C) * gamma (+ 0), then optionally
// nn.SiLU: per pixel, LPP lanes x NV 16-byte vectors cover the Cp channels.  Each lane group handles kPix
// pixels per pass with all their loads issued before any arithmetic (the kernel is a pure HBM stream: 2 x
// tensor bytes, nothing to reuse), and walks the tensor with a grid stride.
constexpr int kNormPix = 4;
template <int LPP, int NV>
__global__ void __launch_bounds__(256)
vae_norm_kernel(const __nv_bfloat16* __restrict__ in, int in_frames, int in_t0, __nv_bfloat16* __restrict__ out,
                int out_frames, int out_t0, int T, int pixels, int Cp, float scale,
                const __nv_bfloat16* __restrict__ gamma, int silu) {
  griddep_wait();
  const int sub = threadIdx.x % LPP;
  const int nvec = Cp / 8;
  const int total = T * pixels;  // < 2^31 (checked on the host)
  const int groups = (gridDim.x * blockDim.x) / LPP;
  // gamma is the same for every pixel: keep this lane's slices in registers
  uint4 g4[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = sub + i * LPP;
    g4[i] = vi < nvec ? __ldg(reinterpret_cast<const uint4*>(gamma) + vi) : make_uint4(0, 0, 0, 0);
  }
  // uniform trip count: every lane of a warp must reach the shuffles below, out-of-range groups just idle
  const int n_iters = (total + groups * kNormPix - 1) / (groups * kNormPix);
  for (int it = 0; it < n_iters; ++it) {
    const int base = (blockIdx.x * blockDim.x + threadIdx.x) / LPP + it * groups * kNormPix;
    uint4 v[kNormPix][NV];
    const uint4* src[kNormPix];
    uint4* dst[kNormPix];
    bool live[kNormPix];
#pragma unroll
    for (int k = 0; k < kNormPix; ++k) {
      const int gp = base + k * groups;
      live[k] = gp < total;
      const int t = live[k] ? gp / pixels : 0;
      const int px = live[k] ? gp - t * pixels : 0;
      src[k] = reinterpret_cast<const uint4*>(in + (static_cast<int64_t>((in_t0 + t) % in_frames) * pixels + px) * Cp);
      dst[k] = reinterpret_cast<uint4*>(out + (static_cast<int64_t>((out_t0 + t) % out_frames) * pixels + px) * Cp);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = sub + i * LPP;
        v[k][i] = (live[k] && vi < nvec) ? src[k][vi] : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int k = 0; k < kNormPix; ++k) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[k][i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = bf16_lo(w[e]), b = bf16_hi(w[e]);
          ss += a * a + b * b;
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      // x.norm(2, dim) comes back as a bf16 tensor, clamp_min(1e-12), then the bf16 division
      // (one exact reciprocal per pixel; x * (1/d) and x / d can differ in the last fp32 bit, which survives
      // the following bf16 rounding for about one element in 2^15)
      const float inv = __frcp_rn(fmaxf(bf16_round(sqrtf(ss)), 1e-12f));
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = sub + i * LPP;
        if (!(live[k] && vi < nvec)) continue;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[k][i]);
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g4[i]);
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float y[2] = {bf16_lo(w[e]), bf16_hi(w[e])};
          const float gg[2] = {bf16_lo(gw[e]), bf16_hi(gw[e])};
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            float a = bf16_round(y[q] * inv);
            a = bf16_round(a * scale);
            a = bf16_round(a * gg[q]);
            if (silu) a = silu_ref(a);
            y[q] = a;
          }
          o[e] = pack_bf16x2(y[0], y[1]);
        }
        dst[k][vi] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// nn.Upsample(scale_factor=(2, 2), mode='nearest') per frame (vae.py:57-63, 76-83)
__global__ void vae_upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int T, int H, int W,
                                      int nvec) {
  griddep_wait();
  const int64_t total = static_cast<int64_t>(T) * 2 * H * 2 * W * nvec;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    int64_t r = i / nvec;
    const int w2 = static_cast<int>(r % (2 * W)); r /= 2 * W;
    const int h2 = static_cast<int>(r % (2 * H));
    const int t = static_cast<int>(r / (2 * H));
    out[i] = in[((static_cast<int64_t>(t) * H + (h2 >> 1)) * W + (w2 >> 1)) * nvec + v];
  }
}

// out[c][r] = in[r][c]   (V^T for the attention block's P V product)
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld_in,
                                      __nv_bfloat16* __restrict__ out, int64_t ld_out, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][33];
  griddep_wait();
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[static_cast<int64_t>(r) * ld_in + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[static_cast<int64_t>(c) * ld_out + r] = tile[threadIdx.x][i];
  }
}

// P = softmax(logits * scale) over the first cols_valid columns of each row, written as bf16 with the
// padding columns [cols_valid, cols_pad) set to zero.  One warp per row.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ logits, int64_t ld, __nv_bfloat16* __restrict__ out, int64_t ldo,
                    int rows, int cols_valid, int cols_pad, float scale) {
  griddep_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* lr = logits + static_cast<int64_t>(row) * ld;
  float mx = -INFINITY;
  for (int c = lane; c < cols_valid; c += 32) mx = fmaxf(mx, lr[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < cols_valid; c += 32) sum += expf((lr[c] - mx) * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  __nv_bfloat16* orow = out + static_cast<int64_t>(row) * ldo;
  for (int c = lane; c < cols_pad; c += 32)
    orow[c] = __float2bfloat16(c < cols_valid ? expf((lr[c] - mx) * scale) * inv : 0.f);
}

// WanVAE_.cached_decode prologue (vae.py:573-579): z / (1/std) + mean, then conv2 (1x1x1, zc -> zc), written
// channels-last with the channel dimension zero-padded to Cp.  z is the reference's [zc, T, h, w] layout.
__global__ void vae_latent_in_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ mean,
                                     const __nv_bfloat16* __restrict__ inv_std, const __nv_bfloat16* __restrict__ w,
                                     const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out,
                                     int out_frames, int out_t0, int T, int zc, int64_t hw, int Cp) {
  griddep_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(T) * hw) return;
  const int t = static_cast<int>(i / hw);
  const int64_t px = i - static_cast<int64_t>(t) * hw;
  float u[32];
  for (int c = 0; c < zc; ++c) {
    float a = __bfloat162float(z[(static_cast<int64_t>(c) * T + t) * hw + px]);
    a = bf16_round(a / __bfloat162float(inv_std[c]));
    u[c] = bf16_round(a + __bfloat162float(mean[c]));
  }
  __nv_bfloat16* o = out + (static_cast<int64_t>((out_t0 + t) % out_frames) * hw + px) * Cp;
  for (int n = 0; n < zc; ++n) {
    float acc = 0.f;
    for (int c = 0; c < zc; ++c) acc += u[c] * __bfloat162float(w[n * zc + c]);
    o[n] = __float2bfloat16(acc + __bfloat162float(b[n]));
  }
  for (int n = zc; n < Cp; ++n) o[n] = __float2bfloat16(0.f);
}

// WanVAEWrapper.decode_to_pixel epilogue (utils/wan_wrapper.py:112): .float().clamp_(-1, 1), channel-first
__global__ void vae_pixel_out_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int T,
                                     int64_t hw, int Cp) {
  griddep_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(T) * hw) return;
  const int t = static_cast<int>(i / hw);
  const int64_t px = i - static_cast<int64_t>(t) * hw;
  const __nv_bfloat16* s = in + i * Cp;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    out[(static_cast<int64_t>(t) * 3 + c) * hw + px] = fminf(1.0f, fmaxf(-1.0f, __bfloat162float(s[c])));
}

}  // namespace llb

using namespace llb;

extern "C" int llb_vae_norm(const void* in, int in_frames, int in_t0, void* out, int out_frames, int out_t0, int T,
                            int64_t pixels, int Cp, int C, const void* gamma, int silu, void* stream) {
  LLB_CHECK_ARG(in && out && gamma && T > 0 && pixels > 0 && in_frames >= T && out_frames >= T, "vae_norm: bad arguments");
  LLB_CHECK_ARG(Cp % 64 == 0 && Cp <= 512 && C > 0 && C <= Cp, "vae_norm: Cp=%d C=%d unsupported", Cp, C);
  LLB_CHECK_ARG(static_cast<int64_t>(T) * pixels < (1ll << 30), "vae_norm: too many pixels");
  const float scale = sqrtf(static_cast<float>(C));  // python float dim ** 0.5, used as an fp32 scalar
  const int nvec = Cp / 8;
  const int lpp = nvec <= 8 ? 8 : (nvec <= 16 ? 16 : 32);
  const int64_t groups = (static_cast<int64_t>(T) * pixels + kNormPix - 1) / kNormPix;
  const int64_t blocks = (groups * lpp + 255) / 256;
  const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
  const unsigned grid = static_cast<unsigned>(blocks < static_cast<int64_t>(sms) * 16 ? blocks : sms * 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const auto* i = static_cast<const __nv_bfloat16*>(in);
  auto* o = static_cast<__nv_bfloat16*>(out);
  const auto* g = static_cast<const __nv_bfloat16*>(gamma);
  const int px = static_cast<int>(pixels);
  if (lpp == 8) vae_norm_kernel<8, 1><<<grid, 256, 0, s>>>(i, in_frames, in_t0, o, out_frames, out_t0, T, px, Cp, scale, g, silu);
  else if (lpp == 16) vae_norm_kernel<16, 1><<<grid, 256, 0, s>>>(i, in_frames, in_t0, o, out_frames, out_t0, T, px, Cp, scale, g, silu);
  else if (nvec <= 32) vae_norm_kernel<32, 1><<<grid, 256, 0, s>>>(i, in_frames, in_t0, o, out_frames, out_t0, T, px, Cp, scale, g, silu);
  else vae_norm_kernel<32, 2><<<grid, 256, 0, s>>>(i, in_frames, in_t0, o, out_frames, out_t0, T, px, Cp, scale, g, silu);
  LLB_LAUNCH_CHECK("vae_norm_kernel");
  return LLB_OK;
}

extern "C" int llb_vae_upsample2x(const void* in, void* out, int T, int H, int W, int Cp, void* stream) {
  LLB_CHECK_ARG(in && out && T > 0 && H > 0 && W > 0 && Cp % 8 == 0, "vae_upsample2x: bad arguments");
  const int64_t total = static_cast<int64_t>(T) * 4 * H * W * (Cp / 8);
  const int64_t blocks = (total + 255) / 256;
  const unsigned grid = static_cast<unsigned>(blocks < 148 * 32 ? blocks : 148 * 32);
  vae_upsample2x_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(in), static_cast<uint4*>(out), T, H, W, Cp / 8);
  LLB_LAUNCH_CHECK("vae_upsample2x_kernel");
  return LLB_OK;
}

extern "C" int llb_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int rows, int cols,
                                  void* stream) {
  LLB_CHECK_ARG(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "transpose: bad arguments");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_bf16_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), ld_in, static_cast<__nv_bfloat16*>(out), ld_out, rows, cols);
  LLB_LAUNCH_CHECK("transpose_bf16_kernel");
  return LLB_OK;
}

extern "C" int llb_softmax_rows(const float* logits, int64_t ld, void* out, int64_t ldo, int rows, int cols_valid,
                                int cols_pad, float scale, void* stream) {
  LLB_CHECK_ARG(logits && out && rows > 0 && cols_valid > 0 && cols_pad >= cols_valid && ld >= cols_valid &&
                    ldo >= cols_pad, "softmax_rows: bad arguments");
  softmax_rows_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, ld, static_cast<__nv_bfloat16*>(out), ldo, rows, cols_valid, cols_pad, scale);
  LLB_LAUNCH_CHECK("softmax_rows_kernel");
  return LLB_OK;
}

extern "C" int llb_vae_latent_in(const void* z, const void* mean, const void* inv_std, const void* w, const void* b,
                                 void* out, int out_frames, int out_t0, int T, int zc, int64_t hw, int Cp,
                                 void* stream) {
  LLB_CHECK_ARG(z && mean && inv_std && w && b && out && T > 0 && hw > 0 && out_frames >= T, "vae_latent_in: bad arguments");
  LLB_CHECK_ARG(zc > 0 && zc <= 32 && Cp >= zc && Cp % 8 == 0, "vae_latent_in: zc=%d Cp=%d unsupported", zc, Cp);
  const int64_t n = static_cast<int64_t>(T) * hw;
  vae_latent_in_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(z), static_cast<const __nv_bfloat16*>(mean),
      static_cast<const __nv_bfloat16*>(inv_std), static_cast<const __nv_bfloat16*>(w),
      static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(out), out_frames, out_t0, T, zc, hw, Cp);
  LLB_LAUNCH_CHECK("vae_latent_in_kernel");
  return LLB_OK;
}

extern "C" int llb_vae_pixel_out(const void* in, float* out, int T, int64_t hw, int Cp, void* stream) {
  LLB_CHECK_ARG(in && out && T > 0 && hw > 0 && Cp >= 3, "vae_pixel_out: bad arguments");
  const int64_t n = static_cast<int64_t>(T) * hw;
  vae_pixel_out_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), out, T, hw, Cp);
  LLB_LAUNCH_CHECK("vae_pixel_out_kernel");
  return LLB_OK;
}

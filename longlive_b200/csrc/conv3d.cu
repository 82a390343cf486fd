// llb_conv3d - causal 3-D convolution on channels-last activations as an implicit GEMM on tcgen05.
//
//   out[t, h, w, :] = bias + sum over taps (dt, dh, dw) of  W[:, tap, :] . in[t + dt - (kt-1), h + dh - kh/2, w + dw - kw/2, :]
//
// This is the convolution of the reference's streaming VAE decoder (wan/modules/vae.py:17-36 CausalConv3d,
// and the per-frame nn.Conv2d of Resample, :66-100, as the kt = 1 case): zero padding in h / w, causal in t,
// where the frames before the current call are the two-frame history the callers keep (feat_cache, :202-220).
// Here the history is not a separate tensor: the input is a ring of frames, the new frames are written at
// ring positions in_t0 .. in_t0 + T - 1 by the producing kernel and the taps at t - 1, t - 2 simply read the
// positions before them (zero-initialised ring == zero padding at the start of a stream).
//
// Structure: the persistent tcgen05 GEMM of gemm_bf16.cu with a different A operand.  One CTA tile is a
// 16 (h) x 8 (w) pixel patch of one frame (M = 128 rows, row = h_local * 8 + w_local) times BN output channels;
// the K loop runs over (dt, dw, channel chunk, dh).  The A operand of the three vertical taps dh = 0, 1, 2 is
// ONE TMA box: the patch plus a one-pixel halo above and below, [18 h x 8 w pixels x CK channels], loaded once
// per (dt, dw, chunk).  With rows ordered h-major, eight consecutive rows (one 128-byte-swizzle atom) are the
// eight pixels of one image row, so the window of tap dh is the same shared-memory tile entered one atom
// further down: its UMMA descriptor is base + dh * atom.  Coordinates outside the image (also negative) are
// zero-filled by the TMA unit, which is the spatial zero padding.  The kernels are bound by the bytes an SM can
// take in (64 B/clk), so loading the activations once per three taps is what buys speed:
// 18.4 + 3 x 24 KB instead of 3 x (16 + 24) KB per (dt, dw, chunk) at BN = 192.
// Weights come from a tap-major [Cout, taps*Cin] matrix, one [BN x CK] slice per (tap, chunk), in a second
// shared-memory ring.  The MMA warp and the 8 epilogue warps are the GEMM's; the epilogue maps tile rows back to
// pixels, adds the residual (x + h of ResidualBlock, vae.py:220) after rounding conv + bias to bf16 as the
// reference does, and can emit the RMS_norm + SiLU of the result for the next convolution.
#include <stdlib.h>

#include "llb_common.cuh"
#include "llb_host.h"

namespace llb {

constexpr int kConvTH = 16, kConvTW = 8;  // pixel patch = 128 GEMM rows, h-major
constexpr int kConvEpiWarps = 8;
constexpr int kConvThreads = 64 + kConvEpiWarps * 32;
constexpr int kConvEpiStageBytesPerWarp = 32 * 64;

// BN: output channels per tile (64 / 96 / 128 / 192).  CK: channels per chunk: 64 (128-byte swizzle rows) or
// 32 (64-byte rows) - the latter for layers whose channel count is a multiple of 32 only (the 96-channel
// full-resolution stage), so that no zero padding is moved or multiplied.  CPS: chunks per pipeline stage -
// with 32-channel chunks one chunk is only two short MMAs per tap, too little work per barrier round trip, so
// a stage then carries all three chunks of the 96 channels.
// MT: patches per CTA tile, stacked vertically (MT * 16 image rows in ONE halo box): each weight slice then
// feeds MT MMAs, halving the weight bytes per flop where BN is small (two accumulators of BN <= 128 columns,
// double-buffered, still fit the 512 TMEM columns).
template <int BN, int CK, int CPS, int MT>
struct ConvCfg {
  static constexpr int kAStages = MT == 2 ? 2 : 3;
  static constexpr int kPatchH = kConvTH * MT;
  static constexpr int kAtomBytes = 8 * CK * 2;                       // 8 rows = one image row of the patch
  static constexpr int kChunkA = (kPatchH + 2) * 8 * CK * 2;          // patches + halo rows
  static constexpr int kChunkB = BN * CK * 2;
  static constexpr int kStageA = CPS * kChunkA;
  static constexpr int kStageB = CPS * kChunkB;
  static constexpr int kTmemCols = (2 * MT * BN <= 128) ? 128 : (2 * MT * BN <= 256 ? 256 : 512);
  static_assert(2 * MT * BN <= 512, "TMEM columns");
  static constexpr int kSlabs = BN / 32;                      // 32-column epilogue slabs, dealt out alternately
  static constexpr int kMySlabs = (kSlabs + 1) / 2;          // most slabs one epilogue warp handles
  static constexpr int kEpiVecBytesPerWarp = 2 * kMySlabs * 32 * 4;  // fp32 bias and norm gamma of its slabs
  static constexpr int kNormXchgBytes = 2 * 8 * 32 * 4;      // per-row sums of squares, double-buffered by tile parity
  static constexpr int kEpiBytes = kConvEpiWarps * (kConvEpiStageBytesPerWarp + kEpiVecBytesPerWarp) + kNormXchgBytes;
  static constexpr int kFixedBytes = 1024 + kEpiBytes + 256 + kAStages * kStageA;
  static constexpr int kBFit = (232448 - kFixedBytes) / kStageB;
  static constexpr int kBStages = kBFit > 10 ? 10 : kBFit;
  static constexpr int kSmemBytes = kFixedBytes + kBStages * kStageB;
  static_assert(kBStages >= 3 && kSmemBytes <= 232448, "shared memory budget");
  static_assert(2 * kAStages + 2 * kBStages + 5 <= 32, "barrier block is 256 bytes");
  static_assert(kStageA % 1024 == 0 && kChunkB % 512 == 0, "swizzle atom alignment");
};

struct ConvParams {
  int H, W, Cin, Cout, T;  // Cin: channels iterated per tap (multiple of CK); Cout: channels produced
  int ld_out;              // channel stride of out / res (>= Cout)
  int kt, kh, kw;
  int in_frames, in_t0;
  __nv_bfloat16* out;
  int out_frames, out_t0, out_t_step;
  const __nv_bfloat16* res;
  int res_frames, res_t0;
  const __nv_bfloat16* bias;
  int tiles_h, tiles_w, num_n_tiles;
  // fused RMS_norm (+SiLU) of the result into a second ring (kNorm kernels; requires one n-tile)
  __nv_bfloat16* norm_out;
  int norm_frames, norm_t0, norm_silu;
  float norm_scale;  // sqrt(real channel count)
  const __nv_bfloat16* norm_gamma;
};

// kNorm: the epilogue additionally writes RMS_norm(result) (* gamma, optionally SiLU) into a second ring -
// the input of the NEXT convolution - so the normalisation costs no pass over HBM of its own.  The two
// epilogue warps of a TMEM lane quadrant each see half of a pixel's channels; they exchange their partial
// sums of squares through shared memory and a 64-thread named barrier.
template <int BN, int CK, int CPS, bool kNorm, int MT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3d_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
              const ConvParams p) {
  using Cfg = ConvCfg<BN, CK, CPS, MT>;
  constexpr int kBStages = Cfg::kBStages;
  constexpr int kConvAStages = Cfg::kAStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;                                    // kConvAStages halo tiles
  const uint32_t b_base = a_base + kConvAStages * Cfg::kStageA;         // kBStages weight slices
  const uint32_t epi_base = b_base + kBStages * Cfg::kStageB;
  uint8_t* epi_gen = smem_gen + (epi_base - smem_base);
  const uint32_t bar_base = epi_base + Cfg::kEpiBytes;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (kConvAStages + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * kConvAStages + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (2 * kConvAStages + kBStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kConvAStages + 2 * kBStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kConvAStages + 2 * kBStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kConvAStages + 2 * kBStages + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(epi_gen + Cfg::kEpiBytes + 8 * (2 * kConvAStages + 2 * kBStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_frame = p.tiles_h * p.tiles_w;
  const int num_m_tiles = p.T * tiles_per_frame;
  const int num_tiles = num_m_tiles * p.num_n_tiles;
  const int cgroups = p.Cin / (CK * CPS);  // chunk groups per tap

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_in);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < kConvAStages; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int s = 0; s < kBStages; ++s) {
      mbar_init(bfull_bar(s), 1);
      mbar_init(bempty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kConvEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      // bytes of one halo box: (16 + kh - 1) image rows of 8 pixels
      const uint32_t a_bytes = static_cast<uint32_t>(CPS) * (Cfg::kPatchH + p.kh - 1) * 8 * CK * 2;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_idx = tile % num_m_tiles;
        const int n_idx = tile / num_m_tiles;
        const int t = m_idx / tiles_per_frame;
        const int rem = m_idx - t * tiles_per_frame;
        const int h0 = (rem / p.tiles_w) * Cfg::kPatchH;
        const int w0 = (rem % p.tiles_w) * kConvTW;
        for (int dt = 0; dt < p.kt; ++dt) {
          // causal: tap dt reads frame t + dt - (kt - 1) of the stream = ring slot before the new frames
          int tin = (p.in_t0 + t + dt - (p.kt - 1)) % p.in_frames;
          if (tin < 0) tin += p.in_frames;
          for (int dw = 0; dw < p.kw; ++dw) {
            for (int cg = 0; cg < cgroups; ++cg) {
              mbar_wait(aempty_bar(as), aph ^ 1);
              mbar_arrive_expect_tx(afull_bar(as), a_bytes);
#pragma unroll
              for (int j = 0; j < CPS; ++j)
                tma_load_4d(a_base + as * Cfg::kStageA + j * Cfg::kChunkA, &tmap_in, afull_bar(as),
                            (cg * CPS + j) * CK, w0 + dw - p.kw / 2, h0 - p.kh / 2, tin);
              if (++as == kConvAStages) { as = 0; aph ^= 1; }
              for (int dh = 0; dh < p.kh; ++dh) {
                const int tap = (dt * p.kh + dh) * p.kw + dw;
                mbar_wait(bempty_bar(bs), bph ^ 1);
                mbar_arrive_expect_tx(bfull_bar(bs), Cfg::kStageB);
#pragma unroll
                for (int j = 0; j < CPS; ++j)
                  tma_load_2d(b_base + bs * Cfg::kStageB + j * Cfg::kChunkB, &tmap_w, bfull_bar(bs),
                              tap * p.Cin + (cg * CPS + j) * CK, n_idx * BN);
                if (++bs == kBStages) { bs = 0; bph ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    int it = 0;
    const int groups = p.kt * p.kw * cgroups;  // halo tiles per output tile
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * (MT * BN);
      for (int g = 0; g < groups; ++g) {
        mbar_wait(afull_bar(as), aph);
        const uint32_t sa = a_base + as * Cfg::kStageA;
        for (int dh = 0; dh < p.kh; ++dh) {
          mbar_wait(bfull_bar(bs), bph);
          tc_fence_after();
          const uint32_t sb = b_base + bs * Cfg::kStageB;
          // the window of vertical tap dh starts dh image rows (= dh swizzle atoms) into the halo tile
          const uint64_t da = CK == 64 ? umma_desc_kmajor(sa + dh * Cfg::kAtomBytes)
                                       : umma_desc_kmajor_sw64(sa + dh * Cfg::kAtomBytes);
          const uint64_t db = CK == 64 ? umma_desc_kmajor(sb) : umma_desc_kmajor_sw64(sb);
          if (elect_one()) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)  // patch mt sits 16 image rows (16 atoms) further down the halo tile
#pragma unroll
              for (int j = 0; j < CPS; ++j)
#pragma unroll
                for (int k = 0; k < CK / 16; ++k)
                  umma_ss(d_tmem + mt * BN, da + ((mt * kConvTH * Cfg::kAtomBytes + j * Cfg::kChunkA) >> 4) + 2 * k,
                          db + j * (Cfg::kChunkB >> 4) + 2 * k, idesc, (g | dh | j | k) != 0);
            umma_commit(bempty_bar(bs));
            if (dh == p.kh - 1) {
              umma_commit(aempty_bar(as));
              if (g == groups - 1) umma_commit(tfull_bar(acc));
            }
          }
          __syncwarp();
          if (++bs == kBStages) { bs = 0; bph ^= 1; }
        }
        if (++as == kConvAStages) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    constexpr int kMySlabs = Cfg::kMySlabs;
    uint8_t* my_stage = epi_gen + (warp - 2) * kConvEpiStageBytesPerWarp;
    float* my_bias = reinterpret_cast<float*>(epi_gen + kConvEpiWarps * kConvEpiStageBytesPerWarp +
                                              (warp - 2) * Cfg::kEpiVecBytesPerWarp);
    float* my_gamma = my_bias + kMySlabs * 32;
    float* xchg = reinterpret_cast<float*>(epi_gen + kConvEpiWarps * (kConvEpiStageBytesPerWarp + Cfg::kEpiVecBytesPerWarp));
    const bool has_res = p.res != nullptr;
    const bool has_out = p.out != nullptr;
    const int seg = lane & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_idx = tile % num_m_tiles;
      const int n_idx = tile / num_m_tiles;
      const int t = m_idx / tiles_per_frame;
      const int rem = m_idx - t * tiles_per_frame;
      const int h0 = (rem / p.tiles_w) * Cfg::kPatchH;
      const int w0 = (rem % p.tiles_w) * kConvTW;
      const int t_out = (p.out_t0 + t * p.out_t_step) % p.out_frames;
      const int t_res = has_res ? (p.res_t0 + t) % p.res_frames : 0;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
#pragma unroll
      for (int i = 0; i < kMySlabs; ++i) {
        const int cg = n_idx * BN + (2 * i + h) * 32 + lane;
        const bool in = 2 * i + h < Cfg::kSlabs && cg < p.Cout;
        my_bias[i * 32 + lane] = (in && p.bias != nullptr) ? __bfloat162float(p.bias[cg]) : 0.f;
        if constexpr (kNorm) my_gamma[i * 32 + lane] = in ? __bfloat162float(p.norm_gamma[cg]) : 0.f;
      }
      __syncwarp();
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      // pixel of each of the four rows this lane handles in the transposed (coalesced) phase
      int64_t pix[4];
      bool row_ok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = q * 32 + i * 8 + (lane >> 2);  // row of the 16 x 8 patch: h-major
        const int ph = h0 + mt * kConvTH + (row >> 3), pw = w0 + (row & 7);
        row_ok[i] = ph < p.H && pw < p.W;
        pix[i] = static_cast<int64_t>(ph) * p.W + pw;
      }
      uint4 yfin[kNorm ? kMySlabs : 1][4];
      float ss[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * (MT * BN) + mt * BN + h * 32;
#pragma unroll
      for (int c = 0; c < kMySlabs; ++c) {
        if (2 * c + h >= Cfg::kSlabs) continue;  // warp-uniform
        const int col0 = n_idx * BN + (2 * c + h) * 32;
        uint32_t v[32];
        tmem_ld32(t_row + c * 64, v);
        tmem_wait_ld();
        if (2 * (c + 1) + h >= Cfg::kSlabs && mt == MT - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t packed[4];
          const float4 b0 = *reinterpret_cast<const float4*>(my_bias + c * 32 + g * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(my_bias + c * 32 + g * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = g * 8 + e * 2;
            packed[e] = pack_bf16x2(__uint_as_float(v[j]) + bb[2 * e], __uint_as_float(v[j + 1]) + bb[2 * e + 1]);
          }
          *reinterpret_cast<uint4*>(my_stage + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
              make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        __syncwarp();
        const int gcol = col0 + seg * 8;
        const bool col_ok = gcol < p.Cout;
        uint4 yv[4], xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2);
          yv[i] = *reinterpret_cast<const uint4*>(my_stage + r * 64 + ((seg ^ ((r >> 1) & 3)) << 4));
          xv[i] = make_uint4(0, 0, 0, 0);
          if (row_ok[i] && col_ok && has_res)
            xv[i] = *reinterpret_cast<const uint4*>(p.res + (static_cast<int64_t>(t_res) * p.H * p.W + pix[i]) * p.ld_out + gcol);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 y = yv[i];
          if (has_res) {
            const uint32_t* yy = reinterpret_cast<const uint32_t*>(&yv[i]);
            const uint32_t* xx = reinterpret_cast<const uint32_t*>(&xv[i]);
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = pack_bf16x2(bf16_lo(xx[e]) + bf16_lo(yy[e]), bf16_hi(xx[e]) + bf16_hi(yy[e]));
            y = make_uint4(o[0], o[1], o[2], o[3]);
          }
          const bool ok = row_ok[i] && col_ok;
          if (ok && has_out)
            *reinterpret_cast<uint4*>(p.out + (static_cast<int64_t>(t_out) * p.H * p.W + pix[i]) * p.ld_out + gcol) = y;
          if constexpr (kNorm) {
            if (!col_ok) y = make_uint4(0, 0, 0, 0);
            yfin[c][i] = y;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(&y);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = bf16_lo(w[e]), b = bf16_hi(w[e]);
              ss[i] += a * a + b * b;
            }
          }
        }
        __syncwarp();
      }
      if constexpr (kNorm) {
        // sum of squares of each pixel over ALL its channels: the 4 lanes sharing a row, then the partner warp
        const int par = (it * MT + mt) & 1;
        float* mine = xchg + (par * 8 + q * 2 + h) * 32;
        const float* theirs = xchg + (par * 8 + q * 2 + (h ^ 1)) * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], 1);
          ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], 2);
          if (seg == 0) mine[i * 8 + (lane >> 2)] = ss[i];
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        float inv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          inv[i] = __frcp_rn(fmaxf(bf16_round(sqrtf(ss[i] + theirs[i * 8 + (lane >> 2)])), 1e-12f));
        const int t_n = (p.norm_t0 + t) % p.norm_frames;
#pragma unroll
        for (int c = 0; c < kMySlabs; ++c) {
          if (2 * c + h >= Cfg::kSlabs) continue;
          const int gcol = n_idx * BN + (2 * c + h) * 32 + seg * 8;
          if (gcol >= p.Cout) continue;
          const float4 g0 = *reinterpret_cast<const float4*>(my_gamma + c * 32 + seg * 8);
          const float4 g1 = *reinterpret_cast<const float4*>(my_gamma + c * 32 + seg * 8 + 4);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (!row_ok[i]) continue;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(&yfin[c][i]);
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float y2[2] = {bf16_lo(w[e]), bf16_hi(w[e])};
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                float a = bf16_round(y2[k] * inv[i]);
                a = bf16_round(a * p.norm_scale);
                a = bf16_round(a * gg[2 * e + k]);
                if (p.norm_silu) a = __fdividef(a, 1.0f + __expf(-a));
                y2[k] = a;
              }
              o[e] = pack_bf16x2(y2[0], y2[1]);
            }
            *reinterpret_cast<uint4*>(p.norm_out + (static_cast<int64_t>(t_n) * p.H * p.W + pix[i]) * p.ld_out + gcol) =
                make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      }  // mt
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int CK, int CPS, bool kNorm, int MT>
static int launch_conv_n(const CUtensorMap& ti, const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, CK, CPS, MT>;
  LLB_SET_MAX_SMEM((conv3d_kernel<BN, CK, CPS, kNorm, MT>), Cfg::kSmemBytes);
  const int sms = device_sm_count();
  LLB_CHECK_ARG(sms > 0, "no CUDA device");
  const int tiles = p.T * p.tiles_h * p.tiles_w * p.num_n_tiles;
  const int grid = tiles < sms ? tiles : sms;
  conv3d_kernel<BN, CK, CPS, kNorm, MT><<<grid, kConvThreads, Cfg::kSmemBytes, stream>>>(ti, tw, p);
  LLB_LAUNCH_CHECK("conv3d_kernel");
  return LLB_OK;
}

template <int BN, int CK, int CPS>
static int launch_conv(const CUtensorMap& ti, const CUtensorMap& tw, const ConvParams& p, int mt, cudaStream_t stream) {
  if constexpr (BN <= 128) {
    if (mt == 2)
      return p.norm_out != nullptr ? launch_conv_n<BN, CK, CPS, true, 2>(ti, tw, p, stream)
                                   : launch_conv_n<BN, CK, CPS, false, 2>(ti, tw, p, stream);
  }
  return p.norm_out != nullptr ? launch_conv_n<BN, CK, CPS, true, 1>(ti, tw, p, stream)
                               : launch_conv_n<BN, CK, CPS, false, 1>(ti, tw, p, stream);
}

template <int CK, int CPS>
static int dispatch_conv(int bn, const CUtensorMap& ti, const CUtensorMap& tw, const ConvParams& p, int mt,
                         cudaStream_t s) {
  switch (bn) {
    case 192: return launch_conv<192, CK, CPS>(ti, tw, p, mt, s);
    case 128: return launch_conv<128, CK, CPS>(ti, tw, p, mt, s);
    case 96: return launch_conv<96, CK, CPS>(ti, tw, p, mt, s);
    default: return launch_conv<64, CK, CPS>(ti, tw, p, mt, s);
  }
}

}  // namespace llb

extern "C" int llb_conv3d(const llb_conv3d_desc* d, void* stream) {
  using namespace llb;
  LLB_CHECK_ARG(d && d->in && d->weight && (d->out || d->norm_out), "conv3d: null tensor");
  LLB_CHECK_ARG(d->H > 0 && d->W > 0 && d->T > 0 && d->in_frames > 0 && (d->out == nullptr || d->out_frames > 0),
                "conv3d: bad shape");
  LLB_CHECK_ARG(d->ld_in > 0 && d->ld_in % 8 == 0 && d->ld_out > 0 && d->ld_out % 8 == 0,
                "conv3d: channel strides must be multiples of 8 (ld_in=%d ld_out=%d)", d->ld_in, d->ld_out);
  LLB_CHECK_ARG(d->Cin > 0 && d->Cin % 32 == 0 && d->Cin <= d->ld_in && d->Cout > 0 && d->Cout % 32 == 0 &&
                    d->Cout <= d->ld_out,
                "conv3d: channel counts must be multiples of 32 within the strides (Cin=%d Cout=%d)", d->Cin, d->Cout);
  LLB_CHECK_ARG((d->kt == 1 || d->kt == 3) && (d->kh == 1 || d->kh == 3) && d->kw == d->kh,
                "conv3d: kernel %dx%dx%d unsupported", d->kt, d->kh, d->kw);
  LLB_CHECK_ARG(d->in_frames >= d->T + d->kt - 1, "conv3d: input ring of %d frames too short for T=%d, kt=%d",
                d->in_frames, d->T, d->kt);
  LLB_CHECK_ARG(d->out == nullptr || (d->out_t_step >= 1 && d->out_frames >= (d->T - 1) * d->out_t_step + 1),
                "conv3d: output ring too short");
  LLB_CHECK_ARG(d->res == nullptr || d->res_frames >= d->T, "conv3d: residual ring too short");
  LLB_CHECK_ARG(d->in != d->out, "conv3d: in-place convolution is not supported");

  const int ck = d->Cin % 64 == 0 ? 64 : 32;
  const int bn = d->Cout % 192 == 0 ? 192 : (d->Cout % 128 == 0 ? 128 : (d->Cout % 96 == 0 ? 96 : 64));
  ConvParams p;
  p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout; p.T = d->T;
  p.ld_out = d->ld_out;
  p.kt = d->kt; p.kh = d->kh; p.kw = d->kw;
  p.in_frames = d->in_frames; p.in_t0 = d->in_t0;
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.out_frames = d->out_frames > 0 ? d->out_frames : 1; p.out_t0 = d->out_t0; p.out_t_step = d->out_t_step;
  p.res = static_cast<const __nv_bfloat16*>(d->res);
  p.res_frames = d->res_frames > 0 ? d->res_frames : 1; p.res_t0 = d->res_t0;
  p.bias = static_cast<const __nv_bfloat16*>(d->bias);
  // two stacked patches per CTA when the accumulators fit (BN <= 128) and the image is tall enough to fill the SMs
  const char* mt_env = getenv("LLB_CONV_MT");
  int mt = (bn <= 128 && d->H >= 64) ? 2 : 1;
  if (mt_env != nullptr && (mt_env[0] == '1' || (mt_env[0] == '2' && bn <= 128))) mt = mt_env[0] - '0';
  p.tiles_h = (d->H + kConvTH * mt - 1) / (kConvTH * mt);
  p.tiles_w = (d->W + kConvTW - 1) / kConvTW;
  p.num_n_tiles = (d->Cout + bn - 1) / bn;
  p.norm_out = static_cast<__nv_bfloat16*>(d->norm_out);
  p.norm_frames = d->norm_frames > 0 ? d->norm_frames : 1;
  p.norm_t0 = d->norm_t0;
  p.norm_silu = d->norm_silu;
  p.norm_scale = sqrtf(static_cast<float>(d->norm_channels > 0 ? d->norm_channels : 1));
  p.norm_gamma = static_cast<const __nv_bfloat16*>(d->norm_gamma);
  if (d->norm_out != nullptr) {
    LLB_CHECK_ARG(p.num_n_tiles == 1, "conv3d: fused norm needs all %d output channels in one tile (<= 192)", d->Cout);
    LLB_CHECK_ARG(d->norm_gamma && d->norm_channels > 0 && d->norm_channels <= d->Cout && d->norm_frames >= d->T &&
                      d->norm_out != d->in, "conv3d: bad fused-norm arguments");
  }

  CUtensorMap ti, tw;
  // one box = the 16 x 8 patch plus its halo rows for the vertical taps
  int rc = make_tmap_4d_bf16(&ti, d->in, d->in_frames, d->H, d->W, d->ld_in, kConvTH * mt + d->kh - 1, kConvTW, ck, 2 * ck);
  if (rc) return rc;
  const int64_t kdim = static_cast<int64_t>(d->kt) * d->kh * d->kw * d->Cin;
  rc = make_tmap_2d_bf16_sw(&tw, d->weight, d->Cout, kdim, kdim, bn, ck, 2 * ck);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ck == 64) return dispatch_conv<64, 1>(bn, ti, tw, p, mt, s);
  return (d->Cin / 32) % 3 == 0 ? dispatch_conv<32, 3>(bn, ti, tw, p, mt, s) : dispatch_conv<32, 1>(bn, ti, tw, p, mt, s);
}

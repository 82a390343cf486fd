// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16, M = 128, K = 16) issued back to back by one thread,
// for the operand forms / N the attention kernel uses.  One CTA per SM, operands = whatever is in shared memory / TMEM
// (timing only).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I longlive_b200/csrc tools/micro/mma_rate.cu -o tools/micro/mma_rate
#include "llb_common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace llb;

constexpr int kTile = 128 * 128 * 2;
constexpr int kBox = 128 * 64 * 2;

// mode 0: SS N=128 (QK full)      1: SS N=64 (QK half)     2: TS N=128 MN-major B (PV)
// mode 3: old step  = [8 PV + 8 QK128] x 2 chains          4: new step = [8 QK64] + [8 PV + 8 QK64] x 2 chains
// mode 5: SS N=64 with the A operand alternating between two Q tiles every MMA
// mode 6: SS N=256 (two key tiles at once, for reference)
__global__ void mma_rate_kernel(int mode, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q0 = base, q1 = base + kTile, k0 = base + 2 * kTile, v0 = base + 3 * kTile, k1 = base + 4 * kTile;
  const uint32_t bar = base + 6 * kTile;
  const uint32_t slot = bar + 16;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  // some finite data
  for (int i = threadIdx.x; i < 6 * kTile / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  constexpr uint32_t id128 = umma_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t id64 = umma_idesc_bf16(128, 64, 0, 0);
  constexpr uint32_t id256 = umma_idesc_bf16(128, 256, 0, 0);
  constexpr uint32_t idpv = umma_idesc_bf16(128, 128, 0, 1);
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    uint32_t ph = 0;
    auto qk = [&](uint32_t d, uint32_t qa, uint32_t kb, uint32_t idesc) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t o = (kk >> 2) * kBox + (kk & 3) * 32;
        umma_ss(d, umma_desc_kmajor(qa + o), umma_desc_kmajor(kb + o), idesc, kk != 0);
      }
    };
    auto pv = [&](uint32_t d, uint32_t pa, uint32_t vb) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) umma_ts(d, pa + kk * 8, umma_desc_mnmajor(vb + kk * 2048, kBox), idpv, 1u);
    };
    for (int pass = 0; pass < 2; ++pass) {  // pass 0 = warm-up
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        for (int r = 0; r < reps; ++r) {
          switch (mode) {
            case 0: qk(tmem, q0, k0, id128); break;
            case 1: qk(tmem, q0, k0, id64); break;
            case 2: pv(tmem + 256, tmem, v0); break;
            case 3:
              pv(tmem + 256, tmem, v0); qk(tmem, q0, k0, id128);
              pv(tmem + 384, tmem + 128, v0); qk(tmem + 128, q1, k0, id128);
              break;
            case 4:
              qk(tmem + 64, q0, k0 + 8192, id64);
              pv(tmem + 256, tmem, v0); qk(tmem, q0, k0, id64);
              qk(tmem + 192, q1, k0 + 8192, id64);
              pv(tmem + 384, tmem + 128, v0); qk(tmem + 128, q1, k0, id64);
              break;
            case 5:
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const uint32_t o = (kk >> 2) * kBox + (kk & 3) * 32;
                umma_ss(tmem + (kk & 1) * 128, umma_desc_kmajor(((kk & 1) ? q1 : q0) + o), umma_desc_kmajor(k0 + o), id64, 1u);
              }
              break;
            case 6:
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const uint32_t o = (kk >> 2) * kBox + (kk & 3) * 32;
                // N = 256: rows 0..127 from k0, 128..255 from k1 (k1 = k0 + 2 tiles; SBO walks 8-row atoms, so this only
                // times the instruction - the B rows beyond 128 come from whatever follows k0's box)
                umma_ss(tmem, umma_desc_kmajor(q0 + o), umma_desc_kmajor(k0 + o), id256, kk != 0);
              }
              break;
          }
        }
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, ph);
      ph ^= 1;
      tc_fence_after();
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  (void)k1;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  const int smem = 6 * kTile + 2048;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const char* names[] = {"SS N=128 (QK)", "SS N=64 (QK half)", "TS N=128 MN-major B (PV)", "old step: 2 x [8 PV + 8 QK128]",
                         "new step: 2 x [8 QK64 | 8 PV + 8 QK64]", "SS N=64, A alternating", "SS N=256"};
  const int mmas[] = {8, 8, 8, 32, 48, 8, 8};
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 7; ++mode) {
      const int reps = 200;
      mma_rate_kernel<<<grid, 128, smem>>>(mode, reps, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < grid; ++i) avg += h[i];
      avg /= grid;
      printf("{\"grid\": %d, \"mode\": %d, \"what\": \"%s\", \"cycles_per_mma\": %.1f, \"cycles_per_rep\": %.1f}\n", grid, mode,
             names[mode], avg / (reps * mmas[mode]), avg / reps);
    }
  }
  return 0;
}

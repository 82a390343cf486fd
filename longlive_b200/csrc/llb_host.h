// Host-side helpers shared by the C-ABI translation units: error reporting, launch counting,
// and CUtensorMap encoding through the runtime's driver-entry-point lookup (no libcuda link).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/llb200.h"

namespace llb {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define LLB_CHECK_ARG(cond, ...)     \
  do {                               \
    if (!(cond)) {                   \
      llb::set_error(__VA_ARGS__);   \
      return LLB_E_INVALID;          \
    }                                \
  } while (0)

#define LLB_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      llb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                     __LINE__);                                                     \
      return LLB_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

// Called right after a kernel launch.
#define LLB_LAUNCH_CHECK(name)                                                         \
  do {                                                                                 \
    cudaError_t e_ = cudaGetLastError();                                               \
    if (e_ != cudaSuccess) {                                                           \
      llb::set_error("launch of %s failed: %s", name, cudaGetErrorString(e_));         \
      return LLB_E_CUDA;                                                               \
    }                                                                                  \
    llb::g_launches.fetch_add(1, std::memory_order_relaxed);                           \
  } while (0)

// 2-D bf16 row-major tensor [rows, cols] with leading dimension ld (elements), box
// [box_rows x box_cols] with 128-byte swizzle (box_cols * 2 bytes must be 128).
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld, uint32_t box_rows, uint32_t box_cols);

// Same for 1-byte elements (fp8): box_cols must be 128.
int make_tmap_2d_u8(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows, uint32_t box_cols);

// 4-D channels-last activation tensor [frames, H, W, C] (bf16): box [1, box_h, box_w, box_c] with
// 128-byte swizzle (box_c * 2 bytes must be 128); out-of-range coordinates (also negative) read as 0.
int make_tmap_4d_bf16(CUtensorMap* out, const void* base, uint64_t frames, uint64_t H, uint64_t W,
                      uint64_t C, uint32_t box_h, uint32_t box_w, uint32_t box_c, int swizzle_bytes);
// 2-D bf16 map with a 64- or 128-byte swizzle (box_cols * 2 bytes must equal swizzle_bytes).
int make_tmap_2d_bf16_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                         uint64_t ld, uint32_t box_rows, uint32_t box_cols, int swizzle_bytes);

// SM count of the CURRENT device (cached per device ordinal).
int device_sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: opt in once per (kernel
// instantiation, device ordinal), thread-safe.  Pass the kernel in parentheses if its template
// argument list contains commas.
#define LLB_SET_MAX_SMEM(kernel, bytes)                                                             \
  do {                                                                                              \
    static std::atomic<uint64_t> done_{0};                                                          \
    int dev_ = 0;                                                                                   \
    LLB_CUDA(cudaGetDevice(&dev_));                                                                 \
    const uint64_t bit_ = 1ull << (dev_ & 63);                                                      \
    if (!(done_.load(std::memory_order_acquire) & bit_)) {                                          \
      LLB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));   \
      done_.fetch_or(bit_, std::memory_order_release);                                              \
    }                                                                                               \
  } while (0)

// Programmatic dependent launch is opt-in: LLB_PDL=1 (read once).
bool pdl_enabled();

// Launch through cudaLaunchKernelEx with the optional attributes this library uses: a cluster width and
// programmatic stream serialization (the kernel must call griddep_wait() before it touches global
// memory that earlier kernels on the stream read or write).
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                      int cluster_x, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl && pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace llb

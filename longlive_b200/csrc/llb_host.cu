// Host-side support for libllb200.so: error strings, launch counter, tensor-map encoding and the
// integer KV-ring planner (the product implementation of the reference's cache index math).
#include "llb_host.h"

#include <stdlib.h>

#include <stdarg.h>
#include <string.h>

namespace llb {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e =
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d_bf16_sw(out, base, rows, cols, ld, box_rows, box_cols, 128);
}

int make_tmap_2d_bf16_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                         uint64_t ld, uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return LLB_E_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) {
    set_error("tensor map: base/stride must be 16-byte aligned (base=%p ld=%llu)", base,
              (unsigned long long)ld);
    return LLB_E_INVALID;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%llu cols=%llu ld=%llu box=%ux%u",
              (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld,
              box_rows, box_cols);
    return LLB_E_CUDA;
  }
  return LLB_OK;
}

int make_tmap_2d_u8(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return LLB_E_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ld % 16 != 0) {
    set_error("tensor map (u8): base/stride must be 16-byte aligned (base=%p ld=%llu)", base,
              (unsigned long long)ld);
    return LLB_E_INVALID;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(u8) failed (CUresult %d) rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    return LLB_E_CUDA;
  }
  return LLB_OK;
}

int make_tmap_4d_bf16(CUtensorMap* out, const void* base, uint64_t frames, uint64_t H, uint64_t W,
                      uint64_t C, uint32_t box_h, uint32_t box_w, uint32_t box_c, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return LLB_E_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (C * 2) % 16 != 0) {
    set_error("tensor map (4d): base/stride must be 16-byte aligned (base=%p C=%llu)", base, (unsigned long long)C);
    return LLB_E_INVALID;
  }
  cuuint64_t gdim[4] = {C, W, H, frames};
  cuuint64_t gstride[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (CUresult %d) frames=%llu H=%llu W=%llu C=%llu box=%ux%ux%u",
              (int)r, (unsigned long long)frames, (unsigned long long)H, (unsigned long long)W,
              (unsigned long long)C, box_h, box_w, box_c);
    return LLB_E_CUDA;
  }
  return LLB_OK;
}

bool pdl_enabled() {
  static const bool on = [] {
    // Off unless LLB_PDL=1: measured neutral on the 21-frame bench (93.9 vs 93.5 FPS, the forward runs
    // at the board power cap, so hiding launch gaps only lowers the clock), kept for latency-bound uses.
    const char* e = getenv("LLB_PDL");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

int device_sm_count() {
  static std::atomic<int> cache[64];  // zero-initialised; keyed by device ordinal
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  std::atomic<int>& slot = cache[dev & 63];
  int n = slot.load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    slot.store(n, std::memory_order_relaxed);
  }
  return n;
}

}  // namespace llb

extern "C" {

int llb_version(void) { return LLB_VERSION; }
const char* llb_last_error(void) { return llb::g_err; }
int64_t llb_launch_count(void) { return llb::g_launches.load(); }

// ------------------------------------------------------------------------------------------------
// KV ring planner.  Integer contract transcribed from the behaviour of
// wan/modules/causal_model.py:213-246 (roll branch), :291-306 (direct branch), :331-360 (attended
// window) and :849-905 (commit); see SURVEY.md section 8 "a-index".
// ------------------------------------------------------------------------------------------------
static inline int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }
static inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }

int64_t llb_kv_ring_phys(const llb_kv_config* cfg, int64_t rot, int64_t p) {
  const int64_t S = cfg->sink_tokens;
  const int64_t ring = cfg->cache_size - S;
  if (p < S || ring <= 0) return p;
  return S + ((p - S + rot) % ring);
}

// Append the physical image of logical range [lo, hi) as <= 2 contiguous physical ranges.
static int map_logical_range(const llb_kv_config* cfg, int64_t rot, int64_t lo, int64_t hi,
                             int64_t* starts, int64_t* lens, int cap) {
  int n = 0;
  const int64_t S = cfg->sink_tokens;
  const int64_t ring = cfg->cache_size - S;
  if (hi <= lo) return 0;
  // part below the sink boundary maps identically
  if (lo < S) {
    int64_t e = imin(hi, S);
    if (n < cap) { starts[n] = lo; lens[n] = e - lo; }
    n++;
    lo = e;
  }
  if (hi > lo) {
    if (ring <= 0) return -1;
    int64_t len = hi - lo;
    if (len > ring) return -1;
    int64_t p0 = S + ((lo - S + rot) % ring);
    int64_t first = imin(len, S + ring - p0);
    if (n < cap) { starts[n] = p0; lens[n] = first; }
    n++;
    if (len > first) {
      if (n < cap) { starts[n] = S; lens[n] = len - first; }
      n++;
    }
  }
  // merge adjacent ranges
  int m = 0;
  for (int i = 0; i < n && i < cap; ++i) {
    if (m > 0 && starts[m - 1] + lens[m - 1] == starts[i]) {
      lens[m - 1] += lens[i];
    } else {
      starts[m] = starts[i];
      lens[m] = lens[i];
      m++;
    }
  }
  return n > cap ? -1 : m;
}

int llb_kv_ring_plan(const llb_kv_config* cfg, const llb_kv_state* st, int64_t current_start,
                     int64_t num_new, int32_t sink_recache_after_switch, llb_kv_plan* plan) {
  LLB_CHECK_ARG(cfg && st && plan, "kv_ring_plan: null argument");
  LLB_CHECK_ARG(num_new > 0 && current_start >= 0, "kv_ring_plan: bad range");
  memset(plan, 0, sizeof(*plan));
  const int64_t size = cfg->cache_size, S = cfg->sink_tokens, M = cfg->max_attention_size;
  const int64_t G = st->global_end, Le = st->local_end, n = num_new;
  const int64_t current_end = current_start + n;
  const int is_recompute = (current_end <= G) && (current_start > 0);
  int64_t rot = st->rot;
  int64_t Le2, Ls2, ws;
  if (cfg->local_attn_size != -1 && current_end > G && n + Le > size) {
    const int64_t evicted = n + Le - size;
    const int64_t rolled = Le - evicted - S;
    LLB_CHECK_ARG(rolled >= 0 && size - S > 0, "kv_ring_plan: roll larger than the rolling region");
    Le2 = Le + (current_end - G) - evicted;
    Ls2 = Le2 - n;
    ws = is_recompute ? imax(Ls2, S) : Ls2;
    plan->action = 1;
    plan->num_evicted = evicted;
    plan->num_rolled = rolled;
    // the reference memmoves [S+evicted, S+evicted+rolled) down to [S, S+rolled); as a ring this
    // is a rotation of the region by `evicted`.
    rot = (rot + evicted) % (size - S);
  } else {
    Le2 = Le + (current_end - G);
    Ls2 = Le2 - n;
    ws = is_recompute ? imax(Ls2, S) : Ls2;
    if (sink_recache_after_switch) ws = Ls2;
    plan->action = 0;
  }
  LLB_CHECK_ARG(Le2 <= size && Le2 >= 0, "kv_ring_plan: local_end %lld outside cache of %lld",
                (long long)Le2, (long long)size);
  const int64_t off = imax(0, ws - Ls2);
  const int64_t wl = imax(0, Le2 - ws);
  LLB_CHECK_ARG(ws >= 0 || wl == 0, "kv_ring_plan: negative write start");
  plan->is_recompute = is_recompute;
  plan->current_end = current_end;
  plan->local_start = Ls2;
  plan->local_end = Le2;
  plan->write_start = ws;
  plan->write_end = Le2;
  plan->roped_offset = off;
  plan->write_len = wl;
  plan->rot_after = rot;
  // attended window
  if (S > 0) {
    plan->attn_sink_len = S;
    const int64_t budget = M - S;
    plan->attn_window_start = budget > 0 ? imax(S, Le2 - budget) : Le2;
  } else {
    plan->attn_sink_len = 0;
    plan->attn_window_start = imax(0, Le2 - M);
  }
  // physical write segments.  A write that covers the whole rolling region (KV-recache) makes
  // the old rotation irrelevant, so it is reset to keep the ring contiguous.
  if (wl > 0 && ws <= S && ws + wl >= size) rot = 0;
  plan->rot_after = rot;
  {
    int64_t s[LLB_MAX_SEGS], l[LLB_MAX_SEGS];
    int m = map_logical_range(cfg, rot, ws, ws + wl, s, l, LLB_MAX_SEGS);
    LLB_CHECK_ARG(m >= 0 && m <= LLB_MAX_SEGS, "kv_ring_plan: write range maps to %d segments", m);
    plan->n_write_segs = m;
    int64_t src = off;
    for (int i = 0; i < m; ++i) {
      plan->write_src[i] = src;
      plan->write_dst[i] = s[i];
      plan->write_n[i] = l[i];
      src += l[i];
    }
  }
  // physical attention segments: sink part then window part
  {
    int64_t s[LLB_MAX_SEGS + 2], l[LLB_MAX_SEGS + 2];
    int m = 0;
    // The reference reads temp[0:S] as the sink even when less than S tokens are valid (zeros).
    int m1 = map_logical_range(cfg, rot, 0, plan->attn_sink_len, s, l, 2);
    LLB_CHECK_ARG(m1 >= 0, "kv_ring_plan: bad sink range");
    m = m1;
    int m2 = map_logical_range(cfg, rot, plan->attn_window_start, Le2, s + m, l + m, 3);
    LLB_CHECK_ARG(m2 >= 0, "kv_ring_plan: bad window range");
    m += m2;
    // key order is irrelevant to softmax attention: sort by physical start, then merge
    for (int i = 1; i < m; ++i)
      for (int j = i; j > 0 && s[j] < s[j - 1]; --j) {
        int64_t t = s[j]; s[j] = s[j - 1]; s[j - 1] = t;
        t = l[j]; l[j] = l[j - 1]; l[j - 1] = t;
      }
    int k = 0;
    for (int i = 0; i < m; ++i) {
      if (l[i] <= 0) continue;
      if (k > 0 && s[k - 1] + l[k - 1] == s[i]) {
        l[k - 1] += l[i];
      } else {
        s[k] = s[i];
        l[k] = l[i];
        k++;
      }
    }
    LLB_CHECK_ARG(k <= LLB_MAX_SEGS, "kv_ring_plan: too many attention segments");
    plan->n_attn_segs = k;
    for (int i = 0; i < k; ++i) {
      plan->attn_start[i] = s[i];
      plan->attn_len[i] = l[i];
      plan->attn_total += l[i];
    }
  }
  return LLB_OK;
}

int llb_kv_ring_commit(const llb_kv_plan* plan, llb_kv_state* st) {
  LLB_CHECK_ARG(plan && st, "kv_ring_commit: null argument");
  st->rot = plan->rot_after;
  if (!plan->is_recompute) {
    st->global_end = plan->current_end;
    st->local_end = plan->local_end;
  }
  return LLB_OK;
}

}  // extern "C"

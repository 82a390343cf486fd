/*
 * llb200.h — C ABI of libllb200.so, the B200-native (sm_100a) kernel library behind the LongLive
 * frame-level autoregressive denoising hot path.
 *
 * The reference (kpham-augment/LongLive) has no FFI layer: the hot path is PyTorch library calls
 * inside wan/modules/causal_model.py.  Each entry point below replaces one group of those calls;
 * the reference call site it replaces is cited as file:line (relative to the reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers unless noted;
 *   - bf16 tensors are row-major with explicit leading dimensions given in ELEMENTS;
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*), never
 *     synchronises, is CUDA-graph capturable, and returns 0 on success or a negative LLB_E_*;
 *     llb_last_error() returns a thread-local description of the last failure;
 *   - there is no CPU fallback: on a machine without an sm_100 device every compute entry point
 *     fails with LLB_E_CUDA.
 */
#ifndef LLB200_H_
#define LLB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLB_VERSION 100

#define LLB_OK 0
#define LLB_E_INVALID (-1) /* bad argument (shape / alignment / null pointer)          */
#define LLB_E_CUDA (-2)    /* a CUDA runtime / driver call failed                      */
#define LLB_E_UNSUPPORTED (-3)

int llb_version(void);
const char* llb_last_error(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t llb_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * KV ring-buffer index math (host, integer only).
 * Replaces the cache bookkeeping of CausalWanSelfAttention.forward
 * (wan/modules/causal_model.py:206-246, 291-306, 331-360) and CausalWanModel._apply_cache_updates
 * (:849-905).  All quantities are in tokens.  The reference keeps the window in chronological
 * order by memmove-ing it ("roll"); here the rolling region is a ring: logical position p >= sink
 * lives at physical row  sink + (p - sink + rot) % (size - sink), so eviction is `rot += evicted`.
 * ------------------------------------------------------------------------------------------ */
typedef struct llb_kv_state {
  int64_t global_end; /* kv_cache["global_end_index"]                      */
  int64_t local_end;  /* kv_cache["local_end_index"]                       */
  int64_t rot;        /* ring rotation of the rolling region, in tokens    */
} llb_kv_state;

typedef struct llb_kv_config {
  int64_t cache_size;         /* kv_cache["k"].shape[1]                              */
  int64_t sink_tokens;        /* sink_size * frame_seqlen                            */
  int64_t max_attention_size; /* CausalWanSelfAttention.max_attention_size           */
  int32_t local_attn_size;    /* -1 = global attention (never rolls)                 */
} llb_kv_config;

#define LLB_MAX_SEGS 4
typedef struct llb_kv_plan {
  /* reference-observable values (bit-exact contract) */
  int32_t action; /* 0 = direct_insert, 1 = roll_and_insert */
  int32_t is_recompute;
  int64_t current_end;
  int64_t num_evicted, num_rolled;
  int64_t local_start, local_end; /* Ls', Le' */
  int64_t write_start, write_end; /* ws, Le' (logical) */
  int64_t roped_offset, write_len;
  int64_t attn_sink_len;     /* logical [0, attn_sink_len) attended                 */
  int64_t attn_window_start; /* logical [attn_window_start, local_end) attended     */
  /* physical plan for the kernels */
  int64_t rot_after; /* ring rotation to use for this call's writes / reads */
  int32_t n_write_segs;
  int64_t write_src[LLB_MAX_SEGS], write_dst[LLB_MAX_SEGS], write_n[LLB_MAX_SEGS]; /* new[src..] -> physical rows [dst..] */
  int32_t n_attn_segs;
  int64_t attn_start[LLB_MAX_SEGS], attn_len[LLB_MAX_SEGS]; /* physical key-row ranges */
  int64_t attn_total;
} llb_kv_plan;

/* Computes the plan for one model forward with `num_new` tokens starting at `current_start`.
 * Does not modify *st. */
int llb_kv_ring_plan(const llb_kv_config* cfg, const llb_kv_state* st, int64_t current_start,
                     int64_t num_new, int32_t sink_recache_after_switch, llb_kv_plan* plan);
/* Commits the plan after all layers ran (the reference's _apply_cache_updates index update). */
int llb_kv_ring_commit(const llb_kv_plan* plan, llb_kv_state* st);
/* physical row of logical position p */
int64_t llb_kv_ring_phys(const llb_kv_config* cfg, int64_t rot, int64_t p);

/* Per-forward dynamic parameters, resident in DEVICE memory so that one captured CUDA graph
 * serves every chunk: the host rewrites this struct (one small H2D copy) before each replay. */
typedef struct llb_step_params {
  int32_t rope_start_frame; /* current_start // frame_seqlen                                  */
  int32_t n_write_segs;
  int32_t write_src[LLB_MAX_SEGS], write_dst[LLB_MAX_SEGS], write_n[LLB_MAX_SEGS];
  int32_t n_attn_segs;
  int32_t attn_start[LLB_MAX_SEGS], attn_len[LLB_MAX_SEGS];
  int32_t reserved[1];
} llb_step_params;

/* ------------------------------------------------------------------------------------------
 * Optional single-stream head parallelism (Ulysses-style; the reference's only
 * sequence-parallel design is wan/distributed/xdit_context_parallel.py:131-192, for the non-causal
 * model).  Tokens are sharded over ranks for every GEMM / row kernel, heads are sharded for
 * attention and for the KV ring.  The two exchanges per block are FUSED into the producing kernels
 * as direct stores into peer GPUs' (NVLink-mapped, symmetric) buffers:
 *   llb_rmsnorm_rope_append  writes head h of Q / K / V to rank h / heads_per_rank,
 *   llb_attn_fwd             writes output rows to the rank that owns those token rows,
 * followed by llb_peer_barrier.  A null shard pointer means single-GPU operation.
 * ------------------------------------------------------------------------------------------ */
#define LLB_MAX_RANKS 8
typedef struct llb_qkv_shard {
  int32_t n_ranks;
  int32_t heads_per_rank;       /* width of every rank's buffers in heads (the most heads any rank owns) */
  int32_t row0;                 /* global token row of local row 0                           */
  int32_t round_robin;          /* 0: head h -> rank h / heads_per_rank (needs n_heads % n_ranks == 0);
                                 * 1: head h -> rank h % n_ranks, local head h / n_ranks (any n_ranks <= n_heads,
                                 *    e.g. 12 heads over 8 ranks: ranks 0-3 own two heads, ranks 4-7 one)   */
  void* q_peers[LLB_MAX_RANKS]; /* rank r: Q buffer [L_total, heads_per_rank*128]              */
  void* k_peers[LLB_MAX_RANKS]; /* rank r: K ring   [cache_rows, heads_per_rank*128]           */
  void* v_peers[LLB_MAX_RANKS];
} llb_qkv_shard;

typedef struct llb_out_shard {
  int32_t n_ranks;
  int32_t rows_per_rank;          /* token rows owned by each rank                             */
  int32_t head_col0;              /* first output column of this rank's heads (head0 * 128)    */
  int32_t head_col_stride;        /* columns between consecutive local heads; 0 = 128 (contiguous heads),
                                   * n_ranks * 128 for the round-robin head map                  */
  int64_t ld_out;                 /* leading dimension of the peers' output buffers            */
  void* out_peers[LLB_MAX_RANKS]; /* rank r: attention output [rows_per_rank, ld_out]          */
} llb_out_shard;

/* Device-side barrier across ranks: flags_peers (DEVICE array of n_ranks pointers) -> each rank's
 * uint32[n_ranks] flag array in peer-mapped memory; epoch_local: this rank's uint32 counter.  Makes
 * all earlier peer stores of every rank visible before any later kernel of any rank runs. */
int llb_peer_barrier(void* const* flags_peers_dev, int rank, int n_ranks, void* epoch_local, void* stream);

/* ------------------------------------------------------------------------------------------
 * GEMM family: out[M,N] = epilogue(A[M,K] @ W[N,K]^T + bias[N]); bf16 in/out, fp32 accumulate
 * in TMEM (tcgen05.mma, TMA-fed).  Replaces nn.Linear at causal_model.py:90-93,122-126,364,
 * 406-408,492,601-608 and model.py:172-178,193 together with the elementwise ops that follow.
 * ------------------------------------------------------------------------------------------ */
#define LLB_EPI_BIAS 0          /* y = acc + b                                               */
#define LLB_EPI_BIAS_GELU 1     /* gelu_tanh(y)            (ffn.0 + GELU, causal_model.py:407) */
#define LLB_EPI_BIAS_SILU 2     /* silu(y)                 (time_embedding, :606)            */
#define LLB_EPI_BIAS_GATE_RES 3 /* res + y * gate[row / rows_per_gate]   (:456, :467-468)    */
#define LLB_EPI_BIAS_RES 4      /* res + y                 (cross-attn residual, :460)       */
#define LLB_EPI_BIAS_F32 5      /* y written as float32 (out is float*, ldo in floats): attention logits of the VAE decoder */
#define LLB_EPI_BIAS_MUL 6      /* res * y: the gated FFN of the umT5 text encoder, fc1(x) * gelu(gate(x)) with
                                   res = gelu(gate(x)) from a BIAS_GELU_BF16 launch (wan/modules/t5.py:133) */
#define LLB_EPI_GEGLU_BF16 8    /* gated FFN of the umT5 encoder in one launch: W = 256-row tiles of [128 gate rows | the 128 fc1
                                   rows of the same outputs], out [M, N / 2] = bf16(fc1) * gelu_bf16(bf16(gate)) (t5.py:133) */
#define LLB_EPI_BIAS_GELU_BF16 7 /* the umT5 GELU module (wan/modules/t5.py:46-50): the tanh formula evaluated op by
                                   op with a bf16 rounding after each, as the reference's tensor expression does */

int llb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                  int M, int N, int K, int epilogue, const void* bias, const void* gate,
                  int64_t ld_gate, int rows_per_gate, int gate_row0, const void* res, int64_t ld_res,
                  void* stream);

/* Split-K form for skinny problems (the umT5 encoder's N = 4096 projections at 128-256 token rows stream 33-84 MB of
 * weights through only 64 CTAs otherwise): k_splits partial products are written as fp32 to `workspace`
 * (k_splits * M * N floats, caller-owned) and a second launch sums them and applies the epilogue
 * (LLB_EPI_BIAS or LLB_EPI_BIAS_RES; `res` may alias `out`).  With norm_out != NULL (N <= 8192) the second launch also
 * writes norm_out = bf16(bf16(out * rsqrt(mean(out^2) + norm_eps)) * norm_w): the T5LayerNorm that follows the projection
 * (wan/modules/t5.py:57-62, 166-167). */
int llb_gemm_bf16_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int M,
                         int N, int K, int epilogue, const void* bias, const void* res, int64_t ld_res,
                         int k_splits, void* workspace, int64_t workspace_bytes, const void* norm_w,
                         void* norm_out, int64_t ld_norm, float norm_eps, void* stream);

/* Optional FP8 linears (README.md:50 of the reference advertises "FP8 quantization" at 24.8 FPS but
 * ships no code for it, reports.md:24,39).  W8A8 with e4m3 operands on tcgen05 kind::f8f6f4:
 *   A8 [M,K] e4m3 with per-row scale a_scale[M] (dynamic, produced by llb_ln_modulate_fp8 /
 *   llb_quant_rows_fp8), W8 [N,K] e4m3 with per-output-channel scale w_scale[N] (static);
 *   out = epilogue(acc * a_scale[row] * w_scale[col] + bias), same epilogues as llb_gemm_bf16.
 * Leading dimensions of the fp8 operands are in elements (= bytes). */
int llb_gemm_fp8(const void* A8, int64_t lda, const float* a_scale, const void* W8, int64_t ldw,
                 const float* w_scale, void* out, int64_t ldo, int M, int N, int K, int epilogue,
                 const void* bias, const void* gate, int64_t ld_gate, int rows_per_gate, int gate_row0,
                 const void* res, int64_t ld_res, void* stream);
/* llb_ln_modulate whose result is quantised row-wise to e4m3 (scale = amax / 448) instead of stored as bf16 */
int llb_ln_modulate_fp8(const void* x, int64_t ldx, void* out8, int64_t ld8, float* out_scale, int rows,
                        int C, const void* shift, const void* scale, int64_t ld_mod, int rows_per_frame,
                        int row0, const void* ln_w, const void* ln_b, float eps, void* stream);
/* row-wise dynamic e4m3 quantisation of a bf16 matrix [rows, C] */
int llb_quant_rows_fp8(const void* x, int64_t ldx, void* out8, int64_t ld8, float* out_scale, int rows,
                       int C, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense attention  out = softmax(Q K^T * scale) V  per head (head_dim 128), flash-style with S/P/O
 * in TMEM.  Replaces attention()/flash_attention() (wan/modules/attention.py:43-197) at
 * causal_model.py:349-360 (self-attention over sink ++ window, read IN PLACE from the ring) and
 * model.py:189 (cross-attention over the cached text K/V).
 *   q   [Lq, n_heads*128] bf16, k/v [kv_rows, n_heads*128] bf16, out [Lq, n_heads*128] bf16.
 *   The attended keys are the union of physical row ranges listed in seg_dev
 *   (llb_step_params.n_attn_segs / attn_start / attn_len, device memory).
 *   out must be 16-byte aligned with ldo a multiple of 8 (single-GPU launches store it with TMA: rows >= Lq and
 *   columns outside [0, n_heads*128) of a wider buffer are never touched).
 *   variant: 0 = default (single-CTA kernel); 64 = the CTA-pair cta_group::2 kernel (same results, measured slower,
 *   DESIGN.md 4.1).  The other variants of rounds 1-2 were measured and removed (profiles/r02_attn_*.md).
 * ------------------------------------------------------------------------------------------ */
int llb_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                 int64_t ldv, void* out, int64_t ldo, int Lq, int n_heads, int kv_rows,
                 const llb_step_params* seg_dev, float scale, int variant, void* workspace,
                 int64_t workspace_bytes, const llb_out_shard* shard, void* stream);
/* Size of the device workspace llb_attn_fwd needs (partial O / (m,l) / flags of the stream-K
 * split, one slice per SM).  Allocate once per device, zero it once, reuse for every launch on
 * streams that are ordered with respect to each other. */
int64_t llb_attn_workspace_bytes(void);

/* ------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound).
 * ------------------------------------------------------------------------------------------ */
/* LayerNorm(eps, no affine) then x*(1+scale)+shift per frame  (causal_model.py:445, 463-464, 507)
 * or, with ln_w/ln_b non-null and shift/scale null, affine LayerNorm (norm3, :460).
 * shift/scale: [n_frames, ld_mod] rows selected by (row0 + row) / rows_per_frame. */
int llb_ln_modulate(const void* x, int64_t ldx, void* out, int64_t ldo, int rows, int C,
                    const void* shift, const void* scale, int64_t ld_mod, int rows_per_frame,
                    int row0, const void* ln_w, const void* ln_b, float eps, void* stream);

/* Fused WanRMSNorm(q), WanRMSNorm(k) (model.py:78-86), causal_rope_apply (causal_model.py:32-60)
 * and the KV-cache insert (:268-269 / :310-311) in one pass over the fused QKV GEMM output.
 *   qkv [rows, ld_qkv] = q | k | v column blocks of width C = n_heads*128.
 *   q_out [rows, ldq] roped queries; k/v written to the ring rows given by p_dev->write_*.
 *   rope_cs: float2 [LLB_ROPE_MAX_POS][64] (cos, sin) table; token -> (frame, h, w) row-major over
 *   (frames, grid_h, grid_w).  k_cache/v_cache may be null (norm + rope only).
 *   The caller guarantees p_dev->rope_start_frame + frames <= LLB_ROPE_MAX_POS (the reference fails
 *   there too: freqs[0][start:start+f] comes back short, causal_model.py:46-52). */
#define LLB_ROPE_MAX_POS 1024
int llb_rmsnorm_rope_append(const void* qkv, int64_t ld_qkv, void* q_out, int64_t ldq,
                            void* k_cache, void* v_cache, int64_t ld_cache, int rows, int n_heads,
                            const void* wq, const void* wk, float eps, const void* rope_cs,
                            int grid_h, int grid_w, const llb_step_params* p_dev,
                            const llb_qkv_shard* shard, void* stream);

/* Plain WanRMSNorm over C <= 8192 channels: out = bf16(x * rsqrt(mean(x^2)+eps)) * w  (model.py:78-86; also
 * T5LayerNorm, t5.py:57-62). */
int llb_rmsnorm(const void* x, int64_t ldx, void* out, int64_t ldo, int rows, int C,
                const void* w, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small glue kernels of CausalWanModel._forward_inference.
 * ------------------------------------------------------------------------------------------ */
/* patch_embedding Conv3d(k=s=(1,2,2)) as a gather: x [C_in, F, H, W] -> A [F*(H/2)*(W/2), C_in*4]
 * (causal_model.py:599-600, 959-963). */
int llb_patchify(const void* x, void* out, int c_in, int frames, int H, int W, void* stream);
/* unpatchify (causal_model.py:1240-1263): y [F*(H/2)*(W/2), 4*C_out] -> out [C_out, F, H, W]. */
int llb_unpatchify(const void* y, void* out, int c_out, int frames, int H, int W, void* stream);
/* sinusoidal_embedding_1d in fp64 -> bf16 (model.py:15-25): t [n] f32 -> out [n, dim] bf16 */
int llb_sinusoidal(const float* t, void* out, int n, int dim, void* stream);
/* out[l, r, :] = bf16(table[l, r % table_rows, :] + e[r, :]) — the per-layer adaLN table
 * e = modulation + e0 (causal_model.py:440, 506), all layers in one launch. */
int llb_modulation_table(const void* modulation, const void* e0, void* out, int n_layers,
                         int n_frames, int width, void* stream);
int llb_silu(const void* x, void* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Streaming VAE decoder kernels (reference wan/modules/vae.py, called from
 * WanVAEWrapper.decode_to_pixel, utils/wan_wrapper.py:96-117).  Activations are channels-last bf16
 * [frames, H, W, Cp], Cp = channel count padded to a multiple of 64 (pad channels stay exactly 0).
 * A tensor that feeds a causal convolution is a RING of frames: the new frames of a call sit at ring
 * positions t0 .. t0+T-1 (mod frames) and the two positions before them hold the previous call's last
 * two frames - the reference's per-conv feat_cache (vae.py:202-220) without a copy.  A zeroed ring is
 * the zero padding at the start of a stream.
 * ------------------------------------------------------------------------------------------ */
typedef struct llb_conv3d_desc {
  const void* in;       /* [in_frames, H, W, ld_in] bf16 ring */
  int in_frames, in_t0; /* ring length; ring index of the first new frame */
  int H, W;
  int Cin, ld_in;       /* channels convolved (multiple of 32) and the channel stride of `in` (>= Cin) */
  int Cout, ld_out;     /* channels produced (multiple of 32) and the channel stride of `out` / `res`;
                           columns [Cout, ld_out) are left untouched */
  const void* weight;   /* [Cout, kt*kh*kw*Cin] bf16: tap-major, k = ((dt*kh + dh)*kw + dw)*Cin + c */
  const void* bias;     /* [Cout] bf16 or NULL */
  int kt, kh, kw;       /* 3x3x3, 3x1x1, 1x3x3 or 1x1x1; causal in t, zero "same" padding in h / w */
  void* out;            /* [out_frames, H, W, ld_out] bf16; frame t goes to ring index (out_t0 + t*out_t_step) % out_frames */
  int out_frames, out_t0, out_t_step;
  const void* res;      /* optional [res_frames, H, W, ld_out]: out = bf16(res + bf16(conv + bias)); may alias out */
  int res_frames, res_t0;
  int T;                /* frames to produce */
  /* Optional fused RMS_norm (+SiLU) of the result (vae.py:51-54 applied to what `out` holds) into a second
   * ring - the input of the next convolution.  Needs Cout <= 192 (one tile sees a pixel's channels);
   * `out` may then be NULL if nobody reads the un-normalised result. */
  void* norm_out;       /* [norm_frames, H, W, ld_out] or NULL */
  int norm_frames, norm_t0;
  const void* norm_gamma; /* [Cout] bf16 */
  int norm_channels;    /* real channel count C: the result is scaled by sqrt(C) */
  int norm_silu;
} llb_conv3d_desc;
/* CausalConv3d.forward (vae.py:28-36) / the per-frame nn.Conv2d of Resample (vae.py:76-83) as an implicit
 * GEMM on tcgen05: TMA loads one shifted [8 x 16 pixel x 64 (or 32) channel] box per (tap, channel chunk). */
int llb_conv3d(const llb_conv3d_desc* d, void* stream);
/* RMS_norm.forward (vae.py:51-54) over the C real channels of every pixel, optionally followed by SiLU
 * (the RMS_norm -> SiLU pairs of ResidualBlock / head, vae.py:193-199, 415-417); gamma [Cp] bf16, 0 in the pad. */
int llb_vae_norm(const void* in, int in_frames, int in_t0, void* out, int out_frames, int out_t0, int T,
                 int64_t pixels, int Cp, int C, const void* gamma, int silu, void* stream);
/* nn.Upsample(scale_factor=(2,2), mode="nearest") per frame (vae.py:57-63): [T,H,W,Cp] -> [T,2H,2W,Cp] */
int llb_vae_upsample2x(const void* in, void* out, int T, int H, int W, int Cp, void* stream);
/* out[c][r] = in[r][c] (bf16) */
int llb_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int rows, int cols, void* stream);
/* P = softmax(logits * scale) over the first cols_valid columns of each fp32 row, bf16 out, columns
 * [cols_valid, cols_pad) zeroed: the softmax of AttentionBlock's single-head SDPA (vae.py:251-256). */
int llb_softmax_rows(const float* logits, int64_t ld, void* out, int64_t ldo, int rows, int cols_valid,
                     int cols_pad, float scale, void* stream);
/* cached_decode prologue (vae.py:573-579): z [zc, T, h*w] bf16 -> bf16(bf16(z / inv_std) + mean) -> conv2
 * (1x1x1, w [zc, zc], b [zc]) -> channels-last ring [out_frames, h*w, Cp] */
int llb_vae_latent_in(const void* z, const void* mean, const void* inv_std, const void* w, const void* b,
                      void* out, int out_frames, int out_t0, int T, int zc, int64_t hw, int Cp, void* stream);
/* decode_to_pixel epilogue (utils/wan_wrapper.py:112): [T, h*w, Cp] bf16 -> float [T, 3, h*w] clamped to [-1, 1] */
int llb_vae_pixel_out(const void* in, float* out, int T, int64_t hw, int Cp, void* stream);

/* ------------------------------------------------------------------------------------------
 * umT5 text encoder kernels (reference wan/modules/t5.py, called from
 * WanTextEncoder.forward, utils/wan_wrapper.py:43-57, in bf16 after `pipeline.to(dtype=torch.bfloat16)`,
 * inference.py:134).  The block Linears run on llb_gemm_bf16 (no bias; the gated FFN uses BIAS_GELU followed
 * by BIAS_MUL), T5LayerNorm on llb_rmsnorm (same arithmetic as WanRMSNorm; rows up to 8192 wide).
 * Activations are [batch * rows_per_seq, C] bf16; rows_per_seq is the text length rounded as the caller
 * likes (a multiple of 128) - rows >= the sequence's valid length are computed but never influence valid rows.
 * ------------------------------------------------------------------------------------------ */
/* nn.Embedding lookup (t5.py:288): out[b * rows_per_seq + r, :] = table[ids[b * ld_ids + r], :], ids int64 */
int llb_embed_rows(const void* table, int64_t vocab, const void* ids, int64_t ld_ids, void* out, int64_t ldo,
                   int batch, int rows_per_seq, int C, void* stream);
/* T5Attention core (t5.py:96-111), head_dim 64, no 1/sqrt(d) scaling; tcgen05 kernel, rows_per_seq <= 512 (the
 * logits of 128 query rows against 512 keys fill the 512 TMEM columns):
 *   qkv [batch * rows_per_seq, ld_qkv] = q | k | v column blocks of width n_heads * 64 (fused projection output)
 *   logits = bf16(bf16(q . k) + pos_emb[bucket_lut[key - query + lut_center]][head]); keys >= seq_lens[b] get
 *   probability 0 (reference: finfo.min fill); out = bf16(softmax_fp32(logits) V) -> [batch * rows_per_seq, ldo].
 *   pos_emb: the block's T5RelativeEmbedding table [num_buckets, n_heads] bf16 (t5.py:230-247);
 *   bucket_lut: int32 [2 * lut_center + 1], host-evaluated _relative_position_bucket (t5.py:249-268);
 *   max_seq_len: host-known upper bound of seq_lens (sizes the shared-memory key / value stage; <= 0 means
 *   rows_per_seq); lengths above it are clamped. */
int llb_t5_attn(const void* qkv, int64_t ld_qkv, void* out, int64_t ldo, int batch, int rows_per_seq,
                int n_heads, int max_seq_len, const int32_t* seq_lens_dev, const void* pos_emb,
                const int32_t* bucket_lut_dev, int lut_center, void* stream);
/* Final T5LayerNorm (t5.py:294) fused with WanTextEncoder's padding (`u[v:] = 0.0`, wan_wrapper.py:52-53):
 *   out[b, r, :] = r < seq_lens[b] ? norm(x[b * rows_per_seq + r, :]) : 0   for r < rows_out */
int llb_t5_final_norm(const void* x, int64_t ldx, void* out, int64_t ldo, int batch, int rows_per_seq,
                      int rows_out, int C, const void* w, float eps, const int32_t* seq_lens_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LLB200_H_ */

/* tmx.h — C ABI of libtmx.so: the B200 (sm_100a) kernels behind the TweedieMix fusion-sampling
 * hot path.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * The reference (KwonGihyun/TweedieMix) is pure Python with no FFI of its own; each entry point
 * below replaces the arithmetic of a specific span of reference code (cited per function, paths
 * relative to the reference checkout; [D] = diffusers 0.29.2, the reference's pinned dependency).
 * INTEGRATION.md shows the ctypes stubs a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; kernels never allocate, free or
 *     retain pointers past return; workspaces are passed explicitly (tmx_*_workspace_bytes);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), does no host
 *     sync and is CUDA-graph capturable once tmx_init(device) has run;
 *   - return value 0 = success, negative TMX_E* = failure; tmx_last_error() gives a thread-local
 *     message.  There is NO CPU fallback: an unsupported shape/dtype/arch is an error.
 */
#ifndef TMX_H
#define TMX_H

#if defined(__GNUC__)
#define TMX_API __attribute__((visibility("default")))
#else
#define TMX_API
#endif

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMX_VERSION 110            /* 0.1.1 */

/* element types */
#define TMX_F32  0
#define TMX_F16  1
#define TMX_BF16 2

/* error codes */
#define TMX_OK          0
#define TMX_EINVAL     -1          /* bad argument (null pointer, non-positive size, ...) */
#define TMX_ESHAPE     -2          /* shape not supported by the kernel */
#define TMX_EDTYPE     -3          /* dtype not supported */
#define TMX_EALIGN     -4          /* pointer / stride not aligned as required */
#define TMX_EARCH      -5          /* device is not sm_100 */
#define TMX_ECUDA      -6          /* CUDA runtime / driver error (message in tmx_last_error) */
#define TMX_ENOINIT    -7          /* tmx_init(device) was not called for the current device */

/* activation for tmx_groupnorm_fwd */
#define TMX_ACT_NONE 0
#define TMX_ACT_SILU 1

/* memory layout for tmx_groupnorm_fwd */
#define TMX_NCHW 0
#define TMX_NHWC 1

/* rounding mode for the blend kernels */
#define TMX_ROUND_FP32 0           /* all arithmetic in fp32 on the loaded eps */
#define TMX_ROUND_REF  1           /* re-round intermediates to eps' dtype exactly where the
                                      reference's PyTorch type promotion does */

TMX_API int         tmx_version(void);
TMX_API const char* tmx_last_error(void);

/* One-time per-device setup (function attributes, SM count, driver entry points).
 * Must be called outside stream capture.  Idempotent. */
TMX_API int tmx_init(int device);

/* ---------------------------------------------------------------------------------------------
 * k7 — fused CFG combine + Tweedie x0 + mask-weighted concept blend + DDIM update.
 * Replaces fusion_generation/fusion_sampling.py:376-386,430,471-472 (fused phase) and, through
 * `weights`/null `masks`, :392-403 (resampling x0), :407-412 (re-noise), :421-430 (plain CFG),
 * :441-447 (jump).
 *
 *   eps_u      = eps[img][0]
 *   eps_c      = eps_u + g * (eps[img][1+c] - eps_u)                       c = 0..K-1
 *   x0         = sum_c  w_c * m_c(p) * (x - sqrt(1-a_t) * eps_c) / sqrt(a_t)
 *   x_out      = is_last ? x0 : sqrt(a_next) * x0 + sqrt(1-a_next) * eps_u
 *
 * x, x_out, x0_out : fp32 [imgs, C, HW]            (x_out may alias x; x0_out may be NULL)
 * eps              : eps_dtype [imgs, K+1, C, HW]
 * masks            : fp32 [K, HW] shared by all images, or NULL (= all ones)
 * weights          : HOST pointer to K floats, or NULL (= all ones)
 * Requires HW % 8 == 0 and 16-byte aligned pointers.
 * Algorithmic bytes per image: C*HW*(4+4) + (K+1)*C*HW*sizeof(eps) + K*HW*4  (+ C*HW*4 if x0_out).
 */
TMX_API int tmx_tweedie_blend_ddim_fwd(const float* x, const void* eps, const float* masks,
                               const float* weights, float* x_out, float* x0_out,
                               int imgs, int K, int C, int HW,
                               float a_t, float a_next, float g, int is_last,
                               int eps_dtype, int round_mode, void* stream);

/* Concept-parallel (multi-GPU) split of k7, SURVEY §8e.  Linear form (w_c = weights[c] or 1,
 * m_c = masks[c] or 1):
 *   x0 = [ M x - s (1-g) M eps_u - s g sum_c w_c m_c eps_c ] / sqrt(a_t),  M = sum_c w_c m_c, s = sqrt(1-a_t)
 * partial: acc[img][0] = sum_{rows r>0 owned} w_{c(r)} m_{c(r)} * eps_r ; acc[img][1] = eps_u if the
 *          uncond row is owned else 0.  acc is fp32 [imgs, 2, C, HW] and is what gets all-reduced.
 *          R may be 0 (a rank that owns no row of this phase contributes zeros).
 * finish : consumes the all-reduced acc and produces x_out / x0_out on every rank identically.
 * eps_rows : eps_dtype [imgs, R, C, HW] — only the R rows this rank computed;
 * row_ids  : HOST int[R], global row index of each local row (0 = uncond, 1+c = concept c).
 */
TMX_API int tmx_blend_partial_fwd(const void* eps_rows, const float* masks, const float* weights,
                          const int* row_ids, float* acc, int imgs, int R, int K, int C, int HW,
                          int eps_dtype, void* stream);
TMX_API int tmx_blend_finish_fwd(const float* x, const float* acc, const float* masks, const float* weights,
                         float* x_out, float* x0_out, int imgs, int K, int C, int HW,
                         float a_t, float a_next, float g, int is_last, void* stream);

/* ---------------------------------------------------------------------------------------------
 * k4 / k5 — GroupNorm (+ optional per-(n,c) additive bias before the norm, + optional SiLU).
 * Replaces [D] ResnetBlock2D's  norm1 -> SiLU,  (+temb) -> norm2 -> SiLU  (body mirrored in the
 * reference at video_gen/utils_attn.py:391-431), conv_norm_out -> SiLU, and Transformer2DModel.norm.
 *
 *   y = act( (x + add[n,c] - mean[n,g]) * rstd[n,g] * gamma[c] + beta[c] )
 *
 * x, y      : dtype [N, C, HW] (TMX_NCHW) or [N, HW, C] (TMX_NHWC); y may alias x
 * gamma/beta: fp32 [C];  add: fp32 [N, C] or NULL
 * workspace : tmx_groupnorm_workspace_bytes(N, C, HW, G, layout) bytes, 16-byte aligned; its first 8 KiB
 *             (per-row arrival counters, returned to zero by every launch, and per-row barrier generation
 *             words, monotonic) must be ZERO before the first launch — one memset at allocation is enough;
 *             N <= 1024; one launch at a time per workspace (launches on the same stream are fine)
 * NHWC 16-bit activations that fit in one wave of shared memory (N * floor(#SM/N) CTAs x <= 200 KB) run as ONE
 * cooperative launch that reads x once and writes y once; larger ones as two launches (stats, apply).
 * Statistics in fp32 (Chan/Welford merge, deterministic), one rounding on store.
 * Requires C % G == 0, C % 8 == 0 (NHWC) or HW*(C/G) % 8 == 0 (NCHW).
 * Algorithmic bytes: 2 * N*C*HW * sizeof(dtype).
 */
TMX_API size_t tmx_groupnorm_workspace_bytes(int N, int C, int HW, int G, int layout);
TMX_API int    tmx_groupnorm_fwd(const void* x, const float* gamma, const float* beta, const float* add,
                         void* y, void* workspace, int N, int C, int HW, int G, float eps,
                         int act, int layout, int dtype, void* stream);

/* Tuning / test hook: 0 = defaults (one CTA or cluster per (sample, group) where that slab fits in shared memory, else the fused
 * cooperative single launch where the activation fits in one wave of shared memory, else two launches), 1 = always two launches,
 * 2 = fused kernel through a plain (non-cooperative) launch, 3 = per-group slab kernel off, 4 = on again, 5 = on without clusters. */
TMX_API int tmx_groupnorm_set_variant(int v);

/* Same as tmx_groupnorm_fwd (NHWC, 16-bit) over the channel concatenation [x1 | x2] WITHOUT materialising it: x1 is
 * [N, HW, C1], x2 is [N, HW, C - C1] (C1 % 8 == 0), y is [N, HW, C].  Replaces the torch.cat([hidden, skip], dim=1) that feeds
 * norm1 of every up-block ResnetBlock2D ([D] unet_2d_blocks.py) — 9 copies of up to 126 MB per U-Net forward. */
TMX_API int tmx_groupnorm_cat_fwd(const void* x1, const void* x2, int C1, const float* gamma, const float* beta, const float* add,
                          void* y, void* workspace, int N, int C, int HW, int G, float eps, int act, int dtype, void* stream);
/* Number of kernels tmx_groupnorm_fwd launches for this shape (1 fused, 2 stats + apply): launch accounting. */
TMX_API int tmx_groupnorm_launches(int N, int C, int HW, int layout, int dtype);

/* k6 — residual add:  y = (a + b) * inv_scale    ([D] ResnetBlock2D tail, output_scale_factor;
 * also the three residual adds of BasicTransformerBlock).  n elements, n % 8 == 0; y may alias. */
TMX_API int tmx_resadd_fwd(const void* a, const void* b, void* y, size_t n, float inv_scale,
                   int dtype, void* stream);

/* k6b — convolution bias folded into the ResNet tail:  y[r,c] = (a[r,c] + bias[c] + b[r,c]) * inv_scale over an
 * NHWC [rows, C] tensor; b may be NULL (plain per-channel bias add).  Replaces the separate ATen bias-add
 * kernel behind every [D] nn.Conv2d of ResnetBlock2D (conv2 + conv_shortcut biases travel in `bias`) and
 * the `(input + hidden) / output_scale_factor` tail.  fp16 / bf16, C % 8 == 0; bias fp32 [C]; y may alias a or b. */
TMX_API int tmx_bias_resadd_fwd(const void* a, const void* b, const float* bias, void* y, size_t rows, int C,
                        float inv_scale, int dtype, void* stream);

/* k6c — transformer residual add fused with the LayerNorm that consumes it ([D] BasicTransformerBlock:
 * `hidden = attn_output + hidden; norm_hidden = self.norm2(hidden)` and the two analogous pairs):
 *   h = round_dtype(a + b);  n = LayerNorm(h) * gamma + beta.
 * a, b, h_out, n_out : dtype [rows, D]; h_out may alias a or b, n_out must be distinct.  D % 8 == 0, D <= 2048.
 * Algorithmic bytes: 4 * rows * D * sizeof(dtype) (vs 5 for the two separate kernels). */
TMX_API int tmx_resadd_layernorm_fwd(const void* a, const void* b, const float* gamma, const float* beta,
                             void* h_out, void* n_out, size_t rows, int D, float eps, int dtype, void* stream);

/* k8 — LayerNorm over the last dimension ([D] BasicTransformerBlock.norm1/2/3 = F.layer_norm, fp32
 * statistics; 210 sites per U-Net forward).  x, y : dtype [rows, D] (y may alias x); gamma, beta : fp32 [D].
 * fp16 / bf16, D % 8 == 0, D <= 2048.  One pass: algorithmic bytes 2 * rows * D * sizeof(dtype). */
TMX_API int tmx_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y,
                      size_t rows, int D, float eps, int dtype, void* stream);

/* k9a — GEGLU gating  y[m, j] = x[m, j] * gelu_erf(x[m, F + j])  ([D] diffusers GEGLU.forward inside
 * BasicTransformerBlock.ff; 70 sites per U-Net forward).  x : dtype [rows, 2F], y : dtype [rows, F],
 * fp16 / bf16, F % 8 == 0.  Algorithmic bytes: 3 * rows * F * sizeof(dtype). */
TMX_API int tmx_geglu_fwd(const void* x, void* y, size_t rows, int F, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * k13 / k14 — the data-movement ops of the U-Net's up path on NHWC 16-bit tensors (bit-exact copies, 128-bit accesses):
 *   tmx_cat_channels_fwd        y[r, 0:Ca] = a[r, :], y[r, Ca:Ca+Cb] = b[r, :] over rows = N*H*W pixels
 *                               ([D] CrossAttnUpBlock2D / UpBlock2D: torch.cat([hidden_states, res_hidden_states], dim=1))
 *   tmx_upsample_nearest2x_fwd  y[n, 2i+di, 2j+dj, :] = x[n, i, j, :]          x [N, H, W, C] -> y [N, 2H, 2W, C]
 *                               ([D] Upsample2D: F.interpolate(scale_factor=2.0, mode="nearest"))
 * Ca, Cb, C multiples of 8; dtype TMX_F16 or TMX_BF16; the output must not alias an input.
 * Algorithmic bytes: 2 * rows * (Ca + Cb) * 2 (cat);  5 * N*H*W*C * 2 (upsample: 1 read + 4 writes).
 */
TMX_API int tmx_cat_channels_fwd(const void* a, const void* b, void* y, size_t rows, int Ca, int Cb, int dtype, void* stream);
TMX_API int tmx_upsample_nearest2x_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * k1 / k2 — scaled-dot-product attention forward, head dim 64, no mask, non-causal:
 *   O[b,i,h,:] = softmax_j( scale * Q[b,i,h,:] . K[b,j,h,:] ) V[b,j,h,:]
 * tcgen05.mma (kind::f16) with fp32 accumulators in TMEM, operands staged by TMA.
 * Replaces the einsum/softmax/einsum of fusion_generation/utils_custom.py:91-105 and
 * utils_lora.py:99-113 (and the xformers call diffusers makes for attn1 in the custom variant).
 *
 * q, o : dtype [B, Nq, H, 64]; k, v : dtype [B, Nk, H, 64] — i.e. diffusers' [B, N, H*D] tensors
 * before head_to_batch_dim, so no permute is needed on either side.  Row strides (elements)
 * between consecutive tokens are given explicitly so that q/k/v may be slices of a fused QKV
 * projection output: *_stride_n >= H*64, multiple of 8; batch stride = N * stride_n.
 * dtype TMX_F16 or TMX_BF16.  Nq, Nk >= 1 (tails are masked).
 * Algorithmic FLOPs: 4 * B * H * Nq * Nk * 64.
 *
 * Three kernels behind this entry point: k1 (long K/V streams: persistent, stream-K scheduled over (query-tile pair, K/V tile)
 * iterations — units cut along K/V between CTAs are merged through a per-device workspace allocated by tmx_init, so launches of
 * ONE process on DIFFERENT streams of a device must not overlap), k2t (64 < Nk <= 80, the 77 text tokens of attn2: three query tiles
 * in flight per SM) and k2s (other Nk <= 128: streaming warp-level kernel).
 */
TMX_API int tmx_attn_fwd(const void* q, const void* k, const void* v, void* o,
                 int B, int H, int Nq, int Nk, int D,
                 int64_t q_stride_n, int64_t k_stride_n, int64_t v_stride_n, int64_t o_stride_n,
                 float scale, int dtype, void* stream);

/* Tuning / test hook.  0 = defaults; 1 / 2 = one / two 128-row query tiles per step of k1 (and short K/V stays on k1);
 * 11 / 12 = one / two softmax threads per row; 20 / 21 / 22 = units split along K/V between CTAs never / by the cost model / always;
 * 30 .. 33 = short K/V (Nk <= 128) on k1 / k2s with 16 rows per warp / k2s with 32 rows / k2t where it applies (default). */
TMX_API int tmx_attn_set_variant(int nq);

/* ---------------------------------------------------------------------------------------------
 * k3 — per-row routed projection: one weight matrix and / or one set of rank-r LoRA factors per batch row
 *   w != NULL        :  y[b]  = x[b] @ W[b]^T                                   (grouped tcgen05 GEMM)
 *   lora_* != NULL   :  y[b] += per segment s: (x[b] @ down[b][s]^T) @ up[b][cols of s]^T   (rank-r delta)
 * With w == NULL the call only ADDS the LoRA deltas to a y that already holds the shared-weight projection.
 * Replaces the per-row nn.Linear + torch.cat of utils_custom.py:64-82 (concept K/V weights) and the rank-r LoRA
 * deltas of utils_lora.py:65-79,113-119 / model_lora.py:28-48.
 * x : dtype [B, M, Kin] contiguous;  y : dtype [B, M, Nout] contiguous;  B <= 16, Kin % 64 == 0, Nout % 8 == 0
 * w : HOST array of B device pointers to dtype [Nout, Kin] (all non-NULL), or NULL
 * lora_down / lora_up : HOST arrays of B device pointers to dtype [nseg*rank, Kin] / [Nout, rank]; a NULL ENTRY
 *   (in both arrays) leaves that batch row untouched (row 0, the unconditional row, is never routed); NULL arrays =
 *   no LoRA.  Output column n belongs to segment n / (Nout / nseg) — a packed q|k|v projection is nseg = 3 — and
 *   uses rows [s*rank, (s+1)*rank) of down;  nseg*rank in {4, 8, 12, 16}.
 * fp32 accumulation, one rounding per store.  Algorithmic FLOPs: 2*B*M*Kin*Nout (+ 2*B*M*rank*(nseg*Kin + Nout)).
 */
TMX_API int tmx_routed_linear_fwd(const void* x, const void* const* w, const void* const* lora_down,
                          const void* const* lora_up, void* y, int B, int M, int Kin, int Nout,
                          int rank, int nseg, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * k10 — dense projection with a fused epilogue (persistent tcgen05 GEMM, TMA in and out):
 *   TMX_EPI_NONE  :  y[M, N]   = x[M, K] @ W[N, K]^T (+ bias[N]) (+ residual[M, N])
 *   TMX_EPI_GEGLU :  y[M, N/2] = (a_v + bias_v) * gelu_erf(a_g + bias_g), where W (and bias) hold their rows INTERLEAVED in
 *                    blocks of 32: rows [64j, 64j+32) = value rows 32j.., rows [64j+32, 64j+64) = gate rows 32j.. of the
 *                    original [2F, K] GEGLU projection (value = first F rows, gate = last F rows)
 * Replaces to_q / to_k / to_v / to_out[0] of the hooked attention forward (fusion_generation/utils_custom.py:63-89,106;
 * utils_lora.py:65-79,113-121) with the residual add that follows in [D] BasicTransformerBlock folded in, and the [D] GEGLU
 * feed-forward (proj -> value * gelu(gate) -> Linear -> + residual).
 * Optional LoRA tail (utils_lora.py:65-79,113-119 as ONE extra K = 16 MMA step per tile): lora_t = dtype [M, 64] whose
 * columns [0, 16) hold t = x @ down^T of each row's batch row (tmx_lora_t_fwd), lora_up = HOST array of lora_batch device
 * pointers to dtype [N, 64] (columns [0, 16) = the up factors placed in their segment's slots, rest zero) or NULL entries
 * for un-routed batch rows; batch row of x row m = m / lora_rows_per_batch (a multiple of 128).  lora_t == NULL: no tail.
 * x : dtype [M, K], row stride ldx;  W : dtype [N, K] contiguous;  bias : fp32 [N] or NULL;  residual : dtype, row stride ldr,
 * or NULL;  y row stride ldy.  K % 64 == 0, N % 8 == 0 (GEGLU: N % 64 == 0), strides multiples of 8.  fp32 accumulation, one
 * rounding.  Algorithmic FLOPs: 2*M*N*K.
 */
#define TMX_EPI_NONE  0
#define TMX_EPI_GEGLU 1
TMX_API int tmx_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                   int M, int N, int K, int64_t ldx, int64_t ldr, int64_t ldy, int epilogue,
                   const void* lora_t, const void* const* lora_up, int lora_rows_per_batch, int lora_batch,
                   void* workspace, int dtype, void* stream);

/* Workspace of tmx_linear_fwd's split-K tail: when the tile count is not a multiple of the SM count, the tiles of the last,
 * partly filled round are cut along K so that every SM works in it; the fp32 partial accumulators meet in `workspace` and are
 * summed in a fixed order (bit-reproducible).  Allocate tmx_linear_workspace_bytes() once per stream of execution, zero its
 * first 4 KiB ONCE (the kernels re-arm their counters), pass it to every call; NULL disables the split (whole tiles only). */
TMX_API size_t tmx_linear_workspace_bytes(void);

/* Tuning / test hook: tile width of tmx_linear_fwd (0 = heuristic, 128, 192, 256, 320), + 1000 to disable the split-K tail,
 * + 2000 to take it at any K (default: only when K >= 2560). */
TMX_API int tmx_linear_set_variant(int v);

/* t[b*M + m, q] = sum_k x[b, m, k] * down[b][q, k], q < sr = nseg*rank in {4, 8, 12, 16}: the A operand of the LoRA tail
 * (model_lora.py:28-48 `down`).  x : dtype [B, M, K] with row stride ldx;  lora_down : HOST array of B device pointers to
 * dtype [sr, K] (NULL = batch row not routed, its t rows are left untouched);  t : dtype [B*M, 64], zeroed once by the caller
 * (only columns [0, sr) are written).  B <= 16. */
TMX_API int tmx_lora_t_fwd(const void* x, const void* const* lora_down, void* t, int B, int M, int K, int64_t ldx,
                   int sr, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * k11 — video loop (I2VGen-XL stage): classifier-free guidance + v-prediction Tweedie estimate + DDIM update, one pass
 * over the flat latents (position-wise, so the reference's [B,C,F,H,W] <-> [(B F),C,H,W] permutes are not needed):
 *   v = v_u + g (v_c - v_u);  eps = sqrt(a_t) v + sqrt(1-a_t) x;  x0 = sqrt(a_t) x - sqrt(1-a_t) v;
 *   x_next = sqrt(a_next) x0 + sqrt(1-a_next) eps
 * Replaces video_gen/pipeline_i2vgen_xl.py:694-713.  x, v_uncond, v_cond, x_next (, x0_out or NULL): dtype [n], n % 8 == 0;
 * dtype TMX_F16 / TMX_BF16 / TMX_F32; round_mode as for the blend kernels (TMX_ROUND_REF = the fp16 pipeline's roundings).
 * Algorithmic bytes: (4 or 5) * n * sizeof(dtype).
 */
TMX_API int tmx_vpred_cfg_ddim_fwd(const void* x, const void* v_uncond, const void* v_cond, void* x_next, void* x0_out,
                           size_t n, float a_t, float a_next, float guidance, int dtype, int round_mode, void* stream);

/* k12 — frame-0 residual-feature injection, in place, on a ResNet output viewed as [groups, frames, frame_elems]
 * (the reference hard-codes groups = 2, frames = 16): y[g, t >= 1] = interp * y[g, 0] + (1 - interp) * y[g, t];
 * interp = 1 copies frame 0 (`injection_schedule`, video_gen/utils_attn.py:433-443), 0 < interp < 1 is the
 * `injection_schedule2` blend (:445-456).  frame_elems % 8 == 0; dtype TMX_F16 / TMX_BF16.
 */
TMX_API int tmx_frame_inject_fwd(void* y, int groups, int frames, size_t frame_elems, float interp, int dtype, int round_mode,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TMX_H */

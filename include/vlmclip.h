/*
 * vlmclip.h — C ABI of libvlmclip_b200.so, the sm_100a compute library behind the CLIP adapter
 * fine-tuning hot path of Quillboltcode/VLM-CLIP.
 *
 * The reference has no FFI: its seam is the Python nn.Module API.  Each entry point below
 * replaces the PyTorch eager ops the reference issues at the cited file:line; the Python mirror of the
 * reference modules (vlm_clip_b200/*.py) binds them with ctypes.
 *   HF = transformers/models/clip/modeling_clip.py (the third-party module holding the backbone arithmetic).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library never allocates on the hot path and
 *     never synchronises the host.  Work is enqueued on the caller's stream (a CUstream / cudaStream_t
 *     passed as void*).
 *   - return value: 0 = ok, <0 = invalid argument (checked before launch), >0 = cudaError_t.
 *     vlmclip_last_error() returns a thread-local message for the last non-zero return.
 *   - matrices are row-major; "bf16" = __nv_bfloat16 storage; ld* = leading dimension in ELEMENTS.
 *   - entry points are re-entrant (forward on the Python main thread, backward on autograd's thread).
 */
#ifndef VLMCLIP_H_
#define VLMCLIP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLMCLIP_ABI_VERSION 5

/* activation selector for GEMM epilogues and adapter kernels */
enum {
  VLMCLIP_ACT_NONE = 0,
  VLMCLIP_ACT_QUICK_GELU = 1, /* x*sigmoid(1.702x), HF:349 (ACT2FN["quick_gelu"]) */
  VLMCLIP_ACT_GELU_ERF = 2,   /* nn.GELU(), adapter/clip_adapter.py:12, adapter/peclip.py:11 */
  VLMCLIP_ACT_RELU = 3        /* model_t.py:19, model_v.py:22 */
};

/* post-op selector of the fused bottleneck adapter */
enum {
  VLMCLIP_ADAPTER_RESIDUAL_LN = 0, /* LN(up(act(down x)) + x): adapter/clip_adapter.py:17-23,144-150 */
  VLMCLIP_ADAPTER_RESIDUAL = 1,    /* up(act(down x)) + x: adapter/peclip.py:13-18 */
  VLMCLIP_ADAPTER_BLEND_L2 = 2,    /* f = a*up(act(down x)) + (1-a)*x; f/|f|: model_t.py:163-169, model_v.py:280-286 */
  VLMCLIP_ADAPTER_PLAIN = 3        /* up(act(down x)): model_t.py:21-22 */
};

int vlmclip_abi_version(void);
const char* vlmclip_last_error(void);
/* number of kernels launched by this library in the calling process since load (for bench.py gpu_launches) */
int64_t vlmclip_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * Dense layers.  C[M,N] = epilogue(A[M,K] * W[N,K]^T), tcgen05.mma + TMEM accumulators, TMA-fed.
 * Replaces nn.Linear (aten::addmm) at HF:310-312,334 (q/k/v/out), HF:348-350 (fc1/fc2), HF:209 (patch conv as
 * GEMM), HF:784-785 (projections).
 *   epilogue, in order:  v = acc
 *     if row_stats: v = rstd[m] * (v - mean[m] * col_c[n])      (LayerNorm folded into the layer; W holds
 *                                                                gamma*W, col_c[n] = sum_k gamma[k] W[n,k],
 *                                                                bias holds beta.W^T + b; HF:371,380)
 *     if bias:      v += bias[n]
 *     v = act(v)                                                 (HF:349)
 *     if residual:  v += residual[m, n]                          (HF:377,382)
 *   A, W, residual: bf16.  bias, col_c: fp32[N].  row_stats: fp32[M][2] = (mean, rstd).
 *   Instead of row_stats the row statistics can come as partials: stats_part_in fp32 [M][K/32][2] = (mean, M2 =
 *   sum (x - mean)^2) of each 32-column block of the row, combined in the epilogue (Chan) with eps = ln_eps.
 *   stats_part_out (optional, bf16 output, N % 32 == 0): the epilogue writes those partials for the rows it
 *   produces, fp32 [M][N/32][2], so the next LN-folded layer needs no separate statistics pass over HBM.
 *   stats_out + row_counters (optional, with stats_part_out): the CTA that completes the last column tile of a 128-row
 *   block also combines that block's partials into (mean, rstd) rows of stats_out fp32 [M][2] (eps = ln_eps), which
 *   replaces a vlmclip_ln_partials_to_stats launch.  row_counters: ceil(M / 128) 32-bit words, ZERO before the first
 *   call; every launch leaves them zero (one buffer per stream can be reused for ever).
 *   C: bf16 (out_fp32 = 0) or fp32 (out_fp32 = 1).  K % 8 == 0, N % 8 == 0, lda/ldw/ldc/ldr % 8 == 0,
 *   16-byte aligned base pointers.
 * --------------------------------------------------------------------------------------------------------- */
int vlmclip_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc,
                      const float* bias, const void* residual, int64_t ldr, const float* row_stats,
                      const float* col_c, const float* stats_part_in, int npart_in, float ln_eps,
                      float* stats_part_out, float* stats_out, int32_t* row_counters, int M, int N, int K, int act,
                      int out_fp32, void* stream);
/* Two-term residual update of an encoder layer (HF modeling_clip.py:372-373, 382-383 `hidden_states = residual +
 * hidden_states`), in place:  x += A[M,K] * W[N,K]^T + bias, where the residual stream x is stored as TWO bf16 planes,
 * hi = bf16(x) at X and lo = bf16(x - hi) at X + plane_stride elements (both [M, N], leading dimension ldx).  hi is
 * what the next dense layer reads as its bf16 A operand; lo restores 16 mantissa bits on the one tensor that is
 * accumulated over all layers, which takes the towers' end-to-end error from 9e-3 (bf16 stream) to 3.5e-3 (the floor
 * set by the bf16 operand roundings any bf16 execution has).  stats_part_out as in vlmclip_gemm_bf16, computed from
 * the fp32 values before they are split. */
int vlmclip_gemm_bf16_res2(const void* A, int64_t lda, const void* W, int64_t ldw, void* X, int64_t ldx,
                           int64_t plane_stride, const float* bias, float* stats_part_out, float* stats_out,
                           int32_t* row_counters, float ln_eps, int M, int N, int K, void* stream);
/* Split reduction for GEMMs with few output tiles and a long K (the weight gradients dW = dY^T X of the full fine-tune
 * backward, K = all tokens of the batch): C_s[M,N] (fp32) = A[M, K_s] * W[N, K_s]^T for `planes` consecutive slices K_s
 * of K, plane s at C + s * plane_stride floats; (split, output tile) pairs are the work items of the persistent kernel.
 * Returns the number of planes written (>= 1, <= planes) or a negative error code.  vlmclip_sum_planes_f32 adds planes
 * in index order: out[i] = sum_s parts[s * plane_stride + i] (n and plane_stride multiples of 4). */
int vlmclip_gemm_bf16_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc,
                             int64_t plane_stride, int planes, int M, int N, int K, void* stream);
int vlmclip_sum_planes_f32(const float* parts, int64_t plane_stride, int planes, float* out, int64_t n, void* stream);
/* The same split reduction with both operands MN-major: C_s[M,N] = A_t[K_s, M]^T * B_t[K_s, N] (A_t, B_t row-major with K
 * as the row index), i.e. dW = dY^T X read from dY [tokens, N_out] and X [tokens, K_in] as they lie - no transposed
 * copies, no padding of K.  M, N multiples of 8. */
int vlmclip_gemm_bf16_atb_splitk(const void* At, int64_t ldat, const void* Bt, int64_t ldbt, float* C, int64_t ldc,
                                 int64_t plane_stride, int planes, int M, int N, int K, void* stream);
/* out[c] = sum_r x[r, c], x bf16 [R, C] (the bias gradient: column sums of dY); workspace: vlmclip_colsum_bf16_slices(R) * C
 * floats; two deterministic passes. */
int vlmclip_colsum_bf16_slices(int R);
int vlmclip_colsum_bf16(const void* x, int64_t ldx, float* out, float* workspace, int R, int C, void* stream);

/* LayerNorm over the last dimension (eps as given, affine), fp32 statistics.  HF:371,380,562,677.
 *   x: bf16 [M, D] (ldx), y: bf16 [M, D] (ldy).  gamma/beta fp32[D].  stats_out (optional): fp32[M][2]. */
int vlmclip_layernorm_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma,
                           const float* beta, float* stats_out, int M, int D, float eps, void* stream);
/* Same, writing fp32 (used on the few pooled rows that feed the fp32 trainable path: token 0 through
 * final_layer_norm, model_m.py:86,102; CLS through post_layernorm, HF:686).  x rows may be strided (ldx). */
int vlmclip_layernorm_bf16_f32out(const void* x, int64_t ldx, float* y, int64_t ldy, const float* gamma,
                                  const float* beta, int M, int D, float eps, void* stream);
/* Combine the (mean, M2) partials a GEMM epilogue wrote (stats_part_out) into (mean, rstd) per row: fp32 [M][2]. */
int vlmclip_ln_partials_to_stats(const float* partials, float* stats_out, int M, int npart, float eps, void* stream);
/* Row statistics only (mean, rstd) for the LN-folded GEMM epilogue. */
int vlmclip_row_stats_bf16(const void* x, int64_t ldx, float* stats_out, int M, int D, float eps, void* stream);

/* Patch extraction for the vision embedding (HF:209 Conv2d(k=stride=patch) as im2col + GEMM).
 *   pixels: [B,3,H,W] fp32 (pix_bf16=0) or bf16 (pix_bf16=1), NCHW contiguous.
 *   out: bf16 [B*(H/p)*(W/p), Kpad], Kpad = 3*p*p rounded up to a multiple of 64 (zero filled), column index
 *   = c*p*p + i*p + j (matches patch_embedding.weight.view(D,-1), which the caller pads the same way). */
int vlmclip_im2col_patches(const void* pixels, int pix_bf16, void* out, int B, int H, int W, int patch,
                           void* stream);
/* Frame preprocessing fused with patch extraction (process_video.py:14-29 + HF:209): decoded uint8 HWC frames ->
 * (BGR2RGB) -> bilinear resize with OpenCV's 11-bit fixed-point arithmetic -> /255 -> (x - mean) / std -> bf16 im2col
 * rows in the layout of vlmclip_im2col_patches.  frames: n_frames images of [Hs, Ws, 3] uint8, frame_stride bytes
 * apart (any clip / frame nesting that is a constant stride).  ytab [H][3] / xtab [W][3] int32 device tables
 * (source index, weight of it, weight of the next; weights sum to 2048) select the resize; NULL for both = frames
 * are already H x W.  mean / std are per output channel (RGB order). */
int vlmclip_preprocess_patches(const uint8_t* frames, int64_t frame_stride, int Hs, int Ws, int bgr, const int32_t* ytab,
                               const int32_t* xtab, float mean0, float mean1, float mean2, float std0, float std1,
                               float std2, void* out, int n_frames, int H, int W, int patch, void* stream);
/* Temporal mean-pool of per-frame features (not in the reference, defined as
 * get_image_features(frames).view(B, T, P).mean(1)): y[b] = mean_t x[b*T + t]; and its backward dx = dy / T. */
int vlmclip_mean_pool(const float* x, float* y, int B, int T, int P, void* stream);
int vlmclip_mean_pool_bwd(const float* dy, float* dx, int B, int T, int P, void* stream);
/* Assemble vision tokens and apply pre_layrnorm (HF:211-218, HF:677):
 *   x[b,0] = cls + pos[0]; x[b,1+p] = patch[b,p] + pos[1+p]; y = LN(x).
 *   patch: [B*(S-1), D], fp32 (patch_bf16 = 0: the patch GEMM ran with out_fp32 = 1) or bf16 (patch_bf16 = 1: the
 *   patch GEMM's bf16 output, half the traffic; what an autocast reference run produces); the position add and the
 *   LayerNorm are fp32 either way.  cls fp32[D]; pos fp32[S,D]; y bf16 [B*S, D]; y_lo (optional) bf16 [B*S, D]
 *   receives bf16(value - y), the second term of the two-term residual stream (vlmclip_gemm_bf16_res2). */
int vlmclip_vision_embed_ln(const void* patch, int patch_bf16, const float* cls, const float* pos, const float* gamma,
                            const float* beta, void* y, void* y_lo, int B, int S, int D, float eps, void* stream);
/* Text embeddings (HF:234-258): y[b,s] = tok[ids[b,s]] + pos[s].  ids int64 [B,S]; tok fp32[V,D]
 * or bf16 (tok_bf16=1); pos fp32 [>=S, D]; y bf16 [B*S, D]; y_lo optional (see vlmclip_vision_embed_ln).  Out-of-range ids -> return -1 is NOT possible
 * without a sync, so ids are clamped to [0,V) on device (the reference would raise IndexError). */
int vlmclip_text_embed(const int64_t* ids, const void* tok, int tok_bf16, const float* pos, void* y, void* y_lo, int B,
                       int S, int D, int V, void* stream);

/* Multi-head self-attention core (HF:261-279,318-331): softmax(q k^T * scale + mask) v, fp32 softmax.
 *   qkv: bf16 [B*S, 3*D] rows = tokens, columns = [q | k | v], head h at columns h*64..h*64+63 of each third.
 *   out: bf16 [B*S, D].  head_dim fixed at 64 (all CLIP towers).  causal: 0/1 (HF:546-551).
 *   key_mask (optional): uint8 [B,S], 1 = attend (attention_mask), applied to keys as in create_causal_mask.
 *   A query row whose keys are all masked produces zeros. */
int vlmclip_attention_fwd(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H, int causal,
                          float scale, void* stream);
/* Same operator with a caller-provided scratch buffer of vlmclip_attention_fwd_workspace(B, S, H) floats (0 = none
 * needed).  With it, unmasked sequences of 289..384 tokens (225..384 with VLMCLIP_ATTN_SPLIT=1..4; ViT-L/14's S = 257
 * defaults to the mma.sync kernel, which measures faster there) run on the tcgen05 kernel as two key ranges whose
 * partial softmaxes are merged on the device (reference exponent and row sum per row and head live in the workspace
 * between the two launches); workspace = NULL behaves exactly like vlmclip_attention_fwd. */
int64_t vlmclip_attention_fwd_workspace(int B, int S, int H);
int vlmclip_attention_fwd_ws(const void* qkv, void* out, const uint8_t* key_mask, float* workspace, int B, int S,
                             int H, int causal, float scale, void* stream);

/* Single-query attention, head_dim 64: out[b] = softmax(q[b] K^T * scale) V per head, ONE query row per batch element.
 * Replaces nn.MultiheadAttention's core in SharedMHSAttentionAdapter (adapter/clip_adapter.py:114, as called from
 * model_m.py:93-100 with the vision position table as keys/values and only token 0 consumed, model_m.py:102), and
 * serves the CLS-only evaluation of the last vision layer.  q: bf16, row b at q + b*q_stride (elements), heads
 * contiguous 64-wide; k, v: bf16, key j of batch b at k + b*kv_batch_stride + j*kv_row_stride (kv_batch_stride = 0:
 * keys/values shared by the batch); out: bf16 [B, H*64].  S <= 512. */
int vlmclip_attention_1q(const void* q, int64_t q_stride, const void* k, const void* v, int64_t kv_row_stride,
                         int64_t kv_batch_stride, void* out, int B, int S, int H, float scale, void* stream);

/* One frozen tower, all layers, in one call: the launch sequence of towers.NativeClipTowers._encoder (the reference's
 * CLIPEncoder loop, HF modeling_clip.py:355-386 x num_hidden_layers) issued natively.  Same kernels as the per-op
 * entry points above; this exists because issuing ~170 ops per step from the interpreter cost more host time than the
 * GPU needs to run them.  LayerNorm is folded: qkv_w / fc1_w hold bf16(gamma * W), qkv_b / fc1_b hold W beta + b,
 * qkv_c / fc1_c the fold column sums (see vlmclip_gemm_bf16).
 *   x    bf16 [B*S, D], updated in place (the residual stream); qkv bf16 [B*S, 3D], att bf16 [B*S, D],
 *   hid  bf16 [B*S, F], stats fp32 [B*S, 2], part fp32 [B*S, D/32, 2]: workspaces.  head_dim is 64 (D = 64 H). */
typedef struct {
  const void* qkv_w; /* bf16 [3D, D] */
  const float* qkv_b;
  const float* qkv_c;
  const void* out_w; /* bf16 [D, D] */
  const float* out_b;
  const void* fc1_w; /* bf16 [F, D] */
  const float* fc1_b;
  const float* fc1_c;
  const void* fc2_w; /* bf16 [D, F] */
  const float* fc2_b;
} vlmclip_layer_t;
/* x_lo (optional): second plane of the two-term residual stream, bf16 [B*S, D] (see vlmclip_gemm_bf16_res2); NULL keeps
 * the stream in one bf16 plane (half the residual traffic, 2.5x the end-to-end rounding error). */
/* row_counters (optional): ceil(B*S / 128) zeroed 32-bit words (see vlmclip_gemm_bf16); with them the out-proj / fc2
 * epilogues finalise the LayerNorm statistics themselves and the 2 L - 1 ln_partials_to_stats launches disappear. */
int vlmclip_encoder_fwd(const vlmclip_layer_t* layers, int n_layers, void* x, void* x_lo, void* qkv, void* att, void* hid,
                        float* stats, float* part, int32_t* row_counters, const uint8_t* key_mask, int B, int S, int H,
                        int D, int F, float eps, int causal, int act, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused bottleneck adapter, fp32 (trainable weights live in fp32; adapter/clip_adapter.py:4-23,131-150,
 * adapter/peclip.py:6-18, model_t.py:13-33, model_v.py:18-27).
 *   h = act(x W1^T + b1);  u = h W2^T + b2;  y = post(u, x)     x,y: fp32 [R, D] (ldx); W1 [A,D]; W2 [D,A]
 *   post: see VLMCLIP_ADAPTER_*; gamma/beta used by RESIDUAL_LN (eps), alpha by BLEND_L2.
 *   x may be a strided view (ldx >= D), e.g. token 0 of every sequence (model_m.py:102,122).
 *   x_bf16 = 1: x is bf16 (backbone activations), converted on load.
 *   hmask (optional): fp32 [R, A] multiplier applied to h after the activation (the dropout mask of
 *   model_v.py:24, already scaled by 1/(1-p); NULL in eval mode).
 * Backward produces adapter-only gradients (the backbone is frozen: model_m.py:64-70); dx only if dx != NULL
 * (full fine-tune).  Gradients are WRITTEN (not accumulated).  workspace: fp32, size from
 * vlmclip_adapter_bwd_workspace(R, D, A) elements.
 * --------------------------------------------------------------------------------------------------------- */
int vlmclip_adapter_fwd(const void* x, int x_bf16, int64_t ldx, const float* W1, const float* b1,
                        const float* W2, const float* b2, const float* gamma, const float* beta,
                        const float* hmask, float* y, int R, int D, int A, int act, int post, float alpha,
                        float eps, void* stream);
int64_t vlmclip_adapter_bwd_workspace(int R, int D, int A);
int vlmclip_adapter_bwd(const void* x, int x_bf16, int64_t ldx, const float* W1, const float* b1,
                        const float* W2, const float* b2, const float* gamma, const float* beta,
                        const float* hmask, const float* dy, float* dW1, float* db1, float* dW2, float* db2, float* dgamma,
                        float* dbeta, float* dx, float* workspace, int R, int D, int A, int act, int post,
                        float alpha, float eps, void* stream);

/* Small fp32 linear for the projections (HF:784-785; model_m.py:103,123): y[R,N] = x[R,K] W[N,K]^T (+b),
 * and its input gradient dx[R,K] = dy[R,N] W[N,K].  fp32 SIMT (K,N <= 1024-ish, R = batch). */
int vlmclip_linear_f32(const float* x, int64_t ldx, const float* W, const float* b, float* y, int R, int N,
                       int K, void* stream);
int vlmclip_linear_f32_dgrad(const float* dy, const float* W, float* dx, int R, int N, int K, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Contrastive head (model_m.py:146-171): L2-normalise, exp(logit_scale) * T I^T, symmetric cross-entropy,
 * and its gradient w.r.t. the UN-normalised features.
 *   txt, img: fp32 [N, P] un-normalised features of the GLOBAL batch (after all-gather under DP).
 *   row0, nloc: this rank's rows [row0, row0+nloc) — gradients are produced for these rows only, including
 *               both the row-softmax and the column-softmax terms (exact dL_global/d(local features)).
 *   outputs: txt_n, img_n fp32 [N,P] normalised; logits_per_text fp32 [N,N] (optional, may be NULL);
 *            loss fp32[1]; d_txt, d_img fp32 [nloc, P] (optional); d_logit_scale fp32[1] (optional; gradient
 *            w.r.t. the log-scale parameter, local rows' share).
 *   workspace fp32: vlmclip_clip_loss_workspace(N, P) elements.
 * --------------------------------------------------------------------------------------------------------- */
int64_t vlmclip_clip_loss_workspace(int N, int P);
int vlmclip_clip_loss(const float* txt, const float* img, float logit_scale_exp, float* txt_n, float* img_n,
                      float* logits_per_text, float* loss, float* d_txt, float* d_img, float* d_logit_scale,
                      float* workspace, int N, int P, int row0, int nloc, void* stream);

/* The same loss on STRIPS of the logit matrix, three launches instead of eight (csrc/clip_loss.cu).  A rank that owns
 * rows [row0, row0 + nloc) of the global batch computes Zt = its text rows x all images and Zi = its image rows x all
 * texts, which give the text-side LSE of its rows and the image-side LSE of its columns.  Under data parallelism the
 * ranks all-gather their [lse_t | lse_i | loss share] blocks (2 nloc + 1 floats) between the two calls; a single process
 * passes lse_loc straight on.
 *   state     fp32 scratch carried from _fwd to _bwd (vlmclip_clip_loss_state_size(N, P, nloc) floats)
 *   counters  vlmclip_clip_loss_counters(nloc) 32-bit words, ZERO before the first call; every call leaves them zero
 *   _fwd      txt / img [N, P] (all-gathered, un-normalised) -> txt_n / img_n [N, P], lse_loc [2 nloc], loss_share [1]
 *             (the loss when nloc == N), logits_per_text [N, N] (optional, nloc == N only)
 *   _bwd      d_txt / d_img [nloc, P]: gradient of the GLOBAL loss w.r.t. the un-normalised local rows.  lse_all: the
 *             gathered blocks, text-side LSE of global row r at lse_all[(r / rows_per_rank) * lse_stride + r %
 *             rows_per_rank], the image-side one rows_per_rank further.  The strips in `state` may cover more rows than
 *             [row0, row0 + nloc): [strip_row0, strip_row0 + strip_rows), the range the _fwd call was made for.
 *             workspace: vlmclip_clip_loss_bwd_workspace floats.  d_logit_scale (optional): dL/d(log scale), local share. */
int64_t vlmclip_clip_loss_state_size(int N, int P, int nloc);
/* y[i] = a[i] * scalar_dev[0] (the upstream gradient of the scalar loss is a device scalar) */
int vlmclip_scale_f32(const float* a, const float* scalar_dev, float* y, int64_t n, void* stream);
int64_t vlmclip_clip_loss_counters(int nloc);
int64_t vlmclip_clip_loss_bwd_workspace(int N, int P, int nloc);
int vlmclip_clip_loss_fwd(const float* txt, const float* img, float logit_scale_exp, float* txt_n, float* img_n,
                          float* logits_per_text, float* lse_loc, float* loss_share, float* state, int32_t* counters, int N,
                          int P, int row0, int nloc, void* stream);
int vlmclip_clip_loss_bwd(const float* txt_n, const float* img_n, const float* lse_all, int lse_stride, int rows_per_rank,
                          float logit_scale_exp, float* d_txt, float* d_img, float* d_logit_scale, float* state,
                          int32_t* counters, float* workspace, int N, int P, int row0, int nloc, int strip_row0,
                          int strip_rows, void* stream);

/* Class-prompt head (model_t.py:184-187,213-242; model_v.py:340-343): logits[B,C] = scale * f_img f_txt^T,
 * cross-entropy against int64 labels (hard) or fp32 [B,C] probabilities (soft), mean reduction, plus
 * gradients w.r.t. f_img, f_txt.  probs (optional) = softmax(logits).  group > 1: logits are first
 * max-reduced over `group` consecutive prompts per class (predict_with_all_descriptions, model_t.py:264-298;
 * forward only; f_txt is then [C*group, P]).  workspace: fp32 B*C + B elements, needed when loss or
 * gradients are requested. */
int vlmclip_class_head(const float* f_img, const float* f_txt, float scale, const int64_t* labels,
                       const float* soft_labels, float* logits, float* probs, float* loss, float* d_img,
                       float* d_txt, float* workspace, int B, int C, int P, int group, void* stream);

/* Back-propagate a given dlogits [B,C] through logits = scale * f_img f_txt^T (for callers that apply their own
 * criterion to the logits, main.py:78-84).  d_img [B,P] and/or d_txt [C,P] are written. */
int vlmclip_class_head_bwd(const float* f_img, const float* f_txt, const float* dlogits, float scale, float* d_img,
                           float* d_txt, int B, int C, int P, void* stream);

/* L2 normalise rows: y = s / |s| with s = x (+ x2 if given) (model_t.py:160; model_v.py:306-315 where
 * normalise((a+b)/2) == normalise(a+b)).  sum_out (optional) receives s for the backward. */
int vlmclip_l2norm_rows(const float* x, const float* x2, float* y, float* sum_out, int R, int P, void* stream);
int vlmclip_l2norm_rows_bwd(const float* x, const float* dy, float* dx, int R, int P, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Optimiser tail (trainer.py:91-99): global grad-norm, clip to max_norm, AdamW (decoupled weight decay),
 * one launch over a flat fp32 parameter arena, no host sync.  step_size/bias corrections are computed on
 * device from `step` (int32[1], incremented by the kernel).  lr is read from lr_dev (fp32[1]) so a host
 * scheduler can update it asynchronously.  grad_norm_out fp32[1] receives the pre-clip norm.
 * --------------------------------------------------------------------------------------------------------- */
int vlmclip_adamw_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                            const float* lr_dev, float beta1, float beta2, float eps, float weight_decay,
                            float max_norm, int32_t* step, float* grad_norm_out, float* workspace,
                            void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Backbone backward (full fine-tune, BASELINE config 5: `CLIPWithAdapters(freeze_clip=False)`, model_m.py:22,72-75).
 * The reference gets these from autograd over HF modeling_clip.py; here every dense product is put into the
 * C = A W^T form of vlmclip_gemm_bf16 (dgrad: A = dY, W = W^T; wgrad: A = dY^T, W = X^T, fp32 output) by the
 * transposes below, and the row-wise pieces are the kernels of csrc/backward.cu / csrc/attention_bwd.cu.
 * --------------------------------------------------------------------------------------------------------- */
/* dst[c, r] = bf16(src[row(r), c]), r < R; dst[c, R..Rpad) = 0 (Rpad: the GEMM's K, a multiple of 8).  src is fp32
 * (src_f32 = 1: a master weight) or bf16 (an activation / gradient), [.., C] with lds; dst bf16 [C, ldd].
 * Row gather (group_dst > 0): row(r) = (r / group_dst) * group_src + group_off + r % group_dst, e.g. (S-1, S, 1)
 * drops the CLS row of every image (patch-embedding weight gradient, HF:209). */
int vlmclip_transpose_to_bf16(const void* src, int src_f32, int64_t lds, void* dst, int64_t ldd, int R, int Rpad, int C,
                              int group_dst, int group_src, int group_off, void* stream);
/* fp32 -> bf16 (per-step cast of the trainable master weights, and of the fp32 gradient stream for the next GEMM) */
int vlmclip_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
/* acc[i] += float(b[i]), b bf16: a bf16 gradient branch joins the fp32 gradient of a skip connection */
int vlmclip_add_bf16_into_f32(float* acc, const void* b, int64_t n, void* stream);
/* out[r] = sum_c x[r, c], x bf16 [R, C] (ldx): bias gradient = row sums of the transposed output gradient */
int vlmclip_rowsum_bf16(const void* x, int64_t ldx, float* out, int R, int C, void* stream);
/* out[c] = sum_r x[r*ldx + c], fp32: position / class embedding gradients (sum over the batch), HF:214-217,255-256 */
int vlmclip_colsum_f32(const float* x, int64_t ldx, float* out, int R, int64_t C, void* stream);
/* quick_gelu (HF:349) as a stand-alone op (training keeps the pre-activation) and its backward da = dy * g'(a) */
int vlmclip_quick_gelu_bf16(const void* a, void* y, int64_t n, void* stream);
int vlmclip_quick_gelu_bwd_bf16(const void* a, const void* dy, void* da, int64_t n, void* stream);
/* LayerNorm backward (HF:371,380,562,677; adapter LN is handled inside vlmclip_adapter_bwd).
 *   x: the bf16 INPUT of the LayerNorm [M, D] (ldx; statistics are recomputed from it); dy: gradient of its output,
 *   bf16 (dy_f32 = 0) or fp32 (dy_f32 = 1), lddy.  dres (optional, fp32, lddx): gradient arriving over the skip
 *   connection, added to dx.  dx is written as fp32 (dx_f32) and / or bf16 (dx_bf16), both with lddx.
 *   dgamma / dbeta fp32 [D] (optional, WRITTEN); workspace: vlmclip_layernorm_bwd_workspace(M, D) floats. */
int64_t vlmclip_layernorm_bwd_workspace(int M, int D);
int vlmclip_layernorm_bwd(const void* dy, int dy_f32, int64_t lddy, const void* x, int64_t ldx, const float* gamma,
                          const float* dres, float* dx_f32, void* dx_bf16, int64_t lddx, float* dgamma, float* dbeta,
                          float* workspace, int M, int D, float eps, void* stream);
/* Vision tokens without pre_layrnorm (HF:211-218): e[b,0] = cls + pos[0]; e[b,1+p] = patch[b,p] + pos[1+p], bf16.
 * Training keeps e as the saved input of the pre-LayerNorm. */
int vlmclip_vision_embed(const void* patch, const float* cls, const float* pos, void* e, int B, int S, int D,
                         void* stream);
/* token_embedding.weight gradient (HF:248): dtok[ids[r]] += d[r] (fp32 atomics; dtok zeroed by the caller) */
int vlmclip_embed_scatter_add(const float* d, int64_t ldd, const int64_t* ids, float* dtok, int64_t rows, int D, int V,
                              void* stream);
/* Backward of vlmclip_attention_fwd: dqkv [B*S, 3D] (bf16, WRITTEN) from qkv, out = forward output [B*S, D] and its
 * gradient dout.  Probabilities are recomputed (nothing else is kept from the forward).  head_dim 64.
 *   workspace: vlmclip_attention_bwd_workspace(B, S, H) floats (row log-sum-exp and dO.O) -> tensor-core kernels
 *   (mma.sync bf16, S <= 512); NULL -> the fp32 SIMT reference kernel (S <= 288), kept for the parity tests. */
int64_t vlmclip_attention_bwd_workspace(int B, int S, int H);
int vlmclip_attention_bwd(const void* qkv, const void* out, const void* dout, void* dqkv, const uint8_t* key_mask,
                          float* workspace, int B, int S, int H, int causal, float scale, void* stream);
/* dW[N,K] = dy[R,N]^T x[R,K] (fp32): weight gradient of vlmclip_linear_f32 when the projection is trainable */
int vlmclip_linear_f32_wgrad(const float* dy, const float* x, int64_t ldx, float* dW, int R, int N, int K, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Trainable SharedMHSAttentionAdapter (adapter/clip_adapter.py:69-128 + autograd), fp32 rows: the pieces that the
 * fp32 linear kernels above do not cover.  Track M evaluates the adapter on token 0 of every caption against the
 * projected vision position table (model_m.py:93-102): B query rows, one [S, D] key/value table for the batch.
 * --------------------------------------------------------------------------------------------------------- */
/* y = LN(x) on fp32 rows (x row stride ldx, y contiguous [M, D]); stats [M][2] = (mean, rstd) kept for the backward */
int vlmclip_layernorm_f32(const float* x, int64_t ldx, const float* gamma, const float* beta, float* y, float* stats,
                          int M, int D, float eps, void* stream);
/* dx = LN'(dy) (+ dres) [M, D] (optional), dgamma / dbeta [D] (optional, WRITTEN) */
int vlmclip_layernorm_f32_bwd(const float* dy, const float* x, int64_t ldx, const float* stats, const float* gamma,
                              const float* dres, float* dx, float* dgamma, float* dbeta, int M, int D, void* stream);
/* nn.GELU() (exact erf) on fp32 and its backward da = dy * gelu'(a) */
int vlmclip_gelu_f32(const float* a, float* y, int64_t n, void* stream);
int vlmclip_gelu_f32_bwd(const float* a, const float* dy, float* da, int64_t n, void* stream);
/* y = a (* mask) (+ b), mask / b optional: dropout application (mask already scaled by 1/(1-p)) and residual adds */
int vlmclip_fma_mask_f32(const float* a, const float* mask, const float* b, float* y, int64_t n, void* stream);
/* Single-query multi-head attention against a table shared by the batch, head_dim 64 (nn.MultiheadAttention's core,
 * adapter/clip_adapter.py:114): q [B, H*64]; k, v [S, H*64] (row stride ldkv); p_out [B, H, S] = softmax(q k^T scale)
 * (kept for the backward); pmask (optional) [B, H, S] multiplies p (attention dropout); out [B, H*64] = (p*pmask) v.
 * Backward: dq [B, H*64], dk / dv [S, H*64] (summed over the batch, WRITTEN); ds_ws: B*H*S floats of workspace. */
int vlmclip_attn1q_f32_fwd(const float* q, const float* k, const float* v, int64_t ldkv, const float* pmask, float* p_out,
                           float* out, int B, int S, int H, float scale, void* stream);
int vlmclip_attn1q_f32_bwd(const float* dout, const float* q, const float* k, const float* v, int64_t ldkv, const float* p,
                           const float* pmask, float* ds_ws, float* dq, float* dk, float* dv, int B, int S, int H,
                           float scale, void* stream);

/* bf16 <-> fp32 casts with optional row gather (token-0 slice): y[r, :] = x[r*ldx : r*ldx + D] */
int vlmclip_gather_rows_bf16_to_f32(const void* x, int64_t ldx, float* y, int R, int D, void* stream);
/* same from a two-term stream: y[r, :] = float(x[r*ldx ...]) + float(x_lo[r*ldx ...]) (x_lo may be NULL) */
int vlmclip_gather_rows2_bf16_to_f32(const void* x, const void* x_lo, int64_t ldx, float* y, int R, int D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLMCLIP_H_ */

// Contrastive and class-prompt heads, fp32.
//
//  vlmclip_clip_loss   model_m.py:146-171: t^ = t/|t|, i^ = i/|i|, Z = exp(logit_scale) t^ i^T,
//                      loss = (CE(Z, arange) + CE(Z^T, arange)) / 2, and its gradient w.r.t. the un-normalised
//                      features of this rank's rows (G = [(softmax_rows(Z) - I) + (softmax_cols(Z) - I)] / 2N).
//  vlmclip_class_head  model_t.py:184-187,213-298 / model_v.py:340-343: logits = scale f_img f_txt^T,
//                      cross-entropy (hard or soft labels), softmax probabilities, max over prompt groups.
//
// Log-sum-exp reductions are held in registers and combined with warp shuffles; the batch-coupled products are
// 64x64 SIMT tiles (tile_f32.cuh).  Everything is deterministic (no atomics).
#include "../../include/vlmclip.h"
#include "common.cuh"
#include "tile_f32.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

// ---------------------------------------------------------------- L2 normalise
// y = s / |s|, s = x (+ x2): the optional second operand is the context branch of model_v.py:306-315, where
// normalise((a + b) / 2) == normalise(a + b).  sum_out (optional) keeps s for the backward.
__global__ void __launch_bounds__(256)
l2norm_rows_kernel(const float* __restrict__ x, const float* __restrict__ x2, float* __restrict__ y,
                   float* __restrict__ inv_out, float* __restrict__ sum_out, int R, int P) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (int64_t)row * P;
  const float* x2r = x2 != nullptr ? x2 + (int64_t)row * P : nullptr;
  float q = 0.f;
  for (int c = lane; c < P; c += 32) {
    const float v = xr[c] + (x2r != nullptr ? x2r[c] : 0.f);
    q = fmaf(v, v, q);
  }
  const float inv = 1.f / sqrtf(warp_sum(q));
  for (int c = lane; c < P; c += 32) {
    const float v = xr[c] + (x2r != nullptr ? x2r[c] : 0.f);
    y[(int64_t)row * P + c] = v * inv;
    if (sum_out != nullptr) sum_out[(int64_t)row * P + c] = v;
  }
  if (inv_out != nullptr && lane == 0) inv_out[row] = inv;
}

// dx = (dy - y (y . dy)) / |x|, y = x/|x|
__global__ void __launch_bounds__(256)
l2norm_rows_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int R,
                       int P) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (int64_t)row * P;
  const float* gr = dy + (int64_t)row * P;
  float q = 0.f, d = 0.f;
  for (int c = lane; c < P; c += 32) {
    q = fmaf(xr[c], xr[c], q);
    d = fmaf(xr[c], gr[c], d);
  }
  q = warp_sum(q);
  d = warp_sum(d);
  const float inv = 1.f / sqrtf(q);
  const float dot = d * inv;  // y . dy
  for (int c = lane; c < P; c += 32) dx[(int64_t)row * P + c] = (gr[c] - xr[c] * inv * dot) * inv;
}

// ---------------------------------------------------------------- similarity logits
// Z[i][j] = s * sum_p tn[i][p] * in[j][p]
__global__ void __launch_bounds__(TF_THREADS)
sim_logits_kernel(const float* __restrict__ tn, const float* __restrict__ in_, float* __restrict__ Z, float s, int N,
                  int P) {
  const int m0 = blockIdx.x * TF_TILE, n0 = blockIdx.y * TF_TILE;
  tile_gemm_f32(
      P, m0, n0, true, true, [&](int m, int k) { return (m < N && k < P) ? tn[(int64_t)m * P + k] : 0.f; },
      [&](int n, int k) { return (n < N && k < P) ? in_[(int64_t)n * P + k] : 0.f; },
      [&](int m, int n, float v) {
        if (m < N && n < N) Z[(int64_t)m * N + n] = s * v;
      });
}

// row log-sum-exp: one warp per row.  lse_r[i] and the per-row loss term (lse_r[i] - Z[i][i])
__global__ void __launch_bounds__(256)
row_lse_kernel(const float* __restrict__ Z, float* __restrict__ lse_r, int N) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  const float* zr = Z + (int64_t)row * N;
  float m = -INFINITY;
  for (int c = lane; c < N; c += 32) m = fmaxf(m, zr[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < N; c += 32) s += __expf(zr[c] - m);
  s = warp_sum(s);
  if (lane == 0) lse_r[row] = m + logf(s);
}

// column log-sum-exp: block = 32 columns x 8 row-slices, online (max, sum) pairs merged through smem
__global__ void __launch_bounds__(256)
col_lse_kernel(const float* __restrict__ Z, float* __restrict__ lse_c, int N) {
  __shared__ float sm_m[8][33], sm_s[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int slice = threadIdx.x >> 5;
  float m = -INFINITY, s = 0.f;
  if (col < N) {
    for (int r = slice; r < N; r += 8) {
      const float z = Z[(int64_t)r * N + col];
      if (z > m) {
        s = s * __expf(m - z) + 1.f;
        m = z;
      } else {
        s += __expf(z - m);
      }
    }
  }
  sm_m[slice][threadIdx.x & 31] = m;
  sm_s[slice][threadIdx.x & 31] = s;
  __syncthreads();
  if (slice == 0 && col < N) {
    float M = -INFINITY;
    for (int k = 0; k < 8; ++k) M = fmaxf(M, sm_m[k][threadIdx.x]);
    float S = 0.f;
    for (int k = 0; k < 8; ++k)
      if (sm_m[k][threadIdx.x] > -INFINITY) S += sm_s[k][threadIdx.x] * __expf(sm_m[k][threadIdx.x] - M);
    lse_c[col] = M + logf(S);
  }
}

// loss = (1/2N) sum_i [(lse_r[i] - Z[i][i]) + (lse_c[i] - Z[i][i])]   (single block, fixed order)
__global__ void __launch_bounds__(256)
clip_loss_reduce_kernel(const float* __restrict__ Z, const float* __restrict__ lse_r,
                        const float* __restrict__ lse_c, float* __restrict__ loss, int N) {
  __shared__ float part[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += 256) {
    const float d = Z[(int64_t)i * N + i];
    s += (lse_r[i] - d) + (lse_c[i] - d);
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = part[0] / (2.f * (float)N);
}

__device__ __forceinline__ float clip_G(const float* Z, const float* lse_r, const float* lse_c, int N, int i, int j) {
  const float z = Z[(int64_t)i * N + j];
  float g = __expf(z - lse_r[i]) + __expf(z - lse_c[j]);
  if (i == j) g -= 2.f;
  return g / (2.f * (float)N);
}

// blockIdx.z = 0: dtn[i][p] = s * sum_j G[i][j] in[j][p]   (i in local rows)
// blockIdx.z = 1: din[j][p] = s * sum_i G[i][j] tn[i][p]   (j in local rows)
__global__ void __launch_bounds__(TF_THREADS)
clip_grad_kernel(const float* __restrict__ Z, const float* __restrict__ lse_r, const float* __restrict__ lse_c,
                 const float* __restrict__ tn, const float* __restrict__ in_, float* __restrict__ dtn,
                 float* __restrict__ din, float s, int N, int P, int row0, int nloc) {
  const int m0 = blockIdx.x * TF_TILE, n0 = blockIdx.y * TF_TILE;  // m over local rows, n over P
  if (blockIdx.z == 0) {
    tile_gemm_f32(
        N, m0, n0, true, false,
        [&](int m, int j) { return (m < nloc && j < N) ? clip_G(Z, lse_r, lse_c, N, row0 + m, j) : 0.f; },
        [&](int n, int j) { return (n < P && j < N) ? in_[(int64_t)j * P + n] : 0.f; },
        [&](int m, int n, float v) {
          if (m < nloc && n < P) dtn[(int64_t)m * P + n] = s * v;
        });
  } else {
    tile_gemm_f32(
        N, m0, n0, false, false,
        [&](int m, int i) { return (m < nloc && i < N) ? clip_G(Z, lse_r, lse_c, N, i, row0 + m) : 0.f; },
        [&](int n, int i) { return (n < P && i < N) ? tn[(int64_t)i * P + n] : 0.f; },
        [&](int m, int n, float v) {
          if (m < nloc && n < P) din[(int64_t)m * P + n] = s * v;
        });
  }
}

// back through the normalisation for the local rows: d = (dn - xn (xn . dn)) * inv; also the local share of
// dL/d(log scale) = sum over local rows i of sum_j [Prow-δ]Z/2N + sum over local cols j of sum_i [Pcol-δ]Z/2N
__global__ void __launch_bounds__(256)
clip_norm_bwd_kernel(const float* __restrict__ tn, const float* __restrict__ in_, const float* __restrict__ inv_t,
                     const float* __restrict__ inv_i, const float* __restrict__ dtn, const float* __restrict__ din,
                     float* __restrict__ d_txt, float* __restrict__ d_img, int P, int row0, int nloc) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= nloc) return;
  const int lane = threadIdx.x & 31;
  const bool is_img = blockIdx.y == 1;
  const float* xn = (is_img ? in_ : tn) + (int64_t)(row0 + r) * P;
  const float* dn = (is_img ? din : dtn) + (int64_t)r * P;
  const float inv = (is_img ? inv_i : inv_t)[row0 + r];
  float* out = (is_img ? d_img : d_txt) + (int64_t)r * P;
  float d = 0.f;
  for (int c = lane; c < P; c += 32) d = fmaf(xn[c], dn[c], d);
  d = warp_sum(d);
  for (int c = lane; c < P; c += 32) out[c] = (dn[c] - xn[c] * d) * inv;
}

__global__ void __launch_bounds__(256)
clip_dscale_kernel(const float* __restrict__ Z, const float* __restrict__ lse_r, const float* __restrict__ lse_c,
                   float* __restrict__ out, int N, int row0, int nloc) {
  __shared__ float part[256];
  float s = 0.f;
  const int64_t total = (int64_t)nloc * N;
  for (int64_t idx = threadIdx.x; idx < total; idx += 256) {
    const int a = row0 + (int)(idx / N);
    const int b = (int)(idx % N);
    // row term of local row a against every column b; column term of local column a against every row b
    const float zr = Z[(int64_t)a * N + b];
    const float zc = Z[(int64_t)b * N + a];
    float gr = __expf(zr - lse_r[a]);
    float gc = __expf(zc - lse_c[a]);
    if (a == b) {
      gr -= 1.f;
      gc -= 1.f;
    }
    s += gr * zr + gc * zc;
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = part[0] / (2.f * (float)N);
}

// ---------------------------------------------------------------- class-prompt head
// one warp per image row: logits over C*group prompts, group max, softmax, CE, dlogits
__global__ void __launch_bounds__(256)
class_head_rows_kernel(const float* __restrict__ f_img, const float* __restrict__ f_txt, float scale,
                       const int64_t* __restrict__ labels, const float* __restrict__ soft, float* __restrict__ logits,
                       float* __restrict__ probs, float* __restrict__ rowloss, float* __restrict__ dlogits, int B,
                       int C, int P, int group) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const float* fi = f_img + (int64_t)row * P;
  float* lr = logits + (int64_t)row * C;
  // logits (each class handled by the whole warp so reads of f_txt rows are coalesced)
  for (int c = 0; c < C; ++c) {
    float best = -INFINITY;
    for (int g = 0; g < group; ++g) {
      const float* ft = f_txt + ((int64_t)c * group + g) * P;
      float d = 0.f;
      for (int k = lane; k < P; k += 32) d = fmaf(fi[k], ft[k], d);
      d = warp_sum(d) * scale;
      best = fmaxf(best, d);
    }
    if (lane == 0) lr[c] = best;
  }
  __syncwarp();
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, lr[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(lr[c] - m);
  s = warp_sum(s);
  const float lse = m + logf(s);
  float tsum = 0.f, tdot = 0.f;  // soft labels: loss = sum_c t_c (lse - z_c)
  if (soft != nullptr) {
    for (int c = lane; c < C; c += 32) {
      const float t = soft[(int64_t)row * C + c];
      tsum += t;
      tdot = fmaf(t, lr[c], tdot);
    }
    tsum = warp_sum(tsum);
    tdot = warp_sum(tdot);
  }
  int64_t lab = -1;
  if (labels != nullptr) lab = labels[row];
  for (int c = lane; c < C; c += 32) {
    const float pr = __expf(lr[c] - lse);
    if (probs != nullptr) probs[(int64_t)row * C + c] = pr;
    if (dlogits != nullptr) {
      float g = 0.f;
      if (soft != nullptr)
        g = pr * tsum - soft[(int64_t)row * C + c];
      else if (labels != nullptr)
        g = pr - (c == lab ? 1.f : 0.f);
      dlogits[(int64_t)row * C + c] = g / (float)B;
    }
  }
  if (rowloss != nullptr && lane == 0) {
    float l = 0.f;
    if (soft != nullptr)
      l = tsum * lse - tdot;
    else if (labels != nullptr && lab >= 0 && lab < C)
      l = lse - lr[lab];
    rowloss[row] = l;
  }
}

// d_img[b][p] = scale sum_c dl[b][c] f_txt[c][p];  d_txt[c][p] = scale sum_b dl[b][c] f_img[b][p];  loss = mean
__global__ void __launch_bounds__(256)
class_head_grad_kernel(const float* __restrict__ f_img, const float* __restrict__ f_txt,
                       const float* __restrict__ dlogits, const float* __restrict__ rowloss, float scale,
                       float* __restrict__ d_img, float* __restrict__ d_txt, float* __restrict__ loss, int B, int C,
                       int P) {
  const int64_t nimg = (int64_t)B * P, ntxt = (int64_t)C * P;
  const int64_t nwork = (d_img != nullptr || d_txt != nullptr) ? nimg + ntxt : 0;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < nwork;
       idx += (int64_t)gridDim.x * blockDim.x) {
    if (idx < nimg) {
      if (d_img == nullptr) continue;
      const int b = (int)(idx / P), p = (int)(idx % P);
      float s = 0.f;
      for (int c = 0; c < C; ++c) s = fmaf(dlogits[(int64_t)b * C + c], f_txt[(int64_t)c * P + p], s);
      d_img[idx] = scale * s;
    } else {
      if (d_txt == nullptr) continue;
      const int64_t k = idx - nimg;
      const int c = (int)(k / P), p = (int)(k % P);
      float s = 0.f;
      for (int b = 0; b < B; ++b) s = fmaf(dlogits[(int64_t)b * C + c], f_img[(int64_t)b * P + p], s);
      d_txt[k] = scale * s;
    }
  }
  if (loss != nullptr && blockIdx.x == 0) {
    __shared__ float part[256];
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += 256) s += rowloss[b];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = part[0] / (float)B;
  }
}

}  // namespace
}  // namespace vlmclip

using namespace vlmclip;

extern "C" int vlmclip_l2norm_rows(const float* x, const float* x2, float* y, float* sum_out, int R, int P,
                                   void* stream) {
  VLMCLIP_CHECK_ARG(x && y && R > 0 && P > 0, "l2norm_rows: bad arguments");
  count_launch(1);
  l2norm_rows_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, x2, y, nullptr, sum_out, R, P);
  return report_cuda(cudaGetLastError(), "l2norm_rows_kernel launch");
}

extern "C" int vlmclip_l2norm_rows_bwd(const float* x, const float* dy, float* dx, int R, int P, void* stream) {
  VLMCLIP_CHECK_ARG(x && dy && dx && R > 0 && P > 0, "l2norm_rows_bwd: bad arguments");
  count_launch(1);
  l2norm_rows_bwd_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, dy, dx, R, P);
  return report_cuda(cudaGetLastError(), "l2norm_rows_bwd_kernel launch");
}

extern "C" int64_t vlmclip_clip_loss_workspace(int N, int P) {
  // inv_t[N] inv_i[N] lse_r[N] lse_c[N] | Z[N*N] (used when logits_per_text is NULL) | dtn[N*P] din[N*P]
  return 4 * (int64_t)N + (int64_t)N * N + 2 * (int64_t)N * P;
}

extern "C" int vlmclip_clip_loss(const float* txt, const float* img, float logit_scale_exp, float* txt_n,
                                 float* img_n, float* logits_per_text, float* loss, float* d_txt, float* d_img,
                                 float* d_logit_scale, float* workspace, int N, int P, int row0, int nloc,
                                 void* stream) {
  VLMCLIP_CHECK_ARG(txt && img && txt_n && img_n && loss && workspace, "clip_loss: null pointer");
  VLMCLIP_CHECK_ARG(N > 0 && P > 0, "clip_loss: bad dims N=%d P=%d", N, P);
  VLMCLIP_CHECK_ARG(row0 >= 0 && nloc >= 0 && row0 + nloc <= N, "clip_loss: local rows [%d,%d) outside [0,%d)", row0,
                    row0 + nloc, N);
  VLMCLIP_CHECK_ARG((d_txt == nullptr) == (d_img == nullptr), "clip_loss: d_txt and d_img go together");
  cudaStream_t s = (cudaStream_t)stream;
  float* inv_t = workspace;
  float* inv_i = inv_t + N;
  float* lse_r = inv_i + N;
  float* lse_c = lse_r + N;
  float* Z = logits_per_text != nullptr ? logits_per_text : lse_c + N;
  float* dtn = lse_c + N + (int64_t)N * N;
  float* din = dtn + (int64_t)N * P;
  const int tiles_n = (N + TF_TILE - 1) / TF_TILE;
  count_launch(6);
  l2norm_rows_kernel<<<(N + 7) / 8, 256, 0, s>>>(txt, nullptr, txt_n, inv_t, nullptr, N, P);
  l2norm_rows_kernel<<<(N + 7) / 8, 256, 0, s>>>(img, nullptr, img_n, inv_i, nullptr, N, P);
  VLMCLIP_CUDA(cudaGetLastError());
  sim_logits_kernel<<<dim3(tiles_n, tiles_n), TF_THREADS, 0, s>>>(txt_n, img_n, Z, logit_scale_exp, N, P);
  row_lse_kernel<<<(N + 7) / 8, 256, 0, s>>>(Z, lse_r, N);
  col_lse_kernel<<<(N + 31) / 32, 256, 0, s>>>(Z, lse_c, N);
  clip_loss_reduce_kernel<<<1, 256, 0, s>>>(Z, lse_r, lse_c, loss, N);
  VLMCLIP_CUDA(cudaGetLastError());
  if (d_txt != nullptr && nloc > 0) {
    count_launch(2);
    dim3 g((nloc + TF_TILE - 1) / TF_TILE, (P + TF_TILE - 1) / TF_TILE, 2);
    clip_grad_kernel<<<g, TF_THREADS, 0, s>>>(Z, lse_r, lse_c, txt_n, img_n, dtn, din, logit_scale_exp, N, P, row0, nloc);
    clip_norm_bwd_kernel<<<dim3((nloc + 7) / 8, 2), 256, 0, s>>>(txt_n, img_n, inv_t, inv_i, dtn, din, d_txt, d_img,
                                                                P, row0, nloc);
    VLMCLIP_CUDA(cudaGetLastError());
  }
  if (d_logit_scale != nullptr && nloc > 0) {
    count_launch(1);
    clip_dscale_kernel<<<1, 256, 0, s>>>(Z, lse_r, lse_c, d_logit_scale, N, row0, nloc);
  }
  return report_cuda(cudaGetLastError(), "clip_loss launch");
}

extern "C" int vlmclip_class_head(const float* f_img, const float* f_txt, float scale, const int64_t* labels,
                                  const float* soft_labels, float* logits, float* probs, float* loss, float* d_img,
                                  float* d_txt, float* workspace, int B, int C, int P, int group, void* stream) {
  VLMCLIP_CHECK_ARG(f_img && f_txt && logits, "class_head: null f_img/f_txt/logits");
  VLMCLIP_CHECK_ARG(B > 0 && C > 0 && P > 0 && group >= 1, "class_head: bad dims");
  const bool want_grad = d_img != nullptr || d_txt != nullptr;
  VLMCLIP_CHECK_ARG(!(want_grad && group > 1), "class_head: gradients are not defined for group max (forward only)");
  VLMCLIP_CHECK_ARG(!(want_grad || loss) || (labels || soft_labels), "class_head: loss/gradients need labels");
  VLMCLIP_CHECK_ARG(!(want_grad || loss) || workspace, "class_head: loss/gradients need a workspace of B*C + B floats");
  cudaStream_t s = (cudaStream_t)stream;
  float* dl = workspace;
  float* rowloss = workspace ? workspace + (int64_t)B * C : nullptr;
  count_launch(1);
  class_head_rows_kernel<<<(B + 7) / 8, 256, 0, s>>>(f_img, f_txt, scale, labels, soft_labels, logits, probs,
                                                     (loss || want_grad) ? rowloss : nullptr,
                                                     want_grad ? dl : nullptr, B, C, P, group);
  VLMCLIP_CUDA(cudaGetLastError());
  if (want_grad || loss) {
    count_launch(1);
    const int64_t total = (int64_t)(B + C) * P;
    int grid = (int)((total + 255) / 256);
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    if (!want_grad) grid = 1;
    class_head_grad_kernel<<<grid, 256, 0, s>>>(f_img, f_txt, dl, rowloss, scale, want_grad ? d_img : nullptr,
                                                want_grad ? d_txt : nullptr, loss, B, C, P);
  }
  return report_cuda(cudaGetLastError(), "class_head launch");
}

extern "C" int vlmclip_class_head_bwd(const float* f_img, const float* f_txt, const float* dlogits, float scale,
                                      float* d_img, float* d_txt, int B, int C, int P, void* stream) {
  VLMCLIP_CHECK_ARG(f_img && f_txt && dlogits && (d_img || d_txt), "class_head_bwd: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && C > 0 && P > 0, "class_head_bwd: bad dims");
  const int64_t total = (int64_t)(B + C) * P;
  int grid = (int)((total + 255) / 256);
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  count_launch(1);
  class_head_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(f_img, f_txt, dlogits, nullptr, scale, d_img, d_txt,
                                                                 nullptr, B, C, P);
  return report_cuda(cudaGetLastError(), "class_head_bwd launch");
}

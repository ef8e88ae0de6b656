// fp32 row kernels of the TRAINABLE SharedMHSAttentionAdapter path (adapter/clip_adapter.py:99-128 + autograd; Track M
// evaluates it on token 0 of every caption against the projected vision position table, model_m.py:93-102):
//   LayerNorm forward / backward on fp32 rows, exact-erf GELU forward / backward, masked multiply-add (dropout and
//   residual adds), single-query multi-head attention over a table shared by the batch, forward and backward.
// Everything here works on B (batch) or S (table) rows of 512 floats: latency-bound like the rest of the trainable
// path, fp32 on purpose (gradient parity against autograd), deterministic (no atomics).
#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int SA_MAX_S = 512;

// one warp per row; stats[row] = (mean, rstd)
__global__ void __launch_bounds__(256)
ln_f32_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ stats, int M, int D,
                  float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (int64_t)row * ldx;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += xr[c];
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float d = xr[c] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  for (int c = lane; c < D; c += 32) y[(int64_t)row * D + c] = fmaf((xr[c] - mean) * rstd, gamma[c], beta[c]);
  if (lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ dres), g = gamma * dy
__global__ void __launch_bounds__(256)
ln_f32_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, int64_t ldx,
                  const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ dres,
                  float* __restrict__ dx, int M, int D) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  const float* xr = x + (int64_t)row * ldx;
  const float* dyr = dy + (int64_t)row * D;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float g = gamma[c] * dyr[c];
    s1 += g;
    s2 = fmaf(g, (xr[c] - mean) * rstd, s2);
  }
  const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
  for (int c = lane; c < D; c += 32) {
    const float xh = (xr[c] - mean) * rstd;
    float o = rstd * (gamma[c] * dyr[c] - m1 - xh * m2);
    if (dres != nullptr) o += dres[(int64_t)row * D + c];
    dx[(int64_t)row * D + c] = o;
  }
}

// dgamma[c] = sum_r dy[r,c] * xhat[r,c], dbeta[c] = sum_r dy[r,c]   (rows = batch or table length: a short loop)
__global__ void __launch_bounds__(128)
ln_f32_param_grads_kernel(const float* __restrict__ dy, const float* __restrict__ x, int64_t ldx,
                          const float* __restrict__ stats, float* __restrict__ dgamma, float* __restrict__ dbeta, int M,
                          int D) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float g = 0.f, b = 0.f;
  for (int r = 0; r < M; ++r) {
    const float d = dy[(int64_t)r * D + c];
    g = fmaf(d, (x[(int64_t)r * ldx + c] - stats[2 * r]) * stats[2 * r + 1], g);
    b += d;
  }
  dgamma[c] = g;
  dbeta[c] = b;
}

__global__ void __launch_bounds__(256)
gelu_f32_fwd_kernel(const float* __restrict__ a, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = gelu_erf(a[i]);
}
__global__ void __launch_bounds__(256)
gelu_f32_bwd_kernel(const float* __restrict__ a, const float* __restrict__ dy, float* __restrict__ da, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    da[i] = dy[i] * gelu_erf_grad(a[i]);
}
// y = a (* mask) (+ b): dropout application and residual adds
__global__ void __launch_bounds__(256)
fma_mask_f32_kernel(const float* __restrict__ a, const float* __restrict__ mask, const float* __restrict__ b,
                    float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = a[i];
    if (mask != nullptr) v *= mask[i];
    if (b != nullptr) v += b[i];
    y[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Single-query attention against a table shared by the batch (head_dim 64): one CTA of 64 threads per (b, h).
//   p = softmax(q k^T scale) over the S table rows, pd = p * pmask (attention dropout, nn.MultiheadAttention
//   applies it to the probabilities), out = pd V.   p is kept for the backward.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
attn1q_f32_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t ldkv,
                      const float* __restrict__ pmask, float* __restrict__ p_out, float* __restrict__ out, int S, int H,
                      float scale) {
  __shared__ float sq[64];
  __shared__ float sp[SA_MAX_S];
  __shared__ float red[2];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int t = threadIdx.x;
  const int Dm = H * 64;
  sq[t] = q[(int64_t)b * Dm + h * 64 + t] * scale;
  __syncthreads();
  float mx = -INFINITY;
  for (int j = t; j < S; j += 64) {
    const float* kr = k + (int64_t)j * ldkv + h * 64;
    float s = 0.f;
#pragma unroll 16
    for (int d = 0; d < 64; ++d) s = fmaf(sq[d], kr[d], s);
    sp[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if ((t & 31) == 0) red[t >> 5] = mx;
  __syncthreads();
  mx = fmaxf(red[0], red[1]);
  __syncthreads();
  float l = 0.f;
  for (int j = t; j < S; j += 64) {
    const float e = __expf(sp[j] - mx);
    sp[j] = e;
    l += e;
  }
  l = warp_sum(l);
  if ((t & 31) == 0) red[t >> 5] = l;
  __syncthreads();
  const float inv = 1.f / (red[0] + red[1]);
  const int64_t prow = (int64_t)blockIdx.x * S;
  for (int j = t; j < S; j += 64) {
    const float p = sp[j] * inv;
    p_out[prow + j] = p;
    sp[j] = pmask != nullptr ? p * pmask[prow + j] : p;
  }
  __syncthreads();
  float o = 0.f;
  for (int j = 0; j < S; ++j) o = fmaf(sp[j], v[(int64_t)j * ldkv + h * 64 + t], o);
  out[(int64_t)b * Dm + h * 64 + t] = o;
}

// per (b, h): dpd_j = dout . v_j; dp = dpd * pmask; ds_j = p_j (dp_j - sum_i p_i dp_i); dq = scale * sum_j ds_j k_j.
// ds is written for the table-side kernel below.
__global__ void __launch_bounds__(64)
attn1q_f32_bwd_q_kernel(const float* __restrict__ dout, const float* __restrict__ k, const float* __restrict__ v,
                        int64_t ldkv, const float* __restrict__ p, const float* __restrict__ pmask,
                        float* __restrict__ ds_out, float* __restrict__ dq, int S, int H, float scale) {
  __shared__ float sg[64];
  __shared__ float sds[SA_MAX_S];
  __shared__ float red[2];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int t = threadIdx.x;
  const int Dm = H * 64;
  const int64_t prow = (int64_t)blockIdx.x * S;
  sg[t] = dout[(int64_t)b * Dm + h * 64 + t];
  __syncthreads();
  float acc = 0.f;
  for (int j = t; j < S; j += 64) {
    const float* vr = v + (int64_t)j * ldkv + h * 64;
    float dp = 0.f;
#pragma unroll 16
    for (int d = 0; d < 64; ++d) dp = fmaf(sg[d], vr[d], dp);
    if (pmask != nullptr) dp *= pmask[prow + j];
    sds[j] = dp;
    acc = fmaf(p[prow + j], dp, acc);
  }
  acc = warp_sum(acc);
  if ((t & 31) == 0) red[t >> 5] = acc;
  __syncthreads();
  const float dot = red[0] + red[1];
  for (int j = t; j < S; j += 64) {
    const float ds = p[prow + j] * (sds[j] - dot);
    sds[j] = ds;
    ds_out[prow + j] = ds;
  }
  __syncthreads();
  float o = 0.f;
  for (int j = 0; j < S; ++j) o = fmaf(sds[j], k[(int64_t)j * ldkv + h * 64 + t], o);
  dq[(int64_t)b * Dm + h * 64 + t] = o * scale;
}

// per (table row j, head h): dk_j = scale * sum_b ds[b,h,j] q[b,h,:], dv_j = sum_b p[b,h,j] pmask[b,h,j] dout[b,h,:]
__global__ void __launch_bounds__(64)
attn1q_f32_bwd_kv_kernel(const float* __restrict__ q, const float* __restrict__ dout, const float* __restrict__ p,
                         const float* __restrict__ pmask, const float* __restrict__ ds, float* __restrict__ dk,
                         float* __restrict__ dv, int B, int S, int H, float scale) {
  const int j = blockIdx.x / H, h = blockIdx.x % H;
  const int t = threadIdx.x;
  const int Dm = H * 64;
  float ak = 0.f, av = 0.f;
  for (int b = 0; b < B; ++b) {
    const int64_t pi = ((int64_t)b * H + h) * S + j;
    float pd = p[pi];
    if (pmask != nullptr) pd *= pmask[pi];
    ak = fmaf(ds[pi], q[(int64_t)b * Dm + h * 64 + t], ak);
    av = fmaf(pd, dout[(int64_t)b * Dm + h * 64 + t], av);
  }
  dk[(int64_t)j * Dm + h * 64 + t] = ak * scale;
  dv[(int64_t)j * Dm + h * 64 + t] = av;
}

int ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace
}  // namespace vlmclip

using namespace vlmclip;

extern "C" int vlmclip_layernorm_f32(const float* x, int64_t ldx, const float* gamma, const float* beta, float* y,
                                     float* stats, int M, int D, float eps, void* stream) {
  VLMCLIP_CHECK_ARG(x && gamma && beta && y && stats && M > 0 && D > 0 && ldx >= D, "layernorm_f32: bad arguments");
  count_launch(1);
  ln_f32_fwd_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, ldx, gamma, beta, y, stats, M, D, eps);
  return report_cuda(cudaGetLastError(), "ln_f32_fwd_kernel launch");
}

extern "C" int vlmclip_layernorm_f32_bwd(const float* dy, const float* x, int64_t ldx, const float* stats,
                                         const float* gamma, const float* dres, float* dx, float* dgamma, float* dbeta,
                                         int M, int D, void* stream) {
  VLMCLIP_CHECK_ARG(dy && x && stats && gamma && M > 0 && D > 0 && ldx >= D, "layernorm_f32_bwd: bad arguments");
  VLMCLIP_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "layernorm_f32_bwd: dgamma and dbeta go together");
  cudaStream_t s = (cudaStream_t)stream;
  if (dx != nullptr) {
    count_launch(1);
    ln_f32_bwd_kernel<<<(M + 7) / 8, 256, 0, s>>>(dy, x, ldx, stats, gamma, dres, dx, M, D);
  }
  if (dgamma != nullptr) {
    count_launch(1);
    ln_f32_param_grads_kernel<<<(D + 127) / 128, 128, 0, s>>>(dy, x, ldx, stats, dgamma, dbeta, M, D);
  }
  return report_cuda(cudaGetLastError(), "layernorm_f32_bwd launch");
}

extern "C" int vlmclip_gelu_f32(const float* a, float* y, int64_t n, void* stream) {
  VLMCLIP_CHECK_ARG(a && y && n > 0, "gelu_f32: bad arguments");
  count_launch(1);
  gelu_f32_fwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(a, y, n);
  return report_cuda(cudaGetLastError(), "gelu_f32_fwd_kernel launch");
}

extern "C" int vlmclip_gelu_f32_bwd(const float* a, const float* dy, float* da, int64_t n, void* stream) {
  VLMCLIP_CHECK_ARG(a && dy && da && n > 0, "gelu_f32_bwd: bad arguments");
  count_launch(1);
  gelu_f32_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(a, dy, da, n);
  return report_cuda(cudaGetLastError(), "gelu_f32_bwd_kernel launch");
}

extern "C" int vlmclip_fma_mask_f32(const float* a, const float* mask, const float* b, float* y, int64_t n,
                                    void* stream) {
  VLMCLIP_CHECK_ARG(a && y && n > 0, "fma_mask_f32: bad arguments");
  count_launch(1);
  fma_mask_f32_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(a, mask, b, y, n);
  return report_cuda(cudaGetLastError(), "fma_mask_f32_kernel launch");
}

extern "C" int vlmclip_attn1q_f32_fwd(const float* q, const float* k, const float* v, int64_t ldkv, const float* pmask,
                                      float* p_out, float* out, int B, int S, int H, float scale, void* stream) {
  VLMCLIP_CHECK_ARG(q && k && v && p_out && out, "attn1q_f32_fwd: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && H > 0 && S > 0 && S <= SA_MAX_S && ldkv >= H * 64, "attn1q_f32_fwd: bad dims B=%d S=%d H=%d", B,
                    S, H);
  count_launch(1);
  attn1q_f32_fwd_kernel<<<B * H, 64, 0, (cudaStream_t)stream>>>(q, k, v, ldkv, pmask, p_out, out, S, H, scale);
  return report_cuda(cudaGetLastError(), "attn1q_f32_fwd_kernel launch");
}

extern "C" int vlmclip_attn1q_f32_bwd(const float* dout, const float* q, const float* k, const float* v, int64_t ldkv,
                                      const float* p, const float* pmask, float* ds_ws, float* dq, float* dk, float* dv,
                                      int B, int S, int H, float scale, void* stream) {
  VLMCLIP_CHECK_ARG(dout && q && k && v && p && ds_ws && dq && dk && dv, "attn1q_f32_bwd: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && H > 0 && S > 0 && S <= SA_MAX_S && ldkv >= H * 64, "attn1q_f32_bwd: bad dims B=%d S=%d H=%d", B,
                    S, H);
  cudaStream_t s = (cudaStream_t)stream;
  count_launch(2);
  attn1q_f32_bwd_q_kernel<<<B * H, 64, 0, s>>>(dout, k, v, ldkv, p, pmask, ds_ws, dq, S, H, scale);
  attn1q_f32_bwd_kv_kernel<<<S * H, 64, 0, s>>>(q, dout, p, pmask, ds_ws, dk, dv, B, S, H, scale);
  return report_cuda(cudaGetLastError(), "attn1q_f32_bwd launch");
}

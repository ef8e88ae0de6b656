// Optimiser tail of the adapter train step (trainer.py:91-99): clip_grad_norm_(params, max_norm) followed by
// torch.optim.AdamW.step(), fused over one flat fp32 arena (params, grads, exp_avg, exp_avg_sq share offsets).
// Two launches, no host synchronisation, deterministic: (1) per-block partial sums of squares, (2) every
// block re-reduces the partials in the same order, derives the clip coefficient and applies the update.
// Track T/V's optim.Adam(lr) (model_t.py:141-145) is the same kernel with weight_decay = 0, max_norm <= 0.
#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);
namespace {

constexpr int OPT_BLOCK = 256;
constexpr int OPT_MAX_PARTIALS = 1024;

__global__ void __launch_bounds__(OPT_BLOCK)
sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partials, int32_t* step) {
  __shared__ float part[OPT_BLOCK];
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)OPT_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * OPT_BLOCK)
    s = fmaf(g[i], g[i], s);
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = OPT_BLOCK / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = part[0];
    if (blockIdx.x == 0) step[0] += 1;
  }
}

__global__ void __launch_bounds__(OPT_BLOCK)
adamw_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  int64_t n, const float* __restrict__ lr_dev, float beta1, float beta2, float eps, float wd,
                  float max_norm, const int32_t* __restrict__ step, const float* __restrict__ partials, int npart,
                  float* __restrict__ grad_norm_out) {
  __shared__ float s_coef;
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < npart; ++i) tot += partials[i];
    const float norm = sqrtf(tot);
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(max_norm / (norm + 1e-6f), 1.f);
    s_coef = coef;
    if (blockIdx.x == 0 && grad_norm_out != nullptr) grad_norm_out[0] = norm;
  }
  __syncthreads();
  const float coef = s_coef;
  const float lr = lr_dev[0];
  const float t = (float)step[0];
  const float bc1 = 1.f - powf(beta1, t);
  const float bc2 = 1.f - powf(beta2, t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = 1.f - lr * wd;
  for (int64_t i = blockIdx.x * (int64_t)OPT_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * OPT_BLOCK) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = p[i] * decay - step_size * (mi / denom);
  }
}

}  // namespace
}  // namespace vlmclip

using namespace vlmclip;

extern "C" int vlmclip_adamw_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                       int64_t n, const float* lr_dev, float beta1, float beta2, float eps,
                                       float weight_decay, float max_norm, int32_t* step, float* grad_norm_out,
                                       float* workspace, void* stream) {
  VLMCLIP_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && lr_dev && step && workspace,
                    "adamw_clip_step: null pointer");
  VLMCLIP_CHECK_ARG(n > 0, "adamw_clip_step: empty parameter arena");
  int64_t blocks = (n + OPT_BLOCK * 4 - 1) / (OPT_BLOCK * 4);
  if (blocks > OPT_MAX_PARTIALS) blocks = OPT_MAX_PARTIALS;
  if (blocks < 1) blocks = 1;
  cudaStream_t s = (cudaStream_t)stream;
  count_launch(2);
  sumsq_partial_kernel<<<(int)blocks, OPT_BLOCK, 0, s>>>(grads, n, workspace, step);
  VLMCLIP_CUDA(cudaGetLastError());
  adamw_clip_kernel<<<(int)blocks, OPT_BLOCK, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n, lr_dev, beta1, beta2,
                                                      eps, weight_decay, max_norm, step, workspace, (int)blocks,
                                                      grad_norm_out);
  return report_cuda(cudaGetLastError(), "adamw_clip_step launch");
}

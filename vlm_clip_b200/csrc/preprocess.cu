// On-GPU frame preprocessing fused with patch extraction, and the temporal mean-pool of clip features.
//   process_video.py:14-29   BGR2RGB -> cv2.resize(frame, size) (INTER_LINEAR, uint8) -> ToTensor (/255) -> Normalize
//   HF modeling_clip.py:209  Conv2d(kernel = stride = patch) == im2col + GEMM
//   SURVEY.md 8a-12          clip feature = get_image_features(frames).view(B, T, P).mean(1)
// The decoded uint8 HWC frames go straight to the bf16 im2col matrix the patch GEMM reads: the fp32 NCHW pixel tensor
// (154 MB per 256 frames) never exists.  The bilinear resize reproduces OpenCV's 11-bit fixed-point arithmetic
// (coefficient tables are built by the caller, see ops._resize_tables), so resized pixels match cv2 bit for bit
// wherever oracle/preprocess_oracle.py does.
#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

struct PreArgs {
  const uint8_t* src;
  int64_t frame_stride;  // bytes between frames
  int Hs, Ws, bgr;
  const int32_t* ytab;  // [H][3]: source row, weight of it, weight of the next row (11-bit fixed point); null = no resize
  const int32_t* xtab;  // [W][3]
  float mean[3], stdv[3];
  __nv_bfloat16* out;
  int n, H, W, p, Kpad;
};

__global__ void __launch_bounds__(256) preprocess_patches_kernel(const PreArgs a) {
  const int64_t total = (int64_t)a.n * a.H * a.W;
  const int gw = a.W / a.p, gh = a.H / a.p;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % a.W);
    const int64_t t = idx / a.W;
    const int y = (int)(t % a.H);
    const int n = (int)(t / a.H);
    const uint8_t* f = a.src + n * a.frame_stride;
    int v[3];
    if (a.ytab != nullptr) {
      const int sy = a.ytab[y * 3], b0 = a.ytab[y * 3 + 1], b1 = a.ytab[y * 3 + 2];
      const int sx = a.xtab[x * 3], a0 = a.xtab[x * 3 + 1], a1 = a.xtab[x * 3 + 2];
      const int sy1 = min(sy + 1, a.Hs - 1), sx1 = min(sx + 1, a.Ws - 1);
      const uint8_t* r0 = f + (int64_t)sy * a.Ws * 3;
      const uint8_t* r1 = f + (int64_t)sy1 * a.Ws * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int s0 = r0[sx * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;  // horizontal pass
        const int s1 = r1[sx * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
        v[c] = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;  // OpenCV VResizeLinear, 8U
        v[c] = min(max(v[c], 0), 255);
      }
    } else {
      const uint8_t* px = f + ((int64_t)y * a.Ws + x) * 3;
      v[0] = px[0];
      v[1] = px[1];
      v[2] = px[2];
    }
    const int py = y / a.p, i = y - py * a.p;
    const int pxx = x / a.p, j = x - pxx * a.p;
    __nv_bfloat16* dst = a.out + (((int64_t)n * gh + py) * gw + pxx) * a.Kpad + i * a.p + j;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int sc = a.bgr ? 2 - c : c;
      const float q = __fdiv_rn((float)v[sc], 255.f);                   // ToTensor
      const float z = __fdiv_rn(q - a.mean[c], a.stdv[c]);              // Normalize
      dst[c * a.p * a.p] = __float2bfloat16(z);
    }
  }
}

__global__ void __launch_bounds__(256) pad_cols_kernel(__nv_bfloat16* __restrict__ out, int64_t rows, int K, int Kpad) {
  const int padw = Kpad - K;
  const int64_t total = rows * padw;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / padw;
    out[r * Kpad + K + (idx - r * padw)] = __float2bfloat16(0.f);
  }
}

// y[b][p] = (1/T) sum_t x[b*T + t][p]   (fixed order: deterministic)
__global__ void __launch_bounds__(256) mean_pool_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int T, int P) {
  const int64_t total = (int64_t)B * P;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx % P);
    const int64_t b = idx / P;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += x[(b * T + t) * P + p];
    y[idx] = s / (float)T;
  }
}
__global__ void __launch_bounds__(256) mean_pool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int T, int P) {
  const int64_t total = (int64_t)B * T * P;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx % P);
    const int64_t b = idx / ((int64_t)T * P);
    dx[idx] = dy[b * P + p] / (float)T;
  }
}

inline int grid_of(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace
}  // namespace vlmclip

using namespace vlmclip;

extern "C" int vlmclip_preprocess_patches(const uint8_t* frames, int64_t frame_stride, int Hs, int Ws, int bgr,
                                          const int32_t* ytab, const int32_t* xtab, float mean0, float mean1, float mean2,
                                          float std0, float std1, float std2, void* out, int n_frames, int H, int W,
                                          int patch, void* stream) {
  VLMCLIP_CHECK_ARG(frames && out, "preprocess: null pointer");
  VLMCLIP_CHECK_ARG(n_frames > 0 && Hs > 0 && Ws > 0 && patch > 0 && H % patch == 0 && W % patch == 0,
                    "preprocess: bad dims (H=%d W=%d patch=%d)", H, W, patch);
  VLMCLIP_CHECK_ARG(frame_stride >= (int64_t)Hs * Ws * 3, "preprocess: frame_stride smaller than a frame");
  VLMCLIP_CHECK_ARG((ytab == nullptr) == (xtab == nullptr), "preprocess: give both resize tables or none");
  VLMCLIP_CHECK_ARG(ytab != nullptr || (Hs == H && Ws == W), "preprocess: %dx%d frames need resize tables for %dx%d", Hs, Ws,
                    H, W);
  VLMCLIP_CHECK_ARG(std0 != 0.f && std1 != 0.f && std2 != 0.f, "preprocess: zero std");
  PreArgs a;
  a.src = frames;
  a.frame_stride = frame_stride;
  a.Hs = Hs;
  a.Ws = Ws;
  a.bgr = bgr;
  a.ytab = ytab;
  a.xtab = xtab;
  a.mean[0] = mean0; a.mean[1] = mean1; a.mean[2] = mean2;
  a.stdv[0] = std0; a.stdv[1] = std1; a.stdv[2] = std2;
  a.out = (__nv_bfloat16*)out;
  a.n = n_frames;
  a.H = H;
  a.W = W;
  a.p = patch;
  const int K = 3 * patch * patch;
  a.Kpad = (K + 63) / 64 * 64;
  count_launch(1);
  preprocess_patches_kernel<<<grid_of((int64_t)n_frames * H * W), 256, 0, (cudaStream_t)stream>>>(a);
  if (a.Kpad != K) {
    const int64_t rows = (int64_t)n_frames * (H / patch) * (W / patch);
    count_launch(1);
    pad_cols_kernel<<<grid_of(rows * (a.Kpad - K)), 256, 0, (cudaStream_t)stream>>>(a.out, rows, K, a.Kpad);
  }
  return report_cuda(cudaGetLastError(), "preprocess_patches_kernel launch");
}

extern "C" int vlmclip_mean_pool(const float* x, float* y, int B, int T, int P, void* stream) {
  VLMCLIP_CHECK_ARG(x && y && B > 0 && T > 0 && P > 0, "mean_pool: bad arguments");
  count_launch(1);
  mean_pool_kernel<<<grid_of((int64_t)B * P), 256, 0, (cudaStream_t)stream>>>(x, y, B, T, P);
  return report_cuda(cudaGetLastError(), "mean_pool_kernel launch");
}

extern "C" int vlmclip_mean_pool_bwd(const float* dy, float* dx, int B, int T, int P, void* stream) {
  VLMCLIP_CHECK_ARG(dy && dx && B > 0 && T > 0 && P > 0, "mean_pool_bwd: bad arguments");
  count_launch(1);
  mean_pool_bwd_kernel<<<grid_of((int64_t)B * T * P), 256, 0, (cudaStream_t)stream>>>(dy, dx, B, T, P);
  return report_cuda(cudaGetLastError(), "mean_pool_bwd_kernel launch");
}

// Row-wise, HBM-bound kernels of the towers: LayerNorm, row statistics, patch extraction, token
// assembly.  One warp per row, 16-byte vector accesses, fp32 statistics (two-pass in registers).
// References: HF modeling_clip.py:371,380,562,677 (LayerNorm), :202-218 (vision embeddings),
// :234-258 (text embeddings).
#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int ROWS_PER_BLOCK = 8;
constexpr int MAX_VEC = 8;  // 8 x (32 lanes x 8 elements) = D <= 2048 (kernels are instantiated for 2,3,4,8)

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  f[0] = bf16_lo(q.x);
  f[1] = bf16_hi(q.x);
  f[2] = bf16_lo(q.y);
  f[3] = bf16_hi(q.y);
  f[4] = bf16_lo(q.z);
  f[5] = bf16_hi(q.z);
  f[6] = bf16_lo(q.w);
  f[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}
// second term of the two-term (hi + lo) residual stream: bf16(f - hi)
__device__ __forceinline__ uint4 pack8_lo(const float* f, const uint4& hi) {
  float h[8], d[8];
  unpack8(hi, h);
#pragma unroll
  for (int j = 0; j < 8; ++j) d[j] = f[j] - h[j];
  return pack8(d);
}

// mean / rstd of one row held as v[NV][8] per lane
template <int NV>
__device__ __forceinline__ void row_mean_rstd(float (*v)[8], int* nvec_lane_valid, int D, float eps, float& mean,
                                              float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (nvec_lane_valid[i])
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
  s = warp_sum(s);
  mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (nvec_lane_valid[i])
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q = fmaf(d, d, q);
      }
  q = warp_sum(q);
  rstd = rsqrtf(q / (float)D + eps);
}

// MODE 0: y = LN(x) bf16 (+stats)   MODE 1: stats only   MODE 2: y = LN(x) written as fp32
template <int MODE, int NV>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, void* __restrict__ y_, int64_t ldy,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ stats, int M,
                 int D, float eps) {
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  float v[NV][8];
  int valid[NV];
  const __nv_bfloat16* xr = x + (int64_t)row * ldx;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    valid[i] = vi < nvec;
    if (valid[i]) {
      const uint4 q = ld_nc_v4(xr + vi * 8);
      unpack8(q, v[i]);
    }
  }
  float mean, rstd;
  row_mean_rstd<NV>(v, valid, D, eps, mean, rstd);
  if (stats != nullptr && lane == 0) {
    stats[2 * (int64_t)row] = mean;
    stats[2 * (int64_t)row + 1] = rstd;
  }
  if (MODE == 0 || MODE == 2) {
    __nv_bfloat16* yr = reinterpret_cast<__nv_bfloat16*>(y_) + (int64_t)row * ldy;
    float* yf = reinterpret_cast<float*>(y_) + (int64_t)row * ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (valid[i]) {
        const int c = (lane + i * 32) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((v[i][j] - mean) * rstd, g[j], b[j]);
        if (MODE == 0) {
          st_v4(yr + c, pack8(o));
        } else {
          *reinterpret_cast<float4*>(yf + c) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(yf + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }
}

// out[m, k] = pixel[b, c, py*p + i, px*p + j],  m = (b, py, px), k = c*p*p + i*p + j; k >= 3*p*p -> 0
template <typename PixT>
__global__ void __launch_bounds__(256)
im2col_kernel(const PixT* __restrict__ pix, __nv_bfloat16* __restrict__ out, int B, int H, int W, int p, int Kpad) {
  const int gw = W / p, gh = H / p;
  const int kvec = Kpad >> 3;
  const int64_t total = (int64_t)B * gh * gw * kvec;
  const int pp = p * p;
  const int K = 3 * pp;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int kv = (int)(idx % kvec);
    const int64_t m = idx / kvec;
    const int px = (int)(m % gw);
    const int py = (int)((m / gw) % gh);
    const int b = (int)(m / ((int64_t)gw * gh));
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kv * 8 + e;
      if (k < K) {
        const int c = k / pp;
        const int r = k - c * pp;
        const int i = r / p;
        const int j = r - i * p;
        f[e] = (float)pix[(((int64_t)b * 3 + c) * H + (py * p + i)) * W + (px * p + j)];
      } else {
        f[e] = 0.f;
      }
    }
    st_v4(out + m * Kpad + kv * 8, pack8(f));
  }
}

// Fast path for patch % 8 == 0: one thread moves 8 horizontally adjacent pixels.  Threads are ordered along
// image rows, so global reads are fully coalesced; every write is one full 16-byte bf16 vector.
template <typename PixT>
__global__ void __launch_bounds__(256)
im2col_rows_kernel(const PixT* __restrict__ pix, __nv_bfloat16* __restrict__ out, int B, int H, int W, int p,
                   int Kpad) {
  const int w8 = W >> 3;
  const int gw = W / p, gh = H / p;
  const int64_t total = (int64_t)B * 3 * H * w8;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int x8 = (int)(idx % w8);
    int64_t t = idx / w8;
    const int y = (int)(t % H);
    t /= H;
    const int c = (int)(t % 3);
    const int b = (int)(t / 3);
    const int x = x8 * 8;
    const PixT* src = pix + (((int64_t)b * 3 + c) * H + y) * W + x;
    float f[8];
    if (sizeof(PixT) == 4) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(src));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
      f[0] = a0.x; f[1] = a0.y; f[2] = a0.z; f[3] = a0.w;
      f[4] = a1.x; f[5] = a1.y; f[6] = a1.z; f[7] = a1.w;
    } else {
      const uint4 q = ld_nc_v4(src);
      unpack8(q, f);
    }
    const int py = y / p, i = y - py * p;
    const int px = x / p, j = x - px * p;
    const int64_t m = ((int64_t)b * gh + py) * gw + px;
    st_v4(out + m * Kpad + (c * p * p + i * p + j), pack8(f));
  }
}

// zero the padding columns [K, Kpad) (only needed when 3*p*p is not a multiple of 64)
__global__ void __launch_bounds__(256)
im2col_pad_kernel(__nv_bfloat16* __restrict__ out, int64_t rows, int K, int Kpad) {
  const int padw = Kpad - K;
  const int64_t total = rows * padw;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / padw;
    out[r * Kpad + K + (idx - r * padw)] = __float2bfloat16(0.f);
  }
}

// vision tokens: x[b,0] = cls + pos[0]; x[b,1+q] = patch[b,q] + pos[1+q]; y = LN(x)
// patch rows are fp32 (PATCH_BF16 = false) or bf16 (the patch GEMM's staged bf16 output: half the traffic)
template <int NV, bool PATCH_BF16>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
vision_embed_ln_kernel(const void* __restrict__ patch_v, const float* __restrict__ cls,
                       const float* __restrict__ pos, const float* __restrict__ gamma,
                       const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                       __nv_bfloat16* __restrict__ y_lo, int B, int S, int D, float eps) {
  const int64_t row = blockIdx.x * (int64_t)ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= (int64_t)B * S) return;
  const int lane = threadIdx.x & 31;
  const int s = (int)(row % S);
  const int64_t b = row / S;
  const int nvec = D >> 3;
  const int64_t prow = (b * (S - 1) + (s - 1)) * (int64_t)D;
  const float* src = (s == 0 || PATCH_BF16) ? cls : reinterpret_cast<const float*>(patch_v) + prow;
  const __nv_bfloat16* src16 = reinterpret_cast<const __nv_bfloat16*>(patch_v) + prow;
  const float* pr = pos + (int64_t)s * D;
  float v[NV][8];
  int valid[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    valid[i] = vi < nvec;
    if (valid[i]) {
      float4 a0, a1;
      if (PATCH_BF16 && s != 0) {
        const uint4 raw = ld_nc_v4(src16 + vi * 8);
        a0 = make_float4(bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y));
        a1 = make_float4(bf16_lo(raw.z), bf16_hi(raw.z), bf16_lo(raw.w), bf16_hi(raw.w));
      } else {
        a0 = __ldg(reinterpret_cast<const float4*>(src + vi * 8));
        a1 = __ldg(reinterpret_cast<const float4*>(src + vi * 8 + 4));
      }
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(pr + vi * 8));
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(pr + vi * 8 + 4));
      v[i][0] = a0.x + p0.x;
      v[i][1] = a0.y + p0.y;
      v[i][2] = a0.z + p0.z;
      v[i][3] = a0.w + p0.w;
      v[i][4] = a1.x + p1.x;
      v[i][5] = a1.y + p1.y;
      v[i][6] = a1.z + p1.z;
      v[i][7] = a1.w + p1.w;
    }
  }
  float mean, rstd;
  row_mean_rstd<NV>(v, valid, D, eps, mean, rstd);
  __nv_bfloat16* yr = y + row * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (valid[i]) {
      const int c = (lane + i * 32) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf((v[i][j] - mean) * rstd, gv[j], bv[j]);
      const uint4 hi = pack8(o);
      st_v4(yr + c, hi);
      if (y_lo != nullptr) st_v4(y_lo + row * D + c, pack8_lo(o, hi));
    }
  }
}

template <typename TokT>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
text_embed_kernel(const int64_t* __restrict__ ids, const TokT* __restrict__ tok, const float* __restrict__ pos,
                  __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ y_lo, int B, int S, int D, int V) {
  const int64_t row = blockIdx.x * (int64_t)ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= (int64_t)B * S) return;
  const int lane = threadIdx.x & 31;
  const int s = (int)(row % S);
  int64_t id = ids[row];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  const TokT* tr = tok + id * D;
  const float* pr = pos + (int64_t)s * D;
  for (int c = lane * 8; c < D; c += 256) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (float)tr[c + j] + __ldg(pr + c + j);
    const uint4 hi = pack8(o);
    st_v4(y + row * D + c, hi);
    if (y_lo != nullptr) st_v4(y_lo + row * D + c, pack8_lo(o, hi));
  }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ x_lo, int64_t ldx,
                   float* __restrict__ y, int R, int D) {
  const int64_t total = (int64_t)R * D;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / D;
    const int c = (int)(idx - r * D);
    float v = __bfloat162float(x[r * ldx + c]);
    if (x_lo != nullptr) v += __bfloat162float(x_lo[r * ldx + c]);
    y[idx] = v;
  }
}

// instantiate the row kernels for the per-lane vector counts that occur (D = 512, 768, 1024, <= 2048)
#define VLMCLIP_DISPATCH_NV(D, CALL)          \
  do {                                        \
    const int _nv = ((D) + 255) / 256;        \
    if (_nv <= 2) {                           \
      constexpr int NV = 2;                   \
      CALL;                                   \
    } else if (_nv == 3) {                    \
      constexpr int NV = 3;                   \
      CALL;                                   \
    } else if (_nv == 4) {                    \
      constexpr int NV = 4;                   \
      CALL;                                   \
    } else {                                  \
      constexpr int NV = 8;                   \
      CALL;                                   \
    }                                         \
  } while (0)

// (mean, M2) partials of 32-column blocks -> (mean, rstd) per row (Chan's parallel combination); 4 lanes per row
__global__ void __launch_bounds__(256)
ln_partials_to_stats_kernel(const float2* __restrict__ part, float2* __restrict__ stats, int M, int npart, float eps) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * 64 + (threadIdx.x >> 2);
  const int sub = threadIdx.x & 3;
  float msum = 0.f;
  float2 q[16];  // up to 64 partials (D <= 2048) spread over 4 lanes
  int cnt = 0;
  if (row < M) {
    for (int i = sub; i < npart; i += 4) {
      q[cnt] = __ldg(part + (int64_t)row * npart + i);
      msum += q[cnt].x;
      ++cnt;
    }
  }
  msum += __shfl_xor_sync(0xffffffffu, msum, 1);
  msum += __shfl_xor_sync(0xffffffffu, msum, 2);
  const float mean = msum / (float)npart;
  float m2 = 0.f;
  for (int i = 0; i < cnt; ++i) {
    const float d = q[i].x - mean;
    m2 += q[i].y + 32.f * d * d;
  }
  m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
  m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
  if (row < M && sub == 0) stats[row] = make_float2(mean, rsqrtf(m2 / (32.f * (float)npart) + eps));
}

int grid_for(int64_t total, int block) {
  int64_t g = (total + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 32;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace
}  // namespace vlmclip

using namespace vlmclip;

extern "C" int vlmclip_layernorm_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma,
                                      const float* beta, float* stats_out, int M, int D, float eps, void* stream) {
  VLMCLIP_CHECK_ARG(x && y && gamma && beta, "layernorm: null pointer");
  VLMCLIP_CHECK_ARG(M > 0 && D > 0 && D % 8 == 0 && D <= MAX_VEC * 256, "layernorm: D=%d must be a multiple of 8, <= %d",
                    D, MAX_VEC * 256);
  VLMCLIP_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= D && ldy >= D, "layernorm: bad leading dimension");
  VLMCLIP_CHECK_ARG((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)gamma % 16 == 0 &&
                        (uintptr_t)beta % 16 == 0,
                    "layernorm: pointers must be 16-byte aligned");
  const int grid = (M + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
  count_launch(1);
  VLMCLIP_DISPATCH_NV(D, (layernorm_kernel<0, NV><<<grid, ROWS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
                             (const __nv_bfloat16*)x, ldx, y, ldy, gamma, beta, stats_out, M, D, eps)));
  return report_cuda(cudaGetLastError(), "layernorm_kernel launch");
}

extern "C" int vlmclip_layernorm_bf16_f32out(const void* x, int64_t ldx, float* y, int64_t ldy, const float* gamma,
                                             const float* beta, int M, int D, float eps, void* stream) {
  VLMCLIP_CHECK_ARG(x && y && gamma && beta, "layernorm_f32out: null pointer");
  VLMCLIP_CHECK_ARG(M > 0 && D > 0 && D % 8 == 0 && D <= MAX_VEC * 256, "layernorm_f32out: D=%d must be a multiple of 8, <= %d",
                    D, MAX_VEC * 256);
  VLMCLIP_CHECK_ARG(ldx % 8 == 0 && ldy % 4 == 0 && ldx >= D && ldy >= D, "layernorm_f32out: bad leading dimension");
  VLMCLIP_CHECK_ARG((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)gamma % 16 == 0 &&
                        (uintptr_t)beta % 16 == 0,
                    "layernorm_f32out: pointers must be 16-byte aligned");
  const int grid = (M + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
  count_launch(1);
  VLMCLIP_DISPATCH_NV(D, (layernorm_kernel<2, NV><<<grid, ROWS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
                             (const __nv_bfloat16*)x, ldx, y, ldy, gamma, beta, nullptr, M, D, eps)));
  return report_cuda(cudaGetLastError(), "layernorm_kernel<f32out> launch");
}

extern "C" int vlmclip_row_stats_bf16(const void* x, int64_t ldx, float* stats_out, int M, int D, float eps,
                                      void* stream) {
  VLMCLIP_CHECK_ARG(x && stats_out, "row_stats: null pointer");
  VLMCLIP_CHECK_ARG(M > 0 && D > 0 && D % 8 == 0 && D <= MAX_VEC * 256, "row_stats: D=%d must be a multiple of 8, <= %d",
                    D, MAX_VEC * 256);
  VLMCLIP_CHECK_ARG(ldx % 8 == 0 && ldx >= D && (uintptr_t)x % 16 == 0, "row_stats: bad ldx/alignment");
  const int grid = (M + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
  count_launch(1);
  VLMCLIP_DISPATCH_NV(D, (layernorm_kernel<1, NV><<<grid, ROWS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
                             (const __nv_bfloat16*)x, ldx, nullptr, 0, nullptr, nullptr, stats_out, M, D, eps)));
  return report_cuda(cudaGetLastError(), "row_stats kernel launch");
}

extern "C" int vlmclip_im2col_patches(const void* pixels, int pix_bf16, void* out, int B, int H, int W, int patch,
                                      void* stream) {
  VLMCLIP_CHECK_ARG(pixels && out, "im2col: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && patch > 0 && H % patch == 0 && W % patch == 0, "im2col: H=%d W=%d not divisible by patch=%d",
                    H, W, patch);
  const int K = 3 * patch * patch;
  const int Kpad = (K + 63) / 64 * 64;
  VLMCLIP_CHECK_ARG((uintptr_t)out % 16 == 0, "im2col: out must be 16-byte aligned");
  if (patch % 8 == 0 && W % 8 == 0 && (uintptr_t)pixels % 16 == 0) {
    const int64_t rows = (int64_t)B * (H / patch) * (W / patch);
    const int64_t total8 = (int64_t)B * 3 * H * (W / 8);
    count_launch(1);
    if (pix_bf16)
      im2col_rows_kernel<__nv_bfloat16><<<grid_for(total8, 256), 256, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)pixels, (__nv_bfloat16*)out, B, H, W, patch, Kpad);
    else
      im2col_rows_kernel<float><<<grid_for(total8, 256), 256, 0, (cudaStream_t)stream>>>(
          (const float*)pixels, (__nv_bfloat16*)out, B, H, W, patch, Kpad);
    if (Kpad != K) {
      count_launch(1);
      im2col_pad_kernel<<<grid_for(rows * (Kpad - K), 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out, rows,
                                                                                          K, Kpad);
    }
    return report_cuda(cudaGetLastError(), "im2col_rows_kernel launch");
  }
  const int64_t total = (int64_t)B * (H / patch) * (W / patch) * (Kpad / 8);
  count_launch(1);
  if (pix_bf16)
    im2col_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)pixels, (__nv_bfloat16*)out, B, H, W, patch, Kpad);
  else
    im2col_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)pixels,
                                                                                 (__nv_bfloat16*)out, B, H, W,
                                                                                 patch, Kpad);
  return report_cuda(cudaGetLastError(), "im2col_kernel launch");
}

extern "C" int vlmclip_vision_embed_ln(const void* patch, int patch_bf16, const float* cls, const float* pos,
                                       const float* gamma, const float* beta, void* y, void* y_lo, int B, int S, int D,
                                       float eps, void* stream) {
  VLMCLIP_CHECK_ARG(patch && cls && pos && gamma && beta && y, "vision_embed_ln: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && S > 1 && D % 8 == 0 && D <= MAX_VEC * 256, "vision_embed_ln: bad dims B=%d S=%d D=%d", B, S, D);
  VLMCLIP_CHECK_ARG((uintptr_t)patch % 16 == 0 && (uintptr_t)cls % 16 == 0 && (uintptr_t)pos % 16 == 0 &&
                        (uintptr_t)y % 16 == 0 && (uintptr_t)y_lo % 16 == 0 && (uintptr_t)gamma % 16 == 0 &&
                        (uintptr_t)beta % 16 == 0,
                    "vision_embed_ln: pointers must be 16-byte aligned");
  const int64_t rows = (int64_t)B * S;
  const int grid = (int)((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK);
  count_launch(1);
  if (patch_bf16) {
    VLMCLIP_DISPATCH_NV(D, (vision_embed_ln_kernel<NV, true><<<grid, ROWS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
                               patch, cls, pos, gamma, beta, (__nv_bfloat16*)y, (__nv_bfloat16*)y_lo, B, S, D, eps)));
  } else {
    VLMCLIP_DISPATCH_NV(D, (vision_embed_ln_kernel<NV, false><<<grid, ROWS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
                               patch, cls, pos, gamma, beta, (__nv_bfloat16*)y, (__nv_bfloat16*)y_lo, B, S, D, eps)));
  }
  return report_cuda(cudaGetLastError(), "vision_embed_ln_kernel launch");
}

extern "C" int vlmclip_text_embed(const int64_t* ids, const void* tok, int tok_bf16, const float* pos, void* y, void* y_lo,
                                  int B, int S, int D, int V, void* stream) {
  VLMCLIP_CHECK_ARG(ids && tok && pos && y, "text_embed: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && S > 0 && D % 8 == 0 && V > 0, "text_embed: bad dims");
  VLMCLIP_CHECK_ARG((uintptr_t)y % 16 == 0 && (uintptr_t)y_lo % 16 == 0, "text_embed: y / y_lo must be 16-byte aligned");
  const int64_t rows = (int64_t)B * S;
  const int grid = (int)((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK);
  count_launch(1);
  if (tok_bf16)
    text_embed_kernel<__nv_bfloat16><<<grid, ROWS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
        ids, (const __nv_bfloat16*)tok, pos, (__nv_bfloat16*)y, (__nv_bfloat16*)y_lo, B, S, D, V);
  else
    text_embed_kernel<float><<<grid, ROWS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(
        ids, (const float*)tok, pos, (__nv_bfloat16*)y, (__nv_bfloat16*)y_lo, B, S, D, V);
  return report_cuda(cudaGetLastError(), "text_embed_kernel launch");
}

extern "C" int vlmclip_gather_rows2_bf16_to_f32(const void* x, const void* x_lo, int64_t ldx, float* y, int R, int D,
                                                void* stream) {
  VLMCLIP_CHECK_ARG(x && y && R > 0 && D > 0 && ldx >= D, "gather_rows: bad arguments");
  count_launch(1);
  gather_rows_kernel<<<grid_for((int64_t)R * D, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)x_lo, ldx, y, R, D);
  return report_cuda(cudaGetLastError(), "gather_rows_kernel launch");
}

extern "C" int vlmclip_gather_rows_bf16_to_f32(const void* x, int64_t ldx, float* y, int R, int D, void* stream) {
  return vlmclip_gather_rows2_bf16_to_f32(x, nullptr, ldx, y, R, D, stream);
}

extern "C" int vlmclip_ln_partials_to_stats(const float* partials, float* stats_out, int M, int npart, float eps,
                                            void* stream) {
  VLMCLIP_CHECK_ARG(partials && stats_out && M > 0 && npart > 0 && npart <= 64, "ln_partials_to_stats: bad arguments");
  count_launch(1);
  return report_cuda(launch_pdl(ln_partials_to_stats_kernel, dim3((M + 63) / 64), dim3(256), 0, (cudaStream_t)stream, 1,
                                (const float2*)partials, (float2*)stats_out, M, npart, eps),
                     "ln_partials_to_stats_kernel launch");
}

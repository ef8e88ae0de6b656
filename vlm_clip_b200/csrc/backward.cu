// Row / elementwise kernels of the backbone BACKWARD (full fine-tune, BASELINE config 5): the reference gets these
// from autograd over HF modeling_clip.py when `freeze_clip=False` (model_m.py:22,72-75):
//   LayerNorm backward (HF:371,380,562,677), quick_gelu forward/backward (HF:349), bias gradients (row sums of the
//   transposed output gradient), layout transposes that put the activation-gradient products into the
//   C = A W^T form of the tcgen05 GEMM, embedding gradients (HF:202-218, 234-258).
// All HBM-bound: one warp per row or one thread per 8 elements, 16-byte accesses, fp32 arithmetic, fixed-order
// reductions (the only atomics are the token-embedding scatter, where rows collide by construction).
#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

int grid_cap(int64_t total, int block, int per_sm = 32) {
  int64_t g = (total + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

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
__device__ __forceinline__ void load8_f32(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8_f32(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(f[4], f[5], f[6], f[7]);
}

// dst[c, r] = bf16(src[row(r), c]) for r < R, 0 for R <= r < Rpad.  row(r) = r, or with a row gather
// (group_dst > 0): row(r) = (r / group_dst) * group_src + group_off + r % group_dst  (drop the CLS row of every image)
template <typename T>
__global__ void __launch_bounds__(256)
transpose_to_bf16_kernel(const T* __restrict__ src, int64_t lds, __nv_bfloat16* __restrict__ dst, int64_t ldd, int R,
                         int Rpad, int C, int group_dst, int group_src, int group_off) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      const int64_t sr = group_dst > 0 ? (int64_t)(r / group_dst) * group_src + group_off + (r % group_dst) : r;
      v = to_f(src[sr * lds + c]);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < Rpad) dst[(int64_t)c * ldd + r] = __float2bfloat16(tile[tx][i]);
  }
}

// Fast path (C % 8 == 0, 16-byte aligned rows): 64 x 64 tiles, every global access a 16-byte vector, 128 contiguous
// bytes per 8 lanes on both sides.  The tile sits in shared memory as 64 rows x 8 chunks of 16 B with the chunk
// index XOR-swizzled by (row / 8), so that the column gathers of the write phase (8 rows r = 8k + i of one column,
// k = lane % 8) fall into 8 different 4-bank groups.
template <typename T>
__global__ void __launch_bounds__(256)
transpose64_to_bf16_kernel(const T* __restrict__ src, int64_t lds, __nv_bfloat16* __restrict__ dst, int64_t ldd, int R,
                           int Rpad, int C, int group_dst, int group_src, int group_off) {
  __shared__ __align__(16) __nv_bfloat16 tile[64 * 64];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int q = threadIdx.x + 256 * p;
    const int r = q >> 3, ch = q & 7;
    const int gr = r0 + r, gc = c0 + ch * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (gr < R && gc < C) {
      const int64_t sr = group_dst > 0 ? (int64_t)(gr / group_dst) * group_src + group_off + (gr % group_dst) : gr;
      if (sizeof(T) == 4) {
        float f[8];
        load8_f32(reinterpret_cast<const float*>(src) + sr * lds + gc, f);
        v = pack8(f);
      } else {
        v = ld_nc_v4(reinterpret_cast<const __nv_bfloat16*>(src) + sr * lds + gc);
      }
    }
    *reinterpret_cast<uint4*>(tile + r * 64 + ((ch ^ (r >> 3)) << 3)) = v;
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int q = threadIdx.x + 256 * p;
    const int c = q >> 3, k = q & 7;
    if (c0 + c < C && r0 + 8 * k < Rpad) {
      const int pc = (((c >> 3) ^ k) << 3) + (c & 7);  // swizzled column of rows 8k .. 8k+7
      const unsigned short* t16 = reinterpret_cast<const unsigned short*>(tile) + (8 * k) * 64 + pc;
      uint4 o;
      o.x = (uint32_t)t16[0] | ((uint32_t)t16[64] << 16);
      o.y = (uint32_t)t16[128] | ((uint32_t)t16[192] << 16);
      o.z = (uint32_t)t16[256] | ((uint32_t)t16[320] << 16);
      o.w = (uint32_t)t16[384] | ((uint32_t)t16[448] << 16);
      st_v4(dst + (int64_t)(c0 + c) * ldd + r0 + 8 * k, o);
    }
  }
}

__global__ void __launch_bounds__(256)
cast_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n, int vec_ok) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t n8 = n >> 3;
    for (int64_t i = t0; i < n8; i += stride) {
      float f[8];
      load8_f32(src + i * 8, f);
      st_v4(dst + i * 8, pack8(f));
    }
    for (int64_t i = (n8 << 3) + t0; i < n; i += stride) dst[i] = __float2bfloat16(src[i]);
  } else {
    for (int64_t i = t0; i < n; i += stride) dst[i] = __float2bfloat16(src[i]);
  }
}

// out[i] = sum_s parts[s * stride + i], s in index order (second pass of a split reduction)
__global__ void __launch_bounds__(256)
sum_planes_f32_kernel(const float* __restrict__ parts, int64_t stride, int planes, float* __restrict__ out, int64_t n4) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += step) {
    float4 a = __ldg(reinterpret_cast<const float4*>(parts) + i);
    for (int s = 1; s < planes; ++s) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(parts + (int64_t)s * stride) + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
  }
}

// acc[i] += float(b[i]): a bf16 gradient branch joins the fp32 gradient of a skip connection
__global__ void __launch_bounds__(256)
add_bf16_into_f32_kernel(float* __restrict__ acc, const __nv_bfloat16* __restrict__ b, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) acc[i] += __bfloat162float(b[i]);
}

// out[r] = sum_c x[r, c]: one CTA per row (rows are long: the token dimension of a transposed gradient)
__global__ void __launch_bounds__(256)
rowsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, float* __restrict__ out, int C) {
  __shared__ float red[8];
  const __nv_bfloat16* xr = x + (int64_t)blockIdx.x * ldx;
  float s = 0.f;
  const int c8 = C & ~7;
  for (int c = threadIdx.x * 8; c < c8; c += 256 * 8) {
    float f[8];
    unpack8(ld_nc_v4(xr + c), f);
    s += ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));
  }
  for (int c = c8 + threadIdx.x; c < C; c += 256) s += __bfloat162float(xr[c]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    out[blockIdx.x] = t;
  }
}

// out[c] = sum_r x[r * ldx + c]  (fp32; r runs over the batch: position / class embedding gradients)
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ out, int R, int64_t C) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += x[(int64_t)r * ldx + c];
  out[c] = s;
}

// part[slice][c] = sum over the rows of the slice of x[r, c], x bf16 [R, C]: first pass of the bias gradient (column sums
// of dY) without a transposed copy of dY.  Block = 32 column vectors (8 columns each) x 8 row lanes; fixed order.
__global__ void __launch_bounds__(256)
colsum_bf16_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, float* __restrict__ part, int R, int C,
                           int rows_per_slice) {
  __shared__ float red[8][32][9];
  const int cv = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cv) * 8;
  const int r0 = blockIdx.y * rows_per_slice;
  const int r1 = min(R, r0 + rows_per_slice);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    for (int r = r0 + rl; r < r1; r += 8) {
      float f[8];
      unpack8(ld_nc_v4(x + (int64_t)r * ldx + c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cv][j] = acc[j];
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
      for (int k = 0; k < 8; ++k) s += red[k][cv][j];
      part[(int64_t)blockIdx.y * C + c + j] = s;
    }
  }
}

// quick_gelu (HF:349): y = a * sigmoid(1.702 a), the same one-MUFU form as the fused GEMM epilogue
__global__ void __launch_bounds__(256)
quick_gelu_kernel(const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ y, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(ld_nc_v4(a + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = quick_gelu(f[j]);
    st_v4(y + i * 8, pack8(f));
  }
}
// da = dy * d/da [a sigmoid(1.702 a)] = dy * s (1 + 1.702 a (1 - s))
__global__ void __launch_bounds__(256)
quick_gelu_bwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ dy,
                      __nv_bfloat16* __restrict__ da, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8], g[8];
    unpack8(ld_nc_v4(a + i * 8), f);
    unpack8(ld_nc_v4(dy + i * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = fmaf(0.5f, fast_tanh(0.851f * f[j]), 0.5f);
      g[j] *= s * fmaf(1.702f * f[j], 1.f - s, 1.f);
    }
    st_v4(da + i * 8, pack8(g));
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm backward.  y = xhat * gamma + beta, xhat = (x - mean) * rstd.
//   dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)),  g = gamma * dy     (+ dres: the skip connection's gradient)
//   dgamma = sum_rows dy * xhat,  dbeta = sum_rows dy
// One warp per row, statistics recomputed from the saved bf16 input (two-pass in registers); each lane keeps the
// dgamma / dbeta contributions of its own columns in registers across the rows it visits; the CTA's warps are
// combined through shared memory and written as one partial row per CTA, summed by ln_bwd_reduce_kernel.
// ---------------------------------------------------------------------------------------------------------
constexpr int LNB_WARPS = 4;  // 128 threads at ~150 registers: three CTAs = 12 warps per SM (8 warps x 1 CTA ran at 0.31 of HBM peak)
constexpr int LNB_MAXD = 2048;

// combine the per-lane column sums of a CTA's warps: upper half -> smem, lower half adds, all threads sum the rest
template <int NV>
__device__ __forceinline__ void lnb_block_combine(const float (&acc)[NV][8], const int (&valid)[NV],
                                                  float (*red)[LNB_MAXD], float* __restrict__ dst, int D, int warp,
                                                  int lane) {
  __syncthreads();
  if (warp >= LNB_WARPS / 2) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (valid[i])
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp - LNB_WARPS / 2][(lane + i * 32) * 8 + j] = acc[i][j];
  }
  __syncthreads();
  if (warp < LNB_WARPS / 2) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (valid[i])
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp][(lane + i * 32) * 8 + j] += acc[i][j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += LNB_WARPS * 32) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_WARPS / 2; ++w) t += red[w][c];
    dst[c] = t;
  }
}

template <int NV, bool DY_F32>
__global__ void __launch_bounds__(LNB_WARPS * 32, NV <= 3 ? 3 : (NV == 4 ? 2 : 1))
layernorm_bwd_kernel(const void* __restrict__ dy_, int64_t lddy, const __nv_bfloat16* __restrict__ x, int64_t ldx,
                     const float* __restrict__ gamma, const float* __restrict__ dres, float* __restrict__ dx32,
                     __nv_bfloat16* __restrict__ dx16, int64_t lddx, float* __restrict__ partial, int M, int D,
                     float eps) {
  __shared__ float red[LNB_WARPS / 2][LNB_MAXD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  float ag[NV][8], ab[NV][8];
  int valid[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    valid[i] = (lane + i * 32) < nvec;
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = 0.f;
  }
  const float invD = 1.f / (float)D;
  for (int64_t row = blockIdx.x * (int64_t)LNB_WARPS + warp; row < M; row += (int64_t)gridDim.x * LNB_WARPS) {
    float v[NV][8], g[NV][8];
    const __nv_bfloat16* xr = x + row * ldx;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (valid[i]) {
        unpack8(ld_nc_v4(xr + (lane + i * 32) * 8), v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    const float mean = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (valid[i])
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[i][j] -= mean;
          q = fmaf(v[i][j], v[i][j], q);
        }
    const float rstd = rsqrtf(warp_sum(q) * invD + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (valid[i]) {
        const int c = (lane + i * 32) * 8;
        float dy[8], gm[8];
        if (DY_F32)
          load8_f32(reinterpret_cast<const float*>(dy_) + row * lddy + c, dy);
        else
          unpack8(ld_nc_v4(reinterpret_cast<const __nv_bfloat16*>(dy_) + row * lddy + c), dy);
        load8_f32(gamma + c, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[i][j] *= rstd;  // xhat
          ag[i][j] = fmaf(dy[j], v[i][j], ag[i][j]);
          ab[i][j] += dy[j];
          g[i][j] = gm[j] * dy[j];
          s1 += g[i][j];
          s2 = fmaf(g[i][j], v[i][j], s2);
        }
      }
    const float m1 = warp_sum(s1) * invD, m2 = warp_sum(s2) * invD;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (valid[i]) {
        const int c = (lane + i * 32) * 8;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - m1 - v[i][j] * m2);
        if (dres != nullptr) {
          float r[8];
          load8_f32(dres + row * lddx + c, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        if (dx32 != nullptr) store8_f32(dx32 + row * lddx + c, o);
        if (dx16 != nullptr) st_v4(dx16 + row * lddx + c, pack8(o));
      }
  }
  if (partial == nullptr) return;
  lnb_block_combine<NV>(ag, valid, red, partial + ((int64_t)blockIdx.x * 2 + 0) * D, D, warp, lane);
  lnb_block_combine<NV>(ab, valid, red, partial + ((int64_t)blockIdx.x * 2 + 1) * D, D, warp, lane);
}

// dgamma[c] = sum_blocks partial[b][0][c], dbeta[c] = sum_blocks partial[b][1][c]
__global__ void __launch_bounds__(256)
ln_bwd_reduce_kernel(const float* __restrict__ partial, int nblk, int D, float* __restrict__ dgamma,
                     float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * D) return;
  float t = 0.f;
  for (int b = 0; b < nblk; ++b) t += partial[(int64_t)b * 2 * D + c];
  if (c < D)
    dgamma[c] = t;
  else
    dbeta[c - D] = t;
}

// vision tokens WITHOUT the LayerNorm (training keeps the pre-LN rows for the backward):
//   e[b,0] = cls + pos[0]; e[b,1+q] = patch[b,q] + pos[1+q]      (HF:211-218)
__global__ void __launch_bounds__(256)
vision_embed_kernel(const __nv_bfloat16* __restrict__ patch, const float* __restrict__ cls,
                    const float* __restrict__ pos, __nv_bfloat16* __restrict__ e, int B, int S, int D) {
  const int dv = D >> 3;
  const int64_t total = (int64_t)B * S * dv;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % dv) * 8;
    const int64_t row = idx / dv;
    const int s = (int)(row % S);
    const int64_t b = row / S;
    float a[8], p[8];
    if (s == 0)
      load8_f32(cls + c, a);
    else
      unpack8(ld_nc_v4(patch + (b * (S - 1) + (s - 1)) * (int64_t)D + c), a);
    load8_f32(pos + (int64_t)s * D + c, p);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += p[j];
    st_v4(e + row * D + c, pack8(a));
  }
}

// dtok[ids[r]] += d[r]   (token_embedding.weight gradient, HF:248; rows collide whenever a token repeats)
__global__ void __launch_bounds__(256)
embed_scatter_add_kernel(const float* __restrict__ d, int64_t ldd, const int64_t* __restrict__ ids,
                         float* __restrict__ dtok, int64_t rows, int D, int V) {
  const int64_t total = rows * D;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / D;
    const int c = (int)(idx - r * D);
    const float v = d[r * ldd + c];
    if (v == 0.f) continue;  // Track M: only token 0 of every caption carries gradient
    int64_t id = ids[r];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    atomicAdd(dtok + id * D + c, v);
  }
}

}  // namespace
}  // namespace vlmclip

using namespace vlmclip;

#define VLMCLIP_BWD_DISPATCH_NV(D, CALL)      \
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

extern "C" int vlmclip_transpose_to_bf16(const void* src, int src_f32, int64_t lds, void* dst, int64_t ldd, int R,
                                         int Rpad, int C, int group_dst, int group_src, int group_off, void* stream) {
  VLMCLIP_CHECK_ARG(src && dst && R > 0 && C > 0, "transpose: bad arguments");
  VLMCLIP_CHECK_ARG(Rpad >= R && ldd >= Rpad && lds >= C, "transpose: Rpad=%d R=%d ldd=%lld lds=%lld C=%d", Rpad, R,
                    (long long)ldd, (long long)lds, C);
  VLMCLIP_CHECK_ARG(group_dst == 0 || (group_dst > 0 && group_src >= group_dst + group_off && group_off >= 0),
                    "transpose: bad row gather (%d, %d, %d)", group_dst, group_src, group_off);
  count_launch(1);
  const int esz = src_f32 ? 4 : 2;
  if (C % 8 == 0 && Rpad % 8 == 0 && ldd % 8 == 0 && (lds * esz) % 16 == 0 && (uintptr_t)src % 16 == 0 &&
      (uintptr_t)dst % 16 == 0 && (C + 63) / 64 <= 65535) {
    dim3 grid64((Rpad + 63) / 64, (C + 63) / 64);
    if (src_f32)
      transpose64_to_bf16_kernel<float><<<grid64, 256, 0, (cudaStream_t)stream>>>(
          (const float*)src, lds, (__nv_bfloat16*)dst, ldd, R, Rpad, C, group_dst, group_src, group_off);
    else
      transpose64_to_bf16_kernel<__nv_bfloat16><<<grid64, 256, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)src, lds, (__nv_bfloat16*)dst, ldd, R, Rpad, C, group_dst, group_src, group_off);
    return report_cuda(cudaGetLastError(), "transpose64_to_bf16_kernel launch");
  }
  dim3 grid((Rpad + 31) / 32, (C + 31) / 32);
  VLMCLIP_CHECK_ARG(grid.y <= 65535, "transpose: C=%d too large", C);
  if (src_f32)
    transpose_to_bf16_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, lds, (__nv_bfloat16*)dst,
                                                                            ldd, R, Rpad, C, group_dst, group_src, group_off);
  else
    transpose_to_bf16_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)src, lds, (__nv_bfloat16*)dst, ldd, R, Rpad, C, group_dst, group_src, group_off);
  return report_cuda(cudaGetLastError(), "transpose_to_bf16_kernel launch");
}

extern "C" int vlmclip_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  VLMCLIP_CHECK_ARG(src && dst && n > 0, "cast: bad arguments");
  const int vec_ok = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  count_launch(1);
  cast_f32_to_bf16_kernel<<<grid_cap((n + 7) / 8, 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n,
                                                                                       vec_ok);
  return report_cuda(cudaGetLastError(), "cast_f32_to_bf16_kernel launch");
}

extern "C" int vlmclip_sum_planes_f32(const float* parts, int64_t plane_stride, int planes, float* out, int64_t n, void* stream) {
  VLMCLIP_CHECK_ARG(parts && out && planes >= 1 && n > 0 && n % 4 == 0 && plane_stride % 4 == 0,
                    "sum_planes: n and plane_stride must be positive multiples of 4");
  VLMCLIP_CHECK_ARG((uintptr_t)parts % 16 == 0 && (uintptr_t)out % 16 == 0, "sum_planes: pointers must be 16-byte aligned");
  count_launch(1);
  sum_planes_f32_kernel<<<grid_cap(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(parts, plane_stride, planes, out, n / 4);
  return report_cuda(cudaGetLastError(), "sum_planes_f32_kernel launch");
}

extern "C" int vlmclip_add_bf16_into_f32(float* acc, const void* b, int64_t n, void* stream) {
  VLMCLIP_CHECK_ARG(acc && b && n > 0, "add_bf16_into_f32: bad arguments");
  count_launch(1);
  add_bf16_into_f32_kernel<<<grid_cap(n, 256), 256, 0, (cudaStream_t)stream>>>(acc, (const __nv_bfloat16*)b, n);
  return report_cuda(cudaGetLastError(), "add_bf16_into_f32_kernel launch");
}

extern "C" int vlmclip_rowsum_bf16(const void* x, int64_t ldx, float* out, int R, int C, void* stream) {
  VLMCLIP_CHECK_ARG(x && out && R > 0 && C > 0 && ldx >= C, "rowsum: bad arguments");
  VLMCLIP_CHECK_ARG(ldx % 8 == 0 && (uintptr_t)x % 16 == 0, "rowsum: x must be 16-byte aligned with ldx %% 8 == 0");
  count_launch(1);
  rowsum_bf16_kernel<<<R, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, out, C);
  return report_cuda(cudaGetLastError(), "rowsum_bf16_kernel launch");
}

extern "C" int vlmclip_colsum_f32(const float* x, int64_t ldx, float* out, int R, int64_t C, void* stream) {
  VLMCLIP_CHECK_ARG(x && out && R > 0 && C > 0 && ldx >= C, "colsum: bad arguments");
  count_launch(1);
  colsum_f32_kernel<<<(unsigned)((C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, out, R, C);
  return report_cuda(cudaGetLastError(), "colsum_f32_kernel launch");
}

// out[c] = sum_r x[r, c], x bf16 [R, C] (ldx elements between rows), C % 8 == 0; workspace: vlmclip_colsum_bf16_slices(R) * C floats
extern "C" int vlmclip_colsum_bf16_slices(int R) {
  const int s = (R + 511) / 512;
  return s < 1 ? 1 : (s > 128 ? 128 : s);
}
extern "C" int vlmclip_colsum_bf16(const void* x, int64_t ldx, float* out, float* workspace, int R, int C, void* stream) {
  VLMCLIP_CHECK_ARG(x && out && workspace && R > 0 && C > 0 && C % 8 == 0 && ldx >= C && ldx % 8 == 0,
                    "colsum_bf16: C and ldx must be positive multiples of 8");
  VLMCLIP_CHECK_ARG((uintptr_t)x % 16 == 0, "colsum_bf16: x must be 16-byte aligned");
  const int slices = vlmclip_colsum_bf16_slices(R);
  const int rows_per_slice = (R + slices - 1) / slices;
  count_launch(2);
  colsum_bf16_partial_kernel<<<dim3((C / 8 + 31) / 32, slices), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, ldx, workspace, R, C, rows_per_slice);
  colsum_f32_kernel<<<(unsigned)((C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(workspace, C, out, slices, C);
  return report_cuda(cudaGetLastError(), "colsum_bf16 launch");
}

extern "C" int vlmclip_quick_gelu_bf16(const void* a, void* y, int64_t n, void* stream) {
  VLMCLIP_CHECK_ARG(a && y && n > 0 && n % 8 == 0, "quick_gelu: n must be a positive multiple of 8");
  VLMCLIP_CHECK_ARG((uintptr_t)a % 16 == 0 && (uintptr_t)y % 16 == 0, "quick_gelu: pointers must be 16-byte aligned");
  count_launch(1);
  quick_gelu_kernel<<<grid_cap(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (__nv_bfloat16*)y,
                                                                           n / 8);
  return report_cuda(cudaGetLastError(), "quick_gelu_kernel launch");
}

extern "C" int vlmclip_quick_gelu_bwd_bf16(const void* a, const void* dy, void* da, int64_t n, void* stream) {
  VLMCLIP_CHECK_ARG(a && dy && da && n > 0 && n % 8 == 0, "quick_gelu_bwd: n must be a positive multiple of 8");
  VLMCLIP_CHECK_ARG((uintptr_t)a % 16 == 0 && (uintptr_t)dy % 16 == 0 && (uintptr_t)da % 16 == 0,
                    "quick_gelu_bwd: pointers must be 16-byte aligned");
  count_launch(1);
  quick_gelu_bwd_kernel<<<grid_cap(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)a, (const __nv_bfloat16*)dy, (__nv_bfloat16*)da, n / 8);
  return report_cuda(cudaGetLastError(), "quick_gelu_bwd_kernel launch");
}

static int ln_bwd_blocks(int M) {
  const int need = (M + LNB_WARPS - 1) / LNB_WARPS;
  const int cap = sm_count() * 3;
  return need < cap ? need : cap;
}

extern "C" int64_t vlmclip_layernorm_bwd_workspace(int M, int D) { return (int64_t)ln_bwd_blocks(M) * 2 * D; }

extern "C" int vlmclip_layernorm_bwd(const void* dy, int dy_f32, int64_t lddy, const void* x, int64_t ldx,
                                     const float* gamma, const float* dres, float* dx_f32, void* dx_bf16, int64_t lddx,
                                     float* dgamma, float* dbeta, float* workspace, int M, int D, float eps,
                                     void* stream) {
  VLMCLIP_CHECK_ARG(dy && x && gamma, "layernorm_bwd: null pointer");
  VLMCLIP_CHECK_ARG(M > 0 && D > 0 && D % 8 == 0 && D <= LNB_MAXD, "layernorm_bwd: D=%d must be a multiple of 8, <= %d", D,
                    LNB_MAXD);
  VLMCLIP_CHECK_ARG(ldx % 8 == 0 && lddy % 8 == 0 && ldx >= D && lddy >= D, "layernorm_bwd: bad ldx / lddy");
  VLMCLIP_CHECK_ARG((dx_f32 == nullptr && dx_bf16 == nullptr && dres == nullptr) || (lddx % 8 == 0 && lddx >= D),
                    "layernorm_bwd: bad lddx");
  VLMCLIP_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr) && (dgamma == nullptr || workspace != nullptr),
                    "layernorm_bwd: dgamma, dbeta and workspace go together");
  VLMCLIP_CHECK_ARG((uintptr_t)dy % 16 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)gamma % 16 == 0 &&
                        (uintptr_t)dres % 16 == 0 && (uintptr_t)dx_f32 % 16 == 0 && (uintptr_t)dx_bf16 % 16 == 0,
                    "layernorm_bwd: pointers must be 16-byte aligned");
  const int nblk = ln_bwd_blocks(M);
  float* partial = dgamma != nullptr ? workspace : nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  count_launch(1);
  if (dy_f32) {
    VLMCLIP_BWD_DISPATCH_NV(D, (layernorm_bwd_kernel<NV, true><<<nblk, LNB_WARPS * 32, 0, s>>>(
                                   dy, lddy, (const __nv_bfloat16*)x, ldx, gamma, dres, dx_f32, (__nv_bfloat16*)dx_bf16,
                                   lddx, partial, M, D, eps)));
  } else {
    VLMCLIP_BWD_DISPATCH_NV(D, (layernorm_bwd_kernel<NV, false><<<nblk, LNB_WARPS * 32, 0, s>>>(
                                   dy, lddy, (const __nv_bfloat16*)x, ldx, gamma, dres, dx_f32, (__nv_bfloat16*)dx_bf16,
                                   lddx, partial, M, D, eps)));
  }
  if (dgamma != nullptr) {
    count_launch(1);
    ln_bwd_reduce_kernel<<<(2 * D + 255) / 256, 256, 0, s>>>(partial, nblk, D, dgamma, dbeta);
  }
  return report_cuda(cudaGetLastError(), "layernorm_bwd launch");
}

extern "C" int vlmclip_vision_embed(const void* patch, const float* cls, const float* pos, void* e, int B, int S, int D,
                                    void* stream) {
  VLMCLIP_CHECK_ARG(patch && cls && pos && e, "vision_embed: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && S > 1 && D % 8 == 0, "vision_embed: bad dims B=%d S=%d D=%d", B, S, D);
  VLMCLIP_CHECK_ARG((uintptr_t)patch % 16 == 0 && (uintptr_t)cls % 16 == 0 && (uintptr_t)pos % 16 == 0 &&
                        (uintptr_t)e % 16 == 0,
                    "vision_embed: pointers must be 16-byte aligned");
  count_launch(1);
  vision_embed_kernel<<<grid_cap((int64_t)B * S * (D / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)patch, cls, pos, (__nv_bfloat16*)e, B, S, D);
  return report_cuda(cudaGetLastError(), "vision_embed_kernel launch");
}

extern "C" int vlmclip_embed_scatter_add(const float* d, int64_t ldd, const int64_t* ids, float* dtok, int64_t rows,
                                         int D, int V, void* stream) {
  VLMCLIP_CHECK_ARG(d && ids && dtok && rows > 0 && D > 0 && V > 0 && ldd >= D, "embed_scatter_add: bad arguments");
  count_launch(1);
  embed_scatter_add_kernel<<<grid_cap(rows * D, 256), 256, 0, (cudaStream_t)stream>>>(d, ldd, ids, dtok, rows, D, V);
  return report_cuda(cudaGetLastError(), "embed_scatter_add_kernel launch");
}

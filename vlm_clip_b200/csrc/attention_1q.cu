// Single-query attention: one query row per (batch, head) against S keys, head_dim 64.
//   out[b, h*64:(h+1)*64] = softmax(q[b,h] . K[b or shared][:, h]^T * scale) V[...][:, h]
// Two users: (1) the cross-modal adapter (adapter/clip_adapter.py:99-128 as called by model_m.py:93-100): the keys and
// values come from the vision position table, the same for every caption (kv_batch_stride = 0), and only token 0 of
// each caption is consumed downstream (model_m.py:102); (2) the CLS-only evaluation of the last vision layer
// (model_m.py:122 keeps token 0 of last_hidden_state), where K and V are columns of the fused qkv activation.
// One warp per (b, h): each lane scores the keys l, l+32, ... against the query held in registers (fp32 softmax, as
// HF:272), then the lanes own two output dimensions each and accumulate p_j V[j] with p_j broadcast by shuffle.
#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int Q1_WARPS = 8;
constexpr int Q1_MAX_KEYS_PER_LANE = 16;  // S <= 512

__global__ void __launch_bounds__(Q1_WARPS * 32)
attention_1q_kernel(const __nv_bfloat16* __restrict__ q, int64_t q_stride, const __nv_bfloat16* __restrict__ k,
                    const __nv_bfloat16* __restrict__ v, int64_t kv_row_stride, int64_t kv_batch_stride,
                    __nv_bfloat16* __restrict__ out, int64_t out_stride, int B, int S, int H, float scale_log2e) {
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x * Q1_WARPS + warp;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh - b * H;
  const __nv_bfloat16* qp = q + (int64_t)b * q_stride + h * 64;
  const __nv_bfloat16* kp = k + (int64_t)b * kv_batch_stride + h * 64;
  const __nv_bfloat16* vp = v + (int64_t)b * kv_batch_stride + h * 64;

  float qf[64];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 raw = ld_nc_v4(qp + c * 8);
    qf[c * 8 + 0] = bf16_lo(raw.x); qf[c * 8 + 1] = bf16_hi(raw.x);
    qf[c * 8 + 2] = bf16_lo(raw.y); qf[c * 8 + 3] = bf16_hi(raw.y);
    qf[c * 8 + 4] = bf16_lo(raw.z); qf[c * 8 + 5] = bf16_hi(raw.z);
    qf[c * 8 + 6] = bf16_lo(raw.w); qf[c * 8 + 7] = bf16_hi(raw.w);
  }
  float sc[Q1_MAX_KEYS_PER_LANE];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < Q1_MAX_KEYS_PER_LANE; ++i) {
    const int j = lane + i * 32;
    float s = -INFINITY;
    if (j < S) {
      const __nv_bfloat16* kr = kp + (int64_t)j * kv_row_stride;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 raw = ld_nc_v4(kr + c * 8);
        acc = fmaf(qf[c * 8 + 0], bf16_lo(raw.x), acc); acc = fmaf(qf[c * 8 + 1], bf16_hi(raw.x), acc);
        acc = fmaf(qf[c * 8 + 2], bf16_lo(raw.y), acc); acc = fmaf(qf[c * 8 + 3], bf16_hi(raw.y), acc);
        acc = fmaf(qf[c * 8 + 4], bf16_lo(raw.z), acc); acc = fmaf(qf[c * 8 + 5], bf16_hi(raw.z), acc);
        acc = fmaf(qf[c * 8 + 6], bf16_lo(raw.w), acc); acc = fmaf(qf[c * 8 + 7], bf16_hi(raw.w), acc);
      }
      s = acc * scale_log2e;
    }
    sc[i] = s;
    m = fmaxf(m, s);
  }
  m = warp_max(m);
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < Q1_MAX_KEYS_PER_LANE; ++i) {
    sc[i] = (lane + i * 32 < S) ? exp2f(sc[i] - m) : 0.f;
    l += sc[i];
  }
  l = warp_sum(l);
  // lanes own output dimensions 2*lane, 2*lane+1
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int i = 0; i < Q1_MAX_KEYS_PER_LANE; ++i) {
    if (i * 32 < S) {  // warp-uniform
      const int nj = min(32, S - i * 32);
      const __nv_bfloat16* vrow = vp + (int64_t)(i * 32) * kv_row_stride + lane * 2;
      int jj = 0;
      for (; jj + 8 <= nj; jj += 8) {  // eight independent 128-byte row loads in flight per warp
        uint32_t raw[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) raw[e] = *reinterpret_cast<const uint32_t*>(vrow + (int64_t)(jj + e) * kv_row_stride);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float pj = __shfl_sync(0xffffffffu, sc[i], jj + e);
          o0 = fmaf(pj, bf16_lo(raw[e]), o0);
          o1 = fmaf(pj, bf16_hi(raw[e]), o1);
        }
      }
      for (; jj < nj; ++jj) {
        const float pj = __shfl_sync(0xffffffffu, sc[i], jj);
        const uint32_t raw = *reinterpret_cast<const uint32_t*>(vrow + (int64_t)jj * kv_row_stride);
        o0 = fmaf(pj, bf16_lo(raw), o0);
        o1 = fmaf(pj, bf16_hi(raw), o1);
      }
    }
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  *reinterpret_cast<uint32_t*>(out + (int64_t)b * out_stride + h * 64 + lane * 2) = pack_bf16x2(o0 * inv, o1 * inv);
}

}  // namespace

// Output row b at out + b * out_stride elements (the C entry point below: out_stride = H * 64; the key-range split of
// attention_pp.cu: the tail rows of every sequence inside the [B*S, D] attention output).
int attention_1q_strided(const void* q, int64_t q_stride, const void* k, const void* v, int64_t kv_row_stride,
                         int64_t kv_batch_stride, void* out, int64_t out_stride, int B, int S, int H, float scale,
                         cudaStream_t stream) {
  VLMCLIP_CHECK_ARG(q && k && v && out, "attention_1q: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && H > 0 && S > 0 && S <= 32 * Q1_MAX_KEYS_PER_LANE, "attention_1q: bad dims B=%d S=%d H=%d", B, S, H);
  VLMCLIP_CHECK_ARG(q_stride % 8 == 0 && kv_row_stride % 8 == 0 && kv_batch_stride % 8 == 0 && out_stride % 2 == 0 &&
                        (uintptr_t)q % 16 == 0 && (uintptr_t)k % 16 == 0 && (uintptr_t)v % 4 == 0 && (uintptr_t)out % 4 == 0,
                    "attention_1q: strides must be multiples of 8 elements and pointers 16-byte aligned");
  count_launch(1);
  const int grid = (B * H + Q1_WARPS - 1) / Q1_WARPS;
  return report_cuda(launch_pdl(attention_1q_kernel, dim3(grid), dim3(Q1_WARPS * 32), 0, stream, 1,
                                (const __nv_bfloat16*)q, q_stride, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                                kv_row_stride, kv_batch_stride, (__nv_bfloat16*)out, out_stride, B, S, H,
                                scale * 1.4426950408889634f),
                     "attention_1q_kernel launch");
}

}  // namespace vlmclip

extern "C" int vlmclip_attention_1q(const void* q, int64_t q_stride, const void* k, const void* v, int64_t kv_row_stride,
                                    int64_t kv_batch_stride, void* out, int B, int S, int H, float scale, void* stream) {
  return vlmclip::attention_1q_strided(q, q_stride, k, v, kv_row_stride, kv_batch_stride, out, (int64_t)H * 64, B, S, H,
                                       scale, (cudaStream_t)stream);
}

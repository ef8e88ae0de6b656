// Backward of the attention core (full fine-tune only; autograd of HF modeling_clip.py:261-279 in the reference):
//   P = softmax(Q K^T * scale + mask),  O = P V
//   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - D),  D_i = dO_i . O_i,  dQ = scale * dS K,  dK = scale * dS^T Q
// First correct version, fp32 SIMT: one CTA per (batch, head) keeps Q, K, V, dO of the head in shared memory (bf16,
// rows padded to 144 B so that 16-byte row reads by the lanes of a warp are conflict free) and recomputes P from the
// saved activations (nothing but qkv and O is kept from the forward).  Phase 1: one warp per query row (scores,
// soft-max statistics, dQ); phase 2: one warp per key row (dK, dV) with the row statistics of phase 1.
// 7 x 2 S^2 64 FLOPs per head on the FMA pipe: ~16 TFLOP/s at best, i.e. several ms per ViT-B/16 layer at batch 256.
// The product path is the tensor-core pair of kernels in attention_bwd_mma.cu (selected by passing a workspace);
// this kernel remains as an independent implementation the parity tests compare it with.
#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int AB_WARPS = 8;
constexpr int AB_LD = 72;      // bf16 elements per shared-memory row (64 + 8 padding)
constexpr int AB_MAX_JT = 9;   // S <= 288

__device__ __forceinline__ void load_row(const __nv_bfloat16* row, uint32_t (&r)[32]) {
  const uint4* p = reinterpret_cast<const uint4*>(row);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 v = p[c];
    r[4 * c] = v.x;
    r[4 * c + 1] = v.y;
    r[4 * c + 2] = v.z;
    r[4 * c + 3] = v.w;
  }
}
// dot of a register-resident 64-vector (packed bf16 pairs) with a shared-memory row
__device__ __forceinline__ float dot64(const uint32_t (&a)[32], const __nv_bfloat16* row) {
  const uint4* p = reinterpret_cast<const uint4*>(row);
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 v = p[c];
    acc0 = fmaf(bf16_lo(a[4 * c]), bf16_lo(v.x), acc0);
    acc1 = fmaf(bf16_hi(a[4 * c]), bf16_hi(v.x), acc1);
    acc0 = fmaf(bf16_lo(a[4 * c + 1]), bf16_lo(v.y), acc0);
    acc1 = fmaf(bf16_hi(a[4 * c + 1]), bf16_hi(v.y), acc1);
    acc0 = fmaf(bf16_lo(a[4 * c + 2]), bf16_lo(v.z), acc0);
    acc1 = fmaf(bf16_hi(a[4 * c + 2]), bf16_hi(v.z), acc1);
    acc0 = fmaf(bf16_lo(a[4 * c + 3]), bf16_lo(v.w), acc0);
    acc1 = fmaf(bf16_hi(a[4 * c + 3]), bf16_hi(v.w), acc1);
  }
  return acc0 + acc1;
}

__global__ void __launch_bounds__(AB_WARPS * 32)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                     const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqkv,
                     const uint8_t* __restrict__ key_mask, int S, int Spad, int H, int causal, float scale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + (size_t)S * AB_LD;
  __nv_bfloat16* sV = sK + (size_t)S * AB_LD;
  __nv_bfloat16* sdO = sV + (size_t)S * AB_LD;
  float* sD = reinterpret_cast<float*>(sdO + (size_t)S * AB_LD);  // [Spad] dO_i . O_i
  float* sL = sD + Spad;                                           // [Spad] log-sum-exp of the scaled scores
  float* sP = sL + Spad;                                           // [warps][2][Spad]
  uint8_t* sM = reinterpret_cast<uint8_t*>(sP + (size_t)AB_WARPS * 2 * Spad);  // [Spad] 1 = key visible

  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int Dm = H * 64, D3 = 3 * Dm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tok0 = (int64_t)b * S;

  for (int idx = threadIdx.x; idx < S * 8; idx += AB_WARPS * 32) {
    const int s = idx >> 3, ch = idx & 7;
    const __nv_bfloat16* g = qkv + (tok0 + s) * D3 + h * 64 + ch * 8;
    *reinterpret_cast<uint4*>(sQ + s * AB_LD + ch * 8) = ld_nc_v4(g);
    *reinterpret_cast<uint4*>(sK + s * AB_LD + ch * 8) = ld_nc_v4(g + Dm);
    *reinterpret_cast<uint4*>(sV + s * AB_LD + ch * 8) = ld_nc_v4(g + 2 * Dm);
    *reinterpret_cast<uint4*>(sdO + s * AB_LD + ch * 8) = ld_nc_v4(dout + (tok0 + s) * Dm + h * 64 + ch * 8);
  }
  for (int s = threadIdx.x; s < Spad; s += AB_WARPS * 32)
    sM[s] = (s < S && (key_mask == nullptr || key_mask[tok0 + s] != 0)) ? 1 : 0;
  __syncthreads();

  float* myP = sP + (size_t)warp * 2 * Spad;
  float* myDS = myP + Spad;
  const int JT = (S + 31) >> 5;

  // ---------------- phase 1: query rows -> D_i, lse_i, dQ_i ----------------
  for (int i = warp; i < S; i += AB_WARPS) {
    // D_i = dO_i . O_i (O from global: bf16 [B*S, Dm])
    const uint32_t o2 = *reinterpret_cast<const uint32_t*>(out + (tok0 + i) * Dm + h * 64 + 2 * lane);
    const uint32_t d2 = *reinterpret_cast<const uint32_t*>(sdO + i * AB_LD + 2 * lane);
    const float Di = warp_sum(bf16_lo(o2) * bf16_lo(d2) + bf16_hi(o2) * bf16_hi(d2));
    const int nj = causal ? i + 1 : S;
    uint32_t qi[32];
    load_row(sQ + i * AB_LD, qi);
    float sc[AB_MAX_JT];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < AB_MAX_JT; ++t) {
      sc[t] = -INFINITY;
      const int j = lane + 32 * t;
      if (t < JT && j < nj && sM[j]) sc[t] = scale * dot64(qi, sK + j * AB_LD);
      mx = fmaxf(mx, sc[t]);
    }
    mx = warp_max(mx);
    float l = 0.f;
    if (mx != -INFINITY) {
#pragma unroll
      for (int t = 0; t < AB_MAX_JT; ++t) {
        sc[t] = __expf(sc[t] - mx);  // exp(-inf) = 0 for masked keys
        l += sc[t];
      }
    }
    l = warp_sum(l);
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    if (lane == 0) {
      sD[i] = Di;
      sL[i] = l > 0.f ? mx + __logf(l) : INFINITY;  // a fully masked row: exp(s - inf) = 0 in phase 2
    }
    load_row(sdO + i * AB_LD, qi);  // qi now holds dO_i
#pragma unroll
    for (int t = 0; t < AB_MAX_JT; ++t) {
      const int j = lane + 32 * t;
      if (t < JT && j < S) {
        float ds = 0.f;
        if (j < nj && sM[j] && l > 0.f) {
          const float p = sc[t] * inv_l;
          ds = p * (dot64(qi, sV + j * AB_LD) - Di);
        }
        myDS[j] = ds;
      }
    }
    __syncwarp();
    float dq0 = 0.f, dq1 = 0.f;
    for (int j = 0; j < nj; ++j) {
      const float ds = myDS[j];
      const uint32_t k2 = *reinterpret_cast<const uint32_t*>(sK + j * AB_LD + 2 * lane);
      dq0 = fmaf(ds, bf16_lo(k2), dq0);
      dq1 = fmaf(ds, bf16_hi(k2), dq1);
    }
    *reinterpret_cast<uint32_t*>(dqkv + (tok0 + i) * D3 + h * 64 + 2 * lane) = pack_bf16x2(dq0 * scale, dq1 * scale);
    __syncwarp();
  }
  __syncthreads();

  // ---------------- phase 2: key rows -> dK_j, dV_j ----------------
  for (int j = warp; j < S; j += AB_WARPS) {
    float dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
    if (sM[j]) {
      uint32_t kj[32], vj[32];
      load_row(sK + j * AB_LD, kj);
      load_row(sV + j * AB_LD, vj);
      const int i0 = causal ? j : 0;
#pragma unroll
      for (int t = 0; t < AB_MAX_JT; ++t) {
        const int i = lane + 32 * t;
        if (t < JT && i < S) {
          float p = 0.f, ds = 0.f;
          if (i >= i0) {
            p = __expf(scale * dot64(kj, sQ + i * AB_LD) - sL[i]);
            ds = p * (dot64(vj, sdO + i * AB_LD) - sD[i]);
          }
          myP[i] = p;
          myDS[i] = ds;
        }
      }
      __syncwarp();
      for (int i = i0; i < S; ++i) {
        const float p = myP[i], ds = myDS[i];
        const uint32_t g2 = *reinterpret_cast<const uint32_t*>(sdO + i * AB_LD + 2 * lane);
        const uint32_t q2 = *reinterpret_cast<const uint32_t*>(sQ + i * AB_LD + 2 * lane);
        dv0 = fmaf(p, bf16_lo(g2), dv0);
        dv1 = fmaf(p, bf16_hi(g2), dv1);
        dk0 = fmaf(ds, bf16_lo(q2), dk0);
        dk1 = fmaf(ds, bf16_hi(q2), dk1);
      }
      __syncwarp();
    }
    __nv_bfloat16* g = dqkv + (tok0 + j) * D3 + h * 64 + 2 * lane;
    *reinterpret_cast<uint32_t*>(g + Dm) = pack_bf16x2(dk0 * scale, dk1 * scale);
    *reinterpret_cast<uint32_t*>(g + 2 * Dm) = pack_bf16x2(dv0, dv1);
  }
}

}  // namespace
}  // namespace vlmclip

using namespace vlmclip;

namespace vlmclip {
int64_t attention_bwd_mma_workspace(int B, int S, int H);
int attention_bwd_mma(const void* qkv, const void* out, const void* dout, void* dqkv, const uint8_t* key_mask,
                      float* workspace, int B, int S, int H, int causal, float scale, cudaStream_t stream);
}  // namespace vlmclip

extern "C" int64_t vlmclip_attention_bwd_workspace(int B, int S, int H) {
  return vlmclip::attention_bwd_mma_workspace(B, S, H);
}

extern "C" int vlmclip_attention_bwd(const void* qkv, const void* out, const void* dout, void* dqkv,
                                     const uint8_t* key_mask, float* workspace, int B, int S, int H, int causal,
                                     float scale, void* stream) {
  VLMCLIP_CHECK_ARG(qkv && out && dout && dqkv, "attention_bwd: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && H > 0 && S > 0, "attention_bwd: bad dims B=%d S=%d H=%d", B, S, H);
  VLMCLIP_CHECK_ARG((uintptr_t)qkv % 16 == 0 && (uintptr_t)out % 16 == 0 && (uintptr_t)dout % 16 == 0 &&
                        (uintptr_t)dqkv % 16 == 0,
                    "attention_bwd: pointers must be 16-byte aligned");
  if (workspace != nullptr) {  // tensor-core kernels (attention_bwd_mma.cu)
    VLMCLIP_CHECK_ARG(S <= 512, "attention_bwd: S=%d must be <= 512", S);
    return attention_bwd_mma(qkv, out, dout, dqkv, key_mask, workspace, B, S, H, causal, scale, (cudaStream_t)stream);
  }
  // no workspace: the fp32 SIMT kernel of this file (independent implementation, kept for the parity tests)
  VLMCLIP_CHECK_ARG(S <= 32 * AB_MAX_JT, "attention_bwd (SIMT): S=%d must be <= %d", S, 32 * AB_MAX_JT);
  const int Spad = (S + 31) / 32 * 32;
  const size_t smem = (size_t)4 * S * AB_LD * 2 + (size_t)(2 + 2 * AB_WARPS) * Spad * 4 + Spad;
  static bool attr_set = false;  // benign race: the attribute is idempotent
  if (!attr_set) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  VLMCLIP_CHECK_ARG(smem <= 227 * 1024, "attention_bwd: S=%d needs %zu bytes of shared memory", S, smem);
  count_launch(1);
  attention_bwd_kernel<<<B * H, AB_WARPS * 32, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dqkv, key_mask, S,
      Spad, H, causal, scale);
  return report_cuda(cudaGetLastError(), "attention_bwd_kernel launch");
}

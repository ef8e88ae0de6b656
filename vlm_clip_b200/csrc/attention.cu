// Flash-style multi-head self-attention for the CLIP towers (head_dim = 64, S <= 512):
//   out = softmax(q k^T * scale + mask) v      HF modeling_clip.py:261-279 (eager), :318-331 (dispatch)
// Vision: no mask (S = 50 / 197 / 257).  Text: causal AND key-padding mask (HF:546-551), S <= 77.
//
// One CTA per (batch, head, query group); K and V of that head live in shared memory for the whole CTA,
// each warp owns 16 query rows and streams over 64-key blocks with an online (running max / running sum)
// softmax held in registers; quad reductions by warp shuffles.  Scores and probabilities never touch HBM.
// This is the mma.sync m16n8k16 (fp32 accumulate) variant.  The tcgen05 kernel in attention_tc.cu serves S <= 256;
// this one remains for longer sequences (ViT-L/14, S = 257) and as an independent implementation for tests.
#include <cstdlib>

#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int HD = 64;       // head dim
constexpr int KSTRIDE = 72;  // padded smem row (bf16 elements): 144 B -> conflict-free ldmatrix

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// MASKED = false: no causal / key-padding mask (vision): keys >= S are the only thing to exclude, and only the last
// 16-key group can contain them - the per-key mask loads and selects disappear from every other group.
constexpr int QK_GROUP = 4;  // 8-key tiles of S whose mma.sync chains are interleaved

template <bool MASKED>
__global__ void __launch_bounds__(256, 2)
attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                     const uint8_t* __restrict__ key_mask, int S, int H, int causal, float scale_log2e, int Spad,
                     int q_begin) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_wait();
  pdl_trigger();
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sV = sK + (size_t)Spad * KSTRIDE;
  uint8_t* sMask = reinterpret_cast<uint8_t*>(sV + (size_t)Spad * KSTRIDE);

  const int b = blockIdx.x / H;
  const int h = blockIdx.x - b * H;
  const int D = H * HD;
  const int64_t ld = 3 * (int64_t)D;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const __nv_bfloat16* base = qkv + (int64_t)b * S * ld + h * HD;

  // ---- stage K, V (and the key mask) of this head in shared memory ----
  for (int idx = threadIdx.x; idx < Spad * 8; idx += blockDim.x) {
    const int key = idx >> 3;
    const int ch = (idx & 7) * 8;
    if (key < S) {
      cp_async16(sK + key * KSTRIDE + ch, base + (int64_t)key * ld + D + ch);
      cp_async16(sV + key * KSTRIDE + ch, base + (int64_t)key * ld + 2 * D + ch);
    } else {
      *reinterpret_cast<uint4*>(sK + key * KSTRIDE + ch) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sV + key * KSTRIDE + ch) = make_uint4(0, 0, 0, 0);
    }
  }
  for (int key = threadIdx.x; key < Spad; key += blockDim.x) {
    uint8_t ok = key < S ? 1 : 0;
    if (ok && key_mask != nullptr) ok = key_mask[(int64_t)b * S + key] ? 1 : 0;
    sMask[key] = ok;
  }
  cp_async_wait_all();
  __syncthreads();

  // q_begin > 0: only the query rows from q_begin on (the tail rows of a sequence whose full 128-row tiles run on the
  // tcgen05 kernel); every warp helped to stage K / V above, warps without rows leave here
  const int q0 = q_begin + (blockIdx.y * nwarps + warp) * 16;
  if (q0 >= S) return;  // no block-wide synchronisation below this point

  const int quad = lane >> 2;
  const int tq = lane & 3;
  const int row_a = q0 + quad;
  const int row_b = row_a + 8;

  // ---- Q fragments (A operand, 16 x 64) straight from global ----
  uint32_t qf[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int col = ks * 16 + tq * 2;
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(base + (int64_t)row_a * ld + col);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(base + (int64_t)row_b * ld + col);
    qf[ks][0] = row_a < S ? __ldg(pa) : 0u;
    qf[ks][1] = row_b < S ? __ldg(pb) : 0u;
    qf[ks][2] = row_a < S ? __ldg(pa + 4) : 0u;
    qf[ks][3] = row_b < S ? __ldg(pb + 4) : 0u;
  }

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;

  int kmax = S;
  if (causal) kmax = min(S, q0 + 16);
  const int kmax16 = (kmax + 15) & ~15;

  for (int k0 = 0; k0 < kmax16; k0 += 64) {
    const int n16 = min(4, (kmax16 - k0) >> 4);  // 16-key groups in this block (warp-uniform)
    float s[8][4];
    // S = Q K^T of this block, four 8-key tiles at a time: consecutive mma.sync go to DIFFERENT accumulators (the
    // instruction stream is kept in program order, and four back-to-back updates of one accumulator wait a full tensor
    // pipe latency each); the order of the additions into every accumulator is unchanged
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int g = 0; g < 8; g += QK_GROUP) {
      uint32_t kf[QK_GROUP][4];
      const __nv_bfloat16* kp = sK + (k0 + g * 8 + (lane & 7)) * KSTRIDE + (lane >> 3) * 8;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int i = 0; i < QK_GROUP; ++i)
          if (((g + i) >> 1) < n16) ldmatrix_x4(kf[i], kp + i * 8 * KSTRIDE + half * 32);
#pragma unroll
        for (int i = 0; i < QK_GROUP; ++i)
          if (((g + i) >> 1) < n16) mma_bf16_16816(s[g + i], qf[2 * half], kf[i][0], kf[i][1]);
#pragma unroll
        for (int i = 0; i < QK_GROUP; ++i)
          if (((g + i) >> 1) < n16) mma_bf16_16816(s[g + i], qf[2 * half + 1], kf[i][2], kf[i][3]);
      }
    }
    // ---- mask + running max ----
    float bm_a = -INFINITY, bm_b = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if ((nt >> 1) < n16) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = k0 + nt * 8 + tq * 2 + e;
          if (MASKED) {
            const bool kv = sMask[key] != 0;
            const bool va = kv && (!causal || key <= row_a);
            const bool vb = kv && (!causal || key <= row_b);
            s[nt][e] = va ? s[nt][e] : -INFINITY;
            s[nt][2 + e] = vb ? s[nt][2 + e] : -INFINITY;
          } else if (k0 + nt * 8 + 8 > S) {  // warp-uniform: only the group that straddles S has invalid keys
            s[nt][e] = key < S ? s[nt][e] : -INFINITY;
            s[nt][2 + e] = key < S ? s[nt][2 + e] : -INFINITY;
          }
          bm_a = fmaxf(bm_a, s[nt][e]);
          bm_b = fmaxf(bm_b, s[nt][2 + e]);
        }
      }
    }
    bm_a = fmaxf(bm_a, __shfl_xor_sync(0xffffffffu, bm_a, 1));
    bm_a = fmaxf(bm_a, __shfl_xor_sync(0xffffffffu, bm_a, 2));
    bm_b = fmaxf(bm_b, __shfl_xor_sync(0xffffffffu, bm_b, 1));
    bm_b = fmaxf(bm_b, __shfl_xor_sync(0xffffffffu, bm_b, 2));
    const float mn_a = fmaxf(m_a, bm_a), mn_b = fmaxf(m_b, bm_b);
    const float mu_a = mn_a == -INFINITY ? 0.f : mn_a;  // fully-masked-so-far rows stay finite
    const float mu_b = mn_b == -INFINITY ? 0.f : mn_b;
    const float corr_a = fast_exp2((m_a - mu_a) * scale_log2e);
    const float corr_b = fast_exp2((m_b - mu_b) * scale_log2e);
    m_a = mn_a;
    m_b = mn_b;
    l_a *= corr_a;
    l_b *= corr_b;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      o[nd][0] *= corr_a;
      o[nd][1] *= corr_a;
      o[nd][2] *= corr_b;
      o[nd][3] *= corr_b;
    }
    const float off_a = mu_a * scale_log2e, off_b = mu_b * scale_log2e;
    // ---- P = exp2(s*c - m*c), O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk < n16) {
        uint32_t pf[4];
        float p[8];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int nt = kk * 2 + half;
          p[half * 4 + 0] = fast_exp2(fmaf(s[nt][0], scale_log2e, -off_a));
          p[half * 4 + 1] = fast_exp2(fmaf(s[nt][1], scale_log2e, -off_a));
          p[half * 4 + 2] = fast_exp2(fmaf(s[nt][2], scale_log2e, -off_b));
          p[half * 4 + 3] = fast_exp2(fmaf(s[nt][3], scale_log2e, -off_b));
        }
        l_a += p[0] + p[1] + p[4] + p[5];
        l_b += p[2] + p[3] + p[6] + p[7];
        pf[0] = pack_bf16x2(p[0], p[1]);
        pf[1] = pack_bf16x2(p[2], p[3]);
        pf[2] = pack_bf16x2(p[4], p[5]);
        pf[3] = pack_bf16x2(p[6], p[7]);
#pragma unroll
        for (int nd = 0; nd < 8; nd += 2) {
          uint32_t vf[4];
          const int j = lane >> 3;
          const __nv_bfloat16* vp =
              sV + (k0 + kk * 16 + (j & 1) * 8 + (lane & 7)) * KSTRIDE + (nd + (j >> 1)) * 8;
          ldmatrix_x4_trans(vf, vp);
          mma_bf16_16816(o[nd], pf, vf[0], vf[1]);
          mma_bf16_16816(o[nd + 1], pf, vf[2], vf[3]);
        }
      }
    }
  }

  l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
  const float inv_a = l_a > 0.f ? 1.f / l_a : 0.f;
  const float inv_b = l_b > 0.f ? 1.f / l_b : 0.f;
  __nv_bfloat16* ob = out + (int64_t)b * S * D + h * HD;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) {
    const int col = nd * 8 + tq * 2;
    if (row_a < S)
      *reinterpret_cast<uint32_t*>(ob + (int64_t)row_a * D + col) = pack_bf16x2(o[nd][0] * inv_a, o[nd][1] * inv_a);
    if (row_b < S)
      *reinterpret_cast<uint32_t*>(ob + (int64_t)row_b * D + col) = pack_bf16x2(o[nd][2] * inv_b, o[nd][3] * inv_b);
  }
}

}  // namespace
}  // namespace vlmclip

namespace vlmclip {

int attention_fwd_mma_sync_rows(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H, int causal,
                                float scale, int q_begin, cudaStream_t stream) {
  const int nblocks = (S - q_begin + 15) / 16;  // 16-row query blocks from q_begin on
  static const int max_warps = []() {  // A/B switch: query blocks (= warps) per CTA
    const char* e = getenv("VLMCLIP_ATTN_FWD_WARPS");
    return (e != nullptr && e[0] >= '2' && e[0] <= '8') ? e[0] - '0' : 8;
  }();
  const int groups = (nblocks + max_warps - 1) / max_warps;
  int qw = (nblocks + groups - 1) / groups;
  if (qw < 4) qw = 4;  // at least four warps stage K and V (warps without query rows exit after the staging)
  const int Spad = (S + 15) / 16 * 16;
  const size_t smem = (size_t)Spad * KSTRIDE * 2 * 2 + Spad;
  const bool masked = causal != 0 || key_mask != nullptr;
  using KernT = void (*)(const __nv_bfloat16*, __nv_bfloat16*, const uint8_t*, int, int, int, float, int, int);
  KernT kern = masked ? (KernT)attention_fwd_kernel<true> : (KernT)attention_fwd_kernel<false>;
  static size_t smem_set[2] = {0, 0};
  const int vi = masked ? 1 : 0;
  if (smem > smem_set[vi]) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[vi] = smem;
  }
  dim3 grid(B * H, groups);
  count_launch(1);
  return report_cuda(launch_pdl(kern, grid, dim3(qw * 32), smem, stream, 1, (const __nv_bfloat16*)qkv, (__nv_bfloat16*)out,
                                key_mask, S, H, causal, scale * 1.4426950408889634f, Spad, q_begin),
                     "attention_fwd_kernel launch");
}

// mma.sync fallback used for S > 256 (ViT-L/14, S = 257) when the caller gives no workspace for the tcgen05 key-range
// split, and for the short causal text sequences; same contract as vlmclip_attention_fwd.
int attention_fwd_mma_sync(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H, int causal,
                           float scale, cudaStream_t stream) {
  return attention_fwd_mma_sync_rows(qkv, out, key_mask, B, S, H, causal, scale, 0, stream);
}

}  // namespace vlmclip

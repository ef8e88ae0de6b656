// Tensor-core backward of the attention core (head_dim 64, S <= 512), the counterpart of attention.cu:
//   P = softmax(Q K^T c + mask),  O = P V,  D_i = dO_i . O_i
//   dS = P o (dO V^T - D),  dQ = c dS K,  dK = c dS^T Q,  dV = P^T dO
// Two kernels, both shaped like the forward (one warp owns a 16-row block, the other operand of the head lives in
// shared memory, mma.sync m16n8k16 bf16 with fp32 accumulation, scores and probabilities never leave registers):
//   attention_bwd_dq_kernel   warp = 16 QUERY rows; K, V in smem.  Pass 1 recomputes the row log-sum-exp (the forward
//                             keeps nothing but O), pass 2 forms dS block by block and accumulates dQ.  Writes
//                             (lse, D) per row to a small fp32 workspace.
//   attention_bwd_dkv_kernel  warp = 16 KEY rows; Q, dO (+ lse, D) in smem.  Works on the transposed tiles
//                             S^T = K Q^T, dP^T = V dO^T, so P^T / dS^T come out in the accumulator layout that
//                             converts to an A fragment in registers, exactly like P in the forward.
// No atomics: every output row is owned by one warp.  8 matmul-equivalents instead of the minimal 5 (Q K^T is
// recomputed in both kernels and twice in the first) buys independence from any forward by-product.
#include <cstdlib>

#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int HD = 64;
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

// A fragment (16 rows x 64 columns, 4 k-slices) of rows row_a / row_b of a row-major bf16 matrix in global memory
__device__ __forceinline__ void load_a_frag(uint32_t (&f)[4][4], const __nv_bfloat16* base, int64_t ld, int row_a,
                                            int row_b, int S, int tq) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int col = ks * 16 + tq * 2;
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(base + (int64_t)row_a * ld + col);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(base + (int64_t)row_b * ld + col);
    f[ks][0] = row_a < S ? __ldg(pa) : 0u;
    f[ks][1] = row_b < S ? __ldg(pb) : 0u;
    f[ks][2] = row_a < S ? __ldg(pa + 4) : 0u;
    f[ks][3] = row_b < S ? __ldg(pb + 4) : 0u;
  }
}

// acc[nt] (16 x 8 tile nt of a 16 x 8NT block) = A (16 x 64 fragment) . rows [r0 + nt*8, +8) of a smem matrix ^T.
// GROUP tiles at a time: consecutive mma.sync go to DIFFERENT accumulators (the instruction stream is kept in program
// order, and four back-to-back updates of one accumulator wait a full tensor pipe latency each); the order of the
// additions into every accumulator is unchanged.
template <int NT, int AB_GROUP>
__device__ __forceinline__ void block_a_bt(float (&acc)[NT][4], const uint32_t (&af)[4][4], const __nv_bfloat16* sB,
                                           int r0, int n16, int lane) {
  constexpr int GROUP = NT < AB_GROUP ? NT : AB_GROUP;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
  for (int g = 0; g < NT; g += GROUP) {
    uint32_t bf[GROUP][4];
    const __nv_bfloat16* bp = sB + (r0 + g * 8 + (lane & 7)) * KSTRIDE + (lane >> 3) * 8;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int i = 0; i < GROUP; ++i)
        if (((g + i) >> 1) < n16) ldmatrix_x4(bf[i], bp + i * 8 * KSTRIDE + half * 32);
#pragma unroll
      for (int i = 0; i < GROUP; ++i)
        if (((g + i) >> 1) < n16) mma_bf16_16816(acc[g + i], af[2 * half], bf[i][0], bf[i][1]);
#pragma unroll
      for (int i = 0; i < GROUP; ++i)
        if (((g + i) >> 1) < n16) mma_bf16_16816(acc[g + i], af[2 * half + 1], bf[i][2], bf[i][3]);
    }
  }
}

// acc (16 x 64) += T (16 x 8NT block held as accumulator-layout values t[NT][4]) . rows [r0, r0 + 8NT) of a smem matrix
template <int NT>
__device__ __forceinline__ void block_t_b(float (&acc)[8][4], const float (&t)[NT][4], const __nv_bfloat16* sB, int r0,
                                          int n16, int lane) {
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    if (kk < n16) {
      uint32_t tf[4];
      tf[0] = pack_bf16x2(t[2 * kk][0], t[2 * kk][1]);
      tf[1] = pack_bf16x2(t[2 * kk][2], t[2 * kk][3]);
      tf[2] = pack_bf16x2(t[2 * kk + 1][0], t[2 * kk + 1][1]);
      tf[3] = pack_bf16x2(t[2 * kk + 1][2], t[2 * kk + 1][3]);
#pragma unroll
      for (int nd = 0; nd < 8; nd += 2) {
        uint32_t bf[4];
        const int j = lane >> 3;
        const __nv_bfloat16* bp = sB + (r0 + kk * 16 + (j & 1) * 8 + (lane & 7)) * KSTRIDE + (nd + (j >> 1)) * 8;
        ldmatrix_x4_trans(bf, bp);
        mma_bf16_16816(acc[nd], tf, bf[0], bf[1]);
        mma_bf16_16816(acc[nd + 1], tf, bf[2], bf[3]);
      }
    }
  }
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ float dot2(uint32_t a, uint32_t b) {
  return bf16_lo(a) * bf16_lo(b) + bf16_hi(a) * bf16_hi(b);
}

// stage rows [0, Spad) of two row-major [S, 64]-slices (row strides ld0 / ld1) into padded shared memory
__device__ __forceinline__ void stage_two(__nv_bfloat16* s0, __nv_bfloat16* s1, const __nv_bfloat16* g0, int64_t ld0,
                                          const __nv_bfloat16* g1, int64_t ld1, int S, int Spad) {
  for (int idx = threadIdx.x; idx < Spad * 8; idx += blockDim.x) {
    const int r = idx >> 3;
    const int ch = (idx & 7) * 8;
    if (r < S) {
      cp_async16(s0 + r * KSTRIDE + ch, g0 + (int64_t)r * ld0 + ch);
      cp_async16(s1 + r * KSTRIDE + ch, g1 + (int64_t)r * ld1 + ch);
    } else {
      *reinterpret_cast<uint4*>(s0 + r * KSTRIDE + ch) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(s1 + r * KSTRIDE + ch) = make_uint4(0, 0, 0, 0);
    }
  }
}

// ------------------------------------------------------------------------------------------------------ dQ
// MASKED = false (vision): no causal / key-padding mask.  Padding keys (>= S) only matter for the row sum of pass 1 and
// only in the 8-key tile that straddles S: K and V rows >= S are staged as zeros, so their dS columns add nothing to dQ.
// REGS / GROUP: register cap and mma-chain interleave (see the launcher for the numbers behind 168 / 4).
template <bool MASKED, int REGS, int GROUP>
__global__ void __maxnreg__(REGS)
attention_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                        const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqkv,
                        const uint8_t* __restrict__ key_mask, float* __restrict__ ws_lse, float* __restrict__ ws_d, int S,
                        int H, int causal, float scale, int Spad) {
  extern __shared__ __align__(16) uint8_t smem[];
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
  const __nv_bfloat16* obase = out + (int64_t)b * S * D + h * HD;
  const __nv_bfloat16* dobase = dout + (int64_t)b * S * D + h * HD;
  const float c = scale * 1.4426950408889634f;

  stage_two(sK, sV, base + D, ld, base + 2 * D, ld, S, Spad);
  for (int key = threadIdx.x; key < Spad; key += blockDim.x) {
    uint8_t ok = key < S ? 1 : 0;
    if (ok && key_mask != nullptr) ok = key_mask[(int64_t)b * S + key] ? 1 : 0;
    sMask[key] = ok;
  }
  cp_async_wait_all();
  __syncthreads();

  const int q0 = (blockIdx.y * nwarps + warp) * 16;
  if (q0 >= S) return;  // no block-wide synchronisation below this point
  const int quad = lane >> 2, tq = lane & 3;
  const int row_a = q0 + quad, row_b = row_a + 8;

  uint32_t qf[4][4], gf[4][4];  // Q and dO fragments of the warp's 16 rows
  load_a_frag(qf, base, ld, row_a, row_b, S, tq);
  load_a_frag(gf, dobase, D, row_a, row_b, S, tq);
  float d_a = 0.f, d_b = 0.f;
  {
    uint32_t of[4][4];
    load_a_frag(of, obase, D, row_a, row_b, S, tq);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      d_a += dot2(gf[ks][0], of[ks][0]) + dot2(gf[ks][2], of[ks][2]);
      d_b += dot2(gf[ks][1], of[ks][1]) + dot2(gf[ks][3], of[ks][3]);
    }
    d_a = quad_sum(d_a);
    d_b = quad_sum(d_b);
  }

  int kmax = S;
  if (causal) kmax = min(S, q0 + 16);
  const int kmax16 = (kmax + 15) & ~15;

  // ---- pass 1: row log-sum-exp (log2 units) ----
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;
  for (int k0 = 0; k0 < kmax16; k0 += 64) {
    const int n16 = min(4, (kmax16 - k0) >> 4);
    float s[8][4];
    block_a_bt<8, GROUP>(s, qf, sK, k0, n16, lane);
    float bm_a = -INFINITY, bm_b = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if ((nt >> 1) < n16) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = k0 + nt * 8 + tq * 2 + e;
          if (MASKED) {
            const bool kv = sMask[key] != 0;
            s[nt][e] = (kv && (!causal || key <= row_a)) ? s[nt][e] : -INFINITY;
            s[nt][2 + e] = (kv && (!causal || key <= row_b)) ? s[nt][2 + e] : -INFINITY;
          } else if (k0 + nt * 8 + 8 > S) {  // warp-uniform: only the tile that straddles S has padding keys
            s[nt][e] = key < S ? s[nt][e] : -INFINITY;
            s[nt][2 + e] = key < S ? s[nt][2 + e] : -INFINITY;
          }
          bm_a = fmaxf(bm_a, s[nt][e]);
          bm_b = fmaxf(bm_b, s[nt][2 + e]);
        }
      }
    }
    bm_a = quad_max(bm_a);
    bm_b = quad_max(bm_b);
    const float mn_a = fmaxf(m_a, bm_a), mn_b = fmaxf(m_b, bm_b);
    const float mu_a = mn_a == -INFINITY ? 0.f : mn_a, mu_b = mn_b == -INFINITY ? 0.f : mn_b;
    l_a *= fast_exp2((m_a - mu_a) * c);
    l_b *= fast_exp2((m_b - mu_b) * c);
    m_a = mn_a;
    m_b = mn_b;
    const float off_a = mu_a * c, off_b = mu_b * c;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if ((nt >> 1) < n16) {
        l_a += fast_exp2(fmaf(s[nt][0], c, -off_a)) + fast_exp2(fmaf(s[nt][1], c, -off_a));
        l_b += fast_exp2(fmaf(s[nt][2], c, -off_b)) + fast_exp2(fmaf(s[nt][3], c, -off_b));
      }
    }
  }
  l_a = quad_sum(l_a);
  l_b = quad_sum(l_b);
  const float lse_a = l_a > 0.f ? fmaf(m_a, c, log2f(l_a)) : INFINITY;  // fully masked row: exp2(x - inf) = 0
  const float lse_b = l_b > 0.f ? fmaf(m_b, c, log2f(l_b)) : INFINITY;
  if (tq == 0) {
    float* wl = ws_lse + (int64_t)blockIdx.x * Spad;
    float* wd = ws_d + (int64_t)blockIdx.x * Spad;
    if (row_a < S) {
      wl[row_a] = lse_a;
      wd[row_a] = d_a;
    }
    if (row_b < S) {
      wl[row_b] = lse_b;
      wd[row_b] = d_b;
    }
  }

  // ---- pass 2: dS = P o (dO V^T - D), dQ += dS K ----
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  for (int k0 = 0; k0 < kmax16; k0 += 32) {  // 32 keys per block: S and dP tiles of 16 registers each
    const int n16 = min(2, (kmax16 - k0) >> 4);
    float s[4][4], dp[4][4];
    block_a_bt<4, GROUP>(s, qf, sK, k0, n16, lane);
    block_a_bt<4, GROUP>(dp, gf, sV, k0, n16, lane);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if ((nt >> 1) < n16) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float pa = fast_exp2(fmaf(s[nt][e], c, -lse_a));
          float pb = fast_exp2(fmaf(s[nt][2 + e], c, -lse_b));
          if (MASKED) {
            const int key = k0 + nt * 8 + tq * 2 + e;
            const bool kv = sMask[key] != 0;
            pa = (kv && (!causal || key <= row_a)) ? pa : 0.f;
            pb = (kv && (!causal || key <= row_b)) ? pb : 0.f;
          }
          s[nt][e] = pa * (dp[nt][e] - d_a);
          s[nt][2 + e] = pb * (dp[nt][2 + e] - d_b);
        }
      }
    }
    block_t_b<4>(dq, s, sK, k0, n16, lane);
  }
  __nv_bfloat16* gq = dqkv + (int64_t)b * S * ld + h * HD;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) {
    const int col = nd * 8 + tq * 2;
    if (row_a < S)
      *reinterpret_cast<uint32_t*>(gq + (int64_t)row_a * ld + col) = pack_bf16x2(dq[nd][0] * scale, dq[nd][1] * scale);
    if (row_b < S)
      *reinterpret_cast<uint32_t*>(gq + (int64_t)row_b * ld + col) = pack_bf16x2(dq[nd][2] * scale, dq[nd][3] * scale);
  }
}

// ------------------------------------------------------------------------------------------------------ dK, dV
// MASKED = false: nothing to select - padding queries carry lse = +inf (P = 0), padding key rows are never written
template <bool MASKED, int REGS, int GROUP>
__global__ void __maxnreg__(REGS)
attention_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                         __nv_bfloat16* __restrict__ dqkv, const uint8_t* __restrict__ key_mask,
                         const float* __restrict__ ws_lse, const float* __restrict__ ws_d, int S, int H, int causal,
                         float scale, int Spad) {
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sG = sQ + (size_t)Spad * KSTRIDE;  // dO
  float* sL = reinterpret_cast<float*>(sG + (size_t)Spad * KSTRIDE);
  float* sD = sL + Spad;

  const int b = blockIdx.x / H;
  const int h = blockIdx.x - b * H;
  const int D = H * HD;
  const int64_t ld = 3 * (int64_t)D;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const __nv_bfloat16* base = qkv + (int64_t)b * S * ld + h * HD;
  const __nv_bfloat16* dobase = dout + (int64_t)b * S * D + h * HD;
  const float c = scale * 1.4426950408889634f;

  stage_two(sQ, sG, base, ld, dobase, D, S, Spad);
  for (int i = threadIdx.x; i < Spad; i += blockDim.x) {
    sL[i] = i < S ? ws_lse[(int64_t)blockIdx.x * Spad + i] : INFINITY;  // padding queries: P = 0
    sD[i] = i < S ? ws_d[(int64_t)blockIdx.x * Spad + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  const int j0 = (blockIdx.y * nwarps + warp) * 16;
  if (j0 >= S) return;
  const int quad = lane >> 2, tq = lane & 3;
  const int row_a = j0 + quad, row_b = row_a + 8;  // the warp's KEY rows
  bool kv_a = row_a < S, kv_b = row_b < S;
  if (key_mask != nullptr) {
    kv_a = kv_a && key_mask[(int64_t)b * S + row_a] != 0;
    kv_b = kv_b && key_mask[(int64_t)b * S + row_b] != 0;
  }

  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  const int S16 = (S + 15) & ~15;
  for (int i0 = causal ? j0 : 0; i0 < S16; i0 += 32) {  // causal: queries before the warp's first key see none of them
    const int n16 = min(2, (S16 - i0) >> 4);
    float st[4][4], dpt[4][4];  // S^T and dP^T tiles: rows = keys, columns = 32 queries
    {  // the warp's K / V fragments are re-read (L1 hits) per block instead of living in 32 registers across the loop
      uint32_t kf[4][4];
      load_a_frag(kf, base + D, ld, row_a, row_b, S, tq);
      block_a_bt<4, GROUP>(st, kf, sQ, i0, n16, lane);
    }
    {
      uint32_t vf[4][4];
      load_a_frag(vf, base + 2 * D, ld, row_a, row_b, S, tq);
      block_a_bt<4, GROUP>(dpt, vf, sG, i0, n16, lane);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if ((nt >> 1) < n16) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int qi = i0 + nt * 8 + tq * 2 + e;
          const float lse = sL[qi], dd = sD[qi];
          float pa = fast_exp2(fmaf(st[nt][e], c, -lse));
          float pb = fast_exp2(fmaf(st[nt][2 + e], c, -lse));
          if (MASKED) {
            pa = (kv_a && (!causal || row_a <= qi)) ? pa : 0.f;
            pb = (kv_b && (!causal || row_b <= qi)) ? pb : 0.f;
          }
          st[nt][e] = pa;
          st[nt][2 + e] = pb;
          dpt[nt][e] = pa * (dpt[nt][e] - dd);
          dpt[nt][2 + e] = pb * (dpt[nt][2 + e] - dd);
        }
      }
    }
    block_t_b<4>(dv, st, sG, i0, n16, lane);
    block_t_b<4>(dk, dpt, sQ, i0, n16, lane);
  }
  __nv_bfloat16* gk = dqkv + (int64_t)b * S * ld + h * HD + D;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) {
    const int col = nd * 8 + tq * 2;
    if (row_a < S) {
      *reinterpret_cast<uint32_t*>(gk + (int64_t)row_a * ld + col) = pack_bf16x2(dk[nd][0] * scale, dk[nd][1] * scale);
      *reinterpret_cast<uint32_t*>(gk + (int64_t)row_a * ld + D + col) = pack_bf16x2(dv[nd][0], dv[nd][1]);
    }
    if (row_b < S) {
      *reinterpret_cast<uint32_t*>(gk + (int64_t)row_b * ld + col) = pack_bf16x2(dk[nd][2] * scale, dk[nd][3] * scale);
      *reinterpret_cast<uint32_t*>(gk + (int64_t)row_b * ld + D + col) = pack_bf16x2(dv[nd][2], dv[nd][3]);
    }
  }
}

}  // namespace
}  // namespace vlmclip

namespace vlmclip {

int64_t attention_bwd_mma_workspace(int B, int S, int H) {
  const int Spad = (S + 15) / 16 * 16;
  return 2 * (int64_t)B * H * Spad;
}

int attention_bwd_mma(const void* qkv, const void* out, const void* dout, void* dqkv, const uint8_t* key_mask,
                      float* workspace, int B, int S, int H, int causal, float scale, cudaStream_t stream) {
  // Register file: 16 384 registers per SM sub-partition.  At 168 registers per thread (no spills; the 128-register
  // build that would allow one 13-warp CTA per ViT-B/16 head spills and measured 1 079 us against 949) three warps fit a
  // sub-partition, twelve an SM: CTAs of 4 warps (three per SM) or 6 warps (two per SM) fill it, 5 or 7 do not.  Every
  // CTA stages the whole head, so fewer, fuller CTAs win when both shapes waste the same share of warp slots.
  // Measured (B200, us per call, dQ + dK/dV): S = 197, H = 12, B = 256: 949 with 4 warps, 991 with 6 (3 x 5), 1 069 with 3;
  // S = 77 causal, H = 8: 114 / 134 / 112; S = 257, H = 16, B = 64: 626 / 491 / 705.
  // VLMCLIP_ATTN_BWD_WARPS = 3..6 overrides the choice (A/B switch, read once).
  static const int warps_env = []() {
    const char* e = getenv("VLMCLIP_ATTN_BWD_WARPS");
    return (e != nullptr && e[0] >= '3' && e[0] <= '6') ? e[0] - '0' : 0;
  }();
  const int nblocks = (S + 15) / 16;
  int max_warps = warps_env;
  if (max_warps == 0) {
    double best = -1.0;
    for (int w : {4, 6}) {  // filled share of the twelve resident warp slots; ties go to the smaller CTA
      const int g = (nblocks + w - 1) / w, q = (nblocks + g - 1) / g;
      const double fill = (double)nblocks / (g * q) * ((12 / q) * q / 12.0);
      if (fill > best + 1e-9) best = fill, max_warps = w;
    }
  }
  const int groups = (nblocks + max_warps - 1) / max_warps;
  const int qw = (nblocks + groups - 1) / groups;
  const int Spad = nblocks * 16;
  float* ws_lse = workspace;
  float* ws_d = workspace + (int64_t)B * H * Spad;
  const size_t smem_q = (size_t)Spad * KSTRIDE * 2 * 2 + Spad;
  const size_t smem_kv = (size_t)Spad * KSTRIDE * 2 * 2 + (size_t)Spad * 8;
  const bool masked = causal != 0 || key_mask != nullptr;
  using DqT = void (*)(const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, __nv_bfloat16*, const uint8_t*,
                       float*, float*, int, int, int, float, int);
  using DkvT = void (*)(const __nv_bfloat16*, const __nv_bfloat16*, __nv_bfloat16*, const uint8_t*, const float*,
                        const float*, int, int, int, float, int);
  const DqT dq_kern = masked ? attention_bwd_dq_kernel<true, 168, 4> : attention_bwd_dq_kernel<false, 168, 4>;
  const DkvT dkv_kern = masked ? attention_bwd_dkv_kernel<true, 168, 4> : attention_bwd_dkv_kernel<false, 168, 4>;
  static size_t set_q[2] = {0, 0}, set_kv[2] = {0, 0};
  if (smem_q > set_q[masked]) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(dq_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
    set_q[masked] = smem_q;
  }
  if (smem_kv > set_kv[masked]) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(dkv_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));
    set_kv[masked] = smem_kv;
  }
  dim3 grid(B * H, groups);
  count_launch(2);
  dq_kern<<<grid, qw * 32, smem_q, stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)out,
                                             (const __nv_bfloat16*)dout, (__nv_bfloat16*)dqkv, key_mask, ws_lse, ws_d, S, H,
                                             causal, scale, Spad);
  VLMCLIP_CUDA(cudaGetLastError());
  dkv_kern<<<grid, qw * 32, smem_kv, stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dqkv,
                                               key_mask, ws_lse, ws_d, S, H, causal, scale, Spad);
  return report_cuda(cudaGetLastError(), "attention_bwd_mma launch");
}

}  // namespace vlmclip

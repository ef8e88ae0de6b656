// 64x64 output tile, fp32 SIMT, 256 threads (16x16), 4x4 outputs per thread, K chunks of 16 through shared
// memory.  Operands are supplied by element loaders so the same routine serves the small "reduce over the
// batch" GEMMs of the trainable path (adapter weight gradients, similarity logits, loss gradients), all of
// which are tiny next to the frozen towers and must stay in fp32 for loss parity (1e-4).
#pragma once
#include "common.cuh"

namespace vlmclip {

constexpr int TF_TILE = 64;
constexpr int TF_KC = 16;

// C[m0+i][n0+j] = sum_k A(m0+i, k) * B(n0+j, k);  loadA(m,k)/loadB(n,k) must return 0 outside bounds.
// a_kfast / b_kfast: true when consecutive k are contiguous in memory (choose the coalesced thread mapping).
template <class LoadA, class LoadB, class Store>
__device__ __forceinline__ void tile_gemm_f32(int K, int m0, int n0, bool a_kfast, bool b_kfast, LoadA loadA,
                                              LoadB loadB, Store store) {
  __shared__ float As[TF_KC][TF_TILE + 4];
  __shared__ float Bs[TF_KC][TF_TILE + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += TF_KC) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;  // 0..1023
      int mm, kk;
      if (a_kfast) {
        kk = idx & 15;
        mm = idx >> 4;
      } else {
        mm = idx & 63;
        kk = idx >> 6;
      }
      As[kk][mm] = loadA(m0 + mm, k0 + kk);
      int nn, kb;
      if (b_kfast) {
        kb = idx & 15;
        nn = idx >> 4;
      } else {
        nn = idx & 63;
        kb = idx >> 6;
      }
      Bs[kb][nn] = loadB(n0 + nn, k0 + kb);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TF_KC; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) store(m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j]);
}

}  // namespace vlmclip

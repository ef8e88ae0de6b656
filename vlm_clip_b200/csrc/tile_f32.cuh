// TF_TILE x TF_TILE output tile, fp32 SIMT, (TF_TILE/4)^2 threads, 4x4 outputs per thread, K chunks of 16 through
// shared memory.  Operands are supplied by element loaders so the same routine serves the small "reduce over the
// batch" GEMMs of the trainable path (projections, adapter weight gradients, similarity logits, loss gradients), all
// of which are tiny next to the frozen towers and must stay in fp32 for loss parity (1e-4).
// These problems have M = batch = a few hundred rows: 32 x 32 tiles of 64 threads give 100+ CTAs, and the next
// K chunk is fetched into registers while the current one is multiplied (the kernels were L2-latency bound:
// load -> sync -> multiply -> sync exposed ~700 cycles per 16-deep chunk).
#pragma once
#include "common.cuh"

namespace vlmclip {

constexpr int TF_TILE = 32;
constexpr int TF_KC = 16;
constexpr int TF_THREADS = (TF_TILE / 4) * (TF_TILE / 4);       // 64
constexpr int TF_LOADS = TF_TILE * TF_KC / TF_THREADS;          // elements of A (and of B) per thread per chunk: 8

// C[m0+i][n0+j] = sum_k A(m0+i, k) * B(n0+j, k);  loadA(m,k)/loadB(n,k) must return 0 outside bounds.
// a_kfast / b_kfast: true when consecutive k are contiguous in memory (choose the coalesced thread mapping).
template <class LoadA, class LoadB, class Store>
__device__ __forceinline__ void tile_gemm_f32(int K, int m0, int n0, bool a_kfast, bool b_kfast, LoadA loadA,
                                              LoadB loadB, Store store) {
  __shared__ float As[TF_KC][TF_TILE + 4];
  __shared__ float Bs[TF_KC][TF_TILE + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (TF_TILE / 4), ty = tid / (TF_TILE / 4);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // element (mm, kk) of the chunk handled by load slot e of this thread
  auto coord = [&](int e, bool kfast, int& mm, int& kk) {
    const int idx = tid + e * TF_THREADS;
    if (kfast) {
      kk = idx % TF_KC;
      mm = idx / TF_KC;
    } else {
      mm = idx % TF_TILE;
      kk = idx / TF_TILE;
    }
  };
  float ra[TF_LOADS], rb[TF_LOADS];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < TF_LOADS; ++e) {
      int mm, kk;
      coord(e, a_kfast, mm, kk);
      ra[e] = loadA(m0 + mm, k0 + kk);
      coord(e, b_kfast, mm, kk);
      rb[e] = loadB(n0 + mm, k0 + kk);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += TF_KC) {
#pragma unroll
    for (int e = 0; e < TF_LOADS; ++e) {
      int mm, kk;
      coord(e, a_kfast, mm, kk);
      As[kk][mm] = ra[e];
      coord(e, b_kfast, mm, kk);
      Bs[kk][mm] = rb[e];
    }
    __syncthreads();
    if (k0 + TF_KC < K) fetch(k0 + TF_KC);  // in flight while this chunk is multiplied
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

// Symmetric InfoNCE loss of Track M and its gradient, on STRIPS of the logit matrix (model_m.py:146-171):
//
//     t^ = t/|t|, i^ = i/|i|, Z = s t^ i^T, loss = (CE(Z, arange) + CE(Z^T, arange)) / 2
//     G = [(softmax_rows(Z) - I) + (softmax_cols(Z) - I)] / 2N,  dt^ = s G i^,  di^ = s G^T t^,
//     dx = (dx^ - x^ (x^ . dx^)) / |x|
//
// A rank that owns rows [row0, row0 + nloc) of the global batch needs two strips only:
//     problem 0:  Zt[a, b] = s t^[row0 + a] . i^[b]      (its text rows against every image)
//     problem 1:  Zi[a, b] = s i^[row0 + a] . t^[b]      (its image rows against every text = its COLUMNS of Z)
// Row log-sum-exps of Zt are the text-side LSEs of its rows, row LSEs of Zi the image-side (column) LSEs of its
// columns; both are complete on the rank.  The gradient of local row a of problem p is
//     dA[a] = (s / 2N) sum_b [exp(Zp[a,b] - lseA[a]) + exp(Zp[a,b] - lseB[b]) - 2 delta] B[b]
// which needs the OTHER side's LSE for every b: under data parallelism the 2 nloc LSE values (and the rank's share of
// the loss) are all-gathered between the two kernels; a single process has them already.
//
// Three launches (the previous path: eight, on the full N x N matrix on every rank):
//   clip_norm2_kernel        both feature matrices -> unit rows + 1/norm
//   clip_strip_lse_kernel    128 x 128 fp32 register tiles (8 x 8 outputs per thread) over (row block, column split,
//                            problem); running (max, sum) of every row in registers, merged across the 16 threads of a
//                            row with warp shuffles; the last-arriving CTA of a row block merges the column splits in
//                            fixed order and the very last CTA adds up the loss: deterministic, no floating-point atomics
//   clip_strip_grad_kernel   dA = G B with G formed on the fly from the stored strip, the reduction over the N columns
//                            split across CTAs; the last-arriving CTA of a row block adds the splits in index order and
//                            applies the normalisation backward, which needs the full-row dot product
// fp32 SIMT on purpose: the loss must match the reference's fp32 value to 1e-4 at a logit scale of 100, i.e. cosines
// to ~1e-7; the whole path is ~13 GFLOP at N = 4096 (0.02 % of that step).
#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int CL_TILE = 128;     // CTA tile: 128 x 128 outputs
constexpr int CL_KC = 16;        // reduction chunk staged through shared memory
constexpr int CL_THREADS = 256;  // 16 x 16 threads, 8 x 8 outputs each (two 4-row groups x two 4-column groups, 64 apart)
constexpr int CL_PAD = 4;

__global__ void __launch_bounds__(256)
clip_norm2_kernel(const float* __restrict__ txt, const float* __restrict__ img, float* __restrict__ txt_n,
                  float* __restrict__ img_n, float* __restrict__ inv_t, float* __restrict__ inv_i, int N, int P) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  const bool is_img = blockIdx.y == 1;
  const float* xr = (is_img ? img : txt) + (int64_t)row * P;
  float* yr = (is_img ? img_n : txt_n) + (int64_t)row * P;
  float q = 0.f;
  for (int c = lane; c < P; c += 32) q = fmaf(xr[c], xr[c], q);
  const float inv = 1.f / sqrtf(warp_sum(q));
  for (int c = lane; c < P; c += 32) yr[c] = xr[c] * inv;
  if (lane == 0) (is_img ? inv_i : inv_t)[row] = inv;
}

// 128 x 128 x K fp32 tile product.  Thread (tx, ty) = (tid & 15, tid >> 4) owns rows {ty*4 + i, 64 + ty*4 + i} and
// columns {tx*4 + j, 64 + tx*4 + j}, i, j < 4 (acc[8][8]): per k four 16-byte shared loads feed 64 FMAs, and the 16
// threads of a half-warp share their rows, so row reductions are shuffles.  K runs in chunks of 16 through shared memory
// (layout [k][m]); the next chunk is fetched into registers while the current one is multiplied.
// fetchA / fetchB(k0, regs[8]) return the eight elements this thread stages for chunk k0, storeA / storeB put them away.
struct TileSmem {
  float As[CL_KC][CL_TILE + CL_PAD];
  float Bs[CL_KC][CL_TILE + CL_PAD];
};

__device__ __forceinline__ int tile_row(int ty, int i) { return (i < 4 ? 0 : 60) + ty * 4 + i; }  // i >= 4: 64 + ty*4 + i - 4
__device__ __forceinline__ int tile_col(int tx, int j) { return (j < 4 ? 0 : 60) + tx * 4 + j; }

template <class FetchA, class FetchB, class StoreA, class StoreB>
__device__ __forceinline__ void tile128_mainloop(TileSmem& sm, int k_begin, int k_end, float (&acc)[8][8], FetchA fetchA,
                                                 FetchB fetchB, StoreA storeA, StoreB storeB) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float ra[8], rb[8];
  fetchA(k_begin, ra);
  fetchB(k_begin, rb);
  for (int k0 = k_begin; k0 < k_end; k0 += CL_KC) {
    storeA(ra);
    storeB(rb);
    __syncthreads();
    if (k0 + CL_KC < k_end) {
      fetchA(k0 + CL_KC, ra);
      fetchB(k0 + CL_KC, rb);
    }
#pragma unroll
    for (int kk = 0; kk < CL_KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sm.As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sm.As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sm.Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sm.Bs[kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// staging of a k-contiguous operand (row-major [rows, K]): thread t takes a float4 along k of rows t/4 and 64 + t/4
__device__ __forceinline__ void stage_kfast(float (*S)[CL_TILE + CL_PAD], const float (&r)[8]) {
  const int lrow = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    S[lk + j][lrow] = r[j];
    S[lk + j][64 + lrow] = r[4 + j];
  }
}

struct StripParams {
  const float* txt_n;  // [N, P] unit rows
  const float* img_n;
  float* z[2];         // strips [nloc, ldz]: z[0] = Zt, z[1] = Zi
  float* logits;       // optional [N, N]: a copy of Zt when the rank owns every row (logits_per_text)
  float2* part;        // [2][nsplit][nloc] running (max, sum) of a row over one column split
  float* lse[2];       // [nloc] each
  float* diag;         // [2][nloc] Zp[a, row0 + a]
  float* rb_loss;      // [2][row blocks]
  float* loss;         // [1]: this rank's share (1/2N) sum_a [(lse_t[a] - d) + (lse_i[a] - d)]
  unsigned* counters;  // [2 * row blocks + 1], zero on entry, zero again on exit
  float s;
  int N, P, row0, nloc, ldz, nsplit, cols_per_split;
};

__global__ void __launch_bounds__(CL_THREADS, 2)
clip_strip_lse_kernel(const StripParams p) {
  __shared__ TileSmem sm;
  __shared__ unsigned s_last;
  const int prob = blockIdx.z;
  const int rb = blockIdx.x;
  const int split = blockIdx.y;
  const int a0 = rb * CL_TILE;
  const float* A = (prob == 0 ? p.txt_n : p.img_n) + (int64_t)p.row0 * p.P;  // local rows
  const float* B = prob == 0 ? p.img_n : p.txt_n;                           // every row
  float* Z = prob == 0 ? p.z[0] : p.z[1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // staging: float4 along k of rows lrow and 64 + lrow
  const int c_begin = split * p.cols_per_split;
  const int c_end = min(p.N, c_begin + p.cols_per_split);

  float run_m[8], run_l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    run_m[i] = -INFINITY;
    run_l[i] = 0.f;
  }
  for (int b0 = c_begin; b0 < c_end; b0 += CL_TILE) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    auto fetch = [&](const float* base, int r0, int rlim, int k0, float (&r)[8]) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = r0 + lrow + 64 * h;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rlim && k0 + lk < p.P) v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)row * p.P + k0 + lk));
        r[4 * h + 0] = v.x; r[4 * h + 1] = v.y; r[4 * h + 2] = v.z; r[4 * h + 3] = v.w;
      }
    };
    tile128_mainloop(
        sm, 0, p.P, acc, [&](int k0, float (&r)[8]) { fetch(A, a0, p.nloc, k0, r); },
        [&](int k0, float (&r)[8]) { fetch(B, b0, c_end, k0, r); }, [&](const float (&r)[8]) { stage_kfast(sm.As, r); },
        [&](const float (&r)[8]) { stage_kfast(sm.Bs, r); });
    // epilogue of the tile: scale, store the strip, diagonal, running log-sum-exp of the thread's eight rows
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int a = a0 + tile_row(ty, i);
      if (a >= p.nloc) continue;
      float z[8];
      float tm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int b = b0 + tile_col(tx, j);
        z[j] = p.s * acc[i][j];
        if (b < c_end) {
          Z[(int64_t)a * p.ldz + b] = z[j];
          if (prob == 0 && p.logits != nullptr) p.logits[(int64_t)(p.row0 + a) * p.N + b] = z[j];
          if (b == p.row0 + a) p.diag[prob * p.nloc + a] = z[j];
          tm = fmaxf(tm, z[j]);
        } else {
          z[j] = -INFINITY;
        }
      }
      if (tm > run_m[i]) {
        run_l[i] *= __expf(run_m[i] - tm);  // exp(-inf) = 0 on the first visible tile
        run_m[i] = tm;
      }
      if (run_m[i] > -INFINITY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) run_l[i] += __expf(z[j] - run_m[i]);
      }
    }
  }
  // merge the 16 threads that share a row (the tx dimension = one half-warp): (max, sum) pairs through shuffles
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float m = run_m[i], l = run_l[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
      const float mm = fmaxf(m, m2);
      l = (m > -INFINITY ? l * __expf(m - mm) : 0.f) + (m2 > -INFINITY ? l2 * __expf(m2 - mm) : 0.f);
      m = mm;
    }
    const int a = a0 + tile_row(ty, i);
    if (tx == 0 && a < p.nloc) p.part[((int64_t)prob * p.nsplit + split) * p.nloc + a] = make_float2(m, l);
  }
  // ---- last-arriving CTA of this (problem, row block): merge the column splits in index order ----
  __threadfence();
  __syncthreads();
  const int nrb = gridDim.x;
  if (tid == 0) s_last = atomicAdd(&p.counters[prob * nrb + rb], 1u) == (unsigned)(p.nsplit - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last == 0u) return;
  __threadfence();
  float* red = &sm.As[0][0];  // 128 floats of scratch
  if (tid < CL_TILE) {
    const int a = a0 + tid;
    float term = 0.f;
    if (a < p.nloc) {
      float m = -INFINITY, l = 0.f;
      for (int sidx = 0; sidx < p.nsplit; ++sidx) {
        const float2 q = __ldcg(&p.part[((int64_t)prob * p.nsplit + sidx) * p.nloc + a]);
        if (q.x == -INFINITY) continue;
        const float mm = fmaxf(m, q.x);
        l = (m > -INFINITY ? l * __expf(m - mm) : 0.f) + q.y * __expf(q.x - mm);
        m = mm;
      }
      const float lse = m + logf(l);
      (prob == 0 ? p.lse[0] : p.lse[1])[a] = lse;
      term = lse - __ldcg(&p.diag[prob * p.nloc + a]);
    }
    red[tid] = term;
  }
  __syncthreads();
  if (tid == 0) {
    float sum = 0.f;
    for (int i = 0; i < CL_TILE; ++i) sum += red[i];
    p.rb_loss[prob * nrb + rb] = sum;
    p.counters[prob * nrb + rb] = 0u;  // ready for the next launch
    __threadfence();
    // ---- the very last row block adds up the loss in index order ----
    if (atomicAdd(&p.counters[2 * nrb], 1u) == (unsigned)(2 * nrb - 1)) {
      __threadfence();
      float tot = 0.f;
      for (int i = 0; i < 2 * nrb; ++i) tot += __ldcg(&p.rb_loss[i]);
      p.loss[0] = tot / (2.f * (float)p.N);
      p.counters[2 * nrb] = 0u;
    }
  }
}

struct GradParams {
  const float* txt_n;
  const float* img_n;
  const float* inv_t;  // [N] 1 / |t|
  const float* inv_i;
  const float* z[2];      // strips [strip rows, ldz]; local row a is strip row z_row_off + a
  // LSEs of every global row, as the all-gather of each rank's [lse_t(rows_per_rank) | lse_i(rows_per_rank) | ...] block
  // leaves them: text-side LSE of global row r at lse_all[(r / rows_per_rank) * lse_stride + r % rows_per_rank], the
  // image-side one rows_per_rank further
  const float* lse_all;
  int lse_stride, rows_per_rank;
  float* dn;              // [ksplit][2][nloc, P] partial gradients w.r.t. the unit rows (one plane per reduction split)
  float* dotp;            // [2][nloc][ptiles] per-P-tile partial dot products (unit row . summed gradient)
  float* d[2];            // [nloc, P] outputs: d_txt, d_img
  unsigned* counters;     // [2 * row blocks] level 2, then [2 * row blocks * ptiles] level 1; zero on entry and on exit
  float s;
  int N, P, row0, nloc, ldz, z_row_off;
  int ptiles, ksplit, k_per_split;  // gridDim.y = ptiles * ksplit
};

__device__ __forceinline__ float lse_of(const GradParams& p, int side, int r) {
  const int w = r / p.rows_per_rank;
  return __ldg(p.lse_all + (int64_t)w * p.lse_stride + side * p.rows_per_rank + (r - w * p.rows_per_rank));
}

__global__ void __launch_bounds__(CL_THREADS, 2)
clip_strip_grad_kernel(const GradParams p) {
  __shared__ TileSmem sm;
  __shared__ unsigned s_last;
  const int prob = blockIdx.z;
  const int rb = blockIdx.x;
  const int a0 = rb * CL_TILE;
  const int ptile = blockIdx.y % p.ptiles, ks = blockIdx.y / p.ptiles;
  const int c0 = ptile * CL_TILE;  // columns of P
  const int k_begin = ks * p.k_per_split, k_end = min(p.N, k_begin + p.k_per_split);  // this CTA's share of the reduction
  const float* Z = (prob == 0 ? p.z[0] : p.z[1]) + (int64_t)p.z_row_off * p.ldz;
  const int sideA = prob, sideB = 1 - prob;  // problem 0: own text-side LSE, the images' LSE for the columns; 1: swapped
  const float* B = prob == 0 ? p.img_n : p.txt_n;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // A staging: float4 along b (the reduction index) of rows lrow, 64 + lrow
  const int bk = tid >> 5, bn = (tid & 31) * 4;   // B staging: float4 along the P columns of reduction rows bk, 8 + bk
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float my_lse[2];
  bool a_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    a_ok[h] = a0 + lrow + 64 * h < p.nloc;
    my_lse[h] = a_ok[h] ? lse_of(p, sideA, p.row0 + a0 + lrow + 64 * h) : 0.f;
  }
  tile128_mainloop(
      sm, k_begin, k_end, acc,
      [&](int k0, float (&r)[8]) {  // G[a, b] for b = k0 + lk .. + 3, rows lrow and 64 + lrow
        const int b = k0 + lk;
        float lb[4];
        {
          // one division for the four consecutive columns (they share a rank block unless they straddle its end)
          const int w = b / p.rows_per_rank, off = b - w * p.rows_per_rank;
          const float* lp = p.lse_all + (int64_t)w * p.lse_stride + sideB * p.rows_per_rank + off;
          if (off + 3 < p.rows_per_rank && b + 3 < k_end) {
#pragma unroll
            for (int j = 0; j < 4; ++j) lb[j] = __ldg(lp + j);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) lb[j] = b + j < k_end ? lse_of(p, sideB, b + j) : 0.f;
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          const int a = a0 + lrow + 64 * h;
          if (a_ok[h] && b < k_end) v = __ldg(reinterpret_cast<const float4*>(Z + (int64_t)a * p.ldz + b));  // ldz % 4 == 0
          const float zz[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float g = 0.f;
            if (a_ok[h] && b + j < k_end) {
              g = __expf(zz[j] - my_lse[h]) + __expf(zz[j] - lb[j]);
              if (b + j == p.row0 + a) g -= 2.f;
            }
            r[4 * h + j] = g;
          }
        }
      },
      [&](int k0, float (&r)[8]) {  // B[b, c0 + bn .. + 3] for b = k0 + bk and k0 + 8 + bk
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          const int b = k0 + bk + 8 * h;
          if (b < k_end && c0 + bn < p.P) v = __ldg(reinterpret_cast<const float4*>(B + (int64_t)b * p.P + c0 + bn));
          r[4 * h + 0] = v.x; r[4 * h + 1] = v.y; r[4 * h + 2] = v.z; r[4 * h + 3] = v.w;
        }
      },
      [&](const float (&r)[8]) { stage_kfast(sm.As, r); },
      [&](const float (&r)[8]) {
        *reinterpret_cast<float4*>(&sm.Bs[bk][bn]) = make_float4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<float4*>(&sm.Bs[8 + bk][bn]) = make_float4(r[4], r[5], r[6], r[7]);
      });
  const float coef = p.s / (2.f * (float)p.N);
  const int64_t plane = (int64_t)p.nloc * p.P;
  float* dn = p.dn + ((int64_t)ks * 2 + prob) * plane;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int a = a0 + tile_row(ty, i);
    if (a >= p.nloc) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + 64 * h + tx * 4;
      if (c < p.P)
        *reinterpret_cast<float4*>(dn + (int64_t)a * p.P + c) = make_float4(
            coef * acc[i][4 * h + 0], coef * acc[i][4 * h + 1], coef * acc[i][4 * h + 2], coef * acc[i][4 * h + 3]);
    }
  }
  // ---- two-level last-arriver reduction (fixed summation orders: deterministic) ----
  // level 1, per (problem, row block, P tile): the CTA that completes the tile's last reduction split adds the splits in
  // index order into split 0's plane and leaves each row's partial dot product with its unit row;
  // level 2, per (problem, row block): the CTA that completes the block's last P tile goes back through the
  // normalisation, which needs the whole-row dot product.
  __threadfence();
  __syncthreads();
  const int nrb = gridDim.x;
  unsigned* cnt1 = p.counters + 2 * nrb + ((prob * nrb + rb) * p.ptiles + ptile);
  unsigned* cnt2 = p.counters + prob * nrb + rb;
  if (tid == 0) s_last = atomicAdd(cnt1, 1u) == (unsigned)(p.ksplit - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last == 0u) return;
  __threadfence();
  const float* An = (prob == 0 ? p.txt_n : p.img_n) + (int64_t)p.row0 * p.P;
  const float* inv = (prob == 0 ? p.inv_t : p.inv_i) + p.row0;
  float* out = prob == 0 ? p.d[0] : p.d[1];
  float* sum0 = p.dn + (int64_t)prob * plane;  // split 0's plane of this problem
  float* dotp = p.dotp + (int64_t)prob * p.nloc * p.ptiles;
  const int warp = tid >> 5, lane = tid & 31;
  {
    const int c = c0 + lane * 4;  // 32 lanes x 4 columns = the tile's 128 columns
    for (int r = warp; r < CL_TILE; r += CL_THREADS / 32) {
      const int a = a0 + r;
      if (a >= p.nloc) break;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      float pd = 0.f;
      if (c < p.P) {
        for (int k = 0; k < p.ksplit; ++k) {
          const float4 q = __ldcg(reinterpret_cast<const float4*>(p.dn + ((int64_t)k * 2 + prob) * plane + (int64_t)a * p.P + c));
          v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        }
        *reinterpret_cast<float4*>(sum0 + (int64_t)a * p.P + c) = v;
        const float4 x = *reinterpret_cast<const float4*>(An + (int64_t)a * p.P + c);
        pd = x.x * v.x + x.y * v.y + x.z * v.z + x.w * v.w;
      }
      pd = warp_sum(pd);
      if (lane == 0) dotp[(int64_t)a * p.ptiles + ptile] = pd;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    *cnt1 = 0u;
    s_last = atomicAdd(cnt2, 1u) == (unsigned)(p.ptiles - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last == 0u) return;
  __threadfence();
  for (int r = warp; r < CL_TILE; r += CL_THREADS / 32) {
    const int a = a0 + r;
    if (a >= p.nloc) break;
    float dot = 0.f;
    for (int t = 0; t < p.ptiles; ++t) dot += __ldcg(dotp + (int64_t)a * p.ptiles + t);  // index order
    const float iv = inv[a];
    for (int c = lane * 4; c < p.P; c += 128) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(sum0 + (int64_t)a * p.P + c));
      const float4 x = *reinterpret_cast<const float4*>(An + (int64_t)a * p.P + c);
      *reinterpret_cast<float4*>(out + (int64_t)a * p.P + c) =
          make_float4((v.x - x.x * dot) * iv, (v.y - x.y * dot) * iv, (v.z - x.z * dot) * iv, (v.w - x.w * dot) * iv);
    }
  }
  if (tid == 0) *cnt2 = 0u;
}

// dL/d(log scale) of this rank's rows and columns (full fine-tune only): sum over the two strips of (P - delta) Z / 2N
__global__ void __launch_bounds__(256)
clip_strip_dscale_kernel(const float* __restrict__ zt, const float* __restrict__ zi, const float* __restrict__ lse_t,
                         const float* __restrict__ lse_i, float* __restrict__ out, int N, int row0, int nloc, int ldz) {
  __shared__ float part[256];
  float s = 0.f;
  const int64_t total = (int64_t)nloc * N;
  for (int64_t idx = threadIdx.x; idx < total; idx += 256) {
    const int a = (int)(idx / N), b = (int)(idx % N);
    const float z0 = zt[(int64_t)a * ldz + b], z1 = zi[(int64_t)a * ldz + b];
    float g0 = __expf(z0 - lse_t[a]), g1 = __expf(z1 - lse_i[a]);
    if (b == row0 + a) {
      g0 -= 1.f;
      g1 -= 1.f;
    }
    s += g0 * z0 + g1 * z1;
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = part[0] / (2.f * (float)N);
}

// y = a * s[0]: the upstream gradient of the scalar loss (a device scalar) applied to a stored gradient
__global__ void __launch_bounds__(256)
scale_f32_kernel(const float* __restrict__ a, const float* __restrict__ sc, float* __restrict__ y, int64_t n) {
  const float f = __ldg(sc);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a[i] * f;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t align4(int64_t v) { return (v + 3) & ~(int64_t)3; }

// column splits of the strip kernel: enough CTAs for two per SM, at least one 128-column tile each
inline int pick_nsplit(int N, int nloc) {
  const int rbs = ceil_div(nloc, CL_TILE) * 2;
  int want = (2 * sm_count()) / rbs;  // floor: one resident wave
  const int max_split = ceil_div(N, CL_TILE);
  if (want > max_split) want = max_split;
  return want < 1 ? 1 : want;
}

// reduction splits of the gradient kernel (its output tiles alone are few: nloc/128 x P/128 x 2)
struct GradSplit {
  int ptiles, ksplit, k_per_split;
};
inline GradSplit grad_split(int N, int P, int nloc) {
  GradSplit g;
  g.ptiles = ceil_div(P, CL_TILE);
  const int base = ceil_div(nloc, CL_TILE) * g.ptiles * 2;
  int want = (2 * sm_count()) / base;  // floor: all CTAs resident at once (two per SM), no second partial wave
  const int max_split = ceil_div(N, 2 * CL_TILE);  // at least 256 reduction rows per split
  if (want > max_split) want = max_split;
  if (want < 1) want = 1;
  g.k_per_split = ceil_div(ceil_div(N, want), CL_KC) * CL_KC;
  g.ksplit = ceil_div(N, g.k_per_split);
  return g;
}

struct FwdLayout {
  int ldz, nsplit, nrb, cols_per_split;
  int64_t off_inv_t, off_inv_i, off_zt, off_zi, off_part, off_diag, off_rb, total;
};

FwdLayout fwd_layout(int N, int P, int nloc) {
  FwdLayout L;
  L.ldz = (int)align4(N);
  L.cols_per_split = ceil_div(ceil_div(N, pick_nsplit(N, nloc)), CL_TILE) * CL_TILE;  // whole 64-column tiles
  L.nsplit = ceil_div(N, L.cols_per_split);                                            // no empty split
  L.nrb = ceil_div(nloc, CL_TILE);
  int64_t o = 0;
  L.off_inv_t = o; o += align4(N);
  L.off_inv_i = o; o += align4(N);
  L.off_zt = o; o += (int64_t)nloc * L.ldz;
  L.off_zi = o; o += (int64_t)nloc * L.ldz;
  L.off_part = o; o += align4((int64_t)2 * L.nsplit * nloc * 2);
  L.off_diag = o; o += align4((int64_t)2 * nloc);
  L.off_rb = o; o += align4((int64_t)2 * L.nrb);
  L.total = o;
  (void)P;
  return L;
}

}  // namespace
}  // namespace vlmclip

using namespace vlmclip;

// `state`: fp32 scratch that carries the strips, 1/norms and partials from the forward to the backward call
// (vlmclip_clip_loss_state_size floats).  `counters`: vlmclip_clip_loss_counters(nloc) 32-bit words that must be ZERO
// before the first call; every kernel leaves them zero again, so one buffer per stream can be reused for ever.
extern "C" int64_t vlmclip_clip_loss_state_size(int N, int P, int nloc) { return fwd_layout(N, P, nloc).total; }
// strip kernel: 2 nrb + 1; gradient kernel: 2 nrb (level 2) + 2 nrb * ptiles (level 1, ptiles <= 8)
extern "C" int64_t vlmclip_clip_loss_counters(int nloc) { return 21 * (int64_t)ceil_div(nloc, CL_TILE) + 4; }
extern "C" int64_t vlmclip_clip_loss_bwd_workspace(int N, int P, int nloc) {
  const GradSplit g = grad_split(N, P, nloc);
  return (int64_t)g.ksplit * 2 * (int64_t)nloc * P + align4(2 * (int64_t)nloc * g.ptiles);
}

// Forward on the strips of rows [row0, row0 + nloc).  txt / img: the (all-gathered) un-normalised features [N, P];
// written: txt_n / img_n [N, P]; lse_loc [2 * nloc] = text-side LSE of the rows, then image-side LSE of the columns
// [row0, row0 + nloc); loss_share [1] = this strip's share of the loss (the loss itself when nloc == N).
// logits_per_text (optional, needs nloc == N): the full logit matrix.
extern "C" int vlmclip_clip_loss_fwd(const float* txt, const float* img, float logit_scale_exp, float* txt_n, float* img_n,
                                     float* logits_per_text, float* lse_loc, float* loss_share, float* state,
                                     int32_t* counters, int N, int P, int row0, int nloc, void* stream) {
  VLMCLIP_CHECK_ARG(txt && img && txt_n && img_n && lse_loc && loss_share && state && counters, "clip_loss_fwd: null pointer");
  VLMCLIP_CHECK_ARG(N > 0 && P > 0 && P % 4 == 0, "clip_loss_fwd: bad dims N=%d P=%d (P must be a multiple of 4)", N, P);
  VLMCLIP_CHECK_ARG(row0 >= 0 && nloc > 0 && row0 + nloc <= N, "clip_loss_fwd: local rows [%d,%d) outside [0,%d)", row0,
                    row0 + nloc, N);
  VLMCLIP_CHECK_ARG(logits_per_text == nullptr || nloc == N, "clip_loss_fwd: the full logit matrix needs nloc == N");
  VLMCLIP_CHECK_ARG((uintptr_t)txt_n % 16 == 0 && (uintptr_t)img_n % 16 == 0 && (uintptr_t)state % 16 == 0,
                    "clip_loss_fwd: txt_n / img_n / state must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const FwdLayout L = fwd_layout(N, P, nloc);
  count_launch(2);
  clip_norm2_kernel<<<dim3((N + 7) / 8, 2), 256, 0, s>>>(txt, img, txt_n, img_n, state + L.off_inv_t, state + L.off_inv_i, N, P);
  StripParams p;
  p.txt_n = txt_n;
  p.img_n = img_n;
  p.z[0] = state + L.off_zt;
  p.z[1] = state + L.off_zi;
  p.logits = logits_per_text;
  p.part = reinterpret_cast<float2*>(state + L.off_part);
  p.lse[0] = lse_loc;
  p.lse[1] = lse_loc + nloc;
  p.diag = state + L.off_diag;
  p.rb_loss = state + L.off_rb;
  p.loss = loss_share;
  p.counters = reinterpret_cast<unsigned*>(counters);
  p.s = logit_scale_exp;
  p.N = N;
  p.P = P;
  p.row0 = row0;
  p.nloc = nloc;
  p.ldz = L.ldz;
  p.nsplit = L.nsplit;
  p.cols_per_split = L.cols_per_split;
  clip_strip_lse_kernel<<<dim3(L.nrb, L.nsplit, 2), CL_THREADS, 0, s>>>(p);
  return report_cuda(cudaGetLastError(), "clip_loss_fwd launch");
}

// Gradient of the GLOBAL loss w.r.t. the un-normalised rows [row0, row0 + nloc), from the strips a forward call left in
// `state` for rows [strip_row0, strip_row0 + strip_rows) (the same range under data parallelism with an LSE exchange;
// the whole batch when one process emulates the ranks).  lse_all: see GradParams (for a single block: lse_stride is
// irrelevant and rows_per_rank = N, i.e. [lse_t(N) | lse_i(N)]).
extern "C" int vlmclip_clip_loss_bwd(const float* txt_n, const float* img_n, const float* lse_all, int lse_stride,
                                     int rows_per_rank, float logit_scale_exp, float* d_txt, float* d_img,
                                     float* d_logit_scale, float* state, int32_t* counters, float* workspace, int N, int P,
                                     int row0, int nloc, int strip_row0, int strip_rows, void* stream) {
  VLMCLIP_CHECK_ARG(txt_n && img_n && lse_all && d_txt && d_img && state && counters && workspace, "clip_loss_bwd: null pointer");
  VLMCLIP_CHECK_ARG(N > 0 && P > 0 && P % 4 == 0 && P <= 1024 && row0 >= 0 && nloc > 0 && row0 + nloc <= N,
                    "clip_loss_bwd: bad dims (P must be a multiple of 4, at most 1024)");
  VLMCLIP_CHECK_ARG(strip_row0 <= row0 && row0 + nloc <= strip_row0 + strip_rows && strip_row0 >= 0 && strip_row0 + strip_rows <= N,
                    "clip_loss_bwd: rows [%d,%d) are not inside the stored strips [%d,%d)", row0, row0 + nloc, strip_row0,
                    strip_row0 + strip_rows);
  VLMCLIP_CHECK_ARG(rows_per_rank > 0 && N % rows_per_rank == 0 && lse_stride >= 2 * rows_per_rank,
                    "clip_loss_bwd: bad LSE layout (rows_per_rank=%d, stride=%d)", rows_per_rank, lse_stride);
  VLMCLIP_CHECK_ARG(d_logit_scale == nullptr || (rows_per_rank == N || (strip_row0 == row0 && strip_rows == nloc && rows_per_rank == nloc)),
                    "clip_loss_bwd: d_logit_scale needs the local LSEs to be contiguous");
  VLMCLIP_CHECK_ARG((uintptr_t)workspace % 16 == 0 && (uintptr_t)img_n % 16 == 0 && (uintptr_t)txt_n % 16 == 0,
                    "clip_loss_bwd: txt_n / img_n / workspace must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const FwdLayout L = fwd_layout(N, P, strip_rows);
  GradParams g;
  g.txt_n = txt_n;
  g.img_n = img_n;
  g.inv_t = state + L.off_inv_t;
  g.inv_i = state + L.off_inv_i;
  g.z[0] = state + L.off_zt;
  g.z[1] = state + L.off_zi;
  g.lse_all = lse_all;
  g.lse_stride = lse_stride;
  g.rows_per_rank = rows_per_rank;
  const GradSplit gs = grad_split(N, P, nloc);
  g.dn = workspace;
  g.dotp = workspace + (int64_t)gs.ksplit * 2 * (int64_t)nloc * P;
  g.ptiles = gs.ptiles;
  g.ksplit = gs.ksplit;
  g.k_per_split = gs.k_per_split;
  g.d[0] = d_txt;
  g.d[1] = d_img;
  const int nrb = ceil_div(nloc, CL_TILE);
  g.counters = reinterpret_cast<unsigned*>(counters) + 2 * ceil_div(strip_rows, CL_TILE) + 1;
  g.s = logit_scale_exp;
  g.N = N;
  g.P = P;
  g.row0 = row0;
  g.nloc = nloc;
  g.ldz = L.ldz;
  g.z_row_off = row0 - strip_row0;
  count_launch(1);
  clip_strip_grad_kernel<<<dim3(nrb, gs.ptiles * gs.ksplit, 2), CL_THREADS, 0, s>>>(g);
  if (d_logit_scale != nullptr) {
    // the local LSEs: block of this rank (exchange layout) or rows [row0, ..) of the single block
    const int w = row0 / rows_per_rank, r = row0 - w * rows_per_rank;
    const float* lt = lse_all + (int64_t)w * lse_stride + r;
    const float* li = lt + rows_per_rank;
    count_launch(1);
    clip_strip_dscale_kernel<<<1, 256, 0, s>>>(g.z[0] + (int64_t)g.z_row_off * L.ldz, g.z[1] + (int64_t)g.z_row_off * L.ldz, lt,
                                              li, d_logit_scale, N, row0, nloc, L.ldz);
  }
  return report_cuda(cudaGetLastError(), "clip_loss_bwd launch");
}

extern "C" int vlmclip_scale_f32(const float* a, const float* scalar_dev, float* y, int64_t n, void* stream) {
  VLMCLIP_CHECK_ARG(a && scalar_dev && y && n > 0, "scale_f32: bad arguments");
  int64_t grid = (n + 255) / 256;
  if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
  count_launch(1);
  scale_f32_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(a, scalar_dev, y, n);
  return report_cuda(cudaGetLastError(), "scale_f32_kernel launch");
}

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
//   clip_strip_lse_kernel    64 x 64 fp32 register tiles over (row block, column split, problem); running (max, sum) of
//                            every row in registers, merged across the 16 threads of a row with warp shuffles; the
//                            last-arriving CTA of a row block merges the column splits in fixed order and the very last
//                            CTA adds up the loss: deterministic, no floating-point atomics
//   clip_strip_grad_kernel   dA = G B with G formed on the fly from the stored strip; the last-arriving CTA of a row
//                            block (over the column tiles of P) applies the normalisation backward, which needs the
//                            full-row dot product
// fp32 SIMT on purpose: the loss must match the reference's fp32 value to 1e-4 at a logit scale of 100, i.e. cosines
// to ~1e-7; the whole path is ~13 GFLOP at N = 4096 (0.02 % of that step).
#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int CL_TILE = 64;
constexpr int CL_KC = 16;
constexpr int CL_THREADS = 256;  // 16 x 16 threads, 4 x 4 outputs each
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

// acc[i][j] += sum_k A(ty*4+i, k) * B(tx*4+j, k) over k in [0, K), in chunks of 16 through shared memory; the next
// chunk is fetched into registers while the current one is multiplied.  fetchA / fetchB(k0, regs[4]) return the four
// elements this thread stages for chunk k0, storeA / storeB(regs) put them into As / Bs (layout [k][m]).
struct TileSmem {
  float As[CL_KC][CL_TILE + CL_PAD];
  float Bs[CL_KC][CL_TILE + CL_PAD];
};

template <class FetchA, class FetchB, class StoreA, class StoreB>
__device__ __forceinline__ void tile64_mainloop(TileSmem& sm, int K, float (&acc)[4][4], FetchA fetchA, FetchB fetchB,
                                                StoreA storeA, StoreB storeB) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float ra[4], rb[4];
  fetchA(0, ra);
  fetchB(0, rb);
  for (int k0 = 0; k0 < K; k0 += CL_KC) {
    storeA(ra);
    storeB(rb);
    __syncthreads();
    if (k0 + CL_KC < K) {
      fetchA(k0 + CL_KC, ra);
      fetchB(k0 + CL_KC, rb);
    }
#pragma unroll
    for (int kk = 0; kk < CL_KC; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sm.As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sm.Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
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

__global__ void __launch_bounds__(CL_THREADS)
clip_strip_lse_kernel(const StripParams p) {
  __shared__ TileSmem sm;
  __shared__ unsigned s_last;
  const int prob = blockIdx.z;
  const int rb = blockIdx.x;
  const int split = blockIdx.y;
  const int a0 = rb * CL_TILE;
  const float* A = (prob == 0 ? p.txt_n : p.img_n) + (int64_t)p.row0 * p.P;  // local rows
  const float* B = prob == 0 ? p.img_n : p.txt_n;                           // every row
  float* Z = p.z[prob];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // staging: one float4 along k of row lrow
  const int c_begin = split * p.cols_per_split;
  const int c_end = min(p.N, c_begin + p.cols_per_split);

  float run_m[4], run_l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    run_m[i] = -INFINITY;
    run_l[i] = 0.f;
  }
  for (int b0 = c_begin; b0 < c_end; b0 += CL_TILE) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool a_ok = a0 + lrow < p.nloc;
    const bool b_ok = b0 + lrow < c_end;
    const float* ap = A + (int64_t)(a0 + lrow) * p.P + lk;
    const float* bp = B + (int64_t)(b0 + lrow) * p.P + lk;
    auto fetch = [&](const float* src, bool ok, int k0, float (&r)[4]) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok && k0 + lk < p.P) v = __ldg(reinterpret_cast<const float4*>(src + k0));
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    };
    tile64_mainloop(
        sm, p.P, acc, [&](int k0, float (&r)[4]) { fetch(ap, a_ok, k0, r); },
        [&](int k0, float (&r)[4]) { fetch(bp, b_ok, k0, r); },
        [&](const float (&r)[4]) {
#pragma unroll
          for (int j = 0; j < 4; ++j) sm.As[lk + j][lrow] = r[j];
        },
        [&](const float (&r)[4]) {
#pragma unroll
          for (int j = 0; j < 4; ++j) sm.Bs[lk + j][lrow] = r[j];
        });
    // epilogue of the tile: scale, store the strip, diagonal, running log-sum-exp of the thread's four rows
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int a = a0 + ty * 4 + i;
      if (a >= p.nloc) continue;
      float z[4];
      float tm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = b0 + tx * 4 + j;
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
        for (int j = 0; j < 4; ++j) run_l[i] += __expf(z[j] - run_m[i]);
      }
    }
  }
  // merge the 16 threads that share a row (the tx dimension = one half-warp): (max, sum) pairs through shuffles
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float m = run_m[i], l = run_l[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
      const float mm = fmaxf(m, m2);
      l = (m > -INFINITY ? l * __expf(m - mm) : 0.f) + (m2 > -INFINITY ? l2 * __expf(m2 - mm) : 0.f);
      m = mm;
    }
    const int a = a0 + ty * 4 + i;
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
  float* red = &sm.As[0][0];  // 64 floats of scratch
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
      p.lse[prob][a] = lse;
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
  float* dn[2];           // [nloc, P] gradients w.r.t. the unit rows (scratch)
  float* d[2];            // [nloc, P] outputs: d_txt, d_img
  unsigned* counters;     // [2 * row blocks], zero on entry and on exit
  float s;
  int N, P, row0, nloc, ldz, z_row_off;
};

__device__ __forceinline__ float lse_of(const GradParams& p, int side, int r) {
  const int w = r / p.rows_per_rank;
  return __ldg(p.lse_all + (int64_t)w * p.lse_stride + side * p.rows_per_rank + (r - w * p.rows_per_rank));
}

__global__ void __launch_bounds__(CL_THREADS)
clip_strip_grad_kernel(const GradParams p) {
  __shared__ TileSmem sm;
  __shared__ unsigned s_last;
  const int prob = blockIdx.z;
  const int rb = blockIdx.x;
  const int a0 = rb * CL_TILE;
  const int c0 = blockIdx.y * CL_TILE;  // columns of P
  const float* Z = p.z[prob] + (int64_t)p.z_row_off * p.ldz;
  const int sideA = prob, sideB = 1 - prob;  // problem 0: own text-side LSE, the images' LSE for the columns; 1: swapped
  const float* B = prob == 0 ? p.img_n : p.txt_n;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;   // A staging: float4 along b (the reduction index) of row lrow
  const int bk = tid >> 4, bn = (tid & 15) * 4;    // B staging: float4 along the P columns of reduction row bk
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_ok = a0 + lrow < p.nloc;
  const float my_lse = a_ok ? lse_of(p, sideA, p.row0 + a0 + lrow) : 0.f;
  const int my_diag = p.row0 + a0 + lrow;
  const float* zp = Z + (int64_t)(a0 + lrow) * p.ldz + lk;
  tile64_mainloop(
      sm, p.N, acc,
      [&](int k0, float (&r)[4]) {  // G[a, b] for b = k0 + lk .. + 3
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int b = k0 + lk;
        if (a_ok && b < p.N) v = __ldg(reinterpret_cast<const float4*>(zp + k0));  // ldz is a multiple of 4
        const float zz[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float g = 0.f;
          if (a_ok && b + j < p.N) {
            g = __expf(zz[j] - my_lse) + __expf(zz[j] - lse_of(p, sideB, b + j));
            if (b + j == my_diag) g -= 2.f;
          }
          r[j] = g;
        }
      },
      [&](int k0, float (&r)[4]) {  // B[b, c0 + bn .. + 3] for b = k0 + bk
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + bk < p.N && c0 + bn < p.P) v = __ldg(reinterpret_cast<const float4*>(B + (int64_t)(k0 + bk) * p.P + c0 + bn));
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      },
      [&](const float (&r)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sm.As[lk + j][lrow] = r[j];
      },
      [&](const float (&r)[4]) { *reinterpret_cast<float4*>(&sm.Bs[bk][bn]) = make_float4(r[0], r[1], r[2], r[3]); });
  const float coef = p.s / (2.f * (float)p.N);
  float* dn = p.dn[prob];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + ty * 4 + i;
    const int c = c0 + tx * 4;
    if (a < p.nloc && c < p.P)
      *reinterpret_cast<float4*>(dn + (int64_t)a * p.P + c) =
          make_float4(coef * acc[i][0], coef * acc[i][1], coef * acc[i][2], coef * acc[i][3]);
  }
  // ---- last-arriving CTA of this (problem, row block): back through the normalisation, which needs whole rows ----
  __threadfence();
  __syncthreads();
  const int nrb = gridDim.x;
  if (tid == 0) s_last = atomicAdd(&p.counters[prob * nrb + rb], 1u) == gridDim.y - 1 ? 1u : 0u;
  __syncthreads();
  if (s_last == 0u) return;
  __threadfence();
  const float* An = (prob == 0 ? p.txt_n : p.img_n) + (int64_t)p.row0 * p.P;
  const float* inv = (prob == 0 ? p.inv_t : p.inv_i) + p.row0;
  float* out = p.d[prob];
  const int warp = tid >> 5, lane = tid & 31;
  for (int r = warp; r < CL_TILE; r += CL_THREADS / 32) {
    const int a = a0 + r;
    if (a >= p.nloc) break;
    const float* xn = An + (int64_t)a * p.P;
    const float* dr = dn + (int64_t)a * p.P;
    float dot = 0.f;
    for (int c = lane; c < p.P; c += 32) dot = fmaf(xn[c], __ldcg(dr + c), dot);
    dot = warp_sum(dot);
    const float iv = inv[a];
    for (int c = lane; c < p.P; c += 32) out[(int64_t)a * p.P + c] = (__ldcg(dr + c) - xn[c] * dot) * iv;
  }
  if (tid == 0) p.counters[prob * nrb + rb] = 0u;
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

// column splits of the strip kernel: enough CTAs to fill the GPU, at least one 64-column tile each
inline int pick_nsplit(int N, int nloc) {
  const int rbs = ceil_div(nloc, CL_TILE) * 2;
  int want = ceil_div(2 * sm_count(), rbs);
  const int max_split = ceil_div(N, CL_TILE);
  if (want > max_split) want = max_split;
  return want < 1 ? 1 : want;
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
extern "C" int64_t vlmclip_clip_loss_counters(int nloc) { return 4 * (int64_t)ceil_div(nloc, CL_TILE) + 4; }
extern "C" int64_t vlmclip_clip_loss_bwd_workspace(int N, int P, int nloc) {
  (void)N;
  return 2 * (int64_t)nloc * P;
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
  VLMCLIP_CHECK_ARG(N > 0 && P > 0 && P % 4 == 0 && row0 >= 0 && nloc > 0 && row0 + nloc <= N, "clip_loss_bwd: bad dims");
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
  g.dn[0] = workspace;
  g.dn[1] = workspace + (int64_t)nloc * P;
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
  clip_strip_grad_kernel<<<dim3(nrb, ceil_div(P, CL_TILE), 2), CL_THREADS, 0, s>>>(g);
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

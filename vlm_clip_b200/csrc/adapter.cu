// Fused bottleneck adapter, forward and adapter-only backward, fp32.
//
//   h = act(x W1^T + b1) (* hmask);  u = h W2^T + b2;  y = post(u, x)
//
// Reference modules (all the same bottleneck with different act / post):
//   adapter/clip_adapter.py:4-23,131-150  TextAdapter / VisionAdapter: GELU(erf), LN(u + x)
//   adapter/peclip.py:6-18                TextualAdapter: GELU(erf), u + x
//   model_t.py:13-33,163-169              Visual/TextAdapter: ReLU, then a*u + (1-a)*x and L2 normalise
//   model_v.py:18-27,280-286              BaseAdapter: ReLU (+dropout mask), same blend
//
// The backbone is frozen (model_m.py:64-70), so the backward produces only dW1, db1, dW2, db2, dgamma, dbeta
// (dx on request for full fine-tune).  In the hot path the adapter runs on ONE row per sequence (token 0,
// model_m.py:102,122), i.e. R = batch: a latency-bound problem, so one CTA owns 4 rows end to end (down-proj,
// activation, up-proj, residual, LayerNorm all in shared memory / registers) and the batch reductions for the
// weight gradients are a second, output-parallel launch (deterministic, no atomics).
#include "../../include/vlmclip.h"
#include "common.cuh"
#include "tile_f32.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int RB = 4;  // rows per CTA
constexpr int NT = 1024;

struct AdapterArgs {
  const void* x;
  int x_bf16;
  int64_t ldx;
  const float *W1, *b1, *W2, *b2, *gamma, *beta, *hmask;
  float* y;
  int R, D, A, act, post;
  float alpha, eps;
  // backward only
  const float* dy;
  float* dx;
  float *ws_du, *ws_zhat, *ws_x, *ws_h, *ws_dp;
};

__device__ __forceinline__ float act_fwd(float p, int act) {
  if (act == VLMCLIP_ACT_GELU_ERF) return gelu_erf(p);
  if (act == VLMCLIP_ACT_RELU) return fmaxf(p, 0.f);
  if (act == VLMCLIP_ACT_QUICK_GELU) {
    return p / (1.f + __expf(-1.702f * p));
  }
  return p;
}
__device__ __forceinline__ float act_bwd(float p, int act) {
  if (act == VLMCLIP_ACT_GELU_ERF) return gelu_erf_grad(p);
  if (act == VLMCLIP_ACT_RELU) return p > 0.f ? 1.f : 0.f;
  if (act == VLMCLIP_ACT_QUICK_GELU) {
    const float s = 1.f / (1.f + __expf(-1.702f * p));
    return s * (1.f + 1.702f * p * (1.f - s));
  }
  return 1.f;
}

// shared: xs[RB][D], ps[RB][A], hs[RB][A], us[RB][D]; computes rows [r0, r0+RB)
__device__ __forceinline__ void adapter_rows_forward(const AdapterArgs& a, int r0, float* xs, float* ps, float* hs,
                                                     float* us) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D, A = a.A;
  // ---- load x rows (bf16 or fp32, strided) ----
  for (int idx = tid; idx < RB * D; idx += NT) {
    const int r = idx / D, d = idx - r * D;
    float v = 0.f;
    if (r0 + r < a.R) {
      if (a.x_bf16)
        v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.x)[(int64_t)(r0 + r) * a.ldx + d]);
      else
        v = reinterpret_cast<const float*>(a.x)[(int64_t)(r0 + r) * a.ldx + d];
    }
    xs[idx] = v;
  }
  __syncthreads();
  // ---- down projection + activation: each warp owns UNR bottleneck units at a time, so UNR independent weight
  //      rows are in flight per lane (the kernel is L2-latency bound, not FLOP bound) ----
  constexpr int UNR = 4;
  for (int j0 = warp * UNR; j0 < A; j0 += (NT / 32) * UNR) {
    float part[UNR][RB];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int r = 0; r < RB; ++r) part[u][r] = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
      float4 wv[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        wv[u] = (j0 + u < A) ? __ldg(reinterpret_cast<const float4*>(a.W1 + (int64_t)(j0 + u) * D + d))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + r * D + d);
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          part[u][r] = fmaf(wv[u].x, xv.x, fmaf(wv[u].y, xv.y, fmaf(wv[u].z, xv.z, fmaf(wv[u].w, xv.w, part[u][r]))));
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int r = 0; r < RB; ++r) part[u][r] = warp_sum(part[u][r]);
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int j = j0 + u;
        if (j < A) {
          const float bj = a.b1[j];
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const float p = part[u][r] + bj;
            float h = act_fwd(p, a.act);
            if (a.hmask != nullptr && r0 + r < a.R) h *= a.hmask[(int64_t)(r0 + r) * A + j];
            ps[r * A + j] = p;
            hs[r * A + j] = h;
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- up projection: UNR output features per warp iteration ----
  for (int d0 = warp * UNR; d0 < D; d0 += (NT / 32) * UNR) {
    float part[UNR][RB];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int r = 0; r < RB; ++r) part[u][r] = 0.f;
    for (int j = lane * 4; j < A; j += 128) {
      float4 wv[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        wv[u] = (d0 + u < D) ? __ldg(reinterpret_cast<const float4*>(a.W2 + (int64_t)(d0 + u) * A + j))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(hs + r * A + j);
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          part[u][r] = fmaf(wv[u].x, hv.x, fmaf(wv[u].y, hv.y, fmaf(wv[u].z, hv.z, fmaf(wv[u].w, hv.w, part[u][r]))));
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int r = 0; r < RB; ++r) part[u][r] = warp_sum(part[u][r]);
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        if (d0 + u < D) {
          const float bd = a.b2[d0 + u];
#pragma unroll
          for (int r = 0; r < RB; ++r) us[r * D + d0 + u] = part[u][r] + bd;
        }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT) adapter_fwd_kernel(const AdapterArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int D = a.D, A = a.A;
  float* xs = sm;
  float* us = xs + RB * D;
  float* ps = us + RB * D;
  float* hs = ps + RB * A;
  const int r0 = blockIdx.x * RB;
  adapter_rows_forward(a, r0, xs, ps, hs, us);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= RB || r0 + warp >= a.R) return;
  const int r = warp;
  float* yr = a.y + (int64_t)(r0 + r) * D;
  const float* xr = xs + r * D;
  const float* ur = us + r * D;
  if (a.post == VLMCLIP_ADAPTER_RESIDUAL_LN) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += ur[d] + xr[d];
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float t = ur[d] + xr[d] - mean;
      q = fmaf(t, t, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + a.eps);
    for (int d = lane; d < D; d += 32) yr[d] = fmaf((ur[d] + xr[d] - mean) * rstd, a.gamma[d], a.beta[d]);
  } else if (a.post == VLMCLIP_ADAPTER_RESIDUAL) {
    for (int d = lane; d < D; d += 32) yr[d] = ur[d] + xr[d];
  } else if (a.post == VLMCLIP_ADAPTER_BLEND_L2) {
    float q = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float f = a.alpha * ur[d] + (1.f - a.alpha) * xr[d];
      q = fmaf(f, f, q);
    }
    const float inv = 1.f / sqrtf(warp_sum(q));
    for (int d = lane; d < D; d += 32) yr[d] = (a.alpha * ur[d] + (1.f - a.alpha) * xr[d]) * inv;
  } else {
    for (int d = lane; d < D; d += 32) yr[d] = ur[d];
  }
}

// Row phase of the backward: recompute the forward, then du, dp (and dx) for RB rows; stash what the weight
// gradient launch needs.
__global__ void __launch_bounds__(NT) adapter_bwd_rows_kernel(const AdapterArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int D = a.D, A = a.A;
  float* xs = sm;
  float* us = xs + RB * D;    // becomes du after the post-op backward
  float* dxs = us + RB * D;   // direct (post-op) contribution to dx
  float* ps = dxs + RB * D;
  float* hs = ps + RB * A;
  float* dps = hs + RB * A;
  const int r0 = blockIdx.x * RB;
  adapter_rows_forward(a, r0, xs, ps, hs, us);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp < RB) {
    const int r = warp;
    const bool live = r0 + r < a.R;
    const float* xr = xs + r * D;
    float* ur = us + r * D;
    float* dxr = dxs + r * D;
    const float* dyr = a.dy + (int64_t)(r0 + (live ? r : 0)) * D;
    float* zh = a.ws_zhat + (int64_t)(r0 + r) * D;
    if (!live) {
      for (int d = lane; d < D; d += 32) {
        ur[d] = 0.f;
        dxr[d] = 0.f;
      }
    } else if (a.post == VLMCLIP_ADAPTER_RESIDUAL_LN) {
      float s = 0.f;
      for (int d = lane; d < D; d += 32) s += ur[d] + xr[d];
      const float mean = warp_sum(s) / (float)D;
      float q = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float t = ur[d] + xr[d] - mean;
        q = fmaf(t, t, q);
      }
      const float rstd = rsqrtf(warp_sum(q) / (float)D + a.eps);
      float m1 = 0.f, m2 = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float z = (ur[d] + xr[d] - mean) * rstd;
        const float g = dyr[d] * a.gamma[d];
        m1 += g;
        m2 = fmaf(g, z, m2);
        zh[d] = z;
      }
      m1 = warp_sum(m1) / (float)D;
      m2 = warp_sum(m2) / (float)D;
      for (int d = lane; d < D; d += 32) {
        const float z = (ur[d] + xr[d] - mean) * rstd;
        const float g = dyr[d] * a.gamma[d];
        const float dz = rstd * (g - m1 - z * m2);
        ur[d] = dz;
        dxr[d] = dz;
      }
    } else if (a.post == VLMCLIP_ADAPTER_RESIDUAL) {
      for (int d = lane; d < D; d += 32) {
        const float g = dyr[d];
        ur[d] = g;
        dxr[d] = g;
      }
    } else if (a.post == VLMCLIP_ADAPTER_BLEND_L2) {
      float q = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float f = a.alpha * ur[d] + (1.f - a.alpha) * xr[d];
        q = fmaf(f, f, q);
      }
      const float inv = 1.f / sqrtf(warp_sum(q));
      float dot = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float yv = (a.alpha * ur[d] + (1.f - a.alpha) * xr[d]) * inv;
        dot = fmaf(dyr[d], yv, dot);
      }
      dot = warp_sum(dot);
      for (int d = lane; d < D; d += 32) {
        const float yv = (a.alpha * ur[d] + (1.f - a.alpha) * xr[d]) * inv;
        const float df = (dyr[d] - yv * dot) * inv;
        ur[d] = a.alpha * df;
        dxr[d] = (1.f - a.alpha) * df;
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        ur[d] = dyr[d];
        dxr[d] = 0.f;
      }
    }
  }
  __syncthreads();
  // stash du, x for the weight-gradient launch
  for (int idx = tid; idx < RB * D; idx += NT) {
    const int r = idx / D;
    if (r0 + r < a.R) {
      a.ws_du[(int64_t)r0 * D + idx] = us[idx];
      a.ws_x[(int64_t)r0 * D + idx] = xs[idx];
    }
  }
  // ---- dh = du W2, dp = dh * act'(p) * hmask: one thread per bottleneck unit ----
  for (int j = tid; j < A; j += NT) {
    float acc[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
      const float w = __ldg(a.W2 + (int64_t)d * A + j);
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = fmaf(us[r * D + d], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float dp = acc[r] * act_bwd(ps[r * A + j], a.act);
      if (a.hmask != nullptr && r0 + r < a.R) dp *= a.hmask[(int64_t)(r0 + r) * A + j];
      dps[r * A + j] = dp;
      if (r0 + r < a.R) {
        a.ws_dp[(int64_t)(r0 + r) * A + j] = dp;
        a.ws_h[(int64_t)(r0 + r) * A + j] = hs[r * A + j];
      }
    }
  }
  if (a.dx == nullptr) return;
  __syncthreads();
  // ---- dx = post-op part + dp W1: one thread per feature ----
  for (int d = tid; d < D; d += NT) {
    float acc[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = dxs[r * D + d];
#pragma unroll 8
    for (int j = 0; j < A; ++j) {
      const float w = __ldg(a.W1 + (int64_t)j * D + d);
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = fmaf(dps[r * A + j], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r)
      if (r0 + r < a.R) a.dx[(int64_t)(r0 + r) * D + d] = acc[r];
  }
}

// Weight gradients: blockIdx.z = 0: dW2[d][j] = sum_r du[r][d] h[r][j];  1: dW1[j][d] = sum_r dp[r][j] x[r][d]
__global__ void __launch_bounds__(TF_THREADS)
adapter_wgrad_kernel(const float* __restrict__ du, const float* __restrict__ h, const float* __restrict__ dp,
                     const float* __restrict__ x, float* __restrict__ dW1, float* __restrict__ dW2, int R, int D,
                     int A) {
  if (blockIdx.z == 0) {
    const int m0 = blockIdx.x * TF_TILE, n0 = blockIdx.y * TF_TILE;  // m over D, n over A
    if (m0 >= D || n0 >= A) return;
    tile_gemm_f32(
        R, m0, n0, false, false,
        [&](int m, int k) { return (m < D && k < R) ? du[(int64_t)k * D + m] : 0.f; },
        [&](int n, int k) { return (n < A && k < R) ? h[(int64_t)k * A + n] : 0.f; },
        [&](int m, int n, float v) {
          if (m < D && n < A) dW2[(int64_t)m * A + n] = v;
        });
  } else {
    const int m0 = blockIdx.y * TF_TILE, n0 = blockIdx.x * TF_TILE;  // m over A, n over D
    if (m0 >= A || n0 >= D) return;
    tile_gemm_f32(
        R, m0, n0, false, false,
        [&](int m, int k) { return (m < A && k < R) ? dp[(int64_t)k * A + m] : 0.f; },
        [&](int n, int k) { return (n < D && k < R) ? x[(int64_t)k * D + n] : 0.f; },
        [&](int m, int n, float v) {
          if (m < A && n < D) dW1[(int64_t)m * D + n] = v;
        });
  }
}

// Column sums over the batch: db2, dbeta, dgamma over D columns; db1 over A columns.  One block = 32 columns x 8
// row slices, combined through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
adapter_colsum_kernel(const float* __restrict__ du, const float* __restrict__ dy, const float* __restrict__ zhat,
                      const float* __restrict__ dp, float* __restrict__ db1, float* __restrict__ db2,
                      float* __restrict__ dgamma, float* __restrict__ dbeta, int R, int D, int A) {
  __shared__ float red[3][8][33];
  const int cx = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  if (c < D) {
    for (int r = slice; r < R; r += 8) {
      s0 += du[(int64_t)r * D + c];
      if (dgamma != nullptr) {
        const float g = dy[(int64_t)r * D + c];
        s1 += g;
        s2 = fmaf(g, zhat[(int64_t)r * D + c], s2);
      }
    }
  } else if (c - D < A) {
    const int j = c - D;
    for (int r = slice; r < R; r += 8) s0 += dp[(int64_t)r * A + j];
  }
  red[0][slice][cx] = s0;
  red[1][slice][cx] = s1;
  red[2][slice][cx] = s2;
  __syncthreads();
  if (slice == 0) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < 8; ++k) {
      t0 += red[0][k][cx];
      t1 += red[1][k][cx];
      t2 += red[2][k][cx];
    }
    if (c < D) {
      db2[c] = t0;
      if (dgamma != nullptr) {
        dbeta[c] = t1;
        dgamma[c] = t2;
      }
    } else if (c - D < A) {
      db1[c - D] = t0;
    }
  }
}

// y[R,N] = x[R,K] W[N,K]^T (+ b)
__global__ void __launch_bounds__(TF_THREADS)
linear_f32_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ W, const float* __restrict__ b,
                  float* __restrict__ y, int R, int N, int K) {
  const int m0 = blockIdx.x * TF_TILE, n0 = blockIdx.y * TF_TILE;
  tile_gemm_f32(
      K, m0, n0, true, true, [&](int m, int k) { return (m < R && k < K) ? x[(int64_t)m * ldx + k] : 0.f; },
      [&](int n, int k) { return (n < N && k < K) ? W[(int64_t)n * K + k] : 0.f; },
      [&](int m, int n, float v) {
        if (m < R && n < N) y[(int64_t)m * N + n] = v + (b != nullptr ? b[n] : 0.f);
      });
}
// dx[R,K] = dy[R,N] W[N,K]
__global__ void __launch_bounds__(TF_THREADS)
linear_f32_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ W, float* __restrict__ dx, int R,
                        int N, int K) {
  const int m0 = blockIdx.x * TF_TILE, n0 = blockIdx.y * TF_TILE;  // n over K (input features)
  tile_gemm_f32(
      N, m0, n0, true, false, [&](int m, int j) { return (m < R && j < N) ? dy[(int64_t)m * N + j] : 0.f; },
      [&](int n, int j) { return (n < K && j < N) ? W[(int64_t)j * K + n] : 0.f; },
      [&](int m, int n, float v) {
        if (m < R && n < K) dx[(int64_t)m * K + n] = v;
      });
}

// dW[N,K] = dy[R,N]^T x[R,K]   (weight gradient of a trainable projection: full fine-tune, HF:784-785)
__global__ void __launch_bounds__(TF_THREADS)
linear_f32_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, int64_t ldx, float* __restrict__ dW,
                        int R, int N, int K) {
  const int m0 = blockIdx.x * TF_TILE, n0 = blockIdx.y * TF_TILE;  // m over N (output features), n over K
  tile_gemm_f32(
      R, m0, n0, false, false, [&](int m, int r) { return (m < N && r < R) ? dy[(int64_t)r * N + m] : 0.f; },
      [&](int n, int r) { return (n < K && r < R) ? x[(int64_t)r * ldx + n] : 0.f; },
      [&](int m, int n, float v) {
        if (m < N && n < K) dW[(int64_t)m * K + n] = v;
      });
}
int check_adapter_args(const void* x, const float* W1, const float* b1, const float* W2, const float* b2,
                       const float* gamma, const float* beta, int R, int D, int A, int act, int post,
                       int64_t ldx) {
  VLMCLIP_CHECK_ARG(x && W1 && b1 && W2 && b2, "adapter: null x/W1/b1/W2/b2");
  VLMCLIP_CHECK_ARG(R > 0 && D > 0 && A > 0 && D % 4 == 0 && A % 4 == 0, "adapter: R=%d D=%d A=%d (D, A multiples of 4)",
                    R, D, A);
  VLMCLIP_CHECK_ARG(ldx >= D, "adapter: ldx=%lld < D=%d", (long long)ldx, D);
  VLMCLIP_CHECK_ARG(act >= 0 && act <= 3, "adapter: unknown activation %d", act);
  VLMCLIP_CHECK_ARG(post >= 0 && post <= 3, "adapter: unknown post-op %d", post);
  VLMCLIP_CHECK_ARG(post != VLMCLIP_ADAPTER_RESIDUAL_LN || (gamma && beta), "adapter: LN post-op needs gamma/beta");
  VLMCLIP_CHECK_ARG((uintptr_t)W1 % 16 == 0 && (uintptr_t)W2 % 16 == 0, "adapter: weights must be 16-byte aligned");
  return 0;
}

}  // namespace
}  // namespace vlmclip

using namespace vlmclip;

extern "C" int vlmclip_adapter_fwd(const void* x, int x_bf16, int64_t ldx, const float* W1, const float* b1,
                                   const float* W2, const float* b2, const float* gamma, const float* beta,
                                   const float* hmask, float* y, int R, int D, int A, int act, int post, float alpha,
                                   float eps, void* stream) {
  int rc = check_adapter_args(x, W1, b1, W2, b2, gamma, beta, R, D, A, act, post, ldx);
  if (rc) return rc;
  VLMCLIP_CHECK_ARG(y != nullptr, "adapter_fwd: null y");
  AdapterArgs a{};
  a.x = x;
  a.x_bf16 = x_bf16;
  a.ldx = ldx;
  a.W1 = W1;
  a.b1 = b1;
  a.W2 = W2;
  a.b2 = b2;
  a.gamma = gamma;
  a.beta = beta;
  a.hmask = hmask;
  a.y = y;
  a.R = R;
  a.D = D;
  a.A = A;
  a.act = act;
  a.post = post;
  a.alpha = alpha;
  a.eps = eps;
  const size_t smem = (size_t)RB * (2 * D + 2 * A) * sizeof(float);
  VLMCLIP_CHECK_ARG(smem <= 200 * 1024, "adapter_fwd: D=%d A=%d exceed the shared-memory budget", D, A);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(adapter_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  count_launch(1);
  adapter_fwd_kernel<<<(R + RB - 1) / RB, NT, smem, (cudaStream_t)stream>>>(a);
  return report_cuda(cudaGetLastError(), "adapter_fwd_kernel launch");
}

extern "C" int64_t vlmclip_adapter_bwd_workspace(int R, int D, int A) {
  const int64_t Rp = (int64_t)(R + RB - 1) / RB * RB;
  return Rp * (3 * (int64_t)D + 2 * (int64_t)A);
}

extern "C" int vlmclip_adapter_bwd(const void* x, int x_bf16, int64_t ldx, const float* W1, const float* b1,
                                   const float* W2, const float* b2, const float* gamma, const float* beta,
                                   const float* hmask, const float* dy, float* dW1, float* db1, float* dW2,
                                   float* db2, float* dgamma, float* dbeta, float* dx, float* workspace, int R,
                                   int D, int A, int act, int post, float alpha, float eps, void* stream) {
  int rc = check_adapter_args(x, W1, b1, W2, b2, gamma, beta, R, D, A, act, post, ldx);
  if (rc) return rc;
  VLMCLIP_CHECK_ARG(dy && dW1 && db1 && dW2 && db2 && workspace, "adapter_bwd: null dy/grad/workspace pointer");
  VLMCLIP_CHECK_ARG(post != VLMCLIP_ADAPTER_RESIDUAL_LN || (dgamma && dbeta), "adapter_bwd: LN post-op needs dgamma/dbeta");
  const int64_t Rp = (int64_t)(R + RB - 1) / RB * RB;
  AdapterArgs a{};
  a.x = x;
  a.x_bf16 = x_bf16;
  a.ldx = ldx;
  a.W1 = W1;
  a.b1 = b1;
  a.W2 = W2;
  a.b2 = b2;
  a.gamma = gamma;
  a.beta = beta;
  a.hmask = hmask;
  a.R = R;
  a.D = D;
  a.A = A;
  a.act = act;
  a.post = post;
  a.alpha = alpha;
  a.eps = eps;
  a.dy = dy;
  a.dx = dx;
  a.ws_du = workspace;
  a.ws_zhat = a.ws_du + Rp * D;
  a.ws_x = a.ws_zhat + Rp * D;
  a.ws_h = a.ws_x + Rp * D;
  a.ws_dp = a.ws_h + Rp * A;
  const size_t smem = (size_t)RB * (3 * D + 3 * A) * sizeof(float);
  VLMCLIP_CHECK_ARG(smem <= 200 * 1024, "adapter_bwd: D=%d A=%d exceed the shared-memory budget", D, A);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(adapter_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cudaStream_t s = (cudaStream_t)stream;
  count_launch(3);
  adapter_bwd_rows_kernel<<<(R + RB - 1) / RB, NT, smem, s>>>(a);
  VLMCLIP_CUDA(cudaGetLastError());
  dim3 grid((D + TF_TILE - 1) / TF_TILE, (A + TF_TILE - 1) / TF_TILE, 2);
  adapter_wgrad_kernel<<<grid, TF_THREADS, 0, s>>>(a.ws_du, a.ws_h, a.ws_dp, a.ws_x, dW1, dW2, R, D, A);
  VLMCLIP_CUDA(cudaGetLastError());
  const bool ln = post == VLMCLIP_ADAPTER_RESIDUAL_LN;
  adapter_colsum_kernel<<<(D + A + 31) / 32, 256, 0, s>>>(a.ws_du, dy, a.ws_zhat, a.ws_dp, db1, db2,
                                                           ln ? dgamma : nullptr, ln ? dbeta : nullptr, R, D, A);
  return report_cuda(cudaGetLastError(), "adapter_bwd launch");
}

extern "C" int vlmclip_linear_f32(const float* x, int64_t ldx, const float* W, const float* b, float* y, int R,
                                  int N, int K, void* stream) {
  VLMCLIP_CHECK_ARG(x && W && y && R > 0 && N > 0 && K > 0 && ldx >= K, "linear_f32: bad arguments");
  dim3 grid((R + TF_TILE - 1) / TF_TILE, (N + TF_TILE - 1) / TF_TILE);
  count_launch(1);
  linear_f32_kernel<<<grid, TF_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, W, b, y, R, N, K);
  return report_cuda(cudaGetLastError(), "linear_f32_kernel launch");
}

extern "C" int vlmclip_linear_f32_wgrad(const float* dy, const float* x, int64_t ldx, float* dW, int R, int N, int K,
                                        void* stream) {
  VLMCLIP_CHECK_ARG(dy && x && dW && R > 0 && N > 0 && K > 0 && ldx >= K, "linear_f32_wgrad: bad arguments");
  dim3 grid((N + TF_TILE - 1) / TF_TILE, (K + TF_TILE - 1) / TF_TILE);
  count_launch(1);
  linear_f32_wgrad_kernel<<<grid, TF_THREADS, 0, (cudaStream_t)stream>>>(dy, x, ldx, dW, R, N, K);
  return report_cuda(cudaGetLastError(), "linear_f32_wgrad_kernel launch");
}

extern "C" int vlmclip_linear_f32_dgrad(const float* dy, const float* W, float* dx, int R, int N, int K,
                                        void* stream) {
  VLMCLIP_CHECK_ARG(dy && W && dx && R > 0 && N > 0 && K > 0, "linear_f32_dgrad: bad arguments");
  dim3 grid((R + TF_TILE - 1) / TF_TILE, (K + TF_TILE - 1) / TF_TILE);
  count_launch(1);
  linear_f32_dgrad_kernel<<<grid, TF_THREADS, 0, (cudaStream_t)stream>>>(dy, W, dx, R, N, K);
  return report_cuda(cudaGetLastError(), "linear_f32_dgrad_kernel launch");
}

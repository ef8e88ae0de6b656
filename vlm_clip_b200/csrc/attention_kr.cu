// tcgen05 attention for 225 <= S <= 258 without a mask (ViT-L/14: S = 257) in ONE launch: out = softmax(q k^T * scale) v,
// head_dim 64 (longer sequences need a third query tile per unit and no longer fit two pipeline stages: they stay on
// the two-launch split).  HF modeling_clip.py:261-279 (eager_attention_forward), :318-331 (dispatch).
//
// STATUS: EXPERIMENTAL.  Compiles for sm_100a, has NOT run on a GPU yet (the round's GPU budget ended first).  It is
// reachable only with VLMCLIP_ATTN_SPLIT=4; the default for these shapes is the verified two-launch split of
// attention_pp.cu (variant 3).  tests/test_gpu_kernels.py::test_attention_key_range_split_variants_subprocess[4]
// holds it to the oracle and is skipped unless VLMCLIP_RUN_EXPERIMENTAL=1.
//
// Why: the two-launch split (profiles/r01_attention_split.txt: 480 + 384 us at B=512, H=16) loads Q twice, passes the
// first range's output through HBM, and spends a third 128-row tile slot per (batch, head) on the single row
// 257 = 2 * 128 + 1.  Here the two key ranges of a query tile are two INDEPENDENT single-pass softmaxes running side by
// side, one per softmax warpgroup, each with its own S/P buffer and its own O accumulator in TMEM; the epilogue folds
// the two accumulators with exact power-of-two weights, so nothing is rescaled in TMEM and nothing leaves the SM.
//   TMEM (512 columns): S/P of range 0 at [0, nb), of range 1 at [nb, 2 nb), O of range r at [2 nb + 64 r, +64);
//   nb <= 192.  K and V of ALL keys stay in shared memory for the unit (S rounded up to 16 rows each).
//   warp 0      TMA: per unit Q [128 x 64] x mtiles, the tail query rows (one 8-row box), K and V
//   warp 1      tcgen05: sub-tile tau = 2 * tile + r: S_r = Q K_r^T into buffer r, O_r = P_r V_r into accumulator r
//   warp 3      tail rows (S - Sq <= 2 query rows beyond the last full 128-row tile) on the CUDA cores, straight from
//               the swizzled K / V tiles in shared memory, overlapped with the tensor-core tiles of the same unit
//   warps 4-7   softmax of range 0 of every tile, warps 8-11 of range 1 (same single-pass scheme as attention_pp.cu)
//   warps 12-15 epilogue: out = (O_0 2^(off_0 - m) + O_1 2^(off_1 - m)) / (l_0 2^(off_0 - m) + l_1 2^(off_1 - m))
#include <cstdio>
#include <cstdlib>

#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int KR_THREADS = 512;
constexpr int KR_M = 128;
constexpr int KR_HD = 64;
constexpr uint32_t KR_Q_TILE_BYTES = KR_M * KR_HD * 2;  // 16 KB
constexpr int KR_TMEM_COLS = 512;
constexpr int KR_MAX_STAGES = 3;
constexpr int KR_MAX_TAIL = 2;            // query rows handled by the tail warp
constexpr int KR_MAX_KEYS_PER_LANE = 12;  // S <= 384
constexpr int KR_NBAR = 3 * 2 + 5 * 2;    // kv_full / kv_empty per stage, five pairs of per-range barriers

struct KRParams {
  __nv_bfloat16* out;
  int B, S, H, D;
  float scale_log2e;
  int Sq;      // query rows [0, Sq) go through the tensor cores, [Sq, S) to the tail warp
  int n_tail;  // S - Sq
  int mtiles;  // 128-row query tiles per unit
  int key0[2], Sk[2], Npad[2];  // key range r = [key0, key0 + Sk), Npad = Sk rounded up to 16 (MMA N / K extent)
  int nb;                       // TMEM columns per S/P buffer
  int kv_rows;                  // K (and V) rows per unit in shared memory: S rounded up to 16
  int kv_loads;                 // TMA loads per K (and per V): the box limit is 256 rows
  int num_units;
  uint32_t q_bytes, qt_bytes, kv_bytes, stage_bytes;
  int nstage;
  uint32_t out_stage_off;
};

__device__ __forceinline__ void kr_epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// bit j: key k0 + j belongs to the range (k0 + j < Sk)
__device__ __forceinline__ uint32_t kr_key_bits32(int Sk, int k0) {
  const int n = Sk - k0;
  return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
}

// One 32-column chunk of the single-pass softmax (attention_pp.cu: softmax_chunk, without masks): chunk maximum ->
// (rarely) raise the reference exponent by whole octaves and rescale what was already produced -> exp2, truncate to
// bf16 on the integer pipe, row sum over the truncated values, 16 packed columns of P to TMEM.
__device__ __forceinline__ void kr_softmax_chunk(const uint32_t (&cur)[32], int ch, int Sk, float c, float& off,
                                                 bool& has_ref, float (&l4)[4], uint32_t tb) {
  const int k0 = ch * 32;
  uint32_t pk[16];
  if (k0 >= Sk) {
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = 0u;
    tmem_st_32x32b_x16(tb + ch * 16, pk);
    return;
  }
  const bool full = k0 + 32 <= Sk;
  const uint32_t ok = kr_key_bits32(Sk, k0);
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  if (full) {
#pragma unroll
    for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(cur[j]));
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], (ok >> j) & 1u ? __uint_as_float(cur[j]) : -INFINITY);
  }
  const float mcs = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * c;
  if (!has_ref && mcs > -INFINITY) {
    off = mcs;
    has_ref = true;
  }
  const float excess = has_ref ? mcs - off : 0.f;
  const bool need = excess > 16.f;
  if (__any_sync(0xffffffffu, need)) {  // warp-uniform: the TMEM accesses below are warp collectives
    const float d = need ? ceilf(excess) : 0.f;
    const float f = exp2f(-d);  // exact power of two
    tmem_wait_st();
    for (int j = 0; j < ch; ++j) {
      uint32_t old[16];
      tmem_ld_32x32b_x16(tb + j * 16, old);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t lo = __float_as_uint(__uint_as_float(old[i] << 16) * f) & 0xffff0000u;
        const uint32_t hi = __float_as_uint(__uint_as_float(old[i] & 0xffff0000u) * f) & 0xffff0000u;
        old[i] = __byte_perm(lo, hi, 0x7632);
      }
      tmem_st_32x32b_x16(tb + j * 16, old);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) l4[i] *= f;
    off += d;
  }
  if (full) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t e0 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(cur[2 * j]), c, -off))) & 0xffff0000u;
      const uint32_t e1 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(cur[2 * j + 1]), c, -off))) & 0xffff0000u;
      l4[j & 3] += __uint_as_float(e0) + __uint_as_float(e1);
      pk[j] = __byte_perm(e0, e1, 0x7632);
    }
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (k0 + g * 8 >= Sk) {  // warp-uniform: no exp2 for 8-key groups outside the range
#pragma unroll
        for (int j = 0; j < 4; ++j) pk[g * 4 + j] = 0u;
      } else {
#pragma unroll
        for (int j = g * 4; j < g * 4 + 4; ++j) {
          uint32_t e0 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(cur[2 * j]), c, -off))) & 0xffff0000u;
          uint32_t e1 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(cur[2 * j + 1]), c, -off))) & 0xffff0000u;
          e0 = (ok >> (2 * j)) & 1u ? e0 : 0u;
          e1 = (ok >> (2 * j + 1)) & 1u ? e1 : 0u;
          l4[j & 3] += __uint_as_float(e0) + __uint_as_float(e1);
          pk[j] = __byte_perm(e0, e1, 0x7632);
        }
      }
    }
  }
  tmem_st_32x32b_x16(tb + ch * 16, pk);
}

__global__ void __launch_bounds__(KR_THREADS, 1)
attention_kr_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQt,
                    const __grid_constant__ CUtensorMap tmKV, const KRParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage0 = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.nstage * p.stage_bytes);
  uint64_t* kv_full = bars;                   // [3] TMA landed Q, tail Q, K, V of a unit
  uint64_t* kv_empty = bars + KR_MAX_STAGES;  // [3] the unit's last P.V (and the tail warp) finished reading the stage
  uint64_t* s_full = bars + 2 * KR_MAX_STAGES;  // [2] S_r = Q K_r^T complete
  uint64_t* p_full = s_full + 2;                // [2] softmax r wrote P_r (128 arrivals)
  uint64_t* e_done = s_full + 4;                // [2] the epilogue has read range r's row sums (128 arrivals)
  uint64_t* o_full = s_full + 6;                // [2] O_r = P_r V_r complete
  uint64_t* o_free = s_full + 8;                // [2] O_r drained to registers (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + KR_NBAR);
  float* s_l = reinterpret_cast<float*>(bars + KR_NBAR + 2);  // [2][128] row sums, softmax -> epilogue
  float* s_off = s_l + 256;                                   // [2][128] reference exponents
  uint8_t* s_out = smem + p.out_stage_off;  // [128 rows][128 B] bf16 output tile, 16-B chunks XOR-swizzled by row

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmQt);
    tma_prefetch_desc(&tmKV);
    for (int b = 0; b < KR_MAX_STAGES; ++b) {
      mbar_init(&kv_full[b], 1);
      mbar_init(&kv_empty[b], p.n_tail > 0 ? 2 : 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 128);
      mbar_init(&e_done[b], 128);
      mbar_init(&o_full[b], 1);
      mbar_init(&o_free[b], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    __syncwarp();
    tmem_alloc<KR_TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int n_units = (p.num_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // units of this CTA
  const int n_tiles = n_units * p.mtiles;
  const uint32_t o_col = tmem_base + 2u * (uint32_t)p.nb;  // accumulator r at o_col + 64 r

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      if (lane == 0) {
        // ===================== TMA producer =====================
        const int kv_part = p.kv_rows / p.kv_loads;
        for (int u = 0; u < n_units; ++u) {
          const int bh = blockIdx.x + u * gridDim.x;
          const int bb = bh / p.H, h = bh - bb * p.H;
          const int sg = u % p.nstage;
          uint8_t* st = stage0 + sg * p.stage_bytes;
          mbar_wait(&kv_empty[sg], ((u / p.nstage) & 1) ^ 1u);
          mbar_arrive_expect_tx(&kv_full[sg], p.q_bytes + p.qt_bytes + 2 * p.kv_bytes);
          for (int mt = 0; mt < p.mtiles; ++mt)
            tma_load_2d(st + mt * KR_Q_TILE_BYTES, &tmQ, &kv_full[sg], h * KR_HD, bb * p.S + mt * KR_M);
          if (p.n_tail > 0) tma_load_2d(st + p.q_bytes, &tmQt, &kv_full[sg], h * KR_HD, bb * p.S + p.Sq);
          uint8_t* kdst = st + p.q_bytes + p.qt_bytes;
          for (int kl = 0; kl < p.kv_loads; ++kl) {
            tma_load_2d(kdst + kl * kv_part * 128, &tmKV, &kv_full[sg], p.D + h * KR_HD, bb * p.S + kl * kv_part);
            tma_load_2d(kdst + p.kv_bytes + kl * kv_part * 128, &tmKV, &kv_full[sg], 2 * p.D + h * KR_HD,
                        bb * p.S + kl * kv_part);
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // ===================== MMA issuer: sub-tile tau = 2 * tile + range =====================
        const uint32_t idesc_s0 = make_idesc_bf16(KR_M, p.Npad[0]), idesc_s1 = make_idesc_bf16(KR_M, p.Npad[1]);
        const uint32_t idesc_o = make_idesc_bf16_b_mn(KR_M, KR_HD);
        const int n_sub = 2 * n_tiles;
        auto issue_s = [&](int tau) {
          const int t = tau >> 1, r = tau & 1;
          const int u = t / p.mtiles, mt = t - u * p.mtiles;
          const int sg = u % p.nstage;
          uint8_t* st = stage0 + sg * p.stage_bytes;
          mbar_wait(&kv_full[sg], (u / p.nstage) & 1);
          tcgen05_fence_after();
          const uint64_t qd = make_umma_desc_sw128(smem_u32(st + mt * KR_Q_TILE_BYTES));
          const uint64_t kd = make_umma_desc_sw128(smem_u32(st + p.q_bytes + p.qt_bytes + p.key0[r] * 128));
#pragma unroll
          for (int k = 0; k < KR_HD / 16; ++k)
            umma_bf16_ss(tmem_base + r * p.nb, qd + 2u * k, kd + 2u * k, r ? idesc_s1 : idesc_s0, k != 0 ? 1u : 0u);
          umma_commit(&s_full[r]);
        };
        if (n_sub > 0) issue_s(0);
        if (n_sub > 1) issue_s(1);
        for (int tau = 0; tau < n_sub; ++tau) {
          const int t = tau >> 1, r = tau & 1;
          const int u = t / p.mtiles, mt = t - u * p.mtiles;
          const int sg = u % p.nstage;
          uint8_t* st = stage0 + sg * p.stage_bytes;
          mbar_wait(&p_full[r], t & 1);          // P_r(t) is in TMEM and every S_r(t) read has retired
          mbar_wait(&o_free[r], (t & 1) ^ 1u);   // O_r(t-1) has been drained
          tcgen05_fence_after();
          const uint64_t vd = make_umma_desc_mn_sw128(
              smem_u32(st + p.q_bytes + p.qt_bytes + p.kv_bytes + p.key0[r] * 128), p.kv_bytes);
          const int ksteps = p.Npad[r] >> 4;
          for (int k = 0; k < ksteps; ++k)  // 16 keys per step: 8 packed TMEM columns of P, 2048 B of V
            umma_bf16_ts(o_col + 64u * r, tmem_base + r * p.nb + k * 8, vd + static_cast<uint64_t>(k) * (2048u >> 4),
                         idesc_o, k != 0 ? 1u : 0u);
          umma_commit(&o_full[r]);
          if (r == 1 && mt == p.mtiles - 1) umma_commit(&kv_empty[sg]);
          if (tau + 2 < n_sub) issue_s(tau + 2);  // in order behind P_r.V_r(t): may overwrite S/P buffer r
        }
      }
    } else if (warp == 3 && p.n_tail > 0) {
      // ===================== tail query rows on the CUDA cores, from the swizzled K / V tiles =====================
      // Row j of a 128-byte-swizzled tile keeps its 16-byte chunk c at chunk position c ^ (j & 7).  Scores: lane l owns
      // keys l, l + 32, ... (the query chunk is a broadcast read); output: lane l owns dimensions 2 l, 2 l + 1.
      const float c = p.scale_log2e;
      const int vch = lane >> 2;
      const uint32_t voff = static_cast<uint32_t>(lane & 3) * 4u;
      for (int u = 0; u < n_units; ++u) {
        const int bh = blockIdx.x + u * gridDim.x;
        const int bb = bh / p.H, h = bh - bb * p.H;
        const int sg = u % p.nstage;
        const uint8_t* qt = stage0 + sg * p.stage_bytes + p.q_bytes;
        const uint8_t* ks = qt + p.qt_bytes;
        const uint8_t* vs = ks + p.kv_bytes;
        mbar_wait(&kv_full[sg], (u / p.nstage) & 1);
        for (int i = 0; i < p.n_tail; ++i) {
          float sc[KR_MAX_KEYS_PER_LANE];
          float m = -INFINITY;
#pragma unroll
          for (int i2 = 0; i2 < KR_MAX_KEYS_PER_LANE; ++i2) {
            const int j = lane + i2 * 32;
            float s = -INFINITY;
            if (j < p.S) {
              float acc = 0.f;
#pragma unroll
              for (int cc = 0; cc < 8; ++cc) {
                const uint4 qv = *reinterpret_cast<const uint4*>(qt + i * 128 + ((cc ^ i) << 4));
                const uint4 kv = *reinterpret_cast<const uint4*>(ks + j * 128 + ((cc ^ (j & 7)) << 4));
                acc = fmaf(bf16_lo(qv.x), bf16_lo(kv.x), acc); acc = fmaf(bf16_hi(qv.x), bf16_hi(kv.x), acc);
                acc = fmaf(bf16_lo(qv.y), bf16_lo(kv.y), acc); acc = fmaf(bf16_hi(qv.y), bf16_hi(kv.y), acc);
                acc = fmaf(bf16_lo(qv.z), bf16_lo(kv.z), acc); acc = fmaf(bf16_hi(qv.z), bf16_hi(kv.z), acc);
                acc = fmaf(bf16_lo(qv.w), bf16_lo(kv.w), acc); acc = fmaf(bf16_hi(qv.w), bf16_hi(kv.w), acc);
              }
              s = acc * c;
            }
            sc[i2] = s;
            m = fmaxf(m, s);
          }
          m = warp_max(m);
          float l = 0.f;
#pragma unroll
          for (int i2 = 0; i2 < KR_MAX_KEYS_PER_LANE; ++i2) {
            sc[i2] = (lane + i2 * 32 < p.S) ? exp2f(sc[i2] - m) : 0.f;
            l += sc[i2];
          }
          l = warp_sum(l);
          float o0 = 0.f, o1 = 0.f;
#pragma unroll
          for (int i2 = 0; i2 < KR_MAX_KEYS_PER_LANE; ++i2) {
            if (i2 * 32 < p.S) {  // warp-uniform
              const int nj = min(32, p.S - i2 * 32);
              for (int jj = 0; jj < nj; ++jj) {
                const int j = i2 * 32 + jj;
                const float pj = __shfl_sync(0xffffffffu, sc[i2], jj);
                const uint32_t raw = *reinterpret_cast<const uint32_t*>(vs + j * 128 + ((vch ^ (j & 7)) << 4) + voff);
                o0 = fmaf(pj, bf16_lo(raw), o0);
                o1 = fmaf(pj, bf16_hi(raw), o1);
              }
            }
          }
          const float inv = l > 0.f ? 1.f / l : 0.f;
          __nv_bfloat16* orow = p.out + ((int64_t)bb * p.S + p.Sq + i) * p.D + h * KR_HD;
          *reinterpret_cast<uint32_t*>(orow + lane * 2) = pack_bf16x2(o0 * inv, o1 * inv);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&kv_empty[sg]);  // the stage may be refilled once the unit's last P.V has also retired
      }
    }
  } else if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    // ===================== epilogue warpgroup: fold the two accumulators -> bf16 -> smem tile -> row-contiguous stores
    const int wq = warp & 3;  // TMEM lane quarter
    const int r_local = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    for (int t = 0; t < n_tiles; ++t) {
      const int u = t / p.mtiles, mt = t - u * p.mtiles;
      const int bh = blockIdx.x + u * gridDim.x;
      const int bb = bh / p.H, h = bh - bb * p.H;
      const bool warp_valid = (mt * KR_M + wq * 32) < p.Sq;  // warp-uniform
      const uint32_t par = t & 1;
      mbar_wait(&o_full[0], par);
      mbar_wait(&o_full[1], par);
      tcgen05_fence_after();
      mbar_wait(&p_full[0], par);  // already complete (they precede O); order the reads of the row statistics
      mbar_wait(&p_full[1], par);
      const float l0 = s_l[r_local], off0 = s_off[r_local];
      const float l1 = s_l[128 + r_local], off1 = s_off[128 + r_local];
      mbar_arrive(&e_done[0]);
      mbar_arrive(&e_done[1]);
      const bool h0 = l0 > 0.f, h1 = l1 > 0.f;
      const float m = fmaxf(h0 ? off0 : -INFINITY, h1 ? off1 : -INFINITY);
      const float a0 = h0 ? exp2f(off0 - m) : 0.f;
      const float a1 = h1 ? exp2f(off1 - m) : 0.f;
      const float den = l0 * a0 + l1 * a1;
      const float inv = den > 0.f ? __fdividef(1.f, den) : 0.f;
      const float w0 = a0 * inv, w1 = a1 * inv;
      kr_epi_bar_sync();  // the previous tile's stores have read the staging tile
      uint8_t* srow = s_out + r_local * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t oa[32], ob[32];
        if (warp_valid) {
          __syncwarp();
          tmem_ld_32x32b_x32(o_col + lane_off + half * 32, oa);
          tmem_ld_32x32b_x32(o_col + lane_off + 64 + half * 32, ob);
          tmem_wait_ld();
        }
        if (half == 1) {
          tcgen05_fence_before();
          mbar_arrive(&o_free[0]);  // both accumulators are in registers: the next tile's P.V may start
          mbar_arrive(&o_free[1]);
        }
        if (warp_valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t ow[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i0 = q * 8 + 2 * e;
              const float lo = fmaf(__uint_as_float(oa[i0]), w0, __uint_as_float(ob[i0]) * w1);
              const float hi = fmaf(__uint_as_float(oa[i0 + 1]), w0, __uint_as_float(ob[i0 + 1]) * w1);
              ow[e] = pack_bf16x2(lo, hi);
            }
            *reinterpret_cast<uint4*>(srow + (((half * 4 + q) ^ (r_local & 7)) << 4)) =
                make_uint4(ow[0], ow[1], ow[2], ow[3]);
          }
        }
      }
      kr_epi_bar_sync();  // the whole output tile is in smem
      {
        // 4 warps x 32 rows; one store instruction = 4 rows x 128 contiguous bytes (8 lanes per row)
        const int chunk = lane & 7;
        __nv_bfloat16* obase = p.out + ((int64_t)bb * p.S) * p.D + h * KR_HD + chunk * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = wq * 32 + k * 4 + (lane >> 3);
          const int qr = mt * KR_M + r;
          if (qr < p.Sq) {
            const uint4 v = *reinterpret_cast<const uint4*>(s_out + r * 128 + ((chunk ^ (r & 7)) << 4));
            st_v4(obase + (int64_t)qr * p.D, v);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    // ===================== softmax warpgroup r: range r of every tile, one thread per query row =====================
    const int r = (warp - 4) >> 2;
    const int wq = warp & 3;  // TMEM lane quarter
    const int r_local = wq * 32 + lane;
    const float c = p.scale_log2e;
    const int Sk = p.Sk[r];
    const int nch = (p.Npad[r] + 31) >> 5;  // 32-column chunks
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tb = tmem_base + lane_off + r * p.nb;
    for (int t = 0; t < n_tiles; ++t) {
      const int u = t / p.mtiles, mt = t - u * p.mtiles;
      const bool warp_valid = (mt * KR_M + wq * 32) < p.Sq;  // warp-uniform
      const uint32_t par = t & 1;
      mbar_wait(&s_full[r], par);
      tcgen05_fence_after();
      float l = 0.f, off_pub = 0.f;
      if (warp_valid) {
        __syncwarp();
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        float off = 0.f;
        bool has_ref = false;
        uint32_t sa[32], sb[32];
        tmem_ld_32x32b_x32(tb, sa);
        for (int ch = 0; ch < nch; ch += 2) {
          tmem_wait_ld();
          if (ch + 1 < nch) tmem_ld_32x32b_x32(tb + (ch + 1) * 32, sb);
          kr_softmax_chunk(sa, ch, Sk, c, off, has_ref, l4, tb);
          if (ch + 1 < nch) {
            tmem_wait_ld();
            if (ch + 2 < nch) tmem_ld_32x32b_x32(tb + (ch + 2) * 32, sa);
            kr_softmax_chunk(sb, ch + 1, Sk, c, off, has_ref, l4, tb);
          }
        }
        l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        off_pub = off;
      }
      mbar_wait(&e_done[r], par ^ 1u);  // the epilogue has read the statistics of tile t-1
      s_l[r * 128 + r_local] = l;
      s_off[r * 128 + r_local] = off_pub;
      if (warp_valid) tmem_wait_st();
      tcgen05_fence_before();
      mbar_arrive(&p_full[r]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<KR_TMEM_COLS>(tmem_base);
  }
}

// Shapes and shared-memory plan; false when S is outside 225..384 or two pipeline stages do not fit.  Two stages are a
// protocol requirement, not a tuning choice: S of the NEXT unit is issued before the current unit's last P.V, so with
// one stage the issuer would wait for operands that can only be loaded after that P.V released the stage
// (tools/kr_protocol_sim.py reproduces the deadlock).
static bool kr_plan(int B, int S, int H, float scale, KRParams& p, size_t& smem) {
  p.B = B;
  p.S = S;
  p.H = H;
  p.D = H * KR_HD;
  p.scale_log2e = scale * 1.4426950408889634f;
  const int tail = S % KR_M;
  p.n_tail = (tail >= 1 && tail <= KR_MAX_TAIL) ? tail : 0;
  p.Sq = S - p.n_tail;
  p.mtiles = (p.Sq + KR_M - 1) / KR_M;
  p.key0[0] = 0;
  p.key0[1] = 32 * ((S / 2 + 16) / 32);  // 257 -> [0, 128) + [128, 257): 4 + 5 chunks of 32 keys
  p.Sk[0] = p.key0[1];
  p.Sk[1] = S - p.key0[1];
  int nb = 0;
  for (int r = 0; r < 2; ++r) {
    p.Npad[r] = (p.Sk[r] + 15) / 16 * 16;
    nb = max(nb, (p.Npad[r] + 31) / 32 * 32);
  }
  p.nb = nb;
  if (S <= 224 || S > 32 * KR_MAX_KEYS_PER_LANE || 2 * nb + 2 * KR_HD > KR_TMEM_COLS) return false;
  p.kv_rows = (S + 15) / 16 * 16;
  p.kv_loads = p.kv_rows > 256 ? 2 : 1;
  p.num_units = B * H;
  p.q_bytes = (uint32_t)p.mtiles * KR_Q_TILE_BYTES;
  p.qt_bytes = p.n_tail > 0 ? 1024u : 0u;
  p.kv_bytes = (uint32_t)p.kv_rows * 128u;  // kv_rows is a multiple of 16: 1024-byte aligned
  p.stage_bytes = p.q_bytes + p.qt_bytes + 2 * p.kv_bytes;
  const size_t ctrl = KR_NBAR * 8 + 16 + 512 * sizeof(float);
  const size_t budget = 232448 - KR_M * 128 - ctrl - 1024;
  const int nstage = (int)(budget / p.stage_bytes);
  p.nstage = nstage > KR_MAX_STAGES ? KR_MAX_STAGES : nstage;
  if (p.nstage < 2) return false;
  p.out_stage_off = (uint32_t)(((size_t)p.nstage * p.stage_bytes + ctrl + 1023) & ~(size_t)1023);
  smem = (size_t)p.out_stage_off + KR_M * 128;
  return true;
}

}  // namespace

// whether attention_fwd_keyranges takes sequences of S tokens (S = 257: yes; S >= 289: K + V + three query tiles no
// longer fit twice)
bool attention_keyranges_supported(int S) {
  KRParams p;
  size_t smem = 0;
  return kr_plan(1, S, 1, 1.f, p, smem);
}

int attention_fwd_keyranges(const void* qkv, void* out, int B, int S, int H, float scale, cudaStream_t s) {
  KRParams p;
  size_t smem = 0;
  if (!kr_plan(B, S, H, scale, p, smem)) {
    set_last_error("attention (key ranges): S=%d is not supported by the single-launch kernel", S);
    return -1;
  }
  p.out = (__nv_bfloat16*)out;
  CUtensorMap tmQ, tmQt, tmKV;
  const int64_t rows = (int64_t)B * S;
  int rc = make_tmap_bf16(&tmQ, qkv, rows, 3 * (int64_t)p.D, 3 * (int64_t)p.D, KR_M);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmQt, qkv, rows, 3 * (int64_t)p.D, 3 * (int64_t)p.D, 8);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmKV, qkv, rows, 3 * (int64_t)p.D, 3 * (int64_t)p.D, p.kv_rows / p.kv_loads);
  if (rc) return rc;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(attention_kr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const int grid = p.num_units < sm_count() ? p.num_units : sm_count();
  count_launch(1);
  return report_cuda(launch_pdl(attention_kr_kernel, dim3(grid), dim3(KR_THREADS), smem, s, 1, tmQ, tmQt, tmKV, p),
                     "attention_kr_kernel launch");
}

}  // namespace vlmclip

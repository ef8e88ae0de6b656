// tcgen05 attention, ping-pong schedule: out = softmax(q k^T * scale + mask) v, head_dim 64, S <= 224.
//   HF modeling_clip.py:261-279 (eager_attention_forward), :318-331 (dispatch), :546-551 (causal + padding mask)
//
// One persistent CTA per SM walks (batch, head) units.  A unit's K and V are loaded ONCE and shared by its one or two
// 128-query tiles; the tiles of the CTA form one sequence t = 0, 1, 2, ... and tile t belongs to softmax warpgroup
// t & 1 and to TMEM S/P buffer t & 1, so one warpgroup's softmax overlaps the other one's MMAs; O is drained and
// written out by a third warpgroup, off the softmax path.  Measured limits on B200 (tools/mufu_bench.cu and the
// VLMCLIP_ATTN_DEBUG=1 phase timers): MUFU.EX2 issues one warp instruction per 8 clk per scheduler, tcgen05.ld
// delivers about 64 B/clk/SM, so S is streamed from TMEM exactly once (a 208-score row does not fit one thread's
// registers) with the exp2 work hidden behind nothing but the other warpgroup's tile.
//   warp 0     TMA: per unit Q [128*mtiles x 64], K [Npad x 64], V [Npad x 64] (rows of the fused qkv activation, SW128)
//   warp 1     tcgen05: S = Q K^T (SS MMA, M=128, N=Npad, 4 k-steps) -> TMEM buffer t&1, issued right behind P.V(t-2);
//              O = P V (TS MMA: A = P from TMEM, B = V as an MN-major smem operand) -> TMEM columns [448, 512)
//   warps 4-11 softmax, ONE thread per query row, no cross-thread exchange, ONE pass over S: the reference exponent is the
//              maximum of the row's first chunk (raised by whole octaves, exactly, if a later chunk ever exceeds it by
//              2^16), p = exp2(s*c - ref) truncated to bf16 with integer ops (F2FP shares the SFU pipe with MUFU.EX2),
//              row sum over the truncated values, P written over the S columns (tcgen05.st)
//   warps 12-15 epilogue: tcgen05.ld O (frees O for the next P.V), * 1/rowsum, bf16, through a swizzled smem tile so
//              that global stores are row-contiguous
#include <cstdio>
#include <cstdlib>

#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);

namespace {

constexpr int PP_THREADS = 512;  // warps 0-3 control, 4-7 / 8-11 softmax warpgroups, 12-15 epilogue warpgroup
constexpr int PP_M = 128;
constexpr int PP_HD = 64;
constexpr uint32_t PP_Q_TILE_BYTES = PP_M * PP_HD * 2;  // 16 KB
constexpr int PP_TMEM_COLS = 512;
constexpr int PP_O_COL = 448;
constexpr int PP_MAX_STAGES = 4;
constexpr int PP_MAX_NPAD = 224;

struct PPParams {
  const uint8_t* key_mask;
  __nv_bfloat16* out;
  int B, S, H, D;
  int causal;
  float scale_log2e;
  int Npad;    // keys rounded up to a multiple of 16
  int nb;      // TMEM columns per S buffer: Npad rounded up to 32
  int mtiles;  // query tiles per unit (1 or 2)
  int num_units;
  uint32_t q_bytes;      // mtiles * 16 KB
  uint32_t kv_bytes;     // Npad * 128
  uint32_t kv_stride;    // kv_bytes rounded up to 1024
  uint32_t stage_bytes;  // q_bytes + 2 * kv_stride
  int nstage;
  uint32_t out_stage_off;  // 16 KB output staging tile
  int debug;
};

// visibility bits of keys [k0, k0+32) for query row qrow
template <bool GENERAL_MASK>
__device__ __forceinline__ uint32_t key_bits32(const PPParams& p, const uint8_t* km, int k0, int qrow) {
  uint32_t bits = 0u;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int key = k0 + j;
    bool ok = key < p.S;
    if (GENERAL_MASK) {
      if (p.causal) ok = ok && key <= qrow;
      if (km != nullptr && ok) ok = __ldg(km + key) != 0;
    }
    bits |= (ok ? 1u : 0u) << j;
  }
  return bits;
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// running 4-way row maximum over one 32-column chunk of scores
template <bool GENERAL_MASK>
__device__ __forceinline__ void max_chunk(const PPParams& p, const uint32_t (&cur)[32], int k0, int kmax_warp,
                                          const uint8_t* km, int qrow, float (&m4)[4]) {
  if (k0 >= kmax_warp) return;
  if (!GENERAL_MASK && k0 + 32 <= p.S) {
#pragma unroll
    for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(cur[j]));
  } else {
    const uint32_t ok = key_bits32<GENERAL_MASK>(p, km, k0, qrow);
#pragma unroll
    for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], (ok >> j) & 1u ? __uint_as_float(cur[j]) : -INFINITY);
  }
}

// p = exp2(s*c - off) for one 32-column chunk -> bf16 by truncation (integer pipe), 4-way row sum over the truncated
// values, 16 packed columns of P written to TMEM at taddr
template <bool GENERAL_MASK>
__device__ __forceinline__ void exp_chunk(const PPParams& p, const uint32_t (&cur)[32], int k0, int kmax_warp,
                                          const uint8_t* km, int qrow, float c, float off, float (&l4)[4],
                                          uint32_t taddr) {
  uint32_t pk[16];
  if (k0 >= kmax_warp) {
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = 0u;
  } else if (!GENERAL_MASK && k0 + 32 <= p.S) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t e0 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(cur[2 * j]), c, -off))) & 0xffff0000u;
      const uint32_t e1 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(cur[2 * j + 1]), c, -off))) & 0xffff0000u;
      l4[j & 3] += __uint_as_float(e0) + __uint_as_float(e1);
      pk[j] = __byte_perm(e0, e1, 0x7632);
    }
  } else {
    const uint32_t ok = key_bits32<GENERAL_MASK>(p, km, k0, qrow);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (k0 + g * 8 >= kmax_warp) {  // warp-uniform: no exp2 for 8-key groups nobody in the warp sees
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
  tmem_st_32x32b_x16(taddr, pk);
}

// One 32-column chunk of the single-pass softmax: chunk maximum -> (rarely) raise the reference by whole octaves and
// rescale what was already produced -> exp2 / truncate / sum / pack / store.
template <bool GENERAL_MASK>
__device__ __forceinline__ void softmax_chunk(const PPParams& p, const uint32_t (&cur)[32], int ch, int kmax_warp,
                                              const uint8_t* km, int qrow, float c, float& off, bool& has_ref,
                                              float (&l4)[4], uint32_t tb) {
  const int k0 = ch * 32;
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  max_chunk<GENERAL_MASK>(p, cur, k0, kmax_warp, km, qrow, m4);
  const float mcs = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * c;  // -inf: no visible key in this chunk
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
      uint32_t pk[16];
      tmem_ld_32x32b_x16(tb + j * 16, pk);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t lo = __float_as_uint(__uint_as_float(pk[i] << 16) * f) & 0xffff0000u;
        const uint32_t hi = __float_as_uint(__uint_as_float(pk[i] & 0xffff0000u) * f) & 0xffff0000u;
        pk[i] = __byte_perm(lo, hi, 0x7632);
      }
      tmem_st_32x32b_x16(tb + j * 16, pk);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) l4[i] *= f;
    off += d;
  }
  exp_chunk<GENERAL_MASK>(p, cur, k0, kmax_warp, km, qrow, c, off, l4, tb + ch * 16);
}

template <bool GENERAL_MASK>
__global__ void __launch_bounds__(PP_THREADS, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const PPParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage0 = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.nstage * p.stage_bytes);
  uint64_t* kv_full = bars;                     // [4] TMA landed Q, K, V of a unit
  uint64_t* kv_empty = bars + PP_MAX_STAGES;    // [4] the unit's last P.V finished reading the stage
  uint64_t* s_full = bars + 2 * PP_MAX_STAGES;  // [2] S = Q K^T complete (per buffer / warpgroup)
  uint64_t* p_full = s_full + 2;                // [2] softmax wrote P (128 arrivals)
  uint64_t* e_done = s_full + 4;                // [2] the epilogue has read the row sums of the buffer (128 arrivals)
  uint64_t* o_full = s_full + 6;                // [1] O = P V complete, in tile order
  uint64_t* o_free = s_full + 7;                // [1] O drained to registers (128 arrivals), in tile order
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);
  float* s_l = reinterpret_cast<float*>(s_full + 10);  // [2][128] row sums, softmax -> epilogue
  uint8_t* s_out = smem + p.out_stage_off;             // [128 rows][128 B] bf16 O tile, 16-B chunks XOR-swizzled by row

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int b = 0; b < PP_MAX_STAGES; ++b) {
      mbar_init(&kv_full[b], 1);
      mbar_init(&kv_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 128);
      mbar_init(&e_done[b], 128);
    }
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    fence_mbar_init();
  }
  if (warp == 2) {
    __syncwarp();
    tmem_alloc<PP_TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int n_units = (p.num_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // units of this CTA
  const int n_tiles = n_units * p.mtiles;

  // Register allocation is per 128 threads: the control and epilogue warpgroups hand registers to the two softmax
  // warpgroups (128 x 56 + 128 x 104 + 256 x 176 = 65536).  The setmaxnreg sits inside each role branch so ptxas knows
  // which limit applies.
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      for (int u = 0; u < n_units; ++u) {
        const int bh = blockIdx.x + u * gridDim.x;
        const int bb = bh / p.H, h = bh - bb * p.H;
        const int sg = u % p.nstage;
        uint8_t* st = stage0 + sg * p.stage_bytes;
        mbar_wait(&kv_empty[sg], ((u / p.nstage) & 1) ^ 1u);
        mbar_arrive_expect_tx(&kv_full[sg], p.q_bytes + 2 * p.kv_bytes);
        tma_load_2d(st, &tmQ, &kv_full[sg], h * PP_HD, bb * p.S);
        tma_load_2d(st + p.q_bytes, &tmKV, &kv_full[sg], p.D + h * PP_HD, bb * p.S);
        tma_load_2d(st + p.q_bytes + p.kv_stride, &tmKV, &kv_full[sg], 2 * p.D + h * PP_HD, bb * p.S);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc_s = make_idesc_bf16(PP_M, p.Npad);
      const uint32_t idesc_o = make_idesc_bf16_b_mn(PP_M, PP_HD);
      const int ksteps = p.Npad >> 4;
      auto issue_s = [&](int t) {
        const int u = t / p.mtiles, mt = t - u * p.mtiles;
        const int sg = u % p.nstage;
        uint8_t* st = stage0 + sg * p.stage_bytes;
        mbar_wait(&kv_full[sg], (u / p.nstage) & 1);
        tcgen05_fence_after();
        const uint64_t qd = make_umma_desc_sw128(smem_u32(st + mt * PP_Q_TILE_BYTES));
        const uint64_t kd = make_umma_desc_sw128(smem_u32(st + p.q_bytes));
#pragma unroll
        for (int k = 0; k < PP_HD / 16; ++k)
          umma_bf16_ss(tmem_base + (t & 1) * p.nb, qd + 2u * k, kd + 2u * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[t & 1]);
      };
      if (n_tiles > 0) issue_s(0);
      if (n_tiles > 1) issue_s(1);
      for (int t = 0; t < n_tiles; ++t) {
        const int u = t / p.mtiles, mt = t - u * p.mtiles;
        const int sg = u % p.nstage;
        uint8_t* st = stage0 + sg * p.stage_bytes;
        mbar_wait(&p_full[t & 1], (t >> 1) & 1);  // P(t) is in TMEM and every S(t) read has retired
        mbar_wait(o_free, (t & 1) ^ 1u);          // O(t-1) has been drained
        tcgen05_fence_after();
        const uint64_t vd = make_umma_desc_mn_sw128(smem_u32(st + p.q_bytes + p.kv_stride), p.kv_stride);
        for (int k = 0; k < ksteps; ++k)  // 16 keys per step: 8 packed TMEM columns of P, 2048 B of V
          umma_bf16_ts(tmem_base + PP_O_COL, tmem_base + (t & 1) * p.nb + k * 8,
                       vd + static_cast<uint64_t>(k) * (2048u >> 4), idesc_o, k != 0 ? 1u : 0u);
        umma_commit(o_full);
        if (mt == p.mtiles - 1) umma_commit(&kv_empty[sg]);
        if (t + 2 < n_tiles) issue_s(t + 2);  // in order behind P.V(t): may overwrite S/P buffer t & 1
      }
    }
  }
  } else if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    // ===================== epilogue warpgroup: O / rowsum -> bf16 -> smem tile -> row-contiguous stores ==============
    const int wq = warp & 3;  // TMEM lane quarter
    const int r_local = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    long long tk[3] = {0, 0, 0};
    const bool dbg = p.debug == 1 && blockIdx.x == 0 && wq == 0 && lane == 0;
    for (int t = 0; t < n_tiles; ++t) {
      long long t0 = dbg ? clock64() : 0;
      const int u = t / p.mtiles, mt = t - u * p.mtiles;
      const int bh = blockIdx.x + u * gridDim.x;
      const int bb = bh / p.H, h = bh - bb * p.H;
      const bool warp_valid = (mt * PP_M + wq * 32) < p.S;  // warp-uniform
      mbar_wait(o_full, t & 1);
      tcgen05_fence_after();
      if (dbg) { long long x = clock64(); tk[0] += x - t0; t0 = x; }
      uint32_t o[64];
      if (warp_valid) {
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_base + lane_off + PP_O_COL, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld_32x32b_x32(tmem_base + lane_off + PP_O_COL + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
        tmem_wait_ld();
      }
      tcgen05_fence_before();
      mbar_arrive(o_free);  // O is in registers: the next P.V may start
      mbar_wait(&p_full[t & 1], (t >> 1) & 1);  // already complete (it precedes O); orders the read of the row sums
      const float l = s_l[(t & 1) * 128 + r_local];
      mbar_arrive(&e_done[t & 1]);
      epi_bar_sync();  // the previous tile's stores have read the staging tile
      if (dbg) { long long x = clock64(); tk[1] += x - t0; t0 = x; }
      if (warp_valid) {
        const float inv = l > 0.f ? __fdividef(1.f, l) : 0.f;
        uint8_t* srow = s_out + r_local * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(srow + ((q ^ (r_local & 7)) << 4)) = v;
        }
      }
      epi_bar_sync();  // the whole O tile is in smem
      {
        // 4 warps x 32 rows; one store instruction = 4 rows x 128 contiguous bytes (8 lanes per row)
        const int chunk = lane & 7;
        __nv_bfloat16* obase = p.out + ((int64_t)bb * p.S) * p.D + h * PP_HD + chunk * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = wq * 32 + k * 4 + (lane >> 3);
          const int qr = mt * PP_M + r;
          if (qr < p.S) {
            const uint4 v = *reinterpret_cast<const uint4*>(s_out + r * 128 + ((chunk ^ (r & 7)) << 4));
            st_v4(obase + (int64_t)qr * p.D, v);
          }
        }
      }
      if (dbg) { long long x = clock64(); tk[2] += x - t0; t0 = x; }
    }
    if (dbg && n_tiles > 0)
      printf("attn-pp dbg epilogue tiles %d cycles/tile: wait_o %lld ldO+sync %lld scale+store %lld\n", n_tiles,
             tk[0] / n_tiles, tk[1] / n_tiles, tk[2] / n_tiles);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    // ===================== softmax warpgroups: one thread per query row =====================
    const int wg = (warp - 4) >> 2;
    const int wq = warp & 3;  // TMEM lane quarter
    const int r_local = wq * 32 + lane;
    const float c = p.scale_log2e;
    const int nch = p.nb >> 5;  // 32-column chunks
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tb = tmem_base + lane_off + wg * p.nb;

    long long tk[5] = {0, 0, 0, 0, 0};
    long long tw[2] = {0, 0};
    const bool dbgw = p.debug == 1 && blockIdx.x == 0 && wq == 0;
    const bool dbg = dbgw && lane == 0;
    int ntl = 0;
    for (int t = wg; t < n_tiles; t += 2, ++ntl) {
      long long t0 = dbg ? clock64() : 0;
      const int u = t / p.mtiles, mt = t - u * p.mtiles;
      const int bh = blockIdx.x + u * gridDim.x;
      const int bb = bh / p.H;
      const int qrow = mt * PP_M + r_local;
      const bool warp_valid = (mt * PP_M + wq * 32) < p.S;  // warp-uniform
      const uint8_t* km = p.key_mask != nullptr ? p.key_mask + (int64_t)bb * p.S : nullptr;
      const int kmax_warp = p.causal ? min(p.S, mt * PP_M + wq * 32 + 32) : p.S;  // keys any row of this warp sees
      const uint32_t par = (t >> 1) & 1;

      mbar_wait(&s_full[wg], par);
      tcgen05_fence_after();
      if (dbg) { long long x = clock64(); tk[0] += x - t0; t0 = x; }
      float l = 0.f;
      if (warp_valid) {
        __syncwarp();
        // ---- single pass over S (TMEM reads are the scarce resource: ~64 B/clk/SM): softmax is shift invariant, so the
        // reference `off` only has to keep exp2 in range.  It starts as the maximum of the row's first visible chunk and
        // is raised by an INTEGER number of octaves (exact rescale of the P already written and of the row sum) only
        // when a later chunk exceeds it by more than 16 octaves: never for trained CLIP weights, still exact if it does.
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        float off = 0.f;
        bool has_ref = false;
        uint32_t sa[32], sb[32];
        tmem_ld_32x32b_x32(tb, sa);
        for (int ch = 0; ch < nch; ch += 2) {
          long long w0 = dbgw ? clock64() : 0;
          tmem_wait_ld();
          if (ch + 1 < nch) tmem_ld_32x32b_x32(tb + (ch + 1) * 32, sb);
          long long w1 = dbgw ? clock64() : 0;
          softmax_chunk<GENERAL_MASK>(p, sa, ch, kmax_warp, km, qrow, c, off, has_ref, l4, tb);
          long long w2 = dbgw ? clock64() : 0;
          tw[0] += w1 - w0;
          tw[1] += w2 - w1;
          if (ch + 1 < nch) {
            tmem_wait_ld();
            if (ch + 2 < nch) tmem_ld_32x32b_x32(tb + (ch + 2) * 32, sa);
            long long w3 = dbgw ? clock64() : 0;
            softmax_chunk<GENERAL_MASK>(p, sb, ch + 1, kmax_warp, km, qrow, c, off, has_ref, l4, tb);
            long long w4 = dbgw ? clock64() : 0;
            tw[0] += w3 - w2;
            tw[1] += w4 - w3;
          }
        }
        l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      }
      if (dbg) { long long x = clock64(); tk[3] += x - t0; t0 = x; }
      mbar_wait(&e_done[wg], par ^ 1u);  // the epilogue has read the row sums of tile t-2
      s_l[wg * 128 + r_local] = l;
      if (warp_valid) tmem_wait_st();
      tcgen05_fence_before();
      mbar_arrive(&p_full[wg]);
      if (dbg) { long long x = clock64(); tk[4] += x - t0; t0 = x; }
    }
    if (dbg && ntl > 0)
      printf("attn-pp dbg wg %d tiles %d cycles/tile: wait_s %lld softmax %lld (wait+ld %lld, chunks %lld) publish %lld\n", wg, ntl,
             tk[0] / ntl, tk[3] / ntl, tw[0] / ntl, tw[1] / ntl, tk[4] / ntl);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<PP_TMEM_COLS>(tmem_base);
  }
}

template <bool GENERAL_MASK>
int launch_pp(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const PPParams& p, size_t smem, int grid, cudaStream_t s) {
  static size_t smem_set = 0;
  if (smem > smem_set) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(attention_pp_kernel<GENERAL_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cudaError_t e = launch_pdl(attention_pp_kernel<GENERAL_MASK>, dim3(grid), dim3(PP_THREADS), smem, s, 1, tmQ, tmKV, p);
  if (e != cudaSuccess) {
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, attention_pp_kernel<GENERAL_MASK>);
    set_last_error("attention_pp_kernel launch: %s (regs=%d maxThreads=%d static_smem=%zu requested_dyn=%zu grid=%d)",
                   cudaGetErrorString(e), fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, smem, grid);
    return (int)e;
  }
  return 0;
}

}  // namespace

// S <= 224; called by vlmclip_attention_fwd (attention_tc.cu), arguments already validated there
int attention_fwd_pingpong(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H, int causal, float scale,
                           cudaStream_t s) {
  PPParams p;
  p.key_mask = key_mask;
  p.out = (__nv_bfloat16*)out;
  p.B = B;
  p.S = S;
  p.H = H;
  p.D = H * PP_HD;
  p.causal = causal;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.Npad = (S + 15) / 16 * 16;
  p.nb = (p.Npad + 31) / 32 * 32;
  p.mtiles = (S + PP_M - 1) / PP_M;
  p.num_units = B * H;
  p.q_bytes = (uint32_t)p.mtiles * PP_Q_TILE_BYTES;
  p.kv_bytes = (uint32_t)p.Npad * 128u;
  p.kv_stride = (p.kv_bytes + 1023u) & ~1023u;
  p.stage_bytes = p.q_bytes + 2 * p.kv_stride;
  {
    const char* e = getenv("VLMCLIP_ATTN_DEBUG");
    p.debug = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  int nstage = (int)((208u * 1024u) / p.stage_bytes);  // 227 KB - 16 KB output tile - barriers / row sums
  p.nstage = nstage < 2 ? 2 : (nstage > PP_MAX_STAGES ? PP_MAX_STAGES : nstage);
  const size_t ctrl = (2 * PP_MAX_STAGES + 10) * 8 + 256 * sizeof(float) + 16;
  p.out_stage_off = (uint32_t)(((size_t)p.nstage * p.stage_bytes + ctrl + 1023) & ~(size_t)1023);
  const size_t smem = (size_t)p.out_stage_off + PP_M * 128;
  if (smem > 232448) {
    set_last_error("attention: S=%d needs %zu bytes of shared memory", S, smem);
    return -1;
  }
  CUtensorMap tmQ, tmKV;
  const int64_t rows = (int64_t)B * S;
  int rc = make_tmap_bf16(&tmQ, qkv, rows, 3 * (int64_t)p.D, 3 * (int64_t)p.D, PP_M * p.mtiles);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmKV, qkv, rows, 3 * (int64_t)p.D, 3 * (int64_t)p.D, p.Npad);
  if (rc) return rc;
  const int grid = p.num_units < sm_count() ? p.num_units : sm_count();
  count_launch(1);
  const bool general = causal != 0 || key_mask != nullptr;
  return general ? launch_pp<true>(tmQ, tmKV, p, smem, grid, s) : launch_pp<false>(tmQ, tmKV, p, smem, grid, s);
}

}  // namespace vlmclip

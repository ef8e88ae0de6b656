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
//   warp 1     tcgen05: S = Q K^T (SS MMA, M=128, 4 k-steps) -> TMEM buffer t&1; O = P V (TS MMA: A = P from TMEM, B = V as an
//              MN-major smem operand) -> TMEM columns [448, 512).  With more than 128 keys S is issued in two parts, keys
//              [128, Npad) ("hi") and [0, 128) ("lo"), CAN be interleaved with the two halves of P.V(t-2) that free the
//              columns they overwrite: P.V_hi(t-2), S_hi(t), P.V_lo(t-2), S_lo(t), so that the softmax warpgroup starts on
//              S_hi after a fifth of the tensor work it otherwise waits for (VLMCLIP_ATTN_SSPLIT=1; measured slower, see
//              launch_range, and therefore off).  P(t) is written over the first 16 columns of each 32-column S chunk
//              (not compacted at the front), so that a chunk's P lives inside the column range its own half of S owns.
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
// attention_1q.cu: one query row per batch element, output row b at out + b * out_stride
int attention_1q_strided(const void* q, int64_t q_stride, const void* k, const void* v, int64_t kv_row_stride,
                         int64_t kv_batch_stride, void* out, int64_t out_stride, int B, int S, int H, float scale,
                         cudaStream_t stream);
// attention.cu: register-resident mma.sync kernel restricted to the query rows from q_begin on
int attention_fwd_mma_sync_rows(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H, int causal,
                                float scale, int q_begin, cudaStream_t stream);

namespace {

constexpr int PP_THREADS = 512;  // warps 0-3 control, 4-7 / 8-11 softmax warpgroups, 12-15 epilogue warpgroup
constexpr int PP_M = 128;
constexpr int PP_HD = 64;
constexpr uint32_t PP_Q_TILE_BYTES = PP_M * PP_HD * 2;  // 16 KB
constexpr int PP_TMEM_COLS = 512;
constexpr int PP_O_COL = 448;
constexpr int PP_MAX_STAGES = 4;

struct PPParams {
  const uint8_t* key_mask;
  __nv_bfloat16* out;
  int B, S, H, D;
  int causal;
  float scale_log2e;
  int Npad;    // keys rounded up to a multiple of 16
  int nb;      // TMEM columns per S buffer: Npad rounded up to 32
  int n_lo;    // keys [0, n_lo) form the "lo" part of S, [n_lo, Npad) the "hi" part (n_lo = Npad: no split)
  int mtiles;  // query tiles per unit (1 or 2)
  int num_units;
  uint32_t q_bytes;      // mtiles * 16 KB
  uint32_t kv_bytes;     // Npad * 128
  uint32_t kv_stride;    // kv_bytes rounded up to 1024
  uint32_t stage_bytes;  // q_bytes + 2 * kv_stride
  int nstage;
  uint32_t out_stage_off;  // 16 KB output staging tile
  int debug;
  // key-range split for 224 < S <= 432 (ViT-L/14, S = 257): a launch covers keys [key0, key0 + Sk) of every sequence.
  // The first launch (key0 = 0) also writes each row's softmax reference exponent and row sum to `stats`
  // ([B*S][H] float2), the second one (merge = 1) combines its own un-normalised O with the first launch's
  // normalised output in its epilogue: out = (O1 a1 + O2 2^(off2-m)) / (a1 + l2 2^(off2-m)), a1 = l1 2^(off1-m).
  // merge = 1: each thread loads its own row after O has arrived; merge = 2: the tile is staged through shared memory
  // with row-contiguous loads issued before the wait for O (the epilogue warpgroup is serial over tiles).
  // Sq: query rows [0, Sq) of every sequence are computed (Sq = S except when a caller handles tail rows elsewhere).
  int Sk, key0, merge, Sq;
  int q_loads;  // TMA loads per unit for Q (box = 128 * mtiles / q_loads rows; the box dimension limit is 256)
  float2* stats;
};

// visibility bits of keys [k0, k0+32) for query row qrow
template <bool GENERAL_MASK>
__device__ __forceinline__ uint32_t key_bits32(const PPParams& p, const uint8_t* km, int k0, int qrow) {
  uint32_t bits = 0u;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int key = k0 + j;
    bool ok = key < p.Sk;
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
  if (!GENERAL_MASK && k0 + 32 <= p.Sk) {
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
  } else if (!GENERAL_MASK && k0 + 32 <= p.Sk) {
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
__device__ __forceinline__ void softmax_chunk(const PPParams& p, const uint32_t (&cur)[32], int ch, int ch_first, int nch,
                                              int kmax_warp, const uint8_t* km, int qrow, float c, float& off,
                                              bool& has_ref, float (&l4)[4], uint32_t tb) {
  // chunks are processed in the order ch_first, ..., nch - 1, 0, ..., ch_first - 1 (hi part of S first, see the kernel)
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
    for (int j = 0; j < nch; ++j) {
      // already processed: in the hi phase (ch >= ch_first) the chunks [ch_first, ch), in the lo phase those and [0, ch)
      const bool done = ch >= ch_first ? (j >= ch_first && j < ch) : (j >= ch_first || j < ch);
      if (!done) continue;
      uint32_t pk[16];
      tmem_ld_32x32b_x16(tb + j * 32, pk);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t lo = __float_as_uint(__uint_as_float(pk[i] << 16) * f) & 0xffff0000u;
        const uint32_t hi = __float_as_uint(__uint_as_float(pk[i] & 0xffff0000u) * f) & 0xffff0000u;
        pk[i] = __byte_perm(lo, hi, 0x7632);
      }
      tmem_st_32x32b_x16(tb + j * 32, pk);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) l4[i] *= f;
    off += d;
  }
  exp_chunk<GENERAL_MASK>(p, cur, k0, kmax_warp, km, qrow, c, off, l4, tb + ch * 32);
}

// SPLIT: role of the launch in a two-launch key-range split (PPParams::stats / merge), compiled per role so that the
// plain kernel carries none of it: 0 = not split, 1 = first range (publishes the row statistics), 2 = second range
// with merge = 1, 3 = second range with merge = 2
template <bool GENERAL_MASK, int SPLIT>
__global__ void __launch_bounds__(PP_THREADS, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const PPParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage0 = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.nstage * p.stage_bytes);
  uint64_t* kv_full = bars;                     // [4] TMA landed Q, K, V of a unit
  uint64_t* kv_empty = bars + PP_MAX_STAGES;    // [4] the unit's last P.V finished reading the stage
  uint64_t* s_full = bars + 2 * PP_MAX_STAGES;  // [2] the lo part of S = Q K^T (all of it without a split) is complete
  uint64_t* p_full = s_full + 2;                // [2] softmax wrote P (128 arrivals)
  uint64_t* e_done = s_full + 4;                // [2] the epilogue has read the row sums of the buffer (128 arrivals)
  uint64_t* o_full = s_full + 6;                // [1] O = P V complete, in tile order
  uint64_t* o_free = s_full + 7;                // [1] O drained to registers (128 arrivals), in tile order
  uint64_t* s_hi_full = s_full + 8;             // [2] the hi part of S is complete (only used with a split)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 10);
  float* s_l = reinterpret_cast<float*>(s_full + 12);  // [2][128] row sums, softmax -> epilogue
  float* s_off = s_l + 256;                            // [2][128] reference exponents (key-range split only)
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
      mbar_init(&s_hi_full[b], 1);
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
        const uint32_t q_part = p.q_bytes / (uint32_t)p.q_loads;
        for (int ql = 0; ql < p.q_loads; ++ql)
          tma_load_2d(st + ql * q_part, &tmQ, &kv_full[sg], h * PP_HD, bb * p.S + ql * (int)(q_part >> 7));
        tma_load_2d(st + p.q_bytes, &tmKV, &kv_full[sg], p.D + h * PP_HD, bb * p.S + p.key0);
        tma_load_2d(st + p.q_bytes + p.kv_stride, &tmKV, &kv_full[sg], 2 * p.D + h * PP_HD, bb * p.S + p.key0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      const bool split = p.n_lo < p.Npad;
      const uint32_t idesc_lo = make_idesc_bf16(PP_M, p.n_lo);
      const uint32_t idesc_hi = make_idesc_bf16(PP_M, split ? p.Npad - p.n_lo : 16);
      const uint32_t idesc_o = make_idesc_bf16_b_mn(PP_M, PP_HD);
      const int ksteps = p.Npad >> 4;
      const int ksteps_lo = p.n_lo >> 4;
      // S(t), keys [key_b, key_b + N): K rows start key_b * 128 B into the K tile, columns start at key_b of the buffer
      auto issue_s_part = [&](int t, uint32_t idesc, int key_b, uint64_t* done_bar) {
        const int u = t / p.mtiles, mt = t - u * p.mtiles;
        const int sg = u % p.nstage;
        uint8_t* st = stage0 + sg * p.stage_bytes;
        mbar_wait(&kv_full[sg], (u / p.nstage) & 1);
        tcgen05_fence_after();
        const uint64_t qd = make_umma_desc_sw128(smem_u32(st + mt * PP_Q_TILE_BYTES));
        const uint64_t kd = make_umma_desc_sw128(smem_u32(st + p.q_bytes + key_b * 128));
#pragma unroll
        for (int k = 0; k < PP_HD / 16; ++k)
          umma_bf16_ss(tmem_base + (t & 1) * p.nb + key_b, qd + 2u * k, kd + 2u * k, idesc, k != 0 ? 1u : 0u);
        umma_commit(done_bar);
      };
      auto issue_s_hi = [&](int t) {
        if (split) issue_s_part(t, idesc_hi, p.n_lo, &s_hi_full[t & 1]);
      };
      auto issue_s_lo = [&](int t) { issue_s_part(t, idesc_lo, 0, &s_full[t & 1]); };
      for (int t = 0; t < 2 && t < n_tiles; ++t) {
        issue_s_hi(t);
        issue_s_lo(t);
      }
      for (int t = 0; t < n_tiles; ++t) {
        const int u = t / p.mtiles, mt = t - u * p.mtiles;
        const int sg = u % p.nstage;
        uint8_t* st = stage0 + sg * p.stage_bytes;
        mbar_wait(&p_full[t & 1], (t >> 1) & 1);  // P(t) is in TMEM and every S(t) read has retired
        mbar_wait(o_free, (t & 1) ^ 1u);          // O(t-1) has been drained
        tcgen05_fence_after();
        const uint64_t vd = make_umma_desc_mn_sw128(smem_u32(st + p.q_bytes + p.kv_stride), p.kv_stride);
        // 16 keys per step: 8 packed TMEM columns of P at the start of the keys' own S chunk half, 2048 B of V.
        // Key steps of the hi part first: they free the columns S_hi(t+2) overwrites.
        auto pv = [&](int k, uint32_t acc) {
          umma_bf16_ts(tmem_base + PP_O_COL, tmem_base + (t & 1) * p.nb + (k >> 1) * 32 + (k & 1) * 8,
                       vd + static_cast<uint64_t>(k) * (2048u >> 4), idesc_o, acc);
        };
        for (int k = ksteps_lo; k < ksteps; ++k) pv(k, k != ksteps_lo ? 1u : 0u);
        if (t + 2 < n_tiles) issue_s_hi(t + 2);  // in order behind P.V_hi(t): overwrites columns [n_lo, Npad) of buffer t & 1
        for (int k = 0; k < ksteps_lo; ++k) pv(k, (split || k != 0) ? 1u : 0u);
        umma_commit(o_full);
        if (mt == p.mtiles - 1) umma_commit(&kv_empty[sg]);
        if (t + 2 < n_tiles) issue_s_lo(t + 2);  // in order behind P.V_lo(t): overwrites columns [0, n_lo)
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
    // Second key range: the rows and statistics the merge folds in were written by the previous launch and stream from
    // DRAM (the attention output of a ViT-L/14 batch is larger than L2; measured ~3000 clk per tile when loaded on
    // demand by this serial warpgroup).  Each thread requests its row of a later tile into L2 ahead of time.
    constexpr int PF_TILES = 2;
    auto prefetch_merge_rows = [&](int tt) {
      if (tt >= n_tiles) return;
      const int pu = tt / p.mtiles, pmt = tt - pu * p.mtiles;
      const int pbh = blockIdx.x + pu * gridDim.x;
      const int pb = pbh / p.H, ph = pbh - pb * p.H;
      const int pr = pmt * PP_M + r_local;
      if (pr < p.Sq) {
        prefetch_l2(p.out + ((int64_t)pb * p.S + pr) * p.D + ph * PP_HD);
        prefetch_l2(p.stats + ((int64_t)pb * p.S + pr) * p.H + ph);
      }
    };
    if (SPLIT >= 2) {
      for (int tt = 0; tt < PF_TILES; ++tt) prefetch_merge_rows(tt);
    }
    for (int t = 0; t < n_tiles; ++t) {
      long long t0 = dbg ? clock64() : 0;
      if (SPLIT >= 2) prefetch_merge_rows(t + PF_TILES);
      const int u = t / p.mtiles, mt = t - u * p.mtiles;
      const int bh = blockIdx.x + u * gridDim.x;
      const int bb = bh / p.H, h = bh - bb * p.H;
      const bool warp_valid = (mt * PP_M + wq * 32) < p.Sq;  // warp-uniform
      if (SPLIT == 3) {
        // ---- second key range, staged variant: the first launch's tile is copied into the staging tile with the
        // row-contiguous access pattern of the stores below (4 rows x 128 B per warp instruction instead of 32 lines of
        // 16 B), requested before the wait for O; each thread then folds its own row in place, O in two 32-column halves
        const int qr = mt * PP_M + r_local;
        const int chunk = lane & 7;
        float2 st1 = make_float2(0.f, 0.f);
        if (warp_valid && qr < p.Sq) st1 = p.stats[((int64_t)bb * p.S + qr) * p.H + h];  // (off1, l1)
        uint4 pv[8];
        {
          const __nv_bfloat16* pbase = p.out + ((int64_t)bb * p.S) * p.D + h * PP_HD + chunk * 8;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int pr = mt * PP_M + wq * 32 + k * 4 + (lane >> 3);
            pv[k] = make_uint4(0u, 0u, 0u, 0u);
            if (pr < p.Sq) pv[k] = *reinterpret_cast<const uint4*>(pbase + (int64_t)pr * p.D);
          }
        }
        epi_bar_sync();  // the previous tile's stores have read the staging tile
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = wq * 32 + k * 4 + (lane >> 3);
          *reinterpret_cast<uint4*>(s_out + r * 128 + ((chunk ^ (r & 7)) << 4)) = pv[k];
        }
        mbar_wait(o_full, t & 1);
        tcgen05_fence_after();
        if (dbg) { long long x = clock64(); tk[0] += x - t0; t0 = x; }
        mbar_wait(&p_full[t & 1], (t >> 1) & 1);  // already complete (it precedes O); orders the read of the row sums
        const float l = s_l[(t & 1) * 128 + r_local];
        const float off2 = s_off[(t & 1) * 128 + r_local];
        mbar_arrive(&e_done[t & 1]);
        const bool h1 = st1.y > 0.f, h2 = l > 0.f;
        const float m = fmaxf(h1 ? st1.x : -INFINITY, h2 ? off2 : -INFINITY);
        const float a1 = h1 ? st1.y * exp2f(st1.x - m) : 0.f;
        const float s2 = h2 ? exp2f(off2 - m) : 0.f;
        const float den = a1 + l * s2;
        const float inv = den > 0.f ? __fdividef(1.f, den) : 0.f;
        const float w1 = a1 * inv, w2 = s2 * inv;
        epi_bar_sync();  // the first launch's tile is staged
        uint8_t* srow = s_out + r_local * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t oh[32];
          if (warp_valid) {
            __syncwarp();
            tmem_ld_32x32b_x32(tmem_base + lane_off + PP_O_COL + half * 32, oh);
            tmem_wait_ld();
          }
          if (half == 1) {
            tcgen05_fence_before();
            mbar_arrive(o_free);  // O is in registers: the next P.V may start
          }
          if (warp_valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4* slot = reinterpret_cast<uint4*>(srow + (((half * 4 + q) ^ (r_local & 7)) << 4));
              const uint4 pq = *slot;
              const uint32_t pw[4] = {pq.x, pq.y, pq.z, pq.w};
              uint32_t ow[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float lo = fmaf(bf16_lo(pw[e]), w1, __uint_as_float(oh[q * 8 + 2 * e]) * w2);
                const float hi = fmaf(bf16_hi(pw[e]), w1, __uint_as_float(oh[q * 8 + 2 * e + 1]) * w2);
                ow[e] = pack_bf16x2(lo, hi);
              }
              *slot = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
          }
        }
        if (dbg) { long long x = clock64(); tk[1] += x - t0; t0 = x; }
      } else {
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
      const float off2 = SPLIT != 0 ? s_off[(t & 1) * 128 + r_local] : 0.f;
      mbar_arrive(&e_done[t & 1]);
      epi_bar_sync();  // the previous tile's stores have read the staging tile
      if (dbg) { long long x = clock64(); tk[1] += x - t0; t0 = x; }
      if (SPLIT == 2 && warp_valid) {
        // second key range: fold the first launch's normalised rows (already in `out`) into this tile
        const int qr = mt * PP_M + r_local;
        uint8_t* srow = s_out + r_local * 128;
        float a1 = 0.f, s2 = 0.f, inv = 0.f;
        const __nv_bfloat16* prev = p.out + ((int64_t)bb * p.S + (qr < p.Sq ? qr : 0)) * p.D + h * PP_HD;
        if (qr < p.Sq) {
          const float2 st1 = p.stats[((int64_t)bb * p.S + qr) * p.H + h];  // (off1, l1)
          const bool h1 = st1.y > 0.f, h2 = l > 0.f;
          const float m = fmaxf(h1 ? st1.x : -INFINITY, h2 ? off2 : -INFINITY);
          a1 = h1 ? st1.y * exp2f(st1.x - m) : 0.f;
          s2 = h2 ? exp2f(off2 - m) : 0.f;
          const float den = a1 + l * s2;
          inv = den > 0.f ? __fdividef(1.f, den) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 pv = make_uint4(0u, 0u, 0u, 0u);
          if (qr < p.Sq) pv = *reinterpret_cast<const uint4*>(prev + q * 8);
          const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
          uint32_t ow[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = fmaf(bf16_lo(pw[e]), a1, __uint_as_float(o[q * 8 + 2 * e]) * s2) * inv;
            const float hi = fmaf(bf16_hi(pw[e]), a1, __uint_as_float(o[q * 8 + 2 * e + 1]) * s2) * inv;
            ow[e] = pack_bf16x2(lo, hi);
          }
          *reinterpret_cast<uint4*>(srow + ((q ^ (r_local & 7)) << 4)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
      } else if (warp_valid) {
        if (SPLIT == 1) {
          const int qr = mt * PP_M + r_local;
          if (qr < p.Sq) p.stats[((int64_t)bb * p.S + qr) * p.H + h] = make_float2(off2, l);
        }
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
          if (qr < p.Sq) {
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
      const bool warp_valid = (mt * PP_M + wq * 32) < p.Sq;  // warp-uniform
      const uint8_t* km = p.key_mask != nullptr ? p.key_mask + (int64_t)bb * p.S : nullptr;
      const int kmax_warp = p.causal ? min(p.Sk, mt * PP_M + wq * 32 + 32) : p.Sk;  // keys any row of this warp sees
      const uint32_t par = (t >> 1) & 1;

      const int ch_first = p.n_lo < p.Npad ? (p.n_lo >> 5) : 0;  // first chunk of the hi part of S (0: no split)
      mbar_wait(ch_first > 0 ? &s_hi_full[wg] : &s_full[wg], par);
      tcgen05_fence_after();
      if (dbg) { long long x = clock64(); tk[0] += x - t0; t0 = x; }
      float l = 0.f;
      float off_pub = 0.f;
      {
        // ---- single pass over S (TMEM reads are the scarce resource: ~64 B/clk/SM): softmax is shift invariant, so the
        // reference `off` only has to keep exp2 in range.  It starts as the maximum of the row's first visible chunk and
        // is raised by an INTEGER number of octaves (exact rescale of the P already written and of the row sum) only
        // when a later chunk exceeds it by more than 16 octaves: never for trained CLIP weights, still exact if it does.
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        float off = 0.f;
        bool has_ref = false;
        uint32_t sa[32], sb[32];
        // chunks [c0, c1) with the next chunk's tcgen05.ld in flight under the current chunk's arithmetic
        auto run_chunks = [&](int c0, int c1) {
          if (!warp_valid || c0 >= c1) return;
          __syncwarp();
          tmem_ld_32x32b_x32(tb + c0 * 32, sa);
          for (int ch = c0; ch < c1; ch += 2) {
            long long w0 = dbgw ? clock64() : 0;
            tmem_wait_ld();
            if (ch + 1 < c1) tmem_ld_32x32b_x32(tb + (ch + 1) * 32, sb);
            long long w1 = dbgw ? clock64() : 0;
            softmax_chunk<GENERAL_MASK>(p, sa, ch, ch_first, nch, kmax_warp, km, qrow, c, off, has_ref, l4, tb);
            long long w2 = dbgw ? clock64() : 0;
            tw[0] += w1 - w0;
            tw[1] += w2 - w1;
            if (ch + 1 < c1) {
              tmem_wait_ld();
              if (ch + 2 < c1) tmem_ld_32x32b_x32(tb + (ch + 2) * 32, sa);
              long long w3 = dbgw ? clock64() : 0;
              softmax_chunk<GENERAL_MASK>(p, sb, ch + 1, ch_first, nch, kmax_warp, km, qrow, c, off, has_ref, l4, tb);
              long long w4 = dbgw ? clock64() : 0;
              tw[0] += w3 - w2;
              tw[1] += w4 - w3;
            }
          }
        };
        if (ch_first > 0) {
          run_chunks(ch_first, nch);       // hi part: available after P.V_hi(t-2) + S_hi(t) only
          mbar_wait(&s_full[wg], par);     // lo part: issued behind P.V_lo(t-2), long complete by now
          tcgen05_fence_after();
        }
        run_chunks(0, ch_first > 0 ? ch_first : nch);
        l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        off_pub = off;
      }
      if (dbg) { long long x = clock64(); tk[3] += x - t0; t0 = x; }
      mbar_wait(&e_done[wg], par ^ 1u);  // the epilogue has read the row sums of tile t-2
      s_l[wg * 128 + r_local] = l;
      if (SPLIT != 0) s_off[wg * 128 + r_local] = off_pub;
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

template <bool GENERAL_MASK, int SPLIT>
int launch_pp(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const PPParams& p, size_t smem, int grid, cudaStream_t s) {
  static size_t smem_set = 0;
  if (smem > smem_set) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(attention_pp_kernel<GENERAL_MASK, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cudaError_t e = launch_pdl(attention_pp_kernel<GENERAL_MASK, SPLIT>, dim3(grid), dim3(PP_THREADS), smem, s, 1, tmQ, tmKV, p);
  if (e != cudaSuccess) {
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, attention_pp_kernel<GENERAL_MASK, SPLIT>);
    set_last_error("attention_pp_kernel launch: %s (regs=%d maxThreads=%d static_smem=%zu requested_dyn=%zu grid=%d)",
                   cudaGetErrorString(e), fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, smem, grid);
    return (int)e;
  }
  return 0;
}

// One launch over keys [key0, key0 + Sk) of every sequence (Sk <= 224).  stats / merge: see PPParams.
int launch_range(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int Sq, int H, int causal,
                 float scale, int key0, int Sk, float2* stats, int merge, cudaStream_t s) {
  PPParams p;
  p.key_mask = key_mask;
  p.out = (__nv_bfloat16*)out;
  p.B = B;
  p.S = S;
  p.H = H;
  p.D = H * PP_HD;
  p.causal = causal;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.Sk = Sk;
  p.Sq = Sq;
  p.key0 = key0;
  p.merge = merge;
  p.stats = stats;
  p.Npad = (Sk + 15) / 16 * 16;
  p.nb = (p.Npad + 31) / 32 * 32;
  {
    // S in two parts when there are more than 128 keys: OFF by default.  Measured on B200 (profiles/
    // r02_attention_s_split_ab.txt, B = 256, S = 197, H = 12): the wait for S drops from 2.9 k to 1.6 k cycles per tile as
    // intended, but the chunk loop that now runs concurrently with the other half's MMAs slows from 6.0 k to 8.8 k cycles
    // (TMEM port and issue slots are shared with the tensor pipe's operand reads): 166 us against 132 us.
    // VLMCLIP_ATTN_SSPLIT=1 enables it for A/B measurements.
    static const bool ssplit = []() {
      const char* e = getenv("VLMCLIP_ATTN_SSPLIT");
      return e != nullptr && e[0] == '1';
    }();
    p.n_lo = (ssplit && p.Npad > 128) ? 128 : p.Npad;
  }
  p.mtiles = (Sq + PP_M - 1) / PP_M;
  p.q_loads = p.mtiles <= 2 ? 1 : p.mtiles;
  p.num_units = B * H;
  p.q_bytes = (uint32_t)p.mtiles * PP_Q_TILE_BYTES;
  p.kv_bytes = (uint32_t)p.Npad * 128u;
  p.kv_stride = (p.kv_bytes + 1023u) & ~1023u;
  p.stage_bytes = p.q_bytes + 2 * p.kv_stride;
  {
    const char* e = getenv("VLMCLIP_ATTN_DEBUG");
    p.debug = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  int nstage = (int)((206u * 1024u) / p.stage_bytes);  // 227 KB - 16 KB output tile - barriers / row sums / exponents
  p.nstage = nstage < 2 ? 2 : (nstage > PP_MAX_STAGES ? PP_MAX_STAGES : nstage);
  const size_t ctrl = (2 * PP_MAX_STAGES + 12) * 8 + 512 * sizeof(float) + 16;
  p.out_stage_off = (uint32_t)(((size_t)p.nstage * p.stage_bytes + ctrl + 1023) & ~(size_t)1023);
  const size_t smem = (size_t)p.out_stage_off + PP_M * 128;
  if (smem > 232448) {
    set_last_error("attention: S=%d (keys %d..%d) needs %zu bytes of shared memory", S, key0, key0 + Sk, smem);
    return -1;
  }
  CUtensorMap tmQ, tmKV;
  const int64_t rows = (int64_t)B * S;
  int rc = make_tmap_bf16(&tmQ, qkv, rows, 3 * (int64_t)p.D, 3 * (int64_t)p.D, PP_M * p.mtiles / p.q_loads);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmKV, qkv, rows, 3 * (int64_t)p.D, 3 * (int64_t)p.D, p.Npad);
  if (rc) return rc;
  const int grid = p.num_units < sm_count() ? p.num_units : sm_count();
  count_launch(1);
  const bool general = causal != 0 || key_mask != nullptr;
  if (stats != nullptr) {
    if (general) {
      set_last_error("attention: the key-range split does not take a mask");
      return -1;
    }
    return merge == 0   ? launch_pp<false, 1>(tmQ, tmKV, p, smem, grid, s)
           : merge == 1 ? launch_pp<false, 2>(tmQ, tmKV, p, smem, grid, s)
                        : launch_pp<false, 3>(tmQ, tmKV, p, smem, grid, s);
  }
  return general ? launch_pp<true, 0>(tmQ, tmKV, p, smem, grid, s) : launch_pp<false, 0>(tmQ, tmKV, p, smem, grid, s);
}

}  // namespace

// S <= 224; called by vlmclip_attention_fwd (attention_tc.cu), arguments already validated there
int attention_fwd_pingpong(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H, int causal, float scale,
                           cudaStream_t s) {
  return launch_range(qkv, out, key_mask, B, S, S, H, causal, scale, 0, S, nullptr, 0, s);
}

// 224 < S <= 384, no mask (ViT-L/14: S = 257): two launches over the key ranges [0, 208) and [208, S); the second one
// merges in its epilogue (see PPParams).  workspace: 2 * B * S * H floats (reference exponent and row sum per row and
// head).  208 keys, not 224: with three 128-query tiles per unit two pipeline stages of Q + K + V must fit 227 KB.
// A sequence that is a few rows longer than a multiple of 128 (257 = 2 * 128 + 1) would spend a whole 128-row tile
// slot per unit on those rows: they go to the one-warp-per-row kernel of attention_1q.cu instead (variant 2).
//   variant 1: every row on the tcgen05 kernel, merge = 1
//   variant 2: tail rows on the single-query kernel, merge = 2
//   variant 3: every row on the tcgen05 kernel, merge = 2
constexpr int PP_SPLIT_KEYS = 208;
constexpr int PP_MAX_TAIL_ROWS = 2;
int attention_fwd_pingpong_split(const void* qkv, void* out, float* workspace, int B, int S, int H, float scale,
                                 int variant, cudaStream_t s) {
  const int tail = S % PP_M;
  // variant 4: a sequence a few rows longer than a multiple of 128 (257 = 2 * 128 + 1) keeps its full tiles on the
  // tcgen05 kernel and sends the tail rows to the mma.sync kernel (one 16-row block per unit) instead of spending a
  // third 128-row tile slot of BOTH launches on them
  const bool tail_mma = variant == 4 && tail >= 1 && tail <= 16;
  const int Sq = ((variant == 2 && tail >= 1 && tail <= PP_MAX_TAIL_ROWS) || tail_mma) ? S - tail : S;
  float2* stats = reinterpret_cast<float2*>(workspace);
  int rc = launch_range(qkv, out, nullptr, B, S, Sq, H, 0, scale, 0, PP_SPLIT_KEYS, stats, 0, s);
  if (rc) return rc;
  rc = launch_range(qkv, out, nullptr, B, S, Sq, H, 0, scale, PP_SPLIT_KEYS, S - PP_SPLIT_KEYS, stats,
                    variant == 1 ? 1 : 2, s);
  if (rc) return rc;
  if (tail_mma) return attention_fwd_mma_sync_rows(qkv, out, nullptr, B, S, H, 0, scale, Sq, s);
  const int64_t D = (int64_t)H * PP_HD;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  for (int r = Sq; r < S; ++r) {
    rc = attention_1q_strided(base + r * 3 * D, S * 3 * D, base + D, base + 2 * D, 3 * D, S * 3 * D,
                              static_cast<__nv_bfloat16*>(out) + r * D, S * D, B, S, H, scale, s);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace vlmclip

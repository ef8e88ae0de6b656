// Dense layer GEMM for the frozen CLIP towers:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
// Replaces the aten::addmm calls behind nn.Linear at HF modeling_clip.py:310-312,334 (q/k/v/out),
// :348-350 (fc1 / quick_gelu / fc2) and :209 (patch conv as GEMM).
//
// B200 design (persistent over output tiles; one CTA per SM, normally as 2-CTA clusters):
//   PAIR               two CTAs of a cluster share one 256 x 256 tile: each loads its 128 A rows and HALF of the W rows,
//                      the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) for both and multicasts the commits, so
//                      every SM pulls half the W bytes through shared memory.  Single-CTA variants (128 x 256,
//                      128 x 128) remain for M <= 128.
//   warp 0 (1 thread)  TMA producer: A tile [128 x 64] and W tile [BLOCK_N(/2) x 64] per stage,
//                      128-byte swizzle, completion by mbarrier complete_tx.
//   warp 1 (1 thread)  tcgen05.mma issuer (K = 16 x 4 per stage), fp32 accumulators in TMEM, double buffered
//                      (2 x BLOCK_N columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//   warp 2             TMEM allocator / deallocator; lane 0 = panel manager of epilogue group 0.
//   warp 3 (1 thread)  panel manager of epilogue group 1.
//   warps 4..11        epilogue, two groups of 4 warps (each group covers the 128 accumulator rows and
//                      owns every other 64-column quarter of the tile): tcgen05.ld -> registers ->
//                      LN-fold / bias / activation (vectors staged in smem once per tile) -> + residual
//                      (prefetched by TMA into a 64-B-swizzled 128 x 32 panel) -> bf16 written in place
//                      into the panel -> TMA store.  Panel managers chain store -> wait-read -> next
//                      residual load so global traffic of the epilogue is fully asynchronous.
//   EPI                the epilogue is specialised at compile time (fold / fold+gelu / residual / plain / generic):
//                      the fully unrolled run-time-switched version thrashed the instruction cache.
//   Tile order is n-fastest so the CTAs running at one moment share a few A row-blocks through L2.
//   Launched with programmatic dependent launch: the prologue overlaps the previous kernel's tail.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace vlmclip {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 256;
constexpr int EPI_GROUP_THREADS = 128;      // one group = 4 warps = all 128 accumulator rows
constexpr int QUARTER_N = 64;               // staged epilogue unit: 128 rows x 64 columns (one swizzle panel)
constexpr int PANEL_N = 32;                 // staged epilogue unit: 128 rows x 32 columns (64-byte swizzle span)
constexpr uint32_t EBUF_BYTES = BLOCK_M * PANEL_N * 2;  // 8 KB

struct GemmParams {
  void* C;
  const float* bias;
  const __nv_bfloat16* residual;
  const float* row_stats;  // [M][2] mean, rstd
  const float* col_c;      // [N]
  const float* part_in;    // [M][npart_in][2] (mean, M2) of 32-column blocks of the A operand's rows (LN fold)
  float* part_out;         // [M][N/32][2] same statistics of the rows this GEMM writes (for the next LN-folded layer)
  // Fused combine of part_out: the CTA that finishes the LAST column tile of a 128-row block (counted in row_counters,
  // one zero-initialised word per row block, left zero again) turns the block's partials into (mean, rstd) rows of
  // stats_out [M][2] - what a separate ln_partials_to_stats launch did (46 launches per ViT-B/16 step).
  float* stats_out;
  unsigned* row_counters;
  int npart_in;
  float ln_eps;
  int64_t ldc, ldr;
  int M, N, K;
  int act;
  int out_fp32;
  int debug;   // VLMCLIP_GEMM_DEBUG=1: one epilogue warp prints per-phase cycle totals (development aid)
  int staged;  // 1: residual in / result out go through swizzled smem panels and TMA (bf16 output only)
  int int_pack;
  int m_tiles, n_tiles, k_blocks;
  // Split of the reduction dimension across work items (weight-gradient GEMMs: few output tiles, K = all tokens):
  // work item = (split, output tile); split s reduces k-blocks [s * kb_per_split, ...) and writes its own fp32 plane
  // C + s * split_stride (direct fp32 epilogue only); a second pass adds the planes in index order.  ksplit = 1: off.
  int ksplit, kb_per_split;
  int64_t split_stride;
};

// EB = panel buffers per epilogue group.  EB = 1: one 16 KB panel per group (residual in / result out chained through
// it).  EB = 2: two panels per group, so the residual panel of the next quarter is prefetched while the current one is
// processed (used for the residual layers, traded against one pipeline stage of the main loop).
// RP = planes per panel: 1 (bf16 residual / result) or 2 (two-term hi + lo residual stream, EPI_RES2).
template <int BLOCK_N, int STAGES, int EB, bool PAIR = false, int RP = 1>
struct SmemLayout {
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;  // a CTA pair splits W along N
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t PANEL_BYTES = RP * EBUF_BYTES;                    // 8 KB per plane
  static constexpr uint32_t EBUF_OFFSET = STAGES * STAGE_BYTES;               // 2 * EB panels, 1024-aligned
  static constexpr uint32_t VEC_OFFSET = EBUF_OFFSET + 2 * EB * PANEL_BYTES;  // bias[BLOCK_N], col_c[BLOCK_N] fp32
  static constexpr uint32_t BAR_OFFSET = VEC_OFFSET + 2 * BLOCK_N * 4;
  static constexpr uint32_t NUM_BARS = 2 * STAGES + 4 + 4 * EB;
  static constexpr uint32_t DYN_BYTES = BAR_OFFSET + NUM_BARS * 8 + 16;
  static_assert(DYN_BYTES <= 232448, "shared memory budget exceeded");
};

// bf16x2 pack with round-to-nearest-even on the integer pipe (F2FP runs on the SFU pipe, which the quick_gelu
// epilogue already loads with one MUFU.TANH per element)
__device__ __forceinline__ uint32_t pack_bf16x2_int(float lo, float hi) {
  uint32_t a = __float_as_uint(lo), b = __float_as_uint(hi);
  a += 0x7fffu + ((a >> 16) & 1u);
  b += 0x7fffu + ((b >> 16) & 1u);
  return __byte_perm(a, b, 0x7632);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return quick_gelu(v);
  if (act == 2) return gelu_erf(v);
  if (act == 3) return fmaxf(v, 0.0f);
  return v;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// v[0..7] = epilogue(acc) for 8 consecutive columns starting at tile-local column lc
__device__ __forceinline__ void epi_math8(float* v, const float* sBias, const float* sColc, int lc, float a_scale,
                                          float a_shift, bool has_bias, bool has_stats, int act) {
  if (has_stats) {
    // rstd*(acc - mean*c) + b  ==  fma(rstd, acc, fma(-mean*rstd, c, b))
    const float4 c0 = *reinterpret_cast<const float4*>(sColc + lc);
    const float4 c1 = *reinterpret_cast<const float4*>(sColc + lc + 4);
    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    float bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (has_bias) {
      const float4 b0 = *reinterpret_cast<const float4*>(sBias + lc);
      const float4 b1 = *reinterpret_cast<const float4*>(sBias + lc + 4);
      bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
      bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(a_scale, v[j], fmaf(a_shift, cc[j], bb[j]));
  } else if (has_bias) {
    const float4 b0 = *reinterpret_cast<const float4*>(sBias + lc);
    const float4 b1 = *reinterpret_cast<const float4*>(sBias + lc + 4);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += bb[j];
  }
  if (act != 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], act);
  }
}

// PAIR = true: the kernel is launched in clusters of two CTAs (cta_group::2).  Each CTA owns 128 rows of a 256-row
// tile (its own A tile, accumulator, epilogue) and loads HALF of the W tile; the leader CTA's MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both halves.  W bytes per SM and the smem operand traffic halve,
// which also frees shared memory for more pipeline stages.
// EPI selects a compile-time epilogue so each instantiation carries only the code it executes (the fully unrolled
// generic epilogue is ~90 KB of SASS and thrashes the instruction cache: stall_no_inst was a top stall reason):
//   EPI_GENERIC   everything decided at run time (tests, odd combinations, fp32 / direct output)
//   EPI_FOLD      LN fold + bias                     (QKV)
//   EPI_FOLD_ACT  LN fold + bias + quick_gelu        (fc1)
//   EPI_RES       bias + residual + LN partials out  (out-proj, fc2)
//   EPI_PLAIN     bf16 store only                    (patch embedding)
//   EPI_RES2      bias + TWO-TERM residual + LN partials out (out-proj, fc2 of the towers): the residual stream is
//                 kept as x = hi + lo, two bf16 planes [2][M][N]; hi (= bf16(x)) is what the next GEMM reads as its A
//                 operand, lo (= bf16(x - hi)) is touched by these epilogues only.  16 mantissa bits instead of 8 on the
//                 quantity that is accumulated over 2 x L layers: the end-to-end error drops from 9e-3 to the 3.5e-3 of
//                 the bf16 operand roundings (oracle/emulate_bf16.py).  A panel is then [2 planes][128 rows][32 columns],
//                 moved by ONE 3-D TMA box per direction.
enum { EPI_GENERIC = 0, EPI_FOLD = 1, EPI_FOLD_ACT = 2, EPI_RES = 3, EPI_PLAIN = 4, EPI_RES2 = 5 };

// OPMN = true: both operands are MN-major in global memory, A_t [K, M] and B_t [K, N] (row-major, K = rows), i.e.
// C = A_t^T B_t - the weight-gradient form dW = dY^T X with dY [tokens, N_out], X [tokens, K_in] read as they lie, no
// transposed copies.  A stage then holds [64 k rows][64 contiguous m] atoms of 8 KB (128-B swizzle), two per 128 rows of
// M or N; the MMA descriptors are the MN-major ones (LBO = 8 KB between atoms, SBO = 1 KB between 8-row k groups).
template <int BLOCK_N, int STAGES, int EB, bool PAIR, int EPI, bool OPMN = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                    const GemmParams p) {
  constexpr int RP = EPI == EPI_RES2 ? 2 : 1;
  using L = SmemLayout<BLOCK_N, STAGES, EB, PAIR, RP>;
  constexpr uint32_t PANEL_BYTES = L::PANEL_BYTES;
  constexpr int TMEM_COLS = 2 * BLOCK_N;  // double-buffered accumulator (power of two: 256 or 512)
  constexpr int QUARTERS = BLOCK_N / QUARTER_N;
  extern __shared__ __align__(1024) uint8_t smem[];

  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::A_BYTES;
  uint8_t* sE = smem + L::EBUF_OFFSET;
  float* sBias = reinterpret_cast<float*>(smem + L::VEC_OFFSET);
  float* sColc = sBias + BLOCK_N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint64_t* res_full_bar = bars + 2 * STAGES + 4;            // [2][EB] panel is free (and holds the residual)
  uint64_t* e_written_bar = bars + 2 * STAGES + 4 + 2 * EB;  // [2][EB] the group has written its result panel
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);
  volatile uint32_t* s_last_tile = tmem_slot + 1;  // epilogue: "this CTA completed its row block" (fused LN statistics)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;  // 0 = leader
  // epilogue switches: compile-time constants for the specialised instantiations
  const bool k_staged = EPI == EPI_GENERIC ? p.staged != 0 : true;
  const bool k_has_res = EPI == EPI_GENERIC ? p.residual != nullptr : (EPI == EPI_RES || EPI == EPI_RES2);
  const bool k_has_bias = EPI == EPI_GENERIC ? p.bias != nullptr : EPI != EPI_PLAIN;
  const bool k_has_stats = EPI == EPI_GENERIC ? (p.row_stats != nullptr || p.part_in != nullptr)
                                              : (EPI == EPI_FOLD || EPI == EPI_FOLD_ACT);
  const bool k_part_out =
      EPI == EPI_GENERIC ? p.part_out != nullptr : ((EPI == EPI_RES || EPI == EPI_RES2) && p.part_out != nullptr);
  const int k_act = EPI == EPI_GENERIC ? p.act : (EPI == EPI_FOLD_ACT ? 1 : 0);
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // persistent worker index (CTA or CTA pair)
  const int n_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) __trap();  // swizzled tiles need a 1024-B aligned base
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (k_staged) {
      tma_prefetch_desc(&tmC);
      if (k_has_res) tma_prefetch_desc(&tmR);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);  // pair: only the leader arrives (expecting both CTAs' bytes); see the producer
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], PAIR ? 2 * EPI_THREADS : EPI_THREADS);  // pair: both CTAs' epilogues release the leader
    }
    for (int b = 0; b < 2 * EB; ++b) {
      mbar_init(&res_full_bar[b], 1);
      mbar_init(&e_written_bar[b], EPI_GROUP_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR)
      tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    else
      tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  if (PAIR)
    cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  else
    __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // prologue done; everything below touches global memory
  pdl_trigger();  // the next kernel in the stream may start its own prologue as SMs free up

  // pair: tiles are 256 rows tall (m index counts pair tiles); this CTA's 128-row block is 2*m + rank
  const int tiles_mn = (PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
  const int num_tiles = tiles_mn * p.ksplit;  // work items: (reduction split, output tile)

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0;
    for (int tile = unit; tile < num_tiles; tile += n_units) {
      const int split = tile / tiles_mn, t_mn = tile - split * tiles_mn;
      const int m_t = t_mn / p.n_tiles;
      const int n_blk = t_mn - m_t * p.n_tiles;
      const int m_blk = PAIR ? 2 * m_t + (int)cta_rank : m_t;
      const int kb0 = split * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (PAIR) {
          // Both CTAs' loads complete on the LEADER's full barrier; the leader alone arrives, expecting the bytes of both
          // CTAs.  The peer needs no arrival of its own: its loads for the next round of this stage are issued only
          // after its empty barrier fired, i.e. after the leader's MMAs consumed (hence completed) the current phase,
          // and bytes landing before the leader's expect_tx merely drive the tx-count negative for a moment.
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
          if (OPMN) {
            // [64 k][64 m] atoms: inner coordinate = m (or n), outer = k
#pragma unroll
            for (int h = 0; h < BLOCK_M / 64; ++h)
              tma_load_2d_pair(sA + stage * L::A_BYTES + h * 8192, &tmA, &full_bar[stage], m_blk * BLOCK_M + 64 * h, kb * BLOCK_K);
#pragma unroll
            for (int h = 0; h < BLOCK_N / 128; ++h)
              tma_load_2d_pair(sB + stage * L::B_BYTES + h * 8192, &tmB, &full_bar[stage],
                               n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / 2) + 64 * h, kb * BLOCK_K);
          } else {
            tma_load_2d_pair(sA + stage * L::A_BYTES, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d_pair(sB + stage * L::B_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K,
                             n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / 2));
          }
        } else {
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if (OPMN) {
#pragma unroll
            for (int h = 0; h < BLOCK_M / 64; ++h)
              tma_load_2d(sA + stage * L::A_BYTES + h * 8192, &tmA, &full_bar[stage], m_blk * BLOCK_M + 64 * h, kb * BLOCK_K);
#pragma unroll
            for (int h = 0; h < BLOCK_N / 64; ++h)
              tma_load_2d(sB + stage * L::B_BYTES + h * 8192, &tmB, &full_bar[stage], n_blk * BLOCK_N + 64 * h, kb * BLOCK_K);
          } else {
            tma_load_2d(sA + stage * L::A_BYTES, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d(sB + stage * L::B_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          }
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer (pair: leader CTA only) =====================
    if (!PAIR || cta_rank == 0) {
      constexpr uint32_t idesc =
          make_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, BLOCK_N) | (OPMN ? ((1u << 15) | (1u << 16)) : 0u);  // A, B MN-major
      uint32_t stage = 0, phase = 0;
      uint32_t abuf = 0, aphase = 0;
      for (int tile = unit; tile < num_tiles; tile += n_units) {
        mbar_wait(&tempty_bar[abuf], aphase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + abuf * BLOCK_N;
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t a_desc = OPMN ? make_umma_desc_mn_sw128(smem_u32(sA + stage * L::A_BYTES), 8192u)
                                       : make_umma_desc_sw128(smem_u32(sA + stage * L::A_BYTES));
          const uint64_t b_desc = OPMN ? make_umma_desc_mn_sw128(smem_u32(sB + stage * L::B_BYTES), 8192u)
                                       : make_umma_desc_sw128(smem_u32(sB + stage * L::B_BYTES));
          // K-major: advance 16 bf16 = 32 B inside the 128-B swizzle span (+2 in the addr >> 4 field);
          // MN-major: 16 k rows = two 8-row groups of 1 KB (+128)
          constexpr uint32_t kstep = OPMN ? (2048u >> 4) : 2u;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (PAIR)
              umma_bf16_ss_pair(d_tmem, a_desc + kstep * k, b_desc + kstep * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            else
              umma_bf16_ss(d_tmem, a_desc + kstep * k, b_desc + kstep * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (PAIR)
            umma_commit_pair(&empty_bar[stage]);
          else
            umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // accumulator complete -> epilogue(s)
        if (PAIR)
          umma_commit_pair(&tfull_bar[abuf]);
        else
          umma_commit(&tfull_bar[abuf]);
        abuf ^= 1u;
        if (abuf == 0) aphase ^= 1u;
      }
    }
  } else if ((warp == 2 || warp == 3) && lane == 0) {
    // ===================== epilogue panel manager of group g =====================
    // Owns E buffer g: TMA-stores the panel the group has written, then (once the store has read it) refills
    // the buffer with the residual panel of the group's next quarter, or simply hands it back.
    if (k_staged) {
      const int g = warp - 2;
      // panels of this group in processing order: (tile, q, hp) with q = g, g+2 (64-column quarters), hp = 0, 1
      // (32-column halves), keeping only panels whose first column is inside N
      struct QIter {
        int tile, q, hp;
      };
      auto advance = [&](QIter& it) {  // next valid panel, or tile >= num_tiles
        for (;;) {
          if (++it.hp == 2) {
            it.hp = 0;
            it.q += 2;
            if (it.q >= QUARTERS) {
              it.q = g;
              it.tile += n_units;
            }
          }
          if (it.tile >= num_tiles) return;
          const int n_blk = it.tile % p.n_tiles;
          if (n_blk * BLOCK_N + it.q * QUARTER_N + it.hp * PANEL_N < p.N) return;
        }
      };
      auto coords = [&](const QIter& it, int& c0, int& r0) {
        const int m_t = it.tile / p.n_tiles;
        const int n_blk = it.tile - m_t * p.n_tiles;
        c0 = n_blk * BLOCK_N + it.q * QUARTER_N + it.hp * PANEL_N;
        r0 = (PAIR ? 2 * m_t + (int)cta_rank : m_t) * BLOCK_M;
      };
      auto hand_over = [&](const QIter& it, int j) {  // make panel j ready for unit `it`
        uint8_t* ebuf = sE + (g * EB + j) * PANEL_BYTES;
        if (k_has_res) {
          int c0, r0;
          coords(it, c0, r0);
          mbar_arrive_expect_tx(&res_full_bar[g * EB + j], PANEL_BYTES);
          if (RP == 2)
            tma_load_3d(ebuf, &tmR, &res_full_bar[g * EB + j], c0, r0, 0);  // both planes of the stream in one box
          else
            tma_load_2d(ebuf, &tmR, &res_full_bar[g * EB + j], c0, r0);
        } else {
          mbar_arrive(&res_full_bar[g * EB + j]);
        }
      };
      QIter st{unit, g, -1};  // store cursor (advance() moves it onto the first valid panel)
      advance(st);
      QIter ld = st;  // load cursor, runs EB quarters ahead
      for (int j = 0; j < EB && ld.tile < num_tiles; ++j) {
        hand_over(ld, j);
        advance(ld);
      }
      for (uint32_t n = 0; st.tile < num_tiles; ++n) {
        const int j = n % EB;
        uint8_t* ebuf = sE + (g * EB + j) * PANEL_BYTES;
        int c0, r0;
        coords(st, c0, r0);
        mbar_wait(&e_written_bar[g * EB + j], (n / EB) & 1u);  // the group has written quarter n into panel j
        if (RP == 2)
          tma_store_3d(&tmC, ebuf, c0, r0, 0);
        else
          tma_store_2d(&tmC, ebuf, c0, r0);
        tma_store_commit();
        advance(st);
        if (ld.tile < num_tiles) {
          tma_store_wait_read<0>();  // the store has read panel j: refill it for quarter n + EB
          hand_over(ld, j);
          advance(ld);
        }
      }
      tma_store_wait<0>();  // all stores complete before the CTA (and its shared memory) goes away
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue workers =====================
    const int g = (warp - EPI_WARP0) >> 2;  // column group: quarters g, g+2
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int row_in_tile = quarter * 32 + lane;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..255
    const bool has_bias = k_has_bias;
    const bool has_stats = k_has_stats;
    uint32_t abuf = 0, aphase = 0, qseq = 0;  // qseq counts the quarters this group has processed
    long long tk[6] = {0, 0, 0, 0, 0, 0};
    int ntl = 0;
    const bool dbg = p.debug != 0 && blockIdx.x == 0 && (warp == EPI_WARP0 || warp == EPI_WARP0 + 4) && lane == 0;
    // Per-tile scalars (this thread's bias / col_c element for the smem staging, its row's mean / rstd) are fetched one
    // tile ahead, so their global-load latency never sits on the epilogue's critical path.
    float pf_bias = 0.f, pf_colc = 0.f, pf_mean = 0.f, pf_rstd = 1.f;
    auto prefetch_tile = [&](int tile) {
      if (tile >= num_tiles) return;
      const int t_mn = tile % tiles_mn;
      const int m_t = t_mn / p.n_tiles;
      const int n_blk = t_mn - m_t * p.n_tiles;
      const int m_blk = PAIR ? 2 * m_t + (int)cta_rank : m_t;
      const int row = m_blk * BLOCK_M + row_in_tile;
      const int col = n_blk * BLOCK_N + et;
      pf_bias = (has_bias && et < BLOCK_N && col < p.N) ? __ldg(p.bias + col) : 0.f;
      pf_colc = (has_stats && et < BLOCK_N && col < p.N) ? __ldg(p.col_c + col) : 0.f;
      pf_mean = 0.f;
      pf_rstd = 1.f;
      if (has_stats && row < p.M) {
        if (p.part_in != nullptr) {
          // Chan's parallel combination of the per-32-column (mean, M2) partials the producing layer's epilogue left
          const float2* pp = reinterpret_cast<const float2*>(p.part_in) + (int64_t)row * p.npart_in;
          float msum = 0.f, m2 = 0.f;
          for (int i = 0; i < p.npart_in; ++i) msum += __ldg(&pp[i]).x;
          pf_mean = msum / (float)p.npart_in;
          for (int i = 0; i < p.npart_in; ++i) {
            const float2 q = __ldg(&pp[i]);
            const float d = q.x - pf_mean;
            m2 += q.y + 32.f * d * d;
          }
          pf_rstd = rsqrtf(m2 / (32.f * (float)p.npart_in) + p.ln_eps);
        } else {
          const float2 st = __ldg(reinterpret_cast<const float2*>(p.row_stats + 2 * (int64_t)row));
          pf_mean = st.x;
          pf_rstd = st.y;
        }
      }
    };
    prefetch_tile(unit);
    for (int tile = unit; tile < num_tiles; tile += n_units) {
      long long t0 = dbg ? clock64() : 0;
      ++ntl;
      const int split = tile / tiles_mn, t_mn = tile - split * tiles_mn;
      const int m_t = t_mn / p.n_tiles;
      const int n_blk = t_mn - m_t * p.n_tiles;
      const int m_blk = PAIR ? 2 * m_t + (int)cta_rank : m_t;
      const int row = m_blk * BLOCK_M + row_in_tile;
      const bool row_ok = row < p.M;
      const float a_scale = pf_rstd, a_shift = -pf_mean * pf_rstd;
      // ---- stage this tile's bias / col_c once (previous tile's readers are past the first barrier) ----
      named_bar_sync(1, EPI_THREADS);
      if (et < BLOCK_N) {
        sBias[et] = pf_bias;
        sColc[et] = pf_colc;
      }
      prefetch_tile(tile + n_units);
      named_bar_sync(1, EPI_THREADS);
      if (dbg) { long long t = clock64(); tk[0] += t - t0; t0 = t; }

      mbar_wait(&tfull_bar[abuf], aphase);
      tcgen05_fence_after();
      if (dbg) { long long t = clock64(); tk[1] += t - t0; t0 = t; }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + abuf * BLOCK_N;

      if (k_staged) {
        bool released = false;
#pragma unroll 1
        for (int q = g; q < QUARTERS; q += 2) {
#pragma unroll 1
          for (int hp = 0; hp < 2; ++hp) {
            const int c0 = n_blk * BLOCK_N + q * QUARTER_N + hp * PANEL_N;
            if (c0 >= p.N) continue;  // uniform across the CTA
            uint32_t r0[32];
            __syncwarp();
            tmem_ld_32x32b_x32(taddr + q * QUARTER_N + hp * PANEL_N, r0);
            tmem_wait_ld();
            if (dbg) { long long t = clock64(); tk[2] += t - t0; t0 = t; }
            // last TMEM read of this tile by this thread: hand the accumulator back early
            const bool last_q = (q + 2 >= QUARTERS) || (n_blk * BLOCK_N + (q + 2) * QUARTER_N >= p.N);
            const bool last_h = (hp == 1) || (c0 + PANEL_N >= p.N);
            if (last_q && last_h) {
              tcgen05_fence_before();
              (PAIR ? mbar_arrive_leader(&tempty_bar[abuf]) : mbar_arrive(&tempty_bar[abuf]));
              released = true;
            }
            const int pj = qseq % EB;
            mbar_wait(&res_full_bar[g * EB + pj], (qseq / EB) & 1u);  // panel is ours (and holds the residual, if any)
            if (dbg) { long long t = clock64(); tk[3] += t - t0; t0 = t; }
            uint8_t* erow = sE + (g * EB + pj) * PANEL_BYTES + row_in_tile * 64;
            float sh = 0.f, s1 = 0.f, s2 = 0.f;  // shifted one-pass statistics of this row's 32 outputs
            // All shared-memory READS of the panel (residual) and of the staged vectors come first, the four 16-byte
            // result stores last: generic smem pointers may alias as far as the compiler knows, so a store between two
            // chunks would serialise their loads and math (measured: 1.8 k cycles per panel instead of ~0.6 k).
            uint4 rr[4], rl[4];
            uint4* slot[4];
            constexpr int LO = EBUF_BYTES / 16;  // uint4 stride from a hi slot to its lo slot (second plane of the panel)
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
              slot[pc] = reinterpret_cast<uint4*>(erow + ((pc ^ ((row_in_tile >> 1) & 3)) << 4));  // 64-B swizzle
              if (k_has_res) rr[pc] = *slot[pc];
              if (RP == 2) rl[pc] = *(slot[pc] + LO);
            }
            uint4 o[4], ol[4];
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r0[pc * 8 + j]);
              epi_math8(v, sBias, sColc, q * QUARTER_N + hp * PANEL_N + pc * 8, a_scale, a_shift, has_bias, has_stats, k_act);
              if (k_has_res) {
                v[0] += bf16_lo(rr[pc].x);
                v[1] += bf16_hi(rr[pc].x);
                v[2] += bf16_lo(rr[pc].y);
                v[3] += bf16_hi(rr[pc].y);
                v[4] += bf16_lo(rr[pc].z);
                v[5] += bf16_hi(rr[pc].z);
                v[6] += bf16_lo(rr[pc].w);
                v[7] += bf16_hi(rr[pc].w);
              }
              if (RP == 2) {
                v[0] += bf16_lo(rl[pc].x);
                v[1] += bf16_hi(rl[pc].x);
                v[2] += bf16_lo(rl[pc].y);
                v[3] += bf16_hi(rl[pc].y);
                v[4] += bf16_lo(rl[pc].z);
                v[5] += bf16_hi(rl[pc].z);
                v[6] += bf16_lo(rl[pc].w);
                v[7] += bf16_hi(rl[pc].w);
              }
              if (k_part_out) {
                if (pc == 0) sh = v[0];  // shift by the first value: keeps sum((v - sh)^2) - s1^2/n well conditioned
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float d = v[j] - sh;
                  s1 += d;
                  s2 = fmaf(d, d, s2);
                }
              }
              o[pc].x = pack_bf16x2(v[0], v[1]);
              o[pc].y = pack_bf16x2(v[2], v[3]);
              o[pc].z = pack_bf16x2(v[4], v[5]);
              o[pc].w = pack_bf16x2(v[6], v[7]);
              if (RP == 2) {  // lo = bf16(v - hi): the part of the fp32 value the hi plane cannot hold
                ol[pc].x = pack_bf16x2(v[0] - bf16_lo(o[pc].x), v[1] - bf16_hi(o[pc].x));
                ol[pc].y = pack_bf16x2(v[2] - bf16_lo(o[pc].y), v[3] - bf16_hi(o[pc].y));
                ol[pc].z = pack_bf16x2(v[4] - bf16_lo(o[pc].z), v[5] - bf16_hi(o[pc].z));
                ol[pc].w = pack_bf16x2(v[6] - bf16_lo(o[pc].w), v[7] - bf16_hi(o[pc].w));
              }
            }
            if (dbg) { long long t = clock64(); tk[5] += t - t0; t0 = t; }
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
              *slot[pc] = o[pc];
              if (RP == 2) *(slot[pc] + LO) = ol[pc];
            }
            if (k_part_out && row_ok) {
              const float mq = s1 * (1.f / 32.f);
              reinterpret_cast<float2*>(p.part_out)[(int64_t)row * (p.N >> 5) + (c0 >> 5)] =
                  make_float2(sh + mq, fmaxf(s2 - s1 * mq, 0.f));
            }
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store
            mbar_arrive(&e_written_bar[g * EB + pj]);
            ++qseq;
            if (dbg) { long long t = clock64(); tk[4] += t - t0; t0 = t; }
          }
        }
        if (!released) {
          tcgen05_fence_before();
          (PAIR ? mbar_arrive_leader(&tempty_bar[abuf]) : mbar_arrive(&tempty_bar[abuf]));
        }
        if ((EPI == EPI_RES || EPI == EPI_RES2 || EPI == EPI_GENERIC) && k_part_out && p.stats_out != nullptr) {
          // ---- fused LayerNorm statistics: last column tile of this 128-row block? ----
          // The 256 epilogue threads wrote the tile's partials with plain stores.  One barrier orders them before thread 0,
          // whose single gpu-scope fence then publishes them cumulatively ahead of the counter increment (a fence per
          // thread measured +0.5 ms per step: 256 membar.gl per tile).
          named_bar_sync(2, EPI_THREADS);
          if (et == 0) {
            __threadfence();
            const unsigned old = atomicAdd(&p.row_counters[m_blk], 1u);
            const bool last = old == (unsigned)(p.n_tiles - 1);
            if (last) p.row_counters[m_blk] = 0u;  // nobody else touches it any more in this launch
            *s_last_tile = last ? 1u : 0u;
          }
          named_bar_sync(2, EPI_THREADS);
          if (*s_last_tile != 0u && et < BLOCK_M) {
            __threadfence();
            const int r = m_blk * BLOCK_M + et;
            if (r < p.M) {
              // Chan's parallel combination of the (mean, M2) partials of the row's 32-column blocks
              const int np = p.N >> 5;
              const float2* pp = reinterpret_cast<const float2*>(p.part_out) + (int64_t)r * np;
              float msum = 0.f;
              for (int i = 0; i < np; ++i) msum += __ldcg(&pp[i]).x;
              const float mean = msum / (float)np;
              float m2 = 0.f;
              for (int i = 0; i < np; ++i) {
                const float2 q = __ldcg(&pp[i]);
                const float d = q.x - mean;
                m2 += q.y + 32.f * d * d;
              }
              reinterpret_cast<float2*>(p.stats_out)[r] = make_float2(mean, rsqrtf(m2 / (32.f * (float)np) + p.ln_eps));
            }
          }
        }
      } else if (EPI == EPI_GENERIC) {
        // ---- direct path (fp32 output): registers -> global ----
#pragma unroll 1
        for (int c = g; c < BLOCK_N / 32; c += 2) {
          uint32_t r[32];
          __syncwarp();
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_wait_ld();
          const int col0 = n_blk * BLOCK_N + c * 32;
          if (col0 >= p.N) continue;  // warp-uniform
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const int col = col0 + gq * 8;
            if (col >= p.N) break;  // warp-uniform (N % 8 == 0)
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[gq * 8 + j]);
            epi_math8(v, sBias, sColc, c * 32 + gq * 8, a_scale, a_shift, has_bias, has_stats, p.act);
            if (row_ok) {
              if (p.residual != nullptr) {
                const uint4 rr = ld_nc_v4(p.residual + (int64_t)row * p.ldr + col);
                v[0] += bf16_lo(rr.x);
                v[1] += bf16_hi(rr.x);
                v[2] += bf16_lo(rr.y);
                v[3] += bf16_hi(rr.y);
                v[4] += bf16_lo(rr.z);
                v[5] += bf16_hi(rr.z);
                v[6] += bf16_lo(rr.w);
                v[7] += bf16_hi(rr.w);
              }
              if (p.out_fp32) {
                float* out = reinterpret_cast<float*>(p.C) + (int64_t)split * p.split_stride + (int64_t)row * p.ldc + col;
                *reinterpret_cast<float4*>(out) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(out + 4) = make_float4(v[4], v[5], v[6], v[7]);
              } else {
                __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + col;
                uint4 o;
                o.x = pack_bf16x2(v[0], v[1]);
                o.y = pack_bf16x2(v[2], v[3]);
                o.z = pack_bf16x2(v[4], v[5]);
                o.w = pack_bf16x2(v[6], v[7]);
                st_v4(out, o);
              }
            }
          }
        }
        tcgen05_fence_before();
        (PAIR ? mbar_arrive_leader(&tempty_bar[abuf]) : mbar_arrive(&tempty_bar[abuf]));
      }
      abuf ^= 1u;
      if (abuf == 0) aphase ^= 1u;
    }
    if (dbg && ntl > 0)
      printf("gemm dbg warp %d tiles %d cycles/tile: stats+stage %lld wait_acc %lld ldtm %lld wait_panel %lld math %lld "
             "store+fence+arrive %lld\n", warp, ntl, tk[0] / ntl, tk[1] / ntl, tk[2] / ntl, tk[3] / ntl, tk[5] / ntl, tk[4] / ntl);
  }

  tcgen05_fence_before();
  if (PAIR)
    cluster_sync_all();  // the peer may still signal this CTA's barriers / read its smem until both are done
  else
    __syncthreads();
  if (warp == 2) {
    __syncwarp();
    tcgen05_fence_after();
    if (PAIR)
      tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    else
      tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES, int EB, bool PAIR, int EPI, bool OPMN = false>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                GemmParams& p, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, STAGES, EB, PAIR, EPI == EPI_RES2 ? 2 : 1>;
  static bool attr_set = false;  // benign race: setting the attribute twice is harmless
  if (!attr_set) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BLOCK_N, STAGES, EB, PAIR, EPI, OPMN>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
    attr_set = true;
  }
  p.n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  if (!PAIR) {
    const int tiles = p.m_tiles * p.n_tiles * p.ksplit;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    return report_cuda(launch_pdl(gemm_bf16_tn_kernel<BLOCK_N, STAGES, EB, PAIR, EPI, OPMN>, dim3(grid), dim3(GEMM_THREADS),
                                  L::DYN_BYTES, stream, 1, tmA, tmB, tmC, tmR, p),
                       "gemm_bf16_tn_kernel launch");
  }
  // CTA pairs: cluster of 2 along x, one pair per two SMs
  const int tiles = ((p.m_tiles + 1) / 2) * p.n_tiles * p.ksplit;
  const int pairs = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
  return report_cuda(launch_pdl(gemm_bf16_tn_kernel<BLOCK_N, STAGES, EB, PAIR, EPI, OPMN>, dim3(2 * pairs), dim3(GEMM_THREADS),
                                L::DYN_BYTES, stream, 2, tmA, tmB, tmC, tmR, p),
                     "gemm_bf16_tn_kernel<pair> launch");
}

}  // namespace

void count_launch(int n);

}  // namespace vlmclip

using namespace vlmclip;

static int gemm_bf16_impl(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, const float* bias,
                          const void* residual, int64_t ldr, const float* row_stats, const float* col_c,
                          const float* stats_part_in, int npart_in, float ln_eps, float* stats_part_out, float* stats_out,
                          int32_t* row_counters, int M, int N, int K, int act, int out_fp32, int splitk_planes,
                          int64_t splitk_stride, void* stream) {
  VLMCLIP_CHECK_ARG(A && W && C, "gemm: null A/W/C pointer");
  VLMCLIP_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: non-positive dims M=%d N=%d K=%d", M, N, K);
  VLMCLIP_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "gemm: K and N must be multiples of 8 (K=%d N=%d)", K, N);
  VLMCLIP_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, "gemm: lda/ldw/ldc must be multiples of 8");
  VLMCLIP_CHECK_ARG(lda >= K && ldw >= K && ldc >= N, "gemm: leading dimension smaller than row length");
  VLMCLIP_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)C % 16 == 0),
                    "gemm: A/W/C must be 16-byte aligned");
  VLMCLIP_CHECK_ARG(act >= 0 && act <= 3, "gemm: unknown activation %d", act);
  VLMCLIP_CHECK_ARG(!(row_stats && stats_part_in), "gemm: give row_stats or stats_part_in, not both");
  VLMCLIP_CHECK_ARG(((row_stats != nullptr) || (stats_part_in != nullptr)) == (col_c != nullptr),
                    "gemm: LN fold needs col_c together with row_stats / stats_part_in");
  VLMCLIP_CHECK_ARG(stats_part_in == nullptr || (npart_in > 0 && npart_in * 32 == K),
                    "gemm: stats_part_in must hold K/32 = %d partials per row (got %d)", K / 32, npart_in);
  VLMCLIP_CHECK_ARG(stats_part_out == nullptr || (N % 32 == 0 && !out_fp32),
                    "gemm: stats_part_out needs N %% 32 == 0 and bf16 output");
  VLMCLIP_CHECK_ARG((stats_out == nullptr) == (row_counters == nullptr) && (stats_out == nullptr || stats_part_out != nullptr),
                    "gemm: stats_out needs row_counters and stats_part_out");
  if (residual) {
    VLMCLIP_CHECK_ARG(ldr % 8 == 0 && ldr >= N && (uintptr_t)residual % 16 == 0,
                      "gemm: residual must be 16-byte aligned with ldr %% 8 == 0");
  }
  if (bias) VLMCLIP_CHECK_ARG((uintptr_t)bias % 16 == 0, "gemm: bias must be 16-byte aligned");
  if (col_c) VLMCLIP_CHECK_ARG((uintptr_t)col_c % 16 == 0, "gemm: col_c must be 16-byte aligned");

  GemmParams p;
  p.C = C;
  p.bias = bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.row_stats = row_stats;
  p.col_c = col_c;
  p.part_in = stats_part_in;
  p.part_out = stats_part_out;
  p.stats_out = stats_out;
  p.row_counters = reinterpret_cast<unsigned*>(row_counters);
  p.npart_in = npart_in;
  p.ln_eps = ln_eps;
  p.ldc = ldc;
  p.ldr = ldr;
  p.M = M;
  p.N = N;
  p.K = K;
  p.act = act;
  p.out_fp32 = out_fp32;
  p.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  p.k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  p.ksplit = 1;
  p.kb_per_split = p.k_blocks;
  p.split_stride = 0;
  if (splitk_planes > 1) {
    VLMCLIP_CHECK_ARG(out_fp32 && !bias && !residual && !row_stats && !stats_part_in && !stats_part_out && act == 0,
                      "gemm: a split reduction needs the plain fp32 output (no bias / residual / LN fold / activation)");
    p.kb_per_split = (p.k_blocks + splitk_planes - 1) / splitk_planes;
    p.ksplit = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
    p.split_stride = splitk_stride;
  }

  // Tile width: 256 columns feed the tensor core best (96 B of smem operand traffic per clock against 128 B/clk for a
  // 128-wide tile), but when 256-wide tiles leave the last wave of the persistent grid mostly empty (e.g. N = 512,
  // M = 19712: 308 tiles on 148 SMs = 3 waves for 2.08 waves of work) the narrower tile wins.  Estimated cost =
  // waves x relative tile time (a 128-wide tile costs ~0.56 of a 256-wide one).
  bool wide = N > 128;
  if (wide) {
    const int sms = sm_count();
    const long t256 = (long)p.m_tiles * ((N + 255) / 256), t128 = (long)p.m_tiles * ((N + 127) / 128);
    const double c256 = (double)((t256 + sms - 1) / sms), c128 = 0.56 * (double)((t128 + sms - 1) / sms);
    if (c128 < 0.92 * c256) wide = false;
    const char* force = getenv("VLMCLIP_GEMM_NARROW");  // experiment switch
    if (force != nullptr && force[0] == '1') wide = false;
  }
  const int block_n = wide ? 256 : 128;
  p.staged = out_fp32 ? 0 : 1;
  CUtensorMap tmA, tmB, tmC, tmR;
  int rc = make_tmap_bf16(&tmA, A, M, K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, W, N, K, ldw, block_n);
  if (rc) return rc;
  tmC = tmA;
  tmR = tmA;
  if (p.staged) {
    rc = make_tmap_bf16_box(&tmC, C, M, N, ldc, BLOCK_M, PANEL_N);
    if (rc) return rc;
    if (residual) {
      rc = make_tmap_bf16_box(&tmR, residual, M, N, ldr, BLOCK_M, PANEL_N);
      if (rc) return rc;
    }
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  count_launch(1);
  // Two panels per epilogue group (at the price of one main-loop stage) pay off when the main loop of a tile is short
  // (K <= 1024: ~6k tensor-core cycles) and the epilogue has a load -> process -> store chain per panel (residual).
  // VLMCLIP_GEMM_CFG = "41" | "32" overrides for experiments; VLMCLIP_GEMM_INTPACK=1 packs bf16 on the integer pipe.
  static const int cfg_override = []() {
    const char* e = getenv("VLMCLIP_GEMM_CFG");
    return e == nullptr ? 0 : atoi(e);
  }();
  static const int intpack_override = []() {
    const char* e = getenv("VLMCLIP_GEMM_INTPACK");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  p.int_pack = intpack_override;
  static const int dbg_flag = []() {
    const char* e = getenv("VLMCLIP_GEMM_DEBUG");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  p.debug = dbg_flag;
  // CTA pairs for the wide tiles (VLMCLIP_GEMM_CFG=1 forces the single-CTA kernel for comparison)
  const bool pair = wide && cfg_override != 1 && M > BLOCK_M;
  // epilogue specialisation (see the EPI_* enum); anything else takes the generic instantiation
  const bool fold = (row_stats != nullptr || stats_part_in != nullptr) && bias != nullptr;
  int epi = EPI_GENERIC;
  if (p.staged && cfg_override != 2) {
    if (fold && residual == nullptr && stats_part_out == nullptr && act == 0) epi = EPI_FOLD;
    if (fold && residual == nullptr && stats_part_out == nullptr && act == 1) epi = EPI_FOLD_ACT;
    if (!fold && row_stats == nullptr && stats_part_in == nullptr && bias != nullptr && residual != nullptr && act == 0)
      epi = EPI_RES;
    if (!fold && row_stats == nullptr && stats_part_in == nullptr && bias == nullptr && residual == nullptr &&
        stats_part_out == nullptr && act == 0)
      epi = EPI_PLAIN;
  }
#define VLMCLIP_GEMM_LAUNCH(BN, ST, EBN, PR)                                                          \
  switch (epi) {                                                                                      \
    case EPI_FOLD: return launch_gemm<BN, ST, EBN, PR, EPI_FOLD>(tmA, tmB, tmC, tmR, p, s);          \
    case EPI_FOLD_ACT: return launch_gemm<BN, ST, EBN, PR, EPI_FOLD_ACT>(tmA, tmB, tmC, tmR, p, s);  \
    case EPI_RES: return launch_gemm<BN, ST, EBN, PR, EPI_RES>(tmA, tmB, tmC, tmR, p, s);            \
    case EPI_PLAIN: return launch_gemm<BN, ST, EBN, PR, EPI_PLAIN>(tmA, tmB, tmC, tmR, p, s);        \
    default: return launch_gemm<BN, ST, EBN, PR, EPI_GENERIC>(tmA, tmB, tmC, tmR, p, s);             \
  }
  if (pair) {
    rc = make_tmap_bf16(&tmB, W, N, K, ldw, 128);  // each CTA of the pair loads 128 of the 256 W rows of a tile
    if (rc) return rc;
    // short-K residual GEMMs (out-proj) wait on their residual panels (TMA store -> read-done -> TMA load chain): four
    // panels per epilogue group instead of two, paid for with one pipeline stage (measured 65.8 -> 62.6 us; with K = 3072
    // the deeper pipeline wins, 165.6 vs 168.2 us)
    if (epi == EPI_RES && K <= 1024 && cfg_override != 62) return launch_gemm<256, 5, 4, true, EPI_RES>(tmA, tmB, tmC, tmR, p, s);
    VLMCLIP_GEMM_LAUNCH(256, 6, 2, true)
  }
  if (wide) {
    VLMCLIP_GEMM_LAUNCH(256, 4, 2, false)
  }
  VLMCLIP_GEMM_LAUNCH(128, 6, 2, false)
#undef VLMCLIP_GEMM_LAUNCH
}

extern "C" int vlmclip_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc,
                                 const float* bias, const void* residual, int64_t ldr, const float* row_stats,
                                 const float* col_c, const float* stats_part_in, int npart_in, float ln_eps,
                                 float* stats_part_out, float* stats_out, int32_t* row_counters, int M, int N, int K, int act,
                                 int out_fp32, void* stream) {
  return gemm_bf16_impl(A, lda, W, ldw, C, ldc, bias, residual, ldr, row_stats, col_c, stats_part_in, npart_in, ln_eps,
                        stats_part_out, stats_out, row_counters, M, N, K, act, out_fp32, 1, 0, stream);
}

// C_s[M,N] (fp32) = A[M, K_s] W[N, K_s]^T for `planes` consecutive slices K_s of the reduction dimension, plane s at
// C + s * plane_stride floats: the weight-gradient GEMMs of the full-fine-tune backward (dW = dY^T X, K = all tokens)
// have only a handful of output tiles, so the reduction is what gets distributed over the SMs.  The caller adds the
// planes in index order (vlmclip_sum_planes_f32): deterministic.  Returns the number of planes actually written
// (<= planes: no slice is empty), or a negative error code.
extern "C" int vlmclip_gemm_bf16_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc,
                                        int64_t plane_stride, int planes, int M, int N, int K, void* stream) {
  VLMCLIP_CHECK_ARG(planes >= 1 && plane_stride >= (int64_t)(M - 1) * ldc + N, "gemm_splitk: bad planes / plane_stride");
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  const int per = (k_blocks + planes - 1) / planes;
  const int used = (k_blocks + per - 1) / per;
  const int rc = gemm_bf16_impl(A, lda, W, ldw, C, ldc, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, 0.f, nullptr, nullptr,
                                nullptr, M, N, K, 0, 1, planes, plane_stride, stream);
  return rc != 0 ? (rc > 0 ? -rc : rc) : used;
}

// x (two planes: hi at X, lo at X + plane_stride) += A W^T + bias, in place; optional LayerNorm partials of the updated
// rows.  The out-proj / fc2 step of an encoder layer (HF modeling_clip.py:372-383) on the two-term residual stream.
extern "C" int vlmclip_gemm_bf16_res2(const void* A, int64_t lda, const void* W, int64_t ldw, void* X, int64_t ldx,
                                      int64_t plane_stride, const float* bias, float* stats_part_out, float* stats_out,
                                      int32_t* row_counters, float ln_eps, int M, int N, int K, void* stream) {
  VLMCLIP_CHECK_ARG(A && W && X && bias, "gemm_res2: null A/W/X/bias pointer");
  VLMCLIP_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_res2: non-positive dims M=%d N=%d K=%d", M, N, K);
  VLMCLIP_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "gemm_res2: K and N must be multiples of 8 (K=%d N=%d)", K, N);
  VLMCLIP_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && ldx % 8 == 0 && plane_stride % 8 == 0,
                    "gemm_res2: lda/ldw/ldx/plane_stride must be multiples of 8");
  VLMCLIP_CHECK_ARG(lda >= K && ldw >= K && ldx >= N, "gemm_res2: leading dimension smaller than row length");
  VLMCLIP_CHECK_ARG(plane_stride >= (int64_t)(M - 1) * ldx + N, "gemm_res2: the two planes overlap");
  VLMCLIP_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)X % 16 == 0) &&
                        ((uintptr_t)bias % 16 == 0),
                    "gemm_res2: A/W/X/bias must be 16-byte aligned");
  VLMCLIP_CHECK_ARG(stats_part_out == nullptr || N % 32 == 0, "gemm_res2: stats_part_out needs N %% 32 == 0");
  VLMCLIP_CHECK_ARG((stats_out == nullptr) == (row_counters == nullptr) && (stats_out == nullptr || stats_part_out != nullptr),
                    "gemm_res2: stats_out needs row_counters and stats_part_out");

  GemmParams p{};
  p.C = X;
  p.bias = bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(X);
  p.part_out = stats_part_out;
  p.stats_out = stats_out;
  p.row_counters = reinterpret_cast<unsigned*>(row_counters);
  p.ln_eps = ln_eps;
  p.ldc = ldx;
  p.ldr = ldx;
  p.M = M;
  p.N = N;
  p.K = K;
  p.staged = 1;
  p.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  p.k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  p.ksplit = 1;
  p.kb_per_split = p.k_blocks;
  p.split_stride = 0;
  static const int dbg_flag = []() {
    const char* e = getenv("VLMCLIP_GEMM_DEBUG");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  p.debug = dbg_flag;
  // stages / panels-per-group of the pair kernel: 52 (default), 43, 61 (VLMCLIP_GEMM_RES2_CFG, for measurements)
  static const int cfg = []() {
    const char* e = getenv("VLMCLIP_GEMM_RES2_CFG");
    return e == nullptr ? 0 : atoi(e);
  }();

  bool wide = N > 128;
  if (wide) {
    const int sms = sm_count();
    const long t256 = (long)p.m_tiles * ((N + 255) / 256), t128 = (long)p.m_tiles * ((N + 127) / 128);
    const double c256 = (double)((t256 + sms - 1) / sms), c128 = 0.56 * (double)((t128 + sms - 1) / sms);
    if (c128 < 0.92 * c256) wide = false;
  }
  const bool pair = wide && M > BLOCK_M;
  CUtensorMap tmA, tmB, tmX;
  int rc = make_tmap_bf16(&tmA, A, M, K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, W, N, K, ldw, pair ? 128 : (wide ? 256 : 128));
  if (rc) return rc;
  rc = make_tmap_bf16_planes(&tmX, X, M, N, ldx, plane_stride, 2, BLOCK_M, PANEL_N);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  count_launch(1);
  if (pair) {
    if (cfg == 43) return launch_gemm<256, 4, 3, true, EPI_RES2>(tmA, tmB, tmX, tmX, p, s);
    // long K (fc2: 48-64 k-blocks per tile): the main loop hides one panel chain per group, the sixth stage pays more
    // (B/16 fc2 172.0 vs 176.2 us, L/14 fc2 814 vs 828 us); short K (out-proj) needs the second panel (84 vs 104 us)
    if (cfg == 61 || (cfg == 0 && K >= 2048)) return launch_gemm<256, 6, 1, true, EPI_RES2>(tmA, tmB, tmX, tmX, p, s);
    return launch_gemm<256, 5, 2, true, EPI_RES2>(tmA, tmB, tmX, tmX, p, s);
  }
  if (wide) return launch_gemm<256, 4, 1, false, EPI_RES2>(tmA, tmB, tmX, tmX, p, s);
  return launch_gemm<128, 4, 2, false, EPI_RES2>(tmA, tmB, tmX, tmX, p, s);
}

// C_s[M,N] (fp32) = A_t[K_s, M]^T B_t[K_s, N] over `planes` slices K_s of the K rows, plane s at C + s * plane_stride: the
// weight gradient dW = dY^T X read from dY [tokens, N_out] and X [tokens, K_in] as they lie (MN-major operands, no
// transposed copies; K needs no padding: rows past K are zero-filled by TMA).  Returns the planes written or < 0.
extern "C" int vlmclip_gemm_bf16_atb_splitk(const void* At, int64_t ldat, const void* Bt, int64_t ldbt, float* C, int64_t ldc,
                                            int64_t plane_stride, int planes, int M, int N, int K, void* stream) {
  VLMCLIP_CHECK_ARG(At && Bt && C, "gemm_atb: null pointer");
  VLMCLIP_CHECK_ARG(M > 0 && N > 0 && K > 0 && M % 8 == 0 && N % 8 == 0, "gemm_atb: M and N must be positive multiples of 8");
  VLMCLIP_CHECK_ARG(ldat % 8 == 0 && ldbt % 8 == 0 && ldat >= M && ldbt >= N && ldc >= N && ldc % 4 == 0,
                    "gemm_atb: bad leading dimensions");
  VLMCLIP_CHECK_ARG(((uintptr_t)At % 16 == 0) && ((uintptr_t)Bt % 16 == 0) && ((uintptr_t)C % 16 == 0),
                    "gemm_atb: pointers must be 16-byte aligned");
  VLMCLIP_CHECK_ARG(planes >= 1 && plane_stride >= (int64_t)(M - 1) * ldc + N, "gemm_atb: bad planes / plane_stride");
  GemmParams p{};
  p.C = C;
  p.ldc = ldc;
  p.M = M;
  p.N = N;
  p.K = K;
  p.out_fp32 = 1;
  p.staged = 0;
  p.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  p.k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  p.kb_per_split = (p.k_blocks + planes - 1) / planes;
  p.ksplit = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  p.split_stride = plane_stride;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16(&tmA, At, K, M, ldat, 64);  // box: 64 k rows x 64 contiguous m
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, Bt, K, N, ldbt, 64);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  count_launch(1);
  const bool wide = N > 128;
  if (wide && M > BLOCK_M)
    rc = launch_gemm<256, 6, 2, true, EPI_GENERIC, true>(tmA, tmB, tmA, tmA, p, s);
  else if (wide)
    rc = launch_gemm<256, 4, 2, false, EPI_GENERIC, true>(tmA, tmB, tmA, tmA, p, s);
  else
    rc = launch_gemm<128, 6, 2, false, EPI_GENERIC, true>(tmA, tmB, tmA, tmA, p, s);
  return rc != 0 ? (rc > 0 ? -rc : rc) : p.ksplit;
}

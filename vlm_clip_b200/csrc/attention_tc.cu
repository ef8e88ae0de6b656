// tcgen05 attention for the CLIP towers: out = softmax(q k^T * scale + mask) v, head_dim 64, S <= 224.
//   HF modeling_clip.py:261-279 (eager_attention_forward), :318-331 (dispatch), :546-551 (causal + padding mask)
//
// One persistent CTA per SM walks work items (batch, head, 128-query tile).  Per item:
//   TMA        Q [128 x 64], K [Npad x 64], V [Npad x 64] (rows of the fused qkv activation, 128-B swizzle),
//              3-4 smem stages so loads run items ahead of the tensor core
//   tcgen05    S = Q K^T            (SS MMA, M=128, N=Npad, 4 k-steps)   -> TMEM S buffer (it & 1)
//   softmax    256 threads = 2 threads per query row (each owns half of the keys, so the whole row lives in
//              registers and TMEM is read ONCE: tcgen05.ld moves only 64 B/clk/SM, as scarce as the 16 exp2/clk/SM
//              of the SFUs).  Row max exchanged through smem, p = exp2(s*c - m*c), truncated to bf16 with integer ops
//              (F2FP shares the SFU pipe with MUFU.EX2), row sum taken over the TRUNCATED values so the normalisation
//              is consistent with the P the tensor core sees; P is written with tcgen05.st over the S columns.
//   tcgen05    O = P V              (TS MMA: A = P from TMEM, B = V as an MN-major smem operand)  -> TMEM O buffer
//   epilogue   (one item late, so P.V has a whole softmax of time) tcgen05.ld O, * 1/rowsum, bf16, 64 B per thread
// TMEM map (512 columns): S/P buffer 0 [0, NB), S/P buffer 1 [NB, 2 NB), O [448, 512);  NB = round_up(Npad, 32) <= 224.
// tcgen05.mma executes in issue order, so S(i+2) may be issued right behind P.V(i) on the same buffer.
#include <cstdio>
#include <cstdlib>

#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);
int attention_fwd_mma_sync(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H, int causal,
                           float scale, cudaStream_t stream);

namespace {

constexpr int AT_THREADS = 384;  // warpgroup 0: warp0 TMA, warp1 MMA, warp2 TMEM alloc; warpgroups 1-2: softmax/epilogue
constexpr int AT_SM_WARP0 = 4;
constexpr int AT_SM_THREADS = 256;
constexpr int AT_M = 128;
constexpr int AT_HD = 64;
constexpr uint32_t AT_Q_BYTES = AT_M * AT_HD * 2;  // 16 KB
constexpr int AT_TMEM_COLS = 512;
constexpr int AT_O_COL = 448;
constexpr int AT_MAX_STAGES = 4;
constexpr int AT_MAX_NPAD = 224;

struct AttnParams {
  const uint8_t* key_mask;
  __nv_bfloat16* out;
  int B, S, H, D;
  int causal;
  float scale_log2e;
  int Npad;     // keys rounded up to a multiple of 16 (<= 224)
  int nb;       // TMEM columns per S buffer: Npad rounded up to 32
  int mtiles;   // ceil(S / 128)
  int num_items;
  uint32_t kv_bytes;     // Npad * 128 (bytes TMA transfers per K or V tile)
  uint32_t kv_stride;    // kv_bytes rounded up to 1024 (placement)
  uint32_t stage_bytes;  // Q + K + V
  int nstage;            // smem stages (2..4)
  uint32_t out_stage_off;  // byte offset of the 16 KB output staging tile (1024-aligned)
  int debug;             // VLMCLIP_ATTN_DEBUG=1: one warp prints its per-phase cycle totals (development aid)
};

// 16-bit visibility mask of keys [k0, k0+16) for query row qrow
template <bool GENERAL_MASK>
__device__ __forceinline__ uint32_t chunk_mask(const AttnParams& p, const uint8_t* km, int k0, int qrow) {
  uint32_t bits = 0u;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
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

__device__ __forceinline__ void at_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// HALF16 = 16-column chunks owned by one thread (ceil(Npad / 32)); the row's scores stay in registers.
template <int HALF16, bool GENERAL_MASK>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage0 = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.nstage * p.stage_bytes);
  uint64_t* kv_full = bars;                     // [4] TMA landed Q,K,V
  uint64_t* kv_empty = bars + AT_MAX_STAGES;    // [4] P.V finished reading the stage
  uint64_t* s_full = bars + 2 * AT_MAX_STAGES;  // [2] S = Q K^T complete
  uint64_t* p_full = s_full + 2;                // [2] softmax wrote P (256 arrivals)
  uint64_t* o_full = s_full + 4;                // [1] O = P V complete
  uint64_t* o_free = s_full + 5;                // [1] epilogue drained O (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);
  float* s_max = reinterpret_cast<float*>(s_full + 8);  // [2][128] partial row maxima
  float* s_sum = s_max + 256;                           // [2][128] partial row sums
  uint8_t* s_out = smem + p.out_stage_off;              // [128 rows][128 B] bf16 O tile, 16-B chunks XOR-swizzled by row

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 0 && lane == 0) {
    for (int b = 0; b < AT_MAX_STAGES; ++b) {
      mbar_init(&kv_full[b], 1);
      mbar_init(&kv_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], AT_SM_THREADS);
    }
    mbar_init(o_full, 1);
    mbar_init(o_free, AT_SM_THREADS);
    fence_mbar_init();
  }
  if (warp == 2) {
    __syncwarp();
    tmem_alloc<AT_TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = (p.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items of this CTA

  // Registers are re-allocated between warpgroups (register allocation is per 128 threads): the control warpgroup
  // needs almost nothing, the two softmax warpgroups keep half a score row per thread in registers.
  // 128 x 56 + 256 x 224 = 64512 = 384 x 168.  The setmaxnreg sits INSIDE each role branch so ptxas sees which limit
  // governs which code.
  if (warp < AT_SM_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    for (int it = 0; it < n_items; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int sg = it % p.nstage;
      const uint32_t ph = (it / p.nstage) & 1;
      const int mt = item % p.mtiles;
      const int bh = item / p.mtiles;
      const int bb = bh / p.H, h = bh - bb * p.H;
      uint8_t* st = stage0 + sg * p.stage_bytes;
      mbar_wait(&kv_empty[sg], ph ^ 1u);
      mbar_arrive_expect_tx(&kv_full[sg], AT_Q_BYTES + 2 * p.kv_bytes);
      tma_load_2d(st, &tmQ, &kv_full[sg], h * AT_HD, bb * p.S + mt * AT_M);
      tma_load_2d(st + AT_Q_BYTES, &tmKV, &kv_full[sg], p.D + h * AT_HD, bb * p.S);
      tma_load_2d(st + AT_Q_BYTES + p.kv_stride, &tmKV, &kv_full[sg], 2 * p.D + h * AT_HD, bb * p.S);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = make_idesc_bf16(AT_M, p.Npad);
    const uint32_t idesc_o = make_idesc_bf16_b_mn(AT_M, AT_HD);
    const int ksteps = p.Npad >> 4;
    auto issue_s = [&](int it) {
      const int b = it & 1;
      const int sg = it % p.nstage;
      uint8_t* st = stage0 + sg * p.stage_bytes;
      mbar_wait(&kv_full[sg], (it / p.nstage) & 1);
      tcgen05_fence_after();
      const uint64_t qd = make_umma_desc_sw128(smem_u32(st));
      const uint64_t kd = make_umma_desc_sw128(smem_u32(st + AT_Q_BYTES));
#pragma unroll
      for (int k = 0; k < AT_HD / 16; ++k)
        umma_bf16_ss(tmem_base + b * p.nb, qd + 2u * k, kd + 2u * k, idesc_s, k != 0 ? 1u : 0u);
      umma_commit(&s_full[b]);
    };
    if (n_items > 0) issue_s(0);
    if (n_items > 1) issue_s(1);
    for (int it = 0; it < n_items; ++it) {
      const int b = it & 1;
      const int sg = it % p.nstage;
      uint8_t* st = stage0 + sg * p.stage_bytes;
      mbar_wait(&p_full[b], (it >> 1) & 1);  // P(it) is in TMEM (and every S(it) read has retired)
      mbar_wait(o_free, (it & 1) ^ 1u);      // the epilogue has drained O(it-1)
      tcgen05_fence_after();
      const uint64_t vd = make_umma_desc_mn_sw128(smem_u32(st + AT_Q_BYTES + p.kv_stride), p.kv_stride);
      for (int k = 0; k < ksteps; ++k)  // 16 keys per step: 8 packed TMEM columns of P, 2048 B of V
        umma_bf16_ts(tmem_base + AT_O_COL, tmem_base + b * p.nb + k * 8, vd + static_cast<uint64_t>(k) * (2048u >> 4),
                     idesc_o, k != 0 ? 1u : 0u);
      umma_commit(o_full);
      umma_commit(&kv_empty[sg]);
      if (it + 2 < n_items) issue_s(it + 2);  // in-order behind P.V(it): safe to overwrite S/P buffer b
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== softmax / epilogue: 2 threads per query row =====================
    const int hf = (warp - AT_SM_WARP0) >> 2;  // which half of the keys / of the output columns
    const int wq = warp & 3;         // TMEM lane quarter
    const int r_local = wq * 32 + lane;
    const float c = p.scale_log2e;
    const int total16 = p.Npad >> 4;
    const int first16 = (total16 + 1) >> 1;             // chunks of half 0
    const int my_c0 = hf == 0 ? 0 : first16;            // first 16-column chunk of this thread
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;

    long long te[3] = {0, 0, 0};
    const bool dbg = p.debug != 0 && blockIdx.x == 0 && (warp == AT_SM_WARP0 || warp == AT_SM_WARP0 + 4) && lane == 0;
    // Deferred epilogue of item `it`: O / rowsum -> bf16 -> smem tile -> row-contiguous global stores (a thread owns half a
    // row in TMEM; writing it straight out would be 32 scattered 16-B segments per store instruction).
    auto epilogue = [&](int it, float l_total) {
      long long e0 = dbg ? clock64() : 0;
      const int item = blockIdx.x + it * gridDim.x;
      const int mt = item % p.mtiles;
      const int bh = item / p.mtiles;
      const int bb = bh / p.H, h = bh - bb * p.H;
      const bool warp_valid = (mt * AT_M + wq * 32) < p.S;
      mbar_wait(o_full, it & 1);
      tcgen05_fence_after();
      if (dbg) { long long t = clock64(); te[0] += t - e0; e0 = t; }
      if (warp_valid) {
        uint32_t o[32];
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_base + lane_off + AT_O_COL + hf * 32, o);
        tmem_wait_ld();
        if (dbg) { long long t = clock64(); te[1] += t - e0; e0 = t; }
        const float inv = l_total > 0.f ? __fdividef(1.f, l_total) : 0.f;
        uint8_t* srow = s_out + r_local * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(srow + (((hf * 4 + q) ^ (r_local & 7)) << 4)) = v;
        }
      }
      tcgen05_fence_before();
      mbar_arrive(o_free);  // TMEM O is drained: P.V of the next item may start
      at_bar_sync();        // the whole O tile is in smem
      {
        // 8 warps x 16 rows; one store instruction = 4 rows x 128 contiguous bytes (8 lanes per row)
        const int w8 = warp - AT_SM_WARP0;
        const int chunk = lane & 7;
        __nv_bfloat16* obase = p.out + ((int64_t)bb * p.S) * p.D + h * AT_HD + chunk * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = w8 * 16 + k * 4 + (lane >> 3);
          const int qr = mt * AT_M + r;
          if (qr < p.S) {
            const uint4 v = *reinterpret_cast<const uint4*>(s_out + r * 128 + ((chunk ^ (r & 7)) << 4));
            st_v4(obase + (int64_t)qr * p.D, v);
          }
        }
      }
      if (dbg) { long long t = clock64(); te[2] += t - e0; e0 = t; }
    };

    float l_prev = 0.f;
    long long tk[6] = {0, 0, 0, 0, 0, 0};
    for (int it = 0; it < n_items; ++it) {
      long long t0 = dbg ? clock64() : 0;
      const int item = blockIdx.x + it * gridDim.x;
      const int b = it & 1;
      const int mt = item % p.mtiles;
      const int bh = item / p.mtiles;
      const int bb = bh / p.H;
      const int qrow = mt * AT_M + r_local;
      const bool warp_valid = (mt * AT_M + wq * 32) < p.S;  // warp-uniform
      const uint8_t* km = p.key_mask != nullptr ? p.key_mask + (int64_t)bb * p.S : nullptr;
      const uint32_t tb = tmem_base + lane_off + b * p.nb;
      const int kmax_warp = p.causal ? min(p.S, mt * AT_M + wq * 32 + 32) : p.S;  // keys any row of this warp sees

      mbar_wait(&s_full[b], (it >> 1) & 1);
      tcgen05_fence_after();
      if (dbg) { long long t = clock64(); tk[0] += t - t0; t0 = t; }
      uint32_t sv[HALF16 * 16];
      float m = -INFINITY;
      if (warp_valid) {
        __syncwarp();
        // every thread loads HALF16 chunks; a chunk past Npad (odd chunk count) only holds stale columns of the
        // 32-rounded buffer and is masked below like any key >= S
#pragma unroll
        for (int ch = 0; ch < HALF16; ++ch)
          tmem_ld_32x32b_x16(tb + (my_c0 + ch) * 16, *reinterpret_cast<uint32_t(*)[16]>(&sv[ch * 16]));
        tmem_wait_ld();
        if (dbg) { long long t = clock64(); tk[1] += t - t0; t0 = t; }
        // sv stays read-only after the load (writing -inf into it makes ptxas keep a second copy of the row).
        // GENERAL_MASK = false (vision tower): only the chunk straddling S needs a key < S test; every other chunk
        // runs a select-free fast path.  GENERAL_MASK = true (text tower): causal and/or key-padding mask per key.
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int ch = 0; ch < HALF16; ++ch) {
          const int k0 = (my_c0 + ch) * 16;
          if (k0 < kmax_warp) {  // warp-uniform; otherwise the chunk is fully masked (causal) or past the sequence
            if (!GENERAL_MASK && k0 + 16 <= p.S) {
#pragma unroll
              for (int j = 0; j < 16; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(sv[ch * 16 + j]));
            } else {
              const uint32_t okbits = chunk_mask<GENERAL_MASK>(p, km, k0, qrow);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                m4[j & 3] = fmaxf(m4[j & 3], (okbits >> j) & 1u ? __uint_as_float(sv[ch * 16 + j]) : -INFINITY);
            }
          }
        }
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      s_max[hf * 128 + r_local] = m;
      at_bar_sync();  // every S read of this item has retired (both halves) and the partial maxima are visible
      if (dbg) { long long t = clock64(); tk[2] += t - t0; t0 = t; }
      float l = 0.f;
      if (warp_valid) {
        m = fmaxf(m, s_max[(hf ^ 1) * 128 + r_local]);
        const float off = (m == -INFINITY ? 0.f : m) * c;
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        // bf16 by truncation (integer pipe: F2FP shares the SFU pipe with MUFU.EX2); the row sum uses the same truncated
        // values, so numerator and denominator of O see identical probabilities.
#pragma unroll
        for (int ch = 0; ch < HALF16; ++ch) {
          const int k0 = (my_c0 + ch) * 16;
          uint32_t pk[8];
          if (k0 >= kmax_warp) {
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = 0u;
          } else if (!GENERAL_MASK && k0 + 16 <= p.S) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t e0 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(sv[ch * 16 + 2 * j]), c, -off))) & 0xffff0000u;
              const uint32_t e1 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(sv[ch * 16 + 2 * j + 1]), c, -off))) & 0xffff0000u;
              l4[j & 3] += __uint_as_float(e0) + __uint_as_float(e1);
              pk[j] = __byte_perm(e0, e1, 0x7632);
            }
          } else {
            const uint32_t okbits = chunk_mask<GENERAL_MASK>(p, km, k0, qrow);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t e0 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(sv[ch * 16 + 2 * j]), c, -off))) & 0xffff0000u;
              uint32_t e1 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(sv[ch * 16 + 2 * j + 1]), c, -off))) & 0xffff0000u;
              e0 = (okbits >> (2 * j)) & 1u ? e0 : 0u;
              e1 = (okbits >> (2 * j + 1)) & 1u ? e1 : 0u;
              l4[j & 3] += __uint_as_float(e0) + __uint_as_float(e1);
              pk[j] = __byte_perm(e0, e1, 0x7632);
            }
          }
          tmem_st_32x32b_x8(tb + (my_c0 + ch) * 8, pk);
        }
        l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        tmem_wait_st();
      }
      s_sum[hf * 128 + r_local] = l;
      tcgen05_fence_before();
      mbar_arrive(&p_full[b]);
      at_bar_sync();  // partial row sums visible to the partner thread
      if (dbg) { long long t = clock64(); tk[3] += t - t0; t0 = t; }
      const float l_total = l + s_sum[(hf ^ 1) * 128 + r_local];
      if (it > 0) epilogue(it - 1, l_prev);  // one item late: P.V(it-1) had a whole softmax of time to finish
      l_prev = l_total;
      at_bar_sync();  // s_max / s_sum / the output staging tile may be overwritten by the next item
      if (dbg) { long long t = clock64(); tk[4] += t - t0; t0 = t; }
    }
    if (dbg)
      printf("attn dbg warp %d items %d cycles/item: wait_s %lld ld %lld max+bar %lld exp+st+bar %lld epi+bar %lld | epi: wait_o %lld "
             "ldO %lld scale+store %lld\n", warp, n_items, tk[0] / n_items, tk[1] / n_items, tk[2] / n_items, tk[3] / n_items,
             tk[4] / n_items, te[0] / n_items, te[1] / n_items, te[2] / n_items);
    if (n_items > 0) epilogue(n_items - 1, l_prev);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<AT_TMEM_COLS>(tmem_base);
  }
}

template <int HALF16, bool GENERAL_MASK>
int launch_attention(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const AttnParams& p, size_t smem, int grid,
                     cudaStream_t s) {
  static size_t smem_set = 0;
  if (smem > smem_set) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(attention_tc_kernel<HALF16, GENERAL_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  attention_tc_kernel<HALF16, GENERAL_MASK><<<grid, AT_THREADS, smem, s>>>(tmQ, tmKV, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, attention_tc_kernel<HALF16, GENERAL_MASK>);
    set_last_error("attention_tc_kernel<%d> launch: %s (regs=%d maxThreads=%d static_smem=%zu max_dyn_smem=%d requested_dyn=%zu "
                   "threads=%d grid=%d)", HALF16, cudaGetErrorString(e), fa.numRegs, fa.maxThreadsPerBlock,
                   fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, smem, AT_THREADS, grid);
    return (int)e;
  }
  return 0;
}

}  // namespace
}  // namespace vlmclip

using namespace vlmclip;

extern "C" int vlmclip_attention_fwd(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H,
                                     int causal, float scale, void* stream) {
  VLMCLIP_CHECK_ARG(qkv && out, "attention: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && S > 0 && H > 0, "attention: bad dims B=%d S=%d H=%d", B, S, H);
  VLMCLIP_CHECK_ARG(S <= 512, "attention: S=%d exceeds the on-chip K/V limit (512)", S);
  VLMCLIP_CHECK_ARG((uintptr_t)qkv % 16 == 0 && (uintptr_t)out % 16 == 0, "attention: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  // Dispatch.  ViT-L/14 (S = 257) exceeds the two-S-buffer TMEM map of the tcgen05 kernel, and for the short causal
  // text sequences (S = 77: 60 % of a 128-row tile would be padding, every chunk takes the masked path) the mma.sync
  // variant measures faster on B200 (44 us vs 80 us at B=256, H=8); VLMCLIP_ATTN_FORCE_TC=1 forces the tcgen05 kernel.
  static const bool force_tc = []() {
    const char* e = getenv("VLMCLIP_ATTN_FORCE_TC");
    return e != nullptr && e[0] == '1';
  }();
  if (S > AT_MAX_NPAD || (!force_tc && (causal != 0 || key_mask != nullptr) && S <= 128))
    return attention_fwd_mma_sync(qkv, out, key_mask, B, S, H, causal, scale, s);

  AttnParams p;
  p.key_mask = key_mask;
  p.out = (__nv_bfloat16*)out;
  p.B = B;
  p.S = S;
  p.H = H;
  p.D = H * AT_HD;
  p.causal = causal;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.Npad = (S + 15) / 16 * 16;
  p.nb = (p.Npad + 31) / 32 * 32;
  p.mtiles = (S + AT_M - 1) / AT_M;
  p.num_items = B * H * p.mtiles;
  p.kv_bytes = (uint32_t)p.Npad * 128u;
  p.kv_stride = (p.kv_bytes + 1023u) & ~1023u;
  p.stage_bytes = AT_Q_BYTES + 2 * p.kv_stride;
  {
    const char* e = getenv("VLMCLIP_ATTN_DEBUG");
    p.debug = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  int nstage = (int)((207u * 1024u) / p.stage_bytes);  // 227 KB - 16 KB output tile - barriers / exchange arrays
  p.nstage = nstage < 2 ? 2 : (nstage > AT_MAX_STAGES ? AT_MAX_STAGES : nstage);
  const size_t ctrl = (2 * AT_MAX_STAGES + 8) * 8 + 512 * sizeof(float) + 16;
  p.out_stage_off = (uint32_t)(((size_t)p.nstage * p.stage_bytes + ctrl + 1023) & ~(size_t)1023);
  const size_t smem = (size_t)p.out_stage_off + AT_M * 128;
  CUtensorMap tmQ, tmKV;
  const int64_t rows = (int64_t)B * S;
  int rc = make_tmap_bf16(&tmQ, qkv, rows, 3 * (int64_t)p.D, 3 * (int64_t)p.D, AT_M);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmKV, qkv, rows, 3 * (int64_t)p.D, 3 * (int64_t)p.D, p.Npad);
  if (rc) return rc;
  const int grid = p.num_items < sm_count() ? p.num_items : sm_count();
  const int half16 = (p.Npad / 16 + 1) / 2;
  count_launch(1);
  const bool general = causal != 0 || key_mask != nullptr;
#define VLMCLIP_AT_CASE(H16)                                                                  \
  case H16:                                                                                   \
    return general ? launch_attention<H16, true>(tmQ, tmKV, p, smem, grid, s)                 \
                   : launch_attention<H16, false>(tmQ, tmKV, p, smem, grid, s);
  switch (half16) {
    VLMCLIP_AT_CASE(1)
    VLMCLIP_AT_CASE(2)
    VLMCLIP_AT_CASE(3)
    VLMCLIP_AT_CASE(4)
    VLMCLIP_AT_CASE(5)
    VLMCLIP_AT_CASE(6)
    default:
      return general ? launch_attention<7, true>(tmQ, tmKV, p, smem, grid, s)
                     : launch_attention<7, false>(tmQ, tmKV, p, smem, grid, s);
  }
#undef VLMCLIP_AT_CASE
}

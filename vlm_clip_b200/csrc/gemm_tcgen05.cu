// Dense layer GEMM for the frozen CLIP towers:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
// Replaces the aten::addmm calls behind nn.Linear at HF modeling_clip.py:310-312,334 (q/k/v/out),
// :348-350 (fc1 / quick_gelu / fc2) and :209 (patch conv as GEMM).
//
// B200 design (one CTA per SM, persistent over output tiles):
//   warp 0 (1 thread)  TMA producer: A tile [128 x 64] and W tile [BLOCK_N x 64] per stage,
//                      128-byte swizzle, completion by mbarrier complete_tx.
//   warp 1 (1 thread)  tcgen05.mma issuer (cta_group::1, M=128, N=BLOCK_N, K=16 x 4 per stage),
//                      fp32 accumulators in TMEM, double buffered (2 x BLOCK_N columns) so the
//                      epilogue of tile i overlaps the main loop of tile i+1.
//   warp 2             TMEM allocator / deallocator.
//   warps 4..7         epilogue: tcgen05.ld 32 lanes x 32 columns per warp -> registers ->
//                      LN-fold / bias / activation / residual -> bf16 (or fp32) -> global.
//   Tile order is n-fastest so the CTAs running at one moment share a few A row-blocks through L2.
#include "common.cuh"

namespace vlmclip {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 128;

struct GemmParams {
  void* C;
  const float* bias;
  const __nv_bfloat16* residual;
  const float* row_stats;  // [M][2] mean, rstd
  const float* col_c;      // [N]
  int64_t ldc, ldr;
  int M, N, K;
  int act;
  int out_fp32;
  int m_tiles, n_tiles, k_blocks;
};

template <int BLOCK_N, int STAGES>
struct SmemLayout {
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr uint32_t NUM_BARS = 2 * STAGES + 4;
  static constexpr uint32_t TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16;
  static constexpr uint32_t DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return quick_gelu(v);
  if (act == 2) return gelu_erf(v);
  if (act == 3) return fmaxf(v, 0.0f);
  return v;
}

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  constexpr int TMEM_COLS = 2 * BLOCK_N;  // double-buffered accumulator (power of two: 256 or 512)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.n_tiles;
      const int n_blk = tile - m_blk * p.n_tiles;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
        tma_load_2d(sA + stage * L::A_BYTES, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
        tma_load_2d(sB + stage * L::B_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
    uint32_t stage = 0, phase = 0;
    uint32_t abuf = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[abuf], aphase ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + abuf * BLOCK_N;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint64_t a_desc = make_umma_desc_sw128(smem_u32(sA + stage * L::A_BYTES));
        const uint64_t b_desc = make_umma_desc_sw128(smem_u32(sB + stage * L::B_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // advance 16 bf16 = 32 B inside the 128-B swizzle span: +2 in the (addr >> 4) field
          umma_bf16_ss(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(&tfull_bar[abuf]);  // accumulator complete -> epilogue
      abuf ^= 1u;
      if (abuf == 0) aphase ^= 1u;
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t abuf = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.n_tiles;
      const int n_blk = tile - m_blk * p.n_tiles;
      const int row = m_blk * BLOCK_M + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      float mean = 0.f, rstd = 1.f;
      if (p.row_stats != nullptr && row_ok) {
        const float2 st = *reinterpret_cast<const float2*>(p.row_stats + 2 * (int64_t)row);
        mean = st.x;
        rstd = st.y;
      }
      mbar_wait(&tfull_bar[abuf], aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + abuf * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the row-predicated stores
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        tmem_wait_ld();
        const int col0 = n_blk * BLOCK_N + c * 32;
        if (col0 >= p.N) continue;  // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        // columns beyond N inside this chunk (N % 8 == 0): handled per 8-column group
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col0 + g * 8;
          if (col >= p.N) break;  // warp-uniform
          if (p.row_stats != nullptr) {
            const float4 c0 = __ldg(reinterpret_cast<const float4*>(p.col_c + col));
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(p.col_c + col + 4));
            const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g * 8 + j] = rstd * (v[g * 8 + j] - mean * cc[j]);
          }
          if (p.bias != nullptr) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g * 8 + j] += bb[j];
          }
          if (p.act != 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g * 8 + j] = apply_act(v[g * 8 + j], p.act);
          }
          if (row_ok) {
            if (p.residual != nullptr) {
              const uint4 rr = ld_nc_v4(p.residual + (int64_t)row * p.ldr + col);
              v[g * 8 + 0] += bf16_lo(rr.x);
              v[g * 8 + 1] += bf16_hi(rr.x);
              v[g * 8 + 2] += bf16_lo(rr.y);
              v[g * 8 + 3] += bf16_hi(rr.y);
              v[g * 8 + 4] += bf16_lo(rr.z);
              v[g * 8 + 5] += bf16_hi(rr.z);
              v[g * 8 + 6] += bf16_lo(rr.w);
              v[g * 8 + 7] += bf16_hi(rr.w);
            }
            if (p.out_fp32) {
              float* out = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col;
              *reinterpret_cast<float4*>(out) =
                  make_float4(v[g * 8 + 0], v[g * 8 + 1], v[g * 8 + 2], v[g * 8 + 3]);
              *reinterpret_cast<float4*>(out + 4) =
                  make_float4(v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]);
            } else {
              __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + col;
              uint4 o;
              o.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
              o.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
              o.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
              o.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
              st_v4(out, o);
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tempty_bar[abuf]);
      abuf ^= 1u;
      if (abuf == 0) aphase ^= 1u;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// bf16 row-major [rows, cols] (ld elements) -> 2-D map with box [box_rows x 64 cols], 128-B swizzle
int make_tmap_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled entry point not available (driver too old or no GPU)");
    return -2;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (base=%p rows=%lld cols=%lld ld=%lld)", (int)r,
                   base, (long long)rows, (long long)cols, (long long)ld);
    return -3;
  }
  return 0;
}

template <int BLOCK_N, int STAGES>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  static bool attr_set = false;  // benign race: setting the attribute twice is harmless
  if (!attr_set) {
    VLMCLIP_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BLOCK_N, STAGES>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
    attr_set = true;
  }
  p.n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_bf16_tn_kernel<BLOCK_N, STAGES><<<grid, GEMM_THREADS, L::DYN_BYTES, stream>>>(tmA, tmB, p);
  return report_cuda(cudaGetLastError(), "gemm_bf16_tn_kernel launch");
}

}  // namespace

void count_launch(int n);

}  // namespace vlmclip

using namespace vlmclip;

extern "C" int vlmclip_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc,
                                 const float* bias, const void* residual, int64_t ldr, const float* row_stats,
                                 const float* col_c, int M, int N, int K, int act, int out_fp32, void* stream) {
  VLMCLIP_CHECK_ARG(A && W && C, "gemm: null A/W/C pointer");
  VLMCLIP_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: non-positive dims M=%d N=%d K=%d", M, N, K);
  VLMCLIP_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "gemm: K and N must be multiples of 8 (K=%d N=%d)", K, N);
  VLMCLIP_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, "gemm: lda/ldw/ldc must be multiples of 8");
  VLMCLIP_CHECK_ARG(lda >= K && ldw >= K && ldc >= N, "gemm: leading dimension smaller than row length");
  VLMCLIP_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)C % 16 == 0),
                    "gemm: A/W/C must be 16-byte aligned");
  VLMCLIP_CHECK_ARG(act >= 0 && act <= 3, "gemm: unknown activation %d", act);
  VLMCLIP_CHECK_ARG((row_stats == nullptr) == (col_c == nullptr), "gemm: row_stats and col_c go together");
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
  p.ldc = ldc;
  p.ldr = ldr;
  p.M = M;
  p.N = N;
  p.K = K;
  p.act = act;
  p.out_fp32 = out_fp32;
  p.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  p.k_blocks = (K + BLOCK_K - 1) / BLOCK_K;

  const bool wide = N > 128;
  const int block_n = wide ? 256 : 128;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16(&tmA, A, M, K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, W, N, K, ldw, block_n);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  count_launch(1);
  if (wide) return launch_gemm<256, 4>(tmA, tmB, p, s);
  return launch_gemm<128, 6>(tmA, tmB, p, s);
}

// microbenchmark: cost of the softmax chunk loop of attention_pp_kernel for ONE warp per scheduler:
//   mode 0: exp2 mix on registers only        mode 1: + tcgen05.ld x32 of the next chunk (prefetch pattern)
//   mode 2: + tcgen05.st x16 of the packed P  mode 3: both (the kernel's loop)   mode 4: both + 16 FMNMX3 max + vote
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../vlm_clip_b200/csrc/common.cuh"
using namespace vlmclip;
namespace vlmclip { void set_last_error(const char*, ...) {} int report_cuda(cudaError_t, const char*) { return 0; } }

template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, long long* cyc, int iters, float c, float off) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = slot + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t sa[32], sb[32];
  for (int j = 0; j < 32; ++j) { sa[j] = __float_as_uint(threadIdx.x * 0.001f + j * 0.01f); sb[j] = sa[j] ^ 0x100; }
  // initialise TMEM columns we read
  for (int ch = 0; ch < 7; ++ch) tmem_st_32x32b_x16(tb + ch * 32, *reinterpret_cast<uint32_t(*)[16]>(&sa[0]));
  for (int ch = 0; ch < 7; ++ch) tmem_st_32x32b_x16(tb + ch * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&sa[16]));
  tmem_wait_st();
  float l4[4] = {0, 0, 0, 0};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 1 || MODE >= 3) tmem_ld_32x32b_x32(tb, sa);
    for (int ch = 0; ch < 7; ++ch) {
      if (MODE == 1 || MODE >= 3) {
        tmem_wait_ld();
        if (ch + 1 < 7) { if (ch & 1) tmem_ld_32x32b_x32(tb + (ch + 1) * 32, sa); else tmem_ld_32x32b_x32(tb + (ch + 1) * 32, sb); }
      }
      uint32_t (&cur)[32] = (ch & 1) ? sb : sa;
      if (MODE == 4) {
        float m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
        for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(cur[j]));
        const float mm = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * c;
        if (__any_sync(0xffffffffu, mm - off > 1e20f)) off += 1.f;
      }
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t e0 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(cur[2 * j]), c, -off))) & 0xffff0000u;
        const uint32_t e1 = __float_as_uint(fast_exp2(fmaf(__uint_as_float(cur[2 * j + 1]), c, -off))) & 0xffff0000u;
        l4[j & 3] += __uint_as_float(e0) + __uint_as_float(e1);
        pk[j] = __byte_perm(e0, e1, 0x7632);
      }
      if (MODE >= 2) tmem_st_32x32b_x16(tb + ch * 16, pk);
      else { cur[0] ^= pk[3]; cur[5] ^= pk[7]; }
    }
    if (MODE >= 2) tmem_wait_st();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l4[0] + l4[1] + l4[2] + l4[3] + __uint_as_float(sa[0]) + __uint_as_float(sb[5]);
  if (threadIdx.x == 0) *cyc = t1 - t0;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const int iters = 200;
#define RUN(M) k<M><<<1, 128>>>(out, cyc, iters, 0.001f, 0.5f); cudaDeviceSynchronize(); \
  printf("mode %d: %.1f cycles per 32-key chunk (one warp per scheduler) err=%s\n", M, (double)*cyc / (iters * 7.0), cudaGetErrorString(cudaGetLastError()));
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
  return 0;
}

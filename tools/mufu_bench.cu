// microbenchmark: MUFU.EX2 issue rate for 1..4 warps per SM sub-partition, alone and in the softmax instruction mix
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float c, float off) {
  float x[32];
  for (int j = 0; j < 32; ++j) x[j] = threadIdx.x * 0.001f + j;
  float l4[4] = {0, 0, 0, 0};
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = ex2(x[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        unsigned e0 = __float_as_uint(ex2(fmaf(x[2 * j], c, -off))) & 0xffff0000u;
        unsigned e1 = __float_as_uint(ex2(fmaf(x[2 * j + 1], c, -off))) & 0xffff0000u;
        l4[j & 3] += __uint_as_float(e0) + __uint_as_float(e1);
        acc ^= __byte_perm(e0, e1, 0x7632);
        x[2 * j] = __uint_as_float(e0) + 1.f;
        x[2 * j + 1] = __uint_as_float(e1) + 1.f;
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 32; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + l4[0] + l4[1] + l4[2] + l4[3] + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const int iters = 1000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, iters, 1.1f, 0.5f); else k<1><<<1, warps * 32>>>(out, cyc, iters, 1.1f, 0.5f);
      cudaDeviceSynchronize();
      printf("mode %d warps/SMSP %d: %.2f cycles per MUFU warp-instr per SMSP (%.2f per warp)\n", mode, warps / 4,
             (double)*cyc / (iters * 32.0 * (warps / 4)), (double)*cyc / (iters * 32.0));
    }
  return 0;
}

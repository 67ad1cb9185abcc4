// Micro-benchmark: TMEM read / write throughput per SM as seen by softmax-style warps (tcgen05.ld / st .32x32b.x32),
// with 1..3 CTAs of 128 threads per SM (each warp reads its own 32-lane quadrant). Answers "what bounds the head_dim-64
// self-attention kernel": a 128x64 fp32 score tile is 32 KB of TMEM reads against 256 tensor cycles of MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rich-text-to-image_b200/csrc tools/micro/tmem_bw.cu -o rich-text-to-image_b200/build/tmem_bw
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace rtti;

template <int MODE>   // 0: ld only, 1: st only, 2: ld + 64 x ex2 per thread (MUFU alongside)
__global__ void __launch_bounds__(128) tmem_kernel(int iters, unsigned long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<128>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = slot + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t r[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) r[i] = threadIdx.x + i;
  tmem_st32(t, r); tmem_st32(t + 32, r + 32); tmem_wait_st();
  __syncthreads();
  float acc = 0.f;
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
      tmem_ld32(t, r); tmem_ld32(t + 32, r + 32);
      tmem_wait_ld_regs32(r); tmem_wait_ld_regs32(r + 32);
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc += ex2_approx(__uint_as_float(r[i]) * 1e-9f);
      } else {
        acc += __uint_as_float(r[0] ^ r[63]);   // static indices only: a dynamic index would move r[] to local memory
      }
    } else {
      r[0] += 1;
      tmem_st32(t, r); tmem_st32(t + 32, r + 32); tmem_wait_st();
    }
  }
  const long long c1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(c1 - c0);
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(slot);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* cyc; float* sink;
  cudaMalloc(&cyc, sizeof(unsigned long long) * sms * 4);
  cudaMalloc(&sink, 4);
  const int iters = 4096;
  const char* names[3] = {"tcgen05.ld only", "tcgen05.st only", "tcgen05.ld + 64 ex2/thread"};
  for (int mode = 0; mode < 3; ++mode)
    for (int per_sm = 1; per_sm <= 3; ++per_sm) {
      const int grid = sms * per_sm;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) tmem_kernel<0><<<grid, 128>>>(iters, cyc, sink);
        if (mode == 1) tmem_kernel<1><<<grid, 128>>>(iters, cyc, sink);
        if (mode == 2) tmem_kernel<2><<<grid, 128>>>(iters, cyc, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      }
      unsigned long long h[1024];
      cudaMemcpy(h, cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < grid; ++i) mean += (double)h[i];
      mean /= grid;
      const double bytes_per_cta = (double)iters * 128 * 64 * 4;   // one 128 x 64 fp32 tile per iteration
      printf("%-28s %d CTA/SM (x128 threads): %.0f cycles per 128x64 fp32 tile per CTA, %.1f B/clk per CTA, %.1f B/clk per SM\n",
             names[mode], per_sm, mean / iters, bytes_per_cta / mean, per_sm * bytes_per_cta / mean);
    }
  return 0;
}

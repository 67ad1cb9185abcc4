// System-scope flag primitives for kernels that synchronise ranks through peer-mapped (NVLink) memory.
#pragma once
#include <cuda_runtime.h>

namespace rtti {

constexpr int PEER_MAX_WORLD = 16;
constexpr long long PEER_SPIN_LIMIT = 8000000LL;  // x >= 200 ns: a few seconds, then the caller raises its error word

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_volatile_f32(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];\n" : "=f"(v) : "l"(p));
  return v;
}
// Spin until *flag >= seq (wrap-safe). Returns false on timeout.
__device__ __forceinline__ bool wait_seq(const unsigned int* flag, unsigned int seq) {
  long long spins = 0;
  while ((int)(ld_acquire_sys(flag) - seq) < 0) {
    __nanosleep(200);
    if (++spins > PEER_SPIN_LIMIT) return false;
  }
  return true;
}

}  // namespace rtti

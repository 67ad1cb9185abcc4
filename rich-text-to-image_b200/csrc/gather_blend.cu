// Fused all-gather + region blend + CFG + Euler update over NVLink peer memory (multi-GPU region parallelism).
//
// After the sharded UNet passes of a step every rank owns the noise predictions of its passes
// ("slots", [n] fp16 each) in a symmetric buffer that all peers have mapped. Instead of
// ncclAllGather followed by a blend kernel, ONE kernel per rank
//   1. publishes "my slots of step s are written" (release store of s to its flag word),
//   2. waits (acquire loads over NVLink) until every owner it reads from has published step s,
//   3. PULLS each slot straight from its owner's memory with 128-bit peer loads while computing
//      the masked region sums + classifier-free guidance + Euler update of
//      models/region_diffusion_sdxl.py:810-845 — replicated on every rank, bit-identical results,
//      so no broadcast of the latents is needed.
// A peer that never publishes makes the wait give up after ~4 s and raise the error word flags[1] (checked by the host).
// Slot buffers are double-buffered by step parity, which makes re-use safe without a trailing barrier:
// a rank overwrites parity p at step s+2 only after its step s+1 kernel saw every peer publish s+1,
// and a peer publishes s+1 only after its step-s kernel (the last reader of parity p) has finished.
#include <cuda_fp16.h>

#include "rtti_internal.h"

namespace rtti {

constexpr int GB_MAX_WORLD = 16;
constexpr int GB_MAX_SLOTS = 24;

struct GatherBlendParams {
  const __half* peer_slots[GB_MAX_WORLD];  // per rank: fp16 [2][n_slots][n]
  unsigned int* peer_flags[GB_MAX_WORLD];  // per rank: one uint32 step counter
  int slot_owner[GB_MAX_SLOTS];
  int world, rank, n_slots, n_regions;
  long long n;
  float guidance, dt_sigma;
  unsigned int step_id;
  const float* masks;        // [n_regions, n]
  __half* eps_out;           // [n]
  const __half* latents;     // [n] or null
  __half* latents_out;
  const __half* latents_ref; // [n] or null: needs slots n_regions+1 (C) and n_regions+2 (D)
  __half* latents_ref_out;
};

struct alignas(16) H8 { __half2 v[4]; };

__device__ __forceinline__ H8 ld_peer(const __half* p) {
  // volatile: never served from a stale L1 line (peer memory bypasses L2, B300_MICROARCH.md NVLink section)
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return *reinterpret_cast<H8*>(&r);
}
__device__ __forceinline__ void up8(const H8& h, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ H8 pk8(const float* f) {
  H8 h;
#pragma unroll
  for (int i = 0; i < 4; ++i) h.v[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return h;
}

__global__ void __launch_bounds__(128) gather_blend_kernel(const GatherBlendParams p) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p.peer_flags[p.rank]), "r"(p.step_id) : "memory");
  }
  if (threadIdx.x < p.world && threadIdx.x != p.rank) {
    unsigned int v;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p.peer_flags[threadIdx.x]) : "memory");
      if ((int)(v - p.step_id) < 0) {
        __nanosleep(500);
        if (++spins > 8000000LL) {  // ~4 s: a peer never published (crashed / diverged) — flag it, do not hang the GPU
          p.peer_flags[p.rank][1] = 0xDEADu;
          break;
        }
      }
    } while ((int)(v - p.step_id) < 0);
  }
  __syncthreads();
  const long long v8 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v8 * 8 >= p.n) return;
  const size_t par = (size_t)(p.step_id & 1u) * p.n_slots * p.n;
  auto slot = [&](int s) { return p.peer_slots[p.slot_owner[s]] + par + (size_t)s * p.n + v8 * 8; };
  float eu[8], msum[8], et[8];
  up8(ld_peer(slot(0)), eu);
#pragma unroll
  for (int i = 0; i < 8; ++i) { msum[i] = 0.f; et[i] = 0.f; }
  for (int r = 0; r < p.n_regions; ++r) {
    float e[8];
    up8(ld_peer(slot(1 + r)), e);
    const float4 m0 = *reinterpret_cast<const float4*>(p.masks + (size_t)r * p.n + v8 * 8);
    const float4 m1 = *reinterpret_cast<const float4*>(p.masks + (size_t)r * p.n + v8 * 8 + 4);
    const float m[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) { msum[i] += m[i]; et[i] = fmaf(e[i], m[i], et[i]); }
  }
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float u = eu[i] * msum[i]; o[i] = u + p.guidance * (et[i] - u); }
  const H8 oh = pk8(o);
  *reinterpret_cast<H8*>(p.eps_out + v8 * 8) = oh;
  if (p.latents != nullptr) {
    float x[8], e16[8];
    up8(*reinterpret_cast<const H8*>(p.latents + v8 * 8), x);
    up8(oh, e16);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(e16[i], p.dt_sigma, x[i]);
    *reinterpret_cast<H8*>(p.latents_out + v8 * 8) = pk8(x);
  }
  if (p.latents_ref != nullptr) {
    float c[8], d[8], x[8], e16[8];
    up8(ld_peer(slot(p.n_regions + 1)), c);
    up8(ld_peer(slot(p.n_regions + 2)), d);
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = c[i] + p.guidance * (d[i] - c[i]);
    up8(pk8(c), e16);
    up8(*reinterpret_cast<const H8*>(p.latents_ref + v8 * 8), x);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(e16[i], p.dt_sigma, x[i]);
    *reinterpret_cast<H8*>(p.latents_ref_out + v8 * 8) = pk8(x);
  }
}

}  // namespace rtti

using namespace rtti;

extern "C" int rtti_gather_blend_step(const void* const* peer_slots, void* const* peer_flags, int world, int rank,
                                      const int* slot_owner, int n_slots, int n_regions, const float* masks,
                                      long long n, float guidance, void* eps_out, const void* latents,
                                      void* latents_out, const void* latents_ref, void* latents_ref_out,
                                      float dt_sigma, unsigned int step_id, void* stream) {
  if (!peer_slots || !peer_flags || !slot_owner || !masks || !eps_out) return RTTI_ERR_ARG;
  if (world < 1 || world > GB_MAX_WORLD || rank < 0 || rank >= world) return RTTI_ERR_ARG;
  if (n_regions < 1 || n_slots < n_regions + 1 || n_slots > GB_MAX_SLOTS || n < 8) return RTTI_ERR_ARG;
  if (n % 8 != 0) return RTTI_ERR_SHAPE;
  if ((latents == nullptr) != (latents_out == nullptr)) return RTTI_ERR_ARG;
  if ((latents_ref == nullptr) != (latents_ref_out == nullptr)) return RTTI_ERR_ARG;
  if (latents_ref != nullptr && n_slots < n_regions + 3) return RTTI_ERR_ARG;
  GatherBlendParams p{};
  for (int r = 0; r < world; ++r) {
    if (!peer_slots[r] || !peer_flags[r]) return RTTI_ERR_ARG;
    if ((uintptr_t)peer_slots[r] & 15) return RTTI_ERR_ALIGN;
    p.peer_slots[r] = (const __half*)peer_slots[r];
    p.peer_flags[r] = (unsigned int*)peer_flags[r];
  }
  for (int s = 0; s < n_slots; ++s) {
    if (slot_owner[s] < 0 || slot_owner[s] >= world) return RTTI_ERR_ARG;
    p.slot_owner[s] = slot_owner[s];
  }
  p.world = world; p.rank = rank; p.n_slots = n_slots; p.n_regions = n_regions; p.n = n;
  p.guidance = guidance; p.dt_sigma = dt_sigma; p.step_id = step_id;
  p.masks = masks; p.eps_out = (__half*)eps_out;
  p.latents = (const __half*)latents; p.latents_out = (__half*)latents_out;
  p.latents_ref = (const __half*)latents_ref; p.latents_ref_out = (__half*)latents_ref_out;
  const long long nv = n / 8;
  gather_blend_kernel<<<(int)((nv + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? RTTI_OK : RTTI_ERR_CUDA;
}

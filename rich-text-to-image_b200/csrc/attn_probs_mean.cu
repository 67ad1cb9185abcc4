// Self-attention token-map capture for sm_100a: accum[q, k] += mean_h softmax(scale Q_h K_h^T)[q, k].
//
// The reference materialises the full probability tensor of every attention call and averages it
// over heads on every call (models/attention_processor.py:1157-1159, 166-171, 1181), then the
// token-map hook copies the conditional row to the CPU and sums it there
// (models/region_diffusion_sdxl.py:986-992). Here the flash kernel (attn_fwd.cu) leaves only the
// per-row log-sum-exp; this kernel recomputes the 128x128 score tiles on tcgen05 tensor cores,
// loops over the heads inside the CTA (so the head mean needs no atomics and is deterministic) and
// adds the tile into an fp32 accumulator that stays on the device.
//
// CTA = one (128 query rows) x (128 keys) tile, all heads. Warps 0-3: exp + accumulate (thread = row =
// TMEM lane); warp 4: TMA producer; warp 5: MMA issuer. S is double-buffered in TMEM so the MMA of
// head h+1 overlaps the exponentials of head h.
#include "ptx.cuh"
#include "rtti_internal.h"

namespace rtti {

struct ProbsMeanParams {
  int heads, head_dim, n_q, n_k, ksteps_qk;
  float scale_log2, inv_heads;
  const float* lse;  // [heads, n_q]
  float* accum;      // [n_q, n_k]
};

template <int NDCH>
struct PMCfg {
  static constexpr int NSTAGE = (NDCH == 3) ? 1 : 2;
  static constexpr int TILE = 128 * 128;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = NSTAGE * NDCH * TILE;
  static constexpr int OFF_BAR = 2 * NSTAGE * NDCH * TILE;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
};

template <int NDCH>
__global__ void __launch_bounds__(192, 1)
attn_probs_mean_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                       const ProbsMeanParams p) {
  using C = PMCfg<NDCH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* full = bars;         // [NSTAGE] TMA -> MMA
  uint64_t* empty = bars + 2;    // [NSTAGE] MMA -> TMA
  uint64_t* s_full = bars + 4;   // [2] MMA -> exp warps
  uint64_t* s_empty = bars + 6;  // [2] exp warps -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, q0 = blockIdx.y * 128;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k);
    for (int i = 0; i < C::NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 128); }
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      for (int h = 0; h < p.heads; ++h) {
        const int st = h % C::NSTAGE;
        mbar_wait(&empty[st], ((h / C::NSTAGE) & 1) ^ 1);
        mbar_expect_tx(&full[st], 2 * NDCH * C::TILE);
#pragma unroll
        for (int c = 0; c < NDCH; ++c) {
          tma_load_4d(smem + C::OFF_Q + (st * NDCH + c) * C::TILE, &tm_q, &full[st], 64 * c, h, q0, 0);
          tma_load_4d(smem + C::OFF_K + (st * NDCH + c) * C::TILE, &tm_k, &full[st], 64 * c, h, k0, 0);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc_f16(128, 128, 0, 0);
      const uint32_t smem_base = smem_u32(smem);
      for (int h = 0; h < p.heads; ++h) {
        const int st = h % C::NSTAGE, sb = h & 1;
        mbar_wait(&full[st], (h / C::NSTAGE) & 1);
        mbar_wait(&s_empty[sb], ((h >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kk = 0; kk < p.ksteps_qk; ++kk) {
          const int c = kk >> 2, k16 = kk & 3;
          const uint64_t da = umma_desc_sw128(smem_base + C::OFF_Q + (st * NDCH + c) * C::TILE + k16 * 32, 0, 1024);
          const uint64_t db = umma_desc_sw128(smem_base + C::OFF_K + (st * NDCH + c) * C::TILE + k16 * 32, 0, 1024);
          mma_f16_ss(tmem + 128 * sb, da, db, IDESC, kk > 0);
        }
        tc_commit(&empty[st]);
        tc_commit(&s_full[sb]);
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const bool row_ok = (q0 + row) < p.n_q;
    float acc[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) acc[i] = 0.f;
    for (int h = 0; h < p.heads; ++h) {
      const int sb = h & 1;
      const float neg_lse = row_ok ? -p.lse[static_cast<size_t>(h) * p.n_q + q0 + row] : 0.f;
      mbar_wait(&s_full[sb], (h >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t s[32];
        tmem_ld32(tlane + 128 * sb + 32 * c, s);
        tmem_wait_ld_regs32(s);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(s[2 * i]), p.scale_log2, neg_lse));
          const float e1 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), p.scale_log2, neg_lse));
          // the reference averages fp16 probabilities (attention_processor.py:405, 1181)
          const float2 r = __half22float2(__floats2half2_rn(e0, e1));
          acc[32 * c + 2 * i] += r.x;
          acc[32 * c + 2 * i + 1] += r.y;
        }
      }
      tc_fence_before();
      mbar_arrive(&s_empty[sb]);
    }
    if (row_ok) {
      float* dst = p.accum + static_cast<size_t>(q0 + row) * p.n_k + k0;
      const int ncol = min(128, p.n_k - k0);
      if (ncol == 128 && (p.n_k & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float4 v = *reinterpret_cast<float4*>(dst + 4 * i);
          v.x += acc[4 * i] * p.inv_heads; v.y += acc[4 * i + 1] * p.inv_heads;
          v.z += acc[4 * i + 2] * p.inv_heads; v.w += acc[4 * i + 3] * p.inv_heads;
          *reinterpret_cast<float4*>(dst + 4 * i) = v;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i < ncol) dst[i] += acc[i] * p.inv_heads;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<256>(tmem);
}

template <int NDCH>
static int launch_pm(const CUtensorMap& tq, const CUtensorMap& tk, const ProbsMeanParams& p, dim3 grid,
                     cudaStream_t stream) {
  using C = PMCfg<NDCH>;
  auto kern = attn_probs_mean_kernel<NDCH>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
      return RTTI_ERR_CUDA;
    configured = true;
  }
  kern<<<grid, 192, C::SMEM_BYTES, stream>>>(tq, tk, p);
  return cudaGetLastError() == cudaSuccess ? RTTI_OK : RTTI_ERR_CUDA;
}

}  // namespace rtti

using namespace rtti;

extern "C" int rtti_attn_probs_mean_accum(const void* q, const void* k, const float* lse, float* accum, int heads,
                                          int head_dim, int n_q, int n_k, long long q_rs, long long k_rs,
                                          float scale, void* stream) {
  if (!q || !k || !lse || !accum) return RTTI_ERR_ARG;
  if (heads < 1 || n_q < 1 || n_k < 1) return RTTI_ERR_ARG;
  if (head_dim < 8 || head_dim > 192 || (head_dim % 8) != 0) return RTTI_ERR_SHAPE;
  if (((uintptr_t)q | (uintptr_t)k) & 15) return RTTI_ERR_ALIGN;
  if ((q_rs | k_rs) & 7) return RTTI_ERR_ALIGN;
  int rc = rtti_arch_ok();
  if (rc != RTTI_OK) return rc;
  ProbsMeanParams p{};
  p.heads = heads; p.head_dim = head_dim; p.n_q = n_q; p.n_k = n_k;
  p.ksteps_qk = (head_dim + 15) / 16;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.inv_heads = 1.f / (float)heads;
  p.lse = lse; p.accum = accum;
  CUtensorMap tq, tk;
  if ((rc = make_head_map(&tq, q, head_dim, heads, n_q, 1, (long long)n_q * q_rs, q_rs, 128)) != RTTI_OK) return rc;
  if ((rc = make_head_map(&tk, k, head_dim, heads, n_k, 1, (long long)n_k * k_rs, k_rs, 128)) != RTTI_OK) return rc;
  dim3 grid((n_k + 127) / 128, (n_q + 127) / 128, 1);
  const int ndch = (head_dim + 63) / 64;
  cudaStream_t st = (cudaStream_t)stream;
  if (ndch == 1) return launch_pm<1>(tq, tk, p, grid, st);
  if (ndch == 2) return launch_pm<2>(tq, tk, p, grid, st);
  return launch_pm<3>(tq, tk, p, grid, st);
}

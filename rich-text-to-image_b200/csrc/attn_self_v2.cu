// Self-attention forward, head_dim <= 64, software-pipelined for sm_100a ("v2").
//
// Same math and C ABI as attn_fwd.cu (it is selected inside rtti_attn_fwd); different schedule. v1 runs
// QK^T -> softmax -> PV strictly in sequence per CTA and relies on a second CTA per SM to fill the
// tensor pipe; at head_dim 64 that leaves the kernel ~2x off its real bound, which is the MUFU (exp2) pipe:
// 128x128 exponentials per tile = 1024 cycles on 16 lanes/SM vs 512 cycles of MMA. v2 hides the MMAs
// completely behind the exponentials inside ONE CTA per SM:
//   * S is double-buffered in TMEM (S0 [0,128), S1 [128,256), O [256,320)): the MMA warp issues QK^T of tile
//     j+1 before the softmax of tile j has finished, then PV of tile j as soon as P_j is written;
//   * 8 softmax warps (2 threads per query row, 64 keys each) keep 2 warps per SM sub-partition so the MUFU
//     pipe stays busy across TMEM-load and barrier latencies; the row max is combined through shared memory;
//   * P_j overwrites its own S buffer as packed fp16 and feeds the PV MMA from TMEM (TS operand);
//   * the lazy O rescale (rare) waits for PV_{j-1} on its own barrier, because QK^T_{j} no longer implies it.
// Warps: 0-7 softmax/epilogue, 8 TMA producer, 9 MMA issuer + TMEM owner. 3-stage K/V ring.
#include <cstdlib>

#include "ptx.cuh"
#include "rtti_internal.h"

namespace rtti {

struct AttnV2Params {
  int batch, heads, head_dim, n_q, n_k, n_k_tiles, ksteps_qk;
  float scale_log2;
  int8_t qk_src[64];
  float* lse;
};

namespace v2 {
constexpr int KT = 128;
constexpr int NSTAGE = 3;
constexpr int Q_TILE = 128 * 128;   // bytes
constexpr int KV_TILE = KT * 128;
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_TILE;
constexpr int OFF_V = OFF_K + NSTAGE * KV_TILE;
constexpr int OFF_O = OFF_V + NSTAGE * KV_TILE;
constexpr int OFF_BAR = OFF_O + Q_TILE;
constexpr int OFF_RED = OFF_BAR + 256;              // float red_max[2 parities][2 halves][128] + red_l[2][128]
constexpr int SMEM_BYTES = OFF_RED + (2 * 2 * 128 + 2 * 128) * 4 + 1024;
constexpr uint32_t S_COL0 = 0, S_COL1 = 128, O_COL = 256;
constexpr int THREADS = 320;
}  // namespace v2

__global__ void __launch_bounds__(v2::THREADS, 1)
attn_self_v2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                    const __grid_constant__ AttnV2Params p) {
  using namespace v2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [3]
  uint64_t* v_full = bars + 4;        // [3]
  uint64_t* kv_empty = bars + 7;      // [3]
  uint64_t* s_full = bars + 10;       // [2]
  uint64_t* p_full = bars + 12;       // [2]
  uint64_t* pv_done = bars + 14;      // 1, one phase per key tile
  uint64_t* o_full = bars + 15;       // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* red_max = reinterpret_cast<float*>(smem + OFF_RED);   // [parity][half][row]
  float* red_l = red_max + 2 * 2 * 128;                        // [half][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int b_qk = p.qk_src[b];
  const int nt = p.n_k_tiles;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
    mbar_init(q_full, 1);
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 256); }
    mbar_init(pv_done, 1); mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_TILE);
      tma_load_4d(smem + OFF_Q, &tm_q, q_full, 0, h, q0, b_qk);
      for (int j = 0; j < nt; ++j) {
        const int st = j % NSTAGE;
        mbar_wait(&kv_empty[st], ((j / NSTAGE) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], KV_TILE);
        tma_load_4d(smem + OFF_K + st * KV_TILE, &tm_k, &k_full[st], 0, h, j * KT, b_qk);
        mbar_expect_tx(&v_full[st], KV_TILE);
        tma_load_4d(smem + OFF_V + st * KV_TILE, &tm_v, &v_full[st], 0, h, j * KT, b);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC_QK = umma_idesc_f16(128, KT, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_f16(128, 64, 0, 1);
      const uint32_t smem_base = smem_u32(smem);
      auto issue_qk = [&](int j) {
        const int st = j % NSTAGE;
        mbar_wait(&k_full[st], (j / NSTAGE) & 1);
        tc_fence_after();
        const uint32_t s_col = (j & 1) ? S_COL1 : S_COL0;
        for (int kk = 0; kk < p.ksteps_qk; ++kk) {
          const uint64_t da = umma_desc_sw128(smem_base + OFF_Q + kk * 32, 0, 1024);
          const uint64_t db = umma_desc_sw128(smem_base + OFF_K + st * KV_TILE + kk * 32, 0, 1024);
          mma_f16_ss(tmem + s_col, da, db, IDESC_QK, kk > 0);
        }
        tc_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < nt; ++j) {
        // S[(j+1)&1] held P_{j-1}; PV_{j-1} was issued in the previous iteration and the pipe is in-order
        if (j + 1 < nt) issue_qk(j + 1);
        const int st = j % NSTAGE;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        mbar_wait(&v_full[st], (j / NSTAGE) & 1);
        tc_fence_after();
        const uint32_t p_col = (j & 1) ? S_COL1 : S_COL0;
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
          const uint64_t db = umma_desc_sw128(smem_base + OFF_V + st * KV_TILE + kk * 2048, KV_TILE, 1024);
          mma_f16_ts(tmem + O_COL, tmem + p_col + kk * 8, db, IDESC_PV, (j > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&kv_empty[st]);
        tc_commit(pv_done);
        if (j == nt - 1) tc_commit(o_full);
      }
    }
  } else {
    // ------------------------------------------------------------- softmax + epilogue: 2 threads per query row
    const int half = warp >> 2;                       // keys [64*half, 64*half + 64) of every tile
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const bool row_ok = (q0 + row) < p.n_q;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < nt; ++j) {
      const uint32_t s_col = ((j & 1) ? S_COL1 : S_COL0);
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      float s[64];
      tmem_ld32(tlane + s_col + 64 * half, reinterpret_cast<uint32_t*>(s));
      tmem_ld32(tlane + s_col + 64 * half + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s));
      tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s) + 32);
      const int valid = p.n_k - j * KT - 64 * half;   // valid keys among my 64
      if (valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) s[i] = -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < 64; ++i) mx = fmaxf(mx, s[i]);
      float* rm = red_max + (j & 1) * 256;
      rm[half * 128 + row] = mx;
      asm volatile("bar.sync 1, 256;\n" ::: "memory");   // also: every thread has finished reading S_j
      mx = fmaxf(mx, rm[(half ^ 1) * 128 + row]);
      const float mxs = mx * p.scale_log2;
      if (j == 0) {
        m_ref = mxs;
      } else {
        const bool need = mxs > m_ref + 8.f;            // identical for both threads of a row
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(pv_done, (j - 1) & 1);               // O is being accumulated by PV_{j-1}
          tc_fence_after();
          const float alpha = need ? ex2_approx(m_ref - mxs) : 1.f;
          if (need) m_ref = mxs;
          l *= alpha;
          uint32_t o[32];
          tmem_ld32(tlane + O_COL + 32 * half, o);
          tmem_wait_ld_regs32(o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tlane + O_COL + 32 * half, o);
        }
      }
      float rowsum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e0 = ex2_approx(fmaf(s[2 * i], p.scale_log2, -m_ref));
        const float e1 = ex2_approx(fmaf(s[2 * i + 1], p.scale_log2, -m_ref));
        rowsum += e0 + e1;
        pk[i] = pack_half2(e0, e1);
      }
      l += rowsum;
      tmem_st32(tlane + s_col + 32 * half, pk);          // packed P_j over S_j: keys 64*half.. -> columns 32*half..
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[j & 1]);
    }
    // ---- epilogue
    red_l[half * 128 + row] = l;
    mbar_wait(o_full, 0);
    tc_fence_after();
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    const float l_tot = l + red_l[(half ^ 1) * 128 + row];
    const float inv_l = 1.f / l_tot;
    {
      uint8_t* otile = smem + OFF_O + row * 128;
      uint32_t o[32];
      tmem_ld32(tlane + O_COL + 32 * half, o);
      tmem_wait_ld_regs32(o);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 w;
        w.x = pack_half2(__uint_as_float(o[8 * v + 0]) * inv_l, __uint_as_float(o[8 * v + 1]) * inv_l);
        w.y = pack_half2(__uint_as_float(o[8 * v + 2]) * inv_l, __uint_as_float(o[8 * v + 3]) * inv_l);
        w.z = pack_half2(__uint_as_float(o[8 * v + 4]) * inv_l, __uint_as_float(o[8 * v + 5]) * inv_l);
        w.w = pack_half2(__uint_as_float(o[8 * v + 6]) * inv_l, __uint_as_float(o[8 * v + 7]) * inv_l);
        const int chunk = half * 4 + v;
        *reinterpret_cast<uint4*>(otile + ((chunk ^ (row & 7)) << 4)) = w;
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    if (threadIdx.x == 0) {
      tma_store_4d(&tm_o, smem + OFF_O, 0, h, q0, b);
      tma_store_commit();
      tma_store_wait_all();
    }
    if (p.lse != nullptr && row_ok && half == 0)
      p.lse[(static_cast<size_t>(b) * p.heads + h) * p.n_q + q0 + row] = m_ref + log2f(l_tot);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

int launch_attn_self_v2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                        int batch, int heads, int head_dim, int n_q, int n_k, float scale_log2, const int8_t* qk_src,
                        float* lse, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_self_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM_BYTES) != cudaSuccess)
      return RTTI_ERR_CUDA;
    configured = true;
  }
  AttnV2Params p{};
  p.batch = batch; p.heads = heads; p.head_dim = head_dim; p.n_q = n_q; p.n_k = n_k;
  p.n_k_tiles = (n_k + v2::KT - 1) / v2::KT;
  p.ksteps_qk = (head_dim + 15) / 16;
  p.scale_log2 = scale_log2;
  for (int i = 0; i < 64; ++i) p.qk_src[i] = qk_src[i];
  p.lse = lse;
  dim3 grid((n_q + 127) / 128, heads, batch);
  attn_self_v2_kernel<<<grid, v2::THREADS, v2::SMEM_BYTES, stream>>>(tq, tk, tv, to, p);
  return cudaGetLastError() == cudaSuccess ? RTTI_OK : RTTI_ERR_CUDA;
}

}  // namespace rtti

// =====================================================================================================
// v3: the same MMA look-ahead with 64-key tiles, ONE thread per query row and TWO CTAs per SM.
//
// v2 hides the MMAs but its 8 softmax warps run in lock-step (the two threads of a row exchange their max
// through a named barrier every tile), so on every SM sub-partition both resident warps are in the MUFU burst
// or outside it at the same time and the MUFU pipe idles ~45 % of the time. v3 removes the exchange (a thread
// owns a whole row of a 64-key tile) and lets two independent CTAs per SM de-phase naturally: while one CTA's
// warp on a sub-partition loads S / takes the max / stores P, the other CTA's warp keeps the MUFU pipe busy.
// TMEM per CTA: S0 [0,64), S1 [64,128), O [128,192) -> 256-column allocation, two CTAs fill the 512 columns.
namespace rtti {
constexpr int V3_POLY_DEFAULT = 0;
constexpr bool V3_ILP_DEFAULT = false;
constexpr bool V3_PF_DEFAULT = false;
namespace v3 {
constexpr int KT = 64;
constexpr int NSTAGE = 4;
constexpr int Q_TILE = 128 * 128;
constexpr int KV_TILE = KT * 128;   // 8 KB
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_TILE;
constexpr int OFF_V = OFF_K + NSTAGE * KV_TILE;
constexpr int OFF_O = OFF_V + NSTAGE * KV_TILE;
constexpr int OFF_BAR = OFF_O + Q_TILE;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;   // ~98 KB -> 2 CTAs / SM
constexpr uint32_t O_COL = 128;
constexpr int THREADS = 192;

// 2^x for x <= 0 on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], cubic
// minimax polynomial for 2^f (max relative error 7.5e-5 = 2^-13.7, six times below the fp16 rounding of P), and
// n added into the exponent field. Inputs below -126 (masked keys: -inf) are clamped -> 2^-126, which is 0 in fp16.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;                 // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float f = x - (t - 12582912.f);
  float pl = fmaf(0.0551716685f, f, 0.2426111251f);
  pl = fmaf(pl, f, 0.6932609677f);
  pl = fmaf(pl, f, 0.9999280572f);
  return __uint_as_float(__float_as_uint(pl) + (__float_as_uint(t) << 23));
}
}  // namespace v3

// POLY: every POLY-th exponential of a row is evaluated with v3::exp2_poly instead of ex2.approx (0 = none). The
// softmax warps are bound by the 16-lane/SM MUFU pipe; moving a fraction of the exponentials to the FMA pipe
// (FlashAttention-4's trick) shortens the MUFU burst of every tile.
// ILP: row max and row sum are reduced as four independent chains (the straight loops compile to 31-deep FMNMX3 and
// 32-deep FADD dependency chains; a softmax warp is latency-bound, not throughput-bound — see DESIGN.md §3.1).
// PF: software-pipelined softmax — the TMEM load of S_{j+1} is issued before the exponentials of tile j and lands
// half-way through them, and its row max is computed in the same basic block as the second half of the exponentials,
// so the load latency and the max chain of the next tile hide under the MUFU work of the current one.
// X2: scale/shift and row sums with the packed fp32 instructions of sm_100 (fma.rn.f32x2 / add.rn.f32x2 -> FFMA2 / FADD2):
// 64 fewer issue slots per tile in a warp whose every instruction costs about its issue time. Same products and
// roundings per element; only the order of the row-sum additions differs.
template <int POLY, bool ILP, bool PF, bool X2 = false>
__global__ void __launch_bounds__(v3::THREADS, 2)
attn_self_v3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                    const __grid_constant__ AttnV2Params p) {
  using namespace v3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [4]
  uint64_t* v_full = bars + 5;        // [4]
  uint64_t* kv_empty = bars + 9;      // [4]
  uint64_t* s_full = bars + 13;       // [2]
  uint64_t* p_full = bars + 15;       // [2]
  uint64_t* pv_done = bars + 17;
  uint64_t* o_full = bars + 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int b_qk = p.qk_src[b];
  const int nt = p.n_k_tiles;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
    mbar_init(q_full, 1);
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); }
    mbar_init(pv_done, 1); mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_TILE);
      tma_load_4d(smem + OFF_Q, &tm_q, q_full, 0, h, q0, b_qk);
      for (int j = 0; j < nt; ++j) {
        const int st = j % NSTAGE;
        mbar_wait(&kv_empty[st], ((j / NSTAGE) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], KV_TILE);
        tma_load_4d(smem + OFF_K + st * KV_TILE, &tm_k, &k_full[st], 0, h, j * KT, b_qk);
        mbar_expect_tx(&v_full[st], KV_TILE);
        tma_load_4d(smem + OFF_V + st * KV_TILE, &tm_v, &v_full[st], 0, h, j * KT, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t IDESC_QK = umma_idesc_f16(128, KT, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_f16(128, 64, 0, 1);
      const uint32_t smem_base = smem_u32(smem);
      auto issue_qk = [&](int j) {
        const int st = j % NSTAGE;
        mbar_wait(&k_full[st], (j / NSTAGE) & 1);
        tc_fence_after();
        for (int kk = 0; kk < p.ksteps_qk; ++kk) {
          const uint64_t da = umma_desc_sw128(smem_base + OFF_Q + kk * 32, 0, 1024);
          const uint64_t db = umma_desc_sw128(smem_base + OFF_K + st * KV_TILE + kk * 32, 0, 1024);
          mma_f16_ss(tmem + 64u * (j & 1), da, db, IDESC_QK, kk > 0);
        }
        tc_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_qk(j + 1);   // S[(j+1)&1] held P_{j-1}; PV_{j-1} precedes it in the in-order tensor pipe
        const int st = j % NSTAGE;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        mbar_wait(&v_full[st], (j / NSTAGE) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
          const uint64_t db = umma_desc_sw128(smem_base + OFF_V + st * KV_TILE + kk * 2048, KV_TILE, 1024);
          mma_f16_ts(tmem + O_COL, tmem + 64u * (j & 1) + kk * 8, db, IDESC_PV, (j > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&kv_empty[st]);
        tc_commit(pv_done);
        if (j == nt - 1) tc_commit(o_full);
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const bool row_ok = (q0 + row) < p.n_q;
    float m_ref = -INFINITY, l = 0.f;
    if constexpr (PF) {
      float sA[64], sB[64];
      auto row_max = [&](const float (&s)[64]) {
        float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
        for (int i = 4; i < KT; i += 4) {
#pragma unroll
          for (int c = 0; c < 4; ++c) m4[c] = fmaxf(m4[c], s[i + c]);
        }
        return fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      };
      auto load = [&](int j, float (&s)[64]) {          // asynchronous: the registers are not valid until land()
        mbar_wait(&s_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        tmem_ld32(tlane + 64u * (j & 1), reinterpret_cast<uint32_t*>(s));
        tmem_ld32(tlane + 64u * (j & 1) + 32, reinterpret_cast<uint32_t*>(s) + 32);
      };
      auto land = [&](int j, float (&s)[64]) {
        tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s));
        tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s) + 32);
        const int valid = p.n_k - j * KT;
        if (valid < KT) {
#pragma unroll
          for (int i = 0; i < KT; ++i)
            if (i >= valid) s[i] = -INFINITY;
        }
      };
      auto tile = [&](int j, float (&cur)[64], float (&nxt)[64], float mx) -> float {
        const uint32_t s_col = 64u * (j & 1);
        const float mxs = mx * p.scale_log2;
        if (j == 0) {
          m_ref = mxs;
        } else {
          const bool need = mxs > m_ref + 8.f;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(pv_done, (j - 1) & 1);   // O is being accumulated by PV_{j-1}
            tc_fence_after();
            const float alpha = need ? ex2_approx(m_ref - mxs) : 1.f;
            if (need) m_ref = mxs;
            l *= alpha;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[16];
              tmem_ld16(tlane + O_COL + 16 * c, o);
              tmem_wait_ld_regs16(o);
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(tlane + O_COL + 16 * c, o);
            }
          }
        }
        const bool has_next = j + 1 < nt;
        if (has_next) load(j + 1, nxt);
        float rs4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e0 = ex2_approx(fmaf(cur[2 * i], p.scale_log2, -m_ref));
          const float e1 = ex2_approx(fmaf(cur[2 * i + 1], p.scale_log2, -m_ref));
          rs4[i & 3] += e0 + e1;
          pk[i] = pack_half2(e0, e1);
        }
        if (has_next) land(j + 1, nxt);
        const float mx_next = row_max(nxt);    // unconditional: same basic block as the exponentials below (unused if !has_next)
#pragma unroll
        for (int i = 16; i < 32; ++i) {
          const float e0 = ex2_approx(fmaf(cur[2 * i], p.scale_log2, -m_ref));
          const float e1 = ex2_approx(fmaf(cur[2 * i + 1], p.scale_log2, -m_ref));
          rs4[i & 3] += e0 + e1;
          pk[i] = pack_half2(e0, e1);
        }
        l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
        tmem_st32(tlane + s_col, pk);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[j & 1]);
        return mx_next;
      };
#pragma unroll
      for (int i = 0; i < 64; ++i) sB[i] = 0.f;
      load(0, sA);
      land(0, sA);
      float mx = row_max(sA);
      for (int j = 0; j < nt; j += 2) {
        mx = tile(j, sA, sB, mx);
        if (j + 1 < nt) mx = tile(j + 1, sB, sA, mx);
      }
    } else {
      for (int j = 0; j < nt; ++j) {
        const uint32_t s_col = 64u * (j & 1);
        mbar_wait(&s_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        float s[64];
        tmem_ld32(tlane + s_col, reinterpret_cast<uint32_t*>(s));
        tmem_ld32(tlane + s_col + 32, reinterpret_cast<uint32_t*>(s) + 32);
        tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s));
        tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s) + 32);
        const int valid = p.n_k - j * KT;
        if (valid < KT) {
  #pragma unroll
          for (int i = 0; i < KT; ++i)
            if (i >= valid) s[i] = -INFINITY;
        }
        float mx;
        if (ILP) {
          float m4[4] = {s[0], s[1], s[2], s[3]};
  #pragma unroll
          for (int i = 4; i < KT; i += 4) {
  #pragma unroll
            for (int c = 0; c < 4; ++c) m4[c] = fmaxf(m4[c], s[i + c]);
          }
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        } else {
          mx = s[0];
  #pragma unroll
          for (int i = 1; i < KT; ++i) mx = fmaxf(mx, s[i]);
        }
        const float mxs = mx * p.scale_log2;
        if (j == 0) {
          m_ref = mxs;
        } else {
          const bool need = mxs > m_ref + 8.f;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(pv_done, (j - 1) & 1);   // O is being accumulated by PV_{j-1}
            tc_fence_after();
            const float alpha = need ? ex2_approx(m_ref - mxs) : 1.f;
            if (need) m_ref = mxs;
            l *= alpha;
  #pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[16];
              tmem_ld16(tlane + O_COL + 16 * c, o);
              tmem_wait_ld_regs16(o);
  #pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(tlane + O_COL + 16 * c, o);
            }
          }
        }
        uint32_t pk[32];
        if constexpr (X2) {
          const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_ref, -m_ref);
          float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float2 x = fma_f32x2(make_float2(s[2 * i], s[2 * i + 1]), sc2, nm2);
            const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
            acc[i & 1] = add_f32x2(acc[i & 1], e);
            pk[i] = pack_half2(e.x, e.y);
          }
          l += (acc[0].x + acc[0].y) + (acc[1].x + acc[1].y);
        } else {
          float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x0 = fmaf(s[2 * i], p.scale_log2, -m_ref);
            const float x1 = fmaf(s[2 * i + 1], p.scale_log2, -m_ref);
            const float e0 = (POLY > 0 && (2 * i) % POLY == POLY - 1) ? exp2_poly(x0) : ex2_approx(x0);
            const float e1 = (POLY > 0 && (2 * i + 1) % POLY == POLY - 1) ? exp2_poly(x1) : ex2_approx(x1);
            rs4[ILP ? (i & 3) : 0] += e0 + e1;
            pk[i] = pack_half2(e0, e1);
          }
          l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
        }
        tmem_st32(tlane + s_col, pk);          // packed P_j over the first 32 columns of its own S buffer
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[j & 1]);
      }
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.f / l;
    uint8_t* otile = smem + OFF_O + row * 128;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t o[32];
      tmem_ld32(tlane + O_COL + 32 * hh, o);
      tmem_wait_ld_regs32(o);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 w;
        w.x = pack_half2(__uint_as_float(o[8 * v + 0]) * inv_l, __uint_as_float(o[8 * v + 1]) * inv_l);
        w.y = pack_half2(__uint_as_float(o[8 * v + 2]) * inv_l, __uint_as_float(o[8 * v + 3]) * inv_l);
        w.z = pack_half2(__uint_as_float(o[8 * v + 4]) * inv_l, __uint_as_float(o[8 * v + 5]) * inv_l);
        w.w = pack_half2(__uint_as_float(o[8 * v + 6]) * inv_l, __uint_as_float(o[8 * v + 7]) * inv_l);
        const int chunk = hh * 4 + v;
        *reinterpret_cast<uint4*>(otile + ((chunk ^ (row & 7)) << 4)) = w;
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;\n" ::: "memory");
    if (threadIdx.x == 0) {
      tma_store_4d(&tm_o, smem + OFF_O, 0, h, q0, b);
      tma_store_commit();
      tma_store_wait_all();
    }
    if (p.lse != nullptr && row_ok)
      p.lse[(static_cast<size_t>(b) * p.heads + h) * p.n_q + q0 + row] = m_ref + log2f(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<256>(tmem);
}

int launch_attn_self_v3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                        int batch, int heads, int head_dim, int n_q, int n_k, float scale_log2, const int8_t* qk_src,
                        float* lse, cudaStream_t stream) {
  // RTTI_ATTN_POLY = 0 | 4 | 8: fraction 1/k of the exponentials on the FMA pipe; RTTI_ATTN_ILP / RTTI_ATTN_PF = 0 | 1 (read once)
  static const int poly = [] { const char* e = getenv("RTTI_ATTN_POLY"); return e ? atoi(e) : V3_POLY_DEFAULT; }();
  static const bool ilp = [] { const char* e = getenv("RTTI_ATTN_ILP"); return e ? atoi(e) != 0 : V3_ILP_DEFAULT; }();
  using KernelT = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnV2Params);
  static const bool pf = [] { const char* e = getenv("RTTI_ATTN_PF"); return e ? atoi(e) != 0 : V3_PF_DEFAULT; }();
  static const bool x2 = [] { const char* e = getenv("RTTI_ATTN_X2"); return e ? atoi(e) != 0 : false; }();   // unverified on hardware
  static const KernelT kernel = [] {
    KernelT k = ilp ? attn_self_v3_kernel<0, true, false> : attn_self_v3_kernel<0, false, false>;
    if (poly == 4) k = ilp ? attn_self_v3_kernel<4, true, false> : attn_self_v3_kernel<4, false, false>;
    if (poly == 8) k = ilp ? attn_self_v3_kernel<8, true, false> : attn_self_v3_kernel<8, false, false>;
    if (pf) k = attn_self_v3_kernel<0, true, true>;
    if (x2) k = attn_self_v3_kernel<0, false, false, true>;
    return k;
  }();
  static const bool configured =
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v3::SMEM_BYTES) == cudaSuccess;
  if (!configured) return RTTI_ERR_CUDA;
  AttnV2Params p{};
  p.batch = batch; p.heads = heads; p.head_dim = head_dim; p.n_q = n_q; p.n_k = n_k;
  p.n_k_tiles = (n_k + v3::KT - 1) / v3::KT;
  p.ksteps_qk = (head_dim + 15) / 16;
  p.scale_log2 = scale_log2;
  for (int i = 0; i < 64; ++i) p.qk_src[i] = qk_src[i];
  p.lse = lse;
  dim3 grid((n_q + 127) / 128, heads, batch);
  kernel<<<grid, v3::THREADS, v3::SMEM_BYTES, stream>>>(tq, tk, tv, to, p);
  return cudaGetLastError() == cudaSuccess ? RTTI_OK : RTTI_ERR_CUDA;
}
}  // namespace rtti

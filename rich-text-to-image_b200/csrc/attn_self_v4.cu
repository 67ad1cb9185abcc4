// EXPERIMENTAL (round-2 candidate, NOT verified on hardware yet, never selected unless RTTI_ATTN_V4=1):
// self-attention for head_dim <= 64 with THREE score buffers in TMEM and a quarter-tile software pipeline.
//
// Why: in v3 a softmax warp is a serial chain per 64-key tile — wait S, tcgen05.ld of 8 KB
// (32 lanes x 64 fp32 columns; the TMEM read port moves ~16 B/cycle per SM sub-partition, i.e. ~512 cycles), row max,
// 64 exponentials (512 cycles of the 4-lane-per-sub-partition MUFU pipe), P store, arrive. The two pipes that matter,
// TMEM read and MUFU, are used one after the other, so the two warps that share a sub-partition reach ~65 % of either.
// Prefetching S_{j+1} under the exponentials of tile j failed with v3's look-ahead of one tile (S_{j+1} is not ready
// that early: RTTI_ATTN_PF=1, 456 vs 690 TFLOP/s) and spilled registers.
//
// v4 changes two things:
//   * S is triple-buffered (3 x 64 columns + 64 for O = the same 256-column allocation, still 2 CTAs/SM); the MMA warp
//     runs QK^T two tiles ahead, so S_{j+1} has long landed when softmax_j starts;
//   * the softmax loop works in quarters of 16 columns: the tcgen05.ld of quarter q of S_{j+1} is issued into the
//     registers that quarter q-1 of S_j just vacated, lands while quarter q of S_j is exponentiated, and its partial row
//     max is taken in the same basic block as the next quarter's exponentials. Live registers: 64 + 16 + packed P.
// Everything else (TMA producer, in-place fp16 P, TS-operand PV MMA, lazy rescale, epilogue) is v3's.
//
// Scheduling note (checked with cuobjdump -sass: order of LDTM vs MUFU.EX2): ptxas is free to re-order the PTX stream and,
// with a straight-line tile body, sinks the quarter loads behind the 64 exponentials (`asm volatile` only orders the
// statements for the front end). Each quarter's exponentials therefore sit in their own basic block behind an
// always-true-but-opaque branch; the SASS of both unrolled tile bodies is then LDTM, 16 MUFU, LDTM, 16 MUFU, ... as
// intended. Cost to look at on hardware: 168 registers with 72 bytes of spill stores per thread.
#include <cstdlib>
#include <type_traits>

#include "ptx.cuh"
#include "rtti_internal.h"

namespace rtti {

struct AttnV4Params {
  int batch, heads, head_dim, n_q, n_k, n_k_tiles, ksteps_qk;
  float scale_log2;
  int8_t qk_src[64];
  float* lse;
};

namespace v4 {
constexpr int KT = 64;
constexpr int NSTAGE = 4;
constexpr int NSBUF = 3;
constexpr int Q_TILE = 128 * 128;
constexpr int KV_TILE = KT * 128;   // 8 KB
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_TILE;
constexpr int OFF_V = OFF_K + NSTAGE * KV_TILE;
constexpr int OFF_O = OFF_V + NSTAGE * KV_TILE;
constexpr int OFF_BAR = OFF_O + Q_TILE;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;   // ~98 KB -> 2 CTAs / SM
constexpr uint32_t O_COL = 64 * NSBUF;             // 192
constexpr int THREADS = 192;

// volatile: keeps the exponentials in program order relative to the (volatile) tcgen05.ld / wait::ld statements — the
// scheduler otherwise hoists all 64 MUFU ops of a tile above the quarter loads and the pipelining is gone (seen in SASS).
__device__ __forceinline__ float ex2_ordered(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace v4

__global__ void __launch_bounds__(v4::THREADS, 2)
attn_self_v4_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                    const __grid_constant__ AttnV4Params p) {
  using namespace v4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [4]
  uint64_t* v_full = bars + 5;        // [4]
  uint64_t* kv_empty = bars + 9;      // [4]
  uint64_t* s_full = bars + 13;       // [3]
  uint64_t* p_full = bars + 16;       // [3]
  uint64_t* pv_done = bars + 19;
  uint64_t* o_full = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

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
    for (int i = 0; i < NSBUF; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); }
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
          mma_f16_ss(tmem + 64u * (j % NSBUF), da, db, IDESC_QK, kk > 0);
        }
        tc_commit(&s_full[j % NSBUF]);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_qk(0);
      if (nt > 1) issue_qk(1);
      if (nt > 2) issue_qk(2);
      for (int j = 0; j < nt; ++j) {
        const int st = j % NSTAGE;
        const int sb = j % NSBUF;
        mbar_wait(&p_full[sb], (j / NSBUF) & 1);
        tc_fence_after();
        mbar_wait(&v_full[st], (j / NSTAGE) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
          const uint64_t db = umma_desc_sw128(smem_base + OFF_V + st * KV_TILE + kk * 2048, KV_TILE, 1024);
          mma_f16_ts(tmem + O_COL, tmem + 64u * sb + kk * 8, db, IDESC_PV, (j > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&kv_empty[st]);
        tc_commit(pv_done);
        if (j == nt - 1) tc_commit(o_full);
        // S buffer sb held P_j: QK^T_{j+3} may overwrite it once PV_j has read it — the tensor pipe is in order
        if (j + NSBUF < nt) issue_qk(j + NSBUF);
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const bool row_ok = (q0 + row) < p.n_q;
    float m_ref = -INFINITY, l = 0.f;
    float sA[64], sB[64];
    // always true, but not provably so: every quarter's exponentials become their own basic block, which keeps ptxas from
    // sinking the tcgen05.ld of the next quarter behind them (see the header)
    const bool opaque = p.batch != -0x5eed;

    auto mask_q = [&](int j, int q, float (&s)[64]) {   // keys beyond n_k in the last tile
      const int valid = p.n_k - j * KT;
      if (valid < KT) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (16 * q + i >= valid) s[16 * q + i] = -INFINITY;
      }
    };
    auto max_q = [&](int q, const float (&s)[64]) {     // two chains of 8
      float a = s[16 * q], c = s[16 * q + 1];
#pragma unroll
      for (int i = 2; i < 16; i += 2) { a = fmaxf(a, s[16 * q + i]); c = fmaxf(c, s[16 * q + i + 1]); }
      return fmaxf(a, c);
    };
    // packed fp32 scale/shift and row sums (FFMA2 / FADD2): half the issue slots of the scalar form
    auto exp_q = [&](int q, const float (&s)[64], uint32_t (&pk)[32], float2 (&rs)[2]) {
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_ref, -m_ref);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 x = fma_f32x2(make_float2(s[16 * q + 2 * i], s[16 * q + 2 * i + 1]), sc2, nm2);
        const float2 e = make_float2(ex2_ordered(x.x), ex2_ordered(x.y));
        rs[i & 1] = add_f32x2(rs[i & 1], e);
        pk[8 * q + i] = pack_half2(e.x, e.y);
      }
    };
    // one tile: exponentiate S_j held in `cur` (row max mx), meanwhile bring S_{j+1} into `nxt` quarter by quarter
    // HAS_NEXT is a compile-time tag (std::true_type / std::false_type): no branches between the quarter steps, so the
    // partial max of S_{j+1} and the next quarter's exponentials are one basic block for the instruction scheduler.
    auto tile = [&](auto has_next_tag, int j, float (&cur)[64], float (&nxt)[64], float mx) -> float {
      constexpr bool has_next = decltype(has_next_tag)::value;
      const uint32_t s_col = 64u * (j % NSBUF);
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
      const uint32_t n_col = 64u * ((j + 1) % NSBUF);
      uint32_t pk[32];
      float2 rs[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      float mq0 = -INFINITY, mq1 = -INFINITY, mq2 = -INFINITY, mq3 = -INFINITY;
      if constexpr (has_next) {
        mbar_wait(&s_full[(j + 1) % NSBUF], ((j + 1) / NSBUF) & 1);   // issued two tiles ago: normally no wait
        tc_fence_after();
        tmem_ld16(tlane + n_col, reinterpret_cast<uint32_t*>(nxt));
      }
      if (opaque) exp_q(0, cur, pk, rs);
      if constexpr (has_next) {
        tmem_wait_ld_regs16(reinterpret_cast<uint32_t*>(nxt));
        mask_q(j + 1, 0, nxt);
        tmem_ld16(tlane + n_col + 16, reinterpret_cast<uint32_t*>(nxt) + 16);
        mq0 = max_q(0, nxt);
      }
      if (opaque) exp_q(1, cur, pk, rs);
      if constexpr (has_next) {
        tmem_wait_ld_regs16(reinterpret_cast<uint32_t*>(nxt) + 16);
        mask_q(j + 1, 1, nxt);
        tmem_ld16(tlane + n_col + 32, reinterpret_cast<uint32_t*>(nxt) + 32);
        mq1 = max_q(1, nxt);
      }
      tmem_st16(tlane + s_col, pk);          // first half of packed P_j (S_j is entirely in registers by now)
      if (opaque) exp_q(2, cur, pk, rs);
      if constexpr (has_next) {
        tmem_wait_ld_regs16(reinterpret_cast<uint32_t*>(nxt) + 32);
        mask_q(j + 1, 2, nxt);
        tmem_ld16(tlane + n_col + 48, reinterpret_cast<uint32_t*>(nxt) + 48);
        mq2 = max_q(2, nxt);
      }
      if (opaque) exp_q(3, cur, pk, rs);
      l += (rs[0].x + rs[0].y) + (rs[1].x + rs[1].y);
      tmem_st16(tlane + s_col + 16, pk + 16);   // second half: P_j occupies the first 32 columns of its own S buffer
      if constexpr (has_next) {
        tmem_wait_ld_regs16(reinterpret_cast<uint32_t*>(nxt) + 48);
        mask_q(j + 1, 3, nxt);
        mq3 = max_q(3, nxt);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[j % NSBUF]);
      return fmaxf(fmaxf(mq0, mq1), fmaxf(mq2, mq3));
    };

    // prologue: S_0 in full
    mbar_wait(&s_full[0], 0);
    tc_fence_after();
#pragma unroll
    for (int q = 0; q < 4; ++q) tmem_ld16(tlane + 16 * q, reinterpret_cast<uint32_t*>(sA) + 16 * q);
#pragma unroll
    for (int q = 0; q < 4; ++q) tmem_wait_ld_regs16(reinterpret_cast<uint32_t*>(sA) + 16 * q);
    float mx = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) { mask_q(0, q, sA); mx = fmaxf(mx, max_q(q, sA)); }
    {
      constexpr std::true_type T{};
      constexpr std::false_type F{};
      int j = 0;
      for (; j + 2 < nt; j += 2) {
        mx = tile(T, j, sA, sB, mx);
        mx = tile(T, j + 1, sB, sA, mx);
      }
      if (j + 1 < nt) {            // two tiles left
        mx = tile(T, j, sA, sB, mx);
        tile(F, j + 1, sB, sA, mx);
      } else {                     // one tile left
        tile(F, j, sA, sB, mx);
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

int launch_attn_self_v4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                        int batch, int heads, int head_dim, int n_q, int n_k, float scale_log2, const int8_t* qk_src,
                        float* lse, cudaStream_t stream) {
  static const bool configured =
      cudaFuncSetAttribute(attn_self_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v4::SMEM_BYTES) == cudaSuccess;
  if (!configured) return RTTI_ERR_CUDA;
  AttnV4Params p{};
  p.batch = batch; p.heads = heads; p.head_dim = head_dim; p.n_q = n_q; p.n_k = n_k;
  p.n_k_tiles = (n_k + v4::KT - 1) / v4::KT;
  p.ksteps_qk = (head_dim + 15) / 16;
  p.scale_log2 = scale_log2;
  for (int i = 0; i < 64; ++i) p.qk_src[i] = qk_src[i];
  p.lse = lse;
  dim3 grid((n_q + 127) / 128, heads, batch);
  attn_self_v4_kernel<<<grid, v4::THREADS, v4::SMEM_BYTES, stream>>>(tq, tk, tv, to, p);
  return cudaGetLastError() == cudaSuccess ? RTTI_OK : RTTI_ERR_CUDA;
}

}  // namespace rtti

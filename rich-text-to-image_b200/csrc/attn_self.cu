// Self-attention forward for head_dim <= 64 on sm_100a: tcgen05 MMAs with TMEM accumulators, TMA-fed shared memory,
// and the region sampler's self-attention INJECTION fused in as "grouped PV".
//
// Reference semantics (models/attention_processor.py:1157-1163): a pass either computes P = softmax(scale Q K^T) or is
// handed `real_attn_probs` = the P of the reference pass (hooks models/region_diffusion_sdxl.py:1018-1029, 1064-1082)
// and only evaluates P @ V. All passes of a denoising step run in one batch here, so the entries that share a score
// source form a GROUP: one CTA computes S and the softmax ONCE per (query tile, head, group) and issues P @ V_b for every
// member b of the group (NV accumulators side by side in TMEM). For the 5-region SDXL step that is 1 + 5 MMAs per key
// tile instead of 5 x (1 + 1), and — more important at head_dim 64 — one row of exponentials instead of five: the MUFU
// pipe (16 ex2/clk/SM: 512 cycles per 128x64 tile against 256 tensor cycles) is what bounds the plain kernel.
//
// Schedule of one CTA (128 query rows of one head; 5 warps: 0-3 softmax/epilogue, 4 = one control thread that issues
// both the TMA loads and the MMAs):
//   TMEM   S  fp32 scores of ONE 64-key tile (64 columns), P packed-fp16 probabilities (32 columns per buffer),
//          O_v fp32 accumulators (64 columns per group member).
//            plain  (NV = 1): two allocations, 128 (S, O) + 32 (P) columns -> THREE CTAs per SM (480 of 512 columns)
//            groups (NV <= 6): one 512-column allocation: S | P0 | P1 | O_0..O_5 -> one CTA per SM
//   smem   Q tile 16 KB (later the output staging tile) | K ring KST x 8 KB | V ring KST x NV x 8 KB
//   * a softmax thread owns one query row: tcgen05.ld of its 64 scores, then IMMEDIATELY releases S (`s_free`), so the
//     control thread issues Q K_{j+1}^T while the exponentials of tile j are still being evaluated (register double
//     buffering: the look-ahead of a second S buffer without its TMEM columns);
//   * P_j is written back to TMEM (tcgen05.st) and feeds P V as a TS MMA (A operand from TMEM): shared-memory
//     bandwidth is spent on K and V only (a first version staged P in shared memory: 80 KB of smem traffic per tile,
//     slower; profiles/r02_kernels_selfattn_psmem_first.jsonl);
//   * groups double-buffer P, so the softmax of tile j+1 overlaps the NV P V_j MMAs (640 tensor cycles for 5 members);
//   * no "stage empty" barriers: the control thread learns that Q K_j^T has completed from `s_free` (the softmax warps
//     read S_j only after it) and that P V_{j-1} has completed from `p_full` (they write P_j only after it), so it
//     refills the K stage of tile j and the V stage of tile j-1 at those points;
//   * lazy rescale of O (only when a row max grows by more than 2^8) on all NV accumulators.
#include <stdlib.h>

#include "ptx.cuh"
#include "rtti_internal.h"

#ifndef SA_POLY_DEFAULT
#define SA_POLY_DEFAULT 4
#endif

namespace rtti {

struct AttnSelfParams {
  int heads, head_dim, n_q, n_k, n_k_tiles, ksteps_qk;
  float scale_log2;
  float* lse;            // [batch, heads, n_q] log2-domain log-sum-exp of the scaled scores (optional)
  int8_t qk[64];         // group g: batch entry supplying Q and K
  int8_t nv[64];         // group g: number of members (1..NV)
  int8_t vent[64][6];    // group g: batch entries supplying V / receiving O
};

namespace sa {
constexpr int KT = 64;
constexpr int Q_TILE = 128 * 128;   // bytes
constexpr int KV_TILE = KT * 128;   // 8 KB
template <int NV> struct Cfg {
  static constexpr int KST = 3;                                // K / V ring depth
  static constexpr int PBUF = NV == 1 ? 1 : 2;                 // P buffers in TMEM
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_TILE;
  static constexpr int OFF_V = OFF_K + KST * KV_TILE;
  static constexpr int SPLIT = NV == 1 ? 1 : 2;                // softmax threads per query row
  static constexpr int THREADS = 32 * (4 * SPLIT + 1);
  static constexpr int OFF_BAR = OFF_V + KST * NV * KV_TILE;
  static constexpr int OFF_RED = OFF_BAR + 128;                // groups: row-max / row-sum exchange, 3 x 2 x 128 floats
  static constexpr int SMEM_BYTES = OFF_RED + (SPLIT == 2 ? 3 * 2 * 128 * 4 : 0) + 1024;   // + alignment slack
  static constexpr uint32_t COLS_A = NV == 1 ? 128 : 512;      // S, (P), O
  static constexpr uint32_t COL_P = 64;                        // groups: P buffers at [64, 128)
  static constexpr uint32_t COL_O = NV == 1 ? 64 : 128;
  static constexpr int MAX_REGS = NV == 1 ? 136 : 224;         // 65536 / (160 threads x 3 CTAs) resp. / 288 threads, 8-register granules
};
// 2^x for x <= ~8 on the FMA / ALU pipes, two lanes at a time with the packed fp32 instructions of sm_100: round-to-nearest
// split x = n + f (magic-number add), cubic minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, below the
// fp16 rounding of P), n added into the exponent field. Inputs below -126 (masked keys: -inf) clamp to 2^-126 -> 0 in fp16.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f); x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);   // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float2 t = add_f32x2(x, magic);
  const float2 f = add_f32x2(x, make_float2(12582912.f - t.x, 12582912.f - t.y));
  float2 pl = fma_f32x2(make_float2(0.0551716685f, 0.0551716685f), f, make_float2(0.2426111251f, 0.2426111251f));
  pl = fma_f32x2(pl, f, make_float2(0.6932609677f, 0.6932609677f));
  pl = fma_f32x2(pl, f, make_float2(0.9999280572f, 0.9999280572f));
  return make_float2(__uint_as_float(__float_as_uint(pl.x) + (__float_as_uint(t.x) << 23)),
                     __uint_as_float(__float_as_uint(pl.y) + (__float_as_uint(t.y) << 23)));
}
}  // namespace sa

template <int NV, int POLY>
__global__ void __launch_bounds__(sa::Cfg<NV>::THREADS) __maxnreg__((sa::Cfg<NV>::MAX_REGS))
attn_self_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                 const __grid_constant__ AttnSelfParams p) {
  using namespace sa;
  using C = Cfg<NV>;
  constexpr int KST = C::KST, PBUF = C::PBUF, CW = 4 * C::SPLIT;   // CW: index of the control warp
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [3]
  uint64_t* v_full = bars + 4;        // [3]
  uint64_t* s_full = bars + 7;        // QK^T_j landed in TMEM
  uint64_t* s_free = bars + 8;        // all 128 rows of S_j are in registers
  uint64_t* p_full = bars + 9;        // [2] P_j is in TMEM buffer j % PBUF. One barrier PER BUFFER: with two buffers the
                                      // softmax warps may finish tile j+1 before the control warp has observed p_full(j)
                                      // (it can sit in a v_full wait, e.g. cold L2) — on a single barrier the phase would
                                      // flip twice and the parity wait would never return (seen as a hang after an L2 flush)
  uint64_t* pv_done = bars + 11;      // [2] P V_j has completed (P buffer j % PBUF reusable, O consistent)
  uint64_t* o_full = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);   // [2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int grp = blockIdx.z;
  const int b_qk = p.qk[grp];
  const int nv = NV == 1 ? 1 : p.nv[grp];
  const int nt = p.n_k_tiles;

  if (warp == CW) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
      mbar_init(q_full, 1);
      for (int i = 0; i < KST; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); }
      mbar_init(s_full, 1); mbar_init(s_free, 128 * C::SPLIT); mbar_init(&p_full[0], 128 * C::SPLIT); mbar_init(&p_full[1], 128 * C::SPLIT);
      mbar_init(&pv_done[0], 1); mbar_init(&pv_done[1], 1); mbar_init(o_full, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc_keep_permit<C::COLS_A>(&tmem_slot[0]);
    if (NV == 1) tmem_alloc_keep_permit<32>(&tmem_slot[1]);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot[0];
  const uint32_t tmem_p = NV == 1 ? tmem_slot[1] : tmem + C::COL_P;   // groups: buffer b at tmem_p + 32 b
  const uint32_t tmem_o = tmem + C::COL_O;

  if (warp == CW) {
    // ------------------------------------------------------------- control warp: TMA loads + MMA issue.
    // All 32 lanes run the loop (uniform control flow, operands in uniform registers); `el` predicates the issue.
    const uint32_t el = elect_one() ? 1u : 0u;
    constexpr uint32_t IDESC_QK = umma_idesc_f16(128, KT, 0, 0);
    constexpr uint32_t IDESC_PV = umma_idesc_f16(128, 64, 0, 1);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t dq0 = umma_desc_sw128(smem_base + C::OFF_Q, 0, 1024);
    const uint64_t dk0 = umma_desc_sw128(smem_base + C::OFF_K, 0, 1024);
    const uint64_t dv0 = umma_desc_sw128(smem_base + C::OFF_V, KV_TILE, 1024);
    int vent[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) vent[v] = p.vent[grp][v < nv ? v : 0];
    auto load_k = [&](int j) {
      const int st = j % KST;
      mbar_expect_tx_p(&k_full[st], KV_TILE, el);
      tma_load_4d_p(smem_base + C::OFF_K + st * KV_TILE, &tm_k, &k_full[st], 0, h, j * KT, b_qk, el);
    };
    auto load_v = [&](int j) {
      const int st = j % KST;
      mbar_expect_tx_p(&v_full[st], nv * KV_TILE, el);
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (v < nv) tma_load_4d_p(smem_base + C::OFF_V + (st * NV + v) * KV_TILE, &tm_v, &v_full[st], 0, h, j * KT, vent[v], el);
    };
    auto issue_qk = [&](int j) {
      const int st = j % KST;
      mbar_wait(&k_full[st], (j / KST) & 1);
      tc_fence_after();
      const uint64_t dk = dk0 + static_cast<uint64_t>((st * KV_TILE) >> 4);
      for (int kk = 0; kk < p.ksteps_qk; ++kk)
        mma_f16_ss_p(tmem, dq0 + 2 * kk, dk + 2 * kk, IDESC_QK, kk > 0, el);   // +32 bytes per 16-element k step
      tc_commit_p(s_full, el);
    };
    mbar_expect_tx_p(q_full, Q_TILE, el);
    tma_load_4d_p(smem_base + C::OFF_Q, &tm_q, q_full, 0, h, q0, b_qk, el);
    for (int j = 0; j < KST && j < nt; ++j) load_k(j);
    for (int j = 0; j < KST - 1 && j < nt; ++j) load_v(j);
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < nt; ++j) {
      if (j + 1 < nt) {
        mbar_wait(s_free, j & 1);           // Q K_j^T has completed and S_j lives in the softmax warps' registers
        tc_fence_after();
        issue_qk(j + 1);
        if (j + KST < nt) load_k(j + KST);  // into the stage of K_j
      }
      const int st = j % KST;
      mbar_wait(&v_full[st], (j / KST) & 1);
      mbar_wait(&p_full[j % PBUF], (j / PBUF) & 1);   // P_j written; the softmax warps saw P V_{j-PBUF} complete before writing it
      tc_fence_after();
      // the V stage of tile j-1 is free once P V_{j-1} has completed: known from p_full(j) with one P buffer; with two
      // buffers p_full(j) only implies P V_{j-2}, so the refill trails by one more tile
      if (PBUF == 1) { if (j + KST - 1 < nt) load_v(j + KST - 1); }
      else if (j >= 1 && j + KST - 2 < nt) load_v(j + KST - 2);
      const uint32_t pa = tmem_p + 32u * (j % PBUF);
      const uint64_t dvs = dv0 + static_cast<uint64_t>((st * NV * KV_TILE) >> 4);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (v < nv) {
#pragma unroll
          for (int kk = 0; kk < KT / 16; ++kk)   // V is MN-major: 16 keys further = +2048 bytes
            mma_f16_ts_p(tmem_o + 64u * v, pa + kk * 8, dvs + ((v * KV_TILE + kk * 2048) >> 4), IDESC_PV,
                         (j > 0 || kk > 0) ? 1u : 0u, el);
        }
      }
      tc_commit_p(&pv_done[j % PBUF], el);
      if (j == nt - 1) tc_commit_p(o_full, el);
    }
  } else {
    // ------------------------------------------------------------- softmax + epilogue
    // plain: thread == query row; groups: SPLIT = 2 threads per row (warps w and w + 4 share TMEM lane quadrant w & 3), each
    // owning NCOL = 32 of the 64 key columns and 32 of the 64 output columns: a group CTA has an SM to itself, and one
    // softmax warp per scheduler cannot hide its own TMEM / barrier / MUFU latencies (measured 1900 cycles per key tile
    // against 768 tensor cycles); two per scheduler overlap each other. The row max is exchanged through shared memory.
    constexpr int SPLIT = C::SPLIT, NCOL = KT / SPLIT;
    const int quad = warp & 3, part = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tlane = tmem + lane_off + NCOL * part;
    const bool row_ok = (q0 + row) < p.n_q;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    float* red = reinterpret_cast<float*>(smem + C::OFF_RED);   // [2 tile parities][SPLIT][128] row maxima, then [SPLIT][128] row sums
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < nt; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float s[NCOL];
#pragma unroll
      for (int c = 0; c < NCOL / 32; ++c) tmem_ld32(tlane + 32 * c, reinterpret_cast<uint32_t*>(s) + 32 * c);
#pragma unroll
      for (int c = 0; c < NCOL / 32; ++c) tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s) + 32 * c);
      tc_fence_before();
      mbar_arrive(s_free);
      const int valid = p.n_k - j * KT - NCOL * part;
      if (valid < NCOL) {
#pragma unroll
        for (int i = 0; i < NCOL; ++i)
          if (i >= valid) s[i] = -INFINITY;
      }
      float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
      for (int i = 4; i < NCOL; i += 4) {
#pragma unroll
        for (int c = 0; c < 4; ++c) m4[c] = fmaxf(m4[c], s[i + c]);
      }
      float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      if (SPLIT == 2) {
        float* rm = red + (j & 1) * 256;
        rm[part * 128 + row] = mx;
        asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quad) : "memory");   // the two warps that share these 32 rows
        mx = fmaxf(mx, rm[(part ^ 1) * 128 + row]);
      }
      const float mxs = mx * p.scale_log2;
      if (j == 0) {
        m_ref = mxs;
      } else {
        const bool need = mxs > m_ref + 8.f;   // lazy rescale: keeps P <= 2^8 in fp16 (identical in both threads of a row)
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(&pv_done[(j - 1) % PBUF], ((j - 1) / PBUF) & 1);   // O is being accumulated by P V_{j-1}
          tc_fence_after();
          const float alpha = need ? ex2_approx(m_ref - mxs) : 1.f;
          if (need) m_ref = mxs;
          l *= alpha;
          for (int c = 0; c < (4 / SPLIT) * nv; ++c) {   // this thread's NCOL output columns of every accumulator
            const uint32_t oc = tmem_o + lane_off + 64u * (c / (4 / SPLIT)) + NCOL * part + 16 * (c % (4 / SPLIT));
            uint32_t o[16];
            tmem_ld16(oc, o);
            tmem_wait_ld_regs16(o);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(oc, o);
          }
        }
      }
      const uint32_t tp = tmem_p + lane_off + 32u * (j % PBUF) + (NCOL / 2) * part;
      const float2 nm2 = make_float2(-m_ref, -m_ref);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      uint32_t pk[NCOL / 2];
#pragma unroll
      for (int i = 0; i < NCOL / 2; ++i) {
        const float2 x = fma_f32x2(make_float2(s[2 * i], s[2 * i + 1]), sc2, nm2);
        // POLY > 0: every POLY-th pair is evaluated on the FMA pipe (sa::exp2_poly2) instead of the MUFU pipe
        const float2 e = (POLY > 0 && i % (POLY > 0 ? POLY : 1) == POLY - 1) ? exp2_poly2(x)
                                                                             : make_float2(ex2_approx(x.x), ex2_approx(x.y));
        if (i & 1) acc1 = add_f32x2(acc1, e); else acc0 = add_f32x2(acc0, e);
        pk[i] = pack_half2(e.x, e.y);
      }
      if (j >= PBUF) {   // the P buffer of this tile was last read by P V_{j-PBUF}: waited for AFTER the exponentials,
        mbar_wait(&pv_done[j % PBUF], ((j / PBUF) - 1) & 1);   // when that MMA has long completed
        tc_fence_after();
      }
      if (SPLIT == 1) tmem_st32(tp, pk); else tmem_st16(tp, pk);
      l += (acc0.x + acc0.y) + (acc1.x + acc1.y);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[j % PBUF]);
    }
    // ---- epilogue: O_v / l -> fp16 -> swizzled staging tile (the Q tile, dead by now) -> TMA store, one member at a time
    if (SPLIT == 2) {   // row sum = the two column halves (the rescale factors applied to l were identical in both)
      float* rl = red + 512;
      rl[part * 128 + row] = l;
      asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quad) : "memory");
      l += rl[(part ^ 1) * 128 + row];
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.f / l;
    uint8_t* orow = smem + C::OFF_Q + row * 128;
    const int sw = row & 7;
    for (int v = 0; v < nv; ++v) {
      if (v > 0) {
        if (threadIdx.x == 0) tma_store_wait_read();   // the previous member's store has drained the staging tile
        asm volatile("bar.sync 5, %0;\n" ::"n"(128 * SPLIT) : "memory");
      }
#pragma unroll
      for (int hh = 0; hh < 2 / SPLIT; ++hh) {
        const int cb = SPLIT == 2 ? part : hh;          // which 32-column half of the 64 output columns
        uint32_t o[32];
        tmem_ld32(tmem_o + lane_off + 64u * v + 32 * cb, o);
        tmem_wait_ld_regs32(o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack_half2(__uint_as_float(o[8 * q + 0]) * inv_l, __uint_as_float(o[8 * q + 1]) * inv_l);
          w.y = pack_half2(__uint_as_float(o[8 * q + 2]) * inv_l, __uint_as_float(o[8 * q + 3]) * inv_l);
          w.z = pack_half2(__uint_as_float(o[8 * q + 4]) * inv_l, __uint_as_float(o[8 * q + 5]) * inv_l);
          w.w = pack_half2(__uint_as_float(o[8 * q + 6]) * inv_l, __uint_as_float(o[8 * q + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + (((cb * 4 + q) ^ sw) << 4)) = w;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      asm volatile("bar.sync 5, %0;\n" ::"n"(128 * SPLIT) : "memory");
      const int b = p.vent[grp][v];
      if (threadIdx.x == 0) {
        tma_store_4d(&tm_o, smem + C::OFF_Q, 0, h, q0, b);
        tma_store_commit();
      }
      if (p.lse != nullptr && row_ok && part == 0)
        p.lse[(static_cast<size_t>(b) * p.heads + h) * p.n_q + q0 + row] = m_ref + log2f(l);
    }
    if (threadIdx.x == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CW) {
    tmem_dealloc<C::COLS_A>(tmem);
    if (NV == 1) tmem_dealloc<32>(tmem_p);
  }
}

// RTTI_ATTN_POLY = 0 | 2 | 3 | 4 (read once; default 4): fraction 1/POLY of the exponentials of the PLAIN kernel is evaluated by
// a polynomial on the FMA pipe instead of MUFU.EX2 (the plain kernel is bound by the 16-lane/SM MUFU pipe). Measured on B200
// (profiles/r02_kernels_selfattn_poly_sweep.jsonl): 1/4 -> +4.4 % / +4.8 % at the 64^2 / 32^2 level, 1/2 -> -1 % (issue-bound).
static const int g_poly = [] { const char* e = getenv("RTTI_ATTN_POLY"); const int v = e ? atoi(e) : SA_POLY_DEFAULT; return (v == 2 || v == 3 || v == 4) ? v : 0; }();

template <int NV, int POLY>
static int launch_variant(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                          const AttnSelfParams& p, int n_groups, cudaStream_t stream) {
  using C = sa::Cfg<NV>;
  static const bool configured =
      cudaFuncSetAttribute(attn_self_kernel<NV, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) == cudaSuccess;
  if (!configured) return RTTI_ERR_CUDA;
  dim3 grid((p.n_q + 127) / 128, p.heads, n_groups);
  attn_self_kernel<NV, POLY><<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(tq, tk, tv, to, p);
  return cudaGetLastError() == cudaSuccess ? RTTI_OK : RTTI_ERR_CUDA;
}

template <int NV>
static int launch_class(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                        const AttnSelfParams& p, int n_groups, cudaStream_t stream) {
  if (NV == 1 && g_poly == 2) return launch_variant<NV, (NV == 1 ? 2 : 0)>(tq, tk, tv, to, p, n_groups, stream);
  if (NV == 1 && g_poly == 3) return launch_variant<NV, (NV == 1 ? 3 : 0)>(tq, tk, tv, to, p, n_groups, stream);
  if (NV == 1 && g_poly == 4) return launch_variant<NV, (NV == 1 ? 4 : 0)>(tq, tk, tv, to, p, n_groups, stream);
  return launch_variant<NV, 0>(tq, tk, tv, to, p, n_groups, stream);
}

// Library-owned side stream + fork/join events, one set per device (created on first use, never destroyed).
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; };
static SideStream* side_stream() {
  static SideStream per_dev[16];
  static bool ready[16] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (!ready[dev]) {
    SideStream& s = per_dev[dev];
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ready[dev] = true;
  }
  return &per_dev[dev];
}

// Entries are grouped by their score source; single entries go to the 3-CTA/SM plain kernel, groups of 2..6 members to
// the 512-column group kernel (larger groups are split into balanced chunks). At most two launches, groups first.
int launch_attn_self(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                     int batch, int heads, int head_dim, int n_q, int n_k, float scale_log2, const int8_t* qk_src,
                     float* lse, int max_group, cudaStream_t stream) {
  AttnSelfParams base{};
  base.heads = heads; base.head_dim = head_dim; base.n_q = n_q; base.n_k = n_k;
  base.n_k_tiles = (n_k + sa::KT - 1) / sa::KT;
  base.ksteps_qk = (head_dim + 15) / 16;
  base.scale_log2 = scale_log2;
  base.lse = lse;
  if (max_group < 1) max_group = 1;
  if (max_group > 6) max_group = 6;
  AttnSelfParams cls[2] = {base, base};   // plain, groups
  int n[2] = {0, 0};
  for (int s = 0; s < batch; ++s) {
    int members[64], m = 0;
    for (int b = 0; b < batch; ++b)
      if (qk_src[b] == s) members[m++] = b;
    if (m == 0) continue;
    const int chunks = (m + max_group - 1) / max_group;   // balanced split of groups larger than max_group
    for (int c = 0, off = 0; c < chunks; ++c) {
      const int take = m / chunks + (c < m % chunks ? 1 : 0);
      const int k = take == 1 ? 0 : 1;
      AttnSelfParams& q = cls[k];
      const int g = n[k]++;
      if (g >= 64) return RTTI_ERR_ARG;
      q.qk[g] = (int8_t)s; q.nv[g] = (int8_t)take;
      for (int i = 0; i < take; ++i) q.vent[g][i] = (int8_t)members[off + i];
      off += take;
    }
  }
  int rc = RTTI_OK;
  if (n[1] && n[0]) {
    // Both classes occur (an injection step): run them CONCURRENTLY. Each has a poor tail on its own (a group CTA owns a
    // whole SM: 160 / 320 CTAs on 148 SMs at the 32^2 / 64^2 level; the plain kernel 480 / 960 CTAs on 444 slots), and the
    // block scheduler starts the second kernel's CTAs as soon as the first one has none left to dispatch, so each
    // kernel's last wave is filled by the other. Fork / join through events on a library-owned side stream: capturable
    // into a CUDA graph (the events become graph edges), no host synchronisation.
    SideStream* ss = side_stream();
    if (ss == nullptr) return RTTI_ERR_CUDA;
    if (cudaEventRecord(ss->fork, stream) != cudaSuccess || cudaStreamWaitEvent(ss->stream, ss->fork, 0) != cudaSuccess)
      return RTTI_ERR_CUDA;
    if ((rc = launch_class<6>(tq, tk, tv, to, cls[1], n[1], stream)) != RTTI_OK) return rc;
    if ((rc = launch_class<1>(tq, tk, tv, to, cls[0], n[0], ss->stream)) != RTTI_OK) return rc;
    if (cudaEventRecord(ss->join, ss->stream) != cudaSuccess || cudaStreamWaitEvent(stream, ss->join, 0) != cudaSuccess)
      return RTTI_ERR_CUDA;
    return RTTI_OK;
  }
  if (n[1] && (rc = launch_class<6>(tq, tk, tv, to, cls[1], n[1], stream)) != RTTI_OK) return rc;
  if (n[0] && (rc = launch_class<1>(tq, tk, tv, to, cls[0], n[0], stream)) != RTTI_OK) return rc;
  return rc;
}

}  // namespace rtti

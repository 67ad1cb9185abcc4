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
//   TMEM   S [0,64) fp32 scores of ONE 64-key tile, O_v [64 + 64 v, +64) fp32 accumulators  -> 128 / 256 / 512 columns
//   smem   Q tile 16 KB | P tile 16 KB (fp16, K-major SWIZZLE_128B, also the output staging tile) | K ring 2 x 8 KB |
//          V ring 2 x NV x 8 KB
//   * a softmax thread owns one query row: tcgen05.ld of its 64 scores, then IMMEDIATELY releases S (`s_free`), so the
//     MMA warp issues Q K_{j+1}^T while the exponentials of tile j are still being evaluated (register double buffering:
//     the look-ahead of a second S buffer without its TMEM columns);
//   * P_j goes to shared memory in the UMMA K-major swizzled layout (8 x st.shared.v4 per row) and feeds P V as an SS MMA;
//   * with 128 TMEM columns and 64 KB of shared memory THREE CTAs share an SM (NV = 1): 12 softmax warps keep the MUFU
//     pipe busy across each other's TMEM-load / max / store phases (the previous schedule: 8 warps, 53 % XU utilisation);
//   * no "stage empty" barriers: the control thread learns that Q K_j^T has completed from `s_free` (the softmax warps
//     read S_j only after it) and that P V_{j-1} has completed from `p_full` (they write P_j only after it), so it
//     refills the K stage of tile j with tile j+2 and the V stage of tile j-1 with tile j+1 at those points;
//   * lazy rescale of O (only when a row max grows by more than 2^8) as before, on all NV accumulators.
#include "ptx.cuh"
#include "rtti_internal.h"

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
constexpr int KSTAGE = 2;
constexpr int Q_TILE = 128 * 128;   // bytes
constexpr int KV_TILE = KT * 128;   // 8 KB
constexpr int THREADS = 160;
template <int NV> struct Cfg {
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_P = OFF_Q + Q_TILE;
  static constexpr int OFF_K = OFF_P + Q_TILE;
  static constexpr int OFF_V = OFF_K + KSTAGE * KV_TILE;
  static constexpr int OFF_BAR = OFF_V + KSTAGE * NV * KV_TILE;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;   // + alignment slack
  static constexpr uint32_t TMEM_COLS = (64 + 64 * NV) <= 128 ? 128 : ((64 + 64 * NV) <= 256 ? 256 : 512);
  static constexpr int MIN_CTAS = NV == 1 ? 3 : (NV <= 3 ? 2 : 1);
  static constexpr int MAX_REGS = NV == 1 ? 136 : (NV <= 3 ? 200 : 255);   // 65536 / (160 threads x MIN_CTAS), 8-register granules
};
}  // namespace sa

template <int NV>
__global__ void __launch_bounds__(sa::THREADS) __maxnreg__((sa::Cfg<NV>::MAX_REGS))
attn_self_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                 const __grid_constant__ AttnSelfParams p) {
  using namespace sa;
  using C = Cfg<NV>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* v_full = bars + 3;        // [2]
  uint64_t* s_full = bars + 5;        // QK^T_j landed in TMEM
  uint64_t* s_free = bars + 6;        // all 128 rows of S_j are in registers
  uint64_t* p_full = bars + 7;        // P_j is in shared memory (and P V_{j-1} has completed)
  uint64_t* pv_done = bars + 8;       // P V_j has completed (P tile reusable, O consistent)
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int grp = blockIdx.z;
  const int b_qk = p.qk[grp];
  const int nv = NV == 1 ? 1 : p.nv[grp];
  const int nt = p.n_k_tiles;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
      mbar_init(q_full, 1);
      for (int i = 0; i < KSTAGE; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); }
      mbar_init(s_full, 1); mbar_init(s_free, 128); mbar_init(p_full, 128); mbar_init(pv_done, 1); mbar_init(o_full, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<C::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------- control thread: TMA loads + MMA issue
    if (lane == 0) {
      constexpr uint32_t IDESC_QK = umma_idesc_f16(128, KT, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_f16(128, 64, 0, 1);
      const uint32_t smem_base = smem_u32(smem);
      auto load_k = [&](int j) {
        const int st = j % KSTAGE;
        mbar_expect_tx(&k_full[st], KV_TILE);
        tma_load_4d(smem + C::OFF_K + st * KV_TILE, &tm_k, &k_full[st], 0, h, j * KT, b_qk);
      };
      auto load_v = [&](int j) {
        const int st = j % KSTAGE;
        mbar_expect_tx(&v_full[st], nv * KV_TILE);
        for (int v = 0; v < nv; ++v)
          tma_load_4d(smem + C::OFF_V + (st * NV + v) * KV_TILE, &tm_v, &v_full[st], 0, h, j * KT, p.vent[grp][v]);
      };
      auto issue_qk = [&](int j) {
        const int st = j % KSTAGE;
        mbar_wait(&k_full[st], (j / KSTAGE) & 1);
        tc_fence_after();
        for (int kk = 0; kk < p.ksteps_qk; ++kk) {
          const uint64_t da = umma_desc_sw128(smem_base + C::OFF_Q + kk * 32, 0, 1024);
          const uint64_t db = umma_desc_sw128(smem_base + C::OFF_K + st * KV_TILE + kk * 32, 0, 1024);
          mma_f16_ss(tmem, da, db, IDESC_QK, kk > 0);
        }
        tc_commit(s_full);
      };
      mbar_expect_tx(q_full, Q_TILE);
      tma_load_4d(smem + C::OFF_Q, &tm_q, q_full, 0, h, q0, b_qk);
      load_k(0);
      if (nt > 1) load_k(1);
      load_v(0);
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) {
          mbar_wait(s_free, j & 1);       // Q K_j^T has completed and S_j lives in the softmax warps' registers
          tc_fence_after();
          issue_qk(j + 1);
          if (j + 2 < nt) load_k(j + 2);  // into the stage of K_j
        }
        const int st = j % KSTAGE;
        mbar_wait(&v_full[st], (j / KSTAGE) & 1);
        mbar_wait(p_full, j & 1);         // P_j written; the softmax warps saw pv_done(j-1) before writing it
        tc_fence_after();
        if (j + 1 < nt) load_v(j + 1);    // into the stage of V_{j-1}
        for (int v = 0; v < nv; ++v) {
#pragma unroll
          for (int kk = 0; kk < KT / 16; ++kk) {
            const uint64_t da = umma_desc_sw128(smem_base + C::OFF_P + kk * 32, 0, 1024);
            const uint64_t db = umma_desc_sw128(smem_base + C::OFF_V + (st * NV + v) * KV_TILE + kk * 2048, KV_TILE, 1024);
            mma_f16_ss(tmem + 64u + 64u * v, da, db, IDESC_PV, (j > 0 || kk > 0) ? 1u : 0u);
          }
        }
        tc_commit(pv_done);
        if (j == nt - 1) tc_commit(o_full);
      }
    }
  } else {
    // ------------------------------------------------------------- softmax + epilogue: thread == query row == TMEM lane
    const int row = warp * 32 + lane;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const bool row_ok = (q0 + row) < p.n_q;
    uint8_t* prow = smem + C::OFF_P + row * 128;
    const int sw = row & 7;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < nt; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float s[64];
      tmem_ld32(tlane, reinterpret_cast<uint32_t*>(s));
      tmem_ld32(tlane + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s));
      tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s) + 32);
      tc_fence_before();
      mbar_arrive(s_free);
      const int valid = p.n_k - j * KT;
      if (valid < KT) {
#pragma unroll
        for (int i = 0; i < KT; ++i)
          if (i >= valid) s[i] = -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < KT; ++i) mx = fmaxf(mx, s[i]);
      const float mxs = mx * p.scale_log2;
      if (j > 0) {   // P tile and (for a rescale) the O accumulators are free once P V_{j-1} has completed
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
      }
      if (j == 0) {
        m_ref = mxs;
      } else {
        const bool need = mxs > m_ref + 8.f;   // lazy rescale: keeps P <= 2^8 in fp16
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2_approx(m_ref - mxs) : 1.f;
          if (need) m_ref = mxs;
          l *= alpha;
          for (int c = 0; c < 4 * nv; ++c) {
            uint32_t o[16];
            tmem_ld16(tlane + 64u + 16 * c, o);
            tmem_wait_ld_regs16(o);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tlane + 64u + 16 * c, o);
          }
          tmem_wait_st();
        }
      }
      const float2 nm2 = make_float2(-m_ref, -m_ref);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 8; ++c) {   // 8 keys -> one 16-byte chunk of the swizzled P row
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = fma_f32x2(make_float2(s[8 * c + 2 * i], s[8 * c + 2 * i + 1]), sc2, nm2);
          const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          if (i & 1) acc1 = add_f32x2(acc1, e); else acc0 = add_f32x2(acc0, e);
          w[i] = pack_half2(e.x, e.y);
        }
        *reinterpret_cast<uint4*>(prow + ((c ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      l += (acc0.x + acc0.y) + (acc1.x + acc1.y);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O_v / l -> fp16 -> swizzled staging tile (the P tile) -> TMA store, one member at a time
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.f / l;
    for (int v = 0; v < nv; ++v) {
      if (v > 0) {
        if (threadIdx.x == 0) tma_store_wait_read();   // the previous member's store has drained the staging tile
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tmem_ld32(tlane + 64u + 64u * v + 32 * hh, o);
        tmem_wait_ld_regs32(o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack_half2(__uint_as_float(o[8 * q + 0]) * inv_l, __uint_as_float(o[8 * q + 1]) * inv_l);
          w.y = pack_half2(__uint_as_float(o[8 * q + 2]) * inv_l, __uint_as_float(o[8 * q + 3]) * inv_l);
          w.z = pack_half2(__uint_as_float(o[8 * q + 4]) * inv_l, __uint_as_float(o[8 * q + 5]) * inv_l);
          w.w = pack_half2(__uint_as_float(o[8 * q + 6]) * inv_l, __uint_as_float(o[8 * q + 7]) * inv_l);
          *reinterpret_cast<uint4*>(prow + (((hh * 4 + q) ^ sw) << 4)) = w;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      const int b = p.vent[grp][v];
      if (threadIdx.x == 0) {
        tma_store_4d(&tm_o, smem + C::OFF_P, 0, h, q0, b);
        tma_store_commit();
      }
      if (p.lse != nullptr && row_ok)
        p.lse[(static_cast<size_t>(b) * p.heads + h) * p.n_q + q0 + row] = m_ref + log2f(l);
    }
    if (threadIdx.x == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<C::TMEM_COLS>(tmem);
}

template <int NV>
static int launch_class(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                        const AttnSelfParams& p, int n_groups, cudaStream_t stream) {
  using C = sa::Cfg<NV>;
  static const bool configured =
      cudaFuncSetAttribute(attn_self_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) == cudaSuccess;
  if (!configured) return RTTI_ERR_CUDA;
  dim3 grid((p.n_q + 127) / 128, p.heads, n_groups);
  attn_self_kernel<NV><<<grid, sa::THREADS, C::SMEM_BYTES, stream>>>(tq, tk, tv, to, p);
  return cudaGetLastError() == cudaSuccess ? RTTI_OK : RTTI_ERR_CUDA;
}

// Entries are grouped by their score source; groups of 1 / 2-3 / 4-6 members go to the 128 / 256 / 512-column kernel
// (larger groups are split). One launch per class that occurs (at most three), heaviest class first.
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
  AttnSelfParams cls[3] = {base, base, base};   // NV = 1, 3, 6
  int n[3] = {0, 0, 0};
  for (int s = 0; s < batch; ++s) {
    int members[64], m = 0;
    for (int b = 0; b < batch; ++b)
      if (qk_src[b] == s) members[m++] = b;
    if (m == 0) continue;
    const int chunks = (m + max_group - 1) / max_group;   // balanced split of groups larger than max_group
    for (int c = 0, off = 0; c < chunks; ++c) {
      const int take = m / chunks + (c < m % chunks ? 1 : 0);
      const int k = take == 1 ? 0 : (take <= 3 ? 1 : 2);
      AttnSelfParams& q = cls[k];
      const int g = n[k]++;
      if (g >= 64) return RTTI_ERR_ARG;
      q.qk[g] = (int8_t)s; q.nv[g] = (int8_t)take;
      for (int i = 0; i < take; ++i) q.vent[g][i] = (int8_t)members[off + i];
      off += take;
    }
  }
  int rc = RTTI_OK;
  if (n[2] && (rc = launch_class<6>(tq, tk, tv, to, cls[2], n[2], stream)) != RTTI_OK) return rc;
  if (n[1] && (rc = launch_class<3>(tq, tk, tv, to, cls[1], n[1], stream)) != RTTI_OK) return rc;
  if (n[0] && (rc = launch_class<1>(tq, tk, tv, to, cls[0], n[0], stream)) != RTTI_OK) return rc;
  return rc;
}

}  // namespace rtti

// Cross-attention forward (77 text keys, head_dim <= 64) as a persistent streaming kernel for sm_100a.
//
// Reference: models/attention_processor.py:1143-1169 (Q K^T -> softmax -> P V) with the font-size token re-weighting of
// :387-399 (E = exp(S - max) in fp32; E[:, :, pos] *= |fs|; P = E / sum E; P[:, :, pos] *= sign(fs); cast) for the batch
// entries of pass B (hooks models/region_diffusion_sdxl.py:1112-1140).
//
// The op is HBM-bound (71 FLOP/B): per (batch, head, 128-query tile) "unit" it streams 16 KB of Q in
// and 16 KB of O out around ~0.4 MFLOP; K and V (77 x 64 fp16 each) come from L2. The first kernel (attn_fwd.cu, one CTA
// per query tile x a few heads, two Q buffers, 2 CTAs/SM) reached 0.25-0.37 of the HBM roof: its warps waited for TMA
// loads (ncu: s_full / o_full waits dominate) because only ~one unit per CTA was ever in flight. This kernel keeps
// FIVE units in flight per SM:
//   * persistent CTAs (one per SM), each owning a contiguous range of units (query tile fastest, so consecutive units
//     share K/V and hit L2);
//   * a 5-stage ring of {Q, K, V} tiles (36 KB per stage) filled by TMA, refilled as soon as P V of a unit has completed;
//   * TMEM: two S/P buffers and two O accumulators, two softmax warpgroups alternating units, so Q K^T of unit i+1 and
//     the softmax of unit i overlap the epilogue (TMEM -> fp16 -> swizzled staging -> TMA store) of unit i-1;
//   * one control warp issues all TMA loads and MMAs with warp-uniform control flow and predicated issue (ptx.cuh).
// The whole row (80 columns) lives in registers: P is normalised BEFORE the fp16 rounding exactly as the reference does.
#include "ptx.cuh"
#include "rtti_internal.h"

namespace rtti {

struct AttnCrossParams {
  int batch, heads, n_q, n_k, q_tiles, ksteps_qk;
  int n_units;
  float scale_log2;
  unsigned long long fs_mask;   // batch entries that get the font-size re-weighting
  const int* word_pos;
  const float* font_size;
  int n_fs;
};

namespace ca {
constexpr int KT = 80;
constexpr int NQ = 5;                       // ring depth
constexpr int Q_TILE = 128 * 128;           // 16 KB
constexpr int KV_TILE = KT * 128;           // 10 KB
constexpr int STAGE = Q_TILE + 2 * KV_TILE; // 36 KB
constexpr int OFF_O = NQ * STAGE;           // two 16 KB staging tiles
constexpr int OFF_BAR = OFF_O + 2 * Q_TILE;
constexpr int OFF_FS = OFF_BAR + 256;
constexpr int SMEM_BYTES = OFF_FS + 128 * 4 + 1024;
constexpr int THREADS = 288;                // 2 softmax warpgroups + 1 control warp
constexpr int CW = 8;
}  // namespace ca

__global__ void __maxnreg__(168)   // 9 warps are checked as 12 (register allocation is verified per 4 sub-partitions): 65536 / (12 x 32) = 170
attn_cross_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                  const __grid_constant__ AttnCrossParams p) {
  using namespace ca;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full = bars;               // [NQ] stage loaded
  uint64_t* empty = bars + NQ;         // [NQ] P V of the unit in this stage has completed
  uint64_t* s_full = bars + 2 * NQ;    // [2]  Q K^T landed in S_g
  uint64_t* p_full = bars + 2 * NQ + 2;    // [2]  P written to TMEM by group g
  uint64_t* o_full = bars + 2 * NQ + 4;    // [2]  P V complete: O_g ready, S_g / P_g reusable
  uint64_t* o_free = bars + 2 * NQ + 6;    // [2]  group g has read O_g
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NQ + 8);
  float* fs_w = reinterpret_cast<float*>(smem + OFF_FS);   // signed font-size weight per key (1 = untouched)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous, balanced range of units for this CTA
  const int per = p.n_units / gridDim.x, rem = p.n_units % gridDim.x;
  const int u0 = blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem);
  const int n = per + (blockIdx.x < rem ? 1 : 0);

  if (warp == CW) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
      for (int i = 0; i < NQ; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 128);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  if (threadIdx.x < 128) fs_w[threadIdx.x] = 1.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0 && p.n_fs > 0) {
    // duplicates in word_pos: last write wins, as the reference's advanced-index assignment (attention_processor.py:393-396)
    for (int i = 0; i < p.n_fs; ++i) {
      const int pos = p.word_pos[i];
      if (pos >= 0 && pos < p.n_k) fs_w[pos] = p.font_size[i];
    }
  }
  __syncthreads();
  const uint32_t tmem = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  if (warp == CW) {
    // ------------------------------------------------------------- control warp: TMA loads + MMA issue
    const uint32_t el = elect_one() ? 1u : 0u;
    constexpr uint32_t IDESC_QK = umma_idesc_f16(128, KT, 0, 0);
    constexpr uint32_t IDESC_PV = umma_idesc_f16(128, 64, 0, 1);
    const uint64_t dq0 = umma_desc_sw128(smem_base, 0, 1024);
    const uint64_t dk0 = umma_desc_sw128(smem_base + Q_TILE, 0, 1024);
    const uint64_t dv0 = umma_desc_sw128(smem_base + Q_TILE + KV_TILE, KV_TILE, 1024);
    auto load = [&](int i) {
      const int u = u0 + i, s = i % NQ;
      const int qt = u % p.q_tiles, bh = u / p.q_tiles;
      const int h = bh % p.heads, b = bh / p.heads;
      const uint32_t base = smem_base + s * STAGE;
      mbar_expect_tx_p(&full[s], STAGE, el);
      tma_load_4d_p(base, &tm_q, &full[s], 0, h, qt * 128, b, el);
      tma_load_4d_p(base + Q_TILE, &tm_k, &full[s], 0, h, 0, b, el);
      tma_load_4d_p(base + Q_TILE + KV_TILE, &tm_v, &full[s], 0, h, 0, b, el);
    };
    for (int i = 0; i < NQ && i < n; ++i) load(i);
    for (int i = 0; i <= n; ++i) {
      if (i < n) {
        const int g = i & 1, s = i % NQ;
        mbar_wait(&full[s], (i / NQ) & 1);
        if (i >= 2) mbar_wait(&o_full[g], ((i - 2) >> 1) & 1);   // P V of unit i-2 complete: S_g / P_g are free
        tc_fence_after();
        const uint64_t so = static_cast<uint64_t>((s * STAGE) >> 4);
        for (int kk = 0; kk < p.ksteps_qk; ++kk)
          mma_f16_ss_p(tmem + 128u * g, dq0 + so + 2 * kk, dk0 + so + 2 * kk, IDESC_QK, kk > 0, el);
        tc_commit_p(&s_full[g], el);
      }
      if (i >= 1) {
        const int j = i - 1, g = j & 1, s = j % NQ;
        mbar_wait(&p_full[g], (j >> 1) & 1);
        if (j >= 2) mbar_wait(&o_free[g], ((j - 2) >> 1) & 1);   // the epilogue of unit j-2 has read O_g
        tc_fence_after();
        const uint64_t so = static_cast<uint64_t>((s * STAGE) >> 4);
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk)
          mma_f16_ts_p(tmem + 256u + 64u * g, tmem + 128u * g + kk * 8, dv0 + so + ((kk * 2048) >> 4), IDESC_PV, kk > 0, el);
        tc_commit_p(&o_full[g], el);
        tc_commit_p(&empty[s], el);
      }
      if (i >= 2 && i - 2 + NQ < n) {   // the stage of unit i-2 was released by its P V, issued one iteration ago
        mbar_wait(&empty[(i - 2) % NQ], ((i - 2) / NQ) & 1);
        load(i - 2 + NQ);
      }
    }
  } else {
    // ------------------------------------------------------------- softmax + epilogue: group g takes units g, g+2, ...
    const int g = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t ts = tmem + lane_off + 128u * g;          // S / P of this group
    const uint32_t to = tmem + lane_off + 256u + 64u * g;    // O of this group
    uint8_t* orow = smem + OFF_O + g * Q_TILE + row * 128;
    const int sw = row & 7;
    const bool leader = quad == 0 && lane == 0;
    const int bar_id = 1 + g;
    for (int i = g, it = 0; i < n; i += 2, ++it) {
      const int u = u0 + i;
      const int qt = u % p.q_tiles, bh = u / p.q_tiles;
      const int h = bh % p.heads, b = bh / p.heads;
      const bool use_fs = p.n_fs > 0 && ((p.fs_mask >> b) & 1ull);
      mbar_wait(&s_full[g], it & 1);
      tc_fence_after();
      float s[KT];
      tmem_ld32(ts, reinterpret_cast<uint32_t*>(s));
      tmem_ld32(ts + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tmem_ld16(ts + 64, reinterpret_cast<uint32_t*>(s) + 64);
      tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s));
      tmem_wait_ld_regs32(reinterpret_cast<uint32_t*>(s) + 32);
      tmem_wait_ld_regs16(reinterpret_cast<uint32_t*>(s) + 64);
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k >= p.n_k) s[k] = -INFINITY;
      float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
      for (int k = 4; k < KT; k += 4) {
#pragma unroll
        for (int c = 0; c < 4; ++c) m4[c] = fmaxf(m4[c], s[k + c]);
      }
      const float m_ref = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
      float rowsum = 0.f;
      if (use_fs) {
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          s[k] = ex2_approx(fmaf(s[k], p.scale_log2, -m_ref)) * fabsf(fs_w[k]);
          rowsum += s[k];
        }
        const float inv = 1.f / rowsum;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          const float w = fs_w[k];
          s[k] = s[k] * inv * (w > 0.f ? 1.f : (w < 0.f ? -1.f : 0.f));
        }
      } else {
        float r4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          s[k] = ex2_approx(fmaf(s[k], p.scale_log2, -m_ref));
          r4[k & 3] += s[k];
        }
        rowsum = (r4[0] + r4[1]) + (r4[2] + r4[3]);
        const float inv = 1.f / rowsum;
#pragma unroll
        for (int k = 0; k < KT; ++k) s[k] *= inv;
      }
      uint32_t pk[KT / 2];
#pragma unroll
      for (int k = 0; k < KT / 2; ++k) pk[k] = pack_half2(s[2 * k], s[2 * k + 1]);
      tmem_st32(ts, pk);                     // P over the first 40 columns of its own S buffer (S is in registers)
      tmem_st8(ts + 32, pk + 32);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
      // ---- epilogue: O (already normalised through P) -> fp16 -> swizzled staging tile -> TMA store
      mbar_wait(&o_full[g], it & 1);
      tc_fence_after();
      if (leader) tma_store_wait_read();     // the previous store of this group has drained its staging tile
      asm volatile("bar.sync %0, 128;\n" ::"r"(bar_id) : "memory");
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tmem_ld32(to + 32 * hh, o);
        tmem_wait_ld_regs32(o);
        if (hh == 1) {                       // O_g is in registers: the control warp may accumulate the next unit into it
          tc_fence_before();
          mbar_arrive(&o_free[g]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack_half2(__uint_as_float(o[8 * q + 0]), __uint_as_float(o[8 * q + 1]));
          w.y = pack_half2(__uint_as_float(o[8 * q + 2]), __uint_as_float(o[8 * q + 3]));
          w.z = pack_half2(__uint_as_float(o[8 * q + 4]), __uint_as_float(o[8 * q + 5]));
          w.w = pack_half2(__uint_as_float(o[8 * q + 6]), __uint_as_float(o[8 * q + 7]));
          *reinterpret_cast<uint4*>(orow + (((hh * 4 + q) ^ sw) << 4)) = w;
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;\n" ::"r"(bar_id) : "memory");
      if (leader) {
        tma_store_4d(&tm_o, smem + OFF_O + g * Q_TILE, 0, h, qt * 128, b);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CW) tmem_dealloc<512>(tmem);
}

// Maps: Q / O with 128-row boxes, K / V with 80-row boxes (rows beyond n_k are zero-filled by TMA and masked in-kernel).
int launch_attn_cross(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                      int batch, int heads, int head_dim, int n_q, int n_k, float scale_log2, unsigned long long fs_mask,
                      const int* word_pos, const float* font_size, int n_fs, cudaStream_t stream) {
  static const bool configured =
      cudaFuncSetAttribute(attn_cross_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ca::SMEM_BYTES) == cudaSuccess;
  if (!configured) return RTTI_ERR_CUDA;
  static const int n_sm = [] {
    int dev = 0, v = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  AttnCrossParams p{};
  p.batch = batch; p.heads = heads; p.n_q = n_q; p.n_k = n_k;
  p.q_tiles = (n_q + 127) / 128;
  p.ksteps_qk = (head_dim + 15) / 16;
  p.n_units = batch * heads * p.q_tiles;
  p.scale_log2 = scale_log2;
  p.fs_mask = n_fs > 0 ? fs_mask : 0ull;
  p.word_pos = word_pos; p.font_size = font_size; p.n_fs = n_fs;
  const int grid = p.n_units < n_sm ? p.n_units : n_sm;
  attn_cross_kernel<<<grid, ca::THREADS, ca::SMEM_BYTES, stream>>>(tq, tk, tv, to, p);
  return cudaGetLastError() == cudaSuccess ? RTTI_OK : RTTI_ERR_CUDA;
}

}  // namespace rtti

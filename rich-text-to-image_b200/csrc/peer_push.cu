// Producer -> consumers hand-off of one activation over NVLink peer memory (multi-GPU region parallelism).
//
// On feature-injection steps every region pass E_j consumes, in each of the 70 self-attention layers, the Q and K of the
// reference pass D (the reference stores D's probabilities and replaces those of the region passes with them:
// models/region_diffusion_sdxl.py:1018-1061), plus one resnet feature map (resnet.py:639-641). Round 1 REPLICATED
// pass D on every rank that owned a region pass, so those ranks ran two passes per step while the others ran one.
// Here pass D runs on ONE rank, which pushes each layer's [tokens, 2C] Q|K slab (10.5 MB at the 64^2 level, 5.2 MB at
// 32^2: 0.42 GB per step) into a per-layer receive buffer of every rank that owns a region pass:
//
//   * peer_push_kernel — copies `rows` x `row_bytes` from a (strided) local tensor into the same offset of up to 15 peer
//     buffers with 128-bit stores; the last CTA to finish publishes the event's sequence number to every destination's
//     flag word (st.release.sys). It runs on a side stream next to the producer's own attention kernel.
//   * peer_wait_kernel — one thread on the consumer spins (ld.acquire.sys on LOCAL memory) until the flag reaches the
//     event's sequence number; the consumer's attention kernel, stream-ordered behind it, TMA-loads Q and K from the
//     receive buffer and V from its own projection.
//
// Sequence numbers are relative to one UNet pass (event 1, 2, ...) and added to the sequence base word flags[8], which
// rtti_peer_seq_advance moves at the end of the pass — identical kernel arguments every step, so both sides live inside
// the per-rank CUDA graph of the pass. Receive buffers are single-buffered: the producer starts step s+1 only after its
// gather+blend kernel of step s has seen the noise predictions of every region pass, which their owners publish after
// their UNet pass — the last reader of the buffers — has completed. A peer that never arrives trips a ~4 s timeout that
// raises the error word flags[1] (sticky: later waits return at once), never a hung GPU.
#include "peer_sync.cuh"
#include "rtti_internal.h"

namespace rtti {

constexpr int PUSH_MAX_DST = 15;
constexpr int PUSH_SEQ_BASE_WORD = 8;

struct PeerPushParams {
  uint4* dst[PUSH_MAX_DST];           // peer-mapped receive buffers (already offset to this event's region)
  unsigned int* dst_flags[PUSH_MAX_DST];
  int n_dst;
  const uint8_t* src;
  long long src_row_stride;           // bytes
  int rows, row_vec;                  // row_vec = row_bytes / 16
  unsigned int* flags;                // local: [1] error, [3] CTA arrival counter, [8] sequence base
  unsigned int seq;
};

__global__ void __launch_bounds__(256) peer_push_kernel(const PeerPushParams p) {
  const long long total = (long long)p.rows * p.row_vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / p.row_vec), c = (int)(i % p.row_vec);
    const uint4 v = *reinterpret_cast<const uint4*>(p.src + (long long)r * p.src_row_stride + (long long)c * 16);
#pragma unroll 1
    for (int d = 0; d < p.n_dst; ++d) p.dst[d][i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int arrived = atomicAdd(&p.flags[3], 1u);
    if (arrived == gridDim.x - 1) {   // last CTA: every store of this event is visible system-wide
      p.flags[3] = 0u;
      const unsigned int seq = p.seq + *reinterpret_cast<const volatile unsigned int*>(p.flags + PUSH_SEQ_BASE_WORD);
      __threadfence_system();
      for (int d = 0; d < p.n_dst; ++d) st_release_sys(p.dst_flags[d], seq);
    }
  }
}

__global__ void peer_wait_kernel(unsigned int* flags, unsigned int seq_rel) {
  const unsigned int seq = seq_rel + *reinterpret_cast<const volatile unsigned int*>(flags + PUSH_SEQ_BASE_WORD);
  if (*reinterpret_cast<const volatile unsigned int*>(flags + 1) != 0u) return;   // an earlier wait timed out: do not stall again
  bool ok = false;
  for (int round = 0; round < 8 && !ok; ++round) ok = wait_seq(flags, seq);   // 8 x ~4 s: the producer may be capturing its CUDA graph
  if (!ok) flags[1] = 0xDEADu;
}

}  // namespace rtti

using namespace rtti;

extern "C" int rtti_peer_push(const void* src, long long src_row_stride_bytes, int rows, int row_bytes, void* const* dst,
                              void* const* dst_flags, int n_dst, void* flags_local, unsigned int seq, void* stream) {
  if (!src || !dst || !dst_flags || !flags_local || rows < 1 || row_bytes < 16) return RTTI_ERR_ARG;
  if (n_dst < 1 || n_dst > PUSH_MAX_DST) return RTTI_ERR_ARG;
  if (row_bytes % 16 != 0 || src_row_stride_bytes % 16 != 0 || src_row_stride_bytes < row_bytes) return RTTI_ERR_SHAPE;
  if (((uintptr_t)src | (uintptr_t)flags_local) & 15) return RTTI_ERR_ALIGN;
  PeerPushParams p{};
  for (int d = 0; d < n_dst; ++d) {
    if (!dst[d] || !dst_flags[d]) return RTTI_ERR_ARG;
    if (((uintptr_t)dst[d] & 15) || ((uintptr_t)dst_flags[d] & 3)) return RTTI_ERR_ALIGN;
    p.dst[d] = (uint4*)dst[d];
    p.dst_flags[d] = (unsigned int*)dst_flags[d];
  }
  p.n_dst = n_dst;
  p.src = (const uint8_t*)src;
  p.src_row_stride = src_row_stride_bytes;
  p.rows = rows; p.row_vec = row_bytes / 16;
  p.flags = (unsigned int*)flags_local;
  p.seq = seq;
  // a modest grid: the kernel shares the GPU with the producer's own attention kernel and is NVLink-bound
  long long blocks = ((long long)rows * p.row_vec + 256 * 8 - 1) / (256 * 8);
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  peer_push_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? RTTI_OK : RTTI_ERR_CUDA;
}

extern "C" int rtti_peer_wait(void* flags_local, unsigned int seq, void* stream) {
  if (!flags_local) return RTTI_ERR_ARG;
  if ((uintptr_t)flags_local & 3) return RTTI_ERR_ALIGN;
  peer_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned int*)flags_local, seq);
  return cudaGetLastError() == cudaSuccess ? RTTI_OK : RTTI_ERR_CUDA;
}

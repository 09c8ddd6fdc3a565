// Tensor-parallel collectives fused with their consumers, over NVLink peer memory (no NCCL call, no separate reduction kernel).
//
// a row-parallel linear is followed by an all-reduce of the [hidden] partial sums and the residual add; the vocabulary-
// parallel lm_head by an all-gather of (value, index) candidates and the token bookkeeping.  Every rank's partials live in a
// SYMMETRIC buffer (same layout on every GPU, mapped into every peer: torch.distributed._symmetric_memory, i.e. cuMem + NVLink P2P),
// and ONE small kernel per collective does: signal the peers (release store into THEIR flag array), wait for everybody's signal
// (acquire loads of the local flag array), pull the peers' partials through NVLink (16 KB per peer), reduce in rank order
// (bit-identical result on every rank), add the residual / pick the token.  A 16 KB message is pure latency: this costs one NVLink
// round trip instead of an NCCL launch + ring/tree steps + a separate residual kernel.
//
// Buffer layout (symmetric, per rank):  int flags[world][TP_FLAGS]  |  float slots[n_slots][slot_floats]
//   flags[r][i] on rank q = "rank r's contribution to collective i of the current step is complete" (written by rank r).
// Sequence values: (epoch << 20) | (step << 8) | (i + 1); `epoch` changes per request, `step` is the decoder's device-side step
// counter, so a CUDA-graph replay produces fresh values without host involvement.  A slot is rewritten one full step later: by
// then every rank has passed all later collectives of the previous step, which needed the reader's own later contributions.
#include "common.cuh"
#include "srgpt_b200.h"

namespace srgpt {
namespace tp {

constexpr int TP_FLAGS = 128;  // collectives per step (2 x layers + 1 <= 128)
constexpr int MAX_WORLD = 16;

struct Peers {
  unsigned long long base[MAX_WORLD];  // peer-mapped base address of every rank's symmetric buffer (own rank included)
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {  // peer memory: never from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned int seq_value(const int* epoch, const int* step, int idx) {
  const int e = *reinterpret_cast<const volatile int*>(epoch), st = *reinterpret_cast<const volatile int*>(step);  // never from a cached copy
  return ((unsigned int)(e & 0xFFF) << 20) | ((unsigned int)(st & 0xFFF) << 8) | (unsigned int)(idx + 1);
}

// signal + wait: thread 0 of block 0 tells every peer, thread 0 of every block waits for every peer
__device__ __forceinline__ void exchange_flags(const Peers& peers, int rank, int world, int idx, const int* epoch, const int* step) {
  // programmatic dependent launch on both sides: the NEXT kernel (a weight-streaming GEMV) may start now and prefetch its rows
  // during the NVLink round trip below; this kernel itself started early and now waits for the GEMV that wrote the partial sums
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // only NOW are the step / epoch counters final (the previous step's token kernel increments *step; reading them before the wait
  // made an early-started kernel signal a stale sequence value and its peer spin forever: found on the 2-GPU box in eager mode)
  const unsigned int seq = seq_value(epoch, step, idx);
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();  // the partial written by the previous kernel of this stream is visible to the peers before the flag
      for (int r = 0; r < world; ++r)
        st_release_sys(reinterpret_cast<unsigned int*>(peers.base[r]) + rank * TP_FLAGS + idx, seq);
    }
    const unsigned int* mine = reinterpret_cast<const unsigned int*>(peers.base[rank]);
    for (int r = 0; r < world; ++r) {
      unsigned int spins = 0;
      while (ld_acquire_sys(mine + r * TP_FLAGS + idx) != seq) {
        if (++spins > (1u << 28)) __trap();  // a lost peer turns into a launch error, not a hang
      }
    }
  }
  __syncthreads();
}

// h = bf16(bf16(sum over ranks of partial_r) + h): all-reduce + the residual add of modeling_llama.py:668,682 in one kernel
__global__ void __launch_bounds__(256)
allreduce_residual_kernel(const Peers peers, int rank, int world, long long slot_off_bytes, int idx, const int* __restrict__ epoch,
                          const int* __restrict__ step, bf16* __restrict__ h, int n) {
  exchange_flags(peers, rank, world, idx, epoch, step);
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < world; ++r) {  // rank order: the same fp32 sum on every rank
    const float4 v = ld_peer_f4(reinterpret_cast<const float4*>(peers.base[r] + slot_off_bytes) + (i >> 2));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const uint2 hv = *reinterpret_cast<const uint2*>(h + i);
  uint2 o;
  o.x = pack_bf16x2(bf16_round(acc.x) + bf16_lo(hv.x), bf16_round(acc.y) + bf16_hi(hv.x));
  o.y = pack_bf16x2(bf16_round(acc.z) + bf16_lo(hv.y), bf16_round(acc.w) + bf16_hi(hv.y));
  *reinterpret_cast<uint2*>(h + i) = o;
}

// all-gather of the ranks' (best value, global index) + arg max (lowest index on ties) + the bookkeeping of lm_head_finalize_kernel
__global__ void __launch_bounds__(256)
allgather_pick_kernel(const Peers peers, int rank, int world, long long slot_off_bytes, int idx, const int* __restrict__ epoch, const bf16* __restrict__ embed_table,
                      bf16* __restrict__ next_x, int K, long long* __restrict__ out_ids, int* step, int* pos) {
  __shared__ int s_tok;
  exchange_flags(peers, rank, world, idx, epoch, step);
  if (threadIdx.x == 0) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int r = 0; r < world; ++r) {
      const int* cand = reinterpret_cast<const int*>(peers.base[r] + slot_off_bytes);
      int vb, ib;
      asm volatile("ld.relaxed.sys.global.v2.s32 {%0,%1}, [%2];" : "=r"(vb), "=r"(ib) : "l"(cand) : "memory");
      const float v = __int_as_float(vb);
      if (v > bv || (v == bv && ib < bi)) { bv = v; bi = ib; }
    }
    if (bi == 0x7fffffff) bi = 0;
    s_tok = bi;
    out_ids[*step] = (long long)bi;
  }
  __syncthreads();
  const int tok = s_tok;
  if (embed_table != nullptr && next_x != nullptr) {
    const uint4* src = reinterpret_cast<const uint4*>(embed_table + (size_t)tok * K);
    for (int c = threadIdx.x; c < (K >> 3); c += blockDim.x) reinterpret_cast<uint4*>(next_x)[c] = src[c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *step += 1;
    *pos += 1;
  }
}

}  // namespace tp
}  // namespace srgpt

using namespace srgpt;

static int fill_peers(tp::Peers* p, const unsigned long long* peer_bases, int world) {
  if (peer_bases == nullptr || world < 1 || world > tp::MAX_WORLD) return -1;
  for (int r = 0; r < tp::MAX_WORLD; ++r) p->base[r] = r < world ? peer_bases[r] : 0ull;
  return 0;
}

extern "C" __attribute__((visibility("default"))) long long srgpt_tp_comm_bytes(int world, int n_slots, int slot_floats) {
  if (world < 1 || world > tp::MAX_WORLD || n_slots < 1 || slot_floats < 1) return -1;
  return (long long)world * tp::TP_FLAGS * 4 + (long long)n_slots * slot_floats * 4;
}
extern "C" __attribute__((visibility("default"))) long long srgpt_tp_comm_slot_offset(int world, int slot, int slot_floats) {
  if (world < 1 || slot < 0 || slot_floats < 1) return -1;
  return (long long)world * tp::TP_FLAGS * 4 + (long long)slot * slot_floats * 4;
}

// peer_bases: HOST array [world] of the peer-mapped base addresses of the symmetric buffer (own rank included); collective `idx`
// (< 128) of the step uses slot `slot_off_bytes` (srgpt_tp_comm_slot_offset) in every rank's buffer.
extern "C" __attribute__((visibility("default"))) int srgpt_tp_allreduce_residual_bf16(const unsigned long long* peer_bases, int rank, int world, long long slot_off_bytes,
                                                                                       int idx, const int* epoch, const int* step, void* h, int n, void* stream) {
  tp::Peers peers;
  SRGPT_CHECK_ARG(fill_peers(&peers, peer_bases, world) == 0 && rank >= 0 && rank < world && idx >= 0 && idx < tp::TP_FLAGS);
  SRGPT_CHECK_ARG(epoch && step && h && n > 0 && (n % 4) == 0 && (slot_off_bytes % 16) == 0 && (reinterpret_cast<uintptr_t>(h) & 7) == 0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ceil_div(n / 4, 256));
  cfg.blockDim = dim3(256);
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tp::allreduce_residual_kernel, peers, rank, world, slot_off_bytes, idx, epoch, step, reinterpret_cast<bf16*>(h), n));
  return SRGPT_OK;
}

extern "C" __attribute__((visibility("default"))) int srgpt_tp_allgather_pick_token(const unsigned long long* peer_bases, int rank, int world, long long slot_off_bytes,
                                                                                    int idx, const int* epoch, const void* embed_table, void* next_x, int K,
                                                                                    long long* out_ids, int* step, int* pos, void* stream) {
  tp::Peers peers;
  SRGPT_CHECK_ARG(fill_peers(&peers, peer_bases, world) == 0 && rank >= 0 && rank < world && idx >= 0 && idx < tp::TP_FLAGS);
  SRGPT_CHECK_ARG(epoch && out_ids && step && pos && (slot_off_bytes % 8) == 0);
  SRGPT_CHECK_ARG((embed_table == nullptr) == (next_x == nullptr));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(256);
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tp::allgather_pick_kernel, peers, rank, world, slot_off_bytes, idx, epoch, reinterpret_cast<const bf16*>(embed_table),
                                      reinterpret_cast<bf16*>(next_x), K, out_ids, step, pos));
  return SRGPT_OK;
}

// Decode-time weight-streaming kernels (one new token): y = W · x with W [N, K] bf16 read exactly
// once from HBM.  At batch 1 the whole Llama decode step is bound by these reads (15 GB / token for
// Llama-3-8B), so this kernel is written against the HBM roofline.
//
// Design (measured history in profiles/: v1 persistent warps 0.70, v2 bulk-copy rings + CTA-level K split
// 0.48, this version — see DESIGN.md "decode GEMV"):
//   * one warp = one PAIR of weight rows, read with 8 independent 16-byte streaming loads in flight per
//     lane (ld.global.nc.L1::no_allocate).  A warp is latency-bound by design (4 KB in flight), the chip
//     is saturated by having >= 2000 warps resident; CTAs are NOT persistent, so the hardware scheduler
//     balances the row pairs dynamically and no warp ever owns 2 pairs while another owns 1;
//   * the first 8 loads of every warp are issued BEFORE anything that depends on the previous kernel;
//     programmatic dependent launch (griddepcontrol.launch_dependents / .wait) lets the next kernel's
//     CTAs take the SM slots the current kernel frees in its tail, so launch latency, the first HBM
//     round trip and the RMSNorm prologue overlap the previous kernel instead of adding ~6 us each;
//   * x (RMS-normalised in the prologue when requested) is staged once per CTA in shared memory;
//   * the pair is chosen so the fused epilogue is warp-local: (gate_i, up_i) for SwiGLU, (d, d + hd/2)
//     of one head for RoPE + KV-cache append, two vocabulary rows for lm_head + argmax.
// Rounding points follow the reference's bf16 torch ops (modeling_llama.py:429-431,186-191,221,668,682).
#include <stdlib.h>

#include "common.cuh"
#include "srgpt_b200.h"

namespace srgpt {
namespace gemv {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int SPRE_MAX = 2;                       // up to 2 x 4 chunks per row per lane in shared memory
constexpr int SPRE_WARP_BYTES = 2 * 4 * 32 * 16;  // per batch: 2 rows x 4 chunks x 32 lanes x 16 B = 4 KB
enum { MODE_LM = 3 };

struct Params {
  const bf16* x;
  const bf16* W;
  int ldw;
  bf16* y;
  int N, K;
  const bf16* norm_weight;
  float eps;
  const bf16* residual;
  // QKV_ROPE
  int n_heads, n_kv_heads, hd;
  const bf16* cos_tab;
  const bf16* sin_tab;
  const int* pos;
  bf16* kv_pages;
  const int* page_table;
  int page_size;
  // LM
  float* logits_out;
  float* part_val;
  int* part_idx;
  unsigned long long* trace;  // optional timeline record (srgpt_trace_begin)
  int spre;                   // number of extra 4-chunk batches per row staged in shared memory before the wait (0..SPRE_MAX)
  int l2pf;                   // 1: the rest of the warp's two rows is requested into L2 (bulk prefetch) before the dependency wait
  // tensor parallelism (srgpt_gemv_tp_bf16): this rank's slice of the heads / of the reduction dimension
  int kv_heads_total;         // KV-cache row = kv_heads_total * hd elements (0: n_kv_heads); the rank writes heads [kv_head_off, +n_kv_heads)
  int kv_head_off;
  float* y_f32;               // PLAIN mode: un-rounded fp32 partial dot products go here instead of bf16 y (+ residual)
};

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// stage x (optionally RMS-normalised) into shared memory as bf16.
// Fast path (K <= 2 * THREADS * 8 = 4096, the two RMSNorm-fused kernels of a layer): every thread keeps its <= 2
// chunks of x in registers between the sum of squares and the scaling, and the (static) norm weights `nw` were
// fetched before the dependency wait -> one L2 round trip on the critical path instead of three.
__device__ __forceinline__ void stage_x(const bf16* __restrict__ x, const bf16* __restrict__ norm_weight, float eps, int K,
                                        bf16* sx, float* red, const uint4* nw_pre, bool nw_pre_valid) {
  const int nchunk = K >> 3;
  if (norm_weight == nullptr) {
    for (int c = threadIdx.x; c < nchunk; c += THREADS)
      reinterpret_cast<uint4*>(sx)[c] = reinterpret_cast<const uint4*>(x)[c];
  } else if (nw_pre_valid) {
    uint4 xr[2];
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = threadIdx.x + k * THREADS;
      xr[k] = make_uint4(0, 0, 0, 0);
      if (c < nchunk) xr[k] = reinterpret_cast<const uint4*>(x)[c];
      float f[8];
      unpack8(xr[k], f);
#pragma unroll
      for (int t = 0; t < 8; ++t) sq += f[t] * f[t];
    }
    const float rstd = rsqrtf(block_sum(sq, red) / (float)K + eps);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = threadIdx.x + k * THREADS;
      if (c < nchunk) {
        float f[8], w[8], o[8];
        unpack8(xr[k], f);
        unpack8(nw_pre[k], w);
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = w[t] * bf16_round(f[t] * rstd);
        reinterpret_cast<uint4*>(sx)[c] = pack8(o);
      }
    }
  } else {
    float sq = 0.f;
    for (int c = threadIdx.x; c < nchunk; c += THREADS) {
      float f[8];
      unpack8(reinterpret_cast<const uint4*>(x)[c], f);
#pragma unroll
      for (int t = 0; t < 8; ++t) sq += f[t] * f[t];
    }
    const float rstd = rsqrtf(block_sum(sq, red) / (float)K + eps);
    for (int c = threadIdx.x; c < nchunk; c += THREADS) {
      float f[8], w[8], o[8];
      unpack8(reinterpret_cast<const uint4*>(x)[c], f);
      unpack8(reinterpret_cast<const uint4*>(norm_weight)[c], w);
#pragma unroll
      for (int t = 0; t < 8; ++t) o[t] = w[t] * bf16_round(f[t] * rstd);
      reinterpret_cast<uint4*>(sx)[c] = pack8(o);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float dot8(const uint4& w, const float* xf) {
  float f[8];
  unpack8(w, f);
  float a = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) a = fmaf(f[t], xf[t], a);
  return a;
}

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

template <int MODE>
__device__ __forceinline__ void pair_rows(const Params& p, int pi, int& r0, int& r1) {
  if (MODE == SRGPT_GEMV_QKV_ROPE) {
    const int half = p.hd >> 1;
    const int head = pi / half, j = pi - head * half;
    r0 = head * p.hd + j;
    r1 = r0 + half;
  } else {
    r0 = 2 * pi;
    r1 = r0 + 1;
    if (MODE == MODE_LM && r1 >= p.N) r1 = r0;  // odd vocabulary: the last pair streams row r0 twice
  }
}

// PRE = number of 4-chunk batches (per row) requested BEFORE the dependency wait: 1 -> 8 loads per lane (3 CTAs/SM),
// 2 -> 16 loads (2 CTAs/SM), 4 -> 32 loads = a whole K=4096 row pair per warp (1 CTA/SM).  The small matrices of a
// layer (qkv 50 MB, o_proj 33 MB) run behind a kernel that leaves HBM idle (decode attention / the previous tail), so
// the more of their weights is in flight before the dependency resolves, the less of them is exposed afterwards.
template <int MODE, int PRE>
__global__ void __launch_bounds__(THREADS, PRE == 1 ? 3 : (PRE == 2 ? 2 : 1)) decode_gemv_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ float red[32];
  __shared__ float sv[WARPS];
  __shared__ int si[WARPS];
  bf16* sx = reinterpret_cast<bf16*>(smem_raw);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  trace_mark(p.trace, 0);
  const int npairs = (MODE == MODE_LM) ? ((p.N + 1) >> 1) : (p.N >> 1);
  const int pi = blockIdx.x * WARPS + warp;
  const bool active = pi < npairs;
  int r0 = 0, r1 = 0;
  if (active) pair_rows<MODE>(p, pi, r0, r1);
  const uint4* p0 = reinterpret_cast<const uint4*>(p.W + (size_t)r0 * p.ldw);
  const uint4* p1 = reinterpret_cast<const uint4*>(p.W + (size_t)r1 * p.ldw);
  const uint4* px = reinterpret_cast<const uint4*>(sx);
  const int nchunk = p.K >> 3;
  const int nchunk_all = nchunk;

  // ---- first 8*PRE loads per lane: weights do not depend on the previous kernel
  constexpr int NPRE = 4 * PRE;
  uint4 u0[NPRE], u1[NPRE];
  int c = lane;
  const bool first_full = active && (c + 32 * (NPRE - 1) < nchunk);
  if (first_full) {
#pragma unroll
    for (int i = 0; i < NPRE; ++i) {
      u0[i] = ld_stream16(p0 + c + 32 * i);
      u1[i] = ld_stream16(p1 + c + 32 * i);
    }
  }
  // ---- a second, register-free prefetch level: the next p.spre batches of both rows go to shared memory with
  //      cp.async (every lane later reads back exactly the 16-byte slots it filled, so no barrier is needed).
  //      Occupancy stays at 3 CTAs/SM, unlike the deeper register prefetch (PRE = 2/4) that was measured and rejected.
  uint8_t* spre_base = smem_raw + (size_t)p.K * 2 + (size_t)warp * ((size_t)p.spre * SPRE_WARP_BYTES);
  int n_spre = 0;
  if (first_full) {
    for (int b = 0; b < p.spre; ++b) {
      const int cb = c + 32 * NPRE + 128 * b;
      if (cb + 96 >= nchunk) break;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(spre_base + ((b * 2 + 0) * 4 + i) * 512 + lane * 16);
        const uint32_t d1 = (uint32_t)__cvta_generic_to_shared(spre_base + ((b * 2 + 1) * 4 + i) * 512 + lane * 16);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(p0 + cb + 32 * i) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(p1 + cb + 32 * i) : "memory");
      }
      ++n_spre;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // ---- third level, no registers and no shared memory: everything of the two rows that the levels above did not request goes
  //      to L2 with one bulk prefetch per row (cp.async.bulk.prefetch.L2, SASS UBLKPF).  A kernel whose CTAs are resident while
  //      its producer still runs (o_proj under the decode attention, gate/up under o_proj's tail, every kernel across the
  //      ~1 us dependency release) then keeps HBM streaming through what used to be idle gaps and later reads L2 hits.
  //      MEASURED SLOWER on the in-graph timeline (profiles/r02_ab_decode_l2pf.txt: step 2.74 -> 2.93 ms; gate/up 35.4 -> 37.3 us,
  //      down 20.4 -> 22.0 us, lm_head 146 -> 156 us exposed): the prefetch and the demand loads of the same rows race and both
  //      reach DRAM.  Off by default; SRGPT_GEMV_L2PF=1 keeps the experiment reproducible.
  if (p.l2pf && active && lane < 2) {
    const int c_req = first_full ? (32 * NPRE + 128 * n_spre) : 0;  // chunks per row already requested above
    if (c_req < nchunk) {
      const char* rowp = reinterpret_cast<const char*>(lane == 0 ? p0 : p1) + (size_t)c_req * 16;
      uint32_t bytes = (uint32_t)(nchunk - c_req) * 16u;
      while (bytes > 0) {  // pieces of <= 16 KB
        const uint32_t n = bytes > 16384u ? 16384u : bytes;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rowp), "r"(n) : "memory");
        rowp += n;
        bytes -= n;
      }
    }
  }
  // norm weights are static too: fetch them before the wait when a thread owns at most 2 chunks of x
  uint4 nw_pre[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
  const bool nw_pre_valid = (p.norm_weight != nullptr) && (nchunk_all <= 2 * THREADS);
  if (nw_pre_valid) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int cc = threadIdx.x + k * THREADS;
      if (cc < nchunk_all) nw_pre[k] = reinterpret_cast<const uint4*>(p.norm_weight)[cc];
    }
  }
  pdl_launch_dependents();
  pdl_wait();  // activations written by earlier kernels are visible from here on
  trace_mark(p.trace, 1);

  stage_x(p.x, p.norm_weight, p.eps, p.K, sx, red, nw_pre, nw_pre_valid);

  float a0 = 0.f, a1 = 0.f;
  if (active) {
    if (first_full) {
#pragma unroll
      for (int i = 0; i < NPRE; ++i) {
        float xf[8];
        unpack8(px[c + 32 * i], xf);
        a0 += dot8(u0[i], xf);
        a1 += dot8(u1[i], xf);
      }
      c += 32 * NPRE;
      if (n_spre > 0) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        for (int b = 0; b < n_spre; ++b) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float xf[8];
            unpack8(px[c + 32 * i], xf);
            const uint4 w0 = *reinterpret_cast<const uint4*>(spre_base + ((b * 2 + 0) * 4 + i) * 512 + lane * 16);
            const uint4 w1 = *reinterpret_cast<const uint4*>(spre_base + ((b * 2 + 1) * 4 + i) * 512 + lane * 16);
            a0 += dot8(w0, xf);
            a1 += dot8(w1, xf);
          }
          c += 128;
        }
      }
    }
    for (; c + 96 < nchunk; c += 128) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        u0[i] = ld_stream16(p0 + c + 32 * i);
        u1[i] = ld_stream16(p1 + c + 32 * i);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float xf[8];
        unpack8(px[c + 32 * i], xf);
        a0 += dot8(u0[i], xf);
        a1 += dot8(u1[i], xf);
      }
    }
    for (; c < nchunk; c += 32) {
      float xf[8];
      unpack8(px[c], xf);
      a0 += dot8(ld_stream16(p0 + c), xf);
      a1 += dot8(ld_stream16(p1 + c), xf);
    }
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  float best = -INFINITY;
  int besti = 0x7fffffff;
  if (active && lane == 0) {
    if (MODE == SRGPT_GEMV_PLAIN && p.y_f32 != nullptr) {
      // row-parallel linear of a tensor-parallel rank: the partial sums over this rank's K slice, reduced across ranks afterwards
      *reinterpret_cast<float2*>(p.y_f32 + r0) = make_float2(a0, a1);
    } else if (MODE == SRGPT_GEMV_PLAIN) {
      float y0 = bf16_round(a0), y1 = bf16_round(a1);
      if (p.residual != nullptr) {
        y0 += e2f(p.residual[r0]);
        y1 += e2f(p.residual[r1]);
      }
      *reinterpret_cast<uint32_t*>(p.y + r0) = pack_bf16x2(y0, y1);
    } else if (MODE == SRGPT_GEMV_SWIGLU) {
      const float g = bf16_round(a0), u = bf16_round(a1);
      p.y[pi] = f2e(bf16_round(silu(g)) * u);
    } else if (MODE == SRGPT_GEMV_QKV_ROPE) {
      const int half = p.hd >> 1;
      const int head = pi / half, j = pi - head * half;
      float v0 = bf16_round(a0), v1 = bf16_round(a1);
      const int pos = *p.pos;
      if (head < p.n_heads + p.n_kv_heads) {
        const float cs = e2f(p.cos_tab[(size_t)pos * half + j]);
        const float sn = e2f(p.sin_tab[(size_t)pos * half + j]);
        const float o0 = bf16_round(bf16_round(v0 * cs) + bf16_round(-v1 * sn));
        const float o1 = bf16_round(bf16_round(v1 * cs) + bf16_round(v0 * sn));
        v0 = o0;
        v1 = o1;
      }
      if (head < p.n_heads) {
        p.y[r0] = f2e(v0);
        p.y[r1] = f2e(v1);
      } else {
        const int page = p.page_table[pos / p.page_size], slot = pos % p.page_size;
        const bool is_v = head >= p.n_heads + p.n_kv_heads;
        const int kh = head - p.n_heads - (is_v ? p.n_kv_heads : 0) + p.kv_head_off;
        const int kv_row = (p.kv_heads_total > 0 ? p.kv_heads_total : p.n_kv_heads) * p.hd;
        bf16* dst = p.kv_pages + (((size_t)page * 2 + (is_v ? 1 : 0)) * p.page_size + slot) * kv_row + kh * p.hd;
        dst[j] = f2e(v0);
        dst[j + half] = f2e(v1);
      }
    } else {  // MODE_LM: logits = lm_head(h).float() -> bf16 rounding first (modeling_llama.py:1044-1045)
      a0 = bf16_round(a0);
      a1 = bf16_round(a1);
      if (p.logits_out != nullptr) {
        p.logits_out[r0] = a0;
        if (r1 != r0) p.logits_out[r1] = a1;
      }
      best = a0;
      besti = r0;
      if (r1 != r0 && better(a1, r1, best, besti)) { best = a1; besti = r1; }
    }
  }
  if (MODE == MODE_LM) {
    if (lane == 0) { sv[warp] = best; si[warp] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < WARPS; ++w)
        if (better(sv[w], si[w], best, besti)) { best = sv[w]; besti = si[w]; }
      p.part_val[blockIdx.x] = best;
      p.part_idx[blockIdx.x] = besti;
    }
  }
  trace_mark(p.trace, 2);
}

__global__ void __launch_bounds__(256)
lm_head_finalize_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int nparts,
                        const bf16* __restrict__ embed_table, bf16* __restrict__ next_x, int K, long long* __restrict__ out_ids,
                        int* step, int* pos) {
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ int s_tok;
  pdl_launch_dependents();
  pdl_wait();
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x)
    if (better(part_val[i], part_idx[i], best, bi)) { best = part_val[i]; bi = part_idx[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (better(sv[w], si[w], best, bi)) { best = sv[w]; bi = si[w]; }
    if (bi == 0x7fffffff) bi = 0;  // all-NaN logits: never feed the sentinel to the embedding gather of the next step
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

// ---- tensor-parallel helpers ------------------------------------------------------------------------
// vocabulary-parallel lm_head: this rank's best (bf16-rounded logit, GLOBAL row index) -> best[0] = value bits, best[1] = index
__global__ void __launch_bounds__(256)
lm_head_local_best_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int nparts, int index_base, int* __restrict__ best) {
  __shared__ float sv[8];
  __shared__ int si[8];
  pdl_launch_dependents();
  pdl_wait();
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x)
    if (better(part_val[i], part_idx[i], bv, bi)) { bv = part_val[i]; bi = part_idx[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
    best[0] = __float_as_int(bv);
    best[1] = (bi == 0x7fffffff) ? index_base : bi + index_base;
  }
}

// after the all-gather of every rank's (value, index): the global arg max (lowest index on ties, like torch.argmax), then the same
// bookkeeping as lm_head_finalize_kernel (token id, next embedding row, ++step, ++pos); identical on every rank
__global__ void __launch_bounds__(256)
tp_pick_token_kernel(const int* __restrict__ best_all, int world, const bf16* __restrict__ embed_table, bf16* __restrict__ next_x, int K,
                     long long* __restrict__ out_ids, int* step, int* pos) {
  __shared__ int s_tok;
  if (threadIdx.x == 0) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int r = 0; r < world; ++r) {
      const float v = __int_as_float(best_all[2 * r]);
      const int i = best_all[2 * r + 1];
      if (better(v, i, bv, bi)) { bv = v; bi = i; }
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

// h = bf16(bf16(sum of the ranks' partial dot products) + h): the rounding points of `residual + o_proj(x)` (modeling_llama.py:668,682)
__global__ void __launch_bounds__(256) tp_residual_add_kernel(bf16* __restrict__ h, const float* __restrict__ partial, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) h[i] = f2e(bf16_round(partial[i]) + e2f(h[i]));
}

// ---- host side ------------------------------------------------------------------------------------
static int grid_for(int npairs) { return ceil_div(npairs, WARPS); }

static void pdl_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int block, int smem, cudaStream_t st) {
  cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
}

static int spre_default() {
  static const int v = [] {
    const char* e = getenv("SRGPT_GEMV_SPRE");
    const int x = (e != nullptr && e[0] != 0) ? atoi(e) : 1;
    return x < 0 ? 0 : (x > SPRE_MAX ? SPRE_MAX : x);
  }();
  return v;
}

template <int MODE, int PRE>
static int launch_pre(const Params& p, int npairs, cudaStream_t st) {
  const int smem = p.K * 2 + WARPS * spre_default() * SPRE_WARP_BYTES;
  static int configured_smem = 0;
  if (smem > 48 * 1024 && smem > configured_smem) {
    SRGPT_CHECK_CUDA(cudaFuncSetAttribute(decode_gemv_kernel<MODE, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  pdl_config(cfg, attr, grid_for(npairs), THREADS, smem, st);
  Params q = p;
  q.trace = trace_next_slot();
  q.spre = spre_default();
  static const int l2pf = [] {
    const char* v = getenv("SRGPT_GEMV_L2PF");
    return (v != nullptr && v[0] != 0) ? atoi(v) : 0;
  }();
  // 1: every GEMV (measured slower, see the kernel); 2: only o_proj, whose CTAs are resident while the latency-bound decode
  // attention leaves HBM idle (its 33 MB fit L2 many times over); 3: o_proj and the qkv GEMV
  const bool small_plain = MODE == SRGPT_GEMV_PLAIN && p.residual != nullptr && (long long)p.N * p.K <= (32LL << 20);
  q.l2pf = (l2pf == 1 || (l2pf >= 2 && small_plain) || (l2pf == 3 && MODE == SRGPT_GEMV_QKV_ROPE)) ? 1 : 0;
  SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_gemv_kernel<MODE, PRE>, q));
  return SRGPT_OK;
}

// Prefetch depth.  Measured on the in-graph timeline (profiles/r01_decode_trace_pre{1,2,4}.txt): deeper prefetch for the
// small matrices (PRE 2 / 4) trims o_proj by ~0.9 us but costs occupancy (2 / 1 CTAs per SM), so the NEXT kernel can no
// longer co-reside and start early: the step gets slower (2.732 / 2.746 / 2.833 ms).  Default 1; SRGPT_GEMV_PRE_SMALL
// keeps the experiment reproducible.
template <int MODE>
static int launch(const Params& p, int npairs, cudaStream_t st) {
  static const int pre_small = [] {
    const char* v = getenv("SRGPT_GEMV_PRE_SMALL");
    const int x = (v != nullptr && v[0] != 0) ? atoi(v) : 1;
    return (x == 1 || x == 2 || x == 4) ? x : 1;
  }();
  const bool small = (MODE == SRGPT_GEMV_PLAIN || MODE == SRGPT_GEMV_QKV_ROPE) && ((size_t)p.N * p.K * 2 <= (size_t)64 << 20);
  if (small && pre_small == 4 && (p.K >> 3) >= 512) return launch_pre<MODE, 4>(p, npairs, st);
  if (small && pre_small >= 2 && (p.K >> 3) >= 256) return launch_pre<MODE, 2>(p, npairs, st);
  return launch_pre<MODE, 1>(p, npairs, st);
}

}  // namespace gemv
}  // namespace srgpt

using namespace srgpt;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" __attribute__((visibility("default"))) int srgpt_gemv_bf16(const void* x, const void* W, int ldw, void* y, int N, int K, const void* norm_weight, float eps,
                               const void* residual, int mode, int n_heads, int n_kv_heads, int head_dim, const void* cos_tab,
                               const void* sin_tab, const int* pos, void* kv_pages, const int* page_table, int page_size,
                               void* stream) {
  SRGPT_CHECK_ARG(x && W && y && N > 0 && K > 0);
  SRGPT_CHECK_ARG((N % 2) == 0 && (K % 8) == 0 && (ldw % 8) == 0 && ldw >= K);
  SRGPT_CHECK_ARG(K * 2 <= 200 * 1024);
  SRGPT_CHECK_ARG(aligned16(x) && aligned16(W) && (reinterpret_cast<uintptr_t>(y) & 3) == 0);
  SRGPT_CHECK_ARG(norm_weight == nullptr || aligned16(norm_weight));
  SRGPT_CHECK_ARG(mode >= SRGPT_GEMV_PLAIN && mode <= SRGPT_GEMV_QKV_ROPE);
  SRGPT_CHECK_ARG(x != y);  // x is read by late CTAs while early ones already write y
  gemv::Params p = {};
  p.x = reinterpret_cast<const bf16*>(x);
  p.W = reinterpret_cast<const bf16*>(W);
  p.ldw = ldw;
  p.y = reinterpret_cast<bf16*>(y);
  p.N = N; p.K = K;
  p.norm_weight = reinterpret_cast<const bf16*>(norm_weight);
  p.eps = eps;
  p.residual = reinterpret_cast<const bf16*>(residual);
  p.n_heads = n_heads; p.n_kv_heads = n_kv_heads; p.hd = head_dim;
  p.cos_tab = reinterpret_cast<const bf16*>(cos_tab);
  p.sin_tab = reinterpret_cast<const bf16*>(sin_tab);
  p.pos = pos;
  p.kv_pages = reinterpret_cast<bf16*>(kv_pages);
  p.page_table = page_table;
  p.page_size = page_size;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (mode) {
    case SRGPT_GEMV_PLAIN:
      p.hd = 2;
      return gemv::launch<SRGPT_GEMV_PLAIN>(p, N / 2, st);
    case SRGPT_GEMV_SWIGLU:
      SRGPT_CHECK_ARG(residual == nullptr);
      p.hd = 2;
      return gemv::launch<SRGPT_GEMV_SWIGLU>(p, N / 2, st);
    case SRGPT_GEMV_QKV_ROPE:
      SRGPT_CHECK_ARG(residual == nullptr && n_heads > 0 && n_kv_heads > 0 && head_dim > 0 && (head_dim % 2) == 0);
      SRGPT_CHECK_ARG(N == (n_heads + 2 * n_kv_heads) * head_dim);
      SRGPT_CHECK_ARG(cos_tab && sin_tab && pos && kv_pages && page_table && page_size > 0);
      return gemv::launch<SRGPT_GEMV_QKV_ROPE>(p, N / 2, st);
  }
  return SRGPT_ERR_INVALID;
}

// Tensor-parallel variants of the decode GEMV (SURVEY.md §8e "optional TP", BASELINE config c5): a rank owns n_heads q heads and
// n_kv_heads kv heads of the fused qkv projection (column parallel; K/V rows land in the FULL-layout cache at kv_head_off), and a K
// slice of o_proj / down_proj (row parallel): PLAIN mode with partial_f32 != NULL writes the un-rounded fp32 partial sums that the
// ranks then all-reduce.  Everything else is srgpt_gemv_bf16.
extern "C" __attribute__((visibility("default"))) int srgpt_gemv_tp_bf16(const void* x, const void* W, int ldw, void* y, int N, int K, const void* norm_weight, float eps,
                                                                         int mode, int n_heads, int n_kv_heads, int head_dim, const void* cos_tab, const void* sin_tab,
                                                                         const int* pos, void* kv_pages, const int* page_table, int page_size, int kv_heads_total,
                                                                         int kv_head_off, float* partial_f32, void* stream) {
  SRGPT_CHECK_ARG(x && W && N > 0 && K > 0 && (y != nullptr || partial_f32 != nullptr));
  SRGPT_CHECK_ARG((N % 2) == 0 && (K % 8) == 0 && (ldw % 8) == 0 && ldw >= K && K * 2 <= 200 * 1024);
  SRGPT_CHECK_ARG(aligned16(x) && aligned16(W) && (norm_weight == nullptr || aligned16(norm_weight)));
  SRGPT_CHECK_ARG(mode == SRGPT_GEMV_PLAIN || mode == SRGPT_GEMV_QKV_ROPE);
  gemv::Params p = {};
  p.x = reinterpret_cast<const bf16*>(x);
  p.W = reinterpret_cast<const bf16*>(W);
  p.ldw = ldw;
  p.y = reinterpret_cast<bf16*>(y);
  p.N = N; p.K = K;
  p.norm_weight = reinterpret_cast<const bf16*>(norm_weight);
  p.eps = eps;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (mode == SRGPT_GEMV_PLAIN) {
    SRGPT_CHECK_ARG(partial_f32 != nullptr && (reinterpret_cast<uintptr_t>(partial_f32) & 7) == 0);
    p.hd = 2;
    p.y_f32 = partial_f32;
    return gemv::launch<SRGPT_GEMV_PLAIN>(p, N / 2, st);
  }
  SRGPT_CHECK_ARG(y != nullptr && (reinterpret_cast<uintptr_t>(y) & 3) == 0 && x != y);
  SRGPT_CHECK_ARG(n_heads > 0 && n_kv_heads > 0 && head_dim > 0 && (head_dim % 2) == 0 && N == (n_heads + 2 * n_kv_heads) * head_dim);
  SRGPT_CHECK_ARG(cos_tab && sin_tab && pos && kv_pages && page_table && page_size > 0);
  SRGPT_CHECK_ARG(kv_heads_total >= n_kv_heads && kv_head_off >= 0 && kv_head_off + n_kv_heads <= kv_heads_total);
  p.n_heads = n_heads; p.n_kv_heads = n_kv_heads; p.hd = head_dim;
  p.cos_tab = reinterpret_cast<const bf16*>(cos_tab);
  p.sin_tab = reinterpret_cast<const bf16*>(sin_tab);
  p.pos = pos;
  p.kv_pages = reinterpret_cast<bf16*>(kv_pages);
  p.page_table = page_table;
  p.page_size = page_size;
  p.kv_heads_total = kv_heads_total;
  p.kv_head_off = kv_head_off;
  return gemv::launch<SRGPT_GEMV_QKV_ROPE>(p, N / 2, st);
}

extern "C" __attribute__((visibility("default"))) int srgpt_tp_residual_add_bf16(void* h, const float* partial, int n, void* stream) {
  SRGPT_CHECK_ARG(h && partial && n > 0);
  gemv::tp_residual_add_kernel<<<ceil_div(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<bf16*>(h), partial, n);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
}

// Vocabulary-parallel lm_head of one rank: rows [index_base, index_base + V_local) of the table.  best = device int[2]
// {bf16-rounded best logit (float bits), its GLOBAL row index}; all-gather the pairs, then srgpt_tp_pick_token.
extern "C" __attribute__((visibility("default"))) int srgpt_lm_head_local_best_bf16(const void* x, const void* W_local, int ldw, int V_local, int K, const void* norm_weight,
                                                                                    float eps, void* workspace, int index_base, int* best, void* stream) {
  SRGPT_CHECK_ARG(x && W_local && workspace && best && V_local > 0 && K > 0 && index_base >= 0);
  SRGPT_CHECK_ARG((K % 8) == 0 && (ldw % 8) == 0 && ldw >= K && K * 2 <= 200 * 1024);
  SRGPT_CHECK_ARG(aligned16(x) && aligned16(W_local) && (norm_weight == nullptr || aligned16(norm_weight)));
  const int npairs = (V_local + 1) / 2;
  const int g = gemv::grid_for(npairs);
  gemv::Params p = {};
  p.x = reinterpret_cast<const bf16*>(x);
  p.W = reinterpret_cast<const bf16*>(W_local);
  p.ldw = ldw; p.N = V_local; p.K = K;
  p.norm_weight = reinterpret_cast<const bf16*>(norm_weight);
  p.eps = eps;
  p.hd = 2;
  p.part_val = reinterpret_cast<float*>(workspace);
  p.part_idx = reinterpret_cast<int*>(p.part_val + g);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = gemv::launch<gemv::MODE_LM>(p, npairs, st);
  if (rc != SRGPT_OK) return rc;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  gemv::pdl_config(cfg, attr, 1, 256, 0, st);
  SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemv::lm_head_local_best_kernel, (const float*)p.part_val, (const int*)p.part_idx, g, index_base, best));
  return SRGPT_OK;
}

extern "C" __attribute__((visibility("default"))) int srgpt_tp_pick_token(const int* best_all, int world, const void* embed_table, void* next_x, int K, long long* out_ids,
                                                                          int* step, int* pos, void* stream) {
  SRGPT_CHECK_ARG(best_all && world > 0 && out_ids && step && pos);
  SRGPT_CHECK_ARG((embed_table == nullptr) == (next_x == nullptr));
  SRGPT_CHECK_ARG(embed_table == nullptr || ((K % 8) == 0 && aligned16(embed_table) && aligned16(next_x)));
  gemv::tp_pick_token_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(best_all, world, reinterpret_cast<const bf16*>(embed_table),
                                                                                  reinterpret_cast<bf16*>(next_x), K, out_ids, step, pos);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
}

extern "C" __attribute__((visibility("default"))) long long srgpt_lm_head_workspace(int V) {
  if (V <= 0) return -1;
  const int g = gemv::grid_for((V + 1) / 2);
  return (long long)g * (long long)(sizeof(float) + sizeof(int));
}

extern "C" __attribute__((visibility("default"))) int srgpt_lm_head_argmax_bf16(const void* x, const void* W, int ldw, int V, int K, const void* norm_weight, float eps,
                                         float* logits_out, void* workspace, const void* embed_table, void* next_x,
                                         long long* out_ids, int* step, int* pos, void* stream) {
  SRGPT_CHECK_ARG(x && W && workspace && out_ids && step && pos && V > 0 && K > 0);
  SRGPT_CHECK_ARG((K % 8) == 0 && (ldw % 8) == 0 && ldw >= K && K * 2 <= 200 * 1024);
  SRGPT_CHECK_ARG(aligned16(x) && aligned16(W) && (norm_weight == nullptr || aligned16(norm_weight)));
  SRGPT_CHECK_ARG((embed_table == nullptr) == (next_x == nullptr));
  SRGPT_CHECK_ARG(embed_table == nullptr || (aligned16(embed_table) && aligned16(next_x)));
  const int npairs = (V + 1) / 2;
  const int g = gemv::grid_for(npairs);
  gemv::Params p = {};
  p.x = reinterpret_cast<const bf16*>(x);
  p.W = reinterpret_cast<const bf16*>(W);
  p.ldw = ldw; p.N = V; p.K = K;
  p.norm_weight = reinterpret_cast<const bf16*>(norm_weight);
  p.eps = eps;
  p.hd = 2;
  p.logits_out = logits_out;
  p.part_val = reinterpret_cast<float*>(workspace);
  p.part_idx = reinterpret_cast<int*>(p.part_val + g);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = gemv::launch<gemv::MODE_LM>(p, npairs, st);
  if (rc != SRGPT_OK) return rc;
  // finalize: also a programmatic dependent (its launch latency hides behind the lm_head kernel)
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  gemv::pdl_config(cfg, attr, 1, 256, 0, st);
  SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemv::lm_head_finalize_kernel, (const float*)p.part_val, (const int*)p.part_idx, g,
                                      reinterpret_cast<const bf16*>(embed_table), reinterpret_cast<bf16*>(next_x), K, out_ids, step, pos));
  return SRGPT_OK;
}

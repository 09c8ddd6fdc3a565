// Decode-time weight-streaming kernels (one new token): y = W · x with W [N, K] bf16 read exactly
// once from HBM.  At batch 1 the whole Llama decode step is bound by these reads (15 GB / token for
// Llama-3-8B), so these kernels are written against the HBM roofline:
//   * persistent grid (2 CTAs / SM), x staged once per CTA in shared memory (with the fused
//     LlamaRMSNorm prologue when requested);
//   * each warp owns a PAIR of weight rows and keeps 8 independent 16-byte streaming loads in flight
//     per lane (ld.global.nc.L1::no_allocate);
//   * the pair is chosen so the fused epilogue is warp-local: (gate_i, up_i) for SwiGLU,
//     (d, d + hd/2) of one head for RoPE.
// Rounding points follow the reference's bf16 torch ops (modeling_llama.py:429-431,186-191,221,668,682).
#include "common.cuh"
#include "srgpt_b200.h"

namespace srgpt {
namespace gemv {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

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
};

// stage x (optionally RMS-normalised) into shared memory as bf16
__device__ __forceinline__ void stage_x(const bf16* __restrict__ x, const bf16* __restrict__ norm_weight, float eps, int K,
                                        bf16* sx, float* red) {
  const int nchunk = K >> 3;
  if (norm_weight == nullptr) {
    for (int c = threadIdx.x; c < nchunk; c += THREADS)
      reinterpret_cast<uint4*>(sx)[c] = reinterpret_cast<const uint4*>(x)[c];
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

// two rows at once: returns the two dot products (valid in every lane)
__device__ __forceinline__ void dot_pair(const bf16* __restrict__ w0, const bf16* __restrict__ w1, const bf16* sx, int K, int lane,
                                         float& a0, float& a1) {
  const uint4* p0 = reinterpret_cast<const uint4*>(w0);
  const uint4* p1 = reinterpret_cast<const uint4*>(w1);
  const uint4* px = reinterpret_cast<const uint4*>(sx);
  const int nchunk = K >> 3;
  a0 = 0.f;
  a1 = 0.f;
  int c = lane;
  for (; c + 96 < nchunk; c += 128) {
    uint4 u0[4], u1[4];
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
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) gemv_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ float red[32];
  bf16* sx = reinterpret_cast<bf16*>(smem_raw);
  stage_x(p.x, p.norm_weight, p.eps, p.K, sx, red);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = p.N >> 1;
  const int half = p.hd >> 1;
  for (int pi = blockIdx.x * WARPS + warp; pi < npairs; pi += gridDim.x * WARPS) {
    int r0, r1;
    if (MODE == SRGPT_GEMV_QKV_ROPE) {
      const int head = pi / half, j = pi % half;
      r0 = head * p.hd + j;
      r1 = r0 + half;
    } else {
      r0 = 2 * pi;
      r1 = r0 + 1;
    }
    float a0, a1;
    dot_pair(p.W + (size_t)r0 * p.ldw, p.W + (size_t)r1 * p.ldw, sx, p.K, lane, a0, a1);
    if (lane == 0) {
      if (MODE == SRGPT_GEMV_PLAIN) {
        float y0 = bf16_round(a0), y1 = bf16_round(a1);
        if (p.residual != nullptr) {
          y0 += __bfloat162float(p.residual[r0]);
          y1 += __bfloat162float(p.residual[r1]);
        }
        *reinterpret_cast<uint32_t*>(p.y + r0) = pack_bf16x2(y0, y1);
      } else if (MODE == SRGPT_GEMV_SWIGLU) {
        const float g = bf16_round(a0), u = bf16_round(a1);
        p.y[pi] = __float2bfloat16_rn(bf16_round(silu(g)) * u);
      } else {
        const int head = pi / half, j = pi % half;
        float v0 = bf16_round(a0), v1 = bf16_round(a1);
        const int pos = *p.pos;
        if (head < p.n_heads + p.n_kv_heads) {
          const float c = __bfloat162float(p.cos_tab[(size_t)pos * half + j]);
          const float s = __bfloat162float(p.sin_tab[(size_t)pos * half + j]);
          const float o0 = bf16_round(bf16_round(v0 * c) + bf16_round(-v1 * s));
          const float o1 = bf16_round(bf16_round(v1 * c) + bf16_round(v0 * s));
          v0 = o0;
          v1 = o1;
        }
        if (head < p.n_heads) {
          p.y[r0] = __float2bfloat16_rn(v0);
          p.y[r1] = __float2bfloat16_rn(v1);
        } else {
          const int page = p.page_table[pos / p.page_size], slot = pos % p.page_size;
          const bool is_v = head >= p.n_heads + p.n_kv_heads;
          const int kh = head - p.n_heads - (is_v ? p.n_kv_heads : 0);
          bf16* dst = p.kv_pages + (((size_t)page * 2 + (is_v ? 1 : 0)) * p.page_size + slot) * p.n_kv_heads * p.hd + kh * p.hd;
          dst[j] = __float2bfloat16_rn(v0);
          dst[j + half] = __float2bfloat16_rn(v1);
        }
      }
    }
  }
}

// ---- lm_head + argmax --------------------------------------------------------------------------
struct LmParams {
  const bf16* x;
  const bf16* W;
  int ldw;
  int V, K;
  const bf16* norm_weight;
  float eps;
  float* logits_out;
  float* part_val;
  int* part_idx;
};

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__global__ void __launch_bounds__(THREADS, 2) lm_head_kernel(const LmParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ float red[32];
  __shared__ float sv[WARPS];
  __shared__ int si[WARPS];
  bf16* sx = reinterpret_cast<bf16*>(smem_raw);
  stage_x(p.x, p.norm_weight, p.eps, p.K, sx, red);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = (p.V + 1) >> 1;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int pi = blockIdx.x * WARPS + warp; pi < npairs; pi += gridDim.x * WARPS) {
    const int r0 = 2 * pi;
    const int r1 = (r0 + 1 < p.V) ? r0 + 1 : r0;  // odd V: last pair reads row r0 twice
    float a0, a1;
    dot_pair(p.W + (size_t)r0 * p.ldw, p.W + (size_t)r1 * p.ldw, sx, p.K, lane, a0, a1);
    // logits = lm_head(h).float(): bf16 rounding first (modeling_llama.py:1044-1045)
    a0 = bf16_round(a0);
    a1 = bf16_round(a1);
    if (lane == 0) {
      if (p.logits_out != nullptr) {
        p.logits_out[r0] = a0;
        if (r1 != r0) p.logits_out[r1] = a1;
      }
      if (better(a0, r0, best, bi)) { best = a0; bi = r0; }
      if (r1 != r0 && better(a1, r1, best, bi)) { best = a1; bi = r1; }
    }
  }
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < WARPS; ++w)
      if (better(sv[w], si[w], best, bi)) { best = sv[w]; bi = si[w]; }
    p.part_val[blockIdx.x] = best;
    p.part_idx[blockIdx.x] = bi;
  }
}

__global__ void __launch_bounds__(256)
lm_head_finalize_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int nparts,
                        const bf16* __restrict__ embed_table, bf16* __restrict__ next_x, int K, long long* __restrict__ out_ids,
                        int* step, int* pos) {
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ int s_tok;
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

static int grid_for(int npairs) {
  int g = ceil_div(npairs, WARPS);
  const int cap = 2 * sm_count();
  return g < cap ? g : cap;
}

template <int MODE>
static int launch(const Params& p, cudaStream_t st) {
  const int smem = p.K * 2;
  static int configured_smem = 0;
  if (smem > 48 * 1024 && smem > configured_smem) {
    SRGPT_CHECK_CUDA(cudaFuncSetAttribute(gemv_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  gemv_kernel<MODE><<<grid_for(p.N >> 1), THREADS, smem, st>>>(p);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
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
  SRGPT_CHECK_ARG(x != y);  // x is re-read by late CTAs while early ones already write y
  gemv::Params p;
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
      return gemv::launch<SRGPT_GEMV_PLAIN>(p, st);
    case SRGPT_GEMV_SWIGLU:
      SRGPT_CHECK_ARG(residual == nullptr);
      p.hd = 2;
      return gemv::launch<SRGPT_GEMV_SWIGLU>(p, st);
    case SRGPT_GEMV_QKV_ROPE:
      SRGPT_CHECK_ARG(residual == nullptr && n_heads > 0 && n_kv_heads > 0 && head_dim > 0 && (head_dim % 2) == 0);
      SRGPT_CHECK_ARG(N == (n_heads + 2 * n_kv_heads) * head_dim);
      SRGPT_CHECK_ARG(cos_tab && sin_tab && pos && kv_pages && page_table && page_size > 0);
      return gemv::launch<SRGPT_GEMV_QKV_ROPE>(p, st);
  }
  return SRGPT_ERR_INVALID;
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
  const int g = gemv::grid_for((V + 1) / 2);
  gemv::LmParams p;
  p.x = reinterpret_cast<const bf16*>(x);
  p.W = reinterpret_cast<const bf16*>(W);
  p.ldw = ldw; p.V = V; p.K = K;
  p.norm_weight = reinterpret_cast<const bf16*>(norm_weight);
  p.eps = eps;
  p.logits_out = logits_out;
  p.part_val = reinterpret_cast<float*>(workspace);
  p.part_idx = reinterpret_cast<int*>(p.part_val + g);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int smem = K * 2;
  static int configured_smem = 0;
  if (smem > 48 * 1024 && smem > configured_smem) {
    SRGPT_CHECK_CUDA(cudaFuncSetAttribute(gemv::lm_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  gemv::lm_head_kernel<<<g, gemv::THREADS, smem, st>>>(p);
  SRGPT_CHECK_LAUNCH();
  gemv::lm_head_finalize_kernel<<<1, 256, 0, st>>>(p.part_val, p.part_idx, g, reinterpret_cast<const bf16*>(embed_table),
                                                   reinterpret_cast<bf16*>(next_x), K, out_ids, step, pos);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
}

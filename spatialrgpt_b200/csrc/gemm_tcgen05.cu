// Dense bf16 GEMM for sm_100a: C[M,N] = epilogue(A[M,K] · W[N,K]^T), fp32 accumulation in TMEM.
//
// This is the tensor-core core of the prefill path (SURVEY.md §2c K1,K3,K5,K6,K8,K9,K12,K13,K16,
// K20,K21): every nn.Linear / Conv2d(k=s) / ConvTranspose2d(k=s) the reference runs through
// cuBLAS/cuDNN (e.g. modeling_llama.py:429-431,498,221; base_extractor.py:92-97,158;
// base_projector.py:76-79) is one instantiation of this kernel with a fused epilogue.
//
// Design (Blackwell-native, no library code):
//   * persistent kernel, one CTA per SM, static round-robin over 128x128 output tiles;
//   * warp 0  : TMA producer  (cp.async.bulk.tensor 2D, 128B-swizzled K-major boxes, 6-stage ring);
//   * warp 1  : tcgen05.mma issuer (one elected lane), accumulators double-buffered in TMEM;
//   * warps 2-9: epilogue (tcgen05.ld 32x32b -> registers -> fused bias/activation/residual ->
//                 16-byte global stores); overlaps the next tile's main loop.
//   * mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
// Out-of-bounds rows/cols/K are zero-filled by TMA, so M, N need no padding and K only has to
// be a multiple of 8 elements (16-byte global strides).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "srgpt_b200.h"
#include "tcgen05.cuh"

namespace srgpt {
namespace gemm {

using namespace tc;  // mbarrier / TMA / TMEM / tcgen05 wrappers shared with attention_tc.cu

constexpr int BM = 128, BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int ACC_BUFS = 2;
constexpr int NUM_EPI_WARPS = 8;           // two warps per TMEM lane group, each owns half of the tile's columns
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;  // warp0 TMA, warp1 MMA, warps 2..9 epilogue
constexpr int MAX_STAGES = 6;
// Staged epilogue: every epilogue warp owns STG_BUFS buffers of one 32-row x 32-column bf16 chunk (64-byte rows laid out in the
// TMA SWIZZLE_64B pattern so the row-per-thread 16-byte writes are bank-conflict free); a chunk leaves through ONE TMA store
// (cp.async.bulk.tensor ... global.shared::cta, SASS UTMASTG) instead of 32 lanes x 4 scattered 16-byte STG (each a separate
// half-sector L2 write) - the row-per-thread stores were what paced the short-K SigLIP GEMMs (DESIGN.md "GEMM").
constexpr int STG_CHUNK_BYTES = 32 * 32 * 2;
constexpr int STG_BUFS = 2;
constexpr int STG_BYTES = NUM_EPI_WARPS * STG_BUFS * STG_CHUNK_BYTES;  // 32 KB
constexpr int BAR_BYTES = 512;

// Tile configuration.  BN = 128: 32 KB / stage, 6 stages, 256 TMEM columns.  BN = 256: 48 KB / stage, 4 stages,
// all 512 TMEM columns — 1.33x the FLOPs per byte pulled from L2, which is what bounds 128x128 tiles
// (every SM pulls ~45 B/clk through TMA while the MMA pipe could eat 128 B/clk; see DESIGN.md "GEMM").
template <int BN_>
struct Cfg {
  static constexpr int BN = BN_;
  static constexpr int STAGES = BN_ == 128 ? 6 : 4;
  static constexpr int B_STAGE_BYTES = BN_ * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int TMEM_COLS = ACC_BUFS * BN_;  // 256 / 512 (power of two >= 32)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align slack*/ + BAR_BYTES;
};



__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
// same copy delivered to the same shared-memory offset (and mbarrier offset) of every CTA in `cta_mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int x, int y, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(x), "r"(y), "h"(cta_mask)
      : "memory");
}
// L2 prefetch of one box (no shared-memory destination, no barrier): turns the later TMA load into an L2 hit
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tmap, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y) : "memory");
}
// TMA store of one box from shared memory (bulk async-group completion); rows / columns outside the tensor are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, uint32_t smem_src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src),
               "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source (the buffers may then be rewritten)
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// same, but the arrive is delivered to the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
               : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (ignored for swizzled K-major; 1) | [32,46) SBO>>4 (8 rows * 128 B)
//   [46,48) version=1 | [61,64) layout_type=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN
__device__ __forceinline__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (SRGPT_UMMA_FMT << 7) | (SRGPT_UMMA_FMT << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct Params {
  int M, N, K;
  int ldc;                 // elements
  const bf16* bias;        // [N] or null
  const bf16* residual;    // [*, ldr] or null
  int ldr;
  int res_row_mod;         // >0: residual row = row % res_row_mod (broadcast position embeddings)
  void* C;
  int epilogue;
  int out_fp32;
  int gm;                  // rasterisation: m-units per group (see unit_to_tile)
  int l2_prefetch;         // k-blocks the producer prefetches into L2 ahead of its loads (0 = off)
  int res_prefetch;        // 1: residual rows are requested before the accumulator wait
  int staged;              // 1: bf16 output leaves through shared memory + TMA stores (tmap_c valid)
  int res_tma;             // 1 (pair kernel only): the residual tile is TMA-loaded into the staging buffers (tmap_r valid)
};

// Tile order.  Units are walked group by group; a group is `gm` vertically adjacent m-units x ALL n-tiles, inside a group the
// m-unit varies fastest.  The CTAs (unit = cid, cid + ncl, ...) therefore work on ~148 neighbouring m-tiles of one weight
// tile, and a group's activation rows (gm * 128 * K * 2 bytes, sized to fit L2) are re-read from L2 - not HBM - for every
// n-tile.  gm = all m-units is the plain "m fastest" order, right when the whole activation fits L2 (Llama prompts); for the
// ViT tower (65536 x 4304 activations = 564 MB) it re-streamed A from HBM once per n-tile (fc2: 2.8 GB for a 0.56 GB problem).
__device__ __forceinline__ void unit_to_tile(int unit, int tiles_mu, int tiles_n, int gm, int& mu, int& nt) {
  const int per_group = gm * tiles_n;
  const int g = unit / per_group;
  const int rem = unit - g * per_group;
  const int g0 = g * gm;
  const int gsz = min(gm, tiles_mu - g0);
  nt = rem / gsz;
  mu = g0 + rem - nt * gsz;
}

template <int EPI, bool RES_STAGED = false>
__device__ __forceinline__ void apply_epilogue(float* v /*32 accumulators*/, const Params& p, int row, int col0, const uint4* pre_res = nullptr) {
  // v[j] is the fp32 accumulator of column col0 + j.  Rounding points mirror the reference's
  // sequence of bf16 torch ops (linear -> activation -> residual add), see DESIGN.md.
  if (EPI == SRGPT_EPI_NONE) return;
  if (EPI == SRGPT_EPI_SWIGLU) return;  // handled by the caller (pairs of columns)
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (col0 + j + 8 <= p.N) {
        uint4 b = *reinterpret_cast<const uint4*>(p.bias + col0 + j);
        float f[8];
        unpack8(b, f);
#pragma unroll
        for (int t = 0; t < 8; ++t) v[j + t] += f[t];
      } else {
        for (int t = 0; t < 8; ++t)
          if (col0 + j + t < p.N) v[j + t] += e2f(p.bias[col0 + j + t]);
      }
    }
  }
  if (EPI == SRGPT_EPI_BIAS_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(bf16_round(v[j]));
  } else if (EPI == SRGPT_EPI_BIAS_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(bf16_round(v[j]));
  } else if (EPI == SRGPT_EPI_BIAS_QUICK_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = quick_gelu(bf16_round(v[j]));
  } else if (EPI == SRGPT_EPI_BIAS_RESIDUAL) {
    if (RES_STAGED) {
      // all 32 columns come from the TMA-loaded residual tile (columns beyond N were zero-filled and are clipped by the store)
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float f[8];
        unpack8(pre_res[j >> 3], f);
#pragma unroll
        for (int t = 0; t < 8; ++t) v[j + t] = bf16_round(v[j + t]) + f[t];
      }
    } else if (p.residual != nullptr) {
      const int rrow = p.res_row_mod > 0 ? row % p.res_row_mod : row;
      const bf16* rp = p.residual + (size_t)rrow * p.ldr + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (col0 + j + 8 <= p.N) {
          // the residual chunk was requested before the accumulator wait when the whole 32-column chunk is inside N
          uint4 b = (pre_res != nullptr && p.res_prefetch && col0 + 32 <= p.N) ? pre_res[j >> 3] : *reinterpret_cast<const uint4*>(rp + j);
          float f[8];
          unpack8(b, f);
#pragma unroll
          for (int t = 0; t < 8; ++t) v[j + t] = bf16_round(v[j + t]) + f[t];
        } else {
          for (int t = 0; t < 8; ++t)
            if (col0 + j + t < p.N) v[j + t] = bf16_round(v[j + t]) + e2f(rp[j + t]);
        }
      }
    }
  }
}

// one thread's 32 consecutive accumulator columns of one output row: fused epilogue + 16-byte stores
template <int EPI>
__device__ __forceinline__ void store_chunk(const uint32_t* r, const Params& p, int row, int col0, const uint4* pre_res = nullptr) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (EPI == SRGPT_EPI_SWIGLU) {
    // interleaved weight rows: column 2i = gate_i, 2i+1 = up_i -> out[:, i]
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float g = bf16_round(v[2 * j]), u = bf16_round(v[2 * j + 1]);
      o[j] = bf16_round(silu(g)) * u;
    }
    bf16* cp = reinterpret_cast<bf16*>(p.C) + (size_t)row * p.ldc + (col0 >> 1);
    const int ncols_out = p.N >> 1;
    if ((col0 >> 1) + 16 <= ncols_out) {
      *reinterpret_cast<uint4*>(cp) = pack8(o);
      *reinterpret_cast<uint4*>(cp + 8) = pack8(o + 8);
    } else {
      for (int j = 0; j < 16; ++j)
        if ((col0 >> 1) + j < ncols_out) cp[j] = f2e(o[j]);
    }
  } else {
    apply_epilogue<EPI>(v, p, row, col0, pre_res);
    if (p.out_fp32) {
      float* cp = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col0;
      if (col0 + 32 <= p.N && (p.ldc & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) cp[j] = v[j];
      }
    } else {
      bf16* cp = reinterpret_cast<bf16*>(p.C) + (size_t)row * p.ldc + col0;
      if (col0 + 32 <= p.N) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) *reinterpret_cast<uint4*>(cp + j) = pack8(v + j);
      } else {
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) cp[j] = f2e(v[j]);
      }
    }
  }
}

// Staged variant (bf16 output, every epilogue except SwiGLU): the warp's 32 rows x 32 columns go to its staging buffer in the
// SWIZZLE_64B layout (16-byte chunk j of row r at r*64 + ((j ^ ((r >> 1) & 3)) << 4)) and leave with one TMA store, which also
// clips rows >= M and columns >= N.  `buf` is free again once at most BUFS-1 younger stores of this lane are still reading.
template <int EPI, int BUFS, bool RES_STAGED>
__device__ __forceinline__ void store_chunk_staged(const uint32_t* r, const Params& p, const CUtensorMap* tmap_c, uint8_t* buf, int lane, int row,
                                                   int row0, int col0, const uint4* pre_res) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  uint8_t* rp = buf + lane * 64;
  const int sw = (lane >> 1) & 3;
  if (RES_STAGED) {
    uint4 res[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) res[j] = *reinterpret_cast<const uint4*>(rp + ((j ^ sw) << 4));
    apply_epilogue<EPI, true>(v, p, row, col0, res);
  } else {
    if (row < p.M) apply_epilogue<EPI, false>(v, p, row, col0, pre_res);
    if (lane == 0) bulk_wait_group_read<BUFS - 1>();
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(rp + ((j ^ sw) << 4)) = pack8(v + 8 * j);
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmap_c, smem_u32(buf), col0, row0);
    bulk_commit_group();
  }
}

// Residual rows of one thread's chunks (CPW chunks of 32 columns), requested BEFORE the wait on the accumulator: the addresses
// depend only on the tile index, so the global-load latency (the longest link of the epilogue chain of the short-K ViT
// GEMMs: out_proj ran at 0.38 of peak with the residual against 0.58 without) hides behind the tile's main loop.
template <int EPI, int CPW>
__device__ __forceinline__ void prefetch_residual(uint4 (&pre)[CPW][4], const Params& p, int row, int n0, int c_first) {
  if (EPI != SRGPT_EPI_BIAS_RESIDUAL) return;
  if (p.residual == nullptr || row >= p.M || !p.res_prefetch) return;
  const int rrow = p.res_row_mod > 0 ? row % p.res_row_mod : row;
#pragma unroll
  for (int ci = 0; ci < CPW; ++ci) {
    const int col0 = n0 + (c_first + ci) * 32;
    if (col0 + 32 <= p.N) {
      const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (size_t)rrow * p.ldr + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j) pre[ci][j] = rp[j];
    }
  }
}

// CL = CTAs per cluster (1 or 2).  With CL = 2 the two CTAs of a cluster own vertically adjacent 128-row tiles and SHARE
// the B (weight) tile: each CTA fetches half of its rows and multicasts them into both CTAs' shared memory, so the bytes
// pulled from L2 per CTA and k-block drop from 16+BN/8 KB... (A + B) to A + B/2 — the quantity that bounds this kernel.
// A stage may be refilled only after BOTH CTAs' MMAs have read it, hence the multicast tcgen05.commit on the
// empty barriers (count CL).
template <int EPI, int BN, int CL, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const Params p) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int STAGE_BYTES = C::STAGE_BYTES;
  constexpr int TMEM_COLS = C::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned stage buffers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES;  // epilogue staging (STG_BYTES), 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + STG_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * STAGES;     // [ACC_BUFS]
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC_BUFS;  // [ACC_BUFS]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC_BUFS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  // work unit = CL vertically adjacent tiles x one n-tile; unit u -> m-unit u % tiles_mu, n-tile u / tiles_mu
  const int crank = (CL == 2) ? (int)cluster_ctarank() : 0;
  const int cid = (int)blockIdx.x / CL, ncl = (int)gridDim.x / CL;
  const int tiles_mu = (tiles_m + CL - 1) / CL;
  const int num_units = tiles_mu * tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (p.staged) prefetch_tmap(&tmap_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), CL);  // one tcgen05.commit per CTA sharing the B tile
    }
    for (int a = 0; a < ACC_BUFS; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), EW);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_base_slot), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // the peer's barriers are initialised before anything can arrive on them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      // optional L2 prefetch cursor, p.l2_prefetch k-blocks ahead of the loads (runs on into the next tiles of this CTA).
      // Idea: the ring holds only STAGES * 48 KB in flight, which bounds the ingest at ~60 B/clk/SM (ncu: tensor pipe 62 %
      // active on the ViT GEMMs).  Measured: it makes things worse (see launch()), so it is off unless SRGPT_GEMM_L2PF is set.
      int pf_unit = cid, pf_kb = 0, pf_m0 = 0, pf_n0 = 0;
      bool pf_new = true;
      auto pf_step = [&](bool issue) {
        if (pf_unit >= num_units) return;
        if (pf_new) {
          int mu, nt;
          unit_to_tile(pf_unit, tiles_mu, tiles_n, p.gm, mu, nt);
          pf_m0 = (mu * CL + crank) * BM;
          pf_n0 = nt * BN + (CL == 2 ? crank * (BN / 2) : 0);
          pf_new = false;
        }
        if (issue) {
          tma_prefetch_2d(&tmap_a, pf_kb * BK, pf_m0);
          tma_prefetch_2d(&tmap_b, pf_kb * BK, pf_n0);
        }
        if (++pf_kb == num_kb) { pf_kb = 0; pf_unit += ncl; pf_new = true; }
      };
      for (int i = 0; i < p.l2_prefetch; ++i) pf_step(false);
      for (int unit = cid; unit < num_units; unit += ncl) {
        int mu, nt;
        unit_to_tile(unit, tiles_mu, tiles_n, p.gm, mu, nt);
        const int m0 = (mu * CL + crank) * BM;
        const int n0 = nt * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (p.l2_prefetch > 0) pf_step(true);
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, STAGE_BYTES);  // own A tile + the whole B tile (one half from each CTA when CL == 2)
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_2d(smem_u32(sa), &tmap_a, fb, kb * BK, m0);
          if (CL == 2) {
            constexpr int HALF = BN / 2;
            tma_load_2d_mc(smem_u32(sa + A_STAGE_BYTES + crank * HALF * BK * 2), &tmap_b, fb, kb * BK, n0 + crank * HALF, (uint16_t)0x3);
          } else {
            tma_load_2d(smem_u32(sa + A_STAGE_BYTES), &tmap_b, fb, kb * BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      uint32_t stage = 0, phase = 0;
      uint32_t acc = 0, acc_phase = 0;
      for (int unit = cid; unit < num_units; unit += ncl) {
        mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(b_addr);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advancing 16 bf16 = 32 bytes along K inside the 128B swizzle atom: +2 in the (>>4) address field
            umma_f16(tmem_d, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot when these MMAs retire (in both CTAs that write into it when the B tile is shared)
          if (CL == 2) umma_commit_mc(smem_u32(&empty_bar[stage]), (uint16_t)0x3);
          else umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&tmem_full_bar[acc]));  // accumulator complete -> epilogue
        if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    // With short K (SigLIP: 18 k-blocks) the epilogue of a tile costs as many issue slots as its main loop, and one
    // warp per scheduler cannot hide the TMEM / global latencies -> two warps per lane group, half the columns each.
    const int lg = warp & 3;  // TMEM lane group this warp may access: lanes [32*lg, 32*lg+32)
    // EW / 4 warps share a lane group; each owns a contiguous share of the tile's 32-column chunks
    constexpr int CPW = (BN / 32) / (EW / 4);
    constexpr bool CAN_STAGE = EW == NUM_EPI_WARPS && EPI != SRGPT_EPI_SWIGLU;  // staging is sized for 8 epilogue warps
    const int cpart = (warp - 2) >> 2;
    const bool staged = CAN_STAGE && p.staged != 0;
    uint8_t* stg = stg_base + (warp - 2) * (STG_BUFS * STG_CHUNK_BYTES);
    uint32_t sbuf = 0;
    uint32_t acc = 0, acc_phase = 0;
    for (int unit = cid; unit < num_units; unit += ncl) {
      int mu, nt;
      unit_to_tile(unit, tiles_mu, tiles_n, p.gm, mu, nt);
      const int m0 = (mu * CL + crank) * BM;
      const int n0 = nt * BN;
      const int row = m0 + lg * 32 + lane;
      uint4 pre[EPI == SRGPT_EPI_BIAS_RESIDUAL ? CPW : 1][4];
      if (EPI == SRGPT_EPI_BIAS_RESIDUAL) prefetch_residual<EPI, (EPI == SRGPT_EPI_BIAS_RESIDUAL ? CPW : 1)>(pre, p, row, n0, cpart * CPW);
      mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_phase);
      tcgen05_fence_after();
#pragma unroll
      for (int ci = 0; ci < CPW; ++ci) {
        const int c = cpart * CPW + ci;
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + acc * BN + c * 32 + ((uint32_t)(lg * 32) << 16), r);
        tmem_ld_wait();
        const uint4* pr = EPI == SRGPT_EPI_BIAS_RESIDUAL ? pre[EPI == SRGPT_EPI_BIAS_RESIDUAL ? ci : 0] : nullptr;
        if (CAN_STAGE && staged) {
          store_chunk_staged<EPI, STG_BUFS, false>(r, p, &tmap_c, stg + sbuf * STG_CHUNK_BYTES, lane, row, m0 + lg * 32, col0, pr);
          sbuf ^= 1;
        } else if (row < p.M) {
          store_chunk<EPI>(r, p, row, col0, pr);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
      if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
    }
    if (CAN_STAGE && staged && lane == 0) bulk_wait_group_read<0>();  // shared memory stays valid until every store has read it
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair configuration (tcgen05 cta_group::2): the two CTAs of a cluster (one TPC) compute ONE 256 x 256 tile.  CTA r keeps
// A rows [128 r, 128 r + 128) and HALF of the weight tile (rows [128 r, 128 r + 128) of the 256) in its own shared memory; the
// leader issues M = 256 MMAs that read the A tile of each CTA and BOTH weight halves (the peer's through the pair's shared-
// memory path), and every CTA ends up with its 128 x 256 accumulator in its own TMEM.  Per k-block a CTA now ingests
// 16 + 16 KB instead of 16 + 32 KB for the same 128 x 256 x 64 MACs - the quantity that bounds the 1-CTA kernel (~60 B/clk/SM
// of TMA ingest, tensor pipe 62-80 % active) - and the smaller stage allows 6 stages instead of 4.
//   * both CTAs' TMA loads complete on the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2, barrier address
//     mapped into the leader with mapa); the leader arms it with the bytes of both;
//   * tcgen05.commit.cta_group::2 multicasts the "slot free" / "accumulator ready" arrivals to both CTAs;
//   * the peer's epilogue warps release the accumulator on the leader's tmem_empty barrier (remote arrive).
// ---------------------------------------------------------------------------------------------
constexpr int PAIR_BN = 256;
constexpr bool PAIR_DEFAULT = true;
constexpr int PAIR_B_HALF_BYTES = (PAIR_BN / 2) * BK * 2;                 // 16 KB
constexpr int PAIR_STAGE_BYTES = A_STAGE_BYTES + PAIR_B_HALF_BYTES;       // 32 KB per CTA
// RES_TMA (residual epilogue): the tile's residual rows are TMA-LOADED into the staging buffers while the tile's main loop runs
// (one 32 x 32 box per chunk, 4 buffers per warp = all of a warp's chunks), the epilogue adds in place and the same buffer leaves
// through the TMA store.  That needs 64 KB of staging, paid for with one pipeline stage (5 instead of 6).
template <bool RES_TMA>
struct PairCfg {
  static constexpr int STAGES = RES_TMA ? 5 : 6;
  static constexpr int BUFS = RES_TMA ? 4 : STG_BUFS;
  static constexpr int STG = NUM_EPI_WARPS * BUFS * STG_CHUNK_BYTES;
  static constexpr int SMEM_BYTES = STAGES * PAIR_STAGE_BYTES + STG + 1024 + BAR_BYTES;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are signalled on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar_cluster_addr, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {  // arrives on this barrier offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)0x3)
               : "memory");
}

template <int EPI, int EW, bool RES_TMA>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_r, const Params p) {
  using PC = PairCfg<RES_TMA>;
  constexpr int BN = PAIR_BN, STAGES = PC::STAGES, STAGE_BYTES = PAIR_STAGE_BYTES, TMEM_COLS = ACC_BUFS * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + PC::STG);
  uint64_t* res_bar = bars + 32;  // [NUM_EPI_WARPS * 4] (RES_TMA): one per warp and chunk, at byte 256 of the barrier block
  uint64_t* full_bar = bars;                            // [STAGES]  used in the leader only
  uint64_t* empty_bar = bars + STAGES;                  // [STAGES]  one per CTA (multicast commit)
  uint64_t* tmem_full_bar = bars + 2 * STAGES;          // [ACC_BUFS] one per CTA (multicast commit)
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC_BUFS;  // [ACC_BUFS] used in the leader only: both CTAs' epilogue warps arrive
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC_BUFS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  const int crank = (int)cluster_ctarank();
  const int cid = (int)blockIdx.x / 2, ncl = (int)gridDim.x / 2;
  const int tiles_mu = (tiles_m + 1) / 2;
  const int num_units = tiles_mu * tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < ACC_BUFS; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), 2 * EW);
    }
    if (RES_TMA)
      for (int i = 0; i < NUM_EPI_WARPS * 4; ++i) mbar_init(smem_u32(&res_bar[i]), 1);
    if (p.staged) prefetch_tmap(&tmap_c);
    if (RES_TMA) prefetch_tmap(&tmap_r);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32(tmem_base_slot), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int unit = cid; unit < num_units; unit += ncl) {
        int mu, nt;
        unit_to_tile(unit, tiles_mu, tiles_n, p.gm, mu, nt);
        const int m0 = (mu * 2 + crank) * BM;
        const int n0 = nt * BN + crank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (crank == 0) mbar_expect_tx(smem_u32(&full_bar[stage]), 2 * STAGE_BYTES);  // bytes of both CTAs
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_2d_pair(smem_u32(sa), &tmap_a, fb_leader, kb * BK, m0);
          tma_load_2d_pair(smem_u32(sa + A_STAGE_BYTES), &tmap_b, fb_leader, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      uint32_t stage = 0, phase = 0;
      uint32_t acc = 0, acc_phase = 0;
      for (int unit = cid; unit < num_units; unit += ncl) {
        mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16_pair(tmem_d, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(smem_u32(&empty_bar[stage]));  // slot free in both CTAs once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(smem_u32(&tmem_full_bar[acc]));  // accumulators (one per CTA) complete
        if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..9), both CTAs: rows of this CTA's 128 x 256 accumulator =====================
    const int lg = warp & 3;
    constexpr int CPW = (BN / 32) / (EW / 4);
    constexpr bool CAN_STAGE = EW == NUM_EPI_WARPS && EPI != SRGPT_EPI_SWIGLU;
    constexpr bool RES_STAGED = RES_TMA && EPI == SRGPT_EPI_BIAS_RESIDUAL && CAN_STAGE && CPW == 4;
    const int cpart = (warp - 2) >> 2;
    const bool staged = CAN_STAGE && p.staged != 0;
    uint8_t* stg = stg_base + (warp - 2) * (PC::BUFS * STG_CHUNK_BYTES);
    uint64_t* my_res_bar = res_bar + (warp - 2) * 4;
    uint32_t sbuf = 0, res_phase = 0;  // res_phase: one parity bit per chunk barrier
    uint32_t acc = 0, acc_phase = 0;
    for (int unit = cid; unit < num_units; unit += ncl) {
      int mu, nt;
      unit_to_tile(unit, tiles_mu, tiles_n, p.gm, mu, nt);
      const int m0 = (mu * 2 + crank) * BM;
      const int n0 = nt * BN;
      const int row = m0 + lg * 32 + lane;
      uint4 pre[(EPI == SRGPT_EPI_BIAS_RESIDUAL && !RES_STAGED) ? CPW : 1][4];
      if (RES_STAGED) {
        // the residual boxes of this warp's chunks land in its 4 staging buffers while the main loop of the tile runs
        if (lane == 0) {
          bulk_wait_group_read<0>();  // the previous tile's stores have finished reading the buffers
#pragma unroll
          for (int ci = 0; ci < CPW; ++ci) {
            const int col0 = n0 + (cpart * CPW + ci) * 32;
            if (col0 < p.N) {
              const uint32_t rb = smem_u32(&my_res_bar[ci]);
              mbar_expect_tx(rb, STG_CHUNK_BYTES);  // out-of-bounds rows / columns are zero-filled and still counted
              tma_load_2d(smem_u32(stg + ci * STG_CHUNK_BYTES), &tmap_r, rb, col0, m0 + lg * 32);
            }
          }
        }
        __syncwarp();
      } else if (EPI == SRGPT_EPI_BIAS_RESIDUAL) {
        prefetch_residual<EPI, (EPI == SRGPT_EPI_BIAS_RESIDUAL && !RES_STAGED) ? CPW : 1>(pre, p, row, n0, cpart * CPW);
      }
      mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_phase);
      tcgen05_fence_after();
#pragma unroll
      for (int ci = 0; ci < CPW; ++ci) {
        const int c = cpart * CPW + ci;
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + acc * BN + c * 32 + ((uint32_t)(lg * 32) << 16), r);
        tmem_ld_wait();
        if (RES_STAGED) {
          mbar_wait(smem_u32(&my_res_bar[ci]), (res_phase >> ci) & 1u);
          res_phase ^= 1u << ci;
          store_chunk_staged<EPI, PC::BUFS, true>(r, p, &tmap_c, stg + ci * STG_CHUNK_BYTES, lane, row, m0 + lg * 32, col0, nullptr);
        } else {
          const uint4* pr = (EPI == SRGPT_EPI_BIAS_RESIDUAL && !RES_STAGED) ? pre[(EPI == SRGPT_EPI_BIAS_RESIDUAL && !RES_STAGED) ? ci : 0] : nullptr;
          if (CAN_STAGE && staged) {
            store_chunk_staged<EPI, STG_BUFS, false>(r, p, &tmap_c, stg + sbuf * STG_CHUNK_BYTES, lane, row, m0 + lg * 32, col0, pr);
            sbuf ^= 1;
          } else if (row < p.M) {
            store_chunk<EPI>(r, p, row, col0, pr);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));
      if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
    }
    if (CAN_STAGE && (staged || RES_STAGED) && lane == 0) bulk_wait_group_read<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// "Tall" configuration for short prompts (128 < M <= 384, e.g. the S = 259 rows of config c2) — an EXPERIMENT that is
// correct but slower than the default tiles (see launch()); kept opt-in so the measurement stays reproducible.
// With 128-row tiles every weight tile is pulled from L2 once per row tile and the activation tile once per
// column tile: 1.0 GB of L2->SM traffic for the 235 MB gate/up weights, 95 us instead of the 37 us HBM floor.
// Here ONE CTA owns all M rows (3 x 128-row accumulators, 384 TMEM columns) of a BN-column tile, so weights cross
// L2->SM once, and the 4 CTAs of a cluster (4 neighbouring column tiles) each fetch a quarter of the 48 KB activation
// tile and multicast it to the others: per CTA and k-block 12 KB of A + BN*128 B of W for 3*BN*128*64*2 FLOP.
// ---------------------------------------------------------------------------------------------
constexpr int TALL_MT = 3;
constexpr int TALL_CS = 4;
constexpr int TALL_A_BYTES = TALL_MT * A_STAGE_BYTES;        // 48 KB
constexpr int TALL_A_SLICE_ROWS = TALL_MT * BM / TALL_CS;    // 96 rows loaded (and multicast) by each CTA

template <int BN_>
struct TallCfg {
  static constexpr int B_BYTES = BN_ * BK * 2;
  static constexpr int STAGE_BYTES = TALL_A_BYTES + B_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int TMEM_COLS = (TALL_MT * BN_ <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int EPI, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tall_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  using C = TallCfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int STAGE_BYTES = C::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  const int crank = (int)cluster_ctarank();
  const int cid = (int)blockIdx.x / TALL_CS, ncl = (int)gridDim.x / TALL_CS;
  const int num_units = (tiles_n + TALL_CS - 1) / TALL_CS;  // a unit = 4 neighbouring column tiles, one per CTA

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), TALL_CS);  // every CTA of the cluster reads the shared A tile
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    mbar_init(smem_u32(tmem_empty_bar), NUM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_base_slot), C::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int unit = cid; unit < num_units; unit += ncl) {
        const int n0 = (unit * TALL_CS + crank) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, STAGE_BYTES);  // the whole A tile (4 multicast slices) + my B tile
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_2d_mc(smem_u32(sa + crank * TALL_A_SLICE_ROWS * BK * 2), &tmap_a, fb, kb * BK, crank * TALL_A_SLICE_ROWS, (uint16_t)0xF);
          tma_load_2d(smem_u32(sa + TALL_A_BYTES), &tmap_b, fb, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      uint32_t stage = 0, phase = 0, acc_phase = 0;
      for (int unit = cid; unit < num_units; unit += ncl) {
        mbar_wait(smem_u32(tmem_empty_bar), acc_phase ^ 1);
        tcgen05_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + TALL_A_BYTES);
#pragma unroll
          for (int mt = 0; mt < TALL_MT; ++mt) {
            const uint64_t a_desc = make_smem_desc_sw128(a_addr + mt * A_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_f16(tmem_base + mt * BN, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_mc(smem_u32(&empty_bar[stage]), (uint16_t)0xF);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(tmem_full_bar));
        acc_phase ^= 1;
      }
    }
  } else {
    const int lg = warp & 3;
    const int chalf = (warp - 2) >> 2;
    uint32_t acc_phase = 0;
    for (int unit = cid; unit < num_units; unit += ncl) {
      const int n0 = (unit * TALL_CS + crank) * BN;
      mbar_wait(smem_u32(tmem_full_bar), acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < TALL_MT; ++mt) {
        const int row = mt * BM + lg * 32 + lane;
        if (mt * BM >= p.M) break;  // warp-uniform: no valid rows in this accumulator
#pragma unroll 1
        for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + mt * BN + c * 32 + ((uint32_t)(lg * 32) << 16), r);
          tmem_ld_wait();
          if (row < p.M) store_chunk<EPI>(r, p, row, col0);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tmem_empty_bar));
      acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// "Tall stream-K" configuration for short prompts (M <= 384 rows: the S = 259 prompt of config c2, the B <= 128 last rows of a
// batched lm_head).  These GEMMs stream the weights once and are bound by the bytes every SM has to ingest (smem capacity /
// L2 latency ~ 85 GB/s per SM, DESIGN.md "GEMM"), so the design minimises exactly that:
//   * ONE CTA holds ALL M rows (MT = ceil(M/128) accumulators of 128 x 128 in TMEM), so a weight tile is ingested once, not once
//     per 128-row tile (3x for M = 259);
//   * stream-K: the (n-tile, k-block) units are cut into gridDim.x equal contiguous ranges, one per SM, so every SM ingests the
//     same number of bytes whatever N / 128 is (down_proj: 32 n-tiles x 3 row tiles = 96 CTAs of the default kernel left a third
//     of the chip idle).  A tile whose k-range is split is owned by the CTA that holds its first k-block; the other CTAs write
//     their fp32 partial accumulators to a workspace slot (at most one per CTA: only the FIRST segment of a range can start
//     mid-tile) and raise an epoch flag; the owner adds the partials in CTA order (deterministic) and runs the fused epilogue.
//     Owners only ever wait on FIRST segments of higher CTAs, which wait on nothing: no cycles; all CTAs are co-resident
//     (grid <= number of SMs, one CTA per SM).
// The workspace (flags + one slot per CTA) is registered by the caller with srgpt_gemm_set_workspace; without it this path is off.
// ---------------------------------------------------------------------------------------------
constexpr int TSK_BN = 128;
constexpr int TSK_MAX_CTAS = 256;
constexpr int TSK_FLAG_BYTES = TSK_MAX_CTAS * 4;
constexpr long long TSK_SLOT_FLOATS = 3LL * BM * TSK_BN;  // MT = 3 accumulators of 128 x 128

// BN < 128 (MT = 1 only): narrow weight tiles for the skinny GEMMs of a batched decode step (M <= 128 rows), see launch_tall_sk
template <int MT, int BN = TSK_BN>
struct TskCfg {
  static constexpr int A_BYTES = MT * A_STAGE_BYTES;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;           // 32 / 48 / 64 KB (BN = 128); 24 / 20 KB (BN = 64 / 32)
  static constexpr int STAGES = MT == 3 ? 3 : (MT == 2 ? 4 : (BN == 128 ? 6 : (BN == 64 ? 8 : 9)));  // ~192 KB in flight in every case
  static constexpr int TMEM_COLS = MT == 1 ? (BN < 32 ? 32 : BN) : (MT == 2 ? 256 : 512);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + BAR_BYTES;
};

struct TskWs {
  float* partial;        // [gridDim.x][MT][128][BN] fp32
  unsigned int* flags;   // [gridDim.x], == epoch once CTA c's partial is complete
  unsigned int epoch;
  int whole_tiles;       // 1: CTAs own whole n-tiles (contiguous ranges balanced over the grid): no k-split, no partials, no flags
  int a_stage_tx;        // bytes one A box delivers per m-tile and k-block (the A map's box may hold fewer than 128 rows)
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int EPI, int MT, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tall_sk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p, const TskWs ws) {
  using C = TskCfg<MT, BN>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  const long long total = (long long)tiles_n * num_kb;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  auto range_begin = [&](int c) { return ws.whole_tiles ? (int)((long long)tiles_n * c / G) * num_kb : (int)(total * c / G); };
  const int u_begin = range_begin(cta), u_end = range_begin(cta + 1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    mbar_init(smem_u32(tmem_empty_bar), NUM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_base_slot), C::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer: runs ahead across segments =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int u = u_begin; u < u_end; ++u) {
        const int tile = u / num_kb, kb = u - tile * num_kb;
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_expect_tx(fb, MT * ws.a_stage_tx + C::B_BYTES);
        uint8_t* sa = smem + stage * STAGE_BYTES;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) tma_load_2d(smem_u32(sa + mt * A_STAGE_BYTES), &tmap_a, fb, kb * BK, mt * BM);  // rows >= M zero-filled
        // (a box of fewer than 128 rows leaves the rest of the A stage stale: those accumulator rows are never read)
        tma_load_2d(smem_u32(sa + C::A_BYTES), &tmap_b, fb, kb * BK, tile * BN);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      uint32_t stage = 0, phase = 0, acc_phase = 0;
      int u = u_begin;
      while (u < u_end) {
        const int tile = u / num_kb;
        const int seg_end = min(u_end, (tile + 1) * num_kb);
        mbar_wait(smem_u32(tmem_empty_bar), acc_phase ^ 1);
        tcgen05_fence_after();
        bool first = true;
        for (; u < seg_end; ++u) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + C::A_BYTES);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t a_desc = make_smem_desc_sw128(a_addr + mt * A_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_f16(tmem_base + mt * BN, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (first && k == 0) ? 0u : 1u);
          }
          first = false;
          umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(tmem_full_bar));
        acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int lg = warp & 3, cpart = (warp - 2) >> 2;  // lane group (rows), column half of the 128-wide tile
    uint32_t acc_phase = 0;
    int u = u_begin;
    while (u < u_end) {
      const int tile = u / num_kb;
      const int seg_end = min(u_end, (tile + 1) * num_kb);
      const int k0 = u - tile * num_kb;
      const bool owner = k0 == 0;
      const int n0 = tile * BN;
      // contributors of an owned, split tile: the CTAs after this one up to the one holding the tile's last k-block
      int c_last = cta;
      if (owner && seg_end < (tile + 1) * num_kb) {
        const int last_unit = (tile + 1) * num_kb - 1;
        while (c_last + 1 < G && range_begin(c_last + 1) <= last_unit) ++c_last;
      }
      mbar_wait(smem_u32(tmem_full_bar), acc_phase);
      tcgen05_fence_after();
      if (owner) {
        for (int cc = cta + 1; cc <= c_last; ++cc) {  // partials complete?  (lane 0 polls, the warp follows)
          if (lane == 0) {
            uint32_t spins = 0;
            while (ld_acquire_u32(ws.flags + cc) != ws.epoch) {
              __nanosleep(64);
              if (++spins > (1u << 24)) __trap();  // deadlock breaker
            }
          }
          __syncwarp();
        }
      }
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        if (mt * BM >= p.M) break;  // warp-uniform
        const int trow = lg * 32 + lane;  // row inside the 128-row accumulator
        const int row = mt * BM + trow;
        if (mt * BM + lg * 32 >= p.M) continue;  // warp-uniform: none of this warp's 32 rows exists (M = 32 of a batched decode step:
                                                  // three of the four lane groups have nothing to read, write or add)
        const bool row_ok = row < p.M;
#pragma unroll 1
        for (int ci = 0; ci < 2; ++ci) {
          const int c = cpart * 2 + ci;
          const int col0 = n0 + c * 32;
          if (c * 32 >= BN || col0 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          __syncwarp();  // the row guard below diverges; tcgen05.ld is warp-collective
          tmem_ld_32x32b_x32(tmem_base + mt * BN + c * 32 + ((uint32_t)(lg * 32) << 16), r);
          tmem_ld_wait();
          if (!row_ok) continue;  // rows >= M: zero-filled operand rows, never stored, never exchanged
          if (!owner) {
            float4* dst = reinterpret_cast<float4*>(ws.partial + ((size_t)cta * MT + mt) * (BM * BN) + (size_t)trow * BN + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          } else {
            // fixed order: deterministic sums.  The partials of up to four contributors are requested together (half a chunk = four
            // 16-byte loads each) so the fix-up pays one L2 round trip per half chunk instead of one per contributor
            for (int cc0 = cta + 1; cc0 <= c_last; cc0 += 4) {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                float4 v[4][4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  const int cc = cc0 + b;
                  const float4* src = reinterpret_cast<const float4*>(ws.partial + ((size_t)(cc <= c_last ? cc : cta + 1) * MT + mt) * (BM * BN) +
                                                                      (size_t)trow * BN + c * 32 + half * 16);
#pragma unroll
                  for (int j = 0; j < 4; ++j) v[b][j] = (cc <= c_last) ? __ldcg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const int q = half * 16 + 4 * j;
                    r[q] = __float_as_uint(__uint_as_float(r[q]) + v[b][j].x);
                    r[q + 1] = __float_as_uint(__uint_as_float(r[q + 1]) + v[b][j].y);
                    r[q + 2] = __float_as_uint(__uint_as_float(r[q + 2]) + v[b][j].z);
                    r[q + 3] = __float_as_uint(__uint_as_float(r[q + 3]) + v[b][j].w);
                  }
                }
              }
            }
            store_chunk<EPI>(r, p, row, col0);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tmem_empty_bar));
      if (!owner) {
        // publish the partial: every epilogue thread's stores, then one release store of the epoch
        __threadfence();
        named_bar_sync(1, 32 * NUM_EPI_WARPS);
        if (warp == 2 && lane == 0) st_release_u32(ws.flags + cta, ws.epoch);
      } else if (c_last > cta) {
        // every epilogue warp has consumed the contributors' partials: clear their flags, so that a CUDA-graph REPLAY of this
        // launch (same baked epoch) starts from zeroed flags again
        named_bar_sync(1, 32 * NUM_EPI_WARPS);
        if (warp == 2 && lane == 0)
          for (int cc = cta + 1; cc <= c_last; ++cc) st_release_u32(ws.flags + cc, 0u);
      }
      acc_phase ^= 1;
      u = seg_end;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// [rows, k] row-major bf16 matrix with `ld` elements between rows -> box {BK, box_rows}, 128B swizzle
static int make_tmap(CUtensorMap* tm, const void* ptr, int rows, int k, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable");
    return SRGPT_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, SRGPT_TMAP_DTYPE, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rows=%d k=%d ld=%d ptr=%p)", (int)r, rows, k, ld, ptr);
    return SRGPT_ERR_CUDA;
  }
  return SRGPT_OK;
}

// [rows, cols] row-major bf16 output / residual matrix -> box {32 columns, 32 rows}, 64B swizzle (the staged epilogue's chunk)
static int make_tmap_out(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable");
    return SRGPT_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, SRGPT_TMAP_DTYPE, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (output) failed: CUresult %d (rows=%d cols=%d ld=%d ptr=%p)", (int)r, rows, cols, ld, ptr);
    return SRGPT_ERR_CUDA;
  }
  return SRGPT_OK;
}

// Epilogue warps: 8 by default (two per TMEM lane group).  A 16-warp epilogue (four per lane group, SRGPT_GEMM_EW=16; the
// erf-GELU epilogue's 116 registers do not fit the 96-register budget of an 18-warp CTA and stays at 8) was built to shorten
// the per-tile latency chain of the short-K ViT GEMMs and MEASURED (profiles/r01_microbench_gemm_ew.txt): out_proj +
// residual 0.33 -> 0.37 of peak, but qkv 0.70 -> 0.62 and fc1 0.62 -> 0.56, Llama shapes unchanged -> not the default.
static int env_int(const char* name);

template <int EPI>
struct EpiWarps {
  static constexpr int N = EPI == SRGPT_EPI_BIAS_GELU_ERF ? 8 : 16;
};
static int epi_warps_env() {
  static const int v = env_int("SRGPT_GEMM_EW");  // 16 selects the 16-warp epilogue where it exists
  return v;
}

template <int EPI, int BN, int CL, int EW>
static int launch_cfg_ew(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const Params& p, int grid, cudaStream_t stream);

template <int EPI, int BN, int CL>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const Params& p, int grid, cudaStream_t stream) {
  if (EpiWarps<EPI>::N == 16 && epi_warps_env() == 16) {
    Params q = p;
    q.staged = 0;  // the staging buffers are sized for 8 epilogue warps
    return launch_cfg_ew<EPI, BN, CL, 16>(ta, tb, tc, q, grid, stream);
  }
  return launch_cfg_ew<EPI, BN, CL, 8>(ta, tb, tc, p, grid, stream);
}

template <int EPI, int BN, int CL, int EW>
static int launch_cfg_ew(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const Params& p, int grid, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    SRGPT_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<EPI, BN, CL, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(64 + 32 * EW);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (CL > 1) ? 1 : 0;
  SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tn_kernel<EPI, BN, CL, EW>, ta, tb, tc, p));
  return SRGPT_OK;
}

static int env_int(const char* name) {
  const char* v = getenv(name);
  return (v != nullptr && v[0] != 0) ? atoi(v) : 0;
}

// Configuration choice: 128x256 tiles when that still leaves >= 2 full waves of work units, else 128x128.
// Clusters of 2 with a shared (multicast) B tile are implemented and correct but NOT the default: measured on B200
// (profiles/r01_microbench_gemm_cl{1,2}.jsonl) they change nothing at M=8288 (1404 vs 1404 TFLOP/s — with 128x256
// tiles the kernel is no longer bound by L2->SM bytes) and lose at M=259 (58 vs 38 us: half as many independent
// producers).  SRGPT_GEMM_BN / SRGPT_GEMM_CL force a choice.
static void pick_cfg(int M, int N, int* bn, int* cl) {
  static const int force_bn = env_int("SRGPT_GEMM_BN"), force_cl = env_int("SRGPT_GEMM_CL");
  const int tiles_m = ceil_div(M, BM);
  *cl = (force_cl == 2 && tiles_m >= 2) ? 2 : 1;
  const long units256 = (long)ceil_div(tiles_m, *cl) * ceil_div(N, 256);
  *bn = (force_bn == 128 || force_bn == 256) ? force_bn : (units256 >= 2L * (sm_count() / *cl) ? 256 : 128);
}

template <int EPI, int BN>
static int launch_tall(const void* A, int lda, const void* W, int ldw, const Params& p, cudaStream_t stream) {
  using C = TallCfg<BN>;
  static bool configured = false;
  if (!configured) {
    SRGPT_CHECK_CUDA(cudaFuncSetAttribute(gemm_tall_kernel<EPI, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, A, p.M, p.K, lda, TALL_A_SLICE_ROWS);
  if (rc != SRGPT_OK) return rc;
  rc = make_tmap(&tb, W, p.N, p.K, ldw, BN);
  if (rc != SRGPT_OK) return rc;
  const int units = ceil_div(ceil_div(p.N, BN), TALL_CS);
  const int max_clusters = sm_count() / TALL_CS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((units < max_clusters ? units : max_clusters) * TALL_CS);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TALL_CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tall_kernel<EPI, BN>, ta, tb, p));
  return SRGPT_OK;
}

// ---- tall stream-K: workspace registration + launch
struct TskState {
  void* base = nullptr;
  long long bytes = 0;
  unsigned int epoch = 0;
};
static TskState g_tsk;

static long long tsk_workspace_bytes(int ctas) { return TSK_FLAG_BYTES + (long long)ctas * TSK_SLOT_FLOATS * 4; }

template <int EPI, int MT, int BN>
static int launch_tall_sk_mt(const void* A, int lda, const void* W, int ldw, const Params& p, cudaStream_t stream, int whole_grid) {
  using C = TskCfg<MT, BN>;
  static bool configured = false;
  if (!configured) {
    SRGPT_CHECK_CUDA(cudaFuncSetAttribute(gemm_tall_sk_kernel<EPI, MT, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  // A box: 128 rows, or 32 / 64 when that covers all M rows of a one-tile problem (fewer bytes written to shared memory per k-block)
  // the small boxes only in the (opt-in) whole-tile mode; with the stream-K split they measured the same as the full box
  const int a_box = whole_grid <= 0 ? BM : ((MT == 1 && p.M <= 32) ? 32 : ((MT == 1 && p.M <= 64) ? 64 : BM));
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, A, p.M, p.K, lda, a_box);
  if (rc != SRGPT_OK) return rc;
  rc = make_tmap(&tb, W, p.N, p.K, ldw, BN);
  if (rc != SRGPT_OK) return rc;
  const long long total = (long long)ceil_div(p.N, BN) * ceil_div(p.K, BK);
  int grid = sm_count();
  if (grid > TSK_MAX_CTAS) grid = TSK_MAX_CTAS;
  if ((long long)grid > total) grid = (int)total;
  TskWs ws;
  ws.whole_tiles = whole_grid > 0 ? 1 : 0;
  ws.a_stage_tx = a_box * BK * 2;
  if (whole_grid > 0) {
    grid = whole_grid;
  } else {
    while (tsk_workspace_bytes(grid) > g_tsk.bytes && grid > 1) --grid;  // a small workspace only narrows the grid
  }
  ws.flags = reinterpret_cast<unsigned int*>(g_tsk.base);
  ws.partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g_tsk.base) + TSK_FLAG_BYTES);
  if (++g_tsk.epoch == 0) g_tsk.epoch = 1;  // flags start at 0 (zeroed workspace) and never equal a future epoch
  ws.epoch = g_tsk.epoch;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tall_sk_kernel<EPI, MT, BN>, ta, tb, p, ws));
  return SRGPT_OK;
}

// One-tile-high problems (M <= 128: the projections and the lm_head of a batched decode step) - WHOLE n-tiles per CTA.
// ncu at M = 32 (profiles/r02_ncu_full_tsk_m32_summary.txt): the stream-K split costs a fixed ~18 us per launch whatever the matrix
// (partial stores + fence + flag, owner polls, dependent partial loads: o_proj 27 us against a 5 us HBM floor, down 40 / 18, gate-up
// 52 / 36).  Without a k-split nothing is exchanged; what fills the chip instead is a NARROWER weight tile: BN is chosen per problem from
// {128, 64, 32} by a two-term cost model - HBM time of the weights vs waves x per-SM TMA ingest (~88 GB/s, DESIGN.md "GEMM: what bounds
// it") of one tile's weight + activation bytes - and the grid is the balanced whole-tile count ceil(tiles / waves).
// MEASURED SLOWER (profiles/r02_ab_batched_decode_skinny.txt, 32 sequences: 7.18 ms per step against 6.24 ms with the stream-K split;
// launch list: the BN = 32 o_proj / down GEMMs average 47 us against 34 us): a 32-row weight box is 4 KB, so the nine stages hold 72 KB of
// real data per SM instead of 192 KB and the TMA request rate, not the byte rate, bounds the stream - the cost model above is wrong about
// small boxes.  The stream-K split therefore stays the default; SRGPT_GEMM_TSK_WHOLE=1 keeps this configuration reproducible.
template <int EPI>
static int launch_skinny(const void* A, int lda, const void* W, int ldw, const Params& p, cudaStream_t stream, bool* handled) {
  static const int whole_env = env_int("SRGPT_GEMM_TSK_WHOLE");
  *handled = false;
  if (whole_env <= 0 || p.M > BM) return SRGPT_OK;  // opt-in (SRGPT_GEMM_TSK_WHOLE=1): measured slower, see above
  const int sms = sm_count();
  const double hbm_s = (double)p.N * p.K * 2 / 6.4e12;
  int best_bn = 0, best_grid = 0;
  double best_t = 1e30;
  const int cands[3] = {128, 64, 32};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (EPI == SRGPT_EPI_SWIGLU && bn < 64) continue;  // keep (gate, up) pairs and 16-byte output vectors inside a chunk
    const int tiles = ceil_div(p.N, bn);
    const int waves = ceil_div(tiles, sms);
    const double per_sm = (double)waves * ((double)bn + (double)(p.M < 32 ? 32 : p.M)) * p.K * 2 / 88e9;
    const double t = (hbm_s > per_sm ? hbm_s : per_sm) * (1.0 + 0.02 * i);  // ties -> the wider tile
    if (t < best_t) { best_t = t; best_bn = bn; best_grid = ceil_div(tiles, waves); }
  }
  *handled = true;
  if (best_bn == 128) return launch_tall_sk_mt<EPI, 1, 128>(A, lda, W, ldw, p, stream, best_grid);
  if (best_bn == 64) return launch_tall_sk_mt<EPI, 1, 64>(A, lda, W, ldw, p, stream, best_grid);
  return launch_tall_sk_mt<EPI, 1, 32>(A, lda, W, ldw, p, stream, best_grid);
}

template <int EPI>
static int launch_tall_sk(const void* A, int lda, const void* W, int ldw, const Params& p, cudaStream_t stream) {
  const int mt = ceil_div(p.M, BM);
  if (mt == 1) {
    bool handled = false;
    const int rc = launch_skinny<EPI>(A, lda, W, ldw, p, stream, &handled);
    if (handled) return rc;
    return launch_tall_sk_mt<EPI, 1, TSK_BN>(A, lda, W, ldw, p, stream, 0);
  }
  if (mt == 2) return launch_tall_sk_mt<EPI, 2, TSK_BN>(A, lda, W, ldw, p, stream, 0);
  return launch_tall_sk_mt<EPI, 3, TSK_BN>(A, lda, W, ldw, p, stream, 0);
}

template <int EPI>
static int launch(const void* A, int lda, const void* W, int ldw, const Params& p, cudaStream_t stream) {
  // short prompts: all rows in one CTA, activation tile multicast across a 4-CTA cluster.  Implemented, parity-tested and
  // MEASURED SLOWER (profiles/r01_microbench_gemm_tall.jsonl: 259x4096x14336 147 vs 82 us, TTFT 17.6 vs 13.4 ms): multicast
  // removes L2 reads but every SM still has to ingest the whole 48 KB activation tile per k-block, and the per-SM TMA ingest
  // rate (~45 B/clk) is what bounds these kernels, not L2 read bandwidth.  Opt-in: SRGPT_GEMM_TALL=1.
  // short prompts, default: the tall stream-K kernel (all rows in one CTA, k-ranges balanced over the SMs) whenever the caller
  // registered a workspace and the weight matrix is big enough to be worth streaming (>= 4 MB); SRGPT_GEMM_TSK=-1 turns it off
  static const int tsk_env = env_int("SRGPT_GEMM_TSK");
  // MEASURED (profiles/r02_ab_gemm_tsk.txt): at 128 < M <= 384 the tall tiles LOSE to the default 128-row tiles (259 x 6144 x 4096:
  // 69 vs 37 us, gate/up 117 vs 97 us, down 118 vs 85 us; TTFT 16.7 vs 12.4 ms) - three 64 KB stages keep fewer bytes in flight than
  // six 32 KB ones, a third of every A stage is zero fill, and the un-overlapped fix-up epilogue stalls the single accumulator set -
  // so that range is opt-in (SRGPT_GEMM_TSK=1).  At M <= 128 (one accumulator, 32 KB stages, the batched-decode projections and the
  // batched lm_head) stream-K is what fills the chip: N / 128 tiles alone leave most SMs idle.
  const bool tsk_default = p.M <= BM && (long long)p.N * p.K * 2 >= (4LL << 20);
  if (tsk_env >= 0 && g_tsk.base != nullptr && p.M <= TALL_MT * BM && (tsk_env > 0 || tsk_default) && g_tsk.bytes >= tsk_workspace_bytes(8))
    return launch_tall_sk<EPI>(A, lda, W, ldw, p, stream);
  static const int tall_on = env_int("SRGPT_GEMM_TALL");
  if (tall_on && p.M > BM && p.M <= TALL_MT * BM && !p.out_fp32) {
    // 128-column tiles when they give every SM a tile, else 64-column tiles (more CTAs pulling weights)
    if (ceil_div(p.N, 128) >= sm_count() - 4) return launch_tall<EPI, 128>(A, lda, W, ldw, p, stream);
    return launch_tall<EPI, 64>(A, lda, W, ldw, p, stream);
  }
  // staged epilogue (shared memory + TMA stores): every bf16 output except SwiGLU; SRGPT_GEMM_DIRECT_EPI=1 restores the
  // row-per-thread global stores (A/B knob).  The residual tile is TMA-loaded too in the pair kernel unless its rows are
  // broadcast (res_row_mod, the position-embedding add of the patch GEMM) or SRGPT_GEMM_NO_RESTMA=1.
  static const int direct_epi = env_int("SRGPT_GEMM_DIRECT_EPI"), no_restma = env_int("SRGPT_GEMM_NO_RESTMA");
  Params ps = p;
  ps.staged = (!direct_epi && !p.out_fp32 && EPI != SRGPT_EPI_SWIGLU) ? 1 : 0;
  ps.res_tma = 0;
  CUtensorMap tc, tr;
  if (ps.staged) {
    int rc = make_tmap_out(&tc, p.C, p.M, p.N, p.ldc);
    if (rc != SRGPT_OK) return rc;
  } else {
    tc = CUtensorMap{};
  }
  tr = tc;
  // CTA-pair kernel (cta_group::2, 256 x 256 tiles per pair) for large problems
  static const int pair_env = env_int("SRGPT_GEMM_PAIR");  // 1: on where eligible, -1: off, 0: default
  // default: at least two waves of 256 x 256 pair tiles and >= 1024 rows (measured, profiles/r01_microbench_gemm_pair.txt:
  // 8288 x 28672 x 4096 0.83 -> 0.90 of the cuBLAS peak, 8192^3 0.85 -> 0.92, ViT qkv 0.67 -> 0.70; M = 259 is faster on
  // the 1-CTA kernel).  SRGPT_GEMM_PAIR_GELU=-1 keeps the GELU-tanh epilogue (SigLIP fc1) off the pair kernel as in round 1.
  static const int pair_gelu = env_int("SRGPT_GEMM_PAIR_GELU");
  const long long pair_units = (long long)ceil_div(ceil_div(p.M, BM), 2) * ceil_div(p.N, PAIR_BN);
  const bool pair_default = PAIR_DEFAULT && p.M >= 1024 && pair_units >= 2 * (sm_count() / 2) &&
                            (EPI != SRGPT_EPI_BIAS_GELU_TANH || pair_gelu >= 0);
  if (pair_env >= 0 && (pair_env > 0 || pair_default)) {
    CUtensorMap ta, tb;
    int rc = make_tmap(&ta, A, p.M, p.K, lda, BM);
    if (rc != SRGPT_OK) return rc;
    rc = make_tmap(&tb, W, p.N, p.K, ldw, PAIR_BN / 2);
    if (rc != SRGPT_OK) return rc;
    Params pg = ps;
    const int tiles_mu = ceil_div(ceil_div(p.M, BM), 2);
    pg.gm = tiles_mu;
    if (2.0 * p.M * p.K > 80e6) {
      const int g = (int)(40e6 / (2.0 * 2 * BM * p.K));
      pg.gm = g < 1 ? 1 : (g < tiles_mu ? g : tiles_mu);
    }
    pg.l2_prefetch = 0;
    const bool ew16 = EpiWarps<EPI>::N == 16 && epi_warps_env() == 16;
    if (ew16) pg.staged = 0;
    constexpr bool HAS_RES = EPI == SRGPT_EPI_BIAS_RESIDUAL;
    const bool res_tma = HAS_RES && pg.staged && !no_restma && p.residual != nullptr && p.res_row_mod == 0;
    if (res_tma) {
      rc = make_tmap_out(&tr, p.residual, p.M, p.N, p.ldr);
      if (rc != SRGPT_OK) return rc;
      pg.res_tma = 1;
    }
    static bool configured = false;
    if (!configured) {
      SRGPT_CHECK_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<EPI, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<false>::SMEM_BYTES));
      SRGPT_CHECK_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<EPI, EpiWarps<EPI>::N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            PairCfg<false>::SMEM_BYTES));
      if (HAS_RES)
        SRGPT_CHECK_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<EPI, 8, HAS_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<HAS_RES>::SMEM_BYTES));
      configured = true;
    }
    const int max_clusters = sm_count() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((pair_units < max_clusters ? pair_units : max_clusters) * 2));
    cfg.blockDim = dim3(64 + 32 * (ew16 ? 16 : 8));
    cfg.dynamicSmemBytes = res_tma ? PairCfg<HAS_RES>::SMEM_BYTES : PairCfg<false>::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (ew16) SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_pair_kernel<EPI, EpiWarps<EPI>::N, false>, ta, tb, tc, tr, pg));
    else if (res_tma) SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_pair_kernel<EPI, 8, HAS_RES>, ta, tb, tc, tr, pg));
    else SRGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_pair_kernel<EPI, 8, false>, ta, tb, tc, tr, pg));
    return SRGPT_OK;
  }
  int bn, cl;
  pick_cfg(p.M, p.N, &bn, &cl);
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, A, p.M, p.K, lda, BM);
  if (rc != SRGPT_OK) return rc;
  rc = make_tmap(&tb, W, p.N, p.K, ldw, bn / cl);  // each CTA of a cluster loads (and multicasts) its share of the B rows
  if (rc != SRGPT_OK) return rc;
  const int tiles_mu = ceil_div(ceil_div(p.M, BM), cl);
  const int units = tiles_mu * ceil_div(p.N, bn);
  // rasterisation group: the whole M when the activation fits L2, else as many m-units as keep a group's rows <= 40 MB
  Params pg = ps;
  static const int gm_env = env_int("SRGPT_GEMM_GM");
  const double a_bytes = 2.0 * p.M * p.K;
  pg.gm = tiles_mu;
  if (a_bytes > 80e6) {
    const int g = (int)(40e6 / (2.0 * cl * BM * p.K));
    pg.gm = g < 1 ? 1 : (g < tiles_mu ? g : tiles_mu);
  }
  if (gm_env > 0) pg.gm = gm_env < tiles_mu ? gm_env : tiles_mu;
  // MEASURED HARMFUL (profiles/r01_microbench_gemm_l2pf.txt: 8288x28672x4096 0.81 -> 0.67 of peak, ViT fc2 0.51 -> 0.29 with
  // a distance of 8 or 16 k-blocks): the prefetch requests compete with the loads for the same L2 -> SM path.  Opt-in only.
  static const int pf_env = env_int("SRGPT_GEMM_L2PF");  // > 0: distance in k-blocks
  pg.l2_prefetch = pf_env > 0 ? pf_env : 0;
  const int max_clusters = sm_count() / cl;
  const int grid = (units < max_clusters ? units : max_clusters) * cl;
  if (cl == 2) return bn == 256 ? launch_cfg<EPI, 256, 2>(ta, tb, tc, pg, grid, stream) : launch_cfg<EPI, 128, 2>(ta, tb, tc, pg, grid, stream);
  return bn == 256 ? launch_cfg<EPI, 256, 1>(ta, tb, tc, pg, grid, stream) : launch_cfg<EPI, 128, 1>(ta, tb, tc, pg, grid, stream);
}

}  // namespace gemm
}  // namespace srgpt

using namespace srgpt;

extern "C" __attribute__((visibility("default"))) long long srgpt_gemm_workspace_bytes(void) {
  int n = sm_count();
  if (n <= 0 || n > gemm::TSK_MAX_CTAS) n = gemm::TSK_MAX_CTAS;
  return gemm::tsk_workspace_bytes(n);
}

extern "C" __attribute__((visibility("default"))) int srgpt_gemm_set_workspace(void* workspace, long long bytes) {
  if (workspace == nullptr || bytes <= 0) {  // unregister
    gemm::g_tsk = gemm::TskState{};
    return SRGPT_OK;
  }
  SRGPT_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0 && bytes >= gemm::tsk_workspace_bytes(8));
  gemm::g_tsk.base = workspace;
  gemm::g_tsk.bytes = bytes;
  gemm::g_tsk.epoch = 0;  // the caller hands over ZEROED memory
  return SRGPT_OK;
}

extern "C" __attribute__((visibility("default"))) int srgpt_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                               const void* bias, const void* residual, int ldr, int res_row_mod, int epilogue,
                               int out_fp32, void* stream) {
  SRGPT_CHECK_ARG(A != nullptr && W != nullptr && C != nullptr);
  SRGPT_CHECK_ARG(M > 0 && N > 0 && K > 0);
  SRGPT_CHECK_ARG(lda >= K && ldw >= K);
  SRGPT_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0);  // 16-byte global strides for TMA
  SRGPT_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  SRGPT_CHECK_ARG((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  SRGPT_CHECK_ARG(epilogue >= SRGPT_EPI_NONE && epilogue <= SRGPT_EPI_BIAS_QUICK_GELU);
  if (epilogue == SRGPT_EPI_SWIGLU) {
    SRGPT_CHECK_ARG((N % 2) == 0 && !out_fp32 && ldc >= N / 2 && (ldc % 8) == 0);
  } else {
    SRGPT_CHECK_ARG(ldc >= N);
    SRGPT_CHECK_ARG(out_fp32 || (ldc % 8) == 0);
  }
  if (residual != nullptr) {
    SRGPT_CHECK_ARG(epilogue == SRGPT_EPI_BIAS_RESIDUAL && ldr >= N && (ldr % 8) == 0);
    SRGPT_CHECK_ARG((reinterpret_cast<uintptr_t>(residual) & 15) == 0);
  }
  if (bias != nullptr) SRGPT_CHECK_ARG((reinterpret_cast<uintptr_t>(bias) & 15) == 0);

  gemm::Params p;
  p.gm = 0; p.l2_prefetch = 0; p.staged = 0; p.res_tma = 0;
  static const bool no_respf = getenv("SRGPT_GEMM_NO_RESPF") != nullptr && getenv("SRGPT_GEMM_NO_RESPF")[0] == '1';
  p.res_prefetch = no_respf ? 0 : 1;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc;
  p.bias = reinterpret_cast<const bf16*>(bias);
  p.residual = reinterpret_cast<const bf16*>(residual);
  p.ldr = ldr; p.res_row_mod = res_row_mod;
  p.C = C; p.epilogue = epilogue; p.out_fp32 = out_fp32;

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (epilogue) {
    case SRGPT_EPI_NONE: return gemm::launch<SRGPT_EPI_NONE>(A, lda, W, ldw, p, st);
    case SRGPT_EPI_BIAS: return gemm::launch<SRGPT_EPI_BIAS>(A, lda, W, ldw, p, st);
    case SRGPT_EPI_BIAS_GELU_TANH: return gemm::launch<SRGPT_EPI_BIAS_GELU_TANH>(A, lda, W, ldw, p, st);
    case SRGPT_EPI_BIAS_GELU_ERF: return gemm::launch<SRGPT_EPI_BIAS_GELU_ERF>(A, lda, W, ldw, p, st);
    case SRGPT_EPI_BIAS_RESIDUAL: return gemm::launch<SRGPT_EPI_BIAS_RESIDUAL>(A, lda, W, ldw, p, st);
    case SRGPT_EPI_SWIGLU: return gemm::launch<SRGPT_EPI_SWIGLU>(A, lda, W, ldw, p, st);
    case SRGPT_EPI_BIAS_QUICK_GELU: return gemm::launch<SRGPT_EPI_BIAS_QUICK_GELU>(A, lda, W, ldw, p, st);
  }
  return SRGPT_ERR_INVALID;
}

// Prefill attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), warp-specialised, persistent.
//
// Replaces SigLIP's eager attention (HF SiglipAttention, head_dim 72, non-causal; call site
// llava/model/multimodal_encoder/vision_encoder.py:119-130) and flash_attn_func / flash_attn_varlen_func
// (modeling_llama.py:540-566; head_dim 128, causal, GQA) for the prompt rows.  The mma.sync kernel in attention.cu
// (110 TFLOP/s) was 73 of the 120 ms of a 64-image tower pass; this is its tensor-memory successor.
//
// One CTA per SM loops over (q tile of 128 rows, head, sequence) work items:
//   warp 0      TMA producer: Q tile once per item, K / V tiles through 2-stage rings.  Tiles come straight out of the
//               fused qkv activation through 3-D tensor maps (channel, head, row): a head of 72 channels is a 64-channel
//               128B-swizzled box + a 16-channel 32B-swizzled box whose channels 72..79 are out of bounds -> zero filled,
//               so neither the GEMM output nor the weights need padding.
//   warp 1      tcgen05.mma issuer.  S = Q K^T (both K-major) into one of two TMEM S buffers; O_j = P_j V_j with P (bf16,
//               written by the softmax warps into a 128B-swizzled K-major tile) as A and the V tile [kv, channels] used
//               AS LOADED as an MN-major B operand (no transpose anywhere).  S_{j+1} is issued before P_j V_j, so the
//               tensor pipe works under the softmax of tile j.
//   warps 2..9  softmax: thread pair (w, w+4) owns one q row, each thread half of the kv columns.  Online softmax in
//               base 2 (ex2.approx on the SFU); every tile's O_j is a FRESH TMEM accumulator that the threads fold into
//               registers (o = o * alpha + O_j), so TMEM is never rescaled and no correction pass exists.
// Measured (B200, 64 images x 16 heads x 1024 tokens, head_dim 72): 0.826 ms = 374 TFLOP/s vs 1.55 ms for the mma.sync kernel
// (0.955 ms before the third K stage);
// ncu: ~1500 clk per 128x128 tile, XU (ex2) pipe 34 %, tensor pipe 21 %, issue slots 36 % - a latency chain per tile
// (TMEM load -> max -> exchange -> 64 ex2 -> pack -> st.shared -> fence -> arrive) on 2 warps per scheduler.  A variant
// with FOUR threads per row (16 softmax warps, 96 registers, two-pass S reads) measured 1.03 ms at head_dim 72 and
// 0.168 vs 0.196 ms at head_dim 128 / 32 prompts; the two-thread version is kept.  Next: ex2 emulation on the FMA pipe
// for a share of the columns and packed f32x2 arithmetic.
// UMMA descriptor encodings (MN-major, 32B swizzle, OOB fill) were verified with tools/umma_probe.
#include <stdlib.h>

#include "common.cuh"
#include "srgpt_b200.h"
#include "tcgen05.cuh"

namespace srgpt {
namespace attn_tc {

using namespace tc;

constexpr int BM = 128;
constexpr int SM_WARPS = 8;
constexpr int NTHREADS = 64 + 32 * SM_WARPS;
constexpr int TMEM_COLS = 512;
constexpr int O_COL = 256;  // S buffers at columns [0, 2*BN), O buffers at 256 and 384

constexpr int KST = 3;  // K ring depth (measured: 0.955 ms -> 0.826 ms against two stages; shared memory has no room for a third V stage)
enum Bar { Q_FULL = 0, Q_EMPTY = 2, K_FULL = 4, K_EMPTY = 7, V_FULL = 10, V_EMPTY = 12, S_FULL = 14, S_EMPTY = 16, P_FULL = 18, P_EMPTY = 20,
           O_FULL = 22, O_EMPTY = 24, NUM_BARS = 26 };

template <int HD, int BN>
struct Geo {
  static_assert(HD == 72 || HD == 128, "head_dim 72 or 128");
  static_assert(BN == 64 || BN == 128, "kv tile 64 or 128");
  static constexpr int W1 = HD == 72 ? 16 : 64;            // channels of the second chunk (the first has 64)
  static constexpr bool TAIL32 = W1 == 16;                 // second chunk is a 32B-swizzled box
  static constexpr int ROWB1 = TAIL32 ? 32 : 128;          // bytes per row of chunk 1
  static constexpr uint64_t F1 = TAIL32 ? UMMA_DESC_SW32 : UMMA_DESC_SW128;
  static constexpr int Q_C0 = BM * 128, Q_BYTES = Q_C0 + BM * ROWB1;
  static constexpr int KV_C0 = BN * 128, KV_BYTES = KV_C0 + BN * ROWB1;
  static constexpr int P_BYTES = BM * BN * 2;
  static constexpr int OFF_K = 2 * Q_BYTES, OFF_V = OFF_K + KST * KV_BYTES, OFF_P = OFF_V + 2 * KV_BYTES, OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int OFF_RED = OFF_BAR + NUM_BARS * 8 + 16;
  static constexpr int SMEM_BYTES = OFF_RED + 3 * 2 * BM * 4 + 1024;  // + alignment slack
  static constexpr int V_ADV1 = 16 * ROWB1;                // bytes per UMMA_K step (16 kv rows) of V chunk 1
  static constexpr int NS = BN / 2;                        // S columns per softmax thread
  static constexpr int OREG = HD == 72 ? 40 : 64;          // O columns per softmax thread
};

struct Maps {
  CUtensorMap q0, q1, k0, k1, v0, v1;  // chunk 0 / chunk 1 boxes of the q, k, v views
};

struct Params {
  const int* cu_seqlens;  // null: `batch` sequences of `seqlen` rows back to back
  int seqlen, batch, n_heads, group, nqt;
  float scale_log2;
  bf16* out;
  int o_ld;
  int s_ahead;  // tiles the S = Q K^T stream may run ahead of P V (1 or 2)
};

struct Work {
  int row_base, seqlen, q0, head, n_tiles;
};

// Work items of this CTA in order: w = blockIdx.x, + gridDim.x, ...; items whose q tile starts past the sequence end are skipped.
template <int BN, bool CAUSAL>
struct WorkIter {
  const Params& p;
  int w, n_work;
  Work k;
  __device__ WorkIter(const Params& p_, int n_work_) : p(p_), w((int)blockIdx.x - (int)gridDim.x), n_work(n_work_) { k.n_tiles = 0; }
  __device__ bool next() {
    while (true) {
      w += gridDim.x;
      if (w >= n_work) return false;
      const int qt = w % p.nqt;
      const int hb = w / p.nqt;
      k.head = hb % p.n_heads;
      const int b = hb / p.n_heads;
      if (p.cu_seqlens != nullptr) {
        k.row_base = p.cu_seqlens[b];
        k.seqlen = p.cu_seqlens[b + 1] - k.row_base;
      } else {
        k.row_base = b * p.seqlen;
        k.seqlen = p.seqlen;
      }
      k.q0 = qt * BM;
      if (k.q0 >= k.seqlen) continue;
      const int kv_end = CAUSAL ? min(k.seqlen, k.q0 + BM) : k.seqlen;
      k.n_tiles = (kv_end + BN - 1) / BN;
      return true;
    }
  }
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

template <int HD, int BN, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS, 1) attn_fwd_tc_kernel(const __grid_constant__ Maps maps, const Params p) {
  using G = Geo<HD, BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_work = p.nqt * p.n_heads * p.batch;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + G::OFF_BAR;
  auto bar = [&](int which) { return bar0 + which * 8; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.q0); prefetch_tmap(&maps.q1); prefetch_tmap(&maps.k0);
    prefetch_tmap(&maps.k1); prefetch_tmap(&maps.v0); prefetch_tmap(&maps.v1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(Q_FULL + i), 1);
      mbar_init(bar(Q_EMPTY + i), 1);
      mbar_init(bar(V_FULL + i), 1);
      mbar_init(bar(V_EMPTY + i), 1);
      mbar_init(bar(S_FULL + i), 1);
      mbar_init(bar(S_EMPTY + i), SM_WARPS);
      mbar_init(bar(P_FULL + i), SM_WARPS);
      mbar_init(bar(P_EMPTY + i), 1);
      mbar_init(bar(O_FULL + i), 1);
      mbar_init(bar(O_EMPTY + i), SM_WARPS);
    }
    for (int i = 0; i < KST; ++i) {
      mbar_init(bar(K_FULL + i), 1);
      mbar_init(bar(K_EMPTY + i), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t t = 0, it = 0;
      WorkIter<BN, CAUSAL> wi(p, n_work);
      while (wi.next()) {
        const Work& k = wi.k;
        const int kvh = k.head / p.group;
        const int qi = it & 1;
        mbar_wait(bar(Q_EMPTY + qi), ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(bar(Q_FULL + qi), G::Q_BYTES);
        tma_load_3d(sbase + qi * G::Q_BYTES, &maps.q0, bar(Q_FULL + qi), 0, k.head, k.row_base + k.q0);
        tma_load_3d(sbase + qi * G::Q_BYTES + G::Q_C0, &maps.q1, bar(Q_FULL + qi), 64, k.head, k.row_base + k.q0);
        for (int j = 0; j < k.n_tiles; ++j, ++t) {
          const int i = t & 1;
          const uint32_t ph = (t >> 1) & 1;
          const int row = k.row_base + j * BN;
          const int ik = t % KST;
          const uint32_t phk = (t / KST) & 1;
          mbar_wait(bar(K_EMPTY + ik), phk ^ 1);
          mbar_expect_tx(bar(K_FULL + ik), G::KV_BYTES);
          tma_load_3d(sbase + G::OFF_K + ik * G::KV_BYTES, &maps.k0, bar(K_FULL + ik), 0, kvh, row);
          tma_load_3d(sbase + G::OFF_K + ik * G::KV_BYTES + G::KV_C0, &maps.k1, bar(K_FULL + ik), 64, kvh, row);
          mbar_wait(bar(V_EMPTY + i), ph ^ 1);
          mbar_expect_tx(bar(V_FULL + i), G::KV_BYTES);
          tma_load_3d(sbase + G::OFF_V + i * G::KV_BYTES, &maps.v0, bar(V_FULL + i), 0, kvh, row);
          tma_load_3d(sbase + G::OFF_V + i * G::KV_BYTES + G::KV_C0, &maps.v1, bar(V_FULL + i), 64, kvh, row);
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Tile stream t = 0, 1, ... runs across work items.  Order: S_0, then for every tile t: S_{t+1} (possibly the first tile of
    // the NEXT item, from the other Q buffer), P_t V_t - the tensor pipe always has the next S queued under the softmax of t.
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BM, BN);
      constexpr uint32_t idesc_v0 = umma_idesc_bf16(BM, 64, 0, 1);
      constexpr uint32_t idesc_v1 = umma_idesc_bf16(BM, G::W1, 0, 1);
      auto issue_s = [&](uint32_t tt, uint32_t item, bool first_of_item, bool last_of_item) {
        const int i = tt & 1, qi = item & 1;
        const uint32_t ph = (tt >> 1) & 1;
        const int ik = tt % KST;
        if (first_of_item) mbar_wait(bar(Q_FULL + qi), (item >> 1) & 1);
        mbar_wait(bar(K_FULL + ik), (tt / KST) & 1);
        mbar_wait(bar(S_EMPTY + i), ph ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + i * BN;
        const uint32_t qa = sbase + qi * G::Q_BYTES, ka = sbase + G::OFF_K + ik * G::KV_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(d, umma_desc(UMMA_DESC_SW128, qa + k * 32), umma_desc(UMMA_DESC_SW128, ka + k * 32), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < G::W1 / 16; ++k)
          umma_f16(d, umma_desc(G::F1, qa + G::Q_C0 + k * 32), umma_desc(G::F1, ka + G::KV_C0 + k * 32), idesc_s, 1u);
        umma_commit(bar(S_FULL + i));
        umma_commit(bar(K_EMPTY + ik));
        if (last_of_item) umma_commit(bar(Q_EMPTY + qi));
      };
      auto issue_pv = [&](uint32_t tt) {
        const int i = tt & 1;
        const uint32_t ph = (tt >> 1) & 1;
        mbar_wait(bar(P_FULL + i), ph);
        mbar_wait(bar(V_FULL + i), ph);
        mbar_wait(bar(O_EMPTY + i), ph ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + O_COL + i * 128;
        const uint32_t pa = sbase + G::OFF_P + i * G::P_BYTES, va = sbase + G::OFF_V + i * G::KV_BYTES;
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {
          const uint64_t a = umma_desc(UMMA_DESC_SW128, pa + (ks >> 2) * (BM * 128) + (ks & 3) * 32);
          umma_f16(d, a, umma_desc(UMMA_DESC_SW128, va + ks * 2048), idesc_v0, ks != 0 ? 1u : 0u);
          umma_f16(d + 64, a, umma_desc(G::F1, va + G::KV_C0 + ks * G::V_ADV1), idesc_v1, ks != 0 ? 1u : 0u);
        }
        umma_commit(bar(O_FULL + i));
        umma_commit(bar(P_EMPTY + i));
        umma_commit(bar(V_EMPTY + i));
      };
      // two cursors over the same tile stream: S = Q K^T is issued p.s_ahead tiles ahead of P V.  Measured in one run
      // (64 images, head_dim 72): one tile ahead 0.826 ms, two tiles ahead (S_{t+2} queued as soon as the softmax warps
      // have pulled S_t out of TMEM) 1.19 ms -> one is the default (SRGPT_ATTN_S_AHEAD=2 selects the other).  The K ring has
      // three stages: with two, 0.955 ms.
      WorkIter<BN, CAUSAL> ws(p, n_work), wp(p, n_work);
      uint32_t ts = 0, its = 0;   // S cursor: global tile index, item index
      int js = 0;                 // tile inside the S cursor's item
      bool s_have = ws.next();
      uint32_t tp = 0;            // P V cursor
      bool p_have = wp.next();
      int jp = 0;
      while (p_have) {
        while (s_have && ts <= tp + p.s_ahead) {  // s_ahead = 2: tiles tp+1 and tp+2 are in the two S buffers while P_tp V_tp is issued
          issue_s(ts, its, js == 0, js + 1 == ws.k.n_tiles);
          ++ts;
          if (++js == ws.k.n_tiles) {
            js = 0;
            ++its;
            s_have = ws.next();
          }
        }
        issue_pv(tp);
        ++tp;
        if (++jp == wp.k.n_tiles) {
          jp = 0;
          p_have = wp.next();
        }
      }
    }
  } else {
    // ===================== softmax warps (2..9) =====================
    const int lg = warp & 3;            // TMEM lane group of this warp
    const int half = (warp - 2) >> 2;   // which half of the kv columns / O columns this thread owns
    const int r = lg * 32 + lane;       // q row inside the tile
    const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
    const float sl = p.scale_log2;
    const uint32_t red_max = sbase + G::OFF_RED;          // [2 tile parities][2 halves][128 rows] floats
    const uint32_t red_l = red_max + 2 * 2 * BM * 4;      // [2 halves][128 rows]
    const uint32_t p_row = sbase + G::OFF_P + (r >> 3) * 1024 + (r & 7) * 128;

    float o[G::OREG];
#pragma unroll
    for (int c = 0; c < G::OREG; ++c) o[c] = 0.f;

    auto fold_o = [&](uint32_t tt, float alpha) {  // o = o * alpha + O_tt (this thread's columns)
      const int i = tt & 1;
      mbar_wait(bar(O_FULL + i), (tt >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t base = tmem_base + O_COL + i * 128 + lane_addr;
      uint32_t v[G::OREG];
      if (HD == 72) {
        if (half == 0) {
          tmem_ld_32x32b_x32(base, v);
          tmem_ld_32x32b_x8(base + 64, v + 32);  // channels 64..71 (the 16-wide tail accumulator)
        } else {
          tmem_ld_32x32b_x32(base + 32, v);
        }
      } else {
        tmem_ld_32x32b_x32(base + half * 64, v);
        tmem_ld_32x32b_x32(base + half * 64 + 32, v + 32);
      }
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(O_EMPTY + i));
      const int n = (HD == 72 && half == 1) ? 32 : G::OREG;
#pragma unroll
      for (int c = 0; c < G::OREG; ++c)
        if (c < n) o[c] = fmaf(o[c], alpha, __uint_as_float(v[c]));
    };
    auto finalize = [&](int row_base, int seqlen, int q0, int head, float l_run) {  // O / l -> global; then o = 0
      st_shared_f32(red_l + (half * BM + r) * 4, l_run);
      named_bar_sync(1 + lg, 64);
      const float l_tot = l_run + ld_shared_f32(red_l + ((half ^ 1) * BM + r) * 4);
      named_bar_sync(1 + lg, 64);  // both halves have read before the next item's finalize overwrites
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      if (q0 + r < seqlen) {
        bf16* orow = p.out + (size_t)(row_base + q0 + r) * p.o_ld + head * HD;
#pragma unroll
        for (int c = 0; c < G::OREG; ++c) o[c] *= inv;
        if (HD == 72) {
          if (half == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(orow + c * 8) = pack8(o + c * 8);
            *reinterpret_cast<uint4*>(orow + 64) = pack8(o + 32);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(orow + 32 + c * 8) = pack8(o + c * 8);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(orow + half * 64 + c * 8) = pack8(o + c * 8);
        }
      }
#pragma unroll
      for (int c = 0; c < G::OREG; ++c) o[c] = 0.f;
    };

    uint32_t t = 0;
    bool pending = false;  // the previous item still has its last P V product to fold and its rows to store
    int pv_row_base = 0, pv_seqlen = 0, pv_q0 = 0, pv_head = 0;
    float pv_l = 0.f, pv_alpha = 0.f;
    WorkIter<BN, CAUSAL> wi(p, n_work);
    while (wi.next()) {
      const Work k = wi.k;
      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
      for (int j = 0; j < k.n_tiles; ++j, ++t) {
        const int i = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        // ---- S tile -> registers
        mbar_wait(bar(S_FULL + i), ph);
        tcgen05_fence_after();
        uint32_t sv[G::NS];
#pragma unroll
        for (int c = 0; c < G::NS; c += 32) tmem_ld_32x32b_x32(tmem_base + i * BN + half * G::NS + c + lane_addr, sv + c);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(S_EMPTY + i));
        float s[G::NS];
#pragma unroll
        for (int c = 0; c < G::NS; ++c) s[c] = __uint_as_float(sv[c]);
        // ---- masks: columns past the sequence end, and the causal diagonal tile
        if (j * BN + BN > k.seqlen || (CAUSAL && j * BN + BN - 1 > k.q0)) {
          const int kv0 = j * BN + half * G::NS, qi = k.q0 + r;
#pragma unroll
          for (int c = 0; c < G::NS; ++c) {
            const int kv = kv0 + c;
            if (kv >= k.seqlen || (CAUSAL && kv > qi)) s[c] = -INFINITY;
          }
        }
        // ---- row max (two threads per row exchange through shared memory); 4 independent chains
        float mx4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
        for (int c = 4; c < G::NS; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], s[c]);
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const uint32_t rm = red_max + (t & 1) * (2 * BM * 4);
        st_shared_f32(rm + (half * BM + r) * 4, mx);
        named_bar_sync(1 + lg, 64);
        mx = fmaxf(mx, ld_shared_f32(rm + ((half ^ 1) * BM + r) * 4));
        const float m_new = fmaxf(m_run, mx);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = ex2_approx((m_run - m_use) * sl);  // m_run = -inf -> 0
        m_run = m_new;
        const float msl = m_use * sl;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < G::NS; ++c) {
          s[c] = ex2_approx(fmaf(s[c], sl, -msl));
          sum4[c & 3] += s[c];
        }
        l_run = fmaf(l_run, alpha, (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
        // ---- P (bf16) -> 128B-swizzled K-major tile: row r, 16-byte chunk c of k-block kb
        mbar_wait(bar(P_EMPTY + i), ph ^ 1);
#pragma unroll
        for (int cc = 0; cc < G::NS / 8; ++cc) {
          const int col = half * G::NS + cc * 8;
          const int kb = col >> 6, cin = (col & 63) >> 3;
          st_shared_v4(p_row + i * G::P_BYTES + kb * (BM * 128) + ((cin ^ (r & 7)) << 4), pack8(s + cc * 8));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(P_FULL + i));
        // ---- fold the previous tile's P V product while the tensor pipe works on this one
        if (j > 0) {
          fold_o(t - 1, alpha_prev);
        } else if (pending) {  // ... which, on the first tile of an item, completes the PREVIOUS item
          fold_o(t - 1, pv_alpha);
          finalize(pv_row_base, pv_seqlen, pv_q0, pv_head, pv_l);
          pending = false;
        }
        alpha_prev = alpha;
      }
      pending = true;
      pv_row_base = k.row_base; pv_seqlen = k.seqlen; pv_q0 = k.q0; pv_head = k.head;
      pv_l = l_run; pv_alpha = alpha_prev;
    }
    if (pending) {
      fold_o(t - 1, pv_alpha);
      finalize(pv_row_base, pv_seqlen, pv_q0, pv_head, pv_l);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// Round 2: PING-PONG kernel for the SigLIP tower (head_dim 72, non-causal, dense batch).
//
// The kernel above keeps ONE q tile in flight: its eight softmax warps walk the per-tile chain S -> max -> (pair exchange) -> ex2 -> P
// -> fold in lock step, so nothing hides their latencies (ncu: tensor pipe 20 %, XU 33 %, issue 35 %; ~3000 clk per 128 x 128 tile
// against 640 clk of MMA and 1024 clk of ex2).  Here a CTA works on a PAIR of q tiles (256 query rows of one head) against ONE K / V
// stream:
//   * softmax group A (warps 2-5) owns q tile A, group B (warps 6-9) q tile B: ONE thread per query row (TMEM lane = row), so there
//     is no cross-thread max exchange and no named barrier; while one group waits (S from the tensor pipe, its TMEM loads, the SFU),
//     the other group's warps on the same schedulers run - the two chains interleave instead of idling;
//   * every K / V tile is loaded once and used by both q tiles (half the TMA / shared-memory traffic per tile product);
//   * two passes over S in TMEM (pass 1: row max; pass 2: ex2 + bf16 P) keep 32 S values live instead of 128, so the 72 fp32
//     output channels of the row fit in registers next to them; O_j is still a FRESH accumulator folded into registers
//     (o = o * alpha + O_j), between the two passes, so P V of tile j never waits for the fold of tile j - 1;
//   * TMEM: S_A | S_B | O_A | O_B = 4 x 128 columns; shared memory: Q_A Q_B (40 KB) + 3 K + 2 V stages (100 KB) + P_A P_B (64 KB).
// MMA issue order per kv tile j:  P_A V_j -> S_A(j+1) -> P_B V_j -> S_B(j+1).
// ---------------------------------------------------------------------------------------------
namespace pp {
constexpr int HD = 72, BN = 128;
using G = Geo<HD, BN>;
enum PBar { PQ_FULL = 0, PQ_EMPTY = 1, PK_FULL = 2, PK_EMPTY = 5, PV_FULL = 8, PV_EMPTY = 10, PS_FULL = 12, PS_EMPTY = 14, PP_FULL = 16, PP_EMPTY = 18,
            PO_FULL = 20, PO_EMPTY = 22, P_NUM_BARS = 24 };
constexpr int OFF_Q = 0, OFF_K = 2 * G::Q_BYTES, OFF_V = OFF_K + KST * G::KV_BYTES, OFF_P = OFF_V + 2 * G::KV_BYTES, OFF_BAR = OFF_P + 2 * G::P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + P_NUM_BARS * 8 + 16 + 1024;

struct Item {
  int row_base, q0, head, n_tiles;
  bool b_active;
};
struct ItemIter {
  const Params& p;
  int w, n_work, n_pairs;
  Item k;
  __device__ ItemIter(const Params& p_) : p(p_), w((int)blockIdx.x - (int)gridDim.x) {
    n_pairs = (p.nqt + 1) >> 1;
    n_work = n_pairs * p.n_heads * p.batch;
  }
  __device__ bool next() {
    w += gridDim.x;
    if (w >= n_work) return false;
    const int pair = w % n_pairs, hb = w / n_pairs;
    k.head = hb % p.n_heads;
    k.row_base = (hb / p.n_heads) * p.seqlen;
    k.q0 = pair * 2 * BM;
    k.b_active = k.q0 + BM < p.seqlen;
    k.n_tiles = (p.seqlen + BN - 1) / BN;
    return true;
  }
};
}  // namespace pp

__global__ void __launch_bounds__(NTHREADS, 1) attn_vit_pp_kernel(const __grid_constant__ Maps maps, const Params p) {
  using namespace pp;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + P_NUM_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  auto bar = [&](int which) { return bar0 + which * 8; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.q0); prefetch_tmap(&maps.q1); prefetch_tmap(&maps.k0);
    prefetch_tmap(&maps.k1); prefetch_tmap(&maps.v0); prefetch_tmap(&maps.v1);
    mbar_init(bar(PQ_FULL), 1);
    mbar_init(bar(PQ_EMPTY), 1);
    for (int i = 0; i < KST; ++i) { mbar_init(bar(PK_FULL + i), 1); mbar_init(bar(PK_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(PV_FULL + i), 1);
      mbar_init(bar(PV_EMPTY + i), 1);
      mbar_init(bar(PS_FULL + i), 1);
      mbar_init(bar(PS_EMPTY + i), 4);   // the four warps of a softmax group
      mbar_init(bar(PP_FULL + i), 4);
      mbar_init(bar(PP_EMPTY + i), 1);
      mbar_init(bar(PO_FULL + i), 1);
      mbar_init(bar(PO_EMPTY + i), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t t = 0, it = 0;
      ItemIter wi(p);
      while (wi.next()) {
        const Item& k = wi.k;
        mbar_wait(bar(PQ_EMPTY), (it & 1) ^ 1);
        mbar_expect_tx(bar(PQ_FULL), (k.b_active ? 2 : 1) * G::Q_BYTES);
        tma_load_3d(sbase + OFF_Q, &maps.q0, bar(PQ_FULL), 0, k.head, k.row_base + k.q0);
        tma_load_3d(sbase + OFF_Q + G::Q_C0, &maps.q1, bar(PQ_FULL), 64, k.head, k.row_base + k.q0);
        if (k.b_active) {
          tma_load_3d(sbase + OFF_Q + G::Q_BYTES, &maps.q0, bar(PQ_FULL), 0, k.head, k.row_base + k.q0 + BM);
          tma_load_3d(sbase + OFF_Q + G::Q_BYTES + G::Q_C0, &maps.q1, bar(PQ_FULL), 64, k.head, k.row_base + k.q0 + BM);
        }
        for (int j = 0; j < k.n_tiles; ++j, ++t) {
          const int row = k.row_base + j * BN;
          const int ik = t % KST, iv = t & 1;
          mbar_wait(bar(PK_EMPTY + ik), ((t / KST) & 1) ^ 1);
          mbar_expect_tx(bar(PK_FULL + ik), G::KV_BYTES);
          tma_load_3d(sbase + OFF_K + ik * G::KV_BYTES, &maps.k0, bar(PK_FULL + ik), 0, k.head, row);
          tma_load_3d(sbase + OFF_K + ik * G::KV_BYTES + G::KV_C0, &maps.k1, bar(PK_FULL + ik), 64, k.head, row);
          mbar_wait(bar(PV_EMPTY + iv), ((t >> 1) & 1) ^ 1);
          mbar_expect_tx(bar(PV_FULL + iv), G::KV_BYTES);
          tma_load_3d(sbase + OFF_V + iv * G::KV_BYTES, &maps.v0, bar(PV_FULL + iv), 0, k.head, row);
          tma_load_3d(sbase + OFF_V + iv * G::KV_BYTES + G::KV_C0, &maps.v1, bar(PV_FULL + iv), 64, k.head, row);
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BM, BN);
      constexpr uint32_t idesc_v0 = umma_idesc_bf16(BM, 64, 0, 1);
      constexpr uint32_t idesc_v1 = umma_idesc_bf16(BM, G::W1, 0, 1);
      // Event-driven issue: four kinds of MMA batches are pending at any time - S_g(k) (needs K tile k and the S_g buffer back from the
      // softmax group) and P_g V(k) (needs P_g(k), V tile k and the O_g buffer back) for g = A, B.  Each is issued the moment its
      // barriers have flipped (non-blocking mbarrier.test_wait), in whatever order the two softmax groups get there: S_g(k+1) goes
      // out as soon as group g has pulled S_g(k) into registers - long before P_g(k) exists - so the next S is ready when the group
      // finishes the tile.  (First version: fixed order P_A V, S_A, P_B V, S_B with blocking waits; the profile showed the softmax
      // warps spinning on S_FULL for a third of all samples.)
      auto test = [&](int which, uint32_t parity) -> bool {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar(which)), "r"(parity) : "memory");
        return ok != 0;
      };
      uint32_t tg[2] = {0, 0};   // S batches issued so far per group (all items)
      uint32_t tpv[2] = {0, 0};  // P V batches issued so far per group
      uint32_t tk = 0, it = 0;   // kv tiles consumed by finished items (K / V ring position of tile 0 of the item), item counter
      auto issue_s = [&](int g, uint32_t tkk) {  // S_g = Q_g K^T of the kv tile at ring position tkk
        const int ik = tkk % KST;
        tcgen05_fence_after();
        const uint32_t d = tmem_base + g * BN;
        const uint32_t qa = sbase + OFF_Q + g * G::Q_BYTES, ka = sbase + OFF_K + ik * G::KV_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(d, umma_desc(UMMA_DESC_SW128, qa + k * 32), umma_desc(UMMA_DESC_SW128, ka + k * 32), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < G::W1 / 16; ++k)
          umma_f16(d, umma_desc(G::F1, qa + G::Q_C0 + k * 32), umma_desc(G::F1, ka + G::KV_C0 + k * 32), idesc_s, 1u);
        umma_commit(bar(PS_FULL + g));
        ++tg[g];
      };
      auto issue_pv = [&](int g, uint32_t tkk) {  // O_g = P_g V of the kv tile at ring position tkk (fresh accumulator)
        const int iv = tkk & 1;
        tcgen05_fence_after();
        const uint32_t d = tmem_base + O_COL + g * 128;
        const uint32_t pa = sbase + OFF_P + g * G::P_BYTES, va = sbase + OFF_V + iv * G::KV_BYTES;
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {
          const uint64_t a = umma_desc(UMMA_DESC_SW128, pa + (ks >> 2) * (BM * 128) + (ks & 3) * 32);
          umma_f16(d, a, umma_desc(UMMA_DESC_SW128, va + ks * 2048), idesc_v0, ks != 0 ? 1u : 0u);
          umma_f16(d + 64, a, umma_desc(G::F1, va + G::KV_C0 + ks * G::V_ADV1), idesc_v1, ks != 0 ? 1u : 0u);
        }
        umma_commit(bar(PO_FULL + g));
        umma_commit(bar(PP_EMPTY + g));
        ++tpv[g];
      };
      ItemIter wi(p);
      while (wi.next()) {
        const Item& k = wi.k;
        const int n = k.n_tiles;
        const int ng = k.b_active ? 2 : 1;
        mbar_wait(bar(PQ_FULL), it & 1);
        int s_done[2] = {0, 0}, pv_done[2] = {0, 0};  // batches issued inside this item
        if (ng == 1) { s_done[1] = n; pv_done[1] = n; }
        uint32_t spins = 0;
        while (pv_done[0] < n || pv_done[1] < n) {
          bool progressed = false;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (g >= ng) continue;
            if (s_done[g] < n) {
              const uint32_t tkk = tk + s_done[g];
              if (test(PK_FULL + tkk % KST, (tkk / KST) & 1) && test(PS_EMPTY + g, (tg[g] & 1) ^ 1)) {
                issue_s(g, tkk);
                ++s_done[g];
                if (s_done[g ^ 1] >= s_done[g]) umma_commit(bar(PK_EMPTY + tkk % KST));          // both q tiles have read K tile k
                if (s_done[0] == n && s_done[1] == n) umma_commit(bar(PQ_EMPTY));                 // the item's last S: Q may be reloaded
                progressed = true;
              }
            }
            if (pv_done[g] < s_done[g] || (pv_done[g] < n && s_done[g] == n)) {
              const uint32_t tkk = tk + pv_done[g];
              if (pv_done[g] < s_done[g] && test(PP_FULL + g, tpv[g] & 1) && test(PV_FULL + (tkk & 1), (tkk >> 1) & 1) &&
                  test(PO_EMPTY + g, (tpv[g] & 1) ^ 1)) {
                issue_pv(g, tkk);
                ++pv_done[g];
                if (pv_done[g ^ 1] >= pv_done[g]) umma_commit(bar(PV_EMPTY + (tkk & 1)));        // both q tiles have read V tile k
                progressed = true;
              }
            }
          }
          if (progressed) spins = 0;
          else if (++spins > (1u << 28)) __trap();  // deadlock breaker
        }
        tk += n;
        ++it;
      }
    }
  } else {
    // ===================== softmax groups: warps 2-5 = q tile A, warps 6-9 = q tile B; one thread per query row =====================
    const int g = (warp - 2) >> 2;
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
    const float sl = p.scale_log2;
    const uint32_t s_addr = tmem_base + g * BN + lane_addr;
    const uint32_t o_addr = tmem_base + O_COL + g * 128 + lane_addr;
    const uint32_t p_row = sbase + OFF_P + g * G::P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
    uint32_t t = 0;  // tiles processed by this group
    ItemIter wi(p);
    while (wi.next()) {
      const Item k = wi.k;
      if (g == 1 && !k.b_active) continue;  // no second q tile in this item: group B sits it out (the MMA warp skips it too)
      float o[72];
#pragma unroll
      for (int c = 0; c < 72; ++c) o[c] = 0.f;
      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
      auto fold = [&](uint32_t tt, float alpha) {  // o = o * alpha + O_tt, in three chunks so only 32 loaded values are live
        mbar_wait(bar(PO_FULL + g), tt & 1);
        tcgen05_fence_after();
        uint32_t v[32];
        tmem_ld_32x32b_x32(o_addr, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] = fmaf(o[c], alpha, __uint_as_float(v[c]));
        tmem_ld_32x32b_x32(o_addr + 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) o[32 + c] = fmaf(o[32 + c], alpha, __uint_as_float(v[c]));
        tmem_ld_32x32b_x8(o_addr + 64, v);  // channels 64..71 (the 16-wide tail accumulator; 72..79 are zero padding)
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(PO_EMPTY + g));
#pragma unroll
        for (int c = 0; c < 8; ++c) o[64 + c] = fmaf(o[64 + c], alpha, __uint_as_float(v[c]));
      };
      for (int j = 0; j < k.n_tiles; ++j, ++t) {
        mbar_wait(bar(PS_FULL + g), t & 1);
        tcgen05_fence_after();
        const bool ragged = j * BN + BN > p.seqlen;  // kv columns past the sequence end (last tile only): the slow, masked path
        const int n_valid = p.seqlen - j * BN;       // valid columns of this tile when ragged
        // ---- pass 1: row max over the 128 columns of S; the TMEM load of chunk c + 1 is in flight while chunk c is reduced
        float mx = -INFINITY;
        {
          uint32_t sa[32], sb[32];
          tmem_ld_32x32b_x32(s_addr, sa);
          tmem_ld_32x32b_x32(s_addr + 32, sb);
          tmem_ld_wait();
          auto red = [&](const uint32_t* sv, int c0) {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (!ragged) {
#pragma unroll
              for (int c = 0; c < 32; c += 2) m4[(c >> 1) & 3] = fmaxf(m4[(c >> 1) & 3], fmaxf(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1])));
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) m4[c & 3] = fmaxf(m4[c & 3], c0 + c < n_valid ? __uint_as_float(sv[c]) : -INFINITY);
            }
            mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          };
          red(sa, 0);
          tmem_ld_32x32b_x32(s_addr + 64, sa);
          red(sb, 32);
          tmem_ld_32x32b_x32(s_addr + 96, sb);
          tmem_ld_wait();
          red(sa, 64);
          red(sb, 96);
        }
        const float m_new = fmaxf(m_run, mx);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = ex2_approx((m_run - m_use) * sl);  // m_run = -inf -> 0
        m_run = m_new;
        const float msl = m_use * sl;
        // ---- pass 2: P = 2^(s * scale - m) -> bf16 -> 128B-swizzled K-major tile, row sum; loads pipelined as in pass 1
        mbar_wait(bar(PP_EMPTY + g), (t & 1) ^ 1);
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
        {
          uint32_t sa[32], sb[32];
          auto emit = [&](const uint32_t* sv, int c0) {
            float e[32];
            if (!ragged) {
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                e[c] = ex2_approx(fmaf(__uint_as_float(sv[c]), sl, -msl));
                sum4[c & 3] += e[c];
              }
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                e[c] = c0 + c < n_valid ? ex2_approx(fmaf(__uint_as_float(sv[c]), sl, -msl)) : 0.f;
                sum4[c & 3] += e[c];
              }
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const int col = c0 + cc * 8;
              const int kb = col >> 6, cin = (col & 63) >> 3;
              st_shared_v4(p_row + kb * (BM * 128) + ((cin ^ (r & 7)) << 4), pack8(e + cc * 8));
            }
          };
          tmem_ld_32x32b_x32(s_addr, sa);
          tmem_ld_32x32b_x32(s_addr + 32, sb);
          tmem_ld_wait();
          emit(sa, 0);
          tmem_ld_32x32b_x32(s_addr + 64, sa);
          emit(sb, 32);
          tmem_ld_32x32b_x32(s_addr + 96, sb);
          tmem_ld_wait();
          // S is in registers for the last time: the tensor pipe may overwrite it with the next tile's S
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(PS_EMPTY + g));
          emit(sa, 64);
          emit(sb, 96);
        }
        l_run = fmaf(l_run, alpha, (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(PP_FULL + g));
        // ---- the PREVIOUS tile's P V product: issued a whole tile ago, so it is complete by now (folding it between the two passes
        //      had the group spin on O_FULL for 12 % of the samples); P_g V(j) waits for this fold through O_EMPTY, ~200 clk
        if (j > 0) fold(t - 1, alpha_prev);
        alpha_prev = alpha;
      }
      // ---- last product of the item, normalisation, store (72 channels = nine 16-byte vectors per row)
      fold(t - 1, alpha_prev);
      const int qrow = k.q0 + g * BM + r;
      if (qrow < p.seqlen) {
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll
        for (int c = 0; c < 72; ++c) o[c] *= inv;
        bf16* orow = p.out + (size_t)(k.row_base + qrow) * p.o_ld + k.head * HD;
#pragma unroll
        for (int c = 0; c < 9; ++c) *reinterpret_cast<uint4*>(orow + c * 8) = pack8(o + c * 8);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int HD, int BN, bool CAUSAL>
static int launch(const void* q, const void* k, const void* v, void* out, int q_ld, int kv_ld, int o_ld, int batch, int seqlen, const int* cu,
                  long long total_rows, int n_heads, int n_kv_heads, float scale, cudaStream_t st) {
  using G = Geo<HD, BN>;
  static bool configured = false;
  if (!configured) {
    SRGPT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD, BN, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    configured = true;
  }
  Maps maps;
  struct View { const void* ptr; int heads, ld, rows; CUtensorMap *m0, *m1; };
  const View views[3] = {{q, n_heads, q_ld, BM, &maps.q0, &maps.q1}, {k, n_kv_heads, kv_ld, BN, &maps.k0, &maps.k1}, {v, n_kv_heads, kv_ld, BN, &maps.v0, &maps.v1}};
  for (const View& vw : views) {
    // (channel, head, row) view of a column slice of the fused qkv activation; channels >= HD are out of bounds -> 0
    const cuuint64_t dims[3] = {(cuuint64_t)HD, (cuuint64_t)vw.heads, (cuuint64_t)total_rows};
    const cuuint64_t strides[2] = {(cuuint64_t)HD * 2, (cuuint64_t)vw.ld * 2};
    const cuuint32_t box0[3] = {64, 1, (cuuint32_t)vw.rows};
    const cuuint32_t box1[3] = {(cuuint32_t)G::W1, 1, (cuuint32_t)vw.rows};
    int rc = encode_tmap_bf16(vw.m0, vw.ptr, 3, dims, strides, box0, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc == 0) rc = encode_tmap_bf16(vw.m1, vw.ptr, 3, dims, strides, box1, G::TAIL32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != 0) {
      set_last_error("attention (tcgen05): cuTensorMapEncodeTiled failed (%d) for ptr=%p heads=%d ld=%d rows=%lld", rc, vw.ptr, vw.heads, vw.ld, total_rows);
      return SRGPT_ERR_CUDA;
    }
  }
  Params p;
  p.cu_seqlens = cu;
  p.seqlen = seqlen;
  p.batch = batch;
  p.n_heads = n_heads;
  p.group = n_heads / n_kv_heads;
  p.nqt = ceil_div(seqlen, BM);
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(out);
  p.o_ld = o_ld;
  static const int s_ahead_env = getenv("SRGPT_ATTN_S_AHEAD") ? atoi(getenv("SRGPT_ATTN_S_AHEAD")) : 0;
  p.s_ahead = (s_ahead_env == 1 || s_ahead_env == 2) ? s_ahead_env : 1;
  if (HD == 72 && BN == 128 && !CAUSAL && cu == nullptr && n_heads == n_kv_heads) {
    // the SigLIP tower: ping-pong kernel (two q tiles per CTA); SRGPT_ATTN_PP=-1 selects the single-tile kernel (A/B knob)
    static const int pp_env = getenv("SRGPT_ATTN_PP") ? atoi(getenv("SRGPT_ATTN_PP")) : 0;
    if (pp_env >= 0) {
      static bool pp_configured = false;
      if (!pp_configured) {
        SRGPT_CHECK_CUDA(cudaFuncSetAttribute(attn_vit_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pp::SMEM_BYTES));
        pp_configured = true;
      }
      const long long n_items = (long long)((p.nqt + 1) / 2) * n_heads * batch;
      const int grid_pp = (int)(n_items < sm_count() ? n_items : sm_count());
      attn_vit_pp_kernel<<<grid_pp, NTHREADS, pp::SMEM_BYTES, st>>>(maps, p);
      SRGPT_CHECK_LAUNCH();
      return SRGPT_OK;
    }
  }
  const long long n_work = (long long)p.nqt * n_heads * batch;
  const int grid = (int)(n_work < sm_count() ? n_work : sm_count());
  attn_fwd_tc_kernel<HD, BN, CAUSAL><<<grid, NTHREADS, G::SMEM_BYTES, st>>>(maps, p);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
}

// Entry used by attention.cu's dispatcher.  Returns SRGPT_ERR_UNSUPPORTED when the shape is not covered (the caller then
// uses the mma.sync kernel).  seqlen = max sequence length when cu != null.
int prefill(const void* q, const void* k, const void* v, void* out, int q_ld, int kv_ld, int o_ld, int batch, int seqlen, const int* cu, long long total_rows,
            int n_heads, int n_kv_heads, int head_dim, float scale, int causal, cudaStream_t st) {
  if ((o_ld % 8) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0 || (q_ld % 8) != 0 || (kv_ld % 8) != 0) return SRGPT_ERR_UNSUPPORTED;
  if (head_dim == 72 && !causal)
    return launch<72, 128, false>(q, k, v, out, q_ld, kv_ld, o_ld, batch, seqlen, cu, total_rows, n_heads, n_kv_heads, scale, st);
  if (head_dim == 128 && causal)
    return launch<128, 64, true>(q, k, v, out, q_ld, kv_ld, o_ld, batch, seqlen, cu, total_rows, n_heads, n_kv_heads, scale, st);
  if (head_dim == 128 && !causal)
    return launch<128, 64, false>(q, k, v, out, q_ld, kv_ld, o_ld, batch, seqlen, cu, total_rows, n_heads, n_kv_heads, scale, st);
  return SRGPT_ERR_UNSUPPORTED;
}

}  // namespace attn_tc
}  // namespace srgpt

// Region-extractor HBM kernels: mask-weight resampling, mask-guided pooling, adaptive average
// pooling, row re-ordering and depth-map preparation.
//
// Reference semantics: llava/model/region_extractor/base_extractor.py:27-84 (MaskPooling),
// :123,145 (AdaptiveAvgPool2d(27)), llava/eval/eval_spatial.py:99-105 (depth map).
//
// Feature-row orders (`order` argument):
//   0 : row-major, row = y*side + x                       (tower features, reference `hres`)
//   2 : 2-level nested 2x2, side = 4P: row = ((y>>2)*P + (x>>2))*16 + (((y>>1)&1)*2 + ((x>>1)&1))*4
//       + ((y&1)*2 + (x&1)).  This is the order in which the two ConvTranspose2d(k=2,s=2) GEMMs emit
//       pixels, so the refinement never needs a pixel-shuffle pass over the 37.7 MB/image tensor.
#include "common.cuh"
#include "tcgen05.cuh"
#include "srgpt_b200.h"

namespace srgpt {

__device__ __forceinline__ int feat_row(int y, int x, int side, int order) {
  if (order == 0) return y * side + x;
  const int P = side >> 2;
  return ((((y >> 2) * P + (x >> 2)) << 4) | (((((y >> 1) & 1) << 1) | ((x >> 1) & 1)) << 2) | (((y & 1) << 1) | (x & 1)));
}

// ATen area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
__device__ __forceinline__ float src_index(float rscale, int dst) {
  // one fused multiply-add, as both ATen builds contract it (gcc -ffp-contract=fast on the host, nvcc -fmad on the device)
  const float s = __fmaf_rn(rscale, (float)dst + 0.5f, -0.5f);
  return s < 0.f ? 0.f : s;
}

template <typename T>
__device__ __forceinline__ float bilinear_tap(const T* __restrict__ img, int IH, int IW, float rs_y, float rs_x, int oy, int ox) {
  const float sy = src_index(rs_y, oy), sx = src_index(rs_x, ox);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < IH - 1 ? 1 : 0), x1 = x0 + (x0 < IW - 1 ? 1 : 0);
  const float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float v00 = (float)img[(size_t)y0 * IW + x0], v01 = (float)img[(size_t)y0 * IW + x1];
  const float v10 = (float)img[(size_t)y1 * IW + x0], v11 = (float)img[(size_t)y1 * IW + x1];
  // same association as ATen's upsample_bilinear2d kernels
  return __fadd_rn(__fmul_rn(ly0, __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01))),
                   __fmul_rn(ly1, __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11))));
}

// ---------------------------------------------------------------------------------------------
// mask weights, two fully parallel passes (a single CTA per mask was latency-bound: 37 us for 8 masks):
//   pass 1  grid (splits, M, n_img): resample -> bf16, per-split partial sums (fixed order -> deterministic)
//   pass 2  same grid: denorm = bf16(bf16(sum) + 1e-8) (base_extractor.py:61), w = bf16(v / denorm)
// ---------------------------------------------------------------------------------------------
constexpr int MW_SPLITS = 16;
// rows of the pooling-weight matrices [n_img, M, L] are padded to a multiple of 8 elements (16 bytes): the TMA view of mask_pool_kernel
// needs 16-byte row strides and L = side^2 is odd for odd sides (27 x 27 tower tokens of a 384-px SigLIP)
__host__ __device__ __forceinline__ int mask_row_ld(int L) { return (L + 7) & ~7; }

// MaskPooling.forward is a chain of four small kernels (taps -> normalise -> pool -> reduce); at one image each lasts 1-8 us, so
// the launch gaps between them were a third of the whole op.  Every kernel after the first is a PROGRAMMATIC DEPENDENT launch: it
// starts while its producer drains, does the work that does not depend on it (index math, L2 prefetch of the feature rows) and
// blocks in griddepcontrol.wait until the producer's writes are visible.
__device__ __forceinline__ void chain_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename T>
__global__ void __launch_bounds__(256)
mask_taps_kernel(const T* __restrict__ masks, bf16* __restrict__ w, float* __restrict__ psum, int M, int IH, int IW, int side,
                 float rscale, int order) {
  __shared__ float red[32];
  chain_launch_dependents();
  const int split = blockIdx.x, m = blockIdx.y, img = blockIdx.z;
  const T* src = masks + ((size_t)img * M + m) * IH * IW;
  const int L = side * side;
  bf16* dst = w + ((size_t)img * M + m) * mask_row_ld(L);
  const int per = (L + MW_SPLITS - 1) / MW_SPLITS;
  const int l_end = min(L, (split + 1) * per);
  float sum = 0.f;
  for (int l = split * per + threadIdx.x; l < l_end; l += blockDim.x) {
    const int oy = l / side, ox = l - oy * side;
    const bf16 v = f2e(bilinear_tap(src, IH, IW, rscale, rscale, oy, ox));
    dst[feat_row(oy, ox, side, order)] = v;
    sum += e2f(v);
  }
  const float total = block_sum(sum, red);
  if (threadIdx.x == 0) psum[((size_t)img * M + m) * MW_SPLITS + split] = total;
}

__global__ void __launch_bounds__(256)
mask_normalise_kernel(const bf16* __restrict__ v, bf16* __restrict__ w, const float* __restrict__ psum, int M, int L) {
  const int split = blockIdx.x, m = blockIdx.y, img = blockIdx.z;
  chain_launch_dependents();
  chain_wait();  // taps + partial sums of the producer are visible
  const float* ps = psum + ((size_t)img * M + m) * MW_SPLITS;
  float total = 0.f;
#pragma unroll
  for (int i = 0; i < MW_SPLITS; ++i) total += ps[i];
  const float denorm = bf16_round(bf16_round(total) + 1e-8f);
  const bf16* src = v + ((size_t)img * M + m) * mask_row_ld(L);
  bf16* dst = w + ((size_t)img * M + m) * mask_row_ld(L);
  const int per = (L + MW_SPLITS - 1) / MW_SPLITS;
  const int l_end = min(L, (split + 1) * per);
  // rows are a permutation of l; normalising the contiguous range [split*per, l_end) of ROWS covers every row once
  for (int r = split * per + threadIdx.x; r < l_end; r += blockDim.x)
    dst[r] = f2e(__fdiv_rn(e2f(src[r]), denorm));
}

// ---------------------------------------------------------------------------------------------
// mask pooling: out[img,m,c] = sum_l w[img,m,l] * x[img,l,c]   (torch.einsum("lc,ml->mc"), base_extractor.py:74-78)
//
// The reference runs this einsum as a bf16 tensor-core GEMM with fp32 accumulation; so do we, but shaped for what
// it is — a [<=16 x L] x [L x C] product whose only cost is streaming x once:
//   * CTA = 128 channels x rows_per_cta rows; features go global -> shared with cp.async (16-byte chunks, 4-stage
//     ring, ~50 KB in flight per CTA, no registers involved), 64 rows per stage;
//   * each of the 8 warps owns 16 channels: per 16 rows one ldmatrix.x4 of the mask weights (A, [m][l] row-major),
//     one ldmatrix.x4.trans of the features (B) and two mma.sync.m16n8k16 (bf16 x bf16 -> fp32);
//   * up to 16 masks per pass; fp32 partials [img][R][M][C] + a tiny deterministic reduce.
// History (profiles/): scalar FFMA 0.11-0.20 of HBM peak, packed FFMA2 + row skipping 0.23-0.49, this version: see
// DESIGN.md.  ~5 instructions per 512 bytes of features per warp instead of ~90.
// ---------------------------------------------------------------------------------------------
constexpr int MP_THREADS = 256;
constexpr int MP_CH = 128;        // channels per CTA
constexpr int MP_MT = 16;         // masks per pass (MMA M)
constexpr int MP_SROWS = 64;      // feature rows per pipeline stage
constexpr int MP_NST = 5;         // pipeline stages
constexpr int MP_MAX_ROWS = 1024; // rows per CTA upper bound (planning only)
constexpr int MP_X_BOX = MP_SROWS * 128;                 // one TMA box: 64 rows x 64 channels, 128-byte rows, SWIZZLE_128B (8 KB)
constexpr int MP_STAGE_X = 2 * MP_X_BOX;                 // channels [0, 64) and [64, 128)
constexpr int MP_STAGE_W = MP_MT * 128;                  // 16 masks x 64 rows of weights, 128-byte rows, SWIZZLE_128B (2 KB)
constexpr int MP_STAGE_BYTES = MP_STAGE_X + MP_STAGE_W;  // 18 KB
constexpr int MP_SMEM_BYTES = MP_NST * MP_STAGE_BYTES + 1024 /*alignment*/ + 64 /*barriers*/;

__device__ __forceinline__ uint32_t mp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mp_tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(x), "r"(y)
               : "memory");
}

// Round 2: the feature / weight tiles arrive by TMA (cp.async.bulk.tensor, one elected thread, SASS UTMALDG) into 128B-swizzled
// tiles that ldmatrix reads conflict-free, 5 stages of 18 KB per CTA, completion on mbarriers.  Round 1 issued 1024 16-byte
// cp.async per stage from all threads and kept 3 stages in flight: ncu showed the kernel waiting on memory (long scoreboard) at 31 %
// of the DRAM throughput (profiles/r02_ncu_full_mask_pool_summary.txt).  Rows past the end of an image belong to the NEXT image in the
// flattened [(n_img L), C] view; their weights are out of bounds in the [(n_img M), L] view of w and arrive as zeros, so they add 0.
__global__ void __launch_bounds__(MP_THREADS, 2)
mask_pool_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, float* __restrict__ partial, int M, int L, int C,
                 int rows_per_cta, int R) {
  extern __shared__ uint8_t mp_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mp_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + MP_NST * MP_STAGE_BYTES);
  const int img = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cbase = blockIdx.y * MP_CH;
  const int l0 = blockIdx.x * rows_per_cta;
  const int nrows = min(rows_per_cta, L - l0);
  const int nstages = (nrows + MP_SROWS - 1) / MP_SROWS;
  const int lr = lane & 7, lmat = lane >> 3, g = lane >> 2, t4 = lane & 3;
  const uint32_t sbase = mp_smem_u32(smem);

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmap_x);
    tc::prefetch_tmap(&tmap_w);
    for (int i = 0; i < MP_NST; ++i) tc::mbar_init(mp_smem_u32(&full_bar[i]), 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  chain_launch_dependents();
  chain_wait();  // the normalised weights of the producer kernel are visible (the feature rows were written long before)

  const int n_pass = (M + MP_MT - 1) / MP_MT;
  const int total = n_pass * nstages;  // global stage index gs = pass * nstages + st; slot = gs % NST, parity = (gs / NST) & 1
  auto issue = [&](int gs) {           // thread 0 only
    const int pass = gs / nstages, st = gs - pass * nstages;
    const uint32_t slot = sbase + (gs % MP_NST) * MP_STAGE_BYTES;
    const uint32_t bar = mp_smem_u32(&full_bar[gs % MP_NST]);
    tc::mbar_expect_tx(bar, MP_STAGE_BYTES);  // out-of-bounds parts are zero-filled and still counted
    const int row = img * L + l0 + st * MP_SROWS;
    mp_tma_load_2d(slot, &tmap_x, bar, cbase, row);
    mp_tma_load_2d(slot + MP_X_BOX, &tmap_x, bar, cbase + 64, row);
    mp_tma_load_2d(slot + MP_STAGE_X, &tmap_w, bar, l0 + st * MP_SROWS, img * M + pass * MP_MT);
  };
  if (threadIdx.x == 0)
    for (int gs = 0; gs < MP_NST && gs < total; ++gs) issue(gs);

  for (int pass = 0; pass < n_pass; ++pass) {
    const int m0 = pass * MP_MT;
    const int mt = min(MP_MT, M - m0);
    float acc[2][4];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
    for (int st = 0; st < nstages; ++st) {
      const int gs = pass * nstages + st;
      tc::mbar_wait(mp_smem_u32(&full_bar[gs % MP_NST]), (uint32_t)(gs / MP_NST) & 1u);
      const uint32_t tx = sbase + (gs % MP_NST) * MP_STAGE_BYTES;
      const uint32_t tw = tx + MP_STAGE_X;
#pragma unroll
      for (int kk = 0; kk < MP_SROWS / 16; ++kk) {
        uint32_t a[4], b[4];
        // A (weights [m][l], 128-byte rows, 16-byte chunk c of row m at ((c ^ (m & 7)) << 4)):
        //   matrices (m0..7,k0..7) (m8..15,k0..7) (m0..7,k8..15) (m8..15,k8..15)
        const int am = lr + (lmat & 1) * 8, ac = kk * 2 + (lmat >> 1);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(tw + am * 128 + ((ac ^ (am & 7)) << 4)));
        // B (features [l][c], transposed on load): (k0..7,c0..7) (k8..15,c0..7) (k0..7,c8..15) (k8..15,c8..15)
        const int bl = kk * 16 + lr + (lmat & 1) * 8, c0 = warp * 16 + (lmat >> 1) * 8;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
                     : "r"(tx + (c0 >> 6) * MP_X_BOX + bl * 128 + ((((c0 & 63) >> 3) ^ (bl & 7)) << 4)));
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." SRGPT_ELEM_PTX "." SRGPT_ELEM_PTX ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(acc[nb][0]), "+f"(acc[nb][1]), "+f"(acc[nb][2]), "+f"(acc[nb][3])
                       : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[2 * nb]), "r"(b[2 * nb + 1]));
      }
      __syncthreads();  // every warp has read the slot: refill it with the stage NST ahead
      if (threadIdx.x == 0 && gs + MP_NST < total) issue(gs + MP_NST);
    }
    // accumulator fragment: c0,c1 -> (m = g, ch = 2*t4, +1); c2,c3 -> (m = g + 8, ...)
    float* pb = partial + (((size_t)img * R + blockIdx.x) * M + m0) * C;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      const int c = cbase + warp * 16 + nb * 8 + 2 * t4;
      if (c < C) {
        if (g < mt) { pb[(size_t)g * C + c] = acc[nb][0]; pb[(size_t)g * C + c + 1] = acc[nb][1]; }
        if (g + 8 < mt) { pb[(size_t)(g + 8) * C + c] = acc[nb][2]; pb[(size_t)(g + 8) * C + c + 1] = acc[nb][3]; }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
mask_pool_reduce_kernel(const float* __restrict__ partial, bf16* __restrict__ out, int M, int C, int R) {
  const int m = blockIdx.y, img = blockIdx.z;
  const int c = blockIdx.x * 256 + threadIdx.x;
  chain_wait();  // (no dependents of its own: the next kernel in the stream is an ordinary launch)
  if (c >= C) return;
  const float* p = partial + ((size_t)img * R * M + m) * C + c;
  const size_t stride = (size_t)M * C;
  float s = 0.f;
  int r = 0;
  for (; r + 8 <= R; r += 8) {  // 8 independent loads in flight, summed in index order (deterministic)
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = p[(size_t)(r + k) * stride];
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
  }
  for (; r < R; ++r) s += p[(size_t)r * stride];
  out[((size_t)img * M + m) * C + c] = f2e(s);
}

static void mask_pool_plan(int n_img, int L, int C, int* R, int* rows_per_cta, int* Q) {
  *Q = ceil_div(C, MP_CH);
  int want = ceil_div(2 * sm_count(), (*Q) * n_img);  // ~2 CTAs per SM
  if (want < 1) want = 1;
  int rpc = ceil_div(ceil_div(L, want), MP_SROWS) * MP_SROWS;  // whole pipeline stages
  if (rpc > MP_MAX_ROWS) rpc = MP_MAX_ROWS;
  if (rpc < MP_SROWS) rpc = MP_SROWS;
  *rows_per_cta = rpc;
  *R = ceil_div(L, rpc);
}

// ---------------------------------------------------------------------------------------------
// adaptive average pool: one CTA per output pixel, threads along channels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160)
adaptive_avgpool_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int side, int out_side, int C, int order) {
  const int img = blockIdx.y;
  const int oy = blockIdx.x / out_side, ox = blockIdx.x % out_side;
  // ATen start_index / end_index: floor(o*in/out), ceil((o+1)*in/out)
  const int ys = (oy * side) / out_side, ye = ((oy + 1) * side + out_side - 1) / out_side;
  const int xs = (ox * side) / out_side, xe = ((ox + 1) * side + out_side - 1) / out_side;
  const float inv = 1.0f / (float)((ye - ys) * (xe - xs));
  const bf16* xb = x + (size_t)img * side * side * C;
  for (int c = threadIdx.x; c < (C >> 3); c += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int yy = ys; yy < ye; ++yy)
      for (int xx = xs; xx < xe; ++xx) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(xb + (size_t)feat_row(yy, xx, side, order) * C + (c << 3)), f);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] += f[t];
      }
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] *= inv;
    *reinterpret_cast<uint4*>(y + ((size_t)img * out_side * out_side + blockIdx.x) * C + (c << 3)) = pack8(acc);
  }
}

__global__ void reorder_rows_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int side, int C, int from_order, int to_order) {
  const int img = blockIdx.y;
  const int yy = blockIdx.x / side, xx = blockIdx.x % side;
  const uint4* src = reinterpret_cast<const uint4*>(x + ((size_t)img * side * side + feat_row(yy, xx, side, from_order)) * C);
  uint4* dst = reinterpret_cast<uint4*>(y + ((size_t)img * side * side + feat_row(yy, xx, side, to_order)) * C);
  for (int c = threadIdx.x; c < (C >> 3); c += blockDim.x) dst[c] = src[c];
}

// ---------------------------------------------------------------------------------------------
// depth map: resize + min/max (pass 1), normalise to u8 x3 (pass 2)
// ---------------------------------------------------------------------------------------------
// Bilinear tap with the EXACT fp32 operation order of ATen's upsample_bilinear2d as compiled (the reference runs it in
// get_depth_map, llava/eval/eval_spatial.py:100): val = fma(w_y0, fma(w_x0, v00, w_x1 * v01), w_y1 * fma(w_x0, v10, w_x1 * v11)).
// The uint8 truncation right after makes a 1-ulp difference visible, so the association matters here (it does not for the
// mask weights, which are rounded to bf16); pinned bit-exactly against torch in tests/test_gpu_ops.py::test_depth_to_u8x3.
__device__ __forceinline__ float bilinear_tap_aten(const float* __restrict__ img, int IH, int IW, float rs_y, float rs_x, int oy, int ox) {
  const float sy = src_index(rs_y, oy), sx = src_index(rs_x, ox);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < IH - 1 ? 1 : 0), x1 = x0 + (x0 < IW - 1 ? 1 : 0);
  const float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float v00 = img[(size_t)y0 * IW + x0], v01 = img[(size_t)y0 * IW + x1];
  const float v10 = img[(size_t)y1 * IW + x0], v11 = img[(size_t)y1 * IW + x1];
  const float r0 = __fmaf_rn(lx0, v00, __fmul_rn(lx1, v01));
  const float r1 = __fmaf_rn(lx0, v10, __fmul_rn(lx1, v11));
  return __fmaf_rn(ly0, r0, __fmul_rn(ly1, r1));
}

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void depth_init_kernel(int* mm) {
  mm[0] = float_to_ordered(INFINITY);
  mm[1] = float_to_ordered(-INFINITY);
}
__global__ void __launch_bounds__(256)
depth_resize_kernel(const float* __restrict__ d, int h, int w, float* __restrict__ tmp, int H, int W, float rs_y, float rs_x, int* mm) {
  __shared__ float smin[8], smax[8];
  float lo = INFINITY, hi = -INFINITY;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
    const float v = bilinear_tap_aten(d, h, w, rs_y, rs_x, p / W, p % W);
    tmp[p] = v;
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { lo = fminf(lo, smin[i]); hi = fmaxf(hi, smax[i]); }
    atomicMin(&mm[0], float_to_ordered(lo));
    atomicMax(&mm[1], float_to_ordered(hi));
  }
}
__global__ void depth_norm_kernel(const float* __restrict__ tmp, const int* __restrict__ mm, unsigned char* __restrict__ out, int n) {
  const float lo = ordered_to_float(mm[0]), hi = ordered_to_float(mm[1]);
  const float range = __fsub_rn(hi, lo);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const float t = __fmul_rn(__fdiv_rn(__fsub_rn(tmp[p], lo), range), 255.0f);
    const unsigned char u = (unsigned char)t;  // numpy astype(uint8): truncation
    out[3 * p] = u;
    out[3 * p + 1] = u;
    out[3 * p + 2] = u;
  }
}

}  // namespace srgpt

using namespace srgpt;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// programmatic dependent launch of `kernel` behind the previous kernel of the stream (plain launch when SRGPT_NO_PDL=1)
template <typename... KArgs, typename... Args>
static cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, int block, int smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

extern "C" __attribute__((visibility("default"))) long long srgpt_mask_weights_workspace(int n_img, int M, int side) {
  if (n_img <= 0 || M <= 0 || side <= 0) return -1;
  return (long long)n_img * M * (MW_SPLITS * (long long)sizeof(float) + (long long)mask_row_ld(side * side) * (long long)sizeof(bf16));
}

extern "C" __attribute__((visibility("default"))) int srgpt_mask_weights(const void* masks, int mask_is_bf16, void* w, void* workspace, int n_img, int M, int IH, int IW,
                                  int side, float rscale, int order, void* stream) {
  SRGPT_CHECK_ARG(masks && w && workspace && n_img > 0 && M > 0 && IH > 0 && IW > 0 && side > 0);
  SRGPT_CHECK_ARG(order == 0 || (order == 2 && (side % 4) == 0));
  dim3 grid(MW_SPLITS, M, n_img);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int L = side * side;
  // workspace: [n_img*M*MW_SPLITS] fp32 partial sums, then [n_img*M*ld(L)] bf16 un-normalised resampled masks (rows padded like w)
  float* psum = reinterpret_cast<float*>(workspace);
  bf16* v = reinterpret_cast<bf16*>(psum + (size_t)n_img * M * MW_SPLITS);
  if (mask_is_bf16)
    mask_taps_kernel<bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(masks), v, psum, M, IH, IW, side, rscale, order);
  else
    mask_taps_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(masks), v, psum, M, IH, IW, side, rscale, order);
  SRGPT_CHECK_LAUNCH();
  SRGPT_CHECK_CUDA(launch_dependent(mask_normalise_kernel, grid, 256, 0, st, (const bf16*)v, reinterpret_cast<bf16*>(w), (const float*)psum, M, L));
  return SRGPT_OK;
}

extern "C" __attribute__((visibility("default"))) long long srgpt_mask_pool_workspace(int n_img, int M, int L, int C) {
  if (n_img <= 0 || M <= 0 || L <= 0 || C <= 0) return -1;
  int R, rpc, Q;
  mask_pool_plan(n_img, L, C, &R, &rpc, &Q);
  return (long long)n_img * R * M * C * (long long)sizeof(float);
}

extern "C" __attribute__((visibility("default"))) int srgpt_mask_pool_bf16(const void* x, const void* w, void* out, void* workspace, int n_img, int M, int L, int C,
                                    void* stream) {
  SRGPT_CHECK_ARG(x && w && out && workspace && n_img > 0 && M > 0 && L > 0 && C > 0);
  SRGPT_CHECK_ARG((C % 8) == 0 && aligned16(x) && aligned16(w));
  int R, rpc, Q;
  mask_pool_plan(n_img, L, C, &R, &rpc, &Q);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static bool configured = false;
  if (!configured) {
    SRGPT_CHECK_CUDA(cudaFuncSetAttribute(mask_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MP_SMEM_BYTES));
    configured = true;
  }
  dim3 grid(R, Q, n_img);
  // tensor maps: features as [(n_img L), C] with boxes of 64 rows x 64 channels, weights as [(n_img M), L] with boxes of 16 masks x 64 rows
  CUtensorMap tm_x, tm_w;
  {
    const cuuint64_t dx[2] = {(cuuint64_t)C, (cuuint64_t)n_img * L}, sx[1] = {(cuuint64_t)C * 2};
    const cuuint32_t bx[2] = {64, (cuuint32_t)MP_SROWS};
    const cuuint64_t dw[2] = {(cuuint64_t)L, (cuuint64_t)n_img * M}, sw[1] = {(cuuint64_t)mask_row_ld(L) * 2};
    const cuuint32_t bw[2] = {(cuuint32_t)MP_SROWS, (cuuint32_t)MP_MT};
    if (tc::encode_tmap_bf16(&tm_x, x, 2, dx, sx, bx, CU_TENSOR_MAP_SWIZZLE_128B) != 0 ||
        tc::encode_tmap_bf16(&tm_w, w, 2, dw, sw, bw, CU_TENSOR_MAP_SWIZZLE_128B) != 0) {
      set_last_error("srgpt_mask_pool_bf16: cuTensorMapEncodeTiled failed (x=%p w=%p L=%d C=%d)", x, w, L, C);
      return SRGPT_ERR_CUDA;
    }
  }
  // behind mask_normalise_kernel (or whatever precedes it in the stream: every kernel of the chain waits before it reads)
  SRGPT_CHECK_CUDA(launch_dependent(mask_pool_kernel, grid, MP_THREADS, MP_SMEM_BYTES, st, tm_x, tm_w, reinterpret_cast<float*>(workspace), M, L, C, rpc, R));
  SRGPT_CHECK_CUDA(launch_dependent(mask_pool_reduce_kernel, dim3(ceil_div(C, 256), M, n_img), 256, 0, st, reinterpret_cast<const float*>(workspace),
                                    reinterpret_cast<bf16*>(out), M, C, R));
  return SRGPT_OK;
}

extern "C" __attribute__((visibility("default"))) int srgpt_adaptive_avgpool_bf16(const void* x, void* y, int n_img, int side, int out_side, int C, int order,
                                           void* stream) {
  SRGPT_CHECK_ARG(x && y && n_img > 0 && side > 0 && out_side > 0 && C > 0 && (C % 8) == 0);
  SRGPT_CHECK_ARG(order == 0 || (order == 2 && (side % 4) == 0));
  SRGPT_CHECK_ARG(aligned16(x) && aligned16(y));
  adaptive_avgpool_kernel<<<dim3(out_side * out_side, n_img), 160, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), side, out_side, C, order);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
}

extern "C" __attribute__((visibility("default"))) int srgpt_reorder_rows_bf16(const void* x, void* y, int n_img, int side, int C, int from_order, int to_order,
                                       void* stream) {
  SRGPT_CHECK_ARG(x && y && x != y && n_img > 0 && side > 0 && C > 0 && (C % 8) == 0);
  SRGPT_CHECK_ARG((from_order == 0 || from_order == 2) && (to_order == 0 || to_order == 2));
  SRGPT_CHECK_ARG(((from_order | to_order) & 2) == 0 || (side % 4) == 0);
  SRGPT_CHECK_ARG(aligned16(x) && aligned16(y));
  reorder_rows_kernel<<<dim3(side * side, n_img), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), side, C, from_order, to_order);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
}

extern "C" __attribute__((visibility("default"))) int srgpt_depth_to_u8x3(const void* depth, int h, int w, void* out, int H, int W, void* workspace, void* stream) {
  SRGPT_CHECK_ARG(depth && out && workspace && h > 0 && w > 0 && H > 0 && W > 0);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* tmp = reinterpret_cast<float*>(workspace);
  int* mm = reinterpret_cast<int*>(tmp + (size_t)H * W);
  // ATen area_pixel_compute_scale with size= given: (float)in / out
  const float rs_y = (float)h / (float)H, rs_x = (float)w / (float)W;
  depth_init_kernel<<<1, 1, 0, st>>>(mm);
  SRGPT_CHECK_LAUNCH();
  int blocks = ceil_div(H * W, 256);
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  depth_resize_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(depth), h, w, tmp, H, W, rs_y, rs_x, mm);
  SRGPT_CHECK_LAUNCH();
  depth_norm_kernel<<<blocks, 256, 0, st>>>(tmp, mm, reinterpret_cast<unsigned char*>(out), H * W);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
}

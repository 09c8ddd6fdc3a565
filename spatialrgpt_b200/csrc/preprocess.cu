// Host preprocessing on the GPU: the image / region preparation in front of generate() - llava/mm_utils.py:421-542
// (process_image / process_images / process_regions) - whose arithmetic lives in third-party code the reference pins:
//   * the HF SiglipImageProcessor of transformers 4.37.2 = PIL `Image.resize(..., BICUBIC)` on uint8, then `image * (1/255)` in
//     float64 -> float32, then `(x - mean) / std` in float32, channels first;
//   * region masks: `cv2.resize(m, (R, R), INTER_NEAREST)` (mm_utils.py:521-523), then the same processor without rescale /
//     normalisation (a same-size resize is a copy) -> float.
// Pillow's resampler (libImaging/Resample.c) is integer arithmetic once its coefficient tables exist: per output pixel a window
// [xmin, xmin + n) of the source row, n <= ksize coefficients in 22-bit fixed point, acc = 2^21 + sum(pixel * k), clip8(acc >> 22);
// the horizontal pass runs first, then the vertical one on its 8-bit result.  The tables depend only on (in_size, out_size) and are
// built on the host in the same double arithmetic (spatialrgpt_b200/preprocess.py), so the kernels below are bit-exact by
// construction: byte work, HBM-bound, one thread per output byte with the window in registers.
#include "common.cuh"
#include "srgpt_b200.h"

namespace srgpt {
namespace prep {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= PRECISION_BITS;
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// in [H, W, C] u8 -> out [H, Wout, C]: out[y, xx, c] = clip8(2^21 + sum_k in[y, xmin[xx] + k, c] * kk[xx, k])
__global__ void __launch_bounds__(256)
resample_h_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int H, int W, int C, int Wout, const int* __restrict__ kk,
                  const int* __restrict__ bounds, int ksize) {
  const long long n = (long long)H * Wout * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int xx = (int)((i / C) % Wout);
    const int y = (int)(i / ((long long)C * Wout));
    const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    const unsigned char* src = in + ((size_t)y * W + xmin) * C + c;
    const int* k = kk + (size_t)xx * ksize;
    int acc = 1 << (PRECISION_BITS - 1);
    for (int x = 0; x < cnt; ++x) acc += (int)src[(size_t)x * C] * k[x];
    out[i] = clip8(acc);
  }
}

// in [H, W, C] u8 -> out [Hout, W, C]
__global__ void __launch_bounds__(256)
resample_v_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int H, int W, int C, int Hout, const int* __restrict__ kk,
                  const int* __restrict__ bounds, int ksize) {
  const long long row = (long long)W * C;
  const long long n = (long long)Hout * row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int yy = (int)(i / row);
    const long long off = i - (long long)yy * row;
    const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    const unsigned char* src = in + (size_t)ymin * row + off;
    const int* k = kk + (size_t)yy * ksize;
    int acc = 1 << (PRECISION_BITS - 1);
    for (int y = 0; y < cnt; ++y) acc += (int)src[(size_t)y * row] * k[y];
    out[i] = clip8(acc);
  }
}

// [H, W, 3] u8 -> [3, H, W] float32: float((double)u8 * scale), then (x - mean[c]) / std[c] in float32 (HF rescale + normalize)
__global__ void __launch_bounds__(256)
normalize_chw_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, int H, int W, int C, double scale, float m0, float m1, float m2, float s0,
                     float s1, float s2, int do_normalize) {
  const long long n = (long long)H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / ((long long)H * W));
    const long long p = i - (long long)c * H * W;
    float v = (float)((double)in[p * C + c] * scale);
    if (do_normalize) {
      const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), s = c == 0 ? s0 : (c == 1 ? s1 : s2);
      v = __fdiv_rn(__fsub_rn(v, m), s);
    }
    out[i] = v;
  }
}

// nearest-neighbour resize with precomputed source indices (cv2.resize INTER_NEAREST: floor(dst * (1 / (dst_size / src_size))))
__global__ void __launch_bounds__(256)
nearest_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, int W, int Hout, int Wout, const int* __restrict__ ys, const int* __restrict__ xs) {
  const long long n = (long long)Hout * Wout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / Wout), x = (int)(i - (long long)y * Wout);
    out[i] = (float)in[(size_t)ys[y] * W + xs[x]];
  }
}

static int grid_for(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = 8LL * sm_count();
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace prep
}  // namespace srgpt

using namespace srgpt;

// One pass of Pillow's 8-bit resampler.  axis 0: vertical ([H, W, C] -> [out_size, W, C]); axis 1: horizontal (-> [H, out_size, C]).
// kk int32 [out_size, ksize] (22-bit fixed point), bounds int32 [out_size, 2] = (first source index, count): device arrays built by
// spatialrgpt_b200.preprocess.resample_coeffs.
extern "C" __attribute__((visibility("default"))) int srgpt_resample_u8(const void* in, void* out, int H, int W, int C, int axis, int out_size, const int* kk,
                                                                       const int* bounds, int ksize, void* stream) {
  SRGPT_CHECK_ARG(in && out && kk && bounds && H > 0 && W > 0 && C > 0 && out_size > 0 && ksize > 0 && (axis == 0 || axis == 1));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned char* ip = reinterpret_cast<const unsigned char*>(in);
  unsigned char* op = reinterpret_cast<unsigned char*>(out);
  if (axis == 1)
    prep::resample_h_kernel<<<prep::grid_for((long long)H * out_size * C), 256, 0, st>>>(ip, op, H, W, C, out_size, kk, bounds, ksize);
  else
    prep::resample_v_kernel<<<prep::grid_for((long long)out_size * W * C), 256, 0, st>>>(ip, op, H, W, C, out_size, kk, bounds, ksize);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
}

// [H, W, C<=3] u8 -> [C, H, W] fp32: x = float(double(u8) * scale); if (do_normalize) x = (x - mean[c]) / std[c].
extern "C" __attribute__((visibility("default"))) int srgpt_u8_to_normalized_chw(const void* in, float* out, int H, int W, int C, double scale, const float* mean3,
                                                                                const float* std3, int do_normalize, void* stream) {
  SRGPT_CHECK_ARG(in && out && H > 0 && W > 0 && C >= 1 && C <= 3 && (!do_normalize || (mean3 && std3)));
  float m[3] = {0, 0, 0}, s[3] = {1, 1, 1};
  if (do_normalize)
    for (int c = 0; c < C; ++c) { m[c] = mean3[c]; s[c] = std3[c]; }
  prep::normalize_chw_kernel<<<prep::grid_for((long long)H * W * C), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const unsigned char*>(in), out, H, W, C, scale, m[0], m[1], m[2], s[0], s[1], s[2], do_normalize);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
}

// [H, W] u8 -> [Hout, Wout] fp32 by index gather (ys [Hout], xs [Wout] device int32)
extern "C" __attribute__((visibility("default"))) int srgpt_resize_nearest_u8(const void* in, float* out, int H, int W, int Hout, int Wout, const int* ys, const int* xs,
                                                                             void* stream) {
  SRGPT_CHECK_ARG(in && out && ys && xs && H > 0 && W > 0 && Hout > 0 && Wout > 0);
  prep::nearest_kernel<<<prep::grid_for((long long)Hout * Wout), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const unsigned char*>(in), out, W,
                                                                                                                Hout, Wout, ys, xs);
  SRGPT_CHECK_LAUNCH();
  return SRGPT_OK;
}

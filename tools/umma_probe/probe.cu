// UMMA / TMA layout probe (development tool, not part of the product library): a single-CTA "dumb executor" that
// runs a host-described list of TMA loads and tcgen05.mma instructions and dumps the TMEM accumulator, so that
// shared-memory descriptor hypotheses (MN-major operands, 32B-swizzled tails, OOB zero fill) can be checked against
// torch.matmul in ONE GPU call.  Build: tools/umma_probe/build.sh; run: python tools/umma_probe/run_probe.py
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

struct Load { int map, ndim, c0, c1, c2; uint32_t smem_off, bytes; };
struct Mma { uint64_t a_desc, b_desc; uint32_t a_off, b_off, d_col, idesc, acc; };
struct Plan {
  int n_loads, n_mma, n_cols, manual_a;   // manual_a: threads copy A [128 x 64] bf16 from `a_src` into smem at a_manual_off (SW128 formula)
  uint32_t a_manual_off;
  const uint16_t* a_src; int a_ld;
  Load loads[8];
  Mma mma[40];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 24)) __trap();
  }
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2,
             const __grid_constant__ CUtensorMap m3, const Plan* __restrict__ planp, float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  __shared__ Plan plan;
  for (int i = threadIdx.x; i < (int)(sizeof(Plan) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(&plan)[i] = reinterpret_cast<const uint32_t*>(planp)[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  const uint32_t sbase = smem_u32(smem);

  if (plan.manual_a) {  // row r, 16-byte chunk c of a [128 x 64] bf16 tile -> SW128 K-major position
    const int r = threadIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(plan.a_src + (size_t)r * plan.a_ld);
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<uint4*>(smem + plan.a_manual_off + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = src[c];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async (tensor core) proxy
  }
  __syncthreads();

  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int i = 0; i < plan.n_loads; ++i) total += plan.loads[i].bytes;
    const uint32_t fb = smem_u32(&bars[0]);
    if (total) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(total) : "memory");
      for (int i = 0; i < plan.n_loads; ++i) {
        const Load& l = plan.loads[i];
        const CUtensorMap* tm = l.map == 0 ? &m0 : (l.map == 1 ? &m1 : (l.map == 2 ? &m2 : &m3));
        if (l.ndim == 2)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(sbase + l.smem_off), "l"(reinterpret_cast<uint64_t>(tm)), "r"(fb), "r"(l.c0), "r"(l.c1) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(sbase + l.smem_off), "l"(reinterpret_cast<uint64_t>(tm)), "r"(fb), "r"(l.c0), "r"(l.c1), "r"(l.c2) : "memory");
      }
      mbar_wait(fb, 0);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int i = 0; i < plan.n_mma; ++i) {
      const Mma& m = plan.mma[i];
      const uint64_t ad = m.a_desc | (uint64_t)(((sbase + m.a_off) & 0x3FFFF) >> 4);
      const uint64_t bd = m.b_desc | (uint64_t)(((sbase + m.b_off) & 0x3FFFF) >> 4);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem_base + m.d_col), "l"(ad), "l"(bd), "r"(m.idesc), "r"(m.acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
  }
  mbar_wait(smem_u32(&bars[1]), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + lane;
  for (int c = 0; c < plan.n_cols; c += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem_base + c + ((uint32_t)(warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(size_t)row * plan.n_cols + c + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// dims/box in elements (innermost first), strides in BYTES for dims 1..ndim-1; swizzle: 0 none, 1 32B, 2 64B, 3 128B
extern "C" int probe_make_map(void* out128, void* ptr, int ndim, const long long* dims, const long long* strides_bytes, const int* box, int swizzle) {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return -100;
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < ndim; ++i) { d[i] = (cuuint64_t)dims[i]; b[i] = (cuuint32_t)box[i]; e[i] = 1; if (i) s[i - 1] = (cuuint64_t)strides_bytes[i - 1]; }
  const CUtensorMapSwizzle sw = swizzle == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = reinterpret_cast<EncodeTiledFn>(sym)(reinterpret_cast<CUtensorMap*>(out128), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)ndim, ptr, d, s, b, e,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

extern "C" int probe_launch(const void* maps4x128, const Plan* plan_dev, float* D) {
  static bool cfg = false;
  const int smem = 160 * 1024;
  if (!cfg) { if (cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1; cfg = true; }
  const CUtensorMap* m = reinterpret_cast<const CUtensorMap*>(maps4x128);
  probe_kernel<<<1, 128, smem>>>(m[0], m[1], m[2], m[3], plan_dev, D);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "probe: %s\n", cudaGetErrorString(e)); return -(int)e - 1000; }
  return 0;
}
extern "C" int probe_plan_size() { return (int)sizeof(Plan); }

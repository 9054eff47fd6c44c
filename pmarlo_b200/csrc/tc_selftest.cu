// tc_selftest.cu -- one 128 x N x Kdim TF32 tile product through tcgen05.mma / TMEM,
// in the two operand layouts the production kernels use.  It exists so that the
// descriptor encodings are validated in isolation (tests/test_gpu_tc.py) before a
// kernel that depends on them is trusted:
//   mode 0: A (128 x Kdim), B (N x Kdim) row-major  -> K-major, no swizzle   (k-means scores)
//   mode 1: A (Kdim x 128), B (Kdim x N) row-major  -> MN-major, 128B swizzle with 32B base (Gram X^T X)
// D (128 x N) fp32 = A B^T resp. A^T B with fp32 accumulation in TMEM.
#include "tc05.cuh"

namespace pmb {

__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          int N, int Kdim, int mode, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // swizzle patterns are functions of the ABSOLUTE shared address: align the tile base by hand
  // (static __shared__ variables precede the dynamic region, which is only 16 B aligned)
  unsigned char* sA = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // 128 * Kdim * 4 bytes
  unsigned char* sB = sA + (size_t)128 * Kdim * 4;            // N * Kdim * 4 bytes (1024-aligned: Kdim % 8 == 0)
  uint32_t cols = 32;
  while (cols < (uint32_t)N) cols <<= 1;

  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, cols);

  uint32_t lbo, sbo;
  if (mode == 0) {
    lbo = 128u;
    sbo = (uint32_t)(Kdim / 4) * 128u;
    for (int i = tid; i < 128 * Kdim; i += 128) {
      const int r = i / Kdim, k = i - r * Kdim;
      *reinterpret_cast<float*>(sA + tc::off_kmajor(r, k, lbo, sbo)) = A[i];
    }
    for (int i = tid; i < N * Kdim; i += 128) {
      const int r = i / Kdim, k = i - r * Kdim;
      *reinterpret_cast<float*>(sB + tc::off_kmajor(r, k, lbo, sbo)) = B[i];
    }
  } else {
    lbo = (uint32_t)Kdim * 128u;
    sbo = 512u;
    for (int i = tid; i < Kdim * 128; i += 128) {
      const int k = i / 128, m = i - k * 128;
      *reinterpret_cast<float*>(sA + tc::off_mnmajor_sw128b32(m, k, lbo, sbo)) = A[i];
    }
    for (int i = tid; i < Kdim * N; i += 128) {
      const int k = i / N, n = i - k * N;
      *reinterpret_cast<float*>(sB + tc::off_mnmajor_sw128b32(n, k, lbo, sbo)) = B[i];
    }
  }
  fence_proxy_async_smem();       // generic-proxy smem writes -> visible to the tensor (async) proxy
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, N, mode, mode);
    const uint64_t layout = mode == 0 ? tc::kLayoutNone : tc::kLayoutSw128Base32;
    for (int s = 0; s < Kdim / 8; ++s) {
      const uint32_t step = mode == 0 ? (uint32_t)s * 2u * lbo : (uint32_t)s * 2u * sbo;
      const uint64_t da = tc::smem_desc(smem_u32(sA) + step, lbo, sbo, layout);
      const uint64_t db = tc::smem_desc(smem_u32(sB) + step, lbo, sbo, layout);
      tc::mma_tf32(tmem, da, db, idesc, s > 0 ? 1u : 0u);
    }
    tc::mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc::fence_after_sync();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[(size_t)row * N + c0 + j] = v[j];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, cols);
}

// Raw variant: the host supplies the two shared-memory operand IMAGES byte for byte plus every descriptor
// field, so that a layout hypothesis (16-bit operands, MN-major swizzles, ...) can be checked from Python
// without recompiling.  kind: 0 = tf32, 1 = f16 (fp16 / bf16 operands, selected by the idesc format fields).
__global__ void __launch_bounds__(128) tc_selftest_raw_kernel(const unsigned char* __restrict__ Aimg, uint32_t a_bytes,
                                                              const unsigned char* __restrict__ Bimg, uint32_t b_bytes,
                                                              int N, int nsteps, uint32_t lbo, uint32_t sbo, uint32_t layout,
                                                              uint32_t step_a, uint32_t step_b, uint32_t idesc, int kind,
                                                              float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* sA = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  unsigned char* sB = sA + ((a_bytes + 1023u) & ~1023u);
  uint32_t cols = 32;
  while (cols < (uint32_t)N) cols <<= 1;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, cols);
  for (uint32_t i = tid; i < a_bytes; i += 128) sA[i] = Aimg[i];
  for (uint32_t i = tid; i < b_bytes; i += 128) sB[i] = Bimg[i];
  fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    for (int s = 0; s < nsteps; ++s) {
      const uint64_t da = tc::smem_desc(smem_u32(sA) + (uint32_t)s * step_a, lbo, sbo, (uint64_t)layout);
      const uint64_t db = tc::smem_desc(smem_u32(sB) + (uint32_t)s * step_b, lbo, sbo, (uint64_t)layout);
      if (kind == 0) {
        tc::mma_tf32(tmem, da, db, idesc, s > 0 ? 1u : 0u);
      } else {
        const uint32_t acc = s > 0 ? 1u : 0u;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    tc::mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc::fence_after_sync();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[(size_t)row * N + c0 + j] = v[j];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, cols);
}

}  // namespace pmb

extern "C" int pmb_tc_selftest_raw(const void* Aimg, uint32_t a_bytes, const void* Bimg, uint32_t b_bytes, int N,
                                   int nsteps, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t step_a,
                                   uint32_t step_b, uint32_t idesc, int kind, float* D, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(Aimg && Bimg && D, "pmb_tc_selftest_raw: null pointer");
  PMB_REQUIRE(N >= 32 && N <= 256 && N % 32 == 0, "pmb_tc_selftest_raw: N must be a multiple of 32 in [32,256]");
  PMB_REQUIRE(nsteps >= 1 && nsteps <= 16 && (kind == 0 || kind == 1), "pmb_tc_selftest_raw: bad nsteps / kind");
  const size_t smem = (size_t)((a_bytes + 1023u) & ~1023u) + b_bytes + 2048;
  PMB_REQUIRE(smem <= 200 * 1024, "pmb_tc_selftest_raw: operand images too large");
  PMB_CUDA(cudaFuncSetAttribute(tc_selftest_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_raw_kernel<<<1, 128, smem, as_stream(stream)>>>(static_cast<const unsigned char*>(Aimg), a_bytes,
                                                              static_cast<const unsigned char*>(Bimg), b_bytes, N, nsteps,
                                                              lbo, sbo, layout, step_a, step_b, idesc, kind, D);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

extern "C" int pmb_tc_selftest(const float* A, const float* B, int N, int Kdim, int mode, float* D,
                               pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(A && B && D, "pmb_tc_selftest: null pointer");
  PMB_REQUIRE(mode == 0 || mode == 1, "pmb_tc_selftest: mode must be 0 or 1");
  PMB_REQUIRE(Kdim >= 8 && Kdim <= 64 && Kdim % 8 == 0, "pmb_tc_selftest: Kdim must be a multiple of 8 in [8,64]");
  PMB_REQUIRE(N >= 32 && N <= 256 && N % 32 == 0, "pmb_tc_selftest: N must be a multiple of 32 in [32,256]");
  const size_t smem = (size_t)(128 + N) * Kdim * 4 + 1024;
  PMB_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(A, B, N, Kdim, mode, D);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

// ubench_mma.cu -- cycles per tcgen05.mma on B200 as a function of kind (tf32 / f16), N, the number of
// K-step instructions per accumulator tile, and how often the issuing warp commits and waits.
// The operands are whatever bytes are in shared memory (zeros): only the timing matters.
// build ON THE GPU BOX (shared cudart, so that no static runtime is embedded):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/ubench_mma tools/ubench_mma.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {   // non-blocking test_wait in a hot loop
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

// One warp issues `tiles` accumulator tiles of `ksteps` MMAs each, round-robin over `nbuf` TMEM buffers;
// commit_every: commit (and, when wait != 0, wait for it) every that many tiles.
// layout: 0 = K-major no swizzle (lbo 128, sbo = ksteps * 256), 2 = K-major 128B swizzle (one 64-wide K atom per 1024 B)
template <int KIND>
__global__ void __launch_bounds__(128) mma_rate(int N, int ksteps, int tiles, int nbuf, int commit_every, int wait, int layout,
                                                int always_acc, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 0) {
    // idesc: D = F32; tf32: A/B format 2; f16: format 0; K-major both
    // layouts 11 / 12: MN-major operands (tf32: 128B swizzle with 32B base, atoms of 4 k-rows; 16-bit: 128B swizzle,
    // atoms of 8 k-rows), atoms contiguous along K -- the Gram kernel's operand layouts
    const bool mn = layout >= 10;
    const uint32_t idesc = (1u << 4) | (KIND == 0 ? ((2u << 7) | (2u << 10)) : ((1u << 7) | (1u << 10))) | (mn ? (3u << 15) : 0u) |
                           ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t lbo = layout == 0 ? 128u : (layout == 11 ? (uint32_t)ksteps * 8u * 128u : (layout == 12 ? (uint32_t)ksteps * 2u * 1024u : 0u));
    const uint32_t sbo = layout == 0 ? (uint32_t)ksteps * 256u : (layout == 11 ? 512u : 1024u);
    const uint64_t ltype = layout == 11 ? 1ull : (layout == 12 ? 2ull : (uint64_t)layout);
    const uint32_t aA = smem_u32(base), aB = aA + 64 * 1024;
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tmem + (uint32_t)((t % nbuf) * N);
      for (int s = 0; s < ksteps; ++s) {
        const uint32_t step = layout == 0 ? (uint32_t)s * 256u : (layout == 11 ? (uint32_t)s * 1024u : (layout == 12 ? (uint32_t)s * 2048u : (uint32_t)s * 32u));
        mma<KIND>(d, smem_desc(aA + step, lbo, sbo, ltype), smem_desc(aB + step, lbo, sbo, ltype), idesc, (s > 0) || always_acc);
      }
      if ((t + 1) % commit_every == 0) {
        commit(&bar);
        if (wait) { mbar_wait(&bar, phase); phase ^= 1u; asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      }
    }
    if (!wait || tiles % commit_every != 0) {
      commit(&bar);
      // drain: every earlier commit has completed a phase; wait for the last one
      const int ncommits = (commit_every > tiles ? 0 : tiles / commit_every) + 1;
      mbar_wait(&bar, (uint32_t)((ncommits - 1) & 1));
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// Tight issue loop: what a production MMA warp can do per chunk -- descriptors precomputed, ring index by mask,
// ONE elect per chunk, both k-step MMAs and the commit inside a single asm block.
__global__ void __launch_bounds__(128) mma_tight(int N, int tiles, int do_commit, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = smem_desc(smem_u32(base), 128u, 512u, 0), db = smem_desc(smem_u32(base) + 64 * 1024, 128u, 512u, 0);
    const uint32_t bar0 = smem_u32(&bar[0]);
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tmem + (uint32_t)((t & 1) * N);
      const uint64_t dbc = db + (uint64_t)((uint32_t)(t & 7) * 512u);   // "chunk" of the centre operand
      if (do_commit)
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %5, 0;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %4, %5, 1;\n\t"
                     "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t}"
                     ::"r"(d), "l"(da), "l"(dbc), "l"(da + 16), "l"(dbc + 16), "r"(idesc), "r"(bar0 + (uint32_t)(t & 3) * 8u) : "memory");
      else
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %5, 0;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %4, %5, 1;\n\t}"
                     ::"r"(d), "l"(da), "l"(dbc), "l"(da + 16), "l"(dbc + 16), "r"(idesc) : "memory");
    }
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar0 + 24u) : "memory");
    // the final commit on bar[3] completes phase (number of earlier commits on bar[3]) & 1
    const int n3 = do_commit ? tiles / 4 : 0;
    mbar_wait(&bar[3], (uint32_t)(n3 & 1));
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// Ring handshake: one MMA warp and `nepi` consumer warps per TMEM buffer, exactly the K6 skeleton without
// producers / finalisers.  N columns per buffer, `ring` buffers, 2 K-step MMAs per buffer.
// epi_work: 0 = consumers only wait and arrive, 1 = they tcgen05.ld their quarter of the buffer (x32 blocks).
__global__ void __launch_bounds__(1024) mma_ring(int N, int ring, int nepi, int tiles, int epi_work, int spin, int batch, long long* cyc, uint32_t* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[8], empty[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], nepi); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const int mma_warp = ring * nepi;     // consumer warps 0 .. ring * nepi - 1, then the MMA warp
  if (warp == mma_warp) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = smem_desc(smem_u32(base), 128u, 512u, 0), db = smem_desc(smem_u32(base) + 64 * 1024, 128u, 512u, 0);
    const uint32_t full0 = smem_u32(&full[0]);
    long long waited = 0;
    const long long t0 = clock64();
    uint32_t s = 0, use = 0;
    for (int t = 0; t < tiles; ++t) {
      const long long w0 = clock64();
      if (spin & 1) mbar_spin(&empty[s], (use & 1u) ^ 1u); else mbar_wait(&empty[s], (use & 1u) ^ 1u);
      waited += clock64() - w0;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d = tmem + s * (uint32_t)N;
      const uint64_t dbc = db + (uint64_t)((uint32_t)(t & 7) * 512u);
      asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                   "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %5, 0;\n\t"
                   "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %4, %5, 1;\n\t"
                   "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t}"
                   ::"r"(d), "l"(da), "l"(dbc), "l"(da + 16), "l"(dbc + 16), "r"(idesc), "r"(full0 + s * 8u) : "memory");
      if (++s == (uint32_t)ring) { s = 0; ++use; }
    }
    const long long t1 = clock64();
    if (lane == 0) { cyc[2 * blockIdx.x] = t1 - t0; cyc[2 * blockIdx.x + 1] = waited; }
  } else if (warp < mma_warp) {
    const int g = warp / nepi, q = warp % nepi;     // buffer and (when nepi == 4) lane quarter
    uint32_t acc = 0;
    int use = 0;
    for (int t = g; t < tiles; t += ring, ++use) {
      if (spin & 2) mbar_spin(&full[g], (uint32_t)(use & 1)); else mbar_wait(&full[g], (uint32_t)(use & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (epi_work) {
        const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * N);
        const int per = N / (nepi >= 4 ? nepi / 4 : 1);      // columns this warp reads
        const int c0 = (nepi >= 4 ? (q / 4) : 0) * per;
        for (int c = 0; c < per; c += 32) {
          uint32_t r[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(tb + (uint32_t)(c0 + c))
              : "memory");
          if (batch <= 1 || ((c / 32) % batch) == batch - 1 || c + 32 >= per) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= r[i];
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[g])) : "memory");
    }
    if (acc == 0x12345u) sink[0] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 2 * 148 * sizeof(long long));
  const size_t smem = 194 * 1024;
  cudaFuncSetAttribute(mma_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(mma_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(mma_tight, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int N = 128; N <= 256; N += 128)
    for (int dc = 0; dc < 2; ++dc) {
      for (int rep = 0; rep < 2; ++rep) {
        mma_tight<<<148, 128, smem>>>(N, 2048, dc, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("tight error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < 148; ++i) mean += (double)h[i];
      printf("tight f16 N=%d ksteps=2 commit=%d: %.0f cycles/tile\n", N, dc, mean / 148 / 2048);
    }
  {
    cudaFuncSetAttribute(mma_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    uint32_t* sink;
    cudaMalloc(&sink, 4);
    struct R { int N, ring, nepi, work, spin, batch; };
    const R rs[] = {{128, 4, 4, 0, 0, 1}, {128, 4, 4, 0, 1, 1}, {128, 4, 4, 0, 3, 1}, {128, 4, 4, 1, 0, 1}, {128, 4, 4, 1, 1, 1},
                    {128, 4, 4, 1, 3, 1}, {128, 4, 4, 1, 0, 2}, {128, 4, 4, 1, 1, 2}, {128, 4, 4, 1, 0, 4}, {128, 4, 8, 1, 0, 1},
                    {128, 4, 8, 1, 1, 2}, {128, 2, 8, 1, 0, 1}, {256, 2, 8, 1, 1, 2}, {128, 4, 1, 0, 3, 1}};
    for (const R& r : rs) {
      const int threads = (r.ring * r.nepi + 1) * 32;
      for (int rep = 0; rep < 2; ++rep) {
        mma_ring<<<148, threads, smem>>>(r.N, r.ring, r.nepi, 4096, r.work, r.spin, r.batch, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ring error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[296];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double mean = 0, w = 0;
      for (int i = 0; i < 148; ++i) { mean += (double)h[2 * i]; w += (double)h[2 * i + 1]; }
      printf("ring N=%d ring=%d consumers/buffer=%d work=%d spin=%d batch=%d: %.0f cycles/buffer (MMA warp waits %.0f), %.0f cycles per 128 columns\n", r.N, r.ring,
             r.nepi, r.work, r.spin, r.batch, mean / 148 / 4096, w / 148 / 4096, mean / 148 / 4096 * 128 / r.N);
    }
  }
  if (getenv("UB_RING_ONLY")) return 0;
  if (getenv("UB_MN_ONLY")) {
    struct CfgM { int kind, N, ksteps, layout; };
    const CfgM cm[] = {{0, 256, 16, 0}, {0, 256, 16, 11}, {0, 128, 16, 11}, {1, 256, 16, 0}, {1, 256, 16, 12}, {1, 128, 16, 12}, {0, 256, 8, 11}, {1, 256, 8, 12}};
    for (const CfgM& c : cm) {
      const int tiles = 1024;
      for (int rep = 0; rep < 2; ++rep) {
        if (c.kind == 0) mma_rate<0><<<148, 128, smem>>>(c.N, c.ksteps, tiles, 2, 1, 0, c.layout, 0, cyc);
        else mma_rate<1><<<148, 128, smem>>>(c.N, c.ksteps, tiles, 2, 1, 0, c.layout, 0, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mn error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < 148; ++i) mean += (double)h[i];
      mean /= 148;
      printf("kind=%s N=%d ksteps=%d layout=%d (%s): %.1f cycles/MMA\n", c.kind == 0 ? "tf32" : "bf16", c.N, c.ksteps, c.layout,
             c.layout >= 10 ? "MN-major" : "K-major", mean / ((double)tiles * c.ksteps));
    }
    return 0;
  }
  struct Cfg { int kind, N, ksteps, nbuf, commit_every, wait, layout, always_acc; };
  const Cfg cfgs[] = {
      {0, 256, 4, 2, 1, 0, 0}, {0, 256, 16, 2, 1, 0, 0}, {0, 256, 4, 2, 1, 1, 0}, {0, 128, 4, 4, 1, 0, 0},
      {1, 256, 2, 2, 1, 0, 0}, {1, 256, 16, 2, 1, 0, 0}, {1, 128, 2, 4, 1, 0, 0}, {1, 128, 2, 4, 1, 1, 0},
      {1, 128, 2, 4, 4, 0, 0}, {1, 128, 16, 4, 1, 0, 0}, {1, 256, 2, 2, 1, 1, 0}, {1, 128, 13, 4, 1, 0, 0},
      {1, 256, 4, 2, 1, 0, 2}, {1, 128, 4, 4, 1, 0, 2}, {0, 256, 8, 2, 1, 0, 2}, {1, 64, 2, 8, 1, 0, 0},
      // always accumulate (no overwrite at tile starts); one buffer only; never commit inside the loop
      {1, 128, 2, 4, 1, 0, 0, 1}, {1, 128, 2, 1, 1, 0, 0, 0}, {1, 128, 2, 1, 1, 0, 0, 1}, {1, 128, 2, 4, 1 << 30, 0, 0, 0},
      {1, 128, 2, 4, 1 << 30, 0, 0, 1}, {1, 128, 1, 4, 1 << 30, 0, 0, 1}, {1, 128, 4, 4, 1 << 30, 0, 0, 1}, {1, 128, 8, 4, 1 << 30, 0, 0, 0},
  };
  for (const Cfg& c : cfgs) {
    const int tiles = 2048;
    for (int rep = 0; rep < 2; ++rep) {
      if (c.kind == 0) mma_rate<0><<<148, 128, smem>>>(c.N, c.ksteps, tiles, c.nbuf, c.commit_every, c.wait, c.layout, c.always_acc, cyc);
      else mma_rate<1><<<148, 128, smem>>>(c.N, c.ksteps, tiles, c.nbuf, c.commit_every, c.wait, c.layout, c.always_acc, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += (double)h[i];
    mean /= 148;
    const double per_mma = mean / ((double)tiles * c.ksteps);
    const double flop = 2.0 * 128 * c.N * (c.kind == 0 ? 8 : 16);
    printf("kind=%s N=%3d ksteps=%2d nbuf=%d commit_every=%d wait=%d layout=%d acc=%d: %.1f cycles/MMA, %.0f cycles/tile, %.0f flop/clk/SM\n",
           c.kind == 0 ? "tf32" : "f16 ", c.N, c.ksteps, c.nbuf, c.commit_every > 9999 ? 0 : c.commit_every, c.wait, c.layout, c.always_acc, per_mma, mean / tiles, flop / per_mma);
  }
  return 0;
}

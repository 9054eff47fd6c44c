// ubench_tmem.cu -- TMEM -> register bandwidth (tcgen05.ld 32x32b) per SM for 4 / 8 / 16 reader warps.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tmem tools/ubench_tmem.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) x ^= r[i];
  return x;
}

// mode 0: ld, wait, ld, wait ...   mode 1: two loads in flight before each wait
__global__ void __launch_bounds__(512) tmem_read(int iters, int mode, long long* cyc, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32 % 512);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      acc ^= ld<32>(base + (uint32_t)((i * 64) & 255));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
      const uint32_t a = ld<32>(base + (uint32_t)((i * 64) & 255));
      const uint32_t b = ld<32>(base + (uint32_t)((i * 64 + 32) & 255));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= a ^ b;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 1024 * 8); cudaMalloc(&sink, 64);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16}) {
      tmem_read<<<148, warps * 32>>>(iters, mode, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)warps * iters * (mode ? 2 : 1) * 4096.0;
      printf("mode %d warps %2d: %.1f bytes/clk/SM (%lld cycles, %.0f cycles per LDTM.x32 per warp)\n", mode, warps,
             bytes / mx, mx, (double)mx / (iters * (mode ? 2 : 1)));
    }
  return 0;
}

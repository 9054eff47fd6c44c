// ubench.cu -- B200 micro-measurements that drive the MLE / eigen kernel designs (not product code):
//   1. fp64 / fp32 FMA issue rate per SM          2. flag all-gather latency between CTAs through L2
//   3. cluster (16 CTA) DSMEM all-gather + barrier.cluster round time
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <typename T, int CH>
__global__ void fma_rate(T* out, int iters, T a, T b) {
  T acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = (T)(threadIdx.x + c);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = acc[c] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c];
  if (s == (T)12345.678) out[0] = s;
}

__global__ void rcp_rate(double* out, int iters, double a) {
  double acc[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) acc[c] = 1.0 + threadIdx.x + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = 1.0 / (acc[c] + a);
  }
  double s = acc[0] + acc[1] + acc[2] + acc[3];
  if (s == 12345.678) out[0] = s;
}

// LL all-gather: every CTA publishes `per` doubles per round, everyone gathers all K.
__device__ __forceinline__ void ll_store(unsigned long long* slot, double v, unsigned int tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (bits & 0xffffffffull) | ((unsigned long long)tag << 32);
  const unsigned long long w1 = (bits >> 32) | ((unsigned long long)tag << 32);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_try(const unsigned long long* slot, unsigned int tag, double& out) {
  unsigned long long w0, w1;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
  out = __longlong_as_double((long long)(((w1 & 0xffffffffull) << 32) | (w0 & 0xffffffffull)));
  return (unsigned int)(w0 >> 32) == tag && (unsigned int)(w1 >> 32) == tag;
}
__global__ void ll_allgather(unsigned long long* xb, int K, int rounds, long long* cyc, double* sink) {
  extern __shared__ double us[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int gwarp = (blockIdx.x * blockDim.x + tid) >> 5;
  double chk = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 1; r <= rounds; ++r) {
    unsigned long long* base = xb + (size_t)(r & 1) * K * 2;
    if (gwarp < K && lane == 0) ll_store(base + (size_t)gwarp * 2, (double)(r + gwarp), (unsigned)r);
    for (int i = tid; i < K; i += blockDim.x) {
      double v;
      unsigned spins = 0;
      while (!ll_try(base + (size_t)i * 2, (unsigned)r, v)) if (++spins > (1u << 22)) __trap();
      us[i] = v;
    }
    __syncthreads();
    chk += us[(tid * 7 + r) % K];
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (chk == 1.2345) sink[0] = chk;
}

// flag-based variant: each CTA writes its values with plain stores, then one release flag; consumers poll flags
__global__ void flag_allgather(double* ub, unsigned int* flags, int K, int rounds, long long* cyc, double* sink) {
  extern __shared__ double us[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  double chk = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 1; r <= rounds; ++r) {
    double* base = ub + (size_t)(r & 1) * K;
    const int row = blockIdx.x * nw + warp;
    if (row < K && lane == 0) base[row] = (double)(r + row);
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * 32), "r"((unsigned)r) : "memory");
    }
    if (tid < gridDim.x) {
      unsigned v, spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + tid * 32) : "memory");
        if (++spins > (1u << 22)) __trap();
      } while (v < (unsigned)r);
    }
    __syncthreads();
    for (int i = tid; i < K; i += blockDim.x) us[i] = __ldcg(base + i);
    __syncthreads();
    chk += us[(tid * 7 + r) % K];
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (chk == 1.2345) sink[0] = chk;
}

// cluster all-gather through DSMEM: each CTA writes its `per` values into every CTA's smem, then barrier.cluster
__global__ void cluster_allgather(int K, int rounds, long long* cyc, double* sink) {
  extern __shared__ double us[];   // 2 x K
  cg::cluster_group cl = cg::this_cluster();
  const int nc = cl.num_blocks(), rank = cl.block_rank();
  const int tid = threadIdx.x;
  const int per = (K + nc - 1) / nc;
  double chk = 0.0;
  cl.sync();
  const long long t0 = clock64();
  for (int r = 1; r <= rounds; ++r) {
    double* buf = us + (size_t)(r & 1) * K;
    // thread t handles (dest cta = t / per, value = t % per)
    for (int t = tid; t < nc * per; t += blockDim.x) {
      const int dst = t / per, k = rank * per + (t - dst * per);
      if (k < K) {
        double* remote = cl.map_shared_rank(buf, dst);
        remote[k] = (double)(r + k);
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    chk += buf[(tid * 7 + r) % K];
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (chk == 1.2345) sink[0] = chk;
  cl.sync();
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, clock attr %d kHz\n", prop.name, sms, clk_khz);
  double* dout; CK(cudaMalloc(&dout, 1024));
  float* fout = reinterpret_cast<float*>(dout);
  const int iters = 4096;
  for (int threads : {256, 1024}) {
    {
      float ms = time_ms([&] { fma_rate<double, 8><<<sms * (2048 / threads), threads>>>(dout, iters, 1.0000001, 1e-9); });
      double ops = (double)sms * 2048 * 8.0 * iters;
      printf("DFMA  threads/CTA=%4d: %.3f ms, %.1f Gfma/s -> %.2f fma/clk/SM at 1.965 GHz, %.2f TFLOP/s\n", threads, ms, ops / ms / 1e6,
             ops / (ms * 1e-3) / sms / 1.965e9, 2 * ops / ms / 1e9);
    }
    {
      float ms = time_ms([&] { fma_rate<float, 8><<<sms * (2048 / threads), threads>>>(fout, iters, 1.0000001f, 1e-9f); });
      double ops = (double)sms * 2048 * 8.0 * iters;
      printf("FFMA  threads/CTA=%4d: %.3f ms, %.2f fma/clk/SM, %.2f TFLOP/s\n", threads, ms, ops / (ms * 1e-3) / sms / 1.965e9, 2 * ops / ms / 1e9);
    }
  }
  {
    float ms = time_ms([&] { fma_rate<double, 2><<<sms, 256>>>(dout, iters, 1.0000001, 1e-9); });
    printf("DFMA latency-bound (8 warps/SM, 2 chains): %.1f cycles per dependent fma pair-step\n", ms * 1e-3 * 1.965e9 / iters);
    ms = time_ms([&] { fma_rate<double, 1><<<sms, 32>>>(dout, iters, 1.0000001, 1e-9); });
    printf("DFMA dependent latency: %.1f cycles\n", ms * 1e-3 * 1.965e9 / iters);
    ms = time_ms([&] { rcp_rate<<<sms * 2, 1024>>>(dout, 512, 0.5); });
    printf("fp64 IEEE 1/x: %.2f per clk per SM\n", (double)sms * 2048 * 4 * 512 / (ms * 1e-3) / sms / 1.965e9);
  }
  // exchanges
  const int K = 1000, rounds = 2000;
  long long* cyc; CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
  long long h[1024];
  {
    unsigned long long* xb; CK(cudaMalloc(&xb, 4 * K * sizeof(double)));
    for (int threads : {256, 1024}) {
      CK(cudaMemset(xb, 0, 4 * K * sizeof(double)));
      int grid = (K * 32 + threads - 1) / threads;
      if (grid > sms) grid = sms;
      int Kk = K; int rr = rounds; double* sink = dout;
      void* args[] = {&xb, &Kk, &rr, &cyc, &sink};
      CK(cudaLaunchCooperativeKernel((void*)ll_allgather, dim3(grid), dim3(threads), args, K * sizeof(double), 0));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("LL all-gather K=%d grid=%d threads=%d: %.0f cycles / round\n", K, grid, threads, (double)mx / rounds);
    }
    double* ub; unsigned int* flags;
    CK(cudaMalloc(&ub, 2 * K * sizeof(double))); CK(cudaMalloc(&flags, 148 * 32 * 4));
    CK(cudaMemset(flags, 0, 148 * 32 * 4));
    {
      int threads = 256, grid = 125; int Kk = K; int rr = rounds; double* sink = dout;
      void* args[] = {&ub, &flags, &Kk, &rr, &cyc, &sink};
      CK(cudaLaunchCooperativeKernel((void*)flag_allgather, dim3(grid), dim3(threads), args, K * sizeof(double), 0));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("flag all-gather K=%d grid=%d threads=%d: %.0f cycles / round\n", K, grid, threads, (double)mx / rounds);
    }
  }
  for (int csz : {8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csz); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 2 * K * sizeof(double);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (csz > 8) CK(cudaFuncSetAttribute(cluster_allgather, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_allgather, K, rounds, cyc, dout);
    if (e != cudaSuccess) { printf("cluster %d launch failed: %s\n", csz, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, csz * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < csz; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("cluster(%d) DSMEM all-gather K=%d: %.0f cycles / round\n", csz, K, (double)mx / rounds);
  }
  printf("done\n");
  return 0;
}

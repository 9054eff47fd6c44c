// kmeans.cu -- K6: nearest-centre assignment fused with the Lloyd accumulation.
//
// Register-tiled SIMT path (tiny / moderate D): one frame per thread, its
// coordinates in registers, centres staged in shared memory as fp32 (broadcast
// float4 reads), direct-difference distances.  Labels must equal the fp64
// oracle's argmin_k sum_d (y_d - c_kd)^2 (first minimum wins): the kernel keeps
// the best and second-best fp32 distances and, whenever the two cannot be
// separated by a rigorous rounding-error bound, re-evaluates that frame against
// all centres in fp64 with the oracle's exact operation order.
//
//   r~ = sqrt(d~) is within gamma of sqrt(sum (x~-c~)^2) and that is within
//   eta = 2^-24 (|y| + max_k |c_k|) of the exact |y - c|  =>  the label is
//   certain iff  sqrt(d1~)(1+gamma) + eta < sqrt(d2~)(1-gamma) - eta.
//
// Accumulation: 32 consecutive frames of a trajectory mostly share labels, so
// each warp does a segmented (run-length) reduction by shuffles and only run
// heads issue fp64 atomics; counts and inertia likewise.
#include "common.cuh"

namespace pmb {

constexpr int kKmThreads = 256;
constexpr int kKmSmemCenters = 64 * 1024;

template <typename TY>
struct KmParams {
  const TY* Y;
  int64_t n;
  int D;
  int64_t ld;
  const double* centers;
  int K;
  int32_t* labels;
  double* sums;
  int64_t* counts;
  double* inertia;
  int64_t* n_rechecked;
  int KT;  // centres per shared-memory tile
};

template <int DP>
__device__ __forceinline__ float dist_direct(const float (&x)[DP], const float* __restrict__ c) {
  float acc = 0.f;
  if constexpr (DP % 4 == 0) {
#pragma unroll
    for (int q = 0; q < DP; q += 4) {
      const float4 cv = *reinterpret_cast<const float4*>(c + q);
      float t;
      t = x[q + 0] - cv.x; acc = fmaf(t, t, acc);
      t = x[q + 1] - cv.y; acc = fmaf(t, t, acc);
      t = x[q + 2] - cv.z; acc = fmaf(t, t, acc);
      t = x[q + 3] - cv.w; acc = fmaf(t, t, acc);
    }
  } else {
#pragma unroll
    for (int q = 0; q < DP; q += 2) {
      const float2 cv = *reinterpret_cast<const float2*>(c + q);
      float t;
      t = x[q + 0] - cv.x; acc = fmaf(t, t, acc);
      t = x[q + 1] - cv.y; acc = fmaf(t, t, acc);
    }
  }
  return acc;
}

template <typename TY, int DP>
__global__ void __launch_bounds__(kKmThreads, 2) kmeans_assign_kernel(KmParams<TY> p) {
  extern __shared__ __align__(16) float s_c[];  // KT x DP
  __shared__ float s_redf[kKmThreads / 32];
  __shared__ double s_redd[kKmThreads / 32];
  __shared__ int s_redi[kKmThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = p.D, K = p.K, KT = p.KT;
  const bool single_tile = (KT >= K);
  const float gamma = (float)(DP + 4) * 5.9604645e-8f;

  float cmax2 = 0.f;  // max_k |c_k|^2 over the centres this thread has staged
  auto stage = [&](int k0) {
    const int kn = (K - k0) < KT ? (K - k0) : KT;
    for (int i = tid; i < kn * DP; i += kKmThreads) {
      const int k = i / DP, q = i - k * DP;
      s_c[i] = (q < D) ? (float)p.centers[(size_t)(k0 + k) * D + q] : 0.f;
    }
    __syncthreads();
    for (int k = tid; k < kn; k += kKmThreads) {
      float nn = 0.f;
      for (int q = 0; q < DP; ++q) nn = fmaf(s_c[k * DP + q], s_c[k * DP + q], nn);
      cmax2 = fmaxf(cmax2, nn);
    }
  };
  auto block_cmax = [&]() -> float {
    float v = cmax2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) s_redf[warp] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < kKmThreads / 32; ++w) r = fmaxf(r, s_redf[w]);
    __syncthreads();
    return sqrtf(r) * 1.0001f;
  };

  float cmax = 0.f;
  if (single_tile) {
    stage(0);
    cmax = block_cmax();
  }

  double inertia_acc = 0.0;
  int recheck_acc = 0;
  const int64_t n_tiles = (p.n + kKmThreads - 1) / kKmThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = tile * kKmThreads + tid;
    const bool valid = row < p.n;
    float x[DP];
    float xn2 = 0.f;
#pragma unroll
    for (int q = 0; q < DP; ++q) {
      x[q] = (valid && q < D) ? (float)p.Y[row * p.ld + q] : 0.f;
      xn2 = fmaf(x[q], x[q], xn2);
    }
    float best = 3.4e38f, second = 3.4e38f;
    int bi = 0;
    if (single_tile) {
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        const float dd = dist_direct<DP>(x, s_c + k * DP);
        if (dd < best) { second = best; best = dd; bi = k; }
        else if (dd < second) second = dd;
      }
    } else {
      cmax2 = 0.f;
      for (int k0 = 0; k0 < K; k0 += KT) {
        __syncthreads();
        stage(k0);
        __syncthreads();
        const int kn = (K - k0) < KT ? (K - k0) : KT;
#pragma unroll 4
        for (int k = 0; k < kn; ++k) {
          const float dd = dist_direct<DP>(x, s_c + k * DP);
          if (dd < best) { second = best; best = dd; bi = k0 + k; }
          else if (dd < second) second = dd;
        }
      }
      __syncthreads();
      cmax = block_cmax();
    }

    // ---- certainty test, fp64 re-check of near ties
    double dbest = (double)best;
    if (valid) {
      const float eta = 5.9604645e-8f * 1.01f * (sqrtf(xn2) * 1.0001f + cmax);
      const float s1 = sqrtf(best) * (1.f + gamma) + eta;
      const float s2 = sqrtf(second) * (1.f - gamma) - eta;
      if (!(s1 < s2) && K > 1) {
        ++recheck_acc;
        double bd = 1.7976931348623157e308;
        int bk = 0;
        for (int k = 0; k < K; ++k) {
          double acc = 0.0;
          const double* c = p.centers + (size_t)k * D;
          for (int q = 0; q < D; ++q) {
            const double t = __dsub_rn((double)p.Y[row * p.ld + q], c[q]);
            acc = __dadd_rn(acc, __dmul_rn(t, t));
          }
          if (acc < bd) { bd = acc; bk = k; }
        }
        bi = bk;
        dbest = bd;
      }
      p.labels[row] = bi;
      inertia_acc += dbest;
    }

    // ---- fused accumulation (segmented by runs of equal labels)
    if (p.sums != nullptr) {
      const int lab = valid ? bi : -1;
      const int prev = __shfl_up_sync(0xffffffffu, lab, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || lab != prev);
      const int my_head = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
      bool take[5];
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        const int oh = __shfl_down_sync(0xffffffffu, my_head, 1 << s);
        take[s] = (lane + (1 << s) < 32) && (oh == my_head);
      }
      const bool is_head = (lane == my_head) && lab >= 0;
      for (int q = 0; q < D; ++q) {
        double v;
        if constexpr (sizeof(TY) == 8) v = valid ? (double)p.Y[row * p.ld + q] : 0.0;
        else v = (double)x[q < DP ? q : 0];
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          const double o = __shfl_down_sync(0xffffffffu, v, 1 << s);
          if (take[s]) v += o;
        }
        if (is_head) atomicAdd(p.sums + (size_t)lab * D + q, v);
      }
      if (is_head && p.counts != nullptr) {
        const unsigned above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
        const int next = above ? (__ffs(above) - 1) : 32;
        atomicAdd(reinterpret_cast<unsigned long long*>(p.counts + lab),
                  (unsigned long long)(next - lane));
      }
    }
  }
  // ---- block reductions for inertia / re-check count
  if (p.inertia != nullptr || p.n_rechecked != nullptr) {
    double v = warp_sum(inertia_acc);
    int r = recheck_acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0) { s_redd[warp] = v; s_redi[warp] = r; }
    __syncthreads();
    if (tid == 0) {
      double tv = 0.0;
      int tr = 0;
      for (int w = 0; w < kKmThreads / 32; ++w) { tv += s_redd[w]; tr += s_redi[w]; }
      if (p.inertia != nullptr) atomicAdd(p.inertia, tv);
      if (p.n_rechecked != nullptr && tr)
        atomicAdd(reinterpret_cast<unsigned long long*>(p.n_rechecked), (unsigned long long)tr);
    }
  }
}

template <typename TY, int DP>
static int launch_assign(KmParams<TY> p, cudaStream_t st) {
  int KT = kKmSmemCenters / (DP * 4);
  if (KT > p.K) KT = p.K;
  p.KT = KT;
  const size_t smem = (size_t)KT * DP * 4;
  PMB_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<TY, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  const int64_t n_tiles = (p.n + kKmThreads - 1) / kKmThreads;
  const int grid = (int)(n_tiles < 2 * kNumSMs ? n_tiles : 2 * kNumSMs);
  kmeans_assign_kernel<TY, DP><<<grid, kKmThreads, smem, st>>>(p);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

template <typename TY>
static int dispatch_assign(KmParams<TY> p, cudaStream_t st) {
  const int D = p.D;
  if (D <= 2) return launch_assign<TY, 2>(p, st);
  if (D <= 4) return launch_assign<TY, 4>(p, st);
  if (D <= 6) return launch_assign<TY, 6>(p, st);
  if (D <= 8) return launch_assign<TY, 8>(p, st);
  if (D <= 12) return launch_assign<TY, 12>(p, st);
  if (D <= 16) return launch_assign<TY, 16>(p, st);
  if (D <= 24) return launch_assign<TY, 24>(p, st);
  if (D <= 32) return launch_assign<TY, 32>(p, st);
  if (D <= 48) return launch_assign<TY, 48>(p, st);
  if (D <= 64) return launch_assign<TY, 64>(p, st);
  set_error("pmb_kmeans_assign: D=%d > 64 is not supported by the register-tiled path", D);
  return PMB_EUNSUPPORTED;
}

__global__ void kmeans_update_kernel(double* __restrict__ centers, const double* __restrict__ sums,
                                     const int64_t* __restrict__ counts, int K, int D,
                                     double* __restrict__ out_shift2) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  double sh = 0.0;
  if (k < K) {
    const int64_t c = counts[k];
    if (c > 0) {
      const double inv = 1.0 / (double)c;
      for (int q = 0; q < D; ++q) {
        const double nv = sums[(size_t)k * D + q] / (double)c;
        const double t = nv - centers[(size_t)k * D + q];
        sh = fma((double)c * t, t, sh);
        centers[(size_t)k * D + q] = nv;
      }
      (void)inv;
    }
  }
  if (out_shift2 != nullptr) {
    sh = warp_sum(sh);
    if ((threadIdx.x & 31) == 0 && sh != 0.0) atomicAdd(out_shift2, sh);
  }
}

// kmeans_tc.cu
bool kmeans_tc_supported(int D, int K);
size_t kmeans_tc_ws_bytes(int64_t n, int D, int K);
int kmeans_tc_assign(const float* Y, int64_t n, int D, int64_t ld, const double* centers, int K, int32_t* labels,
                     const int32_t* hints, double* sums, int64_t* counts, double* inertia, int64_t* n_rechecked, void* ws,
                     float* dbg_scores, cudaStream_t st);

int kmeans_tc_debug_counters(int64_t* out16);
constexpr int64_t kKmTcMinFrames = 16384;   // below this the launch is latency-bound either way

}  // namespace pmb

extern "C" int pmb_debug_counters_kmeans(int64_t* out16) {
  using namespace pmb;
  PMB_REQUIRE(out16 != nullptr, "pmb_debug_counters_kmeans: null pointer");
  return kmeans_tc_debug_counters(out16);
}

extern "C" size_t pmb_kmeans_assign_ws_bytes(int64_t n, int D, int K) {
  if (n <= 0 || !pmb::kmeans_tc_supported(D, K)) return 0;
  return pmb::kmeans_tc_ws_bytes(n, D, K);
}

extern "C" int pmb_kmeans_tc_scores(const float* Y, int64_t n, int D, int64_t ld, const double* centers, int K,
                                    int32_t* labels, float* scores, void* ws, size_t ws_bytes,
                                    pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n > 0 && D > 0 && K > 0 && ld >= D, "pmb_kmeans_tc_scores: bad sizes");
  PMB_REQUIRE(Y && centers && labels && scores && ws, "pmb_kmeans_tc_scores: null pointer");
  PMB_REQUIRE(kmeans_tc_supported(D, K), "pmb_kmeans_tc_scores: D=%d K=%d does not fit the tensor path", D, K);
  PMB_REQUIRE(ws_bytes >= kmeans_tc_ws_bytes(n, D, K), "pmb_kmeans_tc_scores: workspace too small");
  return kmeans_tc_assign(Y, n, D, ld, centers, K, labels, nullptr, nullptr, nullptr, nullptr, nullptr, ws, scores,
                          as_stream(stream));
}

extern "C" int pmb_kmeans_assign(const void* Y, int y_f64, int64_t n, int D, int64_t ld,
                                 const double* centers, int K, int32_t* labels, double* sums,
                                 int64_t* counts, double* inertia, int64_t* n_rechecked,
                                 const int32_t* hints, void* ws, size_t ws_bytes, int impl,
                                 pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n >= 0 && D > 0 && K > 0 && ld >= D, "pmb_kmeans_assign: bad sizes");
  PMB_REQUIRE(impl >= 0 && impl <= 2, "pmb_kmeans_assign: impl must be 0 (auto), 1 (SIMT) or 2 (tcgen05)");
  if (n == 0) return PMB_OK;
  PMB_REQUIRE(Y && centers && labels, "pmb_kmeans_assign: null pointer");
  PMB_REQUIRE((sums == nullptr) == (counts == nullptr), "pmb_kmeans_assign: sums and counts go together");
  {
    const bool tc_ok = !y_f64 && kmeans_tc_supported(D, K) && ws != nullptr && ws_bytes >= kmeans_tc_ws_bytes(n, D, K);
    if (impl == 2 && !tc_ok) {
      set_error("pmb_kmeans_assign: tcgen05 path needs float32 Y, a workspace of pmb_kmeans_assign_ws_bytes "
                "and centres that fit shared memory (D=%d, K=%d)", D, K);
      return PMB_EUNSUPPORTED;
    }
    if (impl == 2 || (impl == 0 && tc_ok && n >= kKmTcMinFrames && (int64_t)D * K >= 2048))
      return kmeans_tc_assign(static_cast<const float*>(Y), n, D, ld, centers, K, labels, hints, sums, counts, inertia,
                              n_rechecked, ws, nullptr, as_stream(stream));
  }
  if (y_f64) {
    KmParams<double> p{static_cast<const double*>(Y), n, D, ld, centers, K, labels, sums, counts,
                       inertia, n_rechecked, 0};
    return dispatch_assign<double>(p, as_stream(stream));
  }
  KmParams<float> p{static_cast<const float*>(Y), n, D, ld, centers, K, labels, sums, counts,
                    inertia, n_rechecked, 0};
  return dispatch_assign<float>(p, as_stream(stream));
}

extern "C" int pmb_kmeans_update(double* centers, const double* sums, const int64_t* counts, int K,
                                 int D, double* out_shift2, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(K > 0 && D > 0, "pmb_kmeans_update: bad sizes");
  PMB_REQUIRE(centers && sums && counts, "pmb_kmeans_update: null pointer");
  kmeans_update_kernel<<<(K + 127) / 128, 128, 0, as_stream(stream)>>>(centers, sums, counts, K, D, out_shift2);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

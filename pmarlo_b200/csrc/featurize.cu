// featurize.cu -- K1: backbone dihedrals (angle / cos / sin) and pair distances.
//
// Bandwidth-bound: 12*A bytes read + 4*F bytes written per frame.  Frames are
// contiguous in memory, so a tile of FT frames is ONE 1-D bulk (TMA) copy into
// shared memory (double-buffered, completion on an mbarrier); features are
// computed from the staged tile into a shared output tile that leaves with one
// bulk store.  Persistent grid: 2 CTAs per SM.
//
// Math follows mdtraj's _dihedral (oracle/featurize.py) in fp32:
//   b1=x1-x0, b2=x2-x1, b3=x3-x2, c1=b2 x b3, c2=b1 x b2,
//   y=(b1.c1)|b2|, x=c1.c2, angle=atan2(y,x), cos=x/hypot, sin=y/hypot.
#include "common.cuh"

namespace pmb {

constexpr int kFeatThreads = 256;   // default; the launch picks a multiple of n_units (see pmb_featurize)
constexpr int kFeatMaxThreads = 512;

struct FeatParams {
  const float* xyz;
  int64_t n_frames;
  int n_atoms;
  const int32_t* units;
  int n_units;
  int n_cols;
  float* out;
  int64_t ld_out;
  int ft;            // frames per tile (multiple of 4)
  int bulk_in_ok;    // xyz base 16B aligned
  int bulk_out_ok;   // out base 16B aligned and ld_out == n_cols
};

// MUFU-based square roots: max relative error 2^-22 (sqrt.approx) resp. 2 ulp (rsqrt.approx), against test
// tolerances of 2e-6 (distances) and 1e-4 rad (angles); the IEEE sequences cost ~10 instructions each and
// the kernel is issue-bound.
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct FeatUnit { int32_t v[8]; };   // kind, 4 atoms, angle / cos / sin (or distance) output columns

__device__ __forceinline__ void feat_unit(const float* __restrict__ fr, const FeatUnit uu,
                                          float* __restrict__ orow) {
  const int32_t* u = uu.v;
  const int kind = u[0];
  const float* p0 = fr + 3 * u[1];
  const float* p1 = fr + 3 * u[2];
  if (kind == 1) {
    float dx = p1[0] - p0[0], dy = p1[1] - p0[1], dz = p1[2] - p0[2];
    float d = fast_sqrt(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    if (u[5] >= 0) orow[u[5]] = d;
    return;
  }
  const float* p2 = fr + 3 * u[3];
  const float* p3 = fr + 3 * u[4];
  float b1x = p1[0] - p0[0], b1y = p1[1] - p0[1], b1z = p1[2] - p0[2];
  float b2x = p2[0] - p1[0], b2y = p2[1] - p1[1], b2z = p2[2] - p1[2];
  float b3x = p3[0] - p2[0], b3y = p3[1] - p2[1], b3z = p3[2] - p2[2];
  float c1x = b2y * b3z - b2z * b3y, c1y = b2z * b3x - b2x * b3z, c1z = b2x * b3y - b2y * b3x;
  float c2x = b1y * b2z - b1z * b2y, c2y = b1z * b2x - b1x * b2z, c2z = b1x * b2y - b1y * b2x;
  float nb2 = fast_sqrt(fmaf(b2x, b2x, fmaf(b2y, b2y, b2z * b2z)));
  float y = (b1x * c1x + b1y * c1y + b1z * c1z) * nb2;
  float x = c1x * c2x + c1y * c2y + c1z * c2z;
  if (u[5] >= 0) {
    float a = atan2f(y, x);
    // (-pi, pi] like pmarlo's _wrap_to_minus_pi_pi (features/builtins.py:11-14)
    if (a <= -3.14159265358979323846f) a += 6.28318530717958647692f;
    orow[u[5]] = a;
  }
  if (u[6] >= 0 || u[7] >= 0) {
    float h2 = fmaf(x, x, y * y);
    float c = 1.0f, s = 0.0f;
    if (h2 > 0.0f) {
      float rh = fast_rsqrt(h2);
      c = x * rh;
      s = y * rh;
    }
    if (u[6] >= 0) orow[u[6]] = c;
    if (u[7] >= 0) orow[u[7]] = s;
  }
}

__global__ void __launch_bounds__(kFeatMaxThreads) featurize_kernel(FeatParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int A3 = p.n_atoms * 3;
  const uint32_t stage_bytes = (uint32_t)p.ft * A3 * 4u;
  const uint32_t stage_stride = (stage_bytes + 127u) & ~127u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                 // 2 mbarriers
  float* in0 = reinterpret_cast<float*>(smem + 128);
  float* in1 = reinterpret_cast<float*>(smem + 128 + stage_stride);
  float* otile = reinterpret_cast<float*>(smem + 128 + 2 * stage_stride);
  const size_t otile_bytes = ((size_t)p.ft * p.n_cols * 4 + 127) & ~(size_t)127;
  int32_t* s_units = reinterpret_cast<int32_t*>(smem + 128 + 2 * stage_stride + otile_bytes);

  int* s_order = s_units + p.n_units * 8;     // unit ids: dihedrals first, then distances (stable)
  __shared__ int s_ndih;
  const int nthreads = blockDim.x;
  for (int i = tid; i < p.n_units * 8; i += nthreads) s_units[i] = p.units[i];
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    int nd = 0;
    for (int u = 0; u < p.n_units; ++u)
      if (s_units[u * 8] != 1) s_order[nd++] = u;
    s_ndih = nd;
    for (int u = 0; u < p.n_units; ++u)
      if (s_units[u * 8] == 1) s_order[nd++] = u;
  }
  __syncthreads();
  const int n_dih = s_ndih;

  const int64_t n_tiles = (p.n_frames + p.ft - 1) / p.ft;
  auto tile_frames = [&](int64_t t) -> int {
    int64_t rem = p.n_frames - t * p.ft;
    return rem < p.ft ? (int)rem : p.ft;
  };
  auto tile_bulk = [&](int64_t t) -> bool { return p.bulk_in_ok && tile_frames(t) == p.ft; };

  int stage = 0;
  uint32_t phase0 = 0, phase1 = 0;
  int64_t tile = blockIdx.x;
  if (tile < n_tiles && tile_bulk(tile) && tid == 0) {
    mbar_expect_tx(&bars[0], stage_bytes);
    bulk_g2s(in0, p.xyz + tile * p.ft * A3, stage_bytes, &bars[0]);
  }
  for (; tile < n_tiles; tile += gridDim.x) {
    const int64_t next = tile + gridDim.x;
    float* in_cur = stage ? in1 : in0;
    if (next < n_tiles && tile_bulk(next) && tid == 0) {
      uint64_t* b = &bars[stage ^ 1];
      mbar_expect_tx(b, stage_bytes);
      bulk_g2s(stage ? in0 : in1, p.xyz + next * p.ft * A3, stage_bytes, b);
    }
    const int nf = tile_frames(tile);
    const int64_t f0 = tile * p.ft;
    if (tile_bulk(tile)) {
      if (stage == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1; }
      else            { mbar_wait(&bars[1], phase1); phase1 ^= 1; }
    } else {
      const float* src = p.xyz + f0 * A3;
      for (int i = tid; i < nf * A3; i += nthreads) in_cur[i] = ldg_stream_f(src + i);
      __syncthreads();
    }
    // the previous tile's bulk store must have finished reading otile
    if (tid == 0) bulk_wait_read<0>();
    __syncthreads();

    // Two phases, dihedrals then distances (a dihedral costs ~3x a distance: mixing the kinds in one pass
    // leaves the distance warps idle while the dihedral warps finish).  In each phase a thread owns ONE unit
    // of that kind (decoded once into registers) and walks the tile's frames with the phase's stride.
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
      const int cnt = phase == 0 ? n_dih : p.n_units - n_dih;
      const int* list = s_order + (phase == 0 ? 0 : n_dih);
      if (cnt == 0) continue;
      if (cnt <= nthreads) {
        const int fl_n = nthreads / cnt;
        const int fl = tid / cnt;
        if (fl < fl_n) {
          const int uid = list[tid - fl * cnt];
          FeatUnit u;
#pragma unroll
          for (int q = 0; q < 8; ++q) u.v[q] = s_units[uid * 8 + q];
          for (int f = fl; f < nf; f += fl_n) feat_unit(in_cur + f * A3, u, otile + f * p.n_cols);
        }
      } else {
        const int work = nf * cnt;
        for (int w = tid; w < work; w += nthreads) {
          const int f = w / cnt;
          const int uid = list[w - f * cnt];
          FeatUnit uu;
#pragma unroll
          for (int q = 0; q < 8; ++q) uu.v[q] = s_units[uid * 8 + q];
          feat_unit(in_cur + f * A3, uu, otile + f * p.n_cols);
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    const uint32_t obytes = (uint32_t)nf * p.n_cols * 4u;
    if (p.bulk_out_ok && (obytes & 15u) == 0) {
      if (tid == 0) {
        bulk_s2g(p.out + f0 * p.ld_out, otile, obytes);
        bulk_commit();
      }
    } else {
      for (int i = tid; i < nf * p.n_cols; i += nthreads) {
        const int f = i / p.n_cols;
        const int c = i - f * p.n_cols;
        p.out[(f0 + f) * p.ld_out + c] = otile[i];
      }
      __syncthreads();
    }
    stage ^= 1;
  }
  if (tid == 0) bulk_wait_all<0>();
}

__global__ void trig_expand_kernel(const double* __restrict__ X, int64_t n, int F,
                                   const uint8_t* __restrict__ periodic,
                                   const int32_t* __restrict__ out_col, double* __restrict__ Xe, int Fe) {
  const int64_t total = n * F;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / F;
    const int j = (int)(e - r * F);
    const double v = X[e];
    double* o = Xe + r * Fe + out_col[j];
    if (periodic[j]) {
      double s, c;
      sincos(v, &s, &c);
      o[0] = c;
      o[1] = s;
    } else {
      o[0] = v;
    }
  }
}

}  // namespace pmb

extern "C" int pmb_trig_expand(const double* X, int64_t n, int F, const uint8_t* periodic,
                               const int32_t* out_col, double* Xe, int Fe, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n >= 0 && F >= 0 && Fe >= F, "pmb_trig_expand: bad sizes");
  if (n == 0 || F == 0) return PMB_OK;
  PMB_REQUIRE(X && periodic && out_col && Xe, "pmb_trig_expand: null pointer");
  const int64_t total = n * F;
  const int64_t blocks = (total + 255) / 256;
  const int grid = (int)(blocks < 16 * kNumSMs ? blocks : 16 * kNumSMs);
  trig_expand_kernel<<<grid, 256, 0, as_stream(stream)>>>(X, n, F, periodic, out_col, Xe, Fe);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

extern "C" int pmb_featurize(const float* xyz, int64_t n_frames, int n_atoms, const int32_t* units,
                             int n_units, int n_cols, float* out, int64_t ld_out,
                             pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n_frames >= 0 && n_atoms > 0 && n_units >= 0 && n_cols >= 0, "pmb_featurize: bad sizes");
  PMB_REQUIRE(ld_out >= n_cols, "pmb_featurize: ld_out < n_cols");
  if (n_frames == 0 || n_cols == 0 || n_units == 0) return PMB_OK;
  PMB_REQUIRE(xyz && units && out, "pmb_featurize: null pointer");

  const size_t frame_in = (size_t)n_atoms * 12, frame_out = (size_t)n_cols * 4;
  const size_t fixed = 128 + 3 * 128 + (size_t)n_units * 36;
  const size_t budget = 100 * 1024;  // 2 CTAs per SM
  int ft = (int)((budget - fixed) / (2 * frame_in + frame_out));
  ft = (ft / 4) * 4;
  if (ft > 64) ft = 64;
  PMB_REQUIRE(ft >= 4, "pmb_featurize: frame too large for shared-memory staging (%d atoms, %d cols)",
              n_atoms, n_cols);
  if ((int64_t)ft > ((n_frames + 3) / 4) * 4) ft = (int)(((n_frames + 3) / 4) * 4);

  FeatParams p;
  p.xyz = xyz; p.n_frames = n_frames; p.n_atoms = n_atoms; p.units = units; p.n_units = n_units;
  p.n_cols = n_cols; p.out = out; p.ld_out = ld_out; p.ft = ft;
  p.bulk_in_ok = ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0);
  p.bulk_out_ok = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && ld_out == n_cols;

  const size_t stage_stride = (((size_t)ft * frame_in) + 127) & ~(size_t)127;
  const size_t otile = (((size_t)ft * frame_out) + 127) & ~(size_t)127;
  const size_t smem = 128 + 2 * stage_stride + otile + (size_t)n_units * 36;
  PMB_CUDA(cudaFuncSetAttribute(featurize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (n_frames + ft - 1) / ft;
  int grid = (int)(n_tiles < 2 * kNumSMs ? n_tiles : 2 * kNumSMs);
  // block size: the largest multiple of n_units that fits 384 threads (all threads own a unit), rounded up
  // to a whole warp; unit lists longer than that use the generic item loop
  int threads = kFeatThreads;
  if (n_units <= 384) {
    threads = (384 / n_units) * n_units;
    threads = ((threads + 31) / 32) * 32;
    if (threads < 128) threads = 128;
  }
  featurize_kernel<<<grid, threads, smem, as_stream(stream)>>>(p);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

// Coarse pass of render_rays, fused: sample placement -> VM lookup -> PE -> CRR field -> compositing.
// Replaces networks/renderer.py:157-188 + networks/pdrf/voxnerf.py:203-259,153-201 for the coarse field.
//
// fp32 SIMT formulation: one thread per (ray, sample); every MLP layer is an outer-product accumulation
// acc[out] += W^T[k][out] * x_k with the transposed weights broadcast from shared memory (LDS.128), so the input
// vector is never materialised and the per-thread state is the output accumulators only.
#include "common.cuh"
#include "coarse_args.cuh"

namespace edn {

constexpr int kCoarseThreads = 256;
constexpr int kCH = 64;    // coarse hidden
constexpr int kCGeo = 15;  // coarse geo feat

struct CoarseSmemLayout {
  // float offsets
  static constexpr int basis = 0;                    // [96][32]
  static constexpr int s0 = basis + 96 * 32;         // [96][64]
  static constexpr int s1 = s0 + 96 * 64;            // [64][16]
  static constexpr int c0 = s1 + 64 * 16;            // [42][64]
  static constexpr int c1 = c0 + 42 * 64;            // [64][64]
  static constexpr int c2 = c1 + 64 * 64;            // [64][4]
  static constexpr int b0 = c2 + 64 * 4;             // [64]
  static constexpr int b1 = b0 + 64;                 // [64]
  static constexpr int b2 = b1 + 64;                 // [4]
  static constexpr int weights_end = b2 + 4;
  // per-block sample scratch (kCoarseThreads entries each)
  static constexpr int sig = weights_end;            // [T]
  static constexpr int rgb = sig + kCoarseThreads;   // [T][3]
  static constexpr int z = rgb + 3 * kCoarseThreads; // [T]
  static constexpr int w = z + kCoarseThreads;       // [T]
  static constexpr int total = w + kCoarseThreads;
};


template <int N>
__device__ __forceinline__ void axpy_row(float (&acc)[N], const float* __restrict__ w_row, float x) {
  const float4* r4 = reinterpret_cast<const float4*>(w_row);
#pragma unroll
  for (int j = 0; j < N / 4; ++j) {
    const float4 w = r4[j];
    acc[4 * j + 0] = fmaf(w.x, x, acc[4 * j + 0]); acc[4 * j + 1] = fmaf(w.y, x, acc[4 * j + 1]);
    acc[4 * j + 2] = fmaf(w.z, x, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(w.w, x, acc[4 * j + 3]);
  }
}

__device__ __forceinline__ void copy_to_smem(float* dst, const float* __restrict__ src, int n, float fill_if_null) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src ? __ldg(src + i) : fill_if_null;
}

template <typename T>
__global__ void __launch_bounds__(kCoarseThreads, 1) coarse_fwd_kernel(const CoarseArgs a) {
  extern __shared__ __align__(16) float smem[];
  using L = CoarseSmemLayout;
  copy_to_smem(smem + L::basis, a.grid.basis_t, 96 * 32, 0.f);
  copy_to_smem(smem + L::s0, a.mlp.sigma0_t, 96 * 64, 0.f);
  copy_to_smem(smem + L::s1, a.mlp.sigma1_t, 64 * 16, 0.f);
  copy_to_smem(smem + L::c0, a.mlp.color0_t, 42 * 64, 0.f);
  copy_to_smem(smem + L::c1, a.mlp.color1_t, 64 * 64, 0.f);
  copy_to_smem(smem + L::c2, a.mlp.color2_t, 64 * 4, 0.f);
  copy_to_smem(smem + L::b0, a.mlp.color0_b, 64, 0.f);
  copy_to_smem(smem + L::b1, a.mlp.color1_b, 64, 0.f);
  copy_to_smem(smem + L::b2, a.mlp.color2_b, 4, 0.f);
  __syncthreads();

  const int S = a.n_samples;
  const int rpb = kCoarseThreads / S;          // rays per block iteration (S <= kCoarseThreads checked on the host)
  const int active = rpb * S;
  const int tid = threadIdx.x;
  const int lr = tid / S, s = tid - lr * S;    // local ray, sample
  const int64_t n_groups = (a.n_rays + rpb - 1) / rpb;

  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int64_t ray = grp * rpb + lr;
    const bool live = (tid < active) && (ray < a.n_rays);
    float zval = 0.f, sig_raw = 0.f, col[3] = {0.f, 0.f, 0.f};
    if (live) {
      const float* rb = a.ray_batch + ray * 11;
      const float o[3] = {__ldg(rb + 0), __ldg(rb + 1), __ldg(rb + 2)};
      const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
      const float near = __ldg(rb + 6), far = __ldg(rb + 7);
      const float vd[3] = {__ldg(rb + 8), __ldg(rb + 9), __ldg(rb + 10)};
      zval = place_sample(a, ray, s, near, far);
      float p[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], zval));   // renderer.py:180

      // ---- sigma_net.0 : [ft(32) | PE(pts)(63)] -> 64, accumulated input by input ------------------------------
      float h[kCH];
#pragma unroll
      for (int j = 0; j < kCH; ++j) h[j] = 0.f;
      {
        float ft[kAppDim];
        vm_sample_point<T>(a.grid, smem + L::basis, p, ft);
#pragma unroll
        for (int k = 0; k < kAppDim; ++k) axpy_row<kCH>(h, smem + L::s0 + k * kCH, ft[k]);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) axpy_row<kCH>(h, smem + L::s0 + (32 + i) * kCH, p[i]);
#pragma unroll 1
      for (int f = 0; f < kPeFreqPts; ++f) {
        const float fr = (float)(1 << f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float sn, cs;
          sincosf(p[i] * fr, &sn, &cs);
          axpy_row<kCH>(h, smem + L::s0 + (35 + 6 * f + i) * kCH, sn);
          axpy_row<kCH>(h, smem + L::s0 + (38 + 6 * f + i) * kCH, cs);
        }
      }
      // ---- sigma_net.1 : relu(h) -> [sigma | geo(15)] -----------------------------------------------------------
      float og[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) og[j] = 0.f;
#pragma unroll
      for (int k = 0; k < kCH; ++k) axpy_row<16>(og, smem + L::s1 + k * 16, fmaxf(h[k], 0.f));
      sig_raw = og[0];
      if (a.feat) {
        float* fo = a.feat + (ray * S + s) * kCGeo;
#pragma unroll
        for (int j = 0; j < kCGeo; ++j) fo[j] = og[1 + j];
      }
      // ---- color_net : [geo(15) | PE(dir)(27)] -> 64 -> 64 -> 3, sigmoid -----------------------------------------
      float c[kCH];
#pragma unroll
      for (int j = 0; j < kCH; ++j) c[j] = smem[L::b0 + j];
#pragma unroll
      for (int k = 0; k < kCGeo; ++k) axpy_row<kCH>(c, smem + L::c0 + k * kCH, og[1 + k]);
#pragma unroll
      for (int i = 0; i < 3; ++i) axpy_row<kCH>(c, smem + L::c0 + (15 + i) * kCH, vd[i]);
#pragma unroll 1
      for (int f = 0; f < kPeFreqDir; ++f) {
        const float fr = (float)(1 << f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float sn, cs;
          sincosf(vd[i] * fr, &sn, &cs);
          axpy_row<kCH>(c, smem + L::c0 + (18 + 6 * f + i) * kCH, sn);
          axpy_row<kCH>(c, smem + L::c0 + (21 + 6 * f + i) * kCH, cs);
        }
      }
#pragma unroll
      for (int j = 0; j < kCH; ++j) c[j] = fmaxf(c[j], 0.f);
      float out4[4] = {smem[L::b2 + 0], smem[L::b2 + 1], smem[L::b2 + 2], 0.f};
#pragma unroll
      for (int half = 0; half < 2; ++half) {     // color_net.1 in two 32-wide halves to bound live registers
        float c2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) c2[j] = smem[L::b1 + half * 32 + j];
#pragma unroll
        for (int k = 0; k < kCH; ++k) axpy_row<32>(c2, smem + L::c1 + k * kCH + half * 32, c[k]);
#pragma unroll
        for (int j = 0; j < 32; ++j) axpy_row<4>(out4, smem + L::c2 + (half * 32 + j) * 4, fmaxf(c2[j], 0.f));
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) col[i] = sigmoidf_(out4[i]);
    }
    smem[L::sig + tid] = sig_raw;
    smem[L::rgb + 3 * tid + 0] = col[0]; smem[L::rgb + 3 * tid + 1] = col[1]; smem[L::rgb + 3 * tid + 2] = col[2];
    smem[L::z + tid] = zval;
    __syncthreads();
    if (tid < rpb) {
      const int64_t r2 = grp * rpb + tid;
      if (r2 < a.n_rays) {
        const float* rb = a.ray_batch + r2 * 11;
        const float dx = __ldg(rb + 3), dy = __ldg(rb + 4), dz = __ldg(rb + 5);
        const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        const bool mask_near = !(a.flags & EDN_FLAG_TRAIN) && a.rmnearplane > 0.f;
        float out[5];
        composite_ray(smem + L::sig + tid * S, smem + L::rgb + 3 * tid * S, smem + L::z + tid * S,
                      a.noise ? a.noise + r2 * (S - 1) : nullptr, S, dnorm, mask_near, a.rmnearplane / 128.0f,
                      (a.flags & EDN_FLAG_RELU_RGB) != 0, smem + L::w + tid * S, out);
        a.rgb[r2 * 3 + 0] = out[0]; a.rgb[r2 * 3 + 1] = out[1]; a.rgb[r2 * 3 + 2] = out[2];
        a.depth[r2] = out[3];
        a.acc[r2] = out[4];
      }
    }
    __syncthreads();
    if (live) {
      a.z_vals[ray * S + s] = zval;
      a.weights[ray * S + s] = smem[L::w + tid];
    }
    __syncthreads();
  }
}

}  // namespace edn

extern "C" int edn_render_coarse_fwd(const edn_vm_grid* grid, const edn_field_mlp* mlp, const float* ray_batch,
                                     const float* t_vals, const float* t_rand, const float* noise, int64_t n_rays,
                                     int32_t n_samples, int32_t flags, float rmnearplane, int32_t precision,
                                     float* z_vals, float* weights, float* rgb, float* depth, float* acc, float* feat,
                                     void* stream) {
  using namespace edn;
  EDN_REQUIRE(mlp && ray_batch && t_vals && z_vals && weights && rgb && depth && acc, "edn_render_coarse_fwd: null pointer");
  EDN_REQUIRE(n_rays >= 0, "edn_render_coarse_fwd: n_rays < 0");
  EDN_REQUIRE(n_samples >= 2 && n_samples <= kCoarseThreads, "edn_render_coarse_fwd: n_samples must be in [2,%d], got %d",
              kCoarseThreads, n_samples);
  EDN_REQUIRE(mlp->hidden == kCH && mlp->geo_feat == kCGeo, "edn_render_coarse_fwd: coarse field must be hidden=64, geo_feat=15");
  EDN_REQUIRE(mlp->sigma0_t && mlp->sigma1_t && mlp->color0_t && mlp->color1_t && mlp->color2_t, "edn_render_coarse_fwd: null weight");
  CoarseArgs a;
  int rc = make_grid_dev(grid, &a.grid);
  if (rc) return rc;
  if (n_rays == 0) return EDN_OK;
  a.mlp = *mlp;
  a.ray_batch = ray_batch; a.t_vals = t_vals; a.t_rand = t_rand; a.noise = noise;
  a.n_rays = n_rays; a.n_samples = n_samples; a.flags = flags; a.rmnearplane = rmnearplane;
  a.z_vals = z_vals; a.weights = weights; a.rgb = rgb; a.depth = depth; a.acc = acc; a.feat = feat;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (precision == EDN_BF16) return launch_coarse_tc(a, grid->dtype, st);
  if (precision == EDN_TC32) return launch_coarse_tc3(a, grid->dtype, st);
  EDN_REQUIRE(precision == EDN_F32, "edn_render_coarse_fwd: bad precision %d", precision);
  const int rpb = kCoarseThreads / n_samples;
  const int64_t n_groups = (n_rays + rpb - 1) / rpb;
  const int grid_x = (int)(n_groups < (int64_t)num_sms() ? n_groups : (int64_t)num_sms());
  const size_t smem = CoarseSmemLayout::total * sizeof(float);
  if (grid->dtype == EDN_F32) {
    EDN_CUDA_OK(cudaFuncSetAttribute(coarse_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    coarse_fwd_kernel<float><<<grid_x, kCoarseThreads, smem, st>>>(a);
  } else if (grid->dtype == EDN_BF16) {
    EDN_CUDA_OK(cudaFuncSetAttribute(coarse_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    coarse_fwd_kernel<__nv_bfloat16><<<grid_x, kCoarseThreads, smem, st>>>(a);
  } else {
    set_error("edn_render_coarse_fwd: bad grid dtype %d", grid->dtype);
    return EDN_E_INVALID;
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

namespace edn {
__global__ void place_samples_kernel(const CoarseArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_rays * a.n_samples) return;
  const int64_t ray = i / a.n_samples;
  const int s = (int)(i - ray * a.n_samples);
  a.z_vals[i] = place_sample(a, ray, s, __ldg(a.ray_batch + ray * 11 + 6), __ldg(a.ray_batch + ray * 11 + 7));
}
}  // namespace edn

extern "C" int edn_place_samples(const float* ray_batch, const float* t_vals, const float* t_rand, int64_t n_rays, int32_t n_samples,
                                 int32_t flags, float* z_vals, void* stream) {
  using namespace edn;
  EDN_REQUIRE(ray_batch && t_vals && z_vals && n_samples >= 1, "edn_place_samples: bad argument");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  CoarseArgs a{};
  a.ray_batch = ray_batch; a.t_vals = t_vals; a.t_rand = t_rand; a.n_rays = n_rays; a.n_samples = n_samples; a.flags = flags;
  a.z_vals = z_vals;
  const int64_t n = n_rays * n_samples;
  place_samples_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

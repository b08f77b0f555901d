// mode = nerf: the vanilla NeRF field ("run_network"), fp32 SIMT.
// Replaces NeRF.mlpforward + NeRF.eval (networks/nerf.py:46-72, 131-162): PE(pts) -> 8 x 256 ReLU MLP with the
// skip-concat of the 63-d input after layer 4 -> alpha_linear / feature_linear -> views_linears (283 -> 128) -> rgb_linear,
// and NeRF.raw2outputs (networks/nerf.py:74-129): sigmoid rgb, sigma in channel 3, compositing, optional white background.
// One CTA per ray, 64-sample row tiles, activations in shared memory, weights streamed from L2 (simt_gemm.cuh).
#include "common.cuh"
#include "simt_gemm.cuh"

namespace edn {
namespace {

constexpr int kPeLd = 68;   // 64 + 4: PE tile row stride

struct NerfSmem {
  static constexpr int A = 0;                               // [64][260]
  static constexpr int Ws = A + kTileM * kLda;              // [2][16][256]
  static constexpr int PE = Ws + 2 * kKc * kFH;             // [64][68]
  static constexpr int bias_ray = PE + kTileM * kPeLd;      // [128]
  static constexpr int total = bias_ray + 128;
};

struct NerfArgs {
  edn_nerf_mlp mlp;
  const float* ray_batch;
  const float* z_vals;
  int64_t n_rays;
  int S;
  float* raw;        // [R][S][4] = rgb(3), sigma
  float* feature;    // [R][S][256] or NULL
  int feature_after_linear;
};

__global__ void __launch_bounds__(kFineThreads, 1) nerf_mlp_f32_kernel(const NerfArgs a) {
  extern __shared__ __align__(16) float smem[];
  using L = NerfSmem;
  const int tid = threadIdx.x;
  float* A = smem + L::A;
  float* Ws = smem + L::Ws;
  float* PEs = smem + L::PE;
  const int S = a.S, n_tiles = (S + kTileM - 1) / kTileM;
  for (int64_t ray = blockIdx.x; ray < a.n_rays; ray += gridDim.x) {
    const float* rb = a.ray_batch + ray * 11;
    const float o[3] = {__ldg(rb + 0), __ldg(rb + 1), __ldg(rb + 2)};
    const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
    if (tid < 128) {   // per-ray view-direction part of views_linears.0: b + W[:, 256:283] . PE(viewdir)
      const float vd[3] = {__ldg(rb + 8), __ldg(rb + 9), __ldg(rb + 10)};
      float b = __ldg(a.mlp.views_b + tid);
      const float* w = a.mlp.views_t + (size_t)256 * 128 + tid;
#pragma unroll
      for (int i = 0; i < 3; ++i) b = fmaf(__ldg(w + i * 128), vd[i], b);
      for (int f = 0; f < kPeFreqDir; ++f) {
        const float fr = (float)(1 << f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float sn, cs;
          sincosf(vd[i] * fr, &sn, &cs);
          b = fmaf(__ldg(w + (3 + 6 * f + i) * 128), sn, b);
          b = fmaf(__ldg(w + (6 + 6 * f + i) * 128), cs, b);
        }
      }
      smem[L::bias_ray + tid] = b;
    }
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int row0 = tile * kTileM, valid = min(kTileM, S - row0);
      __syncthreads();
      {  // PE(pts) of the tile -> PEs[:, 0:63] (col 63 = 0) and A[:, 0:64]; 4 threads per row, 16 columns each
        const int r = tid >> 2, part = tid & 3;
        const float zv = (r < valid) ? a.z_vals[ray * S + row0 + r] : 0.f;
        float p[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], zv));
        for (int c = part * 16; c < part * 16 + 16; ++c) {
          float v = 0.f;
          if (r < valid && c < 63) {
            if (c < 3) v = p[c];
            else {
              const int f = (c - 3) / 6, k = (c - 3) % 6;
              const float ang = p[k % 3] * (float)(1 << f);
              v = (k < 3) ? sinf(ang) : cosf(ang);
            }
          }
          PEs[r * kPeLd + c] = v;
          A[r * kLda + c] = v;
        }
      }
      __syncthreads();
      float acc8[8][8];
      for (int layer = 0; layer < 8; ++layer) {     // pts_linears (nerf.py:135-139)
        if (layer == 0) gemm_tile<8>(A, a.mlp.pts_t[0], 64, acc8, Ws);
        else if (layer == 5) {                       // skip: input = [PE(63) | h(256)]
          gemm_tile<8>(PEs, a.mlp.pts_t[5], 64, acc8, Ws, false, kPeLd);
          gemm_tile<8>(A, a.mlp.pts_t[5] + (size_t)64 * 256, 256, acc8, Ws, true);
        } else gemm_tile<8>(A, a.mlp.pts_t[layer], 256, acc8, Ws);
        store_tile<8, true>(A, acc8, nullptr, a.mlp.pts_b[layer], nullptr, valid);
        __syncthreads();
      }
      if (a.feature && !a.feature_after_linear) {   // extract_feature == "before_linear": h itself
        for (int i = tid; i < valid * 64; i += kFineThreads) {
          const int r = i >> 6, c4 = (i & 63) * 4;
          *reinterpret_cast<float4*>(a.feature + ((size_t)ray * S + row0 + r) * 256 + c4) = *reinterpret_cast<const float4*>(A + r * kLda + c4);
        }
      }
      {  // alpha_linear: 256 -> 1
        float sg[1];
        dot_rows<1>(A, a.mlp.alpha_w, 1, 256, sg);
        if ((tid & 3) == 0 && (tid >> 2) < valid) a.raw[((size_t)ray * S + row0 + (tid >> 2)) * 4 + 3] = sg[0] + __ldg(a.mlp.alpha_b);
      }
      // feature_linear: 256 -> 256, no activation
      gemm_tile<8>(A, a.mlp.feature_t, 256, acc8, Ws);
      store_tile<8, false>(A, acc8, nullptr, a.mlp.feature_b,
                           (a.feature && a.feature_after_linear) ? a.feature + ((size_t)ray * S + row0) * 256 : nullptr, valid);
      __syncthreads();
      {  // views_linears.0: [feature(256) | PE(dir)(27)] -> 128, ReLU (dir part + bias folded into bias_ray)
        float acc4[8][4];
        gemm_tile<4>(A, a.mlp.views_t, 256, acc4, Ws);
        store_tile<4, true>(A, acc4, smem + L::bias_ray, nullptr, nullptr, valid);
      }
      __syncthreads();
      {  // rgb_linear: 128 -> 3
        float c3[3];
        dot_rows<3>(A, a.mlp.rgb_t, 4, 128, c3);
        if ((tid & 3) == 0 && (tid >> 2) < valid) {
          float* o4 = a.raw + ((size_t)ray * S + row0 + (tid >> 2)) * 4;
#pragma unroll
          for (int i = 0; i < 3; ++i) o4[i] = c3[i] + (a.mlp.rgb_b ? __ldg(a.mlp.rgb_b + i) : 0.f);
        }
      }
    }
    __syncthreads();
  }
}

// NeRF.raw2outputs (nerf.py:74-129): one thread per ray, samples in order.
__global__ void nerf_raw2outputs_kernel(const float* __restrict__ raw, const float* __restrict__ z_vals, const float* __restrict__ ray_batch,
                                        const float* __restrict__ noise, int64_t R, int S, int flags, float rmnearplane,
                                        float* __restrict__ weights, float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ acc) {
  const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= R) return;
  const float* rb = ray_batch + ray * 11;
  const float dx = rb[3], dy = rb[4], dz = rb[5];
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const bool mask_near = !(flags & EDN_FLAG_TRAIN) && rmnearplane > 0.f;
  const float near_thr = rmnearplane / 128.0f;
  float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, ac = 0.f;
  const float* rw = raw + ray * S * 4;
  const float* z = z_vals + ray * S;
  for (int s = 0; s < S; ++s) {
    const float nz = (noise && s < S - 1) ? noise[ray * (S - 1) + s] : 0.f;
    const float alpha = alpha_of_sample(rw[4 * s + 3], z[s], s < S - 1 ? z[s + 1] : z[s], nz, dnorm, mask_near, near_thr, s == S - 1);
    const float w = alpha * T;
    weights[ray * S + s] = w;
    cr = fmaf(w, sigmoidf_(rw[4 * s + 0]), cr); cg = fmaf(w, sigmoidf_(rw[4 * s + 1]), cg); cb = fmaf(w, sigmoidf_(rw[4 * s + 2]), cb);
    dep = fmaf(w, z[s], dep);
    ac += w;
    T = T * (1.0f - alpha);
  }
  if (flags & EDN_FLAG_WHITE_BKGD) { cr += 1.f - ac; cg += 1.f - ac; cb += 1.f - ac; }   // nerf.py:126-127
  rgb[ray * 3 + 0] = cr; rgb[ray * 3 + 1] = cg; rgb[ray * 3 + 2] = cb;
  depth[ray] = dep;
  acc[ray] = ac;
}

}  // namespace
}  // namespace edn

extern "C" int edn_nerf_mlp_fwd(const edn_nerf_mlp* mlp, const float* ray_batch, const float* z_vals, int64_t n_rays, int32_t n_samples,
                                int32_t feature_after_linear, float* raw, float* feature, void* stream) {
  using namespace edn;
  EDN_REQUIRE(mlp && ray_batch && z_vals && raw, "edn_nerf_mlp_fwd: null pointer");
  for (int i = 0; i < 8; ++i) EDN_REQUIRE(mlp->pts_t[i] && mlp->pts_b[i], "edn_nerf_mlp_fwd: null pts_linears.%d", i);
  EDN_REQUIRE(mlp->alpha_w && mlp->alpha_b && mlp->feature_t && mlp->feature_b && mlp->views_t && mlp->views_b && mlp->rgb_t,
              "edn_nerf_mlp_fwd: null head weight");
  EDN_REQUIRE(n_samples >= 1, "edn_nerf_mlp_fwd: n_samples < 1");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  NerfArgs a{*mlp, ray_batch, z_vals, n_rays, n_samples, raw, feature, feature_after_linear};
  const size_t smem = NerfSmem::total * sizeof(float);
  EDN_CUDA_OK(cudaFuncSetAttribute(nerf_mlp_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t gx = n_rays < (int64_t)num_sms() ? n_rays : (int64_t)num_sms();
  nerf_mlp_f32_kernel<<<(unsigned)gx, kFineThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_nerf_raw2outputs(const float* raw, const float* z_vals, const float* ray_batch, const float* noise, int64_t n_rays,
                                    int32_t n_samples, int32_t flags, float rmnearplane, float* weights, float* rgb, float* depth,
                                    float* acc, void* stream) {
  using namespace edn;
  EDN_REQUIRE(raw && z_vals && ray_batch && weights && rgb && depth && acc && n_samples >= 1, "edn_nerf_raw2outputs: bad argument");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  nerf_raw2outputs_kernel<<<(unsigned)((n_rays + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      raw, z_vals, ray_batch, noise, n_rays, n_samples, flags, rmnearplane, weights, rgb, depth, acc);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

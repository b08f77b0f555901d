// Fine pass of render_rays, fp32 SIMT parity path (the tcgen05 path is fine_tc.cu).
// Replaces networks/renderer.py:190-217 + networks/pdrf/voxnerf.py:203-259,153-201 for the FVR field:
// VM lookup of the coarse AND fine grids at the merged samples, PE, sigma_net 127->256->129,
// color_net 155->256->256->3, sigmoid, compositing.  One CTA renders one ray at a time in 64-sample row tiles;
// activations stay in shared memory, weights stream from L2 in 16-row K chunks (cp.async double buffer).
#include "common.cuh"
#include "fine_args.cuh"
#include "simt_gemm.cuh"

namespace edn {

constexpr int kFGeo = 128;
constexpr int kMaxS = 512;         // max merged samples per ray

struct FineSmem {
  static constexpr int A = 0;                               // [64][260]
  static constexpr int Ws = A + kTileM * kLda;              // [2][16][256]
  static constexpr int basis_c = Ws + 2 * kKc * kFH;        // [96][32]
  static constexpr int basis_f = basis_c + 96 * 32;         // [96][32]
  static constexpr int bias_ray = basis_f + 96 * 32;        // [256]
  static constexpr int z = bias_ray + kFH;                  // [kMaxS]
  static constexpr int sig = z + kMaxS;                     // [kMaxS]
  static constexpr int w = sig + kMaxS;                     // [kMaxS]
  static constexpr int rgb = w + kMaxS;                     // [kMaxS][3]
  static constexpr int total = rgb + 3 * kMaxS;
};


template <typename T>
__global__ void __launch_bounds__(kFineThreads, 1) fine_fwd_f32_kernel(const FineArgs a) {
  extern __shared__ __align__(16) float smem[];
  using L = FineSmem;
  const int tid = threadIdx.x;
  for (int i = tid; i < 96 * 32; i += kFineThreads) {
    smem[L::basis_c + i] = __ldg(a.gc.basis_t + i);
    smem[L::basis_f + i] = __ldg(a.gf.basis_t + i);
  }
  __syncthreads();
  const int S = a.S;
  const int n_tiles = (S + kTileM - 1) / kTileM;
  float* A = smem + L::A;
  float* Ws = smem + L::Ws;

  for (int64_t ray = blockIdx.x; ray < a.n_rays; ray += gridDim.x) {
    const float* rb = a.ray_batch + ray * 11;
    const float o[3] = {__ldg(rb + 0), __ldg(rb + 1), __ldg(rb + 2)};
    const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
    for (int i = tid; i < S; i += kFineThreads) smem[L::z + i] = a.z_vals[ray * S + i];
    {  // per-ray bias of color_net.0: b0 + W0[:, 128:155] . PE(viewdir)      (voxnerf.py:241-248)
      const float vd[3] = {__ldg(rb + 8), __ldg(rb + 9), __ldg(rb + 10)};
      float b = a.mlp.color0_b ? __ldg(a.mlp.color0_b + tid) : 0.f;
      const float* w = a.mlp.color0_t + (size_t)kFGeo * kFH + tid;
#pragma unroll
      for (int i = 0; i < 3; ++i) b = fmaf(__ldg(w + i * kFH), vd[i], b);
      for (int f = 0; f < kPeFreqDir; ++f) {
        const float fr = (float)(1 << f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float sn, cs;
          sincosf(vd[i] * fr, &sn, &cs);
          b = fmaf(__ldg(w + (3 + 6 * f + i) * kFH), sn, b);
          b = fmaf(__ldg(w + (6 + 6 * f + i) * kFH), cs, b);
        }
      }
      smem[L::bias_ray + tid] = b;
    }
    __syncthreads();

    for (int tile = 0; tile < n_tiles; ++tile) {
      const int row0 = tile * kTileM;
      const int valid = min(kTileM, S - row0);
      // ---- layer-0 input: [ft_coarse(32) | ft_fine(32) | PE(pts)(63) | 0] --------------------------------------
      {
        const int r = tid & 63, job = tid >> 6;       // job 0: coarse grid, 1: fine grid, 2/3: PE halves
        float* arow = A + r * kLda;
        if (r < valid) {
          const float zv = smem[L::z + row0 + r];
          float p[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], zv));
          if (job < 2) {
            float ft[kAppDim];
            if (job == 0) vm_sample_point<T>(a.gc, smem + L::basis_c, p, ft);
            else vm_sample_point<T>(a.gf, smem + L::basis_f, p, ft);
#pragma unroll
            for (int j = 0; j < kAppDim; j += 4)
              *reinterpret_cast<float4*>(arow + job * 32 + j) = make_float4(ft[j], ft[j + 1], ft[j + 2], ft[j + 3]);
          } else {
            const int f0 = (job == 2) ? 0 : 5;
            if (job == 2) { arow[64] = p[0]; arow[65] = p[1]; arow[66] = p[2]; } else { arow[127] = 0.f; }
            for (int f = f0; f < f0 + 5; ++f) {
              const float fr = (float)(1 << f);
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                float sn, cs;
                sincosf(p[i] * fr, &sn, &cs);
                arow[67 + 6 * f + i] = sn;
                arow[70 + 6 * f + i] = cs;
              }
            }
          }
        } else {
          for (int j = job * 32; j < job * 32 + 32; ++j) arow[j] = 0.f;
        }
      }
      __syncthreads();
      float acc8[8][8];
      // ---- sigma_net.0: 128 -> 256, relu -----------------------------------------------------------------------
      gemm_tile<8>(A, a.mlp.sigma0_t, 128, acc8, Ws);
      store_tile<8, true>(A, acc8, nullptr, nullptr, nullptr, valid);
      __syncthreads();
      // ---- sigma_net.1: 256 -> sigma (1) + geo (128), no activation --------------------------------------------
      {
        float sg[1];
        dot_rows<1>(A, a.mlp.sigma1_v, 1, kFH, sg);
        if ((tid & 3) == 0) smem[L::sig + row0 + (tid >> 2)] = sg[0];
      }
      {
        float acc4[8][4];
        gemm_tile<4>(A, a.mlp.sigma1_t, kFH, acc4, Ws);
        store_tile<4, false>(A, acc4, nullptr, nullptr,
                             a.feat ? a.feat + ((size_t)ray * S + row0) * kFGeo : nullptr, valid);
      }
      __syncthreads();
      // ---- color_net.0: [geo(128) | PE(dir)] -> 256, relu (dir part folded into bias_ray) ----------------------
      gemm_tile<8>(A, a.mlp.color0_t, kFGeo, acc8, Ws);
      store_tile<8, true>(A, acc8, smem + L::bias_ray, nullptr, nullptr, valid);
      __syncthreads();
      // ---- color_net.1: 256 -> 256, relu -----------------------------------------------------------------------
      gemm_tile<8>(A, a.mlp.color1_t, kFH, acc8, Ws);
      store_tile<8, true>(A, acc8, nullptr, a.mlp.color1_b, nullptr, valid);
      __syncthreads();
      // ---- color_net.2: 256 -> 3, sigmoid ----------------------------------------------------------------------
      {
        float c3[3];
        dot_rows<3>(A, a.mlp.color2_t, 4, kFH, c3);
        if ((tid & 3) == 0) {
          const int r = row0 + (tid >> 2);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float bb = a.mlp.color2_b ? __ldg(a.mlp.color2_b + i) : 0.f;
            smem[L::rgb + 3 * r + i] = sigmoidf_(c3[i] + bb);
          }
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
      const bool mask_near = !(a.flags & EDN_FLAG_TRAIN) && a.rmnearplane > 0.f;
      float out[5];
      composite_ray(smem + L::sig, smem + L::rgb, smem + L::z, a.noise ? a.noise + ray * (S - 1) : nullptr, S, dnorm,
                    mask_near, a.rmnearplane / 128.0f, (a.flags & EDN_FLAG_RELU_RGB) != 0, smem + L::w, out);
      a.rgb[ray * 3 + 0] = out[0]; a.rgb[ray * 3 + 1] = out[1]; a.rgb[ray * 3 + 2] = out[2];
      a.depth[ray] = out[3];
      a.acc[ray] = out[4];
    }
    __syncthreads();
    for (int i = tid; i < S; i += kFineThreads) a.weights[ray * S + i] = smem[L::w + i];
    __syncthreads();
  }
}

int launch_fine_f32(const FineArgs& a, int grid_dtype, cudaStream_t st) {
  const size_t smem = FineSmem::total * sizeof(float);
  int64_t gx = a.n_rays < (int64_t)num_sms() ? a.n_rays : (int64_t)num_sms();
  if (grid_dtype == EDN_F32) {
    EDN_CUDA_OK(cudaFuncSetAttribute(fine_fwd_f32_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fine_fwd_f32_kernel<float><<<(unsigned)gx, kFineThreads, smem, st>>>(a);
  } else {
    EDN_CUDA_OK(cudaFuncSetAttribute(fine_fwd_f32_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fine_fwd_f32_kernel<__nv_bfloat16><<<(unsigned)gx, kFineThreads, smem, st>>>(a);
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

int launch_fine_tc(const FineArgs& a, int grid_dtype, cudaStream_t st);   // fine_tc.cu

}  // namespace edn

extern "C" int edn_render_fine_fwd(const edn_vm_grid* grid_coarse, const edn_vm_grid* grid_fine, const edn_field_mlp* mlp,
                                   const float* ray_batch, const float* z_vals, const float* noise, int64_t n_rays,
                                   int32_t n_samples, int32_t flags, float rmnearplane, int32_t precision,
                                   float* weights, float* rgb, float* depth, float* acc, float* feat, void* stream) {
  using namespace edn;
  EDN_REQUIRE(mlp && ray_batch && z_vals && weights && rgb && depth && acc, "edn_render_fine_fwd: null pointer");
  EDN_REQUIRE(n_samples >= 2 && n_samples <= kMaxS, "edn_render_fine_fwd: n_samples must be in [2,%d], got %d", kMaxS, n_samples);
  EDN_REQUIRE(mlp->hidden == kFH && mlp->geo_feat == kFGeo, "edn_render_fine_fwd: fine field must be hidden=256, geo_feat=128");
  EDN_REQUIRE(mlp->sigma0_t && mlp->sigma1_t && mlp->sigma1_v && mlp->color0_t && mlp->color1_t && mlp->color2_t,
              "edn_render_fine_fwd: null weight");
  EDN_REQUIRE(grid_coarse && grid_fine && grid_coarse->dtype == grid_fine->dtype, "edn_render_fine_fwd: grids must share a dtype");
  EDN_REQUIRE(grid_fine->dtype == EDN_F32 || grid_fine->dtype == EDN_BF16, "edn_render_fine_fwd: bad grid dtype");
  FineArgs a{};
  int rc = make_grid_dev(grid_coarse, &a.gc);
  if (rc) return rc;
  rc = make_grid_dev(grid_fine, &a.gf);
  if (rc) return rc;
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  a.mlp = *mlp;
  a.ray_batch = ray_batch; a.z_vals = z_vals; a.noise = noise; a.n_rays = n_rays; a.S = n_samples; a.flags = flags;
  a.rmnearplane = rmnearplane; a.weights = weights; a.rgb = rgb; a.depth = depth; a.acc = acc; a.feat = feat;
  a.trace = nullptr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (precision == EDN_F32) return launch_fine_f32(a, grid_fine->dtype, st);
  if (precision == EDN_BF16) return launch_fine_tc(a, grid_fine->dtype, st);
  if (precision == EDN_TC32) {
    EDN_REQUIRE(mlp->tc_blob != nullptr, "edn_render_fine_fwd(tc32): edn_field_mlp.tc_blob is NULL (call edn_pack_fine_tc)");
    return launch_fine_tc3(a, grid_fine->dtype, reinterpret_cast<const uint8_t*>(mlp->tc_blob) + fine_tc3_blob_offset(), st);
  }
  set_error("edn_render_fine_fwd: bad precision %d", precision);
  return EDN_E_INVALID;
}

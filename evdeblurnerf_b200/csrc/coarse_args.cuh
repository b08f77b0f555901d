// Launch arguments and the sample-placement rule shared by the fp32 (coarse.cu) and tcgen05 (coarse_tc.cu) coarse-pass
// kernels.
#pragma once
#include "common.cuh"

namespace edn {

struct CoarseArgs {
  GridDev grid;
  edn_field_mlp mlp;
  const float* ray_batch;
  const float* t_vals;
  const float* t_rand;
  const float* noise;
  int64_t n_rays;
  int n_samples;
  int flags;
  float rmnearplane;
  float* z_vals;
  float* weights;
  float* rgb;
  float* depth;
  float* acc;
  float* feat;
};

// Coarse sample depth of (ray, s): renderer.py:163-178, bit-exact (no FMA contraction).
__device__ __forceinline__ float place_sample(const CoarseArgs& a, int64_t ray, int s, float near, float far) {
  const int S = a.n_samples;
  auto zt = [&](int i) -> float {
    const float t = __ldg(a.t_vals + i);
    if (!(a.flags & EDN_FLAG_LINDISP)) return __fadd_rn(__fmul_rn(near, 1.0f - t), __fmul_rn(far, t));
    return 1.0f / __fadd_rn(__fmul_rn(1.0f / near, 1.0f - t), __fmul_rn(1.0f / far, t));
  };
  const float z = zt(s);
  if (!a.t_rand) return z;
  const float lower = (s == 0) ? z : 0.5f * __fadd_rn(z, zt(s - 1));
  const float upper = (s == S - 1) ? z : 0.5f * __fadd_rn(zt(s + 1), z);
  return __fadd_rn(lower, __fmul_rn(upper - lower, __ldg(a.t_rand + ray * S + s)));
}


int launch_coarse_tc(const CoarseArgs& a, int grid_dtype, cudaStream_t st);   // coarse_tc.cu
int launch_coarse_tc3(const CoarseArgs& a, int grid_dtype, cudaStream_t st);  // coarse_tc.cu: bf16 x 3 parity kernel (EDN_TC32)

}  // namespace edn

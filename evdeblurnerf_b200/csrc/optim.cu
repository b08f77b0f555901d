// Fused Adam sweep (torch.optim.Adam as the reference constructs it, run_nerf.py:272-274: betas (0.9, 0.999), eps 1e-8,
// optional L2 weight decay on color_net weights, run_nerf.py:244-250).  Pure HBM streaming: reads p, g, m, v and writes
// p, m, v once, float4-vectorised; the 36.6 M VM plane parameters dominate (~1 GB per step).
#include "common.cuh"

namespace edn {
namespace {

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float lr_c, float b1, float b2, float eps, float wd, float rsb2) {
  g = fmaf(wd, p, g);
  m = fmaf(b1, m, (1.f - b1) * g);
  v = fmaf(b2, v, (1.f - b2) * g * g);
  p -= lr_c * m / (sqrtf(v) * rsb2 + eps);
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float lr_c, float b1, float b2, float eps, float wd, float rsb2) {
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    adam1(pp.x, gg.x, mm.x, vv.x, lr_c, b1, b2, eps, wd, rsb2);
    adam1(pp.y, gg.y, mm.y, vv.y, lr_c, b1, b2, eps, wd, rsb2);
    adam1(pp.z, gg.z, mm.z, vv.z, lr_c, b1, b2, eps, wd, rsb2);
    adam1(pp.w, gg.w, mm.w, vv.w, lr_c, b1, b2, eps, wd, rsb2);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (n & 3)) { const int64_t i = (n4 << 2) + t; adam1(p[i], g[i], m[i], v[i], lr_c, b1, b2, eps, wd, rsb2); }
}

}  // namespace
}  // namespace edn

extern "C" int edn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int64_t step, void* stream) {
  using namespace edn;
  EDN_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "edn_adam_step: bad argument");
  EDN_REQUIRE((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
               reinterpret_cast<uintptr_t>(exp_avg_sq)) % 16 == 0, "edn_adam_step: buffers must be 16-byte aligned");
  if (n == 0) return EDN_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float lr_c = (float)((double)lr / bc1), rsb2 = (float)(1.0 / sqrt(bc2));
  int64_t blocks = ((n >> 2) + 255) / 256;
  const int64_t cap = 8 * (int64_t)num_sms();
  blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
  adam_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr_c, beta1, beta2, eps,
                                                                                 weight_decay, rsb2);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

// Error plumbing, layout packing and the stand-alone VM lookup entry of the C ABI (include/evdeblur_b200.h).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace edn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %s (%d) at %s", cudaGetErrorString(e), (int)e, what);
  return EDN_E_CUDA;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// [C][H*W] -> [H*W][C] through a 32x32 shared tile (coalesced on both sides).
template <typename T>
__global__ void pack_plane_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int64_t HW) {
  __shared__ float tile[32][33];
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const int64_t p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? src[(int64_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t p = p0 + j;
    const int c = c0 + threadIdx.x;
    if (c < C && p < HW) dst[p * C + c] = (T)tile[threadIdx.x][j];
  }
}

template <typename T>
__global__ void vm_sample_kernel(const GridDev g, const float* __restrict__ pts, float* __restrict__ feat, int64_t n) {
  __shared__ __align__(16) float basis_s[kAppComp * kAppDim];
  for (int i = threadIdx.x; i < kAppComp * kAppDim; i += blockDim.x) basis_s[i] = __ldg(g.basis_t + i);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  float ft[kAppDim];
  vm_sample_point<T>(g, basis_s, p, ft);
#pragma unroll
  for (int j = 0; j < kAppDim; ++j) feat[i * kAppDim + j] = ft[j];
}

}  // namespace edn

extern "C" const char* edn_last_error(void) { return edn::g_err; }
extern "C" int edn_abi_version(void) { return EDN_ABI_VERSION; }

extern "C" int edn_pack_vm_plane(const float* src_chw, void* dst_hwc, int32_t C, int32_t H, int32_t W, int32_t dst_dtype,
                                 void* stream) {
  using namespace edn;
  EDN_REQUIRE(src_chw && dst_hwc && C > 0 && H > 0 && W > 0, "edn_pack_vm_plane: bad argument");
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32)), block(32, 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dst_dtype == EDN_F32) pack_plane_kernel<float><<<grid, block, 0, st>>>(src_chw, (float*)dst_hwc, C, HW);
  else if (dst_dtype == EDN_BF16) pack_plane_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(src_chw, (__nv_bfloat16*)dst_hwc, C, HW);
  else { set_error("edn_pack_vm_plane: bad dtype %d", dst_dtype); return EDN_E_INVALID; }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_vm_sample(const edn_vm_grid* grid, const float* pts, float* feat, int64_t n, void* stream) {
  using namespace edn;
  EDN_REQUIRE(pts && feat, "edn_vm_sample: null pointer");
  GridDev g;
  int rc = make_grid_dev(grid, &g);
  if (rc) return rc;
  if (n <= 0) return n == 0 ? EDN_OK : EDN_E_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (grid->dtype == EDN_F32) vm_sample_kernel<float><<<blocks, 128, 0, st>>>(g, pts, feat, n);
  else if (grid->dtype == EDN_BF16) vm_sample_kernel<__nv_bfloat16><<<blocks, 128, 0, st>>>(g, pts, feat, n);
  else { set_error("edn_vm_sample: bad dtype %d", grid->dtype); return EDN_E_INVALID; }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

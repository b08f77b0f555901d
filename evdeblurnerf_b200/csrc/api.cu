// Error plumbing, layout packing and the stand-alone VM lookup entry of the C ABI (include/evdeblur_b200.h).
#include <cstdarg>
#include <cstdio>

#include <cublas_v2.h>

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

// The library keeps a little per-PROCESS device state (cuBLAS handle, cublasLt workspace, staging buffers of the tensor-core
// launchers): one process per GPU, as the framework is run (torchrun).  The first call that touches such state binds the process to
// its current device; a later call from another device fails loudly instead of reading the first device's buffers.
int bind_device(const char* who) {
  static int bound = -1;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error("%s: cudaGetDevice failed", who); return EDN_E_CUDA; }
  if (bound < 0) bound = dev;
  if (dev != bound) {
    set_error("%s: this process is bound to CUDA device %d (one process per GPU); called with device %d current", who, bound, dev);
    return EDN_E_UNSUPPORTED;
  }
  return EDN_OK;
}

// why blas_handle() returned NULL: the device binding (its message is kept) or cublasCreate
int blas_unavailable() {
  if (int rc = bind_device("cuBLAS handle")) return rc;
  set_error("cublasCreate failed");
  return EDN_E_CUDA;
}

cublasHandle_t blas_handle() {
  static cublasHandle_t h = nullptr;
  if (bind_device("cuBLAS handle") != EDN_OK) return nullptr;
  if (!h && cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) h = nullptr;
  return h;
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

// [H*W][C] -> (+)= [C][H*W]: gradient planes back to the reference parameter layout.
__global__ void unpack_plane_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int64_t HW, int accumulate) {
  __shared__ float tile[32][33];
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t p = p0 + j;
    const int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? src[p * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const int64_t p = p0 + threadIdx.x;
    if (c < C && p < HW) {
      float* d = dst + (int64_t)c * HW + p;
      *d = accumulate ? *d + tile[threadIdx.x][j] : tile[threadIdx.x][j];
    }
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

extern "C" int edn_unpack_vm_plane_grad(const float* src_hwc, float* dst_chw, int32_t C, int32_t H, int32_t W, int32_t accumulate,
                                        void* stream) {
  using namespace edn;
  EDN_REQUIRE(src_hwc && dst_chw && C > 0 && H > 0 && W > 0, "edn_unpack_vm_plane_grad: bad argument");
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32)), block(32, 8);
  unpack_plane_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src_hwc, dst_chw, C, HW, accumulate);
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

// ---- counter-based RNG for the stratified jitter / pdf samples / density noise (renderer.py:176, rays.py:162,
//      voxnerf.py:175).  Philox4x32-10 keyed by (seed, stream); element i draws from counter i: reproducible for a
//      given seed independent of launch geometry, no generator state on the host. --------------------------------------
namespace edn {
namespace {
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__global__ void fill_random_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint32_t stream_id, int normal, float scale) {
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // 4 outputs per thread
  if (i4 * 4 >= n) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)i4, (uint32_t)(i4 >> 32), stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  float v[4];
  if (!normal) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (float)(r[j] >> 8) * (1.0f / 16777216.0f) * scale;      // [0, 1): 24 random bits
  } else {
#pragma unroll
    for (int j = 0; j < 4; j += 2) {                                                              // Box-Muller
      const float u1 = ((float)(r[j] >> 8) + 1.0f) * (1.0f / 16777216.0f);                        // (0, 1]
      const float u2 = (float)(r[j + 1] >> 8) * (1.0f / 16777216.0f);
      const float rad = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincospif(2.0f * u2, &sn, &cs);
      v[j] = rad * cs * scale; v[j + 1] = rad * sn * scale;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i4 * 4 + j < n) out[i4 * 4 + j] = v[j];
}
}  // namespace
}  // namespace edn

extern "C" int edn_fill_random(float* out, int64_t n, uint64_t seed, uint32_t stream_id, int32_t normal, float scale, void* stream) {
  using namespace edn;
  EDN_REQUIRE(out && n >= 0, "edn_fill_random: bad argument");
  if (n == 0) return EDN_OK;
  const int64_t threads = (n + 3) / 4;
  fill_random_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, n, seed, stream_id, normal, scale);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

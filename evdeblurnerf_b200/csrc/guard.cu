// Device-side NaN / Inf guard of render_rays' result dict (networks/renderer.py:259-263).  The reference calls
// torch.isnan(v).any() / torch.isinf(v).any() on every returned tensor (>= 24 host synchronisations per call); here ONE launch
// scans up to EDN_GUARD_MAX_TENSORS tensors and ORs one bit per tensor into two flag words (flags[0]: NaN seen, flags[1]: Inf
// seen) that the host reads lazily.
#include "common.cuh"

namespace edn {
namespace {

struct GuardArgs {
  const float* x[EDN_GUARD_MAX_TENSORS];
  long long start[EDN_GUARD_MAX_TENSORS + 1];   // prefix sums of the tensor sizes in float4 units (rounded up)
  long long n[EDN_GUARD_MAX_TENSORS];           // element counts
  int n_tensors;
};

__global__ void __launch_bounds__(256) guard_kernel(const GuardArgs a, uint32_t* __restrict__ flags) {
  uint32_t nan_bits = 0u, inf_bits = 0u;
  const long long total = a.start[a.n_tensors];
  int t = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    while (i >= a.start[t + 1]) ++t;            // tensors are visited in order: t only grows along the grid-stride walk
    const long long e = (i - a.start[t]) * 4;
    const float* p = a.x[t] + e;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (e + 4 <= a.n[t] && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      for (int k = 0; k < 4; ++k) if (e + k < a.n[t]) v[k] = __ldg(p + k);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (isnan(v[k])) nan_bits |= 1u << t;
      if (isinf(v[k])) inf_bits |= 1u << t;
    }
  }
  nan_bits = __reduce_or_sync(0xffffffffu, nan_bits);
  inf_bits = __reduce_or_sync(0xffffffffu, inf_bits);
  if ((threadIdx.x & 31) == 0) {
    if (nan_bits) atomicOr(flags, nan_bits);
    if (inf_bits) atomicOr(flags + 1, inf_bits);
  }
}

}  // namespace
}  // namespace edn

extern "C" int edn_check_finite(const float* const* tensors_host, const int64_t* sizes_host, int32_t n_tensors, uint32_t* flags,
                                void* stream) {
  using namespace edn;
  EDN_REQUIRE(n_tensors >= 0 && n_tensors <= EDN_GUARD_MAX_TENSORS, "edn_check_finite: at most %d tensors per call", EDN_GUARD_MAX_TENSORS);
  EDN_REQUIRE(flags && (n_tensors == 0 || (tensors_host && sizes_host)), "edn_check_finite: null pointer");
  GuardArgs a{};
  long long off = 0;
  for (int t = 0; t < n_tensors; ++t) {
    EDN_REQUIRE(sizes_host[t] >= 0 && (sizes_host[t] == 0 || tensors_host[t]), "edn_check_finite: bad tensor %d", t);
    a.x[t] = tensors_host[t]; a.n[t] = sizes_host[t]; a.start[t] = off;
    off += (sizes_host[t] + 3) / 4;
  }
  a.start[n_tensors] = off;
  a.n_tensors = n_tensors;
  if (off == 0) return EDN_OK;
  const long long blocks = (off + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 8LL * num_sms() ? blocks : 8LL * num_sms());
  guard_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, flags);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

// Pieces shared by the backward-pass translation units (field_bwd.cu, ray_bwd.cu, loss_bwd.cu): the row-major cuBLAS GEMM
// wrapper for the plain tall contractions and the small elementwise / reduction kernels around them.
#pragma once
#include <cublas_v2.h>

#include "common.cuh"

namespace edn {

cublasHandle_t blas_handle();   // api.cu: process-wide cuBLAS handle (NULL if cublasCreate failed or the process is bound to another device)
int blas_unavailable();         // api.cu: sets the error text for a NULL handle, returns the status code

// Row-major C[M,N] (+)= op(A) op(B).  !ta: A stored [M][K] (lda); ta: A stored [K][M].  !tb: B stored [K][N]; tb: B stored [N][K].
template <typename T> struct CuType;
template <> struct CuType<float> { static constexpr cudaDataType_t v = CUDA_R_32F; };
template <> struct CuType<__nv_bfloat16> { static constexpr cudaDataType_t v = CUDA_R_16BF; };

// element <-> float conversions and 4-wide vector access for the activation storage type of the backward pass
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float4 ldv4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldv4(const __nv_bfloat16* p) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
}
__device__ __forceinline__ void stv4(float* p, const float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void stv4(__nv_bfloat16* p, const float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<const unsigned*>(&a);
  r.y = *reinterpret_cast<const unsigned*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// 16-byte vectors of the storage type: kVec<T> elements
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int n = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int n = 8;
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&t); }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// gemm_lt.cu: Y[M,N] = relu(X[M,K] W^T (+ bias)), row-major, activation fused into the cublasLt epilogue.  W stored [N][K]
// (nn.Linear layout) or, with w_kn, [K][N].  0 = done, 1 = not available (caller falls back to GEMM + relu_bias_kernel), < 0 = error.
// (The RELU_AUX / DRELU epilogues were tried for the backward masks: this cublasLt runs them as a separate "epilogue::globalKernel"
// pass plus a slower GEMM -- 26.8 -> 39.2 ms per training step -- so the backward keeps relu_mask_kernel.)
int lt_relu_linear(cudaDataType_t type, cublasComputeType_t ct, int64_t M, int N, int K, const void* X, int64_t ldx, const void* W,
                   int64_t ldw, bool w_kn, const float* bias, void* Y, int64_t ldy, cudaStream_t st);

struct Gemm {
  cublasHandle_t h;
  cublasComputeType_t ct;
  int operator()(bool ta, bool tb, int64_t M, int N, int64_t K, const float* A, int lda, const float* B, int ldb, float beta,
                 float* C, int ldc) const {
    const float alpha = 1.0f;
    const cublasStatus_t s = cublasGemmEx(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, N, (int)M, (int)K, &alpha,
                                          B, CUDA_R_32F, ldb, A, CUDA_R_32F, lda, &beta, C, CUDA_R_32F, ldc, ct, CUBLAS_GEMM_DEFAULT);
    if (s != CUBLAS_STATUS_SUCCESS) { set_error("cublasGemmEx failed (%d) M=%lld N=%d K=%lld", (int)s, (long long)M, N, (long long)K); return EDN_E_CUDA; }
    return 0;
  }
  // Y = relu(X W^T + bias): fused epilogue when cublasLt offers it, else GEMM + the elementwise pass `fallback`
  template <typename T, typename F>
  int relu_linear(int64_t M, int N, int K, const T* X, int ldx, const T* W, int ldw, const float* bias, T* Y, int ldy, cudaStream_t st,
                  F&& fallback, bool w_kn = false) const {
    const cublasComputeType_t c = (CuType<T>::v == CUDA_R_32F) ? ct : CUBLAS_COMPUTE_32F;
    // exact-fp32 GEMMs (the parity mode) stay on cublasGemmEx: the epilogue-fused fp32 SIMT kernels cublasLt picks are slower there
    const int rc = (CuType<T>::v == CUDA_R_32F && c == CUBLAS_COMPUTE_32F) ? 1 : lt_relu_linear(CuType<T>::v, c, M, N, K, X, ldx, W, ldw, w_kn, bias, Y, ldy, st);
    if (rc <= 0) return rc;
    const int rc2 = run(false, !w_kn, M, N, K, X, ldx, W, ldw, 0.f, Y, ldy);
    if (rc2) return rc2;
    fallback();
    return 0;
  }
  // typed variant: A and B share a storage type (fp32 -> this->ct, bf16 -> fp32 accumulation), C may be wider (fp32 weight gradients)
  template <typename TA, typename TC>
  int run(bool ta, bool tb, int64_t M, int N, int64_t K, const TA* A, int lda, const TA* B, int ldb, float beta, TC* C, int ldc) const {
    const float alpha = 1.0f;
    const cublasComputeType_t c = (CuType<TA>::v == CUDA_R_32F) ? ct : CUBLAS_COMPUTE_32F;
    const cublasStatus_t s = cublasGemmEx(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, N, (int)M, (int)K, &alpha,
                                          B, CuType<TA>::v, ldb, A, CuType<TA>::v, lda, &beta, C, CuType<TC>::v, ldc, c, CUBLAS_GEMM_DEFAULT);
    if (s != CUBLAS_STATUS_SUCCESS) { set_error("cublasGemmEx failed (%d) M=%lld N=%d K=%lld", (int)s, (long long)M, N, (long long)K); return EDN_E_CUDA; }
    return 0;
  }
};


namespace {

// Y[m][0..n) = relu(Y + bias); one 16-byte vector per thread (n and ld multiples of the vector width)
template <typename AT>
__global__ void relu_bias_kernel(AT* __restrict__ Y, int ld, int n, int64_t M, const float* __restrict__ bias) {
  constexpr int V = Vec16<AT>::n;
  const int nq = n / V;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / nq;
  const int j = (int)(t % nq) * V;
  if (m >= M) return;
  Vec16<AT> x;
  x.load(Y + m * ld + j);
#pragma unroll
  for (int i = 0; i < V; ++i) x.v[i] = fmaxf(x.v[i] + (bias ? __ldg(bias + j + i) : 0.f), 0.f);
  x.store(Y + m * ld + j);
}

// Y[m][0..n) += bias   (Y contiguous, row length n)
__global__ void add_bias_kernel(float* __restrict__ Y, int n, int64_t M, const float* __restrict__ bias) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * n) return;
  Y[t] += bias[t % n];
}

// D[m][j] = H[m][j] > 0 ? D[m][j] : 0   (ReLU backward; H is the post-activation)
template <typename AT>
__global__ void relu_mask_kernel(AT* __restrict__ D, const AT* __restrict__ H, int ld, int n, int64_t M) {
  constexpr int V = Vec16<AT>::n;
  const int nq = n / V;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / nq;
  const int j = (int)(t % nq) * V;
  if (m >= M) return;
  Vec16<AT> d, h;
  d.load(D + m * ld + j);
  h.load(H + m * ld + j);
#pragma unroll
  for (int i = 0; i < V; ++i) d.v[i] = h.v[i] > 0.f ? d.v[i] : 0.f;
  d.store(D + m * ld + j);
}

// out[j] += sum_m D[m][j]  (bias gradients)
template <typename AT>
__global__ void colsum_kernel(const AT* __restrict__ D, int ld, int n, int64_t M, float* __restrict__ out) {
  const int j = threadIdx.x;
  if (j >= n) return;
  const int64_t r0 = (int64_t)blockIdx.x * 512, r1 = min(r0 + 512, M);
  float acc = 0.f;
  for (int64_t m = r0; m < r1; ++m) acc += to_f(D[m * ld + j]);
  atomicAdd(out + j, acc);
}

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }


}  // namespace
}  // namespace edn

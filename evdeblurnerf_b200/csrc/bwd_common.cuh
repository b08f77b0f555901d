// Pieces shared by the backward-pass translation units (field_bwd.cu, ray_bwd.cu, loss_bwd.cu): the row-major cuBLAS GEMM
// wrapper for the plain tall contractions and the small elementwise / reduction kernels around them.
#pragma once
#include <cublas_v2.h>

#include "common.cuh"

namespace edn {

cublasHandle_t blas_handle();   // api.cu: process-wide cuBLAS handle (NULL if cublasCreate failed)

// Row-major C[M,N] (+)= op(A) op(B).  !ta: A stored [M][K] (lda); ta: A stored [K][M].  !tb: B stored [K][N]; tb: B stored [N][K].
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
};


namespace {

// Y[m][0..n) = relu(Y + bias)
__global__ void relu_bias_kernel(float* __restrict__ Y, int ld, int n, int64_t M, const float* __restrict__ bias) {
  const int nq = n >> 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / nq;
  const int j = (int)(t % nq) * 4;
  if (m >= M) return;
  float4 v = *reinterpret_cast<float4*>(Y + m * ld + j);
  if (bias) { v.x += __ldg(bias + j); v.y += __ldg(bias + j + 1); v.z += __ldg(bias + j + 2); v.w += __ldg(bias + j + 3); }
  v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  *reinterpret_cast<float4*>(Y + m * ld + j) = v;
}

// Y[m][0..n) += bias   (Y contiguous, row length n)
__global__ void add_bias_kernel(float* __restrict__ Y, int n, int64_t M, const float* __restrict__ bias) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * n) return;
  Y[t] += bias[t % n];
}

// D[m][j] = H[m][j] > 0 ? D[m][j] : 0   (ReLU backward; H is the post-activation)
__global__ void relu_mask_kernel(float* __restrict__ D, const float* __restrict__ H, int ld, int n, int64_t M) {
  const int nq = n >> 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / nq;
  const int j = (int)(t % nq) * 4;
  if (m >= M) return;
  float4 d = *reinterpret_cast<float4*>(D + m * ld + j);
  const float4 h = *reinterpret_cast<const float4*>(H + m * ld + j);
  d.x = h.x > 0.f ? d.x : 0.f; d.y = h.y > 0.f ? d.y : 0.f; d.z = h.z > 0.f ? d.z : 0.f; d.w = h.w > 0.f ? d.w : 0.f;
  *reinterpret_cast<float4*>(D + m * ld + j) = d;
}

// out[j] += sum_m D[m][j]  (bias gradients)
__global__ void colsum_kernel(const float* __restrict__ D, int ld, int n, int64_t M, float* __restrict__ out) {
  const int j = threadIdx.x;
  if (j >= n) return;
  const int64_t r0 = (int64_t)blockIdx.x * 512, r1 = min(r0 + 512, M);
  float acc = 0.f;
  for (int64_t m = r0; m < r1; ++m) acc += D[m * ld + j];
  atomicAdd(out + j, acc);
}

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }


}  // namespace
}  // namespace edn

// fp32 SIMT GEMM building blocks of the parity-path kernels (fine_f32.cu, nerf_f32.cu): a 64-row activation tile in
// shared memory (row stride kLda) times a transposed weight matrix streamed from L2 in 16-row K chunks (cp.async double
// buffer); 256 threads, warp w owns rows 8w..8w+7, lane l owns columns 4l..4l+3 (+128 when N = 256).
#pragma once
#include "common.cuh"

namespace edn {

constexpr int kFineThreads = 256;
constexpr int kTileM = 64;
constexpr int kLda = 260;          // 256 + 4 floats of padding
constexpr int kKc = 16;            // K chunk
constexpr int kFH = 256;           // hidden width

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// acc[8][NT] = A[64][K] (smem, row stride kLda) x Wt[K][N] (global, N = 32*NT contiguous).  Warp w owns rows
// 8w..8w+7, lane l owns columns 4l..4l+3 (+128 for the second quad when NT == 8).
template <int NT>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ A, const float* __restrict__ Wt, int K,
                                          float (&acc)[8][NT], float* __restrict__ Ws, bool accumulate = false, int lda = kLda) {
  constexpr int N = 32 * NT;
  constexpr int kVecPerChunk = kKc * N / 4;                  // float4 per K chunk
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (!accumulate) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
  }
  const int nchunks = K / kKc;
  auto issue = [&](int c) {
    float* dst = Ws + (c & 1) * kKc * kFH;
    const float* src = Wt + (size_t)c * kKc * N;
    for (int v = tid; v < kVecPerChunk; v += kFineThreads) cp_async16(dst + 4 * v, src + 4 * v);
    cp_async_commit();
  };
  issue(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) { issue(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float* ws = Ws + (c & 1) * kKc * kFH;
    const float* a0 = A + (warp * 8) * lda + c * kKc;
#pragma unroll
    for (int kk = 0; kk < kKc; kk += 4) {
      float4 av[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = *reinterpret_cast<const float4*>(a0 + i * lda + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float wv[NT];
#pragma unroll
        for (int h = 0; h < NT / 4; ++h) {
          const float4 w4 = *reinterpret_cast<const float4*>(ws + (kk + q) * N + h * 128 + lane * 4);
          wv[4 * h + 0] = w4.x; wv[4 * h + 1] = w4.y; wv[4 * h + 2] = w4.z; wv[4 * h + 3] = w4.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float ax = (q == 0) ? av[i].x : (q == 1) ? av[i].y : (q == 2) ? av[i].z : av[i].w;
#pragma unroll
          for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(ax, wv[j], acc[i][j]);
        }
      }
    }
    __syncthreads();
  }
}

// A[row][col] <- act(acc + bias[col]); optional raw copy to global (row stride N) for the rows < valid_rows.
template <int NT, bool RELU>
__device__ __forceinline__ void store_tile(float* __restrict__ A, const float (&acc)[8][NT], const float* __restrict__ bias_s,
                                           const float* __restrict__ bias_g, float* __restrict__ gout, int valid_rows) {
  constexpr int N = 32 * NT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int h = 0; h < NT / 4; ++h) {
    const int col = h * 128 + lane * 4;
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias_s) { b[0] = bias_s[col]; b[1] = bias_s[col + 1]; b[2] = bias_s[col + 2]; b[3] = bias_s[col + 3]; }
    if (bias_g) { b[0] += __ldg(bias_g + col); b[1] += __ldg(bias_g + col + 1); b[2] += __ldg(bias_g + col + 2); b[3] += __ldg(bias_g + col + 3); }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = warp * 8 + i;
      float4 v = make_float4(acc[i][4 * h + 0] + b[0], acc[i][4 * h + 1] + b[1], acc[i][4 * h + 2] + b[2], acc[i][4 * h + 3] + b[3]);
      if (gout && row < valid_rows) *reinterpret_cast<float4*>(gout + (size_t)row * N + col) = v;
      if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      *reinterpret_cast<float4*>(A + row * kLda + col) = v;
    }
  }
}

// out[row][0..NO) = A[row][0..K) . Wt[K][ldw]  for the 64 rows: 4 threads per row, strided over k, shuffle-reduced.
template <int NO>
__device__ __forceinline__ void dot_rows(const float* __restrict__ A, const float* __restrict__ Wt, int ldw, int K,
                                         float (&out)[NO]) {
  const int row = threadIdx.x >> 2, part = threadIdx.x & 3;
  float s[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) s[o] = 0.f;
  for (int k = part; k < K; k += 4) {
    const float av = A[row * kLda + k];
#pragma unroll
    for (int o = 0; o < NO; ++o) s[o] = fmaf(av, __ldg(Wt + (size_t)k * ldw + o), s[o]);
  }
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    s[o] += __shfl_xor_sync(0xffffffffu, s[o], 1);
    s[o] += __shfl_xor_sync(0xffffffffu, s[o], 2);
    out[o] = s[o];
  }
}


}  // namespace edn

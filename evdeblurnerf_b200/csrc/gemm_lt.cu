// Y = relu(X W^T (+ bias)) as ONE cublasLt matmul with the activation in the GEMM epilogue -- the forward-recompute layers of the
// backward pass (field_bwd.cu, awp.cu) otherwise pay a separate read-modify-write pass over every hidden activation.
// Library GEMM, like the cublasGemmEx calls next to it (plain tall contractions).  Host side only: descriptors and
// the heuristic's algorithm are cached per problem shape; one lazily allocated workspace per process (single host thread, as the
// reference's training loop is).
#include <cublasLt.h>

#include <map>
#include <tuple>

#include "bwd_common.cuh"

namespace edn {
namespace {

struct LtPlan {
  cublasLtMatmulDesc_t op = nullptr;
  cublasLtMatrixLayout_t a = nullptr, b = nullptr, c = nullptr;
  cublasLtMatmulAlgo_t algo{};
  bool ok = false;
};
using LtKey = std::tuple<int, int, int64_t, int, int, int64_t, int64_t, int64_t, bool, bool>;

constexpr size_t kLtWorkspace = 32u << 20;

void* lt_workspace() {
  static void* ws = nullptr;
  static bool tried = false;
  if (!tried) { tried = true; if (cudaMalloc(&ws, kLtWorkspace) != cudaSuccess) { ws = nullptr; cudaGetLastError(); } }
  return ws;
}

LtPlan make_plan(cublasLtHandle_t lt, cudaDataType_t type, cublasComputeType_t ct, int64_t M, int N, int K, int64_t ldx, int64_t ldw,
                 int64_t ldy, bool has_bias, bool w_kn) {
  LtPlan p;
  const cublasOperation_t ta = w_kn ? CUBLAS_OP_N : CUBLAS_OP_T, tb = CUBLAS_OP_N;
  const cublasLtEpilogue_t ep = has_bias ? CUBLASLT_EPILOGUE_RELU_BIAS : CUBLASLT_EPILOGUE_RELU;
  if (cublasLtMatmulDescCreate(&p.op, ct, CUDA_R_32F) != CUBLAS_STATUS_SUCCESS) return p;
  bool good = cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta)) == CUBLAS_STATUS_SUCCESS &&
              cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb)) == CUBLAS_STATUS_SUCCESS &&
              cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_EPILOGUE, &ep, sizeof(ep)) == CUBLAS_STATUS_SUCCESS;
  // column-major view of the row-major operands: W [N][K] -> K x N (transposed in the product) or W [K][N] -> N x K,
  // X [M][K] -> K x M, Y [M][N] -> N x M
  good = good && cublasLtMatrixLayoutCreate(&p.a, type, (uint64_t)(w_kn ? N : K), (uint64_t)(w_kn ? K : N), ldw) == CUBLAS_STATUS_SUCCESS &&
         cublasLtMatrixLayoutCreate(&p.b, type, (uint64_t)K, (uint64_t)M, ldx) == CUBLAS_STATUS_SUCCESS &&
         cublasLtMatrixLayoutCreate(&p.c, type, (uint64_t)N, (uint64_t)M, ldy) == CUBLAS_STATUS_SUCCESS;
  if (!good) return p;
  cublasLtMatmulPreference_t pref = nullptr;
  if (cublasLtMatmulPreferenceCreate(&pref) != CUBLAS_STATUS_SUCCESS) return p;
  size_t ws = lt_workspace() ? kLtWorkspace : 0;
  cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws, sizeof(ws));
  cublasLtMatmulHeuristicResult_t res{};
  int n = 0;
  const cublasStatus_t s = cublasLtMatmulAlgoGetHeuristic(lt, p.op, p.a, p.b, p.c, p.c, pref, 1, &res, &n);
  cublasLtMatmulPreferenceDestroy(pref);
  if (s == CUBLAS_STATUS_SUCCESS && n > 0) { p.algo = res.algo; p.ok = true; }
  return p;
}

}  // namespace

// 0 = done; 1 = configuration not available through cublasLt (the caller runs GEMM + relu_bias_kernel instead); < 0 = error.
// bias (fp32, length N) is only fused when the storage type is fp32 (the epilogue's bias vector has the type of Y).
int lt_relu_linear(cudaDataType_t type, cublasComputeType_t ct, int64_t M, int N, int K, const void* X, int64_t ldx, const void* W,
                   int64_t ldw, bool w_kn, const float* bias, void* Y, int64_t ldy, cudaStream_t st) {
  if (bias && type != CUDA_R_32F) return 1;
  cublasHandle_t h = blas_handle();
  if (!h) return 1;
  cublasLtHandle_t lt = reinterpret_cast<cublasLtHandle_t>(h);
  static std::map<LtKey, LtPlan> plans;
  const LtKey key{(int)type, (int)ct, M, N, K, ldx, ldw, ldy, bias != nullptr, w_kn};
  auto it = plans.find(key);
  if (it == plans.end()) it = plans.emplace(key, make_plan(lt, type, ct, M, N, K, ldx, ldw, ldy, bias != nullptr, w_kn)).first;
  const LtPlan& p = it->second;
  if (!p.ok) return 1;
  if (bias && cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias)) != CUBLAS_STATUS_SUCCESS) return 1;
  const float alpha = 1.0f, beta = 0.0f;
  void* ws = lt_workspace();
  const cublasStatus_t s = cublasLtMatmul(lt, p.op, &alpha, W, p.a, X, p.b, &beta, Y, p.c, Y, p.c, &p.algo, ws, ws ? kLtWorkspace : 0, st);
  if (s != CUBLAS_STATUS_SUCCESS) { set_error("cublasLtMatmul failed (%d) M=%lld N=%d K=%d", (int)s, (long long)M, N, K); return EDN_E_CUDA; }
  return 0;
}

}  // namespace edn

// Backward of the DP-NeRF rigid blur kernel + render() prologue (rbk.cu; networks/dpnerf/blurmodel.py:129-173, 51-82,
// utils/rigid_warping.py:18-154, renderer.py:423-446, utils/rays.py:104-145): d ray_batch [N*E][11] and d weight [N][E]
// -> gradients of the kernel-net parameters.  "Gradient to r, v, w heads flows back through the whole renderer via pts"
// (SURVEY 8 a1): d ray_batch is what edn_render_field_bwd accumulated.
//
// The SE(3) exponential map, the view-direction normalisation and the NDC projection are differentiated in FORWARD mode
// with 6 tangent lanes (one per component of the motion's rotation / translation parameters) -- one thread per
// (ray, motion) evaluates the 9 ray-batch outputs as dual numbers and contracts their tangents with the upstream gradient.
// The three tiny head MLPs (32 -> 32 -> 3M | E over N rays) are recomputed and back-propagated with plain GEMMs.
#include "bwd_common.cuh"

namespace edn {
namespace {

constexpr int kW = 32;
constexpr int kT = 6;   // tangent lanes

struct Dual {
  float v, d[kT];
  __device__ Dual() {}
  __device__ Dual(float c) : v(c) {
#pragma unroll
    for (int i = 0; i < kT; ++i) d[i] = 0.f;
  }
  friend __device__ Dual operator+(const Dual& a, const Dual& b) {
    Dual r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < kT; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
  }
  friend __device__ Dual operator-(const Dual& a, const Dual& b) {
    Dual r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < kT; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
  }
  friend __device__ Dual operator-(const Dual& a) {
    Dual r; r.v = -a.v;
#pragma unroll
    for (int i = 0; i < kT; ++i) r.d[i] = -a.d[i];
    return r;
  }
  friend __device__ Dual operator*(const Dual& a, const Dual& b) {
    Dual r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < kT; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
  }
  friend __device__ Dual operator/(const Dual& a, const Dual& b) {
    Dual r; const float inv = 1.0f / b.v; r.v = a.v * inv;
#pragma unroll
    for (int i = 0; i < kT; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
  }
};
__device__ Dual dsqrt(const Dual& a) {
  Dual r; r.v = sqrtf(a.v);
  const float k = r.v > 0.f ? 0.5f / r.v : 0.f;      // subgradient 0 at 0, like torch.norm
#pragma unroll
  for (int i = 0; i < kT; ++i) r.d[i] = a.d[i] * k;
  return r;
}
__device__ Dual dsin(const Dual& a) {
  Dual r; float s, c; sincosf(a.v, &s, &c); r.v = s;
#pragma unroll
  for (int i = 0; i < kT; ++i) r.d[i] = a.d[i] * c;
  return r;
}
__device__ Dual dcos(const Dual& a) {
  Dual r; float s, c; sincosf(a.v, &s, &c); r.v = c;
#pragma unroll
  for (int i = 0; i < kT; ++i) r.d[i] = -a.d[i] * s;
  return r;
}

// ray_batch row (o, d, viewdirs = out[0..8]) of a ray (wo, wd): view-direction normalisation + NDC projection, same arithmetic as
// ray_batch_kernel / rbk_warp_ndc_kernel (renderer.py:423-446, utils/rays.py:104-145).
__device__ void finish_ray_dual(const Dual wo[3], const Dual wd[3], float Hf, float Wf, float focal, int ndc, Dual out[9]) {
  const Dual nrm = dsqrt(wd[0] * wd[0] + wd[1] * wd[1] + wd[2] * wd[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) out[6 + i] = wd[i] / nrm;
  if (ndc) {
    const float near = 1.0f;
    const Dual t = -(Dual(near) + wo[2]) / wd[2];
    Dual oo[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) oo[i] = wo[i] + t * wd[i];
    const Dual ox_oz = oo[0] / oo[2], oy_oz = oo[1] / oo[2];
    const float sx = -1.f / (Wf / (2.f * focal)), sy = -1.f / (Hf / (2.f * focal));
    out[0] = Dual(sx) * ox_oz;
    out[1] = Dual(sy) * oy_oz;
    out[2] = Dual(1.f) + Dual(2.f * near) / oo[2];
    out[3] = Dual(sx) * (wd[0] / wd[2] - ox_oz);
    out[4] = Dual(sy) * (wd[1] / wd[2] - oy_oz);
    out[5] = Dual(1.f) - out[2];
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) { out[i] = wo[i]; out[3 + i] = wd[i]; }
  }
}


// ray_batch row (o, d, viewdirs = out[0..8]) of the sub-ray warped by (rot, trn): same arithmetic as rbk_warp_ndc_kernel.
__device__ void warp_ray_dual(const Dual rot[3], const Dual trn[3], const float o[3], const float d[3], float Hf, float Wf, float focal,
                              int ndc, Dual out[9]) {
  const Dual theta = dsqrt(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]) + Dual(1.0e-10f);
  const Dual w[3] = {rot[0] / theta, rot[1] / theta, rot[2] / theta};
  const Dual v[3] = {trn[0] / theta, trn[1] / theta, trn[2] / theta};
  const Dual zero(0.f);
  const Dual Wm[3][3] = {{zero, -w[2], w[1]}, {w[2], zero, -w[0]}, {-w[1], w[0], zero}};
  Dual WW[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) WW[i][j] = Wm[i][0] * Wm[0][j] + Wm[i][1] * Wm[1][j] + Wm[i][2] * Wm[2][j];
  const Dual sn = dsin(theta), cs = dcos(theta), omc = Dual(1.f) - cs, tms = theta - sn;
  Dual wo[3], wd[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Dual p(0.f), ro(0.f), re(0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float eye = (i == j) ? 1.f : 0.f;
      const Dual Rij = Dual(eye) + sn * Wm[i][j] + omc * WW[i][j];
      p = p + (theta * Dual(eye) + omc * Wm[i][j] + tms * WW[i][j]) * v[j];
      ro = ro + Rij * Dual(o[j]);
      re = re + Rij * Dual(o[j] + d[j]);
    }
    wo[i] = ro + p;
    wd[i] = (re + p) - wo[i];
  }
  finish_ray_dual(wo, wd, Hf, Wf, focal, ndc, out);
}

// one thread per (ray, motion): d r, d v [N][3M] (component-major, motion-minor; before the rv_window scale is undone)
__global__ void rbk_warp_bwd_kernel(const float* __restrict__ rays, const float* __restrict__ r_out, const float* __restrict__ v_out,
                                    int64_t N, int M, float rv_window, int H, int W, float focal, int ndc,
                                    const float* __restrict__ d_rb, float* __restrict__ d_r, float* __restrict__ d_v) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = t / M;
  const int mi = (int)(t % M);
  if (n >= N) return;
  const int E = M + 1;
  float o[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i] = rays[n * 6 + 2 * i]; d[i] = rays[n * 6 + 2 * i + 1]; }
  Dual rot[3], trn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    rot[j] = Dual(r_out[n * 3 * M + j * M + mi] * rv_window); rot[j].d[j] = 1.f;
    trn[j] = Dual(v_out[n * 3 * M + j * M + mi] * rv_window); trn[j].d[3 + j] = 1.f;
  }
  Dual out[9];
  warp_ray_dual(rot, trn, o, d, (float)H, (float)W, focal, ndc, out);
  const float* g = d_rb + (n * E + mi + 1) * 11;
  const float gq[9] = {g[0], g[1], g[2], g[3], g[4], g[5], g[8], g[9], g[10]};
  float gin[kT];
#pragma unroll
  for (int k = 0; k < kT; ++k) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) s = fmaf(gq[q], out[q].d[k], s);
    gin[k] = s * rv_window;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    d_r[n * 3 * M + j * M + mi] = gin[j];
    d_v[n * 3 * M + j * M + mi] = gin[3 + j];
  }
}

// edn_build_ray_batch_bwd: one thread per ray, the 6 tangent lanes seed (o, d) themselves
__global__ void ray_batch_bwd_kernel(const float* __restrict__ rays, int64_t R, int H, int W, float focal, int ndc,
                                     const float* __restrict__ d_rb, float* __restrict__ d_rays) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= R) return;
  Dual wo[3], wd[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    wo[i] = Dual(rays[n * 6 + 2 * i]); wo[i].d[i] = 1.f;
    wd[i] = Dual(rays[n * 6 + 2 * i + 1]); wd[i].d[3 + i] = 1.f;
  }
  Dual out[9];
  finish_ray_dual(wo, wd, (float)H, (float)W, focal, ndc, out);
  const float* g = d_rb + n * 11;
  const float gq[9] = {g[0], g[1], g[2], g[3], g[4], g[5], g[8], g[9], g[10]};
#pragma unroll
  for (int k = 0; k < kT; ++k) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) s = fmaf(gq[q], out[q].d[k], s);
    d_rays[n * 6 + 2 * (k % 3) + (k / 3)] = s;
  }
}

// weight = sigmoid(l) / (sum sigmoid(l) + 1e-10)  (blurmodel.py:165-166): d weight -> d l
__global__ void rbk_weight_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ d_weight, int64_t N, int E,
                                      float* __restrict__ d_logits) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float tot = 0.f, dot = 0.f;
  for (int j = 0; j < E; ++j) {
    const float s = sigmoidf_(logits[n * E + j]);
    tot += s;
    dot = fmaf(d_weight ? d_weight[n * E + j] : 0.f, s, dot);
  }
  const float inv = 1.0f / (tot + 1e-10f);
  for (int j = 0; j < E; ++j) {
    const float s = sigmoidf_(logits[n * E + j]);
    const float ds = ((d_weight ? d_weight[n * E + j] : 0.f) - dot * inv) * inv;
    d_logits[n * E + j] = ds * s * (1.f - s);
  }
}

__global__ void gather_embed_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, int64_t N, float* __restrict__ emb) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * kW) return;
  emb[t] = table[idx[t / kW] * kW + (t % kW)];
}
__global__ void scatter_embed_kernel(const float* __restrict__ d_emb, const int64_t* __restrict__ idx, int64_t N, float* __restrict__ d_table) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * kW) return;
  atomicAdd(d_table + idx[t / kW] * kW + (t % kW), d_emb[t]);
}

// rbk_weighted_sum backward: out[n][c] = sum_e w[n][e] x[n*E+e][c]
__global__ void weighted_sum_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ d_out,
                                        int64_t N, int E, int64_t Cn, float* __restrict__ d_x, float* __restrict__ d_w) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (n, e)
  if (t >= N * E) return;
  const int64_t n = t / E;
  const float we = w[t];
  float dot = 0.f;
  for (int64_t c = 0; c < Cn; ++c) {
    const float g = d_out[n * Cn + c];
    if (d_x) d_x[t * Cn + c] = we * g;
    dot = fmaf(g, x[t * Cn + c], dot);
  }
  if (d_w) d_w[t] = dot;
}

}  // namespace
}  // namespace edn

extern "C" int64_t edn_rbk_bwd_workspace_floats(int64_t n_rays, int32_t num_motion) {
  if (n_rays < 0 || num_motion < 0) return -1;
  const int64_t E = num_motion + 1, o = 2 * 3 * (int64_t)num_motion + E;
  return n_rays * (2 * 32 + 4 * 32 + 2 * o);
}

extern "C" int edn_rbk_warp_ndc_bwd(const edn_rbk_params* p, const float* rays, const int64_t* images_idx, int64_t n_rays, int32_t H,
                                    int32_t W, float focal, int32_t ndc, const float* d_ray_batch, const float* d_weight,
                                    const float* d_img_embed, const edn_rbk_grads* g, float* workspace, void* stream) {
  using namespace edn;
  EDN_REQUIRE(p && rays && images_idx && g && workspace, "edn_rbk_warp_ndc_bwd: null pointer");
  const int M = p->num_motion, E = M + 1;
  EDN_REQUIRE(M >= 0 && M < 16, "edn_rbk_warp_ndc_bwd: num_motion must be in [0,16)");
  EDN_REQUIRE(M == 0 || d_ray_batch, "edn_rbk_warp_ndc_bwd: d_ray_batch is required when num_motion > 0");
  EDN_REQUIRE(g->img_embed && g->w_branch_w && g->w_branch_b && g->w_linear_w && g->w_linear_b, "edn_rbk_warp_ndc_bwd: null gradient buffer");
  EDN_REQUIRE(M == 0 || (g->r_branch_w && g->r_branch_b && g->v_branch_w && g->v_branch_b && g->r_linear_w && g->r_linear_b &&
                         g->v_linear_w && g->v_linear_b), "edn_rbk_warp_ndc_bwd: null r/v gradient buffer");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  cublasHandle_t h = blas_handle();
  if (!h) return blas_unavailable();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cublasSetStream(h, st) != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetStream failed"); return EDN_E_CUDA; }
  const Gemm gemm{h, CUBLAS_COMPUTE_32F};
  const int64_t N = n_rays;
  float* base = workspace;
  auto take = [&](int64_t per) { float* q = base; base += per * N; return q; };
  float* emb = take(kW);
  float* d_emb = take(kW);
  float* Hb[3] = {take(kW), take(kW), take(kW)};
  float* dHb = take(kW);
  const int on[3] = {3 * M, 3 * M, E};
  float* out[3] = {take(on[0]), take(on[1]), take(on[2])};
  float* d_out[3] = {take(on[0]), take(on[1]), take(on[2])};
  const float* bw[3] = {p->r_branch_w, p->v_branch_w, p->w_branch_w};
  const float* bb[3] = {p->r_branch_b, p->v_branch_b, p->w_branch_b};
  const float* lw[3] = {p->r_linear_w, p->v_linear_w, p->w_linear_w};
  const float* lb[3] = {p->r_linear_b, p->v_linear_b, p->w_linear_b};
  float* gbw[3] = {g->r_branch_w, g->v_branch_w, g->w_branch_w};
  float* gbb[3] = {g->r_branch_b, g->v_branch_b, g->w_branch_b};
  float* glw[3] = {g->r_linear_w, g->v_linear_w, g->w_linear_w};
  float* glb[3] = {g->r_linear_b, g->v_linear_b, g->w_linear_b};
#define EDN_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
  gather_embed_kernel<<<blocks_for(N * kW, 256), 256, 0, st>>>(p->img_embed, images_idx, N, emb);
  for (int t = (M > 0 ? 0 : 2); t < 3; ++t) {       // recompute the heads
    EDN_RC(gemm(false, true, N, kW, kW, emb, kW, bw[t], kW, 0.f, Hb[t], kW));
    relu_bias_kernel<<<blocks_for(N * (kW / 4), 256), 256, 0, st>>>(Hb[t], kW, kW, N, bb[t]);
    EDN_RC(gemm(false, true, N, on[t], kW, Hb[t], kW, lw[t], kW, 0.f, out[t], on[t]));
    add_bias_kernel<<<blocks_for(N * on[t], 256), 256, 0, st>>>(out[t], on[t], N, lb[t]);
  }
  if (M > 0)
    rbk_warp_bwd_kernel<<<blocks_for(N * M, 128), 128, 0, st>>>(rays, out[0], out[1], N, M, p->rv_window, H, W, focal, ndc, d_ray_batch,
                                                               d_out[0], d_out[1]);
  rbk_weight_bwd_kernel<<<blocks_for(N, 128), 128, 0, st>>>(out[2], d_weight, N, E, d_out[2]);
  bool first = true;
  if (d_img_embed) {       // gradient of the view latents handed to AWP (awp.py:88-94)
    EDN_CUDA_OK(cudaMemcpyAsync(d_emb, d_img_embed, sizeof(float) * (size_t)N * kW, cudaMemcpyDeviceToDevice, st));
    first = false;
  }
  for (int t = (M > 0 ? 0 : 2); t < 3; ++t) {
    EDN_RC(gemm(true, false, on[t], kW, N, d_out[t], on[t], Hb[t], kW, 1.f, glw[t], kW));
    colsum_kernel<<<blocks_for(N, 512), 64, 0, st>>>(d_out[t], on[t], on[t], N, glb[t]);
    EDN_RC(gemm(false, false, N, kW, on[t], d_out[t], on[t], lw[t], kW, 0.f, dHb, kW));
    relu_mask_kernel<<<blocks_for(N * (kW / 4), 256), 256, 0, st>>>(dHb, Hb[t], kW, kW, N);
    EDN_RC(gemm(true, false, kW, kW, N, dHb, kW, emb, kW, 1.f, gbw[t], kW));
    colsum_kernel<<<blocks_for(N, 512), 64, 0, st>>>(dHb, kW, kW, N, gbb[t]);
    EDN_RC(gemm(false, false, N, kW, kW, dHb, kW, bw[t], kW, first ? 0.f : 1.f, d_emb, kW));
    first = false;
  }
#undef EDN_RC
  scatter_embed_kernel<<<blocks_for(N * kW, 256), 256, 0, st>>>(d_emb, images_idx, N, g->img_embed);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_weighted_sum_bwd(const float* x, const float* w, const float* d_out, int64_t n, int32_t n_exposure, int64_t channels,
                                    float* d_x, float* d_w, void* stream) {
  using namespace edn;
  EDN_REQUIRE(x && w && d_out && n_exposure > 0 && channels > 0, "edn_weighted_sum_bwd: bad argument");
  if (n <= 0) return n == 0 ? EDN_OK : EDN_E_INVALID;
  weighted_sum_bwd_kernel<<<blocks_for(n * n_exposure, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, w, d_out, n, n_exposure,
                                                                                                               channels, d_x, d_w);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_build_ray_batch_bwd(const float* rays, int64_t n_rays, int32_t H, int32_t W, float focal, int32_t ndc,
                                       const float* d_ray_batch, float* d_rays, void* stream) {
  using namespace edn;
  EDN_REQUIRE(rays && d_ray_batch && d_rays, "edn_build_ray_batch_bwd: null pointer");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  ray_batch_bwd_kernel<<<blocks_for(n_rays, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rays, n_rays, H, W, focal, ndc, d_ray_batch, d_rays);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

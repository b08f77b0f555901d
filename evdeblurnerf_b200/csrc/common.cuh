// Shared device helpers of the EvDeblurNeRF render path (sm_100a).  No torch headers anywhere under csrc/.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/evdeblur_b200.h"

namespace edn {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int num_sms();
int bind_device(const char* who);   // api.cu: one process per GPU -- binds on first use, EDN_E_UNSUPPORTED from another device

#define EDN_CUDA_OK(expr)                                   \
  do {                                                      \
    int _rc = ::edn::check_cuda((expr), #expr);             \
    if (_rc != 0) return _rc;                               \
  } while (0)

#define EDN_REQUIRE(cond, ...)                              \
  do {                                                      \
    if (!(cond)) {                                          \
      ::edn::set_error(__VA_ARGS__);                        \
      return EDN_E_INVALID;                                 \
    }                                                       \
  } while (0)

constexpr int kAppComp = 96;   // 64 + 16 + 16 (voxnerf.py app_n_comp)
constexpr int kAppDim = 32;    // basis_mat out
constexpr int kPeFreqPts = 10; // multires
constexpr int kPeFreqDir = 4;  // multires_views
constexpr int kPePts = 3 + 6 * kPeFreqPts;  // 63
constexpr int kPeDir = 3 + 6 * kPeFreqDir;  // 27

// Device-side copy of edn_vm_grid with the derived normalisation constants (voxnerf.py:90-91, 204).
struct GridDev {
  const void* plane[3];
  const void* line[3];
  int ph[3], pw[3], ll[3];
  float amin[3];
  float inv[3];  // 2 / (aabb_max - aabb_min), fp32 like invaabbSize
  const float* basis_t;
};

inline int make_grid_dev(const edn_vm_grid* g, GridDev* d) {
  if (!g) { set_error("null edn_vm_grid"); return EDN_E_INVALID; }
  if (g->n_comp[0] != 64 || g->n_comp[1] != 16 || g->n_comp[2] != 16) {
    set_error("app_n_comp must be {64,16,16}, got {%d,%d,%d}", g->n_comp[0], g->n_comp[1], g->n_comp[2]);
    return EDN_E_UNSUPPORTED;
  }
  for (int i = 0; i < 3; ++i) {
    if (!g->plane[i] || !g->line[i]) { set_error("null VM plane/line %d", i); return EDN_E_INVALID; }
    d->plane[i] = g->plane[i];
    d->line[i] = g->line[i];
    d->ph[i] = g->plane_h[i];
    d->pw[i] = g->plane_w[i];
    d->ll[i] = g->line_len[i];
    d->amin[i] = g->aabb_min[i];
    d->inv[i] = 2.0f / (g->aabb_max[i] - g->aabb_min[i]);
  }
  if (!g->basis_t) { set_error("null basis_t"); return EDN_E_INVALID; }
  d->basis_t = g->basis_t;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// grid_sample(bilinear, zeros padding, align_corners=True) tap arithmetic, as F.grid_sample computes it
// (voxnerf.py:145-148).  Offsets are texel indices (row-major [H][W]); an out-of-range tap gets weight 0.
// ---------------------------------------------------------------------------------------------------------------
struct Taps2 { int off[4]; float w[4]; };
struct Taps1 { int off[2]; float w[2]; };

__device__ __forceinline__ float unnormalize(float c, int size) {
  // ((c + 1) / 2) * (size - 1); no FMA contraction so that tap weights match the reference bit for bit.
  return __fmul_rn(__fmul_rn(__fadd_rn(c, 1.0f), 0.5f), (float)(size - 1));
}

__device__ __forceinline__ void plane_taps(float x, float y, int H, int W, Taps2& t) {
  const float ix = unnormalize(x, W), iy = unnormalize(y, H);
  const float x0 = floorf(ix), y0 = floorf(iy);
  const float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
  const float wx0 = x1 - ix, wx1 = ix - x0, wy0 = y1 - iy, wy1 = iy - y0;
  const bool vx0 = (x0 >= 0.0f) && (x0 <= (float)(W - 1)), vx1 = (x1 >= 0.0f) && (x1 <= (float)(W - 1));
  const bool vy0 = (y0 >= 0.0f) && (y0 <= (float)(H - 1)), vy1 = (y1 >= 0.0f) && (y1 <= (float)(H - 1));
  const int ix0 = vx0 ? (int)x0 : 0, ix1 = vx1 ? (int)x1 : 0, iy0 = vy0 ? (int)y0 : 0, iy1 = vy1 ? (int)y1 : 0;
  t.off[0] = iy0 * W + ix0; t.w[0] = (vx0 && vy0) ? __fmul_rn(wx0, wy0) : 0.0f;  // nw
  t.off[1] = iy0 * W + ix1; t.w[1] = (vx1 && vy0) ? __fmul_rn(wx1, wy0) : 0.0f;  // ne
  t.off[2] = iy1 * W + ix0; t.w[2] = (vx0 && vy1) ? __fmul_rn(wx0, wy1) : 0.0f;  // sw
  t.off[3] = iy1 * W + ix1; t.w[3] = (vx1 && vy1) ? __fmul_rn(wx1, wy1) : 0.0f;  // se
}

// Line = [1,C,L,1] sampled at (x = 0, y = v): the x taps collapse to texel 0 with weight 1 (ix = 0).
__device__ __forceinline__ void line_taps(float v, int L, Taps1& t) {
  const float iy = unnormalize(v, L);
  const float y0 = floorf(iy), y1 = y0 + 1.0f;
  const bool v0 = (y0 >= 0.0f) && (y0 <= (float)(L - 1)), v1 = (y1 >= 0.0f) && (y1 <= (float)(L - 1));
  t.off[0] = v0 ? (int)y0 : 0; t.w[0] = v0 ? (y1 - iy) : 0.0f;
  t.off[1] = v1 ? (int)y1 : 0; t.w[1] = v1 ? (iy - y0) : 0.0f;
}

template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  float4 o;
  o.x = __uint_as_float(r.x << 16); o.y = __uint_as_float(r.x & 0xffff0000u);
  o.z = __uint_as_float(r.y << 16); o.w = __uint_as_float(r.y & 0xffff0000u);
  return o;
}

// normalised coordinates of a point: (p - aabb_min) * invaabbSize - 1   (voxnerf.py:204)
__device__ __forceinline__ void normalize_pt(const GridDev& g, const float p[3], float n[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) n[i] = __fsub_rn(__fmul_rn(__fsub_rn(p[i], g.amin[i]), g.inv[i]), 1.0f);
}

// matMode = [[0,1],[0,2],[1,2]], vecMode = [2,1,0]   (voxnerf.py:99-100)
__device__ __forceinline__ void point_taps(const GridDev& g, const float n[3], Taps2 pt[3], Taps1 lt[3]) {
  plane_taps(n[0], n[1], g.ph[0], g.pw[0], pt[0]);
  plane_taps(n[0], n[2], g.ph[1], g.pw[1], pt[1]);
  plane_taps(n[1], n[2], g.ph[2], g.pw[2], pt[2]);
  line_taps(n[2], g.ll[0], lt[0]);
  line_taps(n[1], g.ll[1], lt[1]);
  line_taps(n[0], g.ll[2], lt[2]);
}

// 4 consecutive channels [c, c+4) of (plane_i sampled) * (line_i sampled); C = channel count of component i.
template <typename T>
__device__ __forceinline__ float4 gather4(const T* __restrict__ plane, const T* __restrict__ line, int C, int c,
                                          const Taps2& pt, const Taps1& lt) {
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 v = load4<T>(plane + (size_t)pt.off[k] * C + c);
    p.x = fmaf(v.x, pt.w[k], p.x); p.y = fmaf(v.y, pt.w[k], p.y);
    p.z = fmaf(v.z, pt.w[k], p.z); p.w = fmaf(v.w, pt.w[k], p.w);
  }
  float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float4 v = load4<T>(line + (size_t)lt.off[k] * C + c);
    l.x = fmaf(v.x, lt.w[k], l.x); l.y = fmaf(v.y, lt.w[k], l.y);
    l.z = fmaf(v.z, lt.w[k], l.z); l.w = fmaf(v.w, lt.w[k], l.w);
  }
  return make_float4(p.x * l.x, p.y * l.y, p.z * l.z, p.w * l.w);
}

// ft[32] += basis_t[c..c+3][:] * prod   (basis_t in shared memory, [96][32])
__device__ __forceinline__ void basis_accum4(const float* __restrict__ basis_t_s, int c, const float4 g, float ft[kAppDim]) {
  const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4* row = reinterpret_cast<const float4*>(basis_t_s + (c + q) * kAppDim);
#pragma unroll
    for (int j = 0; j < kAppDim / 4; ++j) {
      const float4 w = row[j];
      ft[4 * j + 0] = fmaf(w.x, gv[q], ft[4 * j + 0]); ft[4 * j + 1] = fmaf(w.y, gv[q], ft[4 * j + 1]);
      ft[4 * j + 2] = fmaf(w.z, gv[q], ft[4 * j + 2]); ft[4 * j + 3] = fmaf(w.w, gv[q], ft[4 * j + 3]);
    }
  }
}

// VoxelNeRFBase.sample for one point held by one thread: ft[32] = basis_mat(plane (.) line)
template <typename T>
__device__ __forceinline__ void vm_sample_point(const GridDev& g, const float* __restrict__ basis_t_s, const float p[3],
                                                float ft[kAppDim]) {
  float n[3];
  normalize_pt(g, p, n);
  Taps2 pt[3];
  Taps1 lt[3];
  point_taps(g, n, pt, lt);
#pragma unroll
  for (int j = 0; j < kAppDim; ++j) ft[j] = 0.f;
  int cbase = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int C = (i == 0) ? 64 : 16;
    const T* pl = reinterpret_cast<const T*>(g.plane[i]);
    const T* ln = reinterpret_cast<const T*>(g.line[i]);
#pragma unroll 4
    for (int c = 0; c < C; c += 4) {
      const float4 gq = gather4<T>(pl, ln, C, c, pt[i], lt[i]);
      basis_accum4(basis_t_s, cbase + c, gq, ft);
    }
    cbase += C;
  }
}

// sigma -> alpha compositing of one ray whose per-sample (sigma_raw, rgb) already sit in shared memory
// (voxnerf.py:153-201).  Called by ONE thread per ray: the transmittance product runs sequentially, sample by
// sample, like torch.cumprod.  Returns the ray sums through out[5] = {r,g,b,depth,acc}; weights -> w_s[S].
__device__ __forceinline__ void composite_ray(const float* __restrict__ sig_s, const float* __restrict__ rgb_s /*[S][3]*/,
                                              const float* __restrict__ z_s, const float* __restrict__ noise,
                                              int S, float dnorm, bool mask_near, float near_thr, bool relu_rgb,
                                              float* __restrict__ w_s, float out[5]) {
  float T = 1.0f, r = 0.f, gg = 0.f, b = 0.f, dep = 0.f, acc = 0.f;
  for (int s = 0; s < S; ++s) {
    float alpha;
    if (s < S - 1) {
      const float dist = __fmul_rn(z_s[s + 1] - z_s[s], dnorm);
      float sg = sig_s[s];
      if (noise) sg = sg + noise[s];
      sg = fmaxf(sg, 0.0f);
      if (mask_near && !(z_s[s + 1] > near_thr)) sg = 0.0f;
      alpha = 1.0f - expf(-__fmul_rn(sg, dist));
    } else {
      alpha = 1.0f;  // voxnerf.py:189
    }
    const float w = alpha * T;
    w_s[s] = w;
    float cr = rgb_s[3 * s + 0], cg = rgb_s[3 * s + 1], cb = rgb_s[3 * s + 2];
    if (relu_rgb) { cr = fmaxf(cr, 0.f); cg = fmaxf(cg, 0.f); cb = fmaxf(cb, 0.f); }
    r = fmaf(w, cr, r); gg = fmaf(w, cg, gg); b = fmaf(w, cb, b);
    dep = fmaf(w, z_s[s], dep);
    acc += w;
    T = T * ((1.0f + 1e-10f) - alpha);  // cumprod of (1 - alpha + 1e-10); 1 + 1e-10 == 1 in fp32, as in torch
  }
  out[0] = r; out[1] = gg; out[2] = b; out[3] = dep; out[4] = acc;
}

// Two-phase variant used by the tensor-core kernels: every row thread computes its own alpha in parallel ...
__device__ __forceinline__ float alpha_of_sample(float sig_raw, float z, float z_next, float noise, float dnorm, bool mask_near,
                                                 float near_thr, bool is_last) {
  if (is_last) return 1.0f;                                 // voxnerf.py:189
  const float dist = __fmul_rn(z_next - z, dnorm);
  float sg = fmaxf(sig_raw + noise, 0.0f);
  if (mask_near && !(z_next > near_thr)) sg = 0.0f;
  return 1.0f - expf(-__fmul_rn(sg, dist));
}
// ... and ONE thread per ray runs the (cheap) sequential transmittance product and the ray sums, in sample order.
__device__ __forceinline__ void composite_from_alpha(const float* __restrict__ alpha_s, const float* __restrict__ rgb_s /*[S][3]*/,
                                                     const float* __restrict__ z_s, int S, bool relu_rgb, float* __restrict__ w_s,
                                                     float out[5]) {
  float T = 1.0f, r = 0.f, gg = 0.f, b = 0.f, dep = 0.f, acc = 0.f;
#pragma unroll 4
  for (int s = 0; s < S; ++s) {
    const float alpha = alpha_s[s];
    const float w = alpha * T;
    w_s[s] = w;
    float cr = rgb_s[3 * s + 0], cg = rgb_s[3 * s + 1], cb = rgb_s[3 * s + 2];
    if (relu_rgb) { cr = fmaxf(cr, 0.f); cg = fmaxf(cg, 0.f); cb = fmaxf(cb, 0.f); }
    r = fmaf(w, cr, r); gg = fmaf(w, cg, gg); b = fmaf(w, cb, b);
    dep = fmaf(w, z_s[s], dep);
    acc += w;
    T = T * (1.0f - alpha);
  }
  out[0] = r; out[1] = gg; out[2] = b; out[3] = dep; out[4] = acc;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace edn

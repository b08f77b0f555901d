// Backward of the adaptive weight proposal (awp.cu; networks/dpnerf/awp.py:79-117, 49-77, mam.py:13-84), train-mode
// BatchNorm included: d ccw [N][E] -> gradients of every AWP parameter, of depth_feature (which flows on into the fine
// field through edn_render_field_bwd's d_feat), of the ray directions and of the view latent.
//
// Same recipe as field_bwd.cu: the forward is recomputed into the workspace (GEMM path of awp.cu, which keeps the four
// layer activations); the per-sample / per-sub-ray MLPs are back-propagated with plain tall cuBLAS GEMMs; everything else
// is hand-written: the output head + BatchNorm backward (two-phase: batch sums, then per ray), the per-primary-ray
// attention backward (softmax over exposures and over samples, the five 1x1 convs, the two bmm's), the feature-integration
// backward with its cumprod over CHANNELS as a division-free suffix recursion, the view-vector backward.
#include "awp_layout.cuh"
#include "bwd_common.cuh"

namespace edn {
namespace {

constexpr int kMaxE = 16, kMaxS = 256;

struct BwdArgs {
  edn_awp_params p;
  edn_awp_grads g;
  AwpWs ws;
  const float* z_vals;
  const float* rays_d;
  int rays_d_stride;
  const float* view_feature;
  int64_t N;
  int E, S;
  float bn_eps;
  double bn_rows;       // rows behind the BatchNorm batch sums (N * E over all ranks)
  const float* d_ccw;   // [N][E]
  float* d_yn;          // [N][E][32]  gradient at the BatchNorm output
  float* d_x;           // [N][E][32]  gradient at the motion features (residual path, then + attention path)
  double* bn_sums;      // [32][2]     sum d_yn, sum d_yn * y_hat
  float* d_xl;          // [NE][S][32]
  float* d_gint;        // [NE][64] (view into dIN)
  float* d_h;           // [NE*S][64]
  float* d_rays_d;
  int d_rays_d_stride;
  float* d_view_feature;
};

// ---- phase 1: output head, leaky ReLU, residual; BatchNorm batch sums ----------------------------------------------------
__global__ void __launch_bounds__(128) awp_out_bwd_kernel(const BwdArgs a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int E = a.E, lane = threadIdx.x & 31;
  const double rows = a.bn_rows > 0.0 ? a.bn_rows : a.ws.stats[64];
  float s1[32], s2[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
  if (n < a.N) {
    float pooled[32], w[kMaxE], dt[kMaxE], tot = 0.f, dot = 0.f;
    for (int c = 0; c < 32; ++c) {
      const double mean = a.ws.stats[2 * c] / rows;
      const double var = a.ws.stats[2 * c + 1] / rows - mean * mean;
      const float inv = rsqrtf((float)var + a.bn_eps), mu = (float)mean;
      float acc = 0.f;
      for (int e = 0; e < E; ++e) {
        const float yn = (a.ws.y[(n * E + e) * 32 + c] - mu) * inv * a.p.bn_weight[c] + a.p.bn_bias[c];
        const float v = a.ws.x[(n * E + e) * 32 + c] + yn;
        acc += v > 0.f ? v : 0.2f * v;
      }
      pooled[c] = acc / (float)E;
    }
    for (int e = 0; e < E; ++e) {
      float t = a.p.w_linear_b[e];
      for (int c = 0; c < 32; ++c) t = fmaf(a.p.w_linear_w[e * 32 + c], pooled[c], t);
      w[e] = sigmoidf_(t);
      tot += w[e];
    }
    for (int e = 0; e < E; ++e) dot = fmaf(a.d_ccw[n * E + e], w[e] / tot, dot);     // sum_e d_ccw_e ccw_e
    for (int e = 0; e < E; ++e) {
      const float dw = (a.d_ccw[n * E + e] - dot) / tot;                              // ccw_e = w_e / tot
      dt[e] = dw * w[e] * (1.f - w[e]);
      atomicAdd(a.g.w_linear_b + e, dt[e]);
    }
    for (int c = 0; c < 32; ++c) {
      float dp = 0.f;
      for (int e = 0; e < E; ++e) { dp = fmaf(dt[e], a.p.w_linear_w[e * 32 + c], dp); atomicAdd(a.g.w_linear_w + e * 32 + c, dt[e] * pooled[c]); }
      dp /= (float)E;
      const double mean = a.ws.stats[2 * c] / rows;
      const double var = a.ws.stats[2 * c + 1] / rows - mean * mean;
      const float inv = rsqrtf((float)var + a.bn_eps), mu = (float)mean;
      for (int e = 0; e < E; ++e) {
        const float yh = (a.ws.y[(n * E + e) * 32 + c] - mu) * inv;
        const float v = a.ws.x[(n * E + e) * 32 + c] + yh * a.p.bn_weight[c] + a.p.bn_bias[c];
        const float dv = v > 0.f ? dp : 0.2f * dp;
        a.d_x[(n * E + e) * 32 + c] = dv;
        a.d_yn[(n * E + e) * 32 + c] = dv;
        s1[c] += dv;
        s2[c] = fmaf(dv, yh, s2[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], o); s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], o); }
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 32; ++c) { atomicAdd(a.bn_sums + 2 * c, (double)s1[c]); atomicAdd(a.bn_sums + 2 * c + 1, (double)s2[c]); }
  }
}

// d gamma = sum d_yn * y_hat, d beta = sum d_yn
__global__ void awp_bn_param_grad_kernel(const BwdArgs a) {
  const int c = threadIdx.x;
  if (c < 32) { atomicAdd(a.g.bn_bias + c, (float)a.bn_sums[2 * c]); atomicAdd(a.g.bn_weight + c, (float)a.bn_sums[2 * c + 1]); }
}

// ---- phase 2: per primary ray, attention backward ---------------------------------------------------------------------------
template <int EMAX, int SMAX>
struct RayBwdSmem {
  float x[EMAX][32], dy[EMAX][32], cf[EMAX][32], dcf[EMAX][32];
  float inter[EMAX][32], dinter[EMAX][32];
  float inter_a[16][EMAX], dinter_a[16][EMAX];
  float xlog[EMAX][16], dxlog[EMAX][16];
  float inter_n[EMAX][16], dinter_n[EMAX][16];
  float x_inter[EMAX][EMAX], dlg_inter[EMAX][EMAX];
  float att[EMAX][SMAX];        // logits, later d att
  float pE[EMAX][SMAX];         // softmax over exposures
  float pS[EMAX][SMAX];         // softmax over samples
  float intra[32][SMAX + 1];    // later d intra (padded: read and written both channel-major and sample-major)
  float intra_b[16][SMAX], dintra_b[16][SMAX];
  float x_intra[EMAX][SMAX];    // later d logits
  float intra_l[SMAX][16], dintra_l[SMAX][16];
  float g_convd[32 * 32], g_conva[16 * 32], g_convb[16 * 32], g_convc[16 * 32], g_convn[16 * 16], g_convl[16 * 16], g_latt[32];
  float red[EMAX];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// EMAX / SMAX bound the exposure / sample counts of this instantiation (shared memory: 88 KB for E <= 8, S <= 128 -> two CTAs per
// SM; 200 KB for the largest one)
constexpr int kRayBwdThreads = 512;   // the kernel is a chain of short shared-memory-latency-bound phases: more warps per CTA hide it
template <int EMAX, int SMAX>
__global__ void __launch_bounds__(kRayBwdThreads) awp_ray_bwd_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) float smraw[];
  RayBwdSmem<EMAX, SMAX>& s = *reinterpret_cast<RayBwdSmem<EMAX, SMAX>*>(smraw);
  const int tid = threadIdx.x, E = a.E, S = a.S, warp = tid >> 5, lane = tid & 31;
  const int64_t n = blockIdx.x;
  const double rows = a.bn_rows > 0.0 ? a.bn_rows : a.ws.stats[64];
  // ---- load / recompute the forward intermediates --------------------------------------------------------------------
  for (int i = tid; i < E * 32; i += kRayBwdThreads) {
    const int e = i >> 5, c = i & 31;
    s.x[e][c] = a.ws.x[(n * E + e) * 32 + c];
    s.inter[e][c] = a.ws.inter[(n * E + e) * 32 + c];
    // BatchNorm backward (train mode): dy = gamma * inv * (d_yn - mean(d_yn) - y_hat * mean(d_yn * y_hat))
    const double mean = a.ws.stats[2 * c] / rows;
    const double var = a.ws.stats[2 * c + 1] / rows - mean * mean;
    const float inv = rsqrtf((float)var + a.bn_eps), mu = (float)mean;
    const float yh = (a.ws.y[(n * E + e) * 32 + c] - mu) * inv;
    const float m1 = (float)(a.bn_sums[2 * c] / rows), m2 = (float)(a.bn_sums[2 * c + 1] / rows);
    s.dy[e][c] = a.p.bn_weight[c] * inv * (a.d_yn[(n * E + e) * 32 + c] - m1 - yh * m2);
  }
  for (int i = tid; i < E * S; i += kRayBwdThreads) s.att[i / S][i % S] = a.ws.att[(n * E + i / S) * S + i % S];
  for (int i = tid; i < 32 * 32; i += kRayBwdThreads) s.g_convd[i] = 0.f;
  for (int i = tid; i < 16 * 32; i += kRayBwdThreads) { s.g_conva[i] = 0.f; s.g_convb[i] = 0.f; s.g_convc[i] = 0.f; }
  for (int i = tid; i < 16 * 16; i += kRayBwdThreads) { s.g_convn[i] = 0.f; s.g_convl[i] = 0.f; }
  if (tid < 32) s.g_latt[tid] = 0.f;
  __syncthreads();
  // softmax over exposures (per sample) and "intra"
  for (int sp = tid; sp < S; sp += kRayBwdThreads) {
    float mx = -INFINITY, pe[EMAX], sum = 0.f;
    for (int e = 0; e < E; ++e) { pe[e] = s.att[e][sp]; mx = fmaxf(mx, pe[e]); }
    for (int e = 0; e < E; ++e) { pe[e] = expf(pe[e] - mx); sum += pe[e]; }
    for (int e = 0; e < E; ++e) s.pE[e][sp] = pe[e] / sum;
  }
  __syncthreads();
  for (int i = tid; i < 32 * S; i += kRayBwdThreads) {        // lanes = channels: coalesced reads of xl
    const int c = i & 31, sp = i >> 5;
    float acc = 0.f;
    for (int e = 0; e < E; ++e) acc = fmaf(a.ws.xl[((n * E + e) * S + sp) * 32 + c], s.pE[e][sp], acc);
    s.intra[c][sp] = acc;
  }
  // softmax over samples (per exposure), one warp per exposure row
  for (int e = warp; e < E; e += kRayBwdThreads / 32) {
    float mx = -INFINITY, sum = 0.f;
    for (int sp = lane; sp < S; sp += 32) mx = fmaxf(mx, s.att[e][sp]);
    mx = warp_max(mx);
    for (int sp = lane; sp < S; sp += 32) { const float v = expf(s.att[e][sp] - mx); s.pS[e][sp] = v; sum += v; }
    sum = warp_sum(sum);
    for (int sp = lane; sp < S; sp += 32) s.pS[e][sp] /= sum;
  }
  for (int i = tid; i < 16 * E; i += kRayBwdThreads) {
    const int k = i / E, e = i % E;
    float acc = 0.f, acc2 = 0.f;
    for (int c = 0; c < 32; ++c) { acc = fmaf(a.p.conva[k * 32 + c], s.inter[e][c], acc); acc2 = fmaf(a.p.convc[k * 32 + c], s.x[e][c], acc2); }
    s.inter_a[k][e] = acc;
    s.xlog[e][k] = acc2;
  }
  __syncthreads();
  for (int i = tid; i < 16 * S; i += kRayBwdThreads) {
    const int k = i / S, sp = i % S;
    float acc = 0.f;
    for (int c = 0; c < 32; ++c) acc = fmaf(a.p.convb[k * 32 + c], s.intra[c][sp], acc);
    s.intra_b[k][sp] = acc;
  }
  if (tid < E) {
    const int e = tid;
    float lg[EMAX], mx = -INFINITY, sum = 0.f;
    for (int e2 = 0; e2 < E; ++e2) {
      float t = 0.f;
      for (int k = 0; k < 16; ++k) t = fmaf(s.xlog[e][k], s.inter_a[k][e2], t);
      lg[e2] = t; mx = fmaxf(mx, t);
    }
    for (int e2 = 0; e2 < E; ++e2) { lg[e2] = expf(lg[e2] - mx); sum += lg[e2]; }
    for (int e2 = 0; e2 < E; ++e2) s.x_inter[e][e2] = lg[e2] / sum;
    for (int k = 0; k < 16; ++k) {
      float t = 0.f;
      for (int c = 0; c < 16; ++c) t = fmaf(a.p.convn[k * 16 + c], s.inter_a[c][e], t);
      s.inter_n[e][k] = t;
    }
  }
  __syncthreads();
  for (int i = tid; i < E * S; i += kRayBwdThreads) {
    const int e = i / S, sp = i % S;
    float t = 0.f;
    for (int k = 0; k < 16; ++k) t = fmaf(s.xlog[e][k], s.intra_b[k][sp], t);
    s.x_intra[e][sp] = t;
  }
  for (int i = tid; i < S * 16; i += kRayBwdThreads) {
    const int sp = i >> 4, k = i & 15;
    float t = 0.f;
    for (int c = 0; c < 16; ++c) t = fmaf(a.p.convl[k * 16 + c], s.intra_b[c][sp], t);
    s.intra_l[sp][k] = t;
  }
  __syncthreads();
  for (int e = warp; e < E; e += kRayBwdThreads / 32) {
    float mx = -INFINITY, sum = 0.f;
    for (int sp = lane; sp < S; sp += 32) mx = fmaxf(mx, s.x_intra[e][sp]);
    mx = warp_max(mx);
    for (int sp = lane; sp < S; sp += 32) { const float v = expf(s.x_intra[e][sp] - mx); s.x_intra[e][sp] = v; sum += v; }
    sum = warp_sum(sum);
    for (int sp = lane; sp < S; sp += 32) s.x_intra[e][sp] /= sum;
  }
  __syncthreads();
  for (int i = tid; i < E * 32; i += kRayBwdThreads) {
    const int e = i >> 5, j = i & 31;
    float t = 0.f;
    if (j < 16) { for (int e2 = 0; e2 < E; ++e2) t = fmaf(s.x_inter[e][e2], s.inter_n[e2][j], t); }
    else { for (int sp = 0; sp < S; ++sp) t = fmaf(s.x_intra[e][sp], s.intra_l[sp][j - 16], t); }
    s.cf[e][j] = t;
  }
  __syncthreads();
  // ---- backward --------------------------------------------------------------------------------------------------------
  // convd: y[e][c] = sum_k convd[c][k] cf[e][k]
  for (int i = tid; i < E * 32; i += kRayBwdThreads) {
    const int e = i >> 5, k = i & 31;
    float t = 0.f;
    for (int c = 0; c < 32; ++c) t = fmaf(s.dy[e][c], a.p.convd_w[c * 32 + k], t);
    s.dcf[e][k] = t;
  }
  for (int i = tid; i < 32 * 32; i += kRayBwdThreads) {
    const int c = i >> 5, k = i & 31;
    float t = 0.f;
    for (int e = 0; e < E; ++e) t = fmaf(s.dy[e][c], s.cf[e][k], t);
    s.g_convd[i] = t;
  }
  __syncthreads();
  // cf[e][j<16] = sum_e2 x_inter[e][e2] inter_n[e2][j];  cf[e][16+j] = sum_sp x_intra[e][sp] intra_l[sp][j]
  for (int i = tid; i < E * E; i += kRayBwdThreads) {
    const int e = i / E, e2 = i % E;
    float t = 0.f;
    for (int j = 0; j < 16; ++j) t = fmaf(s.dcf[e][j], s.inter_n[e2][j], t);
    s.dlg_inter[e][e2] = t;                                  // d x_inter for now
  }
  for (int i = tid; i < E * 16; i += kRayBwdThreads) {
    const int e2 = i >> 4, j = i & 15;
    float t = 0.f;
    for (int e = 0; e < E; ++e) t = fmaf(s.x_inter[e][e2], s.dcf[e][j], t);
    s.dinter_n[e2][j] = t;
  }
  for (int i = tid; i < S * 16; i += kRayBwdThreads) {
    const int sp = i >> 4, j = i & 15;
    float t = 0.f;
    for (int e = 0; e < E; ++e) t = fmaf(s.x_intra[e][sp], s.dcf[e][16 + j], t);
    s.dintra_l[sp][j] = t;
  }
  __syncthreads();
  // softmax backward of x_inter (rows of E) and x_intra (rows of S; d x_intra computed on the fly)
  if (tid < E) {
    const int e = tid;
    float dot = 0.f;
    for (int e2 = 0; e2 < E; ++e2) dot = fmaf(s.x_inter[e][e2], s.dlg_inter[e][e2], dot);
    for (int e2 = 0; e2 < E; ++e2) s.dlg_inter[e][e2] = s.x_inter[e][e2] * (s.dlg_inter[e][e2] - dot);
  }
  for (int e = warp; e < E; e += kRayBwdThreads / 32) {
    float dot = 0.f;
    for (int sp = lane; sp < S; sp += 32) {
      float t = 0.f;
      for (int j = 0; j < 16; ++j) t = fmaf(s.dcf[e][16 + j], s.intra_l[sp][j], t);     // d x_intra[e][sp]
      s.att[e][sp] = t;                                                                   // att is free now: scratch
      dot = fmaf(s.x_intra[e][sp], t, dot);
    }
    dot = warp_sum(dot);
    for (int sp = lane; sp < S; sp += 32) s.x_intra[e][sp] = s.x_intra[e][sp] * (s.att[e][sp] - dot);   // d logits
  }
  __syncthreads();
  // logits: lg_inter[e][e2] = sum_k xlog[e][k] inter_a[k][e2];  lg_intra[e][sp] = sum_k xlog[e][k] intra_b[k][sp]
  for (int i = tid; i < E * 16; i += kRayBwdThreads) {
    const int e = i >> 4, k = i & 15;
    float t = 0.f;
    for (int e2 = 0; e2 < E; ++e2) t = fmaf(s.dlg_inter[e][e2], s.inter_a[k][e2], t);
    for (int sp = 0; sp < S; ++sp) t = fmaf(s.x_intra[e][sp], s.intra_b[k][sp], t);
    s.dxlog[e][k] = t;
  }
  for (int i = tid; i < 16 * E; i += kRayBwdThreads) {
    const int k = i / E, e2 = i % E;
    float t = 0.f;
    for (int e = 0; e < E; ++e) t = fmaf(s.dlg_inter[e][e2], s.xlog[e][k], t);
    for (int k2 = 0; k2 < 16; ++k2) t = fmaf(a.p.convn[k2 * 16 + k], s.dinter_n[e2][k2], t);     // inter_n = convn . inter_a
    s.dinter_a[k][e2] = t;
  }
  for (int i = tid; i < 16 * S; i += kRayBwdThreads) {
    const int k = i / S, sp = i % S;
    float t = 0.f;
    for (int e = 0; e < E; ++e) t = fmaf(s.x_intra[e][sp], s.xlog[e][k], t);
    for (int k2 = 0; k2 < 16; ++k2) t = fmaf(a.p.convl[k2 * 16 + k], s.dintra_l[sp][k2], t);      // intra_l = convl . intra_b
    s.dintra_b[k][sp] = t;
  }
  for (int i = tid; i < 16 * 16; i += kRayBwdThreads) {
    const int k2 = i >> 4, c = i & 15;
    float t = 0.f, u = 0.f;
    for (int e = 0; e < E; ++e) t = fmaf(s.dinter_n[e][k2], s.inter_a[c][e], t);
    for (int sp = 0; sp < S; ++sp) u = fmaf(s.dintra_l[sp][k2], s.intra_b[c][sp], u);
    s.g_convn[i] = t;
    s.g_convl[i] = u;
  }
  __syncthreads();
  // 1x1 convs: xlog = convc . x, inter_a = conva . inter, intra_b = convb . intra
  for (int i = tid; i < 16 * 32; i += kRayBwdThreads) {
    const int k = i >> 5, c = i & 31;
    float tc = 0.f, ta = 0.f, tb = 0.f;
    for (int e = 0; e < E; ++e) { tc = fmaf(s.dxlog[e][k], s.x[e][c], tc); ta = fmaf(s.dinter_a[k][e], s.inter[e][c], ta); }
    for (int sp = 0; sp < S; ++sp) tb = fmaf(s.dintra_b[k][sp], s.intra[c][sp], tb);
    s.g_convc[i] = tc; s.g_conva[i] = ta; s.g_convb[i] = tb;
  }
  __syncthreads();          // g_convb read intra before it is overwritten below
  for (int i = tid; i < E * 32; i += kRayBwdThreads) {
    const int e = i >> 5, c = i & 31;
    float tx = 0.f, ti = 0.f;
    for (int k = 0; k < 16; ++k) { tx = fmaf(a.p.convc[k * 32 + c], s.dxlog[e][k], tx); ti = fmaf(a.p.conva[k * 32 + c], s.dinter_a[k][e], ti); }
    a.d_x[(n * E + e) * 32 + c] += tx;        // residual-path gradient was written by awp_out_bwd_kernel
    s.dinter[e][c] = ti;
  }
  for (int i = tid; i < 32 * S; i += kRayBwdThreads) {
    const int c = i / S, sp = i % S;
    float t = 0.f;
    for (int k = 0; k < 16; ++k) t = fmaf(a.p.convb[k * 32 + c], s.dintra_b[k][sp], t);
    s.intra[c][sp] = t;                         // d intra
  }
  __syncthreads();
  // d pE, d pS -> d att (both softmaxes); att scratch holds d pS first, x_intra holds d pE
  for (int r = warp; r < E * S; r += kRayBwdThreads / 32) {      // one warp per (exposure, sample) row, lanes = channels
    const int e = r / S, sp = r % S;
    const float v = a.ws.xl[((n * E + e) * S + sp) * 32 + lane];
    const float dpe = warp_sum(s.intra[lane][sp] * v), dps = warp_sum(s.dinter[e][lane] * v);
    if (lane == 0) { s.x_intra[e][sp] = dpe; s.att[e][sp] = dps; }
  }
  __syncthreads();
  for (int e = warp; e < E; e += kRayBwdThreads / 32) {           // row dot products of the softmax over samples
    float dot = 0.f;
    for (int sp = lane; sp < S; sp += 32) dot = fmaf(s.pS[e][sp], s.att[e][sp], dot);
    dot = warp_sum(dot);
    if (lane == 0) s.red[e] = dot;
  }
  __syncthreads();
  for (int sp = tid; sp < S; sp += kRayBwdThreads) {
    float dotE = 0.f;
    for (int e = 0; e < E; ++e) dotE = fmaf(s.pE[e][sp], s.x_intra[e][sp], dotE);
    for (int e = 0; e < E; ++e)
      s.att[e][sp] = s.pE[e][sp] * (s.x_intra[e][sp] - dotE) + s.pS[e][sp] * (s.att[e][sp] - s.red[e]);      // d att
  }
  __syncthreads();
  // d xl = d intra * pE + d inter * pS + d att * latt;  d latt += d att * xl
  float glatt = 0.f;      // thread c = tid & 31 accumulates channel c over its (e, sp) subset
  for (int i = tid; i < E * S * 32; i += kRayBwdThreads) {
    const int c = i & 31, r = i >> 5, e = r / S, sp = r % S;
    const int64_t off = ((n * E + e) * S + sp) * 32 + c;
    const float da = s.att[e][sp];
    a.d_xl[off] = s.intra[c][sp] * s.pE[e][sp] + s.dinter[e][c] * s.pS[e][sp] + da * a.p.line_conv_att[c];
    glatt = fmaf(da, a.ws.xl[off], glatt);
  }
  float* part = &s.dintra_b[0][0];      // free since the d intra pass; [warp][channel] partials instead of shared-memory atomics
  part[warp * 32 + lane] = glatt;
  __syncthreads();
  if (tid < 32) {
    float t = 0.f;
    for (int w = 0; w < kRayBwdThreads / 32; ++w) t += part[w * 32 + tid];
    s.g_latt[tid] = t;
  }
  __syncthreads();
  for (int i = tid; i < 32 * 32; i += kRayBwdThreads) atomicAdd(a.g.convd_w + i, s.g_convd[i]);
  for (int i = tid; i < 16 * 32; i += kRayBwdThreads) { atomicAdd(a.g.conva + i, s.g_conva[i]); atomicAdd(a.g.convb + i, s.g_convb[i]); atomicAdd(a.g.convc + i, s.g_convc[i]); }
  for (int i = tid; i < 16 * 16; i += kRayBwdThreads) { atomicAdd(a.g.convn + i, s.g_convn[i]); atomicAdd(a.g.convl + i, s.g_convl[i]); }
  if (tid < 32) atomicAdd(a.g.line_conv_att + tid, s.g_latt[tid]);
}

// ---- motion MLP input: IN[r] = [gint (64) | view latent (32) | PE(viewdir of exposure 0, L = 2) (15) | 0] ----------------------
__global__ void awp_motion_input_kernel(const BwdArgs a, float* __restrict__ IN) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = t / 112;
  const int j = (int)(t % 112);
  if (r >= a.N * a.E) return;
  const int64_t n = r / a.E;
  float v = 0.f;
  if (j < 64) v = a.ws.gint[r * 64 + j];
  else if (j < 96) v = a.view_feature[n * 32 + (j - 64)];
  else if (j < 111) {
    const float* rd = a.rays_d + (n * a.E) * a.rays_d_stride;
    const float nr = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
    const int q = j - 96;
    if (q < 3) v = rd[q] / nr;
    else { const int f = (q - 3) / 6, rem = (q - 3) % 6, i = rem % 3; const float x = rd[i] / nr * (float)(1 << f); v = rem < 3 ? sinf(x) : cosf(x); }
  }
  IN[t] = v;
}

// d IN[:, 64:111] summed over the exposures -> d view latent, d rays_d of exposure 0 (PE + normalisation backward)
__global__ void awp_view_bwd_kernel(const BwdArgs a, const float* __restrict__ dIN) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  float g[47];
#pragma unroll
  for (int j = 0; j < 47; ++j) g[j] = 0.f;
  for (int e = 0; e < a.E; ++e) {
    const float* row = dIN + (n * a.E + e) * 112 + 64;
#pragma unroll
    for (int j = 0; j < 47; ++j) g[j] += row[j];
  }
  if (a.d_view_feature) {
#pragma unroll
    for (int j = 0; j < 32; ++j) a.d_view_feature[n * 32 + j] = g[j];
  }
  if (!a.d_rays_d) return;
  const float* rd = a.rays_d + (n * a.E) * a.rays_d_stride;
  const float nr = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
  float dv[3], dotv = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float v = rd[i] / nr;
    float t = g[32 + i];
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const float fr = (float)(1 << f);
      float sn, cs;
      sincosf(v * fr, &sn, &cs);
      t += fr * (cs * g[32 + 3 + 6 * f + i] - sn * g[32 + 6 + 6 * f + i]);
    }
    dv[i] = t;
    dotv = fmaf(t, v, dotv);
  }
  float* out = a.d_rays_d + (n * a.E) * a.d_rays_d_stride;
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] += (dv[i] - dotv * rd[i] / nr) / nr;      // v = d / |d|
}

// d gint [NE][64] = dIN[:, 0:64] (strided) is read in place by the integration backward.
// ---- feature-integration backward (awp.py:49-77), one block per sub-ray, one thread per sample ---------------------------------
//   al = 1 - exp(-h dist), q = 1 - al, Q[s][c] = prod_{c'<=c} q[s][c'], g[c] = sum_s al[s][c] Q[s-1][c] h[s][c]   (Q[-1] = 1)
__global__ void __launch_bounds__(128) awp_integrate_bwd_kernel(const BwdArgs a, const float* __restrict__ h_all, const float* __restrict__ dIN) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, S = a.S;
  float* Hs = sm;                 // [S][65]
  float* Al = Hs + S * 65;        // [S][65]
  float* Q = Al + S * 65;         // [S][65]
  float* G = Q + S * 65;          // [64]
  float* red = G + 64;            // [4]
  const int64_t sr = blockIdx.x;
  const float* h = h_all + sr * S * 64;
  for (int i = tid; i < S * 64; i += 128) Hs[(i >> 6) * 65 + (i & 63)] = h[i];
  if (tid < 64) G[tid] = dIN[sr * 112 + tid];
  __syncthreads();
  const float* rd = a.rays_d + sr * a.rays_d_stride;
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])), __fmul_rn(rd[2], rd[2])));
  for (int s = tid; s < S; s += 128) {
    const bool has = s < S - 1;
    const float dist = has ? __fmul_rn(a.z_vals[sr * S + s + 1] - a.z_vals[sr * S + s], dnorm) : 0.f;
    float p = 1.0f;
    for (int c = 0; c < 64; ++c) {
      const float al = has ? 1.0f - expf(-__fmul_rn(Hs[s * 65 + c], dist)) : 0.f;
      Al[s * 65 + c] = al;
      p *= (1.0f - al);
      Q[s * 65 + c] = p;
    }
  }
  __syncthreads();
  float ddn = 0.f;
  for (int s = tid; s < S; s += 128) {
    const bool has = s < S - 1;
    const float dz = has ? a.z_vals[sr * S + s + 1] - a.z_vals[sr * S + s] : 0.f;
    const float dist = __fmul_rn(dz, dnorm);
    float T = 0.f, ddist = 0.f;
    float* out = a.d_h + (sr * S + s) * 64;
    for (int c = 63; c >= 0; --c) {
      const float al = Al[s * 65 + c], hv = Hs[s * 65 + c], q = 1.0f - al;
      const float dQ = has ? G[c] * Al[(s + 1) * 65 + c] * Hs[(s + 1) * 65 + c] : 0.f;     // Q[s] feeds the weights of sample s + 1
      const float qn = c < 63 ? 1.0f - Al[s * 65 + c + 1] : 0.f;
      T = dQ + qn * T;
      const float pex = c > 0 ? Q[s * 65 + c - 1] : 1.0f;
      const float qp = s > 0 ? Q[(s - 1) * 65 + c] : 1.0f;
      const float dal = G[c] * qp * hv - pex * T;
      out[c] = G[c] * al * qp + dal * dist * q;
      ddist = fmaf(dal * hv, q, ddist);
    }
    ddn = fmaf(ddist, dz, ddn);
  }
  ddn = warp_sum(ddn);
  if ((tid & 31) == 0) red[tid >> 5] = ddn;
  __syncthreads();
  if (tid == 0 && a.d_rays_d && dnorm > 0.f) {
    const float t = red[0] + red[1] + red[2] + red[3];
    float* out = a.d_rays_d + sr * a.d_rays_d_stride;
    for (int i = 0; i < 3; ++i) out[i] += t * rd[i] / dnorm;
  }
}

// S <= 128: four threads per sample (16 channels each); the suffix recursion over the channels runs locally and is stitched
// across the row's four lanes with three dependent shuffles.
__global__ void __launch_bounds__(512, 2) awp_integrate4_bwd_kernel(const BwdArgs a, const float* __restrict__ h_all, const float* __restrict__ dIN) {
  extern __shared__ __align__(16) float sm4[];
  float (*Qs)[64] = reinterpret_cast<float (*)[64]>(sm4);                 // [128][64]
  float (*AH)[64] = reinterpret_cast<float (*)[64]>(sm4 + 128 * 64);      // [129][64]: al * h of each row (row s needs row s + 1); row 128 = 0
  float* G = sm4 + 128 * 64 + 129 * 64;                                   // [64]
  float* red = G + 64;                                                    // [16]
  const int tid = threadIdx.x, S = a.S, r = tid >> 2, q = tid & 3, c0 = q * 16, warp = tid >> 5, lane = tid & 31;
  const int64_t sr = blockIdx.x;
  const bool valid = r < S, has = r < S - 1;
  if (tid < 64) { G[tid] = dIN[sr * 112 + tid]; AH[128][tid] = 0.f; }
  const float* rd = a.rays_d + sr * a.rays_d_stride;
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])), __fmul_rn(rd[2], rd[2])));
  float h[16], al[16], Q[16];
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(h_all + (sr * S + r) * 64 + c0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    h[i] = v.x; h[i + 1] = v.y; h[i + 2] = v.z; h[i + 3] = v.w;
  }
  const float dz = has ? a.z_vals[sr * S + r + 1] - a.z_vals[sr * S + r] : 0.f;
  const float dist = __fmul_rn(dz, dnorm);
  float p = 1.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    al[i] = has ? 1.0f - expf(-__fmul_rn(h[i], dist)) : 0.f;
    p *= (1.0f - al[i]);
    Q[i] = p;
  }
  float e;
  {
    const float t1 = __shfl_up_sync(0xffffffffu, p, 1), t2 = __shfl_up_sync(0xffffffffu, p, 2), t3 = __shfl_up_sync(0xffffffffu, p, 3);
    e = (q >= 1 ? t1 : 1.0f) * (q >= 2 ? t2 : 1.0f) * (q >= 3 ? t3 : 1.0f);
#pragma unroll
    for (int i = 0; i < 16; ++i) Q[i] *= e;
  }
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    *reinterpret_cast<float4*>(&Qs[r][c0 + i]) = make_float4(Q[i], Q[i + 1], Q[i + 2], Q[i + 3]);
    *reinterpret_cast<float4*>(&AH[r][c0 + i]) = make_float4(al[i] * h[i], al[i + 1] * h[i + 1], al[i + 2] * h[i + 2], al[i + 3] * h[i + 3]);
  }
  __syncthreads();
  // dQ[c] = G[c] al[s+1][c] h[s+1][c];  T_c = dQ_c + q_{c+1} T_{c+1}  (T beyond the last channel = 0)
  float dQ[16], T[16], Mc[16];      // T with zero inflow; Mc = d T_c / d inflow
  const float qn = __shfl_down_sync(0xffffffffu, 1.0f - al[0], 1);     // q of the next quarter's first channel
#pragma unroll
  for (int i = 0; i < 16; ++i) dQ[i] = has ? G[c0 + i] * AH[r + 1][c0 + i] : 0.f;
  {
    float t = 0.f, m = (q < 3) ? qn : 0.f;
#pragma unroll
    for (int i = 15; i >= 0; --i) {
      t = dQ[i] + (i < 15 ? (1.0f - al[i + 1]) * t : 0.f);
      T[i] = t;
      Mc[i] = m;
      m *= (1.0f - al[i]);
    }
  }
  // stitch: inflow of quarter q = T at the first channel of quarter q + 1
  float first = T[0];                       // quarter 3: exact already
  float inflow = 0.f;
#pragma unroll
  for (int step = 2; step >= 0; --step) {
    const float nb = __shfl_down_sync(0xffffffffu, first, 1);
    if (q == step) { inflow = nb; first = T[0] + Mc[0] * nb; }
  }
  float ddist = 0.f;
  float out[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float Tc = T[i] + Mc[i] * inflow;
    const float pex = i > 0 ? Q[i - 1] : e;
    const float qp = r > 0 ? Qs[r - 1][c0 + i] : 1.0f;
    const float qv = 1.0f - al[i];
    const float dal = G[c0 + i] * qp * h[i] - pex * Tc;
    out[i] = G[c0 + i] * al[i] * qp + dal * dist * qv;
    ddist = fmaf(dal * h[i], qv, ddist);
  }
  if (valid) {
#pragma unroll
    for (int i = 0; i < 16; i += 4)
      *reinterpret_cast<float4*>(a.d_h + (sr * S + r) * 64 + c0 + i) = make_float4(out[i], out[i + 1], out[i + 2], out[i + 3]);
  }
  float ddn = valid ? ddist * dz : 0.f;
  ddn = warp_sum(ddn);
  if (lane == 0) red[warp] = ddn;
  __syncthreads();
  if (tid == 0 && a.d_rays_d && dnorm > 0.f) {
    float t = 0.f;
    for (int w = 0; w < 16; ++w) t += red[w];
    float* o = a.d_rays_d + sr * a.d_rays_d_stride;
    for (int i = 0; i < 3; ++i) o[i] += t * rd[i] / dnorm;
  }
}

__global__ void fold_w0_kernel(const float* __restrict__ src, float* __restrict__ dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 32 * 111) return;
  dst[t] += src[(t / 111) * 112 + t % 111];
}

}  // namespace
}  // namespace edn

extern "C" int64_t edn_awp_bwd_workspace_floats(int64_t n_rays, int32_t n_exposure, int32_t n_samples) {
  using namespace edn;
  if (n_rays < 0 || n_exposure < 1 || n_samples < 2) return -1;
  const int64_t NE = n_rays * n_exposure, M = NE * n_samples;
  return awp_ws_floats(n_rays, n_exposure, n_samples, true) + 16 + NE + 2 * NE * 32 /* ccw, d_yn, d_x */ + 2 * 64 + 4 /* bn sums (doubles) */ +
         M * 32 /* d_xl */ + 2 * M * 64 /* d_h, chain ping-pong */ + 2 * NE * 112 /* IN, dIN */ + 2 * NE * 32 /* H0, dH */ + 32 * 112 * 2 /* padded W0 + grad */;
}

extern "C" int64_t edn_awp_bwd_sums_offset_floats(int64_t n_rays, int32_t n_exposure, int32_t n_samples) {
  float* zero = nullptr;
  const edn::AwpBwdHead h = edn::awp_bwd_head(zero, n_rays, n_exposure, n_samples);
  return reinterpret_cast<float*>(h.bn_sums) - zero;
}

extern "C" int edn_awp_bwd(const edn_awp_params* p, const float* depth_feature, const float* z_vals, const float* rays_d,
                           int32_t rays_d_stride, const float* view_feature, int64_t n_rays, int32_t n_exposure, int32_t n_samples,
                           float bn_eps, const edn_awp_options* opt, int32_t forward_in_workspace, const float* d_ccw, const edn_awp_grads* grads,
                           float* d_depth_feature, float* d_rays_d, int32_t d_rays_d_stride, float* d_view_feature, float* workspace,
                           void* stream) {
  using namespace edn;
  EDN_REQUIRE(p && depth_feature && z_vals && rays_d && view_feature && d_ccw && grads && d_depth_feature && workspace, "edn_awp_bwd: null pointer");
  EDN_REQUIRE(n_exposure >= 1 && n_exposure <= kMaxE && n_samples >= 2 && n_samples <= kMaxS, "edn_awp_bwd: need 1 <= E <= %d and 2 <= S <= %d", kMaxE, kMaxS);
  EDN_REQUIRE(opt && (opt->precision == EDN_F32 || opt->precision == EDN_BF16) && opt->phase >= 0 && opt->phase <= 2, "edn_awp_bwd: bad options");
  EDN_REQUIRE(opt->phase == 0 || forward_in_workspace, "edn_awp_bwd: phased (synchronised BatchNorm) backward needs the forward's workspace");
  const int precision = opt->precision, phase = opt->phase;
  for (int l = 0; l < 4; ++l) EDN_REQUIRE(grads->sample_t[l] && grads->sample_b[l], "edn_awp_bwd: null gradient buffer");
  EDN_REQUIRE(grads->motion_w[0] && grads->motion_b[0] && grads->motion_w[1] && grads->motion_b[1] && grads->mam_linear_t && grads->mam_linear_b &&
              grads->line_conv_att && grads->conva && grads->convb && grads->convc && grads->convn && grads->convl && grads->convd_w &&
              grads->bn_weight && grads->bn_bias && grads->w_linear_w && grads->w_linear_b, "edn_awp_bwd: null gradient buffer");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t N = n_rays, NE = n_rays * n_exposure, M = NE * n_samples;
  const int E = n_exposure, S = n_samples;
  const int in_ch = p->input_ch > 0 ? p->input_ch : 128;
  const bool tf32 = precision == EDN_BF16;
  const AwpBwdHead head = awp_bwd_head(workspace, N, E, S);
  float* base = head.next;
  auto take = [&](int64_t n) { float* q = base; base += (n + 3) / 4 * 4; return q; };
  if (!forward_in_workspace) {      // recompute the forward (ccw itself is not needed: scratch)
    int rc = awp_forward(p, depth_feature, z_vals, rays_d, rays_d_stride, view_feature, N, E, S, bn_eps, true, tf32, 0, 0, workspace, head.ccw_tmp, stream);
    if (rc) return rc;
  }
  BwdArgs a{};
  a.p = *p; a.g = *grads; a.ws = awp_ws_carve(workspace, N, E, S, true);
  a.z_vals = z_vals; a.rays_d = rays_d; a.rays_d_stride = rays_d_stride; a.view_feature = view_feature;
  a.N = N; a.E = E; a.S = S; a.bn_eps = bn_eps; a.d_ccw = d_ccw;
  a.bn_rows = (double)(opt->bn_rows_total > 0 ? opt->bn_rows_total : (phase == 0 ? NE : 0));   // 0: the forward's all-reduced count (ws.stats[64])
  a.d_rays_d = d_rays_d; a.d_rays_d_stride = d_rays_d_stride; a.d_view_feature = d_view_feature;
  a.d_yn = head.d_yn;
  a.d_x = head.d_x;
  a.bn_sums = head.bn_sums;
  a.d_xl = take(M * 32);
  a.d_h = take(M * 64);
  float* Dn = take(M * 64);
  float* IN = take(NE * 112);
  float* dIN = take(NE * 112);
  float* H0 = take(NE * 32);
  float* dH = take(NE * 32);
  float* W0p = take(32 * 112);
  float* gW0p = take(32 * 112);

  cublasHandle_t h = blas_handle();
  if (!h) return blas_unavailable();
  if (cublasSetStream(h, st) != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetStream failed"); return EDN_E_CUDA; }
  const Gemm gemm{h, tf32 ? CUBLAS_COMPUTE_32F_FAST_TF32 : CUBLAS_COMPUTE_32F};
#define EDN_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
  // 1. output head + BatchNorm sums (d gamma / d beta take the LOCAL sums; the per-ray backward below the all-reduced ones)
  if (phase != 2) {
    EDN_CUDA_OK(cudaMemsetAsync(a.bn_sums, 0, 64 * sizeof(double), st));
    awp_out_bwd_kernel<<<blocks_for(N, 128), 128, 0, st>>>(a);
    awp_bn_param_grad_kernel<<<1, 32, 0, st>>>(a);
    EDN_CUDA_OK(cudaGetLastError());
    if (phase == 1) return EDN_OK;
  }
  // 2. per-ray attention backward
  {
    auto launch = [&](auto kern, size_t smem) -> int {
      EDN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<(unsigned)N, kRayBwdThreads, smem, st>>>(a);
      return 0;
    };
    int rc2;
    if (E <= 8 && S <= 64) rc2 = launch(awp_ray_bwd_kernel<8, 64>, sizeof(RayBwdSmem<8, 64>));
    else if (E <= 8 && S <= 128) rc2 = launch(awp_ray_bwd_kernel<8, 128>, sizeof(RayBwdSmem<8, 128>));
    else rc2 = launch(awp_ray_bwd_kernel<kMaxE, kMaxS>, sizeof(RayBwdSmem<kMaxE, kMaxS>));
    if (rc2) return rc2;
  }
  // 3. motion MLP (awp.py:104-108) over the N*E sub-rays: x = relu(W1 relu(W0 IN + b0) + b1)
  awp_motion_input_kernel<<<blocks_for(NE * 112, 256), 256, 0, st>>>(a, IN);
  {   // zero-padded copy of motion_w[0] [32][111] -> [32][112]
    EDN_CUDA_OK(cudaMemsetAsync(W0p, 0, sizeof(float) * 32 * 112, st));
    EDN_CUDA_OK(cudaMemcpy2DAsync(W0p, 112 * sizeof(float), p->motion_w[0], 111 * sizeof(float), 111 * sizeof(float), 32, cudaMemcpyDeviceToDevice, st));
    EDN_CUDA_OK(cudaMemsetAsync(gW0p, 0, sizeof(float) * 32 * 112, st));
  }
  EDN_RC(gemm(false, true, NE, 32, 112, IN, 112, W0p, 112, 0.f, H0, 32));
  relu_bias_kernel<<<blocks_for(NE * 8, 256), 256, 0, st>>>(H0, 32, 32, NE, p->motion_b[0]);
  relu_mask_kernel<<<blocks_for(NE * 8, 256), 256, 0, st>>>(a.d_x, a.ws.x, 32, 32, NE);                  // d pre-activation of layer 1
  EDN_RC(gemm(true, false, 32, 32, NE, a.d_x, 32, H0, 32, 1.f, grads->motion_w[1], 32));
  colsum_kernel<<<blocks_for(NE, 512), 64, 0, st>>>(a.d_x, 32, 32, NE, grads->motion_b[1]);
  EDN_RC(gemm(false, false, NE, 32, 32, a.d_x, 32, p->motion_w[1], 32, 0.f, dH, 32));
  relu_mask_kernel<<<blocks_for(NE * 8, 256), 256, 0, st>>>(dH, H0, 32, 32, NE);
  EDN_RC(gemm(true, false, 32, 112, NE, dH, 32, IN, 112, 1.f, gW0p, 112));
  colsum_kernel<<<blocks_for(NE, 512), 64, 0, st>>>(dH, 32, 32, NE, grads->motion_b[0]);
  EDN_RC(gemm(false, false, NE, 112, 32, dH, 32, W0p, 112, 0.f, dIN, 112));
  awp_view_bwd_kernel<<<blocks_for(N, 128), 128, 0, st>>>(a, dIN);
  // 4. feature-integration backward -> d_h (=), d rays_d
  if (S <= 128) {
    const size_t smem4 = sizeof(float) * (128 * 64 + 129 * 64 + 64 + 16);
    EDN_CUDA_OK(cudaFuncSetAttribute(awp_integrate4_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
    awp_integrate4_bwd_kernel<<<(unsigned)NE, 512, smem4, st>>>(a, a.ws.act[3], dIN);
  } else {
    const size_t smem_i = sizeof(float) * (size_t)(S * 3 * 65 + 64 + 4);
    EDN_CUDA_OK(cudaFuncSetAttribute(awp_integrate_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i));
    awp_integrate_bwd_kernel<<<(unsigned)NE, 128, smem_i, st>>>(a, a.ws.act[3], dIN);
  }
  // 5. MAM.linear (mam.py:75): xl = h Wm_t + bm
  EDN_RC(gemm(false, true, M, 64, 32, a.d_xl, 32, p->mam_linear_t, 32, 1.f, a.d_h, 64));
  EDN_RC(gemm(true, false, 64, 32, M, a.ws.act[3], 64, a.d_xl, 32, 1.f, grads->mam_linear_t, 32));
  colsum_kernel<<<blocks_for(M, 512), 64, 0, st>>>(a.d_xl, 32, 32, M, grads->mam_linear_b);
  // 6. sample MLP (awp.py:96-98), layers 3..0
  float* D = a.d_h;
  float* Dnext = Dn;
  for (int l = 3; l >= 0; --l) {
    const float* X = l > 0 ? a.ws.act[l - 1] : depth_feature;
    const int K = l > 0 ? 64 : in_ch;
    relu_mask_kernel<<<blocks_for(M * 16, 256), 256, 0, st>>>(D, a.ws.act[l], 64, 64, M);
    EDN_RC(gemm(true, false, K, 64, M, X, K, D, 64, 1.f, grads->sample_t[l], 64));
    colsum_kernel<<<blocks_for(M, 512), 64, 0, st>>>(D, 64, 64, M, grads->sample_b[l]);
    float* out = l > 0 ? Dnext : d_depth_feature;
    EDN_RC(gemm(false, true, M, K, 64, D, 64, p->sample_t[l], 64, 0.f, out, K));
    Dnext = D;
    D = out;
  }
#undef EDN_RC
  fold_w0_kernel<<<(32 * 111 + 255) / 256, 256, 0, st>>>(gW0p, grads->motion_w[0]);     // [32][112] -> += [32][111]
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

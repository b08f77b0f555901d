// Deformable sparse kernel (DSK) blur model: BlurModel.forward with kernel_type = DSK (networks/pdrf/blurmodel.py:109-224) and its
// backward.  Rows = N rays x n_pt kernel points (20 480 for the headline batch), a 64-wide MLP: a few hundred MFLOP, so the
// contractions are plain fp32 GEMMs (cuBLAS, exact fp32) between hand-written input / head kernels -- the same recipe as the RBK
// head backward (ray_bwd.cu).  Forward and backward share `dsk_recompute` (the backward keeps nothing from the forward).
//
//   x   = [PE(tanh(pattern_pos) * hw (+ noise), scaled by pi / hw) | img_embed[idx] | PE(pixel position)]        blurmodel.py:119-155
//   h   = relu(linears.{2l} h) x num_hidden;  o = linears1.2 relu(linears1.0 [x |] h)                              blurmodel.py:164-166
//   o   = [delta_trans (2, optim_sv_trans) | delta_pos (2) | weight logit]                                         blurmodel.py:168-172
//   new_xy = delta_pos + input_pos; weight = softmax over the points; rays through the offset pixels               blurmodel.py:183-218
#include "bwd_common.cuh"

namespace edn {
namespace {

constexpr float kPi = 3.14159265358979323846f;

struct DskDims {
  int P, L_in, L_sp, pe_in, pe_sp, embed, in_cnl, wide, cat_w, nh, oc;
};
inline DskDims dsk_dims(const edn_dsk_params* p) {
  DskDims d;
  d.P = p->n_pt; d.L_in = p->in_embed; d.L_sp = p->spatial_embed; d.embed = p->embed; d.wide = p->wide; d.nh = p->num_hidden;
  d.pe_in = 2 + 4 * d.L_in;
  d.pe_sp = d.L_sp > 0 ? 2 + 4 * d.L_sp : 0;
  d.in_cnl = d.pe_in + d.embed + d.pe_sp;
  d.cat_w = d.in_cnl + d.wide;
  d.oc = p->optim_sv_trans ? 5 : 3;
  return d;
}
inline int64_t dsk_floats_per_row(const DskDims& d) { return 2 * (int64_t)d.cat_w + (int64_t)(d.nh + 4) * d.wide + 32; }

// [x, sin(2^0 x), cos(2^0 x), ...] for a 2-vector (embedding.py:88-98 with input_dims = 2)
__device__ __forceinline__ void pe2(const float x[2], int L, float* out) {
  out[0] = x[0]; out[1] = x[1];
  for (int f = 0; f < L; ++f) {
    const float fr = (float)(1 << f);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float sn, cs;
      sincosf(x[i] * fr, &sn, &cs);
      out[2 + 4 * f + i] = sn;
      out[4 + 4 * f + i] = cs;
    }
  }
}

// one thread per (ray, point): canonical position -> input row X[m][0..in_cnl) (row stride ld), input_pos[m][2]
__global__ void dsk_input_kernel(const edn_dsk_params p, const DskDims d, const float* __restrict__ rays_x, const float* __restrict__ rays_y,
                                 const int64_t* __restrict__ idx, const float* __restrict__ noise, int64_t N, int H, int W,
                                 float* __restrict__ X, int ld, float* __restrict__ input_pos) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= N * d.P) return;
  const int64_t n = m / d.P;
  const int k = (int)(m % d.P);
  const int64_t img = idx[n];
  const float* pp = p.pattern_pos + ((p.isglobal ? 0 : img) * d.P + k) * 2;
  float pos[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    pos[i] = tanhf(pp[i]) * p.kernel_hwindow;
    if (noise) pos[i] += noise[m * 2 + i];
    input_pos[m * 2 + i] = pos[i];
  }
  float* row = X + m * ld;
  if (d.L_in > 0) {
    const float sc[2] = {pos[0] * (kPi / p.kernel_hwindow), pos[1] * (kPi / p.kernel_hwindow)};
    pe2(sc, d.L_in, row);
  } else {
    row[0] = pos[0]; row[1] = pos[1];
  }
  const float* e = p.img_embed + img * d.embed;
  for (int j = 0; j < d.embed; ++j) row[d.pe_in + j] = e[j];
  if (d.L_sp > 0) {
    const float sp[2] = {rays_x[n] / ((float)W / 2.f / kPi) - kPi, rays_y[n] / ((float)H / 2.f / kPi) - kPi};
    pe2(sp, d.L_sp, row + d.pe_in + d.embed);
  }
}

// scalar elementwise helpers (row strides here are not multiples of 4: no vector access)
__global__ void dsk_bias_act_kernel(float* __restrict__ Y, int ld, int n, int64_t M, const float* __restrict__ bias, int relu) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * n) return;
  const int64_t m = t / n;
  const int j = (int)(t % n);
  const float v = Y[m * ld + j] + bias[j];
  Y[m * ld + j] = relu ? fmaxf(v, 0.f) : v;
}
__global__ void dsk_relu_mask_kernel(float* __restrict__ D, int ldd, const float* __restrict__ Hh, int ldh, int n, int64_t M) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * n) return;
  const int64_t m = t / n;
  const int j = (int)(t % n);
  if (!(Hh[m * ldh + j] > 0.f)) D[m * ldd + j] = 0.f;
}
__global__ void dsk_colsum_kernel(const float* __restrict__ D, int ld, int n, int64_t M, float* __restrict__ out) {
  const int j = threadIdx.x;
  if (j >= n) return;
  const int64_t r0 = (int64_t)blockIdx.x * 256, r1 = min(r0 + 256, M);
  float acc = 0.f;
  for (int64_t m = r0; m < r1; ++m) acc += D[m * ld + j];
  atomicAdd(out + j, acc);
}

struct HeadGeom { float fx, fy, cx, cy; };

// one thread per ray: MLP outputs of its n_pt points -> softmax weights, rays, per-ray share of the align term
__global__ void dsk_head_kernel(const edn_dsk_params p, const DskDims d, const float* __restrict__ O, const float* __restrict__ input_pos,
                                const float* __restrict__ rays_x, const float* __restrict__ rays_y, const int64_t* __restrict__ idx,
                                const float* __restrict__ poses, int64_t N, HeadGeom g, float* __restrict__ new_rays,
                                float* __restrict__ weight, float* __restrict__ align_part) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* pose = poses + n * 12;
  const int64_t img = idx[n];
  float mx = -INFINITY;
  for (int k = 0; k < d.P; ++k) mx = fmaxf(mx, O[(n * d.P + k) * 8 + d.oc - 1]);
  float tot = 0.f;
  for (int k = 0; k < d.P; ++k) tot += expf(O[(n * d.P + k) * 8 + d.oc - 1] - mx);
  for (int k = 0; k < d.P; ++k) {
    const int64_t m = n * d.P + k;
    const float* o = O + m * 8;
    float dt[2] = {0.f, 0.f};
    if (p.optim_sv_trans) { dt[0] = o[0]; dt[1] = o[1]; }
    if (p.pattern_trans) {
      const float* pt = p.pattern_trans + ((p.isglobal ? 0 : img) * d.P + k) * 2;
      dt[0] = pt[0]; dt[1] = pt[1];
    }
    dt[0] *= 0.01f; dt[1] *= 0.01f;
    const float* dp = o + (p.optim_sv_trans ? 2 : 0);
    const float nx = dp[0] + input_pos[m * 2], ny = dp[1] + input_pos[m * 2 + 1];
    if (k == 0 && align_part)
      align_part[n] = (fabsf(nx) + fabsf(ny)) / (2.f * (float)N) + 10.f * (fabsf(dt[0]) + fabsf(dt[1])) / (2.f * (float)N);
    weight[m] = expf(o[d.oc - 1] - mx) / tot;
    const float rx = (rays_x[n] - g.cx + nx) / g.fx;
    const float ry = -(rays_y[n] - g.cy + ny) / g.fy;
    const float dir[3] = {rx - dt[0], ry - dt[1], -1.f};
    float* out = new_rays + m * 6;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      out[2 * i] = dt[0] * pose[i * 4] + dt[1] * pose[i * 4 + 1] + 0.f * pose[i * 4 + 2] + pose[i * 4 + 3];
      out[2 * i + 1] = dir[0] * pose[i * 4] + dir[1] * pose[i * 4 + 1] + dir[2] * pose[i * 4 + 2];
    }
  }
}

__global__ void dsk_align_reduce_kernel(const float* __restrict__ part, int64_t N, float* __restrict__ align) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += 256) acc += (double)part[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) align[0] = (float)sh[0];
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// head backward, one thread per ray: d new_rays / d weight / d align -> dO [M][8], d input_pos [M][2]; pattern_trans gradient
__global__ void dsk_head_bwd_kernel(const edn_dsk_params p, const DskDims d, const float* __restrict__ O, const float* __restrict__ input_pos,
                                    const int64_t* __restrict__ idx, const float* __restrict__ poses, int64_t N, HeadGeom g,
                                    const float* __restrict__ d_new_rays, const float* __restrict__ d_weight,
                                    const float* __restrict__ d_align, float* __restrict__ dO, float* __restrict__ d_input_pos,
                                    float* __restrict__ g_pattern_trans) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* pose = poses + n * 12;
  const int64_t img = idx[n];
  const float ga = d_align ? d_align[0] : 0.f;
  float mx = -INFINITY;
  for (int k = 0; k < d.P; ++k) mx = fmaxf(mx, O[(n * d.P + k) * 8 + d.oc - 1]);
  float tot = 0.f, dot = 0.f;
  for (int k = 0; k < d.P; ++k) tot += expf(O[(n * d.P + k) * 8 + d.oc - 1] - mx);
  if (d_weight)
    for (int k = 0; k < d.P; ++k) dot = fmaf(d_weight[n * d.P + k], expf(O[(n * d.P + k) * 8 + d.oc - 1] - mx) / tot, dot);
  for (int k = 0; k < d.P; ++k) {
    const int64_t m = n * d.P + k;
    const float* o = O + m * 8;
    float* go = dO + m * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) go[j] = 0.f;
    float dt[2] = {0.f, 0.f};
    if (p.optim_sv_trans) { dt[0] = o[0]; dt[1] = o[1]; }
    if (p.pattern_trans) {
      const float* pt = p.pattern_trans + ((p.isglobal ? 0 : img) * d.P + k) * 2;
      dt[0] = pt[0]; dt[1] = pt[1];
    }
    dt[0] *= 0.01f; dt[1] *= 0.01f;
    const float* dp = o + (p.optim_sv_trans ? 2 : 0);
    const float nx = dp[0] + input_pos[m * 2], ny = dp[1] + input_pos[m * 2 + 1];
    float g_dir[2] = {0.f, 0.f}, g_tr[2] = {0.f, 0.f};
    if (d_new_rays) {
      const float* gr = d_new_rays + m * 6;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        g_dir[0] = fmaf(gr[2 * i + 1], pose[i * 4], g_dir[0]);
        g_dir[1] = fmaf(gr[2 * i + 1], pose[i * 4 + 1], g_dir[1]);
        g_tr[0] = fmaf(gr[2 * i], pose[i * 4], g_tr[0]);
        g_tr[1] = fmaf(gr[2 * i], pose[i * 4 + 1], g_tr[1]);
      }
    }
    float g_xy[2] = {g_dir[0] / g.fx, -g_dir[1] / g.fy};
    float g_dt[2] = {g_tr[0] - g_dir[0], g_tr[1] - g_dir[1]};          // w.r.t. the scaled (x 0.01) origin offsets
    if (k == 0) {
      g_xy[0] += ga * sgn(nx) / (2.f * (float)N); g_xy[1] += ga * sgn(ny) / (2.f * (float)N);
      g_dt[0] += ga * 10.f * sgn(dt[0]) / (2.f * (float)N); g_dt[1] += ga * 10.f * sgn(dt[1]) / (2.f * (float)N);
    }
    d_input_pos[m * 2] = g_xy[0]; d_input_pos[m * 2 + 1] = g_xy[1];
    go[(p.optim_sv_trans ? 2 : 0)] = g_xy[0]; go[(p.optim_sv_trans ? 2 : 0) + 1] = g_xy[1];
    if (p.pattern_trans) {
      if (g_pattern_trans) {
        float* gt = g_pattern_trans + ((p.isglobal ? 0 : img) * d.P + k) * 2;
        atomicAdd(gt, g_dt[0] * 0.01f); atomicAdd(gt + 1, g_dt[1] * 0.01f);
      }
    } else if (p.optim_sv_trans) {
      go[0] = g_dt[0] * 0.01f; go[1] = g_dt[1] * 0.01f;
    }
    if (d_weight) {
      const float w = expf(o[d.oc - 1] - mx) / tot;
      go[d.oc - 1] = w * (d_weight[m] - dot);
    }
  }
}

// input backward, one thread per (ray, point): d X row + d input_pos -> pattern_pos and img_embed gradients
__global__ void dsk_input_bwd_kernel(const edn_dsk_params p, const DskDims d, const int64_t* __restrict__ idx, const float* __restrict__ noise,
                                     int64_t N, const float* __restrict__ dX, int ld, const float* __restrict__ d_input_pos,
                                     float* __restrict__ g_pattern_pos, float* __restrict__ g_img_embed) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= N * d.P) return;
  const int64_t n = m / d.P;
  const int k = (int)(m % d.P);
  const int64_t img = idx[n];
  const int64_t po = ((p.isglobal ? 0 : img) * d.P + k) * 2;
  const float* row = dX + m * ld;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float th = tanhf(p.pattern_pos[po + i]);
    float pos = th * p.kernel_hwindow;
    if (noise) pos += noise[m * 2 + i];
    float gp = d_input_pos[m * 2 + i];
    if (d.L_in > 0) {
      const float sc = kPi / p.kernel_hwindow, x = pos * sc;
      float gx = row[i];
      for (int f = 0; f < d.L_in; ++f) {
        const float fr = (float)(1 << f);
        float sn, cs;
        sincosf(x * fr, &sn, &cs);
        gx += fr * (cs * row[2 + 4 * f + i] - sn * row[4 + 4 * f + i]);
      }
      gp += gx * sc;
    } else {
      gp += row[i];
    }
    if (g_pattern_pos) atomicAdd(g_pattern_pos + po + i, gp * p.kernel_hwindow * (1.f - th * th));
  }
  if (g_img_embed)
    for (int j = 0; j < d.embed; ++j) atomicAdd(g_img_embed + img * d.embed + j, row[d.pe_in + j]);
}

struct DskBuffers {
  float* X;            // [M][cat_w]: input row, the last hidden layer behind it when short_cut
  float* Hh[EDN_DSK_MAX_HIDDEN];
  int ldh[EDN_DSK_MAX_HIDDEN];
  float* O0; float* O; float* input_pos; float* align_part;
  float* dX; float* dA; float* dB; float* dO0; float* dO; float* d_input_pos;
};

DskBuffers carve(const DskDims& d, float* ws, int64_t M) {
  DskBuffers b;
  float* base = ws;
  auto take = [&](int64_t per) { float* q = base; base += per * M; return q; };
  b.X = take(d.cat_w);
  for (int l = 0; l < d.nh; ++l) {
    const bool last = (l == d.nh - 1);
    b.Hh[l] = last ? b.X + d.in_cnl : take(d.wide);
    b.ldh[l] = last ? d.cat_w : d.wide;
  }
  b.O0 = take(d.wide); b.O = take(8); b.input_pos = take(2); b.align_part = take(1);
  b.dX = take(d.cat_w); b.dA = take(d.wide); b.dB = take(d.wide); b.dO0 = take(d.wide); b.dO = take(8); b.d_input_pos = take(2);
  return b;
}

int dsk_check(const edn_dsk_params* p, const char* who) {
  EDN_REQUIRE(p && p->img_embed && p->pattern_pos && p->out0_w && p->out0_b && p->out1_w && p->out1_b, "%s: null parameter", who);
  EDN_REQUIRE(p->num_hidden >= 1 && p->num_hidden <= EDN_DSK_MAX_HIDDEN, "%s: num_hidden must be in [1, %d]", who, EDN_DSK_MAX_HIDDEN);
  for (int l = 0; l < p->num_hidden; ++l) EDN_REQUIRE(p->lin_w[l] && p->lin_b[l], "%s: null hidden layer %d", who, l);
  EDN_REQUIRE(p->n_pt >= 1 && p->n_pt <= 64 && p->wide >= 1 && p->wide <= 1024 && p->embed >= 0 && p->in_embed >= 0 && p->in_embed <= 16 &&
              p->spatial_embed >= 0 && p->spatial_embed <= 16 && p->n_img >= 1 && p->kernel_hwindow > 0.f, "%s: bad dimensions", who);
  return EDN_OK;
}

// x -> hidden layers -> linears1: fills b.X, b.Hh, b.O0, b.O (row stride 8), b.input_pos
int dsk_recompute(const edn_dsk_params* p, const DskDims& d, const DskBuffers& b, const Gemm& gemm, const float* rays_x, const float* rays_y,
                  const int64_t* idx, const float* noise, int64_t N, int H, int W, cudaStream_t st) {
  const int64_t M = N * d.P;
  dsk_input_kernel<<<blocks_for(M, 128), 128, 0, st>>>(*p, d, rays_x, rays_y, idx, noise, N, H, W, b.X, d.cat_w, b.input_pos);
#define EDN_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
  const float* prev = b.X; int ldp = d.cat_w, kin = d.in_cnl;
  for (int l = 0; l < d.nh; ++l) {
    EDN_RC(gemm(false, true, M, d.wide, kin, prev, ldp, p->lin_w[l], kin, 0.f, b.Hh[l], b.ldh[l]));
    dsk_bias_act_kernel<<<blocks_for(M * d.wide, 256), 256, 0, st>>>(b.Hh[l], b.ldh[l], d.wide, M, p->lin_b[l], 1);
    prev = b.Hh[l]; ldp = b.ldh[l]; kin = d.wide;
  }
  const float* cat = p->short_cut ? b.X : b.Hh[d.nh - 1];
  const int cat_k = p->short_cut ? d.cat_w : d.wide;
  EDN_RC(gemm(false, true, M, d.wide, cat_k, cat, d.cat_w, p->out0_w, cat_k, 0.f, b.O0, d.wide));
  dsk_bias_act_kernel<<<blocks_for(M * d.wide, 256), 256, 0, st>>>(b.O0, d.wide, d.wide, M, p->out0_b, 1);
  EDN_RC(gemm(false, true, M, d.oc, d.wide, b.O0, d.wide, p->out1_w, d.wide, 0.f, b.O, 8));
  dsk_bias_act_kernel<<<blocks_for(M * d.oc, 256), 256, 0, st>>>(b.O, 8, d.oc, M, p->out1_b, 0);
#undef EDN_RC
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

}  // namespace
}  // namespace edn

extern "C" int64_t edn_dsk_workspace_floats(const edn_dsk_params* p, int64_t n_rays) {
  using namespace edn;
  if (!p || n_rays < 0 || p->num_hidden < 1 || p->num_hidden > EDN_DSK_MAX_HIDDEN || p->n_pt < 1) return -1;
  return dsk_floats_per_row(dsk_dims(p)) * n_rays * p->n_pt + 64;
}

extern "C" int edn_dsk_rays_fwd(const edn_dsk_params* p, const float* rays_x, const float* rays_y, const int64_t* images_idx,
                                const float* poses, const float* noise, int64_t n_rays, int32_t H, int32_t W, float fx, float fy, float cx,
                                float cy, float* new_rays, float* weight, float* align, float* workspace, void* stream) {
  using namespace edn;
  if (int rc = dsk_check(p, "edn_dsk_rays_fwd")) return rc;
  EDN_REQUIRE(rays_x && rays_y && images_idx && poses && new_rays && weight && workspace, "edn_dsk_rays_fwd: null pointer");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  cublasHandle_t h = blas_handle();
  if (!h) return blas_unavailable();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cublasSetStream(h, st) != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetStream failed"); return EDN_E_CUDA; }
  const Gemm gemm{h, CUBLAS_COMPUTE_32F};
  const DskDims d = dsk_dims(p);
  const DskBuffers b = carve(d, workspace, n_rays * d.P);
  if (int rc = dsk_recompute(p, d, b, gemm, rays_x, rays_y, images_idx, noise, n_rays, H, W, st)) return rc;
  dsk_head_kernel<<<blocks_for(n_rays, 128), 128, 0, st>>>(*p, d, b.O, b.input_pos, rays_x, rays_y, images_idx, poses, n_rays,
                                                          HeadGeom{fx, fy, cx, cy}, new_rays, weight, align ? b.align_part : nullptr);
  if (align) dsk_align_reduce_kernel<<<1, 256, 0, st>>>(b.align_part, n_rays, align);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_dsk_rays_bwd(const edn_dsk_params* p, const float* rays_x, const float* rays_y, const int64_t* images_idx,
                                const float* poses, const float* noise, int64_t n_rays, int32_t H, int32_t W, float fx, float fy, float cx,
                                float cy, const float* d_new_rays, const float* d_weight, const float* d_align, const edn_dsk_grads* g,
                                float* workspace, void* stream) {
  using namespace edn;
  if (int rc = dsk_check(p, "edn_dsk_rays_bwd")) return rc;
  EDN_REQUIRE(rays_x && rays_y && images_idx && poses && g && workspace, "edn_dsk_rays_bwd: null pointer");
  EDN_REQUIRE(g->out0_w && g->out0_b && g->out1_w && g->out1_b, "edn_dsk_rays_bwd: null gradient buffer");
  for (int l = 0; l < p->num_hidden; ++l) EDN_REQUIRE(g->lin_w[l] && g->lin_b[l], "edn_dsk_rays_bwd: null gradient buffer (hidden layer %d)", l);
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  cublasHandle_t h = blas_handle();
  if (!h) return blas_unavailable();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cublasSetStream(h, st) != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetStream failed"); return EDN_E_CUDA; }
  const Gemm gemm{h, CUBLAS_COMPUTE_32F};
  const DskDims d = dsk_dims(p);
  const int64_t N = n_rays, M = N * d.P;
  const DskBuffers b = carve(d, workspace, M);
  if (int rc = dsk_recompute(p, d, b, gemm, rays_x, rays_y, images_idx, noise, N, H, W, st)) return rc;
  dsk_head_bwd_kernel<<<blocks_for(N, 128), 128, 0, st>>>(*p, d, b.O, b.input_pos, images_idx, poses, N, HeadGeom{fx, fy, cx, cy}, d_new_rays,
                                                          d_weight, d_align, b.dO, b.d_input_pos, g->pattern_trans);
#define EDN_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
  // linears1.2
  EDN_RC(gemm(true, false, d.oc, d.wide, M, b.dO, 8, b.O0, d.wide, 1.f, g->out1_w, d.wide));
  dsk_colsum_kernel<<<blocks_for(M, 256), 32, 0, st>>>(b.dO, 8, d.oc, M, g->out1_b);
  EDN_RC(gemm(false, false, M, d.wide, d.oc, b.dO, 8, p->out1_w, d.wide, 0.f, b.dO0, d.wide));
  dsk_relu_mask_kernel<<<blocks_for(M * d.wide, 256), 256, 0, st>>>(b.dO0, d.wide, b.O0, d.wide, d.wide, M);
  // linears1.0 on [x |] h
  const float* cat = p->short_cut ? b.X : b.Hh[d.nh - 1];
  const int cat_k = p->short_cut ? d.cat_w : d.wide;
  EDN_RC(gemm(true, false, d.wide, cat_k, M, b.dO0, d.wide, cat, d.cat_w, 1.f, g->out0_w, cat_k));
  dsk_colsum_kernel<<<blocks_for(M, 256), 1024, 0, st>>>(b.dO0, d.wide, d.wide, M, g->out0_b);
  // d [x | h_last] lands in dX (row stride cat_w); without the short cut only its h columns are written
  float* d_cat = p->short_cut ? b.dX : b.dX + d.in_cnl;
  EDN_RC(gemm(false, false, M, cat_k, d.wide, b.dO0, d.wide, p->out0_w, cat_k, 0.f, d_cat, d.cat_w));
  // hidden layers, last to first; dcur points at d h_l (row stride ldc)
  float* dcur = b.dX + d.in_cnl; int ldc = d.cat_w;
  for (int l = d.nh - 1; l >= 0; --l) {
    dsk_relu_mask_kernel<<<blocks_for(M * d.wide, 256), 256, 0, st>>>(dcur, ldc, b.Hh[l], b.ldh[l], d.wide, M);
    const float* prev = l > 0 ? b.Hh[l - 1] : b.X;
    const int ldp = l > 0 ? b.ldh[l - 1] : d.cat_w, kin = l > 0 ? d.wide : d.in_cnl;
    EDN_RC(gemm(true, false, d.wide, kin, M, dcur, ldc, prev, ldp, 1.f, g->lin_w[l], kin));
    dsk_colsum_kernel<<<blocks_for(M, 256), 1024, 0, st>>>(dcur, ldc, d.wide, M, g->lin_b[l]);
    if (l > 0) {
      float* dnext = (dcur == b.dA) ? b.dB : b.dA;
      EDN_RC(gemm(false, false, M, d.wide, d.wide, dcur, ldc, p->lin_w[l], d.wide, 0.f, dnext, d.wide));
      dcur = dnext; ldc = d.wide;
    } else {   // d x: added to the short-cut contribution already in dX's first in_cnl columns
      EDN_RC(gemm(false, false, M, d.in_cnl, d.wide, dcur, ldc, p->lin_w[0], d.in_cnl, p->short_cut ? 1.f : 0.f, b.dX, d.cat_w));
    }
  }
#undef EDN_RC
  dsk_input_bwd_kernel<<<blocks_for(M, 128), 128, 0, st>>>(*p, d, images_idx, noise, N, b.dX, d.cat_w, b.d_input_pos, g->pattern_pos, g->img_embed);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

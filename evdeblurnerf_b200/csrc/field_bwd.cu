// Backward of one PDRF field's render pass (the autograd backward of VoxelNeRFBase.sample + forward + raw2outputs,
// networks/pdrf/voxnerf.py:132-259, at the sample depths the forward pass placed; renderer.py:183-217).
//
// Design: the forward kernels keep every per-sample activation on chip, so nothing is saved for backward.  This pass
// recomputes the activations of a chunk of rays into an HBM workspace as plain row-major matrices and walks the chain
// backwards.  Every contraction here is a plain tall GEMM (M = samples of the chunk, up to 2^19 rows; N, K <= 256) -- the
// weight gradient dW = dY^T X reduces over all samples -- and goes to cuBLAS; everything that is not a GEMM is a hand-written
// kernel in this file: VM plane (.) line products and their scatter-add backward (vector red.global.add.v4.f32 into
// channel-last gradient planes), positional encodings and their backward, ReLU masks, the sigma->alpha compositing backward
// (a division-free suffix recursion, exact when a sample saturates alpha = 1) and the per-ray reduction onto the ray batch.
//
// Buffers per sample m of a chunk (fp32, row-major):
//   P_g [96]   plane (.) line products of grid g            X0 [ldX] = [ft_0 (32) | ft_1 (32) | PE(pts) (63) | 0]
//   H1  [hid]  relu(sigma_net.0)                            SG [ldS] = [geo_feat | sigma | PE(viewdir) (27) | 0]
// Every GEMM operand is 16-byte aligned with K, N multiples of 4 (cuBLAS then picks its sm_100 tensor-op kernels instead of
// the align1 fallbacks): the MLP weights are re-laid-out once per call into zero-padded / row-permuted copies (sigma_net.1:
// geo rows first, sigma row last; color_net.0: a zero column under sigma) and their gradients are folded back at the end.
//   H2, H3 [hid] relu(color_net.0/1)                        RGB [4]  = color_net.2 pre-activation
#include "bwd_common.cuh"

namespace edn {
namespace {

constexpr int kQuads = kAppComp / 4;          // 24 channel quads per sample
constexpr int kSamplesPerBlock = 8;
constexpr int kVmThreads = kQuads * kSamplesPerBlock;

struct GradGrid { float* plane[3]; float* line[3]; };

// ---- sample geometry ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sample_point(const float* __restrict__ rb, const float* __restrict__ z_vals, int64_t idx, int S,
                                             float p[3]) {
  const float* row = rb + (idx / S) * 11;
  const float z = __ldg(z_vals + idx);
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(__ldg(row + i), __fmul_rn(__ldg(row + 3 + i), z));   // as the forward kernels
}

// taps of VM component `comp` (plane over axes matMode[comp], line over vecMode[comp]) with the tap-weight derivatives
struct CompTaps {
  Taps2 pt; Taps1 lt;
  float dwx[4], dwy[4], dl[2];   // d w_k / d ix, d w_k / d iy (plane), d u_k / d iy (line); zero for out-of-range taps
  int ax, ay, av;                // point axes feeding plane x (W), plane y (H), line
  float sx, sy, sv;              // d ix / d n = (W - 1) / 2, ...
};

__device__ __forceinline__ void comp_taps(const GridDev& g, const float n[3], int comp, CompTaps& t, bool want_grad) {
  t.ax = (comp == 2) ? 1 : 0;
  t.ay = (comp == 0) ? 1 : 2;
  t.av = 2 - comp;
  const int H = g.ph[comp], W = g.pw[comp], L = g.ll[comp];
  plane_taps(n[t.ax], n[t.ay], H, W, t.pt);
  line_taps(n[t.av], L, t.lt);
  if (!want_grad) return;
  {
    const float ix = unnormalize(n[t.ax], W), iy = unnormalize(n[t.ay], H);
    const float x0 = floorf(ix), y0 = floorf(iy), x1 = x0 + 1.0f, y1 = y0 + 1.0f;
    const float wx0 = x1 - ix, wx1 = ix - x0, wy0 = y1 - iy, wy1 = iy - y0;
    const bool vx0 = (x0 >= 0.0f) && (x0 <= (float)(W - 1)), vx1 = (x1 >= 0.0f) && (x1 <= (float)(W - 1));
    const bool vy0 = (y0 >= 0.0f) && (y0 <= (float)(H - 1)), vy1 = (y1 >= 0.0f) && (y1 <= (float)(H - 1));
    t.dwx[0] = (vx0 && vy0) ? -wy0 : 0.f; t.dwy[0] = (vx0 && vy0) ? -wx0 : 0.f;
    t.dwx[1] = (vx1 && vy0) ? wy0 : 0.f;  t.dwy[1] = (vx1 && vy0) ? -wx1 : 0.f;
    t.dwx[2] = (vx0 && vy1) ? -wy1 : 0.f; t.dwy[2] = (vx0 && vy1) ? wx0 : 0.f;
    t.dwx[3] = (vx1 && vy1) ? wy1 : 0.f;  t.dwy[3] = (vx1 && vy1) ? wx1 : 0.f;
  }
  {
    const float iy = unnormalize(n[t.av], L);
    const float y0 = floorf(iy), y1 = y0 + 1.0f;
    t.dl[0] = ((y0 >= 0.0f) && (y0 <= (float)(L - 1))) ? -1.f : 0.f;
    t.dl[1] = ((y1 >= 0.0f) && (y1 <= (float)(L - 1))) ? 1.f : 0.f;
  }
  t.sx = 0.5f * (float)(W - 1); t.sy = 0.5f * (float)(H - 1); t.sv = 0.5f * (float)(L - 1);
}

__device__ __forceinline__ void quad_of(int q, int& comp, int& c, int& C) {
  if (q < 16) { comp = 0; c = q * 4; C = 64; }
  else if (q < 20) { comp = 1; c = (q - 16) * 4; C = 16; }
  else { comp = 2; c = (q - 20) * 4; C = 16; }
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 mul4(const float4 a, const float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 scale4(const float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ void fma4(float4& acc, const float4 v, float s) {
  acc.x = fmaf(v.x, s, acc.x); acc.y = fmaf(v.y, s, acc.y); acc.z = fmaf(v.z, s, acc.z); acc.w = fmaf(v.w, s, acc.w);
}

// P[m][96] = plane (.) line products (voxnerf.py:132-149); 24 threads per sample, one channel quad each.
template <typename T, typename AT>
__global__ void __launch_bounds__(kVmThreads) vm_products_kernel(const GridDev g, const float* __restrict__ rb,
                                                                  const float* __restrict__ z_vals, int64_t m0, int64_t Mc, int S,
                                                                  AT* __restrict__ P) {
  const int64_t t = (int64_t)blockIdx.x * kVmThreads + threadIdx.x;
  const int64_t m = t / kQuads;
  const int q = (int)(t % kQuads);
  if (m >= Mc) return;
  float p[3], n[3];
  sample_point(rb, z_vals, m0 + m, S, p);
  normalize_pt(g, p, n);
  int comp, c, C;
  quad_of(q, comp, c, C);
  CompTaps tp;
  comp_taps(g, n, comp, tp, false);
  const float4 v = gather4<T>(reinterpret_cast<const T*>(g.plane[comp]), reinterpret_cast<const T*>(g.line[comp]), C, c, tp.pt, tp.lt);
  stv4(P + m * kAppComp + q * 4, v);
}

// The same for bf16 grids with 12 threads per sample: one 8-channel chunk (a 16-byte load per tap) each -- half the tap
// arithmetic and half the load instructions of the quad version.  Same fp32 blend order as gather4.
constexpr int kOcts = kAppComp / 8;
template <typename AT>
__global__ void __launch_bounds__(kOcts * 16) vm_products8_kernel(const GridDev g, const float* __restrict__ rb,
                                                                  const float* __restrict__ z_vals, int64_t m0, int64_t Mc, int S,
                                                                  AT* __restrict__ P) {
  const int64_t t = (int64_t)blockIdx.x * (kOcts * 16) + threadIdx.x;
  const int64_t m = t / kOcts;
  const int o = (int)(t % kOcts);
  if (m >= Mc) return;
  float p[3], n[3];
  sample_point(rb, z_vals, m0 + m, S, p);
  normalize_pt(g, p, n);
  const int comp = o < 8 ? 0 : (o < 10 ? 1 : 2);
  const int c = (o < 8 ? o : (o < 10 ? o - 8 : o - 10)) * 8, C = comp == 0 ? 64 : 16;
  CompTaps tp;
  comp_taps(g, n, comp, tp, false);
  const __nv_bfloat16* plane = reinterpret_cast<const __nv_bfloat16*>(g.plane[comp]);
  const __nv_bfloat16* line = reinterpret_cast<const __nv_bfloat16*>(g.line[comp]);
  uint4 rp[4], rl[2];
#pragma unroll
  for (int k = 0; k < 4; ++k) rp[k] = __ldg(reinterpret_cast<const uint4*>(plane + (size_t)tp.pt.off[k] * C + c));
#pragma unroll
  for (int k = 0; k < 2; ++k) rl[k] = __ldg(reinterpret_cast<const uint4*>(line + (size_t)tp.lt.off[k] * C + c));
  float pl[8], ln[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { pl[i] = 0.f; ln[i] = 0.f; }
  auto unpack = [](const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  };
  float v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unpack(rp[k], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) pl[i] = fmaf(v[i], tp.pt.w[k], pl[i]);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    unpack(rl[k], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) ln[i] = fmaf(v[i], tp.lt.w[k], ln[i]);
  }
  AT* out = P + m * kAppComp + o * 8;
  stv4(out, make_float4(pl[0] * ln[0], pl[1] * ln[1], pl[2] * ln[2], pl[3] * ln[3]));
  stv4(out + 4, make_float4(pl[4] * ln[4], pl[5] * ln[5], pl[6] * ln[6], pl[7] * ln[7]));
}

// Backward of the products: scatter-add into the channel-last gradient planes / lines and accumulate d pts.
template <typename T, typename AT>
__global__ void __launch_bounds__(kVmThreads) vm_scatter_kernel(const GridDev g, const GradGrid gg, const float* __restrict__ rb,
                                                                 const float* __restrict__ z_vals, int64_t m0, int64_t Mc, int S,
                                                                 const AT* __restrict__ dP, float* __restrict__ dpts,
                                                                 const int* __restrict__ row_list, const AT* __restrict__ dP2) {
  __shared__ float part[kVmThreads][3];      // per-thread coordinate-gradient contributions (no shared-memory atomics: fp32 ones are CAS loops)
  part[threadIdx.x][0] = 0.f; part[threadIdx.x][1] = 0.f; part[threadIdx.x][2] = 0.f;
  const int64_t t = (int64_t)blockIdx.x * kVmThreads + threadIdx.x;
  const int64_t li = t / kQuads;                  // logical row: with a row list (the fine positions of a merged-scatter call, Mc of
  const int q = (int)(t % kQuads);                // them) the actual row of the chunk is row_list[li]
  const int64_t m = (row_list && li < Mc) ? (int64_t)__ldg(row_list + li) : li;
  if (li < Mc) {
    float p[3], n[3];
    sample_point(rb, z_vals, m0 + m, S, p);
    normalize_pt(g, p, n);
    int comp, c, C;
    quad_of(q, comp, c, C);
    CompTaps tp;
    comp_taps(g, n, comp, tp, true);
    const T* plane = reinterpret_cast<const T*>(g.plane[comp]);
    const T* line = reinterpret_cast<const T*>(g.line[comp]);
    float4 v[4], l[2];
    float4 pl = make_float4(0.f, 0.f, 0.f, 0.f), ln = pl;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = load4<T>(plane + (size_t)tp.pt.off[k] * C + c); fma4(pl, v[k], tp.pt.w[k]); }
#pragma unroll
    for (int k = 0; k < 2; ++k) { l[k] = load4<T>(line + (size_t)tp.lt.off[k] * C + c); fma4(ln, l[k], tp.lt.w[k]); }
    float4 dp = ldv4(dP + m * kAppComp + q * 4);
    if (dP2) {
      const float4 d2 = ldv4(dP2 + (m0 + m) * kAppComp + q * 4);
      dp.x += d2.x; dp.y += d2.y; dp.z += d2.z; dp.w += d2.w;
    }
    const float4 dpl = mul4(dp, ln), dln = mul4(dp, pl);
    float gx = 0.f, gy = 0.f, gv = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#if !defined(EDN_SCATTER_ABLATE) || EDN_SCATTER_ABLATE != 2
      if (tp.pt.w[k] != 0.f) atomicAdd(reinterpret_cast<float4*>(gg.plane[comp] + (size_t)tp.pt.off[k] * C + c), scale4(dpl, tp.pt.w[k]));
#endif
      const float dv = dot4(dpl, v[k]);
      gx = fmaf(tp.dwx[k], dv, gx);
      gy = fmaf(tp.dwy[k], dv, gy);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#if !defined(EDN_SCATTER_ABLATE) || EDN_SCATTER_ABLATE != 1      // dev timing ablation: 1 = no line reds, 2 = no plane reds
      if (tp.lt.w[k] != 0.f) atomicAdd(reinterpret_cast<float4*>(gg.line[comp] + (size_t)tp.lt.off[k] * C + c), scale4(dln, tp.lt.w[k]));
#endif
      gv = fmaf(tp.dl[k], dot4(dln, l[k]), gv);
    }
    part[threadIdx.x][tp.ax] = gx * tp.sx;      // (ax, ay, av) is a permutation of the three point axes
    part[threadIdx.x][tp.ay] = gy * tp.sy;
    part[threadIdx.x][tp.av] = gv * tp.sv;
  }
  __syncthreads();
  if (threadIdx.x < kSamplesPerBlock * 3) {
    const int s = threadIdx.x / 3, i = threadIdx.x % 3;
    const int64_t ll = (int64_t)blockIdx.x * kSamplesPerBlock + s;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < kQuads; ++q) acc += part[s * kQuads + q][i];
    if (ll < Mc) {
      const int64_t mm = row_list ? (int64_t)__ldg(row_list + ll) : ll;
      dpts[mm * 4 + i] += acc * g.inv[i];            // n = (p - amin) * inv - 1
    }
  }
}

// The fine pass samples the COARSE grid at all its merged depths, and Nc of them are the coarse pass's own positions: those rows of
// d P (coarse-grid products) are moved here into [ray][coarse index][96] and added to the coarse pass's d P inside ITS scatter, so
// every coarse position is scattered once instead of twice (a fifth of all reds of a c2f step).  One thread per 16-byte piece.
// The rows that stay (the importance samples: order >= n_coarse, S - n_coarse per ray) are listed compactly in row_list, so the
// fine call's coarse-grid scatter runs over full warps of live rows.
template <typename AT>
__global__ void move_rows_kernel(const AT* __restrict__ dP, const int64_t* __restrict__ order, int64_t m0, int64_t Mc, int S, int n_coarse,
                                 AT* __restrict__ moved, int* __restrict__ row_list) {
  constexpr int kPieces = kAppComp * (int)sizeof(AT) / 16;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / kPieces;
  const int pc = (int)(t % kPieces);
  if (m >= Mc) return;
  const int64_t src = __ldg(order + m0 + m);
  if (src >= n_coarse) {
    if (pc == 0) row_list[(m / S) * (S - n_coarse) + (src - n_coarse)] = (int)m;      // chunks start at ray boundaries: m / S = local ray
    return;
  }
  const int64_t ray = (m0 + m) / S;
  const uint4* from = reinterpret_cast<const uint4*>(dP + m * kAppComp);
  uint4* to = reinterpret_cast<uint4*>(moved + (ray * n_coarse + src) * kAppComp);
  to[pc] = from[pc];
}

// Positional encodings (embedding.py:88-98): X0[:, nf : ldX) = [PE(pts) (63) | 0] (dirs = 0), SG[:, geo+1 : ldS) = [PE(viewdir) (27) | 0]
// (dirs = 1; launched after the sigma_net.1 GEMM, whose padded output columns overlap the first PE(viewdir) columns).
// One thread per (sample, group): group 0 writes the 3 identity columns (and the zero padding behind the encoding), group
// 1 + k the sin / cos triplets of frequency 2^k (one sincosf per axis).
template <typename AT>
__global__ void pe_kernel(const float* __restrict__ rb, const float* __restrict__ z_vals, int64_t m0, int64_t Mc, int S,
                          AT* __restrict__ X0, int ldX, int nf, int wp, AT* __restrict__ SG, int ldS, int geo, int dirs) {
  const int L = dirs ? kPeFreqDir : kPeFreqPts, n_pe = 3 + 6 * L;
  const int width = dirs ? ldS - 1 - geo : wp;        // columns to fill: [PE | zero padding]
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / (L + 1);
  const int gq = (int)(t % (L + 1));
  if (m >= Mc) return;
  float x[3];
  if (!dirs) sample_point(rb, z_vals, m0 + m, S, x);
  else {
    const float* row = rb + ((m0 + m) / S) * 11 + 8;
    x[0] = __ldg(row); x[1] = __ldg(row + 1); x[2] = __ldg(row + 2);
  }
  AT* out = dirs ? SG + m * ldS + geo + 1 : X0 + m * ldX + nf;
  if (gq == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = from_f<AT>(x[i]);
    for (int j = n_pe; j < width; ++j) out[j] = from_f<AT>(0.f);
  } else {
    const float fr = (float)(1 << (gq - 1));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float sn, cs;
      sincosf(x[i] * fr, &sn, &cs);
      out[3 + 6 * (gq - 1) + i] = from_f<AT>(sn);
      out[6 + 6 * (gq - 1) + i] = from_f<AT>(cs);
    }
  }
}

// PE(viewdir) is constant along a ray: one block per ray computes the 27 values once (same sincosf arguments as pe_kernel, so the
// stored values are identical) and copies them, with the zero padding, into the S sample rows: SG[m][geo + 1 : ldS).
template <typename AT>
__global__ void __launch_bounds__(128) pe_dirs_ray_kernel(const float* __restrict__ rb, int64_t r0, int S, AT* __restrict__ SG, int ldS, int geo) {
  __shared__ float pe[64];
  const int width = ldS - 1 - geo;       // [PE (27) | zero padding]
  const float* row = rb + (r0 + blockIdx.x) * 11 + 8;
  for (int j = threadIdx.x; j < width && j < 64; j += blockDim.x) pe[j] = 0.f;
  __syncthreads();
  if (threadIdx.x < 3) {
    const int i = threadIdx.x;
    const float x = __ldg(row + i);
    pe[i] = x;
#pragma unroll
    for (int f = 0; f < kPeFreqDir; ++f) {
      float sn, cs;
      sincosf(x * (float)(1 << f), &sn, &cs);
      pe[3 + 6 * f + i] = sn;
      pe[6 + 6 * f + i] = cs;
    }
  }
  __syncthreads();
  AT* out = SG + (int64_t)blockIdx.x * S * ldS + geo + 1;
  for (int idx = threadIdx.x; idx < S * width; idx += blockDim.x) {
    const int sp = idx / width, j = idx - sp * width;
    out[(int64_t)sp * ldS + j] = from_f<AT>(j < 64 ? pe[j] : 0.f);
  }
}

// bf16 storage, PE(pts) only: one thread per sample writes the whole [PE (63) | 0] block as eight 16-byte vectors.  Base
// frequency by sincosf, higher octaves by the double-angle recurrence -- the same recipe as the tensor-core forward (abs
// error <= 2^9 * 1e-7, far below the bf16 resolution of the stored operand).  Needs nf * 2 and ldX * 2 multiples of 16 bytes.
__global__ void pe_pts_bf16_kernel(const float* __restrict__ rb, const float* __restrict__ z_vals, int64_t m0, int64_t Mc, int S,
                                   __nv_bfloat16* __restrict__ X0, int ldX, int nf) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mc) return;
  float pe[64];
  sample_point(rb, z_vals, m0 + m, S, pe);
#pragma unroll
  for (int i = 0; i < 3; ++i) sincosf(pe[i], &pe[3 + i], &pe[6 + i]);
#pragma unroll
  for (int f = 1; f < kPeFreqPts; ++f) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float sp = pe[3 + 6 * (f - 1) + i], cp = pe[6 + 6 * (f - 1) + i];
      pe[3 + 6 * f + i] = 2.0f * sp * cp;
      pe[6 + 6 * f + i] = fmaf(-2.0f * sp, sp, 1.0f);
    }
  }
  pe[63] = 0.f;
  uint4* out = reinterpret_cast<uint4*>(X0 + m * ldX + nf);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 t = __floats2bfloat162_rn(pe[8 * j + 2 * i], pe[8 * j + 2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&t);
    }
    out[j] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// Weight re-layout between the reference nn.Linear tensors and the aligned copies the GEMMs read (to_padded = 1), and the
// fold-back of their gradients (to_padded = 0: ref += padded).  One thread per padded element.
//   mode 0: zero-pad columns            ref [rows][cols_r]      -> pad [rows][cols_p]
//   mode 1: sigma_net.1                 ref [1+geo][cols]       -> pad [rows_p][cols]: rows 0..geo-1 = ref rows 1..geo, row geo = ref row 0
//   mode 2: color_net.0                 ref [rows][geo+27]      -> pad [rows][cols_p]: col geo (under sigma) = 0, PE(dir) cols shifted by one
//   mode 3: zero-pad rows               ref [rows_r][cols]      -> pad [rows_p][cols]
__device__ __forceinline__ bool relayout_map(int t, int mode, int cols_p, int rows_r, int cols_r, int geo, int& src) {
  const int r = t / cols_p, c = t % cols_p;
  int rr = r, cc = c;
  if (mode == 1) rr = r < geo ? r + 1 : (r == geo ? 0 : -1);
  if (mode == 2) cc = c < geo ? c : (c == geo ? -1 : c - 1);
  src = rr * cols_r + cc;
  return rr >= 0 && rr < rows_r && cc >= 0 && cc < cols_r;
}
template <typename AT>
__global__ void relayout_to_padded_kernel(AT* __restrict__ pad, const float* __restrict__ ref, int mode, int rows_p, int cols_p, int rows_r,
                                          int cols_r, int geo) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows_p * cols_p) return;
  int src;
  const bool valid = relayout_map(t, mode, cols_p, rows_r, cols_r, geo, src);
  pad[t] = from_f<AT>(valid ? ref[src] : 0.f);
}
__global__ void relayout_fold_kernel(const float* __restrict__ pad, float* __restrict__ ref, int mode, int rows_p, int cols_p, int rows_r,
                                     int cols_r, int geo) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows_p * cols_p) return;
  int src;
  if (relayout_map(t, mode, cols_p, rows_r, cols_r, geo, src)) ref[src] += pad[t];
}
// dst[m][0..n) = src[m][0..n)  (fp32 -> storage type, different leading dimensions)
template <typename AT>
__global__ void convert_cols_kernel(AT* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, int n, int64_t M) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / n;
  const int j = (int)(t % n);
  if (m < M) dst[m * ldd + j] = from_f<AT>(src[m * lds + j]);
}
template <typename AT>
__global__ void convert_kernel(AT* __restrict__ dst, const float* __restrict__ src, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] = from_f<AT>(src[t]);
}

// dSG[m][geo] = dsig[m]
template <typename AT>
__global__ void set_sigma_grad_kernel(AT* __restrict__ dSG, int ldS, int geo, int64_t M, const float* __restrict__ dsig) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) dSG[m * ldS + geo] = from_f<AT>(dsig[m]);
}

// d pts from the PE(pts) columns of dX0: writes (=) dpts[m][0..2].
template <typename AT>
__global__ void pe_bwd_kernel(const float* __restrict__ rb, const float* __restrict__ z_vals, int64_t m0, int64_t Mc, int S,
                              const AT* __restrict__ dX0, int ldX, int nf, float* __restrict__ dpts) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / 3;
  const int i = (int)(t % 3);
  if (m >= Mc) return;
  float p[3];
  sample_point(rb, z_vals, m0 + m, S, p);
  const AT* gr = dX0 + m * ldX + nf;
  float acc = to_f(gr[i]);
#pragma unroll
  for (int k = 0; k < kPeFreqPts; ++k) {
    const float fr = (float)(1 << k);
    float sn, cs;
    sincosf(p[i] * fr, &sn, &cs);
    acc += fr * (cs * to_f(gr[3 + 6 * k + i]) - sn * to_f(gr[6 + 6 * k + i]));
  }
  dpts[m * 4 + i] = acc;
}

// dH3[m][j] = H3[m][j] > 0 ? sum_c dRGB[m][c] * W2[c][j] : 0     (color_net.2 is [3][hid]: a K = 3 contraction); one 16-byte
// vector of H3 / dH3 per thread
template <typename AT>
__global__ void head_bwd_kernel(const AT* __restrict__ dRGB, int nr, const float* __restrict__ W2, const AT* __restrict__ H3, int hid,
                                int64_t M, AT* __restrict__ dH3, bool w2_vec) {
  constexpr int V = Vec16<AT>::n;
  const int nq = hid / V;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / nq;
  const int j = (int)(t % nq) * V;
  if (m >= M) return;
  const float4 g = ldv4(dRGB + m * nr);
  Vec16<AT> h;
  h.load(H3 + m * hid + j);
  if (w2_vec) {        // W2 16-byte aligned (uniform): the three weight rows as float4 loads
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(W2 + j + i)), b = __ldg(reinterpret_cast<const float4*>(W2 + hid + j + i)),
                   c = __ldg(reinterpret_cast<const float4*>(W2 + 2 * hid + j + i));
      h.v[i] = h.v[i] > 0.f ? g.x * a.x + g.y * b.x + g.z * c.x : 0.f;
      h.v[i + 1] = h.v[i + 1] > 0.f ? g.x * a.y + g.y * b.y + g.z * c.y : 0.f;
      h.v[i + 2] = h.v[i + 2] > 0.f ? g.x * a.z + g.y * b.z + g.z * c.z : 0.f;
      h.v[i + 3] = h.v[i + 3] > 0.f ? g.x * a.w + g.y * b.w + g.z * c.w : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i)
      h.v[i] = h.v[i] > 0.f ? g.x * __ldg(W2 + j + i) + g.y * __ldg(W2 + hid + j + i) + g.z * __ldg(W2 + 2 * hid + j + i) : 0.f;
  }
  h.store(dH3 + m * hid + j);
}

// dSG[m][j] += d_feat[m0 + m][j]   (upstream gradient of feature_map, voxnerf.py:221); 4 columns per thread (geo % 4 == 0),
// scalar tail otherwise
template <typename AT>
__global__ void add_feat_grad_kernel(AT* __restrict__ dSG, int ldS, int geo, int64_t M, const float* __restrict__ d_feat) {
  const int nq = (geo + 3) / 4;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / nq;
  const int j = (int)(t % nq) * 4;
  if (m >= M) return;
  if ((geo & 3) == 0) {
    float4 d = ldv4(dSG + m * ldS + j);
    const float4 f = __ldg(reinterpret_cast<const float4*>(d_feat + m * geo + j));
    d.x += f.x; d.y += f.y; d.z += f.z; d.w += f.w;
    stv4(dSG + m * ldS + j, d);
  } else {
    for (int i = j; i < min(j + 4, geo); ++i) dSG[m * ldS + i] = from_f<AT>(to_f(dSG[m * ldS + i]) + d_feat[m * geo + i]);
  }
}

// Compositing backward (voxnerf.py:153-201), one thread per ray of the chunk.
//   w_i = alpha_i T_i, T_i = prod_{j<i} (1 - alpha_j);  g_i = dL/dw_i = d_rgb . c_i + d_depth z_i + d_acc + d_weights_i
//   dL/dalpha_i = T_i (g_i - S_i),  S_i = sum_{k>i} g_k alpha_k prod_{i<j<k} (1 - alpha_j) = g_{i+1} alpha_{i+1} + (1 - alpha_{i+1}) S_{i+1}
// (no division by 1 - alpha_i: exact when a sample saturates, like torch's cumprod backward).
template <typename AT>
__global__ void composite_bwd_kernel(const AT* __restrict__ SG /* + geo: the sigma column */, int ldS, const float* __restrict__ RGB, int nr, const float* __restrict__ b2,
                                     const float* __restrict__ rb, const float* __restrict__ z_vals, const float* __restrict__ noise,
                                     int64_t r0, int64_t Rc, int S, const float* __restrict__ d_rgb, const float* __restrict__ d_depth,
                                     const float* __restrict__ d_acc, const float* __restrict__ d_weights, float* __restrict__ al,
                                     float* __restrict__ tr, AT* __restrict__ dRGB, float* __restrict__ dsig_out, float* __restrict__ d_rb) {
  const int64_t rl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rl >= Rc) return;
  const int64_t r = r0 + rl;
  const float* z = z_vals + r * S;
  const float* nz = noise ? noise + r * (S - 1) : nullptr;
  const float* row = rb + r * 11;
  const float d[3] = {row[3], row[4], row[5]};
  const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const int64_t mb = rl * S;
  float T = 1.0f;
  for (int s = 0; s < S; ++s) {
    const float a = alpha_of_sample(to_f(SG[(mb + s) * ldS]), z[s], s + 1 < S ? z[s + 1] : 0.f, nz && s + 1 < S ? nz[s] : 0.f, dn, false, 0.f,
                                    s == S - 1);
    al[mb + s] = a; tr[mb + s] = T;
    T = T * (1.0f - a);
  }
  const float gr[3] = {d_rgb ? d_rgb[r * 3] : 0.f, d_rgb ? d_rgb[r * 3 + 1] : 0.f, d_rgb ? d_rgb[r * 3 + 2] : 0.f};
  const float gd = d_depth ? d_depth[r] : 0.f, ga = d_acc ? d_acc[r] : 0.f;
  const float bb[3] = {b2 ? b2[0] : 0.f, b2 ? b2[1] : 0.f, b2 ? b2[2] : 0.f};
  float Ssum = 0.f, ddn = 0.f;
  for (int s = S - 1; s >= 0; --s) {
    const int64_t m = mb + s;
    const float a = al[m], Ti = tr[m], w = a * Ti;
    const float4 x = *reinterpret_cast<const float4*>(RGB + m * nr);
    const float c[3] = {sigmoidf_(x.x + bb[0]), sigmoidf_(x.y + bb[1]), sigmoidf_(x.z + bb[2])};
    const float gw = gr[0] * c[0] + gr[1] * c[1] + gr[2] * c[2] + gd * z[s] + ga + (d_weights ? d_weights[r * S + s] : 0.f);
    stv4(dRGB + m * nr, make_float4(w * gr[0] * c[0] * (1.f - c[0]), w * gr[1] * c[1] * (1.f - c[1]), w * gr[2] * c[2] * (1.f - c[2]), 0.f));
    if (nr > 4) stv4(dRGB + m * nr + 4, make_float4(0.f, 0.f, 0.f, 0.f));
    float dsig = 0.f;
    if (s < S - 1) {
      const float da = Ti * (gw - Ssum);
      const float dz = z[s + 1] - z[s];
      const float dist = __fmul_rn(dz, dn);
      const float sraw = to_f(SG[m * ldS]) + (nz ? nz[s] : 0.f);
      const float sg = fmaxf(sraw, 0.f), one_m = 1.0f - a;
      dsig = sraw > 0.f ? da * dist * one_m : 0.f;
      ddn = fmaf(da * sg * one_m, dz, ddn);
    }
    dsig_out[m] = dsig;
    Ssum = gw * a + (1.0f - a) * Ssum;
  }
  if (dn > 0.f) {
#pragma unroll
    for (int i = 0; i < 3; ++i) d_rb[r * 11 + 3 + i] += ddn * d[i] / dn;     // dists = dz * ||rays_d||  (voxnerf.py:160)
  }
}

// d ray_batch from the per-sample d pts (pts = o + d z) and the PE(viewdir) columns of dSG; one warp per ray.
template <typename AT>
__global__ void ray_reduce_kernel(const float* __restrict__ dpts, const AT* __restrict__ dSG, int ldS, int geo,
                                  const float* __restrict__ rb, const float* __restrict__ z_vals, int64_t r0, int64_t Rc, int S,
                                  float* __restrict__ d_rb) {
  const int64_t rl = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (rl >= Rc) return;
  const int64_t r = r0 + rl;
  float acc[6 + kPeDir];
#pragma unroll
  for (int j = 0; j < 6 + kPeDir; ++j) acc[j] = 0.f;
  for (int s = lane; s < S; s += 32) {
    const int64_t m = rl * S + s;
    const float z = z_vals[r * S + s];
    const float4 g = *reinterpret_cast<const float4*>(dpts + m * 4);
    acc[0] += g.x; acc[1] += g.y; acc[2] += g.z;
    acc[3] = fmaf(z, g.x, acc[3]); acc[4] = fmaf(z, g.y, acc[4]); acc[5] = fmaf(z, g.z, acc[5]);
    const AT* gv = dSG + m * ldS + geo + 1;
#pragma unroll
    for (int j = 0; j < kPeDir; ++j) acc[6 + j] += to_f(gv[j]);
  }
#pragma unroll
  for (int j = 0; j < 6 + kPeDir; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (lane == 0) {
    float* out = d_rb + r * 11;
#pragma unroll
    for (int i = 0; i < 6; ++i) out[i] += acc[i];
    const float* vd = rb + r * 11 + 8;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float g = acc[6 + i];
#pragma unroll
      for (int k = 0; k < kPeFreqDir; ++k) {
        const float fr = (float)(1 << k);
        float sn, cs;
        sincosf(vd[i] * fr, &sn, &cs);
        g += fr * (cs * acc[6 + 3 + 6 * k + i] - sn * acc[6 + 6 + 6 * k + i]);
      }
      out[8 + i] += g;
    }
  }
}

struct Dims {
  int ng, nf, hid, geo, ldX, ldS, kin, cin, sgn, nr, esz;
};
// al = elements per 16 bytes of the activation storage type (4: fp32, 8: bf16); every leading dimension is a multiple of it
inline Dims make_dims(int n_grids, int hidden, int geo, int al) {
  Dims d;
  d.ng = n_grids; d.nf = 32 * n_grids; d.hid = hidden; d.geo = geo;
  d.kin = d.nf + kPePts;                          // 95 | 127
  d.cin = geo + kPeDir;                           // 42 | 155
  d.ldX = (d.kin + al - 1) / al * al;             // 96 | 128
  d.ldS = (1 + d.cin + al - 1) / al * al;         // fp32: 44 | 156, bf16: 48 | 160
  d.sgn = (1 + geo + al - 1) / al * al;           // padded output width of sigma_net.1: 16 | 132 (fp32), 16 | 136 (bf16)
  d.nr = al;                                      // padded width of the rgb head
  d.esz = 16 / al;
  return d;
}
inline int64_t bytes_per_sample(const Dims& d) {
  const int64_t act = (int64_t)kAppComp * d.ng /*P*/ + d.ldX /*X0*/ + d.hid * 3 /*H1 H2 H3*/ + d.ldS /*SG*/ + d.nr /*dRGB*/ + d.hid * 2 /*D1 D2*/ +
                      d.ldS /*dSG*/ + d.nf /*dXf*/ + kAppComp /*dP*/;
  return act * d.esz + (int64_t)sizeof(float) * (d.ldX /*dX0*/ + d.nr /*RGB*/ + 4 /*dpts*/ + 3 /*alpha, T, d sigma*/ + 1 /*row list*/);
}
inline int64_t weight_scratch_bytes(const Dims& d) {     // aligned weight copies (storage type) + fp32 gradients of the re-laid-out ones
  const int64_t relaid = (int64_t)d.hid * d.ldX + (int64_t)d.sgn * d.hid + (int64_t)d.hid * d.ldS + (int64_t)d.nr * d.hid;
  const int64_t plain = (int64_t)d.hid * d.hid + (int64_t)d.ng * kAppDim * kAppComp;
  return (relaid + plain) * d.esz + relaid * (int64_t)sizeof(float) + 256;
}

struct FieldBwdCall {
  const edn_vm_grid* grids[2];
  GridDev gd[2];
  GradGrid gg[2];
  const edn_field_weights* w;
  const edn_field_weights* grad_w;
  const float* ray_batch; const float* z_vals; const float* noise;
  int64_t n_rays; int S;
  const float* d_rgb; const float* d_depth; const float* d_acc; const float* d_weights; const float* d_feat;
  float* d_ray_batch;
  void* workspace; int64_t workspace_bytes;
  cudaStream_t st;
  cublasComputeType_t ct;
  const int64_t* merge_order; int n_coarse; void* moved;      // edn_field_bwd_merge (NULL = every row scattered where it was computed)
};

template <typename AT>
int field_bwd_run(const FieldBwdCall& c) {
  constexpr int al = 16 / (int)sizeof(AT);
  const edn_field_weights* w = c.w;
  const edn_field_weights* grad_w = c.grad_w;
  const int ng = w->n_grids;
  const Dims D = make_dims(ng, w->hidden, w->geo_feat, al);
  const int S = c.S, hid = D.hid, geo = D.geo;
  cudaStream_t st = c.st;
  const float* ray_batch = c.ray_batch; const float* z_vals = c.z_vals;
  const int64_t n_rays = c.n_rays;
  const int64_t per_ray = bytes_per_sample(D) * S;
  int64_t chunk = (c.workspace_bytes - weight_scratch_bytes(D)) / per_ray;
  EDN_REQUIRE(chunk >= 1, "edn_render_field_bwd: workspace too small (%lld bytes, one ray needs %lld)", (long long)c.workspace_bytes,
              (long long)(per_ray + weight_scratch_bytes(D)));
  chunk = chunk < n_rays ? chunk : n_rays;
  if (chunk * S > (1 << 22)) chunk = (1 << 22) / S;         // keep GEMM row counts well inside int range

  cublasHandle_t h = blas_handle();
  if (!h) return blas_unavailable();
  if (cublasSetStream(h, st) != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetStream failed"); return EDN_E_CUDA; }
  const Gemm gemm{h, c.ct};

  char* base = reinterpret_cast<char*>(c.workspace);
  auto carve = [&](int64_t bytes) { char* p = base; base += (bytes + 15) / 16 * 16; return p; };
  // aligned weight copies in the storage type; fp32 gradient accumulators for the re-laid-out ones
  const int nW[4] = {hid * D.ldX, D.sgn * hid, hid * D.ldS, D.nr * hid};
  AT* Wp[4];
  float* gWp[4];
  for (int i = 0; i < 4; ++i) Wp[i] = reinterpret_cast<AT*>(carve((int64_t)nW[i] * sizeof(AT)));
  AT* Wc1 = reinterpret_cast<AT*>(carve((int64_t)hid * hid * sizeof(AT)));
  AT* Wb[2] = {nullptr, nullptr};
  for (int g = 0; g < ng; ++g) Wb[g] = reinterpret_cast<AT*>(carve((int64_t)kAppDim * kAppComp * sizeof(AT)));
  gWp[0] = reinterpret_cast<float*>(carve((int64_t)(nW[0] + nW[1] + nW[2] + nW[3]) * sizeof(float)));
  for (int i = 1; i < 4; ++i) gWp[i] = gWp[i - 1] + nW[i - 1];
  EDN_CUDA_OK(cudaMemsetAsync(gWp[0], 0, sizeof(float) * (size_t)(nW[0] + nW[1] + nW[2] + nW[3]), st));
  relayout_to_padded_kernel<AT><<<blocks_for(nW[0], 256), 256, 0, st>>>(Wp[0], w->sigma0, 0, hid, D.ldX, hid, D.kin, geo);
  relayout_to_padded_kernel<AT><<<blocks_for(nW[1], 256), 256, 0, st>>>(Wp[1], w->sigma1, 1, D.sgn, hid, 1 + geo, hid, geo);
  relayout_to_padded_kernel<AT><<<blocks_for(nW[2], 256), 256, 0, st>>>(Wp[2], w->color0, 2, hid, D.ldS, hid, D.cin, geo);
  relayout_to_padded_kernel<AT><<<blocks_for(nW[3], 256), 256, 0, st>>>(Wp[3], w->color2, 3, D.nr, hid, 3, hid, geo);
  convert_kernel<AT><<<blocks_for(hid * hid, 256), 256, 0, st>>>(Wc1, w->color1, hid * hid);
  for (int g = 0; g < ng; ++g) convert_kernel<AT><<<blocks_for(kAppDim * kAppComp, 256), 256, 0, st>>>(Wb[g], w->basis[g], kAppDim * kAppComp);

  const int64_t Mmax = chunk * S;
  auto take = [&](int64_t per) { return reinterpret_cast<AT*>(carve(per * Mmax * (int64_t)sizeof(AT))); };
  auto takef = [&](int64_t per) { return reinterpret_cast<float*>(carve(per * Mmax * (int64_t)sizeof(float))); };
  AT* P[2] = {take(kAppComp), ng == 2 ? take(kAppComp) : nullptr};
  AT* X0 = take(D.ldX);
  AT* H1 = take(hid);
  AT* H2 = take(hid);
  AT* H3 = take(hid);
  AT* SG = take(D.ldS);
  AT* dRGB = take(D.nr);
  AT* D1 = take(hid);
  AT* D2 = take(hid);
  AT* dSG = take(D.ldS);
  AT* dXf = take(D.nf);            // feature columns of dX0 in the storage type (GEMM operand of the basis_mat backward)
  AT* dP = take(kAppComp);
  float* dX0 = takef(D.ldX);       // fp32: its PE columns are multiplied by up to 2^9 in the PE backward
  float* RGB = takef(D.nr);
  float* dpts = takef(4);
  float* al_ = takef(1);
  float* tr = takef(1);
  float* dsig = takef(1);
  int* row_list = reinterpret_cast<int*>(takef(1));      // merged scatter: the chunk's rows that are not coarse positions

#define EDN_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
  for (int64_t r0 = 0; r0 < n_rays; r0 += chunk) {
    const int64_t Rc = (n_rays - r0) < chunk ? (n_rays - r0) : chunk;
    const int64_t M = Rc * S, m0 = r0 * S;
    // ---- recompute the forward activations -------------------------------------------------------------------------
    for (int g = 0; g < ng; ++g) {
      if (c.grids[g]->dtype == EDN_F32) vm_products_kernel<float, AT><<<blocks_for(M, kSamplesPerBlock), kVmThreads, 0, st>>>(c.gd[g], ray_batch, z_vals, m0, M, S, P[g]);
      else vm_products8_kernel<AT><<<blocks_for(M, 16), kOcts * 16, 0, st>>>(c.gd[g], ray_batch, z_vals, m0, M, S, P[g]);
      EDN_RC(gemm.run(false, true, M, kAppDim, kAppComp, P[g], kAppComp, Wb[g], kAppComp, 0.f, X0 + 32 * g, D.ldX));
    }
    if constexpr (sizeof(AT) == 2) {
      if (D.ldX - D.nf == 64 && D.nf % 8 == 0 && D.ldX % 8 == 0)
        pe_pts_bf16_kernel<<<blocks_for(M, 128), 128, 0, st>>>(ray_batch, z_vals, m0, M, S, reinterpret_cast<__nv_bfloat16*>(X0), D.ldX, D.nf);
      else
        pe_kernel<AT><<<blocks_for(M * (kPeFreqPts + 1), 256), 256, 0, st>>>(ray_batch, z_vals, m0, M, S, X0, D.ldX, D.nf, D.ldX - D.nf, SG, D.ldS, geo, 0);
    } else {
      pe_kernel<AT><<<blocks_for(M * (kPeFreqPts + 1), 256), 256, 0, st>>>(ray_batch, z_vals, m0, M, S, X0, D.ldX, D.nf, D.ldX - D.nf, SG, D.ldS, geo, 0);
    }
    EDN_RC(gemm.relu_linear(M, hid, D.ldX, X0, D.ldX, Wp[0], D.ldX, (const float*)nullptr, H1, hid, st,
                            [&] { relu_bias_kernel<AT><<<blocks_for(M * (hid / al), 256), 256, 0, st>>>(H1, hid, hid, M, nullptr); }));
    EDN_RC(gemm.run(false, true, M, D.sgn, hid, H1, hid, Wp[1], hid, 0.f, SG, D.ldS));                               // [geo | sigma | 0..]
    pe_dirs_ray_kernel<AT><<<(unsigned)Rc, 128, 0, st>>>(ray_batch, r0, S, SG, D.ldS, geo);
    EDN_RC(gemm.relu_linear(M, hid, D.ldS, SG, D.ldS, Wp[2], D.ldS, w->color0_b, H2, hid, st,
                            [&] { relu_bias_kernel<AT><<<blocks_for(M * (hid / al), 256), 256, 0, st>>>(H2, hid, hid, M, w->color0_b); }));
    EDN_RC(gemm.relu_linear(M, hid, hid, H2, hid, Wc1, hid, w->color1_b, H3, hid, st,
                            [&] { relu_bias_kernel<AT><<<blocks_for(M * (hid / al), 256), 256, 0, st>>>(H3, hid, hid, M, w->color1_b); }));
    EDN_RC(gemm.run(false, true, M, D.nr, hid, H3, hid, Wp[3], hid, 0.f, RGB, D.nr));
    // ---- compositing backward ------------------------------------------------------------------------------------------
    composite_bwd_kernel<AT><<<blocks_for(Rc, 64), 64, 0, st>>>(SG + geo, D.ldS, RGB, D.nr, w->color2_b, ray_batch, z_vals, c.noise, r0, Rc, S, c.d_rgb,
                                                               c.d_depth, c.d_acc, c.d_weights, al_, tr, dRGB, dsig, c.d_ray_batch);
    // ---- color_net backward ----------------------------------------------------------------------------------------------
    EDN_RC(gemm.run(true, false, D.nr, hid, M, dRGB, D.nr, H3, hid, 1.f, gWp[3], hid));
    if (grad_w->color2_b) colsum_kernel<AT><<<blocks_for(M, 512), 32, 0, st>>>(dRGB, D.nr, 3, M, grad_w->color2_b);
    head_bwd_kernel<AT><<<blocks_for(M * (hid / al), 256), 256, 0, st>>>(dRGB, D.nr, w->color2, H3, hid, M, D1,
                                                                           (reinterpret_cast<uintptr_t>(w->color2) & 15) == 0 && hid % 4 == 0);   // D1 = dH3
    EDN_RC(gemm.run(true, false, hid, hid, M, D1, hid, H2, hid, 1.f, grad_w->color1, hid));
    if (grad_w->color1_b) colsum_kernel<AT><<<blocks_for(M, 512), 256, 0, st>>>(D1, hid, hid, M, grad_w->color1_b);
    EDN_RC(gemm.run(false, false, M, hid, hid, D1, hid, Wc1, hid, 0.f, D2, hid));                                       // D2 = dH2
    relu_mask_kernel<AT><<<blocks_for(M * (hid / al), 256), 256, 0, st>>>(D2, H2, hid, hid, M);
    EDN_RC(gemm.run(true, false, hid, D.ldS, M, D2, hid, SG, D.ldS, 1.f, gWp[2], D.ldS));
    if (grad_w->color0_b) colsum_kernel<AT><<<blocks_for(M, 512), 256, 0, st>>>(D2, hid, hid, M, grad_w->color0_b);
    EDN_RC(gemm.run(false, false, M, D.ldS, hid, D2, hid, Wp[2], D.ldS, 0.f, dSG, D.ldS));                              // [d geo | 0 | d PE(dir)]
    set_sigma_grad_kernel<AT><<<blocks_for(M, 256), 256, 0, st>>>(dSG, D.ldS, geo, M, dsig);
    if (c.d_feat) add_feat_grad_kernel<AT><<<blocks_for(M * ((geo + 3) / 4), 256), 256, 0, st>>>(dSG, D.ldS, geo, M, c.d_feat + m0 * geo);
    // ---- sigma_net backward ------------------------------------------------------------------------------------------------
    EDN_RC(gemm.run(true, false, D.sgn, hid, M, dSG, D.ldS, H1, hid, 1.f, gWp[1], hid));      // pad rows collect d PE(dir): dropped at fold-back
    EDN_RC(gemm.run(false, false, M, hid, D.sgn, dSG, D.ldS, Wp[1], hid, 0.f, D1, hid));                                // D1 = dH1 (pad rows of W are 0)
    relu_mask_kernel<AT><<<blocks_for(M * (hid / al), 256), 256, 0, st>>>(D1, H1, hid, hid, M);
    EDN_RC(gemm.run(true, false, hid, D.ldX, M, D1, hid, X0, D.ldX, 1.f, gWp[0], D.ldX));
    EDN_RC(gemm.run(false, false, M, D.ldX, hid, D1, hid, Wp[0], D.ldX, 0.f, dX0, D.ldX));
    // ---- inputs: PE(pts), basis_mat, VM grids ----------------------------------------------------------------------------------
    pe_bwd_kernel<float><<<blocks_for(M * 3, 256), 256, 0, st>>>(ray_batch, z_vals, m0, M, S, dX0, D.ldX, D.nf, dpts);
    convert_cols_kernel<AT><<<blocks_for(M * D.nf, 256), 256, 0, st>>>(dXf, D.nf, dX0, D.ldX, D.nf, M);
    for (int g = 0; g < ng; ++g) {
      EDN_RC(gemm.run(true, false, kAppDim, kAppComp, M, dXf + 32 * g, D.nf, P[g], kAppComp, 1.f, grad_w->basis[g], kAppComp));
      EDN_RC(gemm.run(false, false, M, kAppComp, kAppDim, dXf + 32 * g, D.nf, Wb[g], kAppComp, 0.f, dP, kAppComp));
      // fine field, coarse grid (g == 0 of 2): the coarse positions' rows move to the coarse pass; coarse field: they are added back
      const bool fine_call = ng == 2;
      const int64_t* skip = (c.moved && fine_call && g == 0) ? c.merge_order : nullptr;
      const AT* dP2 = (c.moved && !fine_call) ? reinterpret_cast<const AT*>(c.moved) : nullptr;
      int64_t Ms = M;                   // rows the scatter walks
      const int* rows = nullptr;
      if (skip) {
        constexpr int kPieces = kAppComp * (int)sizeof(AT) / 16;
        move_rows_kernel<AT><<<blocks_for(M * kPieces, 256), 256, 0, st>>>(dP, skip, m0, M, S, c.n_coarse, reinterpret_cast<AT*>(c.moved), row_list);
        Ms = Rc * (S - c.n_coarse);
        rows = row_list;
      }
      if (c.grids[g]->dtype == EDN_F32) vm_scatter_kernel<float, AT><<<blocks_for(Ms, kSamplesPerBlock), kVmThreads, 0, st>>>(c.gd[g], c.gg[g], ray_batch, z_vals, m0, Ms, S, dP, dpts, rows, dP2);
      else vm_scatter_kernel<__nv_bfloat16, AT><<<blocks_for(Ms, kSamplesPerBlock), kVmThreads, 0, st>>>(c.gd[g], c.gg[g], ray_batch, z_vals, m0, Ms, S, dP, dpts, rows, dP2);
    }
    ray_reduce_kernel<AT><<<blocks_for(Rc * 32, 256), 256, 0, st>>>(dpts, dSG, D.ldS, geo, ray_batch, z_vals, r0, Rc, S, c.d_ray_batch);
    EDN_CUDA_OK(cudaGetLastError());
  }
#undef EDN_RC
  relayout_fold_kernel<<<blocks_for(nW[0], 256), 256, 0, st>>>(gWp[0], grad_w->sigma0, 0, hid, D.ldX, hid, D.kin, geo);
  relayout_fold_kernel<<<blocks_for(nW[1], 256), 256, 0, st>>>(gWp[1], grad_w->sigma1, 1, D.sgn, hid, 1 + geo, hid, geo);
  relayout_fold_kernel<<<blocks_for(nW[2], 256), 256, 0, st>>>(gWp[2], grad_w->color0, 2, hid, D.ldS, hid, D.cin, geo);
  relayout_fold_kernel<<<blocks_for(nW[3], 256), 256, 0, st>>>(gWp[3], grad_w->color2, 3, D.nr, hid, 3, hid, geo);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

}  // namespace
}  // namespace edn

extern "C" int64_t edn_field_bwd_workspace_bytes(int32_t n_grids, int32_t hidden, int32_t geo_feat, int64_t chunk_rays,
                                                 int32_t n_samples) {
  using namespace edn;
  if (n_grids < 1 || n_grids > 2 || hidden <= 0 || geo_feat <= 0 || chunk_rays <= 0 || n_samples <= 0) return -1;
  const Dims d = make_dims(n_grids, hidden, geo_feat, 4);       // fp32 storage: the larger of the two layouts
  return bytes_per_sample(d) * chunk_rays * n_samples + weight_scratch_bytes(d);
}

extern "C" int edn_render_field_bwd(const edn_vm_grid* grid0, const edn_vm_grid* grid1, const edn_field_weights* w,
                                    const float* ray_batch, const float* z_vals, const float* noise, int64_t n_rays,
                                    int32_t n_samples, int32_t precision, const float* d_rgb, const float* d_depth,
                                    const float* d_acc, const float* d_weights, const float* d_feat,
                                    const edn_field_weights* grad_w, const edn_vm_grid_grad* grad_grid0,
                                    const edn_vm_grid_grad* grad_grid1, float* d_ray_batch, void* workspace,
                                    int64_t workspace_bytes, const edn_field_bwd_merge* merge, void* stream) {
  using namespace edn;
  EDN_REQUIRE(grid0 && w && grad_w && grad_grid0 && ray_batch && z_vals && d_ray_batch && workspace, "edn_render_field_bwd: null pointer");
  EDN_REQUIRE(n_samples >= 2, "edn_render_field_bwd: n_samples must be >= 2");
  const int ng = w->n_grids;
  EDN_REQUIRE(ng == 1 || ng == 2, "edn_render_field_bwd: n_grids must be 1 or 2");
  EDN_REQUIRE(ng == 1 || (grid1 && grad_grid1), "edn_render_field_bwd: n_grids = 2 needs grid1 and grad_grid1");
  EDN_REQUIRE(w->hidden % 8 == 0 && w->hidden > 0 && w->hidden <= 256 && w->geo_feat > 0, "edn_render_field_bwd: unsupported MLP dims");
  EDN_REQUIRE(w->sigma0 && w->sigma1 && w->color0 && w->color1 && w->color2 && w->basis[0] && (ng == 1 || w->basis[1]),
              "edn_render_field_bwd: null weight");
  EDN_REQUIRE(grad_w->sigma0 && grad_w->sigma1 && grad_w->color0 && grad_w->color1 && grad_w->color2 && grad_w->basis[0] &&
                  (ng == 1 || grad_w->basis[1]), "edn_render_field_bwd: null weight gradient");
  EDN_REQUIRE(!w->color0_b == !grad_w->color0_b && !w->color1_b == !grad_w->color1_b && !w->color2_b == !grad_w->color2_b,
              "edn_render_field_bwd: bias / bias-gradient mismatch");
  EDN_REQUIRE(precision == EDN_F32 || precision == EDN_BF16, "edn_render_field_bwd: bad precision");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  FieldBwdCall c{};
  c.grids[0] = grid0; c.grids[1] = grid1;
  const edn_vm_grid_grad* ggr[2] = {grad_grid0, grad_grid1};
  for (int g = 0; g < ng; ++g) {
    int rc = make_grid_dev(c.grids[g], &c.gd[g]);
    if (rc) return rc;
    EDN_REQUIRE(c.grids[g]->dtype == EDN_F32 || c.grids[g]->dtype == EDN_BF16, "edn_render_field_bwd: bad grid dtype");
    for (int i = 0; i < 3; ++i) {
      EDN_REQUIRE(ggr[g]->plane[i] && ggr[g]->line[i], "edn_render_field_bwd: null grid gradient");
      c.gg[g].plane[i] = ggr[g]->plane[i];
      c.gg[g].line[i] = ggr[g]->line[i];
    }
  }
  c.w = w; c.grad_w = grad_w; c.ray_batch = ray_batch; c.z_vals = z_vals; c.noise = noise; c.n_rays = n_rays; c.S = n_samples;
  c.d_rgb = d_rgb; c.d_depth = d_depth; c.d_acc = d_acc; c.d_weights = d_weights; c.d_feat = d_feat; c.d_ray_batch = d_ray_batch;
  c.workspace = workspace; c.workspace_bytes = workspace_bytes; c.st = reinterpret_cast<cudaStream_t>(stream);
  if (merge && merge->moved) {
    EDN_REQUIRE(merge->n_coarse >= 1, "edn_render_field_bwd: merge.n_coarse must be >= 1");
    EDN_REQUIRE(ng == 1 ? merge->n_coarse == n_samples : (merge->order != nullptr && merge->n_coarse < n_samples),
                "edn_render_field_bwd: merge needs the merged order (fine field) / n_coarse == n_samples (coarse field)");
    EDN_REQUIRE((reinterpret_cast<uintptr_t>(merge->moved) & 15) == 0, "edn_render_field_bwd: merge.moved must be 16-byte aligned");
    c.merge_order = merge->order; c.n_coarse = merge->n_coarse; c.moved = merge->moved;
  }
  // EDN_F32: fp32 activations, exact fp32 GEMMs (parity).  EDN_BF16: bf16 activations and GEMM operands, fp32 accumulation and fp32
  // weight gradients (the usual mixed-precision recipe: half the activation traffic, bf16 tensor-core GEMMs).
  c.ct = CUBLAS_COMPUTE_32F;
  return precision == EDN_F32 ? field_bwd_run<float>(c) : field_bwd_run<__nv_bfloat16>(c);
}

// =====================================================================================================================
// mode = nerf: backward of NeRF.mlpforward + NeRF.raw2outputs (networks/nerf.py:46-72, 131-162, 74-129) at pts = o + d * z_vals.
// Same recipe: recompute the 8 x 256 MLP (skip concat after layer 4), alpha / feature heads, view branch and rgb head of a
// chunk of rays as aligned row-major matrices, then walk back.  Per-sample buffers:
//   XH [320] = [PE(pts) (63) | 0 | h_4 (256)]  (the skip input of layer 5 in place)      H_l [256], l = 0..3, 5..7
//   AF [284] = [feature (256) | sigma | PE(viewdir) (27)]   (feature_linear and alpha_linear as ONE GEMM; views_linears reads it
//   through a zero column under sigma)                       HV [128]   RGB [4]
// =====================================================================================================================
namespace edn {
namespace {

constexpr int kNW = 256, kNXH = 320, kNAF = 284, kNAFn = 260, kNHV = 128;

inline int64_t nerf_floats_per_sample() {
  return kNXH * 2 /*XH, dXH*/ + kNW * 7 /*H0-3, H5-7*/ + kNAF * 2 /*AF, dAF*/ + kNHV * 2 /*HV, dHV*/ + 8 /*RGB, dRGB*/ + kNW * 2 /*D1, D2*/ + 4 /*dpts*/ + 3;
}
inline int64_t nerf_weight_scratch_floats() {   // padded weights + gradients: W0p [256][64], W5p [256][320], Waf [260][256], Wvp [128][284], Wrp [4][128], baf [260] (x2)
  return 2 * ((int64_t)kNW * 64 + (int64_t)kNW * kNXH + (int64_t)kNAFn * kNW + (int64_t)kNHV * kNAF + 4 * kNHV + kNAFn);
}

__global__ void add_bias_ld_kernel(float* __restrict__ Y, int ld, int n, int64_t M, const float* __restrict__ bias) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = t / n;
  const int j = (int)(t % n);
  if (m < M) Y[m * ld + j] += bias[j];
}
__global__ void add_vec_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] += src[t];
}
}  // namespace
}  // namespace edn

extern "C" int64_t edn_nerf_bwd_workspace_bytes(int64_t chunk_rays, int32_t n_samples) {
  using namespace edn;
  if (chunk_rays <= 0 || n_samples <= 0) return -1;
  return (nerf_floats_per_sample() * chunk_rays * n_samples + nerf_weight_scratch_floats()) * (int64_t)sizeof(float);
}

extern "C" int edn_nerf_field_bwd(const edn_nerf_weights* w, const float* ray_batch, const float* z_vals, const float* noise,
                                  int64_t n_rays, int32_t n_samples, int32_t flags, int32_t precision, const float* d_rgb,
                                  const float* d_depth, const float* d_acc, const float* d_weights, const float* d_feat,
                                  int32_t feature_after_linear, const edn_nerf_weights* g, float* d_ray_batch, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
  using namespace edn;
  EDN_REQUIRE(w && g && ray_batch && z_vals && d_ray_batch && workspace, "edn_nerf_field_bwd: null pointer");
  EDN_REQUIRE(n_samples >= 2, "edn_nerf_field_bwd: n_samples must be >= 2");
  EDN_REQUIRE(precision == EDN_F32 || precision == EDN_BF16, "edn_nerf_field_bwd: bad precision");
  for (int l = 0; l < 8; ++l) EDN_REQUIRE(w->pts_w[l] && w->pts_b[l] && g->pts_w[l] && g->pts_b[l], "edn_nerf_field_bwd: null pts_linears.%d", l);
  EDN_REQUIRE(w->alpha_w && w->alpha_b && w->feature_w && w->feature_b && w->views_w && w->views_b && w->rgb_w && g->alpha_w && g->alpha_b &&
              g->feature_w && g->feature_b && g->views_w && g->views_b && g->rgb_w && (!w->rgb_b == !g->rgb_b), "edn_nerf_field_bwd: null weight");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  const int S = n_samples;
  const int64_t per_ray = nerf_floats_per_sample() * S * (int64_t)sizeof(float);
  int64_t chunk = (workspace_bytes - nerf_weight_scratch_floats() * (int64_t)sizeof(float)) / per_ray;
  EDN_REQUIRE(chunk >= 1, "edn_nerf_field_bwd: workspace too small (%lld bytes, one ray needs %lld)", (long long)workspace_bytes, (long long)per_ray);
  chunk = chunk < n_rays ? chunk : n_rays;
  if (chunk * S > (1 << 22)) chunk = (1 << 22) / S;
  cublasHandle_t h = blas_handle();
  if (!h) return blas_unavailable();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cublasSetStream(h, st) != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetStream failed"); return EDN_E_CUDA; }
  const Gemm gemm{h, precision == EDN_F32 ? CUBLAS_COMPUTE_32F : CUBLAS_COMPUTE_32F_FAST_TF32};

  float* base = reinterpret_cast<float*>(workspace);
  const int nW[6] = {kNW * 64, kNW * kNXH, kNAFn * kNW, kNHV * kNAF, 4 * kNHV, kNAFn};
  float* Wp[6];
  float* gWp[6];
  for (int i = 0; i < 6; ++i) { Wp[i] = base; base += nW[i]; }
  for (int i = 0; i < 6; ++i) { gWp[i] = base; base += nW[i]; }
  int tot_w = 0;
  for (int i = 0; i < 6; ++i) tot_w += nW[i];
  EDN_CUDA_OK(cudaMemsetAsync(gWp[0], 0, sizeof(float) * (size_t)tot_w, st));
  // aligned copies: pts_linears.0 [256][63] -> [256][64]; pts_linears.5 [256][319] -> [256][320] (zero column after the PE part);
  // [feature_linear; alpha_linear; 0] -> [260][256]; views_linears.0 [128][283] -> [128][284] (zero column under sigma); rgb [3][128] -> [4][128]
  relayout_to_padded_kernel<float><<<blocks_for(nW[0], 256), 256, 0, st>>>(Wp[0], w->pts_w[0], 0, kNW, 64, kNW, 63, 0);
  relayout_to_padded_kernel<float><<<blocks_for(nW[1], 256), 256, 0, st>>>(Wp[1], w->pts_w[5], 2, kNW, kNXH, kNW, 319, 63);
  EDN_CUDA_OK(cudaMemsetAsync(Wp[2], 0, sizeof(float) * (size_t)nW[2], st));
  EDN_CUDA_OK(cudaMemcpyAsync(Wp[2], w->feature_w, sizeof(float) * kNW * kNW, cudaMemcpyDeviceToDevice, st));
  EDN_CUDA_OK(cudaMemcpyAsync(Wp[2] + kNW * kNW, w->alpha_w, sizeof(float) * kNW, cudaMemcpyDeviceToDevice, st));
  relayout_to_padded_kernel<float><<<blocks_for(nW[3], 256), 256, 0, st>>>(Wp[3], w->views_w, 2, kNHV, kNAF, kNHV, 283, 256);
  relayout_to_padded_kernel<float><<<blocks_for(nW[4], 256), 256, 0, st>>>(Wp[4], w->rgb_w, 3, 4, kNHV, 3, kNHV, 0);
  EDN_CUDA_OK(cudaMemsetAsync(Wp[5], 0, sizeof(float) * kNAFn, st));
  EDN_CUDA_OK(cudaMemcpyAsync(Wp[5], w->feature_b, sizeof(float) * kNW, cudaMemcpyDeviceToDevice, st));
  EDN_CUDA_OK(cudaMemcpyAsync(Wp[5] + kNW, w->alpha_b, sizeof(float), cudaMemcpyDeviceToDevice, st));

  const int64_t Mmax = chunk * S;
  auto take = [&](int64_t per) { float* p = base; base += per * Mmax; return p; };
  float* XH = take(kNXH);
  float* H[8];
  for (int l = 0; l < 8; ++l) H[l] = (l == 4) ? XH + 64 : take(kNW);
  const int ldH[8] = {kNW, kNW, kNW, kNW, kNXH, kNW, kNW, kNW};
  float* AF = take(kNAF);
  float* HV = take(kNHV);
  float* RGB = take(4);
  float* dRGB = take(4);
  float* dHV = take(kNHV);
  float* dAF = take(kNAF);
  float* dXH = take(kNXH);
  float* D1 = take(kNW);
  float* D2 = take(kNW);
  float* dpts = take(4);
  float* al = take(1);
  float* tr = take(1);
  float* dsig = take(1);
  // per-layer weights as used by the GEMMs: (pointer, K, ld)
  const float* Wl[8]; int Kl[8];
  for (int l = 0; l < 8; ++l) { Wl[l] = w->pts_w[l]; Kl[l] = kNW; }
  Wl[0] = Wp[0]; Kl[0] = 64; Wl[5] = Wp[1]; Kl[5] = kNXH;
  float* gWl[8];
  for (int l = 0; l < 8; ++l) gWl[l] = g->pts_w[l];
  gWl[0] = gWp[0]; gWl[5] = gWp[1];

#define EDN_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
  for (int64_t r0 = 0; r0 < n_rays; r0 += chunk) {
    const int64_t Rc = (n_rays - r0) < chunk ? (n_rays - r0) : chunk;
    const int64_t M = Rc * S, m0 = r0 * S;
    // ---- forward recompute ----------------------------------------------------------------------------------------------
    pe_kernel<<<blocks_for(M * (kPeFreqPts + 1), 256), 256, 0, st>>>(ray_batch, z_vals, m0, M, S, XH, kNXH, 0, 64, AF, kNAF, 256, 0);
    for (int l = 0; l < 8; ++l) {
      const float* X = (l == 0 || l == 5) ? XH : H[l - 1];
      const int ldx = (l == 0 || l == 5) ? kNXH : ldH[l - 1];
      float* Hl = H[l];
      const int ldh = ldH[l];
      const float* bl = w->pts_b[l];
      EDN_RC(gemm.relu_linear(M, kNW, Kl[l], X, ldx, Wl[l], Kl[l], bl, Hl, ldh, st,
                              [&] { relu_bias_kernel<<<blocks_for(M * (kNW / 4), 256), 256, 0, st>>>(Hl, ldh, kNW, M, bl); }));
    }
    EDN_RC(gemm(false, true, M, kNAFn, kNW, H[7], kNW, Wp[2], kNW, 0.f, AF, kNAF));
    add_bias_ld_kernel<<<blocks_for(M * 257, 256), 256, 0, st>>>(AF, kNAF, 257, M, Wp[5]);
    pe_kernel<<<blocks_for(M * (kPeFreqDir + 1), 256), 256, 0, st>>>(ray_batch, z_vals, m0, M, S, XH, kNXH, 0, 0, AF, kNAF, 256, 1);
    EDN_RC(gemm.relu_linear(M, kNHV, kNAF, AF, kNAF, Wp[3], kNAF, w->views_b, HV, kNHV, st,
                            [&] { relu_bias_kernel<<<blocks_for(M * (kNHV / 4), 256), 256, 0, st>>>(HV, kNHV, kNHV, M, w->views_b); }));
    EDN_RC(gemm(false, true, M, 4, kNHV, HV, kNHV, Wp[4], kNHV, 0.f, RGB, 4));
    // ---- compositing backward (nerf.py:74-129: sigma = channel 3, rgb = sigmoid) ------------------------------------------------
    composite_bwd_kernel<float><<<blocks_for(Rc, 64), 64, 0, st>>>(AF + 256, kNAF, RGB, 4, w->rgb_b, ray_batch, z_vals, noise, r0, Rc, S, d_rgb, d_depth,
                                                           d_acc, d_weights, al, tr, dRGB, dsig, d_ray_batch);
    // ---- heads ------------------------------------------------------------------------------------------------------------
    EDN_RC(gemm(true, false, 4, kNHV, M, dRGB, 4, HV, kNHV, 1.f, gWp[4], kNHV));
    if (g->rgb_b) colsum_kernel<<<blocks_for(M, 512), 32, 0, st>>>(dRGB, 4, 3, M, g->rgb_b);
    head_bwd_kernel<float><<<blocks_for(M * (kNHV / 4), 256), 256, 0, st>>>(dRGB, 4, w->rgb_w, HV, kNHV, M, dHV, (reinterpret_cast<uintptr_t>(w->rgb_w) & 15) == 0);
    EDN_RC(gemm(true, false, kNHV, kNAF, M, dHV, kNHV, AF, kNAF, 1.f, gWp[3], kNAF));
    colsum_kernel<<<blocks_for(M, 512), 128, 0, st>>>(dHV, kNHV, kNHV, M, g->views_b);
    EDN_RC(gemm(false, false, M, kNAF, kNHV, dHV, kNHV, Wp[3], kNAF, 0.f, dAF, kNAF));
    set_sigma_grad_kernel<<<blocks_for(M, 256), 256, 0, st>>>(dAF, kNAF, 256, M, dsig);
    if (d_feat && feature_after_linear) add_feat_grad_kernel<<<blocks_for(M * (kNW / 4), 256), 256, 0, st>>>(dAF, kNAF, kNW, M, d_feat + m0 * kNW);
    EDN_RC(gemm(true, false, kNAFn, kNW, M, dAF, kNAF, H[7], kNW, 1.f, gWp[2], kNW));
    colsum_kernel<<<blocks_for(M, 512), 256, 0, st>>>(dAF, kNAF, kNW, M, g->feature_b);
    colsum_kernel<<<blocks_for(M, 512), 32, 0, st>>>(dAF + kNW, kNAF, 1, M, g->alpha_b);
    EDN_RC(gemm(false, false, M, kNW, kNAFn, dAF, kNAF, Wp[2], kNW, 0.f, D1, kNW));
    if (d_feat && !feature_after_linear) add_feat_grad_kernel<<<blocks_for(M * (kNW / 4), 256), 256, 0, st>>>(D1, kNW, kNW, M, d_feat + m0 * kNW);
    // ---- the 8 x 256 trunk, layers 7..0 (D = gradient at the layer's post-activation) -------------------------------------------------
    float* D = D1;
    int ldD = kNW;
    for (int l = 7; l >= 0; --l) {
      relu_mask_kernel<<<blocks_for(M * (kNW / 4), 256), 256, 0, st>>>(D, H[l], ldD, kNW, M);      // ldD == ldH[l] (320 only for l = 4)
      const float* X = (l == 0 || l == 5) ? XH : H[l - 1];
      const int ldx = (l == 0 || l == 5) ? kNXH : ldH[l - 1];
      EDN_RC(gemm(true, false, kNW, Kl[l], M, D, ldD, X, ldx, 1.f, gWl[l], Kl[l]));
      colsum_kernel<<<blocks_for(M, 512), 256, 0, st>>>(D, ldD, kNW, M, g->pts_b[l]);
      if (l == 5) {          // d [x | h_4] in one GEMM; continue with the h_4 part in place
        EDN_RC(gemm(false, false, M, kNXH, kNW, D, ldD, Wl[5], kNXH, 0.f, dXH, kNXH));
        D = dXH + 64; ldD = kNXH;
      } else if (l == 0) {   // += the first layer's share of d x (the skip connection's share is already there)
        EDN_RC(gemm(false, false, M, 64, kNW, D, ldD, Wl[0], 64, 1.f, dXH, kNXH));
      } else {
        float* out = (D == D1) ? D2 : D1;     // l == 4 comes from dXH + 64 -> D1
        EDN_RC(gemm(false, false, M, kNW, kNW, D, ldD, Wl[l], kNW, 0.f, out, kNW));
        D = out; ldD = kNW;
      }
    }
    pe_bwd_kernel<<<blocks_for(M * 3, 256), 256, 0, st>>>(ray_batch, z_vals, m0, M, S, dXH, kNXH, 0, dpts);
    ray_reduce_kernel<<<blocks_for(Rc * 32, 256), 256, 0, st>>>(dpts, dAF, kNAF, 256, ray_batch, z_vals, r0, Rc, S, d_ray_batch);
    EDN_CUDA_OK(cudaGetLastError());
  }
#undef EDN_RC
  relayout_fold_kernel<<<blocks_for(nW[0], 256), 256, 0, st>>>(gWp[0], g->pts_w[0], 0, kNW, 64, kNW, 63, 0);
  relayout_fold_kernel<<<blocks_for(nW[1], 256), 256, 0, st>>>(gWp[1], g->pts_w[5], 2, kNW, kNXH, kNW, 319, 63);
  add_vec_kernel<<<blocks_for(kNW * kNW, 256), 256, 0, st>>>(g->feature_w, gWp[2], kNW * kNW);
  add_vec_kernel<<<blocks_for(kNW, 256), 256, 0, st>>>(g->alpha_w, gWp[2] + kNW * kNW, kNW);
  relayout_fold_kernel<<<blocks_for(nW[3], 256), 256, 0, st>>>(gWp[3], g->views_w, 2, kNHV, kNAF, kNHV, 283, 256);
  relayout_fold_kernel<<<blocks_for(nW[4], 256), 256, 0, st>>>(gWp[4], g->rgb_w, 3, 4, kNHV, 3, kNHV, 0);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

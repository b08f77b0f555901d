// DP-NeRF adaptive weight proposal (AWP), forward, fp32.
// Replaces AdaptiveWeightProposal.forward + feature_integration (networks/dpnerf/awp.py:79-117, 49-77) and
// MotionAggregationModule / CorrelationModule (networks/dpnerf/mam.py:56-84, 13-53), literally including the quirks
// of SURVEY Appendix B: the feature-integration cumprod runs over CHANNELS, its alpha is padded with zeros.
//   awp_sample_kernel (CTA per sub-ray): per-sample MLP 128->64->64->64->64 (ReLU) on depth_feature, feature integration,
//                                        MAM.linear 64->32, attention logits, softmax-over-samples pooling ("inter").
//   awp_ray_kernel    (CTA per primary ray): motion MLP 111->32->32, softmax-over-exposures pooling ("intra"), the 1x1
//                                        conv attention, convd -> pre-BatchNorm y.
//   awp_bn_partial / awp_bn_final / awp_out_kernel: train-mode BatchNorm1d batch statistics over (rays, exposures), residual +
//                                        leaky ReLU, average pool over exposures, sigmoid(w_linear), normalise.
// precision = EDN_BF16 replaces awp_sample_kernel (fp32 SIMT, the parity path) by the per-sample MLP as four tall TF32 GEMMs
// over all N*E*S samples (cuBLAS; plain [M,128|64] x [64] contractions) + awp_integrate_kernel (integration, attention logits,
// "inter" pooling) -- 15 ms -> ~3 ms on the headline batch -- and keeps the layer activations for the backward pass.
#include "awp_layout.cuh"
#include "bwd_common.cuh"

namespace edn {
namespace {

constexpr int kT = 256;
constexpr int kRows = 64;       // sample tile
constexpr int kLdX = 132, kLdY = 68;
constexpr int kMaxE = 16, kMaxS = 256;

struct AwpSmem {   // float offsets
  static constexpr int X = 0;                          // [64][132]
  static constexpr int Y = X + kRows * kLdX;           // [64][68]
  static constexpr int P = Y + kRows * kLdY;           // [64][68]
  static constexpr int W0 = P + kRows * kLdY;          // [128][64]
  static constexpr int W1 = W0 + 128 * 64;             // 3 x [64][64]
  static constexpr int B = W1 + 3 * 64 * 64;           // 4 x [64]
  static constexpr int Wm = B + 4 * 64;                // [64][32]
  static constexpr int bm = Wm + 64 * 32;              // [32]
  static constexpr int latt = bm + 32;                 // [32]
  static constexpr int xl = latt + 32;                 // [kMaxS][33]
  static constexpr int att = xl + kMaxS * 33;          // [kMaxS]
  static constexpr int carry = att + kMaxS;            // [64]
  static constexpr int gint = carry + 64;              // [64]
  static constexpr int red = gint + 64;                // [64]
  static constexpr int total = red + 64;
};

struct AwpArgs {
  edn_awp_params p;
  const float* depth_feature;  // [NE][S][128]
  const float* z_vals;         // [NE][S]
  const float* rays_d;         // [NE][3]   (row stride rays_d_stride floats)
  int rays_d_stride;
  const float* view_feature;   // [N][32]
  int64_t N;
  int E, S;
  float* gint;   // [NE][64]
  float* inter;  // [NE][32]
  float* xl;     // [NE][S][32]
  float* att;    // [NE][S]
  float* x;      // [N][E][32]
  float* y;      // [N][E][32]
  double* stats; // [32][2] sum, sum of squares
  float* ccw;    // [N][E]
  double bn_rows;  // rows behind the batch sums (N * E, over all ranks when the sums were all-reduced)
};

// Y[r][cg*16..] = relu(bias + X[r][0..K) . Wt[K][64]); 256 threads: row = tid/4, 16 columns per thread
__device__ __forceinline__ void layer64(const float* __restrict__ Xs, int ldx, int K, const float* __restrict__ Wt, const float* __restrict__ bias,
                                        float* __restrict__ Ys, int ldy, bool relu) {
  const int row = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 16;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = bias[c0 + j];
  for (int k = 0; k < K; k += 4) {
    const float4 a4 = *reinterpret_cast<const float4*>(Xs + row * ldx + k);
    const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4* w4 = reinterpret_cast<const float4*>(Wt + (k + q) * 64 + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w = w4[j];
        acc[4 * j + 0] = fmaf(av[q], w.x, acc[4 * j + 0]); acc[4 * j + 1] = fmaf(av[q], w.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(av[q], w.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(av[q], w.w, acc[4 * j + 3]);
      }
    }
  }
  __syncthreads();      // every thread finished reading Xs / Ys before Ys is overwritten (Ys may alias the previous input)
#pragma unroll
  for (int j = 0; j < 16; j += 4)
    *reinterpret_cast<float4*>(Ys + row * ldy + c0 + j) =
        relu ? make_float4(fmaxf(acc[j], 0.f), fmaxf(acc[j + 1], 0.f), fmaxf(acc[j + 2], 0.f), fmaxf(acc[j + 3], 0.f))
             : make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  __syncthreads();
}

__global__ void __launch_bounds__(kT, 1) awp_sample_kernel(const AwpArgs a) {
  extern __shared__ __align__(16) float sm[];
  using L = AwpSmem;
  const int tid = threadIdx.x, S = a.S;
  for (int i = tid; i < 128 * 64; i += kT) sm[L::W0 + i] = a.p.sample_t[0][i];
  for (int l = 1; l < 4; ++l)
    for (int i = tid; i < 64 * 64; i += kT) sm[L::W1 + (l - 1) * 4096 + i] = a.p.sample_t[l][i];
  for (int l = 0; l < 4; ++l)
    for (int i = tid; i < 64; i += kT) sm[L::B + l * 64 + i] = a.p.sample_b[l][i];
  for (int i = tid; i < 64 * 32; i += kT) sm[L::Wm + i] = a.p.mam_linear_t[i];
  for (int i = tid; i < 32; i += kT) { sm[L::bm + i] = a.p.mam_linear_b[i]; sm[L::latt + i] = a.p.line_conv_att[i]; }
  __syncthreads();
  const int64_t NE = a.N * a.E;
  const int n_tiles = (S + kRows - 1) / kRows;
  float* X = sm + L::X; float* Y = sm + L::Y; float* Pb = sm + L::P;
  for (int64_t sr = blockIdx.x; sr < NE; sr += gridDim.x) {
    const float* rd = a.rays_d + sr * a.rays_d_stride;
    const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])), __fmul_rn(rd[2], rd[2])));
    if (tid < 64) { sm[L::carry + tid] = 1.0f; sm[L::gint + tid] = 0.f; }
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int row0 = tile * kRows, valid = min(kRows, S - row0);
      __syncthreads();
      for (int i = tid; i < kRows * 32; i += kT) {          // depth_feature tile -> X (coalesced float4)
        const int r = i >> 5, c4 = (i & 31) * 4;
        const float4 v = r < valid ? __ldg(reinterpret_cast<const float4*>(a.depth_feature + ((size_t)sr * S + row0 + r) * 128 + c4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(X + r * kLdX + c4) = v;
      }
      __syncthreads();
      layer64(X, kLdX, 128, sm + L::W0, sm + L::B, Y, kLdY, true);               // sample_feature_embed_layer (awp.py:96-98)
      layer64(Y, kLdY, 64, sm + L::W1, sm + L::B + 64, X, kLdX, true);
      layer64(X, kLdX, 64, sm + L::W1 + 4096, sm + L::B + 128, Y, kLdY, true);
      layer64(Y, kLdY, 64, sm + L::W1 + 8192, sm + L::B + 192, X, kLdX, true);    // h_local -> X[:, 0:64]
      {  // MAM.linear 64 -> 32 (mam.py:75) : row = tid/4, 8 columns per thread
        const int row = tid >> 2, c0 = (tid & 3) * 8;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = sm[L::bm + c0 + j];
        for (int k = 0; k < 64; ++k) {
          const float av = X[row * kLdX + k];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(av, sm[L::Wm + k * 32 + c0 + j], acc[j]);
        }
        if (row < valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            sm[L::xl + (row0 + row) * 33 + c0 + j] = acc[j];
            a.xl[((size_t)sr * S + row0 + row) * 32 + c0 + j] = acc[j];
          }
        }
      }
      {  // feature integration, part 1: alpha[s][c] = 1 - exp(-h * dist), last sample padded with 0 (awp.py:64-66)
        const int row = tid >> 2, c0 = (tid & 3) * 16, s = row0 + row;
        float dist = 0.f;
        const bool has = (row < valid) && (s < S - 1);
        if (has) dist = __fmul_rn(a.z_vals[sr * S + s + 1] - a.z_vals[sr * S + s], dnorm);
#pragma unroll
        for (int j = 0; j < 16; ++j) Y[row * kLdY + c0 + j] = has ? 1.0f - expf(-__fmul_rn(X[row * kLdX + c0 + j], dist)) : 0.f;
      }
      __syncthreads();
      if (tid < kRows) {   // part 2: cumprod over CHANNELS of (1 - alpha) per sample row (awp.py:68-71)
        float p = 1.0f;
        for (int c = 0; c < 64; ++c) { p *= (1.0f - Y[tid * kLdY + c]); Pb[tid * kLdY + c] = p; }
        if (tid < valid) {  // attention logit: line_conv_att (mam.py:32)
          float t = 0.f;
          for (int c = 0; c < 32; ++c) t = fmaf(sm[L::latt + c], sm[L::xl + (row0 + tid) * 33 + c], t);
          sm[L::att + row0 + tid] = t;
          a.att[sr * S + row0 + tid] = t;
        }
      }
      __syncthreads();
      if (tid < 64) {      // part 3: weights[s][c] = alpha[s][c] * cumprod[s-1][c]; integrate over samples (awp.py:68-75)
        float g = sm[L::gint + tid], prev = sm[L::carry + tid];
        for (int r = 0; r < valid; ++r) {
          g = fmaf(Y[r * kLdY + tid] * prev, X[r * kLdX + tid], g);
          prev = Pb[r * kLdY + tid];
        }
        sm[L::gint + tid] = g;
        sm[L::carry + tid] = prev;
      }
    }
    __syncthreads();
    // "inter": sum_s xl[s][:] * softmax_s(att)   (mam.py:34)
    if (tid < 32) {
      float mx = -INFINITY;
      for (int s = tid; s < S; s += 32) mx = fmaxf(mx, sm[L::att + s]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
      for (int s = tid; s < S; s += 32) sum += expf(sm[L::att + s] - mx);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      float acc = 0.f;
      for (int s = 0; s < S; ++s) acc = fmaf(sm[L::xl + s * 33 + tid], expf(sm[L::att + s] - mx) / sum, acc);
      a.inter[sr * 32 + tid] = acc;
    }
    if (tid < 64) a.gint[sr * 64 + tid] = sm[L::gint + tid];
    __syncthreads();
  }
}

// Feature integration + attention logits + "inter" pooling of one sub-ray from its materialised h_local [S][64] and
// xl [S][32] (GEMM path); same arithmetic and summation order as awp_sample_kernel.  128 threads.
__global__ void __launch_bounds__(128) awp_integrate_kernel(const AwpArgs a, const float* __restrict__ h_all) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, S = a.S;
  float* Hs = sm;                 // [S][65]
  float* Al = Hs + S * 65;        // [S][65]
  float* Q = Al + S * 65;         // [S][65]
  float* xls = Q + S * 65;        // [S][33]
  float* atts = xls + S * 33;     // [S]
  const int64_t sr = blockIdx.x;
  const float* h = h_all + sr * S * 64;
  const float* xl = a.xl + sr * S * 32;
  for (int i = tid; i < S * 64; i += 128) Hs[(i >> 6) * 65 + (i & 63)] = h[i];
  for (int i = tid; i < S * 32; i += 128) xls[(i >> 5) * 33 + (i & 31)] = xl[i];
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
    float t = 0.f;
    for (int c = 0; c < 32; ++c) t = fmaf(__ldg(a.p.line_conv_att + c), xls[s * 33 + c], t);
    atts[s] = t;
    a.att[sr * S + s] = t;
  }
  __syncthreads();
  if (tid < 64) {
    float g = 0.f, prev = 1.0f;
    for (int s = 0; s < S; ++s) {
      g = fmaf(Al[s * 65 + tid] * prev, Hs[s * 65 + tid], g);
      prev = Q[s * 65 + tid];
    }
    a.gint[sr * 64 + tid] = g;
  } else if (tid < 96) {
    const int lane = tid - 64;
    float mx = -INFINITY;
    for (int s = lane; s < S; s += 32) mx = fmaxf(mx, atts[s]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int s = lane; s < S; s += 32) sum += expf(atts[s] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc = fmaf(xls[s * 33 + lane], expf(atts[s] - mx) / sum, acc);
    a.inter[sr * 32 + lane] = acc;
  }
}

// Same computation for S <= 128 with FOUR threads per sample (16 channels each, 512 threads): the channel cumprod is a local
// prefix product + a 3-step shuffle across the row's four lanes, the sums over samples are warp-shuffle + shared-memory
// reductions.  (The products are associated differently from the sequential kernels: last-bit differences.)
__global__ void __launch_bounds__(512, 2) awp_integrate4_kernel(const AwpArgs a, const float* __restrict__ h_all) {
  __shared__ float Qs[128][64];
  __shared__ float atts[128];
  __shared__ float red[16][64];
  __shared__ float redm[16], reds[16];
  const int tid = threadIdx.x, S = a.S, r = tid >> 2, q = tid & 3, c0 = q * 16, warp = tid >> 5, lane = tid & 31;
  const int64_t sr = blockIdx.x;
  const bool valid = r < S, has = r < S - 1;
  const float* rd = a.rays_d + sr * a.rays_d_stride;
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])), __fmul_rn(rd[2], rd[2])));
  float h[16], al[16], Q[16], xl[8];
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(h_all + (sr * S + r) * 64 + c0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    h[i] = v.x; h[i + 1] = v.y; h[i + 2] = v.z; h[i + 3] = v.w;
  }
#pragma unroll
  for (int i = 0; i < 8; i += 4) {
    const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(a.xl + (sr * S + r) * 32 + q * 8 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    xl[i] = v.x; xl[i + 1] = v.y; xl[i + 2] = v.z; xl[i + 3] = v.w;
  }
  const float dist = has ? __fmul_rn(a.z_vals[sr * S + r + 1] - a.z_vals[sr * S + r], dnorm) : 0.f;
  float p = 1.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    al[i] = has ? 1.0f - expf(-__fmul_rn(h[i], dist)) : 0.f;
    p *= (1.0f - al[i]);
    Q[i] = p;
  }
  {   // exclusive product over the lower quarters of the row
    const float t1 = __shfl_up_sync(0xffffffffu, p, 1), t2 = __shfl_up_sync(0xffffffffu, p, 2), t3 = __shfl_up_sync(0xffffffffu, p, 3);
    const float e = (q >= 1 ? t1 : 1.0f) * (q >= 2 ? t2 : 1.0f) * (q >= 3 ? t3 : 1.0f);
#pragma unroll
    for (int i = 0; i < 16; ++i) Q[i] *= e;
  }
#pragma unroll
  for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(&Qs[r][c0 + i]) = make_float4(Q[i], Q[i + 1], Q[i + 2], Q[i + 3]);
  {   // attention logit of the row
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t = fmaf(__ldg(a.p.line_conv_att + q * 8 + i), xl[i], t);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    if (q == 0) { atts[r] = valid ? t : -INFINITY; if (valid) a.att[sr * S + r] = t; }
  }
  __syncthreads();
  // integrate: g[c] = sum_s al[s][c] Q[s-1][c] h[s][c]
  float g[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float qp = r > 0 ? Qs[r - 1][c0 + i] : 1.0f;
    g[i] = valid ? al[i] * qp * h[i] : 0.f;
  }
  // softmax over the samples
  float mx = atts[min(r, 127)];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) redm[warp] = mx;
  __syncthreads();
  mx = redm[0];
#pragma unroll
  for (int w = 1; w < 16; ++w) mx = fmaxf(mx, redm[w]);
  const float ex = valid ? expf(atts[r] - mx) : 0.f;
  float sm = (q == 0) ? ex : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
  if (lane == 0) reds[warp] = sm;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 16; ++w) tot += reds[w];
  const float pr = ex / tot;
  float xi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xi[i] = xl[i] * pr;
  // sums over the rows of the warp (lanes with the same quarter), then over the 16 warps
#pragma unroll
  for (int o = 4; o <= 16; o <<= 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) g[i] += __shfl_xor_sync(0xffffffffu, g[i], o);
#pragma unroll
    for (int i = 0; i < 8; ++i) xi[i] += __shfl_xor_sync(0xffffffffu, xi[i], o);
  }
  if (lane < 4) {
#pragma unroll
    for (int i = 0; i < 16; ++i) red[warp][c0 + i] = g[i];
  }
  __syncthreads();
  if (tid < 64) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) t += red[w][tid];
    a.gint[sr * 64 + tid] = t;
  }
  __syncthreads();
  if (lane < 4) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][q * 8 + i] = xi[i];
  }
  __syncthreads();
  if (tid < 32) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) t += red[w][tid];
    a.inter[sr * 32 + tid] = t;
  }
}

// ---- per primary ray ---------------------------------------------------------------------------------------------------
struct RaySmem {
  float view[48];                 // img_embed (32) | PE(viewdir, L = 2) (15)
  float x[kMaxE][32];             // motion features
  float inter_a[16][kMaxE];
  float xlog[kMaxE][16];
  float inter_n[kMaxE][16];
  float x_inter[kMaxE][kMaxE];
  float cf[kMaxE][32];
  float intra[32][kMaxS + 1];     // padded (written channel-major, read sample-major); later reused flat as intra_l [S][16]
  float intra_b[16][kMaxS];
  float x_intra[kMaxE][kMaxS];
};

constexpr int kRayThreads = 256;   // short shared-memory-latency-bound phases with few busy threads each: 3 CTAs of 8 warps per SM
__global__ void __launch_bounds__(kRayThreads, 3) awp_ray_kernel(const AwpArgs a) {
  extern __shared__ __align__(16) float smraw[];
  RaySmem& s = *reinterpret_cast<RaySmem*>(smraw);
  const int tid = threadIdx.x, E = a.E, S = a.S;
  const int64_t n = blockIdx.x;
  // view vector (awp.py:88-94): latent | PE of the normalised direction of exposure 0
  if (tid < 32) s.view[tid] = a.view_feature[n * 32 + tid];
  if (tid == 32) {
    const float* rd = a.rays_d + (n * E) * a.rays_d_stride;
    const float nr = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
    const float v[3] = {rd[0] / nr, rd[1] / nr, rd[2] / nr};
    for (int i = 0; i < 3; ++i) {
      s.view[32 + i] = v[i];
      for (int f = 0; f < 2; ++f) { s.view[35 + 6 * f + i] = sinf(v[i] * (float)(1 << f)); s.view[38 + 6 * f + i] = cosf(v[i] * (float)(1 << f)); }
    }
  }
  __syncthreads();
  // motion_feature_embed_layer: [gint(64) | view(47)] = 111 -> 32 -> 32, ReLU both (awp.py:104-108)
  for (int i = tid; i < E * 32; i += kRayThreads) {
    const int e = i >> 5, j = i & 31;
    float acc = a.p.motion_b[0][j];
    const float* w = a.p.motion_w[0] + j * 111;
    const float* g = a.gint + (n * E + e) * 64;
    for (int k = 0; k < 64; ++k) acc = fmaf(w[k], g[k], acc);
    for (int k = 0; k < 47; ++k) acc = fmaf(w[64 + k], s.view[k], acc);
    s.cf[e][j] = fmaxf(acc, 0.f);     // staging
  }
  __syncthreads();
  for (int i = tid; i < E * 32; i += kRayThreads) {
    const int e = i >> 5, j = i & 31;
    float acc = a.p.motion_b[1][j];
    const float* w = a.p.motion_w[1] + j * 32;
    for (int k = 0; k < 32; ++k) acc = fmaf(w[k], s.cf[e][k], acc);
    s.x[e][j] = fmaxf(acc, 0.f);
  }
  // "intra": sum_e xl[e][s][:] * softmax_e(att[:, s])   (mam.py:35)
  for (int sp = tid; sp < S; sp += kRayThreads) {           // softmax over exposures -> x_intra (free until the logits below)
    float mx = -INFINITY, pe[kMaxE];
    for (int e = 0; e < E; ++e) { pe[e] = a.att[(n * E + e) * S + sp]; mx = fmaxf(mx, pe[e]); }
    float sum = 0.f;
    for (int e = 0; e < E; ++e) { pe[e] = expf(pe[e] - mx); sum += pe[e]; }
    for (int e = 0; e < E; ++e) s.x_intra[e][sp] = pe[e] / sum;
  }
  __syncthreads();
  for (int i = tid; i < 32 * S; i += kRayThreads) {         // lanes = channels: coalesced reads of xl
    const int c = i & 31, sp = i >> 5;
    float acc = 0.f;
    for (int e = 0; e < E; ++e) acc = fmaf(a.xl[((n * E + e) * S + sp) * 32 + c], s.x_intra[e][sp], acc);
    s.intra[c][sp] = acc;
  }
  __syncthreads();
  // conva on inter (mam.py:37), convc on x (mam.py:40)
  for (int i = tid; i < 16 * E; i += kRayThreads) {
    const int k = i / E, e = i % E;
    float acc = 0.f, acc2 = 0.f;
    for (int c = 0; c < 32; ++c) { acc = fmaf(a.p.conva[k * 32 + c], a.inter[(n * E + e) * 32 + c], acc); acc2 = fmaf(a.p.convc[k * 32 + c], s.x[e][c], acc2); }
    s.inter_a[k][e] = acc;
    s.xlog[e][k] = acc2;
  }
  // convb on intra (mam.py:38)
  for (int i = tid; i < 16 * S; i += kRayThreads) {
    const int k = i / S, sp = i % S;
    float acc = 0.f;
    for (int c = 0; c < 32; ++c) acc = fmaf(a.p.convb[k * 32 + c], s.intra[c][sp], acc);
    s.intra_b[k][sp] = acc;
  }
  __syncthreads();
  // x_inter = softmax_e'(xlog . inter_a), convn (mam.py:41, 44)
  if (tid < E) {
    const int e = tid;
    float lg[kMaxE], mx = -INFINITY, sum = 0.f;
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
  // x_intra logits (mam.py:42): [E][S]
  for (int i = tid; i < E * S; i += kRayThreads) {
    const int e = i / S, sp = i % S;
    float t = 0.f;
    for (int k = 0; k < 16; ++k) t = fmaf(s.xlog[e][k], s.intra_b[k][sp], t);
    s.x_intra[e][sp] = t;
  }
  __syncthreads();
  {  // softmax over samples, one warp per exposure row
    const int warp = tid >> 5, lane = tid & 31;
    for (int e = warp; e < E; e += kRayThreads / 32) {
      float mx = -INFINITY;
      for (int sp = lane; sp < S; sp += 32) mx = fmaxf(mx, s.x_intra[e][sp]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
      for (int sp = lane; sp < S; sp += 32) { const float v = expf(s.x_intra[e][sp] - mx); s.x_intra[e][sp] = v; sum += v; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      for (int sp = lane; sp < S; sp += 32) s.x_intra[e][sp] /= sum;
    }
  }
  // convl on intra_b (mam.py:45): intra_l[s][k] -> reuse s.intra as [S][16]
  float* intra_l = &s.intra[0][0];
  __syncthreads();
  for (int i = tid; i < S * 16; i += kRayThreads) {
    const int sp = i >> 4, k = i & 15;
    float t = 0.f;
    for (int c = 0; c < 16; ++c) t = fmaf(a.p.convl[k * 16 + c], s.intra_b[c][sp], t);
    intra_l[sp * 16 + k] = t;
  }
  __syncthreads();
  // curve features (mam.py:47-50): cf[e] = [x_inter . inter_n | x_intra . intra_l]
  for (int i = tid; i < E * 32; i += kRayThreads) {
    const int e = i >> 5, j = i & 31;
    float t = 0.f;
    if (j < 16) { for (int e2 = 0; e2 < E; ++e2) t = fmaf(s.x_inter[e][e2], s.inter_n[e2][j], t); }
    else { for (int sp = 0; sp < S; ++sp) t = fmaf(s.x_intra[e][sp], intra_l[sp * 16 + (j - 16)], t); }
    s.cf[e][j] = t;
  }
  __syncthreads();
  // convd.0 (mam.py:51, before BatchNorm): y[e][c]; also export x
  for (int i = tid; i < E * 32; i += kRayThreads) {
    const int e = i >> 5, c = i & 31;
    float t = 0.f;
    for (int k = 0; k < 32; ++k) t = fmaf(a.p.convd_w[c * 32 + k], s.cf[e][k], t);
    a.y[(n * E + e) * 32 + c] = t;
    a.x[(n * E + e) * 32 + c] = s.x[e][c];
  }
}

// BatchNorm1d batch statistics over (N, E) per channel (train mode, mam.py:24-27): sum and sum of squares in fp64, as a two-level
// reduction in a FIXED order (64 blocks of contiguous row ranges, then one block over the partials): deterministic, and 20x faster than
// the single block that walked all rows (0.19 ms of the 5.7 ms shipped-configuration forward).
constexpr int kBnBlocks = 64;
__global__ void __launch_bounds__(256) awp_bn_partial_kernel(const float* __restrict__ y, int64_t rows, double* __restrict__ part) {
  const int c = threadIdx.x & 31, slice = threadIdx.x >> 5;      // 32 channels x 8 row slices
  const int64_t per = (rows + kBnBlocks - 1) / kBnBlocks, r0 = blockIdx.x * per, r1 = min(r0 + per, rows);
  double s1 = 0.0, s2 = 0.0;
  for (int64_t r = r0 + slice; r < r1; r += 8) { const double v = y[r * 32 + c]; s1 += v; s2 += v * v; }
  __shared__ double sh[2][8][32];
  sh[0][slice][c] = s1; sh[1][slice][c] = s2;
  __syncthreads();
  if (slice == 0) {
    for (int p = 1; p < 8; ++p) { s1 += sh[0][p][c]; s2 += sh[1][p][c]; }
    part[blockIdx.x * 64 + 2 * c] = s1; part[blockIdx.x * 64 + 2 * c + 1] = s2;
  }
}
__global__ void awp_bn_final_kernel(const double* __restrict__ part, int64_t rows, double* __restrict__ stats) {
  const int t = threadIdx.x;
  if (t < 64) {
    double s = 0.0;
    for (int b = 0; b < kBnBlocks; ++b) s += part[b * 64 + t];
    stats[t] = s;
  }
  if (t == 0) { stats[64] = (double)rows; stats[65] = 0.0; }     // the row count travels (and is all-reduced) with the sums
}

__global__ void awp_out_kernel(const AwpArgs a, float bn_eps) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const int E = a.E;
  const double rows = a.bn_rows > 0.0 ? a.bn_rows : a.stats[64];
  float pooled[32];
  for (int c = 0; c < 32; ++c) {
    const double mean = a.stats[2 * c] / rows;
    const double var = a.stats[2 * c + 1] / rows - mean * mean;      // biased variance (normalisation in train mode)
    const float inv = rsqrtf((float)var + bn_eps), mu = (float)mean;
    float acc = 0.f;
    for (int e = 0; e < E; ++e) {
      const float yn = (a.y[(n * E + e) * 32 + c] - mu) * inv * a.p.bn_weight[c] + a.p.bn_bias[c];
      const float v = a.x[(n * E + e) * 32 + c] + yn;
      acc += v > 0.f ? v : 0.2f * v;                                   // leaky_relu(0.2), mam.py:53
    }
    pooled[c] = acc / (float)E;                                        // adaptive_avg_pool1d over exposures, awp.py:112
  }
  float w[kMaxE], tot = 0.f;
  for (int e = 0; e < E; ++e) {
    float t = a.p.w_linear_b[e];
    for (int c = 0; c < 32; ++c) t = fmaf(a.p.w_linear_w[e * 32 + c], pooled[c], t);
    w[e] = sigmoidf_(t);
    tot += w[e];
  }
  for (int e = 0; e < E; ++e) a.ccw[n * E + e] = w[e] / tot;
}

}  // namespace
}  // namespace edn

extern "C" int64_t edn_awp_workspace_floats(int64_t n_rays, int32_t n_exposure, int32_t n_samples, const edn_awp_options* opt) {
  return edn::awp_ws_floats(n_rays, n_exposure, n_samples, opt && (opt->precision == EDN_BF16 || opt->keep_activations));
}

namespace edn {
int awp_forward(const edn_awp_params* p, const float* depth_feature, const float* z_vals, const float* rays_d, int32_t rays_d_stride,
                const float* view_feature, int64_t n_rays, int32_t n_exposure, int32_t n_samples, float bn_eps, bool gemm_path,
                bool tf32, int phase, int64_t bn_rows_total, float* workspace, float* ccw, void* stream, bool keep_all) {
  EDN_REQUIRE(phase >= 0 && phase <= 2, "edn_awp_fwd: phase must be 0, 1 or 2");
  EDN_REQUIRE(p && depth_feature && z_vals && rays_d && view_feature && workspace && ccw, "edn_awp_fwd: null pointer");
  EDN_REQUIRE(n_exposure >= 1 && n_exposure <= kMaxE && n_samples >= 2 && n_samples <= kMaxS,
              "edn_awp_fwd: need 1 <= E <= %d and 2 <= S <= %d", kMaxE, kMaxS);
  for (int l = 0; l < 4; ++l) EDN_REQUIRE(p->sample_t[l] && p->sample_b[l], "edn_awp_fwd: null sample_feature_embed_layer.%d", l);
  EDN_REQUIRE(p->motion_w[0] && p->motion_b[0] && p->motion_w[1] && p->motion_b[1] && p->mam_linear_t && p->mam_linear_b &&
              p->line_conv_att && p->conva && p->convb && p->convc && p->convn && p->convl && p->convd_w && p->bn_weight && p->bn_bias &&
              p->w_linear_w && p->w_linear_b, "edn_awp_fwd: null weight");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  const int in_ch = p->input_ch > 0 ? p->input_ch : 128;
  EDN_REQUIRE(in_ch % 4 == 0 && in_ch <= 1024, "edn_awp_fwd: input_ch must be a multiple of 4 up to 1024, got %d", in_ch);
  EDN_REQUIRE(in_ch == 128 || gemm_path, "edn_awp_fwd: input_ch = %d runs the materialised path only: set keep_activations", in_ch);
  if (in_ch != 128) tf32 = false;        // the tcgen05 sample MLP (awp_tc.cu) is built for the 128-channel c2f features
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t NE = n_rays * n_exposure;
  AwpArgs a{};
  a.p = *p; a.depth_feature = depth_feature; a.z_vals = z_vals; a.rays_d = rays_d; a.rays_d_stride = rays_d_stride;
  a.view_feature = view_feature; a.N = n_rays; a.E = n_exposure; a.S = n_samples; a.ccw = ccw;
  const AwpWs ws = awp_ws_carve(workspace, n_rays, n_exposure, n_samples, gemm_path);
  a.gint = ws.gint; a.inter = ws.inter; a.xl = ws.xl; a.att = ws.att; a.x = ws.x; a.y = ws.y; a.stats = ws.stats;
  a.bn_rows = (double)(bn_rows_total > 0 ? bn_rows_total : (phase == 0 ? NE : 0));   // 0: read the all-reduced count from the workspace
  if (phase == 2) {      // finish from the (all-reduced) batch sums already in the workspace
    awp_out_kernel<<<(unsigned)((n_rays + 127) / 128), 128, 0, st>>>(a, bn_eps);
    EDN_CUDA_OK(cudaGetLastError());
    return EDN_OK;
  }
  if (!gemm_path) {
    const size_t smem1 = AwpSmem::total * sizeof(float);
    EDN_CUDA_OK(cudaFuncSetAttribute(awp_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    const int64_t g1 = NE < (int64_t)num_sms() ? NE : (int64_t)num_sms();
    awp_sample_kernel<<<(unsigned)g1, kT, smem1, st>>>(a);
  } else if (tf32) {
    // EDN_BF16 precision: the five per-sample contractions on tcgen05, activations on chip (awp_tc.cu)
    EDN_REQUIRE((reinterpret_cast<uintptr_t>(depth_feature) & 15) == 0, "edn_awp_fwd: depth_feature must be 16-byte aligned");
    const int64_t M = NE * n_samples;
    static const bool use_blas = [] { const char* e = getenv("EDN_AWP_BLAS"); return e && e[0] == '1'; }();   // dev switch: round-1 TF32 GEMM chain
    if (use_blas) return awp_forward(p, depth_feature, z_vals, rays_d, rays_d_stride, view_feature, n_rays, n_exposure, n_samples, bn_eps, true,
                                     false, phase, bn_rows_total, workspace, ccw, stream, keep_all);
    int rc = awp_sample_mlp_tc(p, depth_feature, M, ws, keep_all ? 1 : 0, st);
    if (rc) return rc;
    if (n_samples <= 128) {
      awp_integrate4_kernel<<<(unsigned)NE, 512, 0, st>>>(a, ws.act[3]);
    } else {
      const size_t smem_i = sizeof(float) * (size_t)(n_samples * (3 * 65 + 33 + 1));
      EDN_CUDA_OK(cudaFuncSetAttribute(awp_integrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i));
      awp_integrate_kernel<<<(unsigned)NE, 128, smem_i, st>>>(a, ws.act[3]);
    }
  } else {
    EDN_REQUIRE((reinterpret_cast<uintptr_t>(depth_feature) & 15) == 0, "edn_awp_fwd: depth_feature must be 16-byte aligned");
    cublasHandle_t h = blas_handle();
    if (!h) return blas_unavailable();
    if (cublasSetStream(h, st) != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetStream failed"); return EDN_E_CUDA; }
    const Gemm gemm{h, tf32 ? CUBLAS_COMPUTE_32F_FAST_TF32 : CUBLAS_COMPUTE_32F};
    const int64_t M = NE * n_samples;
    EDN_REQUIRE(M < (int64_t)1 << 31, "edn_awp_fwd: too many samples for one GEMM");
    const float* X = depth_feature;
    int K = in_ch;
    for (int l = 0; l < 4; ++l) {
      float* act = ws.act[l];
      const float* bias = p->sample_b[l];
      int rc = gemm.relu_linear(M, 64, K, X, K, p->sample_t[l], 64, bias, act, 64, st,
                                [&] { relu_bias_kernel<<<blocks_for(M * 16, 256), 256, 0, st>>>(act, 64, 64, M, bias); }, /*w_kn=*/true);
      if (rc) return rc;
      X = ws.act[l];
      K = 64;
    }
    int rc = gemm(false, false, M, 32, 64, ws.act[3], 64, p->mam_linear_t, 32, 0.f, ws.xl, 32);
    if (rc) return rc;
    add_bias_kernel<<<blocks_for(M * 32, 256), 256, 0, st>>>(ws.xl, 32, M, p->mam_linear_b);
    if (n_samples <= 128) {
      awp_integrate4_kernel<<<(unsigned)NE, 512, 0, st>>>(a, ws.act[3]);
    } else {
      const size_t smem_i = sizeof(float) * (size_t)(n_samples * (3 * 65 + 33 + 1));
      EDN_CUDA_OK(cudaFuncSetAttribute(awp_integrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i));
      awp_integrate_kernel<<<(unsigned)NE, 128, smem_i, st>>>(a, ws.act[3]);
    }
  }
  const size_t smem2 = sizeof(RaySmem);
  EDN_CUDA_OK(cudaFuncSetAttribute(awp_ray_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  awp_ray_kernel<<<(unsigned)n_rays, kRayThreads, smem2, st>>>(a);
  awp_bn_partial_kernel<<<kBnBlocks, 256, 0, st>>>(a.y, NE, ws.bn_part);
  awp_bn_final_kernel<<<1, 64, 0, st>>>(ws.bn_part, NE, a.stats);
  if (phase == 0) awp_out_kernel<<<(unsigned)((n_rays + 127) / 128), 128, 0, st>>>(a, bn_eps);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}
}  // namespace edn

extern "C" int edn_awp_fwd(const edn_awp_params* p, const float* depth_feature, const float* z_vals, const float* rays_d,
                           int32_t rays_d_stride, const float* view_feature, int64_t n_rays, int32_t n_exposure, int32_t n_samples,
                           float bn_eps, const edn_awp_options* opt, float* workspace, float* ccw, void* stream) {
  using namespace edn;
  EDN_REQUIRE(opt && (opt->precision == EDN_F32 || opt->precision == EDN_BF16), "edn_awp_fwd: bad options");
  return awp_forward(p, depth_feature, z_vals, rays_d, rays_d_stride, view_feature, n_rays, n_exposure, n_samples, bn_eps,
                     opt->precision == EDN_BF16 || opt->keep_activations, opt->precision == EDN_BF16, opt->phase, opt->bn_rows_total,
                     workspace, ccw, stream, opt->keep_activations != 0);
}

extern "C" int64_t edn_awp_stats_offset_floats(int64_t n_rays, int32_t n_exposure, int32_t n_samples) {
  float* zero = nullptr;
  const edn::AwpWs ws = edn::awp_ws_carve(zero, n_rays, n_exposure, n_samples, false);
  return reinterpret_cast<float*>(ws.stats) - zero;
}

// Backward of the loss-path kernels in loss.cu (what autograd does for the reference at run_nerf.py:448-504, 539-594):
//   edn_crf_bwd          <- CRF.forward + encode_rgb / encode_luma   networks/tonemapping.py:59-93, 111-139
//   edn_egm_loss_bwd     <- egm_loss                                 utils/events.py:260-284
//   edn_img2mse_bwd      <- img2mse                                  utils/metrics.py:7
//   edn_tv_loss_app_bwd  <- VoxelNeRFBase.TV_loss_app / TVLoss       networks/pdrf/voxnerf.py:126-130, 306-324
// Upstream loss gradients arrive as DEVICE scalars (d_loss[0]) so no host synchronisation is needed.
#include "common.cuh"

namespace edn {
namespace {

// ---- CRF ----------------------------------------------------------------------------------------------------------------
// One thread per sample runs the 1 -> 16 -> 16 -> 16 -> 1 residual MLP forward and backward for its three colour channels.  The
// parameter gradients are outer products summed over samples: each warp stages its 32 samples' activations / deltas in shared
// memory and the lanes split the 705 parameters (22 each), looping over the 32 samples -- no shared-memory atomics (fp32 ones
// are CAS loops: the first version of this kernel spent 1.4 ms per launch in them) -- then one global atomicAdd per
// parameter and warp.
constexpr int kCrfWarps = 2;
struct CrfStage {          // per warp, per sample (column = lane)
  float in[8][32], h0[16][32], h1[16][32], h2[16][32];
  float d0[16][32], d1[16][32], d2[16][32], dout[32];
};

struct CrfBwdArgs {
  edn_crf_params p;
  edn_crf_grads g;
  const float* x;
  const float* feat;
  int feat_per_channel;
  int flags;
  int64_t M;
  const float* d_out;   // [M][3] or [M][1]
  float* d_x;           // [M][3]
};

__global__ void __launch_bounds__(kCrfWarps * 32) crf_bwd_kernel(const CrfBwdArgs a) {
  __shared__ float w0[16 * 8], b0[16], w1[256], b1[16], w2[256], b2[16], w3[16], b3[1];
  __shared__ CrfStage stage[kCrfWarps];
  const int F = a.p.extra_features, in_ch = 1 + F;
  const bool learn = (a.flags & EDN_CRF_LEARN) && !(a.flags & EDN_CRF_SKIP_LEARN);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (learn) {
    for (int i = threadIdx.x; i < 16 * in_ch; i += blockDim.x) w0[i] = a.p.w0[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { w1[i] = a.p.w1[i]; w2[i] = a.p.w2[i]; }
    for (int i = threadIdx.x; i < 16; i += blockDim.x) { b0[i] = a.p.b0[i]; b1[i] = a.p.b1[i]; b2[i] = a.p.b2[i]; w3[i] = a.p.w3[i]; }
    if (threadIdx.x == 0) b3[0] = a.p.b3[0];
    __syncthreads();
  }
  const bool want_pg = learn && a.g.w0;
  CrfStage& st = stage[warp];
  // parameter p of this lane's share: index space [w0 (16*in_ch) | b0 16 | w1 256 | b1 16 | w2 256 | b2 16 | w3 16 | b3 1]
  const int n_w0 = 16 * in_ch, n_par = n_w0 + 16 + 256 + 16 + 256 + 16 + 16 + 1;
  constexpr int kPerLane = 23;                         // ceil(705 / 32)
  float pg[kPerLane];
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) pg[i] = 0.f;
  const int64_t mIdx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = mIdx < a.M;
  const float third = 1.0f / 3.0f;
  const float luma[3] = {(a.flags & EDN_CRF_LUMA_REC709) ? 0.2126f : (a.flags & EDN_CRF_LUMA_AVG) ? third : 0.299f,
                         (a.flags & EDN_CRF_LUMA_REC709) ? 0.7152f : (a.flags & EDN_CRF_LUMA_AVG) ? third : 0.587f,
                         (a.flags & EDN_CRF_LUMA_REC709) ? 0.0722f : (a.flags & EDN_CRF_LUMA_AVG) ? third : 0.114f};
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    float dxg = 0.f, x0 = 0.f;
    if (valid) {
      x0 = a.x[mIdx * 3 + c];
      const float dy = (a.flags & EDN_CRF_LUMA) ? a.d_out[mIdx] * luma[c] : a.d_out[mIdx * 3 + c];
      const float xg = (a.flags & EDN_CRF_GAMMA) ? powf(x0, 1.0f / a.p.gamma) : x0;
      dxg = dy;
      if (learn) {
        float in[8], h0[16], h1[16], h2[16];
        in[0] = xg;
        for (int f = 0; f < F; ++f)
          in[1 + f] = a.feat ? (a.feat_per_channel ? a.feat[(mIdx * 3 + c) * F + f] : a.feat[mIdx * F + f]) : 0.f;
        for (int j = 0; j < 16; ++j) {
          float s = b0[j];
          for (int k = 0; k < in_ch; ++k) s = fmaf(w0[j * in_ch + k], in[k], s);
          h0[j] = fmaxf(s, 0.f);
        }
        for (int j = 0; j < 16; ++j) {
          float s = b1[j];
#pragma unroll
          for (int k = 0; k < 16; ++k) s = fmaf(w1[j * 16 + k], h0[k], s);
          h1[j] = fmaxf(s, 0.f);
        }
        float o = b3[0];
        for (int j = 0; j < 16; ++j) {
          float s = b2[j];
#pragma unroll
          for (int k = 0; k < 16; ++k) s = fmaf(w2[j * 16 + k], h1[k], s);
          h2[j] = fmaxf(s, 0.f);
          o = fmaf(w3[j], h2[j], o);
        }
        const float y = sigmoidf_(o * 0.1f + xg);
        const float dz = dy * y * (1.f - y);
        const float d_o = 0.1f * dz;
        dxg = dz;
        float d2[16], d1[16], d0[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) d2[j] = h2[j] > 0.f ? d_o * w3[j] : 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float t = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) t = fmaf(d2[j], w2[j * 16 + k], t);
          d1[k] = h1[k] > 0.f ? t : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float t = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) t = fmaf(d1[j], w1[j * 16 + k], t);
          d0[k] = h0[k] > 0.f ? t : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) dxg = fmaf(d0[j], w0[j * in_ch], dxg);
        if (want_pg) {
          for (int k = 0; k < 8; ++k) st.in[k][lane] = k < in_ch ? in[k] : 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            st.h0[j][lane] = h0[j]; st.h1[j][lane] = h1[j]; st.h2[j][lane] = h2[j];
            st.d0[j][lane] = d0[j]; st.d1[j][lane] = d1[j]; st.d2[j][lane] = d2[j];
          }
          st.dout[lane] = d_o;
        }
      }
      float dx = dxg;
      if (a.flags & EDN_CRF_GAMMA) { const float ig = 1.0f / a.p.gamma; dx = dxg * ig * powf(x0, ig - 1.0f); }
      if (a.d_x) a.d_x[mIdx * 3 + c] = dx;
    } else if (want_pg) {
      for (int k = 0; k < 8; ++k) st.in[k][lane] = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { st.h0[j][lane] = 0.f; st.h1[j][lane] = 0.f; st.h2[j][lane] = 0.f; st.d0[j][lane] = 0.f; st.d1[j][lane] = 0.f; st.d2[j][lane] = 0.f; }
      st.dout[lane] = 0.f;
    }
    if (want_pg) {
      __syncwarp();
#pragma unroll 1
      for (int i = 0; i < kPerLane; ++i) {
        const int pi = lane + 32 * i;
        if (pi >= n_par) break;
        const float* A;      // delta row
        const float* B;      // activation row (nullptr: bias -> 1)
        int q = pi;
        if (q < n_w0) { A = st.d0[q / in_ch]; B = st.in[q % in_ch]; }
        else if ((q -= n_w0) < 16) { A = st.d0[q]; B = nullptr; }
        else if ((q -= 16) < 256) { A = st.d1[q >> 4]; B = st.h0[q & 15]; }
        else if ((q -= 256) < 16) { A = st.d1[q]; B = nullptr; }
        else if ((q -= 16) < 256) { A = st.d2[q >> 4]; B = st.h1[q & 15]; }
        else if ((q -= 256) < 16) { A = st.d2[q]; B = nullptr; }
        else if ((q -= 16) < 16) { A = st.dout; B = st.h2[q]; }
        else { A = st.dout; B = nullptr; }
        float t = 0.f;
#pragma unroll 8
        for (int s2 = 0; s2 < 32; ++s2) t = fmaf(A[s2], B ? B[s2] : 1.0f, t);
        pg[i] += t;
      }
      __syncwarp();
    }
  }
  if (want_pg) {
#pragma unroll 1
    for (int i = 0; i < kPerLane; ++i) {
      int q = lane + 32 * i;
      if (q >= n_par) break;
      float* dst;
      if (q < n_w0) dst = a.g.w0 + q;
      else if ((q -= n_w0) < 16) dst = a.g.b0 + q;
      else if ((q -= 16) < 256) dst = a.g.w1 + q;
      else if ((q -= 256) < 16) dst = a.g.b1 + q;
      else if ((q -= 16) < 256) dst = a.g.w2 + q;
      else if ((q -= 256) < 16) dst = a.g.b2 + q;
      else if ((q -= 16) < 16) dst = a.g.w3 + q;
      else dst = a.g.b3;
      atomicAdd(dst, pg[i]);
    }
  }
}

// ---- egm loss: loss = sum_i w_i (pred_i - bii_i)^2 / sum_i w_i ---------------------------------------------------------------
__global__ void egm_loss_bwd_kernel(const float* __restrict__ ls, const float* __restrict__ le, const float* __restrict__ bii,
                                    const uint8_t* __restrict__ mask, const float* __restrict__ cw, int C, int64_t M, float eps,
                                    const float* __restrict__ d_loss, float* __restrict__ d_ls, float* __restrict__ d_le) {
  __shared__ double sh[32];
  __shared__ float inv_den;
  double den = 0.0;
  for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
    float wgt = 1.0f;
    if (mask && cw) wgt = cw[mask[i * 3 + 0] ? 0 : (mask[i * 3 + 1] ? 1 : 2)];
    den += (double)wgt;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = den;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];
    inv_den = (float)(1.0 / t);
  }
  __syncthreads();
  const float g = d_loss[0] * inv_den;
  for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
    int ch = 0;
    float wgt = 1.0f;
    if (mask) {
      ch = mask[i * 3 + 0] ? 0 : (mask[i * 3 + 1] ? 1 : 2);
      if (cw) wgt = cw[ch];
    }
    const float a = le[i * C + ch] + eps, b = ls[i * C + ch] + eps;
    const float dp = 2.0f * wgt * ((logf(a) - logf(b)) - bii[i]) * g;
    for (int c = 0; c < C; ++c) {
      d_le[i * C + c] = (c == ch) ? dp / a : 0.f;
      d_ls[i * C + c] = (c == ch) ? -dp / b : 0.f;
    }
  }
}

__global__ void mse_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, const float* __restrict__ d_loss,
                               float* __restrict__ d_x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d_x[i] = 2.0f * (x[i] - y[i]) * (d_loss[0] / (float)n);
}

// ---- TV: reg(x) = 2 (sum dh^2 / count_h + sum dw^2 / count_w), x = [C][H][W];  grad (+)= scale * d reg / d x ----------------------
__global__ void tv_bwd_kernel(const float* __restrict__ x, int C, int H, int W, float scale, const float* __restrict__ d_loss,
                              float* __restrict__ grad) {
  const int64_t n = (int64_t)C * H * W;
  const float count_h = (float)((int64_t)C * (H - 1) * W);
  const int64_t cwi = (int64_t)C * H * (W - 1);
  const float count_w = (float)(cwi > 1 ? cwi : 1);
  const float kh = H > 1 ? 4.0f * scale * d_loss[0] / count_h : 0.f, kw = 4.0f * scale * d_loss[0] / count_w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const float v = x[i];
    float g = 0.f;
    if (h > 0) g += kh * (v - x[i - W]);
    if (h + 1 < H) g -= kh * (x[i + W] - v);
    if (w > 0) g += kw * (v - x[i - 1]);
    if (w + 1 < W) g -= kw * (x[i + 1] - v);
    grad[i] += g;
  }
}

}  // namespace
}  // namespace edn

extern "C" int edn_crf_bwd(const edn_crf_params* p, const float* x, const float* feat, int32_t feat_per_channel, int32_t flags,
                           int64_t m, const float* d_out, float* d_x, const edn_crf_grads* grads, void* stream) {
  using namespace edn;
  EDN_REQUIRE(p && x && d_out && m >= 0, "edn_crf_bwd: bad argument");
  edn_crf_grads g{};
  const bool learn = (flags & EDN_CRF_LEARN) && !(flags & EDN_CRF_SKIP_LEARN);
  if (learn) {
    EDN_REQUIRE(p->extra_features >= 0 && p->extra_features <= 7, "edn_crf_bwd: extra_features must be in [0,7]");
    EDN_REQUIRE(p->w0 && p->b0 && p->w1 && p->b1 && p->w2 && p->b2 && p->w3 && p->b3, "edn_crf_bwd: null CRF weight");
    if (grads) {
      g = *grads;
      EDN_REQUIRE(g.w0 && g.b0 && g.w1 && g.b1 && g.w2 && g.b2 && g.w3 && g.b3, "edn_crf_bwd: null CRF gradient buffer");
    }
  }
  if (m == 0) return EDN_OK;
  CrfBwdArgs a{*p, g, x, feat, feat_per_channel, flags, m, d_out, d_x};
  crf_bwd_kernel<<<(unsigned)((m + kCrfWarps * 32 - 1) / (kCrfWarps * 32)), kCrfWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_egm_loss_bwd(const float* luma_start, const float* luma_end, const float* bii, const uint8_t* color_mask,
                                const float* color_weight, int32_t channels, int64_t m, float log_eps, const float* d_loss,
                                float* d_luma_start, float* d_luma_end, void* stream) {
  using namespace edn;
  EDN_REQUIRE(luma_start && luma_end && bii && d_loss && d_luma_start && d_luma_end && m > 0, "edn_egm_loss_bwd: bad argument");
  EDN_REQUIRE(channels == 1 || (channels == 3 && color_mask), "edn_egm_loss_bwd: channels must be 1, or 3 with a colour mask");
  egm_loss_bwd_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(luma_start, luma_end, bii, color_mask, color_weight, channels, m,
                                                                           log_eps, d_loss, d_luma_start, d_luma_end);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_img2mse_bwd(const float* x, const float* y, int64_t n, const float* d_loss, float* d_x, void* stream) {
  using namespace edn;
  EDN_REQUIRE(x && y && d_loss && d_x && n > 0, "edn_img2mse_bwd: bad argument");
  mse_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n, d_loss, d_x);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_tv_loss_app_bwd(const float* const planes_chw[3], const float* const lines_chw[3], const int32_t plane_h[3],
                                   const int32_t plane_w[3], const int32_t line_len[3], const int32_t n_comp[3], const float* d_loss,
                                   float* const grad_planes_chw[3], float* const grad_lines_chw[3], void* stream) {
  using namespace edn;
  EDN_REQUIRE(planes_chw && lines_chw && plane_h && plane_w && line_len && n_comp && d_loss && grad_planes_chw && grad_lines_chw,
              "edn_tv_loss_app_bwd: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < 6; ++i) {
    const bool plane = i < 3;
    const int k = plane ? i : i - 3;
    const int Cc = n_comp[k], Hh = plane ? plane_h[k] : line_len[k], Ww = plane ? plane_w[k] : 1;
    const float* x = plane ? planes_chw[k] : lines_chw[k];
    float* gr = plane ? grad_planes_chw[k] : grad_lines_chw[k];
    EDN_REQUIRE(x && gr, "edn_tv_loss_app_bwd: null tensor %d", i);
    const int64_t n = (int64_t)Cc * Hh * Ww;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = 16 * (int64_t)num_sms();
    if (blocks > cap) blocks = cap;
    tv_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, Cc, Hh, Ww, plane ? 1e-2f : 1e-3f, d_loss, gr);
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

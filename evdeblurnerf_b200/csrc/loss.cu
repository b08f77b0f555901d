// Loss-path kernels of the event-blur objective: exposure blending, camera response (CRF), event generation-model
// loss, photometric MSE and the TV regulariser.  All fp32, all tiny or HBM-streaming; one launch each.
//   edn_weighted_sum  <- RigidBlurringModel.rbk_weighted_sum      networks/dpnerf/blurmodel.py:112-127, renderer.py:326-330
//   edn_crf_fwd       <- CRF.forward + encode_rgb / encode_luma   networks/tonemapping.py:59-93, 111-139
//   edn_egm_loss_fwd  <- egm_loss                                 utils/events.py:260-284
//   edn_img2mse       <- img2mse                                  utils/metrics.py:7
//   edn_tv_loss_app   <- VoxelNeRFBase.TV_loss_app / TVLoss       networks/pdrf/voxnerf.py:126-130, 306-324
#include "common.cuh"

namespace edn {
namespace {

// ---- exposure blending: out[n][c] = sum_e w[n][e] * x[n*E + e][c] (sequential over e, like torch's sum over dim 1) ----
__global__ void weighted_sum_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out,
                                    int64_t N, int E, int64_t C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int64_t n = i / C, c = i - n * C;
  float acc = 0.f;
  for (int e = 0; e < E; ++e) acc += x[(n * E + e) * C + c] * w[n * E + e];
  out[i] = acc;
}

// ---- CRF ------------------------------------------------------------------------------------------------------------
struct CrfArgs {
  edn_crf_params p;
  const float* x;      // [M][3]
  const float* feat;   // [M][F] (feat_per_channel = 0) or [M][3][F] (= 1) or NULL (zero padded)
  int feat_per_channel;
  int flags;
  int64_t M;
  float* out;          // [M][3] or [M][1]
};

__global__ void crf_kernel(const CrfArgs a) {
  __shared__ float w0[16 * 8], b0[16], w1[256], b1[16], w2[256], b2[16], w3[16], b3[1];
  const int F = a.p.extra_features, in_ch = 1 + F;
  const bool learn = (a.flags & EDN_CRF_LEARN) && !(a.flags & EDN_CRF_SKIP_LEARN);
  if (learn) {
    for (int i = threadIdx.x; i < 16 * in_ch; i += blockDim.x) w0[i] = a.p.w0[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { w1[i] = a.p.w1[i]; w2[i] = a.p.w2[i]; }
    for (int i = threadIdx.x; i < 16; i += blockDim.x) { b0[i] = a.p.b0[i]; b1[i] = a.p.b1[i]; b2[i] = a.p.b2[i]; w3[i] = a.p.w3[i]; }
    if (threadIdx.x == 0) b3[0] = a.p.b3[0];
    __syncthreads();
  }
  const int64_t mIdx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (mIdx >= a.M) return;
  float y[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float x = a.x[mIdx * 3 + c];
    if (a.flags & EDN_CRF_GAMMA) x = powf(x, 1.0f / a.p.gamma);
    if (learn) {
      float in[8];
      in[0] = x;
      for (int f = 0; f < F; ++f)
        in[1 + f] = a.feat ? (a.feat_per_channel ? a.feat[(mIdx * 3 + c) * F + f] : a.feat[mIdx * F + f]) : 0.f;
      float h0[16], h1[16];
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
        o = fmaf(w3[j], fmaxf(s, 0.f), o);
      }
      x = sigmoidf_(o * 0.1f + x);        // tonemapping.py:88-89
    }
    y[c] = x;
  }
  if (a.flags & EDN_CRF_LUMA) {
    if (a.flags & EDN_CRF_LUMA_REC709) a.out[mIdx] = 0.2126f * y[0] + 0.7152f * y[1] + 0.0722f * y[2];   // tonemapping.py:130-131
    else if (a.flags & EDN_CRF_LUMA_AVG) a.out[mIdx] = (y[0] + y[1] + y[2]) / 3.0f;                      // x.mean(-1), tonemapping.py:132-133
    else a.out[mIdx] = 0.299f * y[0] + 0.587f * y[1] + 0.114f * y[2];                                   // rec601, tonemapping.py:128-129
  } else {
    a.out[mIdx * 3 + 0] = y[0]; a.out[mIdx * 3 + 1] = y[1]; a.out[mIdx * 3 + 2] = y[2];
  }
}

// ---- deterministic single-block reductions (M is a few thousand) --------------------------------------------------------
constexpr int kRedThreads = 1024;
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;   // valid in thread 0
}

__global__ void egm_loss_kernel(const float* __restrict__ ls, const float* __restrict__ le, const float* __restrict__ bii,
                                const uint8_t* __restrict__ mask, const float* __restrict__ cw, int C, int64_t M, float eps,
                                float* __restrict__ out) {
  __shared__ double sh[32];
  double num = 0.0, den = 0.0;
  for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
    int ch = 0;
    float wgt = 1.0f;
    if (mask) {
      ch = mask[i * 3 + 0] ? 0 : (mask[i * 3 + 1] ? 1 : 2);   // one-hot colour mask (events.py:261-267)
      if (cw) wgt = cw[ch];
    }
    const float pred = logf(le[i * C + ch] + eps) - logf(ls[i * C + ch] + eps);
    const float dlt = pred - bii[i];
    num += (double)(dlt * dlt * wgt);
    den += (double)wgt;
  }
  const double n = block_sum(num, sh), d = block_sum(den, sh);
  if (threadIdx.x == 0) out[0] = (float)(n / d);
}

__global__ void mse_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, float* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const float d = x[i] - y[i]; s += (double)(d * d); }
  const double t = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = (float)(t / (double)n);
}

// ---- TV regulariser: sums of squared forward differences along H and W of [C][H][W] tensors -----------------------------------
struct TvDims { int C[6], H[6], W[6]; };
// all six tensors of a field in ONE launch: blockIdx.y = tensor (3 planes, 3 lines), grid-stride over its elements in blockIdx.x
struct TvPtrs { const float* x[6]; };
__global__ void tv_sums_all_kernel(const TvPtrs p, const TvDims d, double* __restrict__ acc /*[6][2]*/) {
  __shared__ double sh[32];
  const int t = blockIdx.y;
  const float* __restrict__ x = p.x[t];
  const int H = d.H[t], W = d.W[t];
  const int64_t n = (int64_t)d.C[t] * H * W;
  double hs = 0.0, ws = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const float v = x[i];
    if (h + 1 < H) { const float dd = x[i + W] - v; hs += (double)(dd * dd); }
    if (w + 1 < W) { const float dd = x[i + 1] - v; ws += (double)(dd * dd); }
  }
  const double th = block_sum(hs, sh), tw = block_sum(ws, sh);
  if (threadIdx.x == 0 && (th != 0.0 || tw != 0.0)) { atomicAdd(acc + 2 * t, th); atomicAdd(acc + 2 * t + 1, tw); }
}
__global__ void tv_finalize_kernel(const double* __restrict__ acc /*[6][2]*/, const TvDims d, float* __restrict__ out) {
  // TV_loss_app = sum_i 1e-2 * reg(plane_i) + 1e-3 * reg(line_i); reg = 2 * (h_tv / count_h + w_tv / count_w)
  float total = 0.f;
  for (int i = 0; i < 6; ++i) {
    const float count_h = (float)((int64_t)d.C[i] * (d.H[i] - 1) * d.W[i]);
    const int64_t cw = (int64_t)d.C[i] * d.H[i] * (d.W[i] - 1);
    const float count_w = (float)(cw > 1 ? cw : 1);
    const float reg = 2.0f * ((float)acc[2 * i] / count_h + (float)acc[2 * i + 1] / count_w);
    total = total + reg * (i < 3 ? 1e-2f : 1e-3f);
  }
  out[0] = total;
}

}  // namespace
}  // namespace edn

extern "C" int edn_weighted_sum(const float* x, const float* w, float* out, int64_t n, int32_t n_exposure, int64_t channels,
                                void* stream) {
  using namespace edn;
  EDN_REQUIRE(x && w && out && n >= 0 && n_exposure > 0 && channels > 0, "edn_weighted_sum: bad argument");
  if (n == 0) return EDN_OK;
  const int64_t tot = n * channels;
  weighted_sum_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, w, out, n, n_exposure, channels);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_crf_fwd(const edn_crf_params* p, const float* x, const float* feat, int32_t feat_per_channel, int32_t flags,
                           int64_t m, float* out, void* stream) {
  using namespace edn;
  EDN_REQUIRE(p && x && out && m >= 0, "edn_crf_fwd: bad argument");
  if (flags & EDN_CRF_LEARN) {
    EDN_REQUIRE(p->extra_features >= 0 && p->extra_features <= 7, "edn_crf_fwd: extra_features must be in [0,7]");
    EDN_REQUIRE((flags & EDN_CRF_SKIP_LEARN) || (p->w0 && p->b0 && p->w1 && p->b1 && p->w2 && p->b2 && p->w3 && p->b3),
                "edn_crf_fwd: null CRF weight");
  }
  if (m == 0) return EDN_OK;
  CrfArgs a{*p, x, feat, feat_per_channel, flags, m, out};
  crf_kernel<<<(unsigned)((m + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_egm_loss_fwd(const float* luma_start, const float* luma_end, const float* bii, const uint8_t* color_mask,
                                const float* color_weight, int32_t channels, int64_t m, float log_eps, float* out, void* stream) {
  using namespace edn;
  EDN_REQUIRE(luma_start && luma_end && bii && out && m > 0, "edn_egm_loss_fwd: bad argument");
  EDN_REQUIRE(channels == 1 || (channels == 3 && color_mask), "edn_egm_loss_fwd: channels must be 1, or 3 with a colour mask");
  egm_loss_kernel<<<1, kRedThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(luma_start, luma_end, bii, color_mask, color_weight,
                                                                              channels, m, log_eps, out);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_img2mse(const float* x, const float* y, int64_t n, float* out, void* stream) {
  using namespace edn;
  EDN_REQUIRE(x && y && out && n > 0, "edn_img2mse: bad argument");
  mse_kernel<<<1, kRedThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n, out);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_tv_loss_app(const float* const planes_chw[3], const float* const lines_chw[3], const int32_t plane_h[3],
                               const int32_t plane_w[3], const int32_t line_len[3], const int32_t n_comp[3], double* workspace,
                               float* out, void* stream) {
  using namespace edn;
  EDN_REQUIRE(planes_chw && lines_chw && plane_h && plane_w && line_len && n_comp && workspace && out, "edn_tv_loss_app: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  EDN_CUDA_OK(cudaMemsetAsync(workspace, 0, 12 * sizeof(double), st));
  TvDims d;
  TvPtrs ptrs;
  int64_t n_max = 0;
  for (int i = 0; i < 6; ++i) {
    const bool plane = i < 3;
    const int k = plane ? i : i - 3;
    d.C[i] = n_comp[k];
    d.H[i] = plane ? plane_h[k] : line_len[k];
    d.W[i] = plane ? plane_w[k] : 1;
    ptrs.x[i] = plane ? planes_chw[k] : lines_chw[k];
    EDN_REQUIRE(ptrs.x[i] != nullptr, "edn_tv_loss_app: null tensor %d", i);
    const int64_t n = (int64_t)d.C[i] * d.H[i] * d.W[i];
    n_max = n > n_max ? n : n_max;
  }
  {
    int64_t blocks = (n_max + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = 2 * (int64_t)num_sms();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    tv_sums_all_kernel<<<dim3((unsigned)blocks, 6), 256, 0, st>>>(ptrs, d, workspace);
  }
  tv_finalize_kernel<<<1, 1, 0, st>>>(workspace, d, out);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

// Event double integral (EDI) prior: bilinear event splat into brightness-increment images + the double integral that
// turns a blurry frame into a sharp mid-exposure estimate.
// Replaces utils/edi.py:7-95 (interpolate_subpixel, brightness_increment_image, inner_double_integral,
// deblur_double_integral) as driven by LLFFEventsDataset.compute_edi_prior (data/loader_events.py:99-131).
// HBM/atomic-bound: one thread per event (4 taps, red.global.add.f32), then one thread per pixel.
#include "common.cuh"

namespace edn {
namespace {

// One event -> up to 4 taps of +/- weight into the segment's image.  Tap rule of interpolate_subpixel (edi.py:7-41): the
// floor tap always, the ceil tap only when it differs from the coordinate (no duplicate for integer coordinates); a tap
// is kept when ref < w, h; negative refs are NOT masked and wrap like numpy's negative indexing (Appendix B quirk).
__global__ void edi_splat_kernel(const float* __restrict__ ex, const float* __restrict__ ey, const float* __restrict__ ep,
                                 const int64_t* __restrict__ seg_start, const int64_t* __restrict__ seg_end, int n_seg, int H, int W,
                                 float c_pos, float c_neg, float* __restrict__ bii /*[n_seg][H][W]*/) {
  const int seg = blockIdx.y;
  const int64_t s0 = seg_start[seg], s1 = seg_end[seg];
  float* img = bii + (size_t)seg * H * W;
  for (int64_t i = s0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s1; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = ex[i], y = ey[i];
    const float sgn = ep[i] > 0.f ? c_pos : -c_neg;
    const float xf = floorf(x), yf = floorf(y), xc = ceilf(x), yc = ceilf(y);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const bool cx = t & 1, cy = t & 2;
      const float xr = cx ? xc : xf, yr = cy ? yc : yf;
      if ((cx && xr == x) || (cy && yr == y)) continue;
      if (!(xr < (float)W) || !(yr < (float)H)) continue;
      const float val = fmaxf(0.f, 1.f - fabsf(xr - x)) * fmaxf(0.f, 1.f - fabsf(yr - y));
      int xi = (int)xr, yi = (int)yr;
      if (xi < 0) xi += W;
      if (yi < 0) yi += H;
      if (xi < 0 || yi < 0) continue;   // numpy would raise; unreachable for in-sensor events
      atomicAdd(img + (size_t)yi * W + xi, val * sgn);
    }
  }
}

// sharp = (2N+1) * blurry / sum_k exp(I_k), I_k = signed partial sums of bii around the mid exposure (edi.py:73-95)
__global__ void edi_deblur_kernel(const float* __restrict__ bii, int n_seg, int64_t HW, int C, const float* __restrict__ blurry,
                                  float* __restrict__ sharp) {
  const int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= HW) return;
  const int N = n_seg / 2;
  float denom = 1.0f;            // exp(0) for the mid-exposure image
  float run = 0.f;
  for (int i = N - 1; i >= 0; --i) { run += bii[(size_t)i * HW + px]; denom += expf(-run); }   // -sum_{k=i}^{N-1}
  run = 0.f;
  for (int i = 0; i < N; ++i) { run += bii[(size_t)(N + i) * HW + px]; denom += expf(run); }    // +sum_{k=N}^{N+i}
  const float scale = (float)(2 * N + 1);
  for (int c = 0; c < C; ++c) sharp[px * C + c] = scale * blurry[px * C + c] / denom;
}

}  // namespace
}  // namespace edn

extern "C" int edn_edi_prior(const float* ev_x, const float* ev_y, const float* ev_p, const int64_t* seg_start,
                             const int64_t* seg_end, int32_t n_seg, int64_t max_seg_events, const float* blurry, int32_t H,
                             int32_t W, int32_t C, float c_pos, float c_neg, float* bii, float* sharp, void* stream) {
  using namespace edn;
  EDN_REQUIRE(ev_x && ev_y && ev_p && seg_start && seg_end && blurry && bii && sharp, "edn_edi_prior: null pointer");
  EDN_REQUIRE(n_seg >= 2 && n_seg % 2 == 0 && H > 0 && W > 0 && C > 0, "edn_edi_prior: n_seg must be even and >= 2");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t HW = (int64_t)H * W;
  EDN_CUDA_OK(cudaMemsetAsync(bii, 0, sizeof(float) * n_seg * HW, st));
  int64_t bx = (max_seg_events + 255) / 256;
  if (bx < 1) bx = 1;
  if (bx > 4 * (int64_t)num_sms()) bx = 4 * (int64_t)num_sms();
  edi_splat_kernel<<<dim3((unsigned)bx, (unsigned)n_seg), 256, 0, st>>>(ev_x, ev_y, ev_p, seg_start, seg_end, n_seg, H, W, c_pos, c_neg, bii);
  edi_deblur_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, st>>>(bii, n_seg, HW, C, blurry, sharp);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

// Ray / event batch generation on the device: what the reference does in its DataLoader workers and -- for the
// event poses -- in numpy / scipy on the host every step.
//   edn_rays_from_pixels   <- get_rays_pix                               utils/rays.py:25-36
//   edn_make_rgb_batch     <- LLFFDataset.__getitem__                    data/loader.py:325-356 (+ unravel_index, utils/misc.py:160)
//   edn_gather_successor   <- gather_successor                           utils/events.py:221-257
//   edn_interpolate_poses  <- LLFFEventsDataset.interpolate_poses        data/loader_events.py:133-148, 175-183; utils/data.py:34-61, 167-183
// All tiny, latency-bound kernels: one thread per ray / event.
#include "common.cuh"

namespace edn {
namespace {

// ox = float(half - cx), oy = float(half - cy): the reference forms these scalars in Python floats before they meet the fp32 tensors
__device__ __forceinline__ void pixel_ray(float x, float y, const float* __restrict__ c2w /*[3][4]*/, float fx, float fy, float ox, float oy,
                                          float* __restrict__ out /*[3][2]*/) {
  // dirs = [(x + (half - cx)) / fx, -(y + (half - cy)) / fy, -1];  rays_d = sum(dirs * c2w[:3,:3], -1);  rays_o = c2w[:3, 3]
  const float d0 = __fdiv_rn(__fadd_rn(x, ox), fx), d1 = -__fdiv_rn(__fadd_rn(y, oy), fy), d2 = -1.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float* r = c2w + 4 * i;
    out[2 * i] = r[3];
    out[2 * i + 1] = __fadd_rn(__fadd_rn(__fmul_rn(d0, r[0]), __fmul_rn(d1, r[1])), __fmul_rn(d2, r[2]));   // torch.sum order, no FMA
  }
}

__global__ void rays_from_pixels_kernel(const float* __restrict__ coords, const float* __restrict__ poses, int broadcast, int64_t n, float fx,
                                        float fy, float ox, float oy, float* __restrict__ rays) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pixel_ray(coords[2 * i], coords[2 * i + 1], poses + (broadcast ? 0 : i * 12), fx, fy, ox, oy, rays + i * 6);
}

__global__ void make_rgb_batch_kernel(const int64_t* __restrict__ ray_ids, int64_t n, const float* __restrict__ images,
                                      const float* __restrict__ poses, int n_img, int H, int W, float fx, float fy, float ox, float oy,
                                      float* __restrict__ rays, float* __restrict__ rays_x, float* __restrict__ rays_y,
                                      int64_t* __restrict__ images_idx, float* __restrict__ rgbsf, float* __restrict__ poses_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = ray_ids[i], hw = (int64_t)H * W;
  const int64_t img = id / hw, rem = id - img * hw;
  const int y = (int)(rem / W), x = (int)(rem - (int64_t)y * W);
  const float* c2w = poses + img * 12;
  pixel_ray((float)x, (float)y, c2w, fx, fy, ox, oy, rays + i * 6);
  if (rays_x) rays_x[i] = (float)x + 0.5f;
  if (rays_y) rays_y[i] = (float)y + 0.5f;
  if (images_idx) images_idx[i] = img;
  if (rgbsf) {
    const float* px = images + (img * hw + rem) * 3;
    rgbsf[3 * i] = px[0]; rgbsf[3 * i + 1] = px[1]; rgbsf[3 * i + 2] = px[2];
  }
  if (poses_out) {
#pragma unroll
    for (int k = 0; k < 12; ++k) poses_out[i * 12 + k] = c2w[k];
  }
}

__global__ void gather_successor_kernel(const int64_t* __restrict__ query_idx, const int64_t* __restrict__ query_hops, int64_t n,
                                        const int64_t* __restrict__ succ, const int32_t* __restrict__ pol, int64_t n_ev,
                                        int64_t* __restrict__ out_idx, int32_t* __restrict__ out_neg, int32_t* __restrict__ out_pos) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t cur = query_idx[i];
  const int64_t hops = query_hops[i];
  int32_t pos = 0, neg = 0;
  bool invalid = false;
  for (int64_t h = 0; h <= hops; ++h) {
    const int64_t nxt = succ[cur];
    if (nxt < 0 || nxt >= n_ev) { invalid = true; break; }      // the reference's final override makes the early exit equivalent
    const int32_t p = pol[nxt];
    pos += p > 0 ? p : 0;
    neg += p < 0 ? p : 0;
    cur = nxt;
  }
  out_idx[i] = invalid ? -1 : cur;
  out_neg[i] = invalid ? 0 : neg;
  out_pos[i] = invalid ? 0 : pos;
}

__device__ __forceinline__ int find_interval(const double* __restrict__ x, int n, double t) {     // x[i] <= t <= x[i+1], i in [0, n-2]
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (x[mid] <= t) lo = mid; else hi = mid; }
  return lo;
}

struct PoseArgs {
  const double* t; int64_t n;
  const double* key_times; const double* key_quats; int n_keys;
  const double* brk; const double* coef; int n_brk;
  double bd_scale; const double* recenter_inv; float* poses;
};

__global__ void interpolate_poses_kernel(const PoseArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  double t = a.t[i];
  t = fmin(fmax(t, a.key_times[0]), a.key_times[a.n_keys - 1]);      // cannot interpolate beyond the known poses (loader_events.py:179)
  // ---- rotation: scipy Slerp = q_i * exp(alpha * log(q_i^-1 q_{i+1})) ---------------------------------------------------
  const int k = find_interval(a.key_times, a.n_keys, t);
  const double alpha = (t - a.key_times[k]) / (a.key_times[k + 1] - a.key_times[k]);
  const double* q0 = a.key_quats + 4 * k;
  const double* q1 = a.key_quats + 4 * (k + 1);
  // relative rotation q0^-1 * q1 (x, y, z, w)
  const double ax = -q0[0], ay = -q0[1], az = -q0[2], aw = q0[3];
  double rx = aw * q1[0] + ax * q1[3] + ay * q1[2] - az * q1[1];
  double ry = aw * q1[1] - ax * q1[2] + ay * q1[3] + az * q1[0];
  double rz = aw * q1[2] + ax * q1[1] - ay * q1[0] + az * q1[3];
  double rw = aw * q1[3] - ax * q1[0] - ay * q1[1] - az * q1[2];
  if (rw < 0.0) { rx = -rx; ry = -ry; rz = -rz; rw = -rw; }             // shortest arc (rotation vector of norm <= pi)
  const double vn = sqrt(rx * rx + ry * ry + rz * rz);
  const double theta = 2.0 * atan2(vn, rw);
  double sx = 0.0, sy = 0.0, sz = 0.0, sw = 1.0;                        // exp(alpha * log(q_rel))
  if (vn > 1e-300) {
    const double hs = sin(0.5 * alpha * theta) / vn;
    sx = rx * hs; sy = ry * hs; sz = rz * hs; sw = cos(0.5 * alpha * theta);
  }
  const double qx = q0[3] * sx + q0[0] * sw + q0[1] * sz - q0[2] * sy;
  const double qy = q0[3] * sy - q0[0] * sz + q0[1] * sw + q0[2] * sx;
  const double qz = q0[3] * sz + q0[0] * sy - q0[1] * sx + q0[2] * sw;
  const double qw = q0[3] * sw - q0[0] * sx - q0[1] * sy - q0[2] * sz;
  double R[3][3];
  R[0][0] = 1 - 2 * (qy * qy + qz * qz); R[0][1] = 2 * (qx * qy - qz * qw); R[0][2] = 2 * (qx * qz + qy * qw);
  R[1][0] = 2 * (qx * qy + qz * qw); R[1][1] = 1 - 2 * (qx * qx + qz * qz); R[1][2] = 2 * (qy * qz - qx * qw);
  R[2][0] = 2 * (qx * qz - qy * qw); R[2][1] = 2 * (qy * qz + qx * qw); R[2][2] = 1 - 2 * (qx * qx + qy * qy);
  // ---- translation: piecewise cubic (the PPoly form of scipy's interp1d(kind="cubic") spline) ---------------------------------
  const int j = find_interval(a.brk, a.n_brk, t);
  const double u = t - a.brk[j];
  double T[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double* c = a.coef + ((size_t)j * 4) * 3 + d;               // [interval][power (highest first)][dim]
    T[d] = ((c[0] * u + c[3]) * u + c[6]) * u + c[9];
  }
  // ---- matrix format [c1, -c0, c2, T] (loader_events.py:137), fp32, T *= bd_scale, recenter (inv(c2w) @ pose) -------------------
  float P[3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    P[r][0] = (float)R[r][1]; P[r][1] = (float)(-R[r][0]); P[r][2] = (float)R[r][2];
    P[r][3] = (float)T[r] * (float)a.bd_scale;
  }
  float* out = a.poses + i * 12;
  if (a.recenter_inv) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double s = a.recenter_inv[4 * r + 3] * (c == 3 ? 1.0 : 0.0);
#pragma unroll
        for (int m = 0; m < 3; ++m) s += a.recenter_inv[4 * r + m] * (double)P[m][c];
        out[4 * r + c] = (float)s;
      }
  } else {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) out[4 * r + c] = P[r][c];
  }
}

}  // namespace
}  // namespace edn

extern "C" int edn_rays_from_pixels(const float* coords, const float* poses, int32_t broadcast_pose, int64_t n, double fx, double fy,
                                    double cx, double cy, int32_t add_halfpix, float* rays, void* stream) {
  using namespace edn;
  if (n == 0) return EDN_OK;
  EDN_REQUIRE(coords && poses && rays && n > 0, "edn_rays_from_pixels: bad argument");
  rays_from_pixels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      coords, poses, broadcast_pose, n, (float)fx, (float)fy, (float)((add_halfpix ? 0.5 : 0.0) - cx), (float)((add_halfpix ? 0.5 : 0.0) - cy), rays);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_make_rgb_batch(const int64_t* ray_ids, int64_t n, const float* images, const float* poses, int32_t n_img, int32_t H,
                                  int32_t W, double fx, double fy, double cx, double cy, float* rays, float* rays_x, float* rays_y,
                                  int64_t* images_idx, float* rgbsf, float* poses_out, void* stream) {
  using namespace edn;
  if (n == 0) return EDN_OK;
  EDN_REQUIRE(ray_ids && poses && rays && n > 0 && n_img > 0 && H > 0 && W > 0, "edn_make_rgb_batch: bad argument");
  EDN_REQUIRE(images || !rgbsf, "edn_make_rgb_batch: rgbsf needs the image stack");
  make_rgb_batch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ray_ids, n, images, poses, n_img, H, W, (float)fx, (float)fy, (float)(0.5 - cx), (float)(0.5 - cy), rays, rays_x, rays_y, images_idx, rgbsf,
      poses_out);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_gather_successor(const int64_t* query_idx, const int64_t* query_hops, int64_t n, const int64_t* successor_map,
                                    const int32_t* polarity, int64_t n_events, int64_t* succ_idx, int32_t* neg_cumsum,
                                    int32_t* pos_cumsum, void* stream) {
  using namespace edn;
  if (n == 0) return EDN_OK;
  EDN_REQUIRE(query_idx && query_hops && successor_map && polarity && succ_idx && neg_cumsum && pos_cumsum && n > 0 && n_events > 0,
              "edn_gather_successor: bad argument");
  gather_successor_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      query_idx, query_hops, n, successor_map, polarity, n_events, succ_idx, neg_cumsum, pos_cumsum);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_interpolate_poses(const double* t, int64_t n, const double* key_times, const double* key_quats, int32_t n_keys,
                                     const double* trans_breaks, const double* trans_coef, int32_t n_breaks, double bd_scale,
                                     const double* recenter_inv, float* poses, void* stream) {
  using namespace edn;
  if (n == 0) return EDN_OK;
  EDN_REQUIRE(t && key_times && key_quats && trans_breaks && trans_coef && poses && n > 0 && n_keys >= 2 && n_breaks >= 2,
              "edn_interpolate_poses: bad argument");
  PoseArgs a{t, n, key_times, key_quats, n_keys, trans_breaks, trans_coef, n_breaks, bd_scale, recenter_inv, poses};
  interpolate_poses_kernel<<<(unsigned)((n + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

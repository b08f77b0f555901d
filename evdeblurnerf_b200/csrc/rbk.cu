// DP-NeRF rigid blur kernel + render() prologue, fused: per primary ray, view latent -> r / v / w heads -> SE(3)
// exponential map per motion -> warped sub-rays -> view directions + NDC -> the [N*E][11] ray batch of render_rays.
// Replaces RigidBlurringModel.forward + rbk_warp (networks/dpnerf/blurmodel.py:129-173, 51-82), SE3Field /
// RigidBody.exp_se3 (utils/rigid_warping.py:18-49, 72-154), NeRFAll.render prologue (networks/renderer.py:423-446)
// and get_ndc_rays (utils/rays.py:104-145).  One thread per primary ray; the ~4K head weights sit in shared memory.
#include "common.cuh"

namespace edn {
namespace {

constexpr int kW = 32;          // latent / branch width
constexpr int kMaxE = 16;

struct RbkArgs {
  edn_rbk_params p;
  const float* rays;        // [N][3][2]
  const int64_t* images_idx;  // [N]
  int64_t N;
  int H, W;
  float focal, near, far;
  int ndc;
  float* new_rays;          // [N][E][3][2] or NULL
  float* weight;            // [N][E]
  float* img_embed;         // [N][32] or NULL
  float* ray_batch;         // [N*E][11] or NULL
};

__device__ __forceinline__ void ndc_ray(float H, float W, float focal, float near, float o[3], float d[3]) {
  // utils/rays.py:104-145 with near = 1 fixed by the caller (renderer.py:437)
  const float t = -(near + o[2]) / d[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = o[i] + t * d[i];
  const float ox_oz = o[0] / o[2], oy_oz = o[1] / o[2];
  const float sx = -1.f / (W / (2.f * focal)), sy = -1.f / (H / (2.f * focal));
  const float o0 = sx * ox_oz, o1 = sy * oy_oz, o2 = 1.f + 2.f * near / o[2];
  const float d0 = sx * (d[0] / d[2] - ox_oz), d1 = sy * (d[1] / d[2] - oy_oz), d2 = 1.f - o2;
  o[0] = o0; o[1] = o1; o[2] = o2; d[0] = d0; d[1] = d1; d[2] = d2;
}

__global__ void rbk_warp_ndc_kernel(const RbkArgs a) {
  extern __shared__ float sm[];
  const int M = a.p.num_motion, E = M + 1;
  // smem: 3 branches (W [32][32] + b [32]) | r_linear W [3M][32] + b | v_linear | w_linear W [E][32] + b
  float* br_w = sm;                       // [3][32*32]
  float* br_b = br_w + 3 * kW * kW;       // [3][32]
  float* r_w = br_b + 3 * kW;             // [3M][32]
  float* r_b = r_w + 3 * M * kW;
  float* v_w = r_b + 3 * M;
  float* v_b = v_w + 3 * M * kW;
  float* w_w = v_b + 3 * M;
  float* w_b = w_w + E * kW;
  const float* bw[3] = {a.p.r_branch_w, a.p.v_branch_w, a.p.w_branch_w};
  const float* bb[3] = {a.p.r_branch_b, a.p.v_branch_b, a.p.w_branch_b};
  for (int i = threadIdx.x; i < kW * kW; i += blockDim.x)
    for (int k = 0; k < 3; ++k) br_w[k * kW * kW + i] = bw[k][i];
  for (int i = threadIdx.x; i < kW; i += blockDim.x)
    for (int k = 0; k < 3; ++k) br_b[k * kW + i] = bb[k][i];
  for (int i = threadIdx.x; i < 3 * M * kW; i += blockDim.x) { r_w[i] = a.p.r_linear_w[i]; v_w[i] = a.p.v_linear_w[i]; }
  for (int i = threadIdx.x; i < 3 * M; i += blockDim.x) { r_b[i] = a.p.r_linear_b[i]; v_b[i] = a.p.v_linear_b[i]; }
  for (int i = threadIdx.x; i < E * kW; i += blockDim.x) w_w[i] = a.p.w_linear_w[i];
  for (int i = threadIdx.x; i < E; i += blockDim.x) w_b[i] = a.p.w_linear_b[i];
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;

  float emb[kW];
  const int64_t idx = a.images_idx[n];
#pragma unroll
  for (int k = 0; k < kW; ++k) emb[k] = a.p.img_embed[idx * kW + k];    // embedding.py:31-32
  if (a.img_embed) {
#pragma unroll
    for (int k = 0; k < kW; ++k) a.img_embed[n * kW + k] = emb[k];
  }
  float o[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i] = a.rays[n * 6 + 2 * i]; d[i] = a.rays[n * 6 + 2 * i + 1]; }

  float rv[2][3 * (kMaxE - 1)];   // r and v heads, flat [3*M] = component-major, motion-minor (blurmodel.py:52-53)
  float wgt[kMaxE];
  for (int head = 0; head < 3; ++head) {
    float h[kW];
    for (int j = 0; j < kW; ++j) {
      float s = br_b[head * kW + j];
#pragma unroll
      for (int k = 0; k < kW; ++k) s = fmaf(br_w[head * kW * kW + j * kW + k], emb[k], s);
      h[j] = fmaxf(s, 0.f);
    }
    if (head < 2) {
      const float* lw = head == 0 ? r_w : v_w;
      const float* lb = head == 0 ? r_b : v_b;
      for (int j = 0; j < 3 * M; ++j) {
        float s = lb[j];
#pragma unroll
        for (int k = 0; k < kW; ++k) s = fmaf(lw[j * kW + k], h[k], s);
        rv[head][j] = s * a.p.rv_window;
      }
    } else {
      float tot = 0.f;
      for (int j = 0; j < E; ++j) {
        float s = w_b[j];
#pragma unroll
        for (int k = 0; k < kW; ++k) s = fmaf(w_w[j * kW + k], h[k], s);
        wgt[j] = sigmoidf_(s);
        tot += wgt[j];
      }
      for (int j = 0; j < E; ++j) a.weight[n * E + j] = wgt[j] / (tot + 1e-10f);
    }
  }

  for (int e = 0; e < E; ++e) {
    float wo[3], wd[3];
    if (e == 0) {                      // use_origin: slot 0 is the un-warped ray
#pragma unroll
      for (int i = 0; i < 3; ++i) { wo[i] = o[i]; wd[i] = d[i]; }
    } else {
      const int mi = e - 1;
      const float rot[3] = {rv[0][0 * M + mi], rv[0][1 * M + mi], rv[0][2 * M + mi]};
      const float trn[3] = {rv[1][0 * M + mi], rv[1][1 * M + mi], rv[1][2 * M + mi]};
      const float theta = sqrtf(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]) + 1.0e-10f;   // rigid_warping.py:24
      const float w[3] = {rot[0] / theta, rot[1] / theta, rot[2] / theta};
      const float v[3] = {trn[0] / theta, trn[1] / theta, trn[2] / theta};
      const float Wm[3][3] = {{0.f, -w[2], w[1]}, {w[2], 0.f, -w[0]}, {-w[1], w[0], 0.f}};
      float WW[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) WW[i][j] = Wm[i][0] * Wm[0][j] + Wm[i][1] * Wm[1][j] + Wm[i][2] * Wm[2][j];
      const float sn = sinf(theta), cs = cosf(theta);
      float R[3][3], p[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float pi = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float eye = (i == j) ? 1.f : 0.f;
          R[i][j] = eye + sn * Wm[i][j] + (1.f - cs) * WW[i][j];                       // exp_so3, rigid_warping.py:97-113
          pi += (theta * eye + (1.f - cs) * Wm[i][j] + (theta - sn) * WW[i][j]) * v[j];  // exp_se3 translation, :75-95
        }
        p[i] = pi;
      }
      float we[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        wo[i] = R[i][0] * o[0] + R[i][1] * o[1] + R[i][2] * o[2] + p[i];
        we[i] = R[i][0] * (o[0] + d[0]) + R[i][1] * (o[1] + d[1]) + R[i][2] * (o[2] + d[2]) + p[i];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) wd[i] = we[i] - wo[i];
    }
    if (a.new_rays) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { a.new_rays[((n * E + e) * 3 + i) * 2] = wo[i]; a.new_rays[((n * E + e) * 3 + i) * 2 + 1] = wd[i]; }
    }
    if (a.ray_batch) {
      const float nrm = sqrtf(wd[0] * wd[0] + wd[1] * wd[1] + wd[2] * wd[2]);
      const float vd[3] = {wd[0] / nrm, wd[1] / nrm, wd[2] / nrm};   // from the pre-NDC direction (renderer.py:425-432)
      if (a.ndc) ndc_ray((float)a.H, (float)a.W, a.focal, 1.0f, wo, wd);
      float* rb = a.ray_batch + (n * E + e) * 11;
      rb[0] = wo[0]; rb[1] = wo[1]; rb[2] = wo[2]; rb[3] = wd[0]; rb[4] = wd[1]; rb[5] = wd[2];
      rb[6] = a.near; rb[7] = a.far; rb[8] = vd[0]; rb[9] = vd[1]; rb[10] = vd[2];
    }
  }
}

// render() prologue alone (no blur kernel): rays [R][3][2] -> ray_batch [R][11]
__global__ void ray_batch_kernel(const float* __restrict__ rays, int64_t R, int H, int W, float focal, float near, float far, int ndc,
                                 float* __restrict__ ray_batch) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= R) return;
  float o[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i] = rays[n * 6 + 2 * i]; d[i] = rays[n * 6 + 2 * i + 1]; }
  const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const float vd[3] = {d[0] / nrm, d[1] / nrm, d[2] / nrm};
  if (ndc) ndc_ray((float)H, (float)W, focal, 1.0f, o, d);
  float* rb = ray_batch + n * 11;
  rb[0] = o[0]; rb[1] = o[1]; rb[2] = o[2]; rb[3] = d[0]; rb[4] = d[1]; rb[5] = d[2];
  rb[6] = near; rb[7] = far; rb[8] = vd[0]; rb[9] = vd[1]; rb[10] = vd[2];
}

}  // namespace
}  // namespace edn

extern "C" int edn_rbk_warp_ndc_fwd(const edn_rbk_params* p, const float* rays, const int64_t* images_idx, int64_t n_rays,
                                    int32_t H, int32_t W, float focal, float near, float far, int32_t ndc, float* new_rays,
                                    float* weight, float* img_embed, float* ray_batch, void* stream) {
  using namespace edn;
  EDN_REQUIRE(p && rays && images_idx && weight, "edn_rbk_warp_ndc_fwd: null pointer");
  EDN_REQUIRE(p->num_motion >= 0 && p->num_motion < kMaxE, "edn_rbk_warp_ndc_fwd: num_motion must be in [0,%d)", kMaxE);
  EDN_REQUIRE(p->img_embed && p->r_branch_w && p->r_branch_b && p->v_branch_w && p->v_branch_b && p->w_branch_w && p->w_branch_b &&
              p->w_linear_w && p->w_linear_b, "edn_rbk_warp_ndc_fwd: null kernel-net weight");
  EDN_REQUIRE(p->num_motion == 0 || (p->r_linear_w && p->r_linear_b && p->v_linear_w && p->v_linear_b),
              "edn_rbk_warp_ndc_fwd: null r/v head weight");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  RbkArgs a{*p, rays, images_idx, n_rays, H, W, focal, near, far, ndc, new_rays, weight, img_embed, ray_batch};
  const int M = p->num_motion, E = M + 1;
  const size_t smem = sizeof(float) * (3 * kW * kW + 3 * kW + 2 * (3 * M * kW + 3 * M) + E * kW + E);
  rbk_warp_ndc_kernel<<<(unsigned)((n_rays + 127) / 128), 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

extern "C" int edn_build_ray_batch(const float* rays, int64_t n_rays, int32_t H, int32_t W, float focal, float near, float far,
                                   int32_t ndc, float* ray_batch, void* stream) {
  using namespace edn;
  EDN_REQUIRE(rays && ray_batch, "edn_build_ray_batch: null pointer");
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  ray_batch_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rays, n_rays, H, W, focal, near, far,
                                                                                                      ndc, ray_batch);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

// Hierarchical sample placement: inverse-CDF sampling + merge with the coarse samples.
// Replaces utils/rays.py:149-193 (sample_pdf) as called at networks/renderer.py:199-203, the torch.sort at
// renderer.py:205 and torch.std at renderer.py:250.
//
// Index contract (DESIGN.md "Numerics contract"): the pdf normaliser and the cdf prefix sums are accumulated in fp64 and
// rounded to fp32 once per entry -- what torch's CPU cumsum does -- so `inds` is platform independent and bit-exact
// against the oracle.  The fp64 sums of these fp32 terms are exact (terms within 2^29 of each other), so the warp-parallel
// scan below returns the same bits as the sequential sum.  The merge is a stable rank sort (ties: coarse sample first).
#include "common.cuh"

namespace edn {

constexpr int kPdfWarpsPerBlock = 4;

struct PdfArgs {
  const float* z0;
  const float* w0;
  const float* u_det;
  const float* u_rand;
  int64_t n_rays;
  int nc, ni;
  float* z_samples;
  int64_t* inds;
  float* z_vals;
  int64_t* order;
  float* z_std;
};

__global__ void __launch_bounds__(kPdfWarpsPerBlock * 32) sample_pdf_merge_kernel(const PdfArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nc = a.nc, ni = a.ni, nb = nc - 1, nt = nc + ni;
  // per-warp scratch: keys[nt] | cdf[nb] | bins[nb]
  float* keys = smem + warp * (nt + 2 * nb);
  float* cdf = keys + nt;
  float* bins = cdf + nb;
  const int64_t ray = (int64_t)blockIdx.x * kPdfWarpsPerBlock + warp;
  if (ray >= a.n_rays) return;
  const float* z0 = a.z0 + ray * nc;
  const float* w0 = a.w0 + ray * nc;

  for (int i = lane; i < nc; i += 32) keys[i] = z0[i];
  __syncwarp();
  for (int i = lane; i < nb; i += 32) bins[i] = 0.5f * __fadd_rn(keys[i + 1], keys[i]);   // z_vals_mid, renderer.py:199
  {
    // weights[..., 1:-1] + 1e-5 -> nc-2 terms; normaliser and prefix sums in fp64, rounded to fp32 once per entry.  The terms are
    // fp32 values within a factor 2^17 of each other (1e-5 <= w + 1e-5 <= ~1), so every fp64 partial sum of up to 1024 of them
    // is EXACT and the association order cannot matter: a warp scan gives the bits of the sequential sum (header comment).
    const int n = nc - 2, per = (n + 31) / 32, lo = lane * per, hi = min(lo + per, n);
    double part = 0.0;
    for (int i = lo; i < hi; ++i) part += (double)__fadd_rn(w0[i + 1], 1e-5f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const float norm = (float)part;
    double loc = 0.0;
    for (int i = lo; i < hi; ++i) loc += (double)__fdiv_rn(__fadd_rn(w0[i + 1], 1e-5f), norm);
    double incl = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    double run = incl - loc;                 // exclusive prefix of this lane's chunk (exact)
    if (lane == 0) cdf[0] = 0.0f;
    for (int i = lo; i < hi; ++i) {
      run += (double)__fdiv_rn(__fadd_rn(w0[i + 1], 1e-5f), norm);
      cdf[i + 1] = (float)run;
    }
  }
  __syncwarp();

  float zsum = 0.f;
  for (int j = lane; j < ni; j += 32) {
    const float u = a.u_rand ? a.u_rand[ray * ni + j] : a.u_det[j];
    // searchsorted(cdf, u, right=True): first index with cdf[idx] > u
    int lo = 0, hi = nb;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    const int ind = lo;
    const int below = max(ind - 1, 0), above = min(ind, nb - 1);
    const float cb = cdf[below], ca = cdf[above], bb = bins[below], ba = bins[above];
    float denom = ca - cb;
    if (denom < 1e-5f) denom = 1.0f;
    const float t = __fdiv_rn(u - cb, denom);
    const float zs = __fadd_rn(bb, __fmul_rn(t, ba - bb));
    keys[nc + j] = zs;
    a.z_samples[ray * ni + j] = zs;
    if (a.inds) a.inds[ray * ni + j] = ind;
    zsum += zs;
  }
  __syncwarp();
  if (a.z_std) {   // torch.std(z_samples, unbiased=False): two-pass
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
    const float mean = zsum / (float)ni;
    float v = 0.f;
    for (int j = lane; j < ni; j += 32) { const float dlt = keys[nc + j] - mean; v = fmaf(dlt, dlt, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) a.z_std[ray] = sqrtf(v / (float)ni);
  }
  // stable rank sort of cat[z_coarse, z_samples] (ties: lower index first, i.e. coarse before fine).  The coarse depths are
  // non-decreasing (stratified placement), so a coarse key's rank among the coarse keys is its index and a sample's count
  // of coarse keys <= it is a binary search: half the comparisons of the plain rank sort.  Unsorted coarse input (generic
  // callers) falls back to the plain O(n^2) rank sort.
  bool sorted = true;
  for (int i = lane; i < nc - 1; i += 32) sorted = sorted && (keys[i] <= keys[i + 1]);
  sorted = __all_sync(0xffffffffu, sorted);
  const float* smp = keys + nc;
  // deterministic u (perturb == 0) is increasing, so the samples come out non-decreasing as well: both halves are sorted runs and
  // the stable merge rank of every key is its own index plus one binary search in the other run
  bool smp_sorted = true;
  for (int j = lane; j < ni - 1; j += 32) smp_sorted = smp_sorted && (smp[j] <= smp[j + 1]);
  smp_sorted = __all_sync(0xffffffffu, smp_sorted) && sorted;
  for (int i = lane; i < nt; i += 32) {
    const float k = keys[i];
    int rank = 0;
    if (smp_sorted) {
      if (i < nc) {                               // coarse key: its index + number of samples strictly below it
        int lo = 0, hi = ni;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (smp[mid] < k) lo = mid + 1; else hi = mid; }
        rank = i + lo;
      } else {                                    // sample: its index among the samples + number of coarse keys <= it
        int lo = 0, hi = nc;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] <= k) lo = mid + 1; else hi = mid; }
        rank = (i - nc) + lo;
      }
    } else if (!sorted) {
      for (int j = 0; j < nt; ++j) {
        const float kj = keys[j];
        rank += (kj < k) || (kj == k && j < i);
      }
    } else if (i < nc) {
      rank = i;
      for (int j = 0; j < ni; ++j) rank += smp[j] < k;
    } else {
      int lo = 0, hi = nc;                       // number of coarse keys <= k
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] <= k) lo = mid + 1; else hi = mid;
      }
      rank = lo;
      const int js = i - nc;
      for (int j = 0; j < ni; ++j) {
        const float kj = smp[j];
        rank += (kj < k) || (kj == k && j < js);
      }
    }
    a.z_vals[ray * nt + rank] = k;
    if (a.order) a.order[ray * nt + rank] = i;
  }
}

}  // namespace edn

extern "C" int edn_sample_pdf_merge(const float* z_vals0, const float* weights0, const float* u_det, const float* u_rand,
                                    int64_t n_rays, int32_t n_samples, int32_t n_importance, float* z_samples,
                                    int64_t* inds, float* z_vals, int64_t* order, float* z_std, void* stream) {
  using namespace edn;
  EDN_REQUIRE(z_vals0 && weights0 && z_samples && z_vals, "edn_sample_pdf_merge: null pointer");
  EDN_REQUIRE(u_det || u_rand, "edn_sample_pdf_merge: need u_det or u_rand");
  EDN_REQUIRE(n_samples >= 3 && n_samples <= 1024, "edn_sample_pdf_merge: n_samples must be in [3,1024], got %d", n_samples);
  EDN_REQUIRE(n_importance >= 1 && n_importance <= 1024, "edn_sample_pdf_merge: n_importance must be in [1,1024], got %d", n_importance);
  if (n_rays <= 0) return n_rays == 0 ? EDN_OK : EDN_E_INVALID;
  PdfArgs a{z_vals0, weights0, u_det, u_rand, n_rays, n_samples, n_importance, z_samples, inds, z_vals, order, z_std};
  const size_t smem = (size_t)kPdfWarpsPerBlock * (n_samples + n_importance + 2 * (n_samples - 1)) * sizeof(float);
  const int64_t blocks = (n_rays + kPdfWarpsPerBlock - 1) / kPdfWarpsPerBlock;
  EDN_CUDA_OK(cudaFuncSetAttribute(sample_pdf_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample_pdf_merge_kernel<<<(unsigned)blocks, kPdfWarpsPerBlock * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

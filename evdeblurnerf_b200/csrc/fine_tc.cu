// Fine pass of render_rays on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands / fp32 accumulation.
// Replaces networks/renderer.py:190-217 + networks/pdrf/voxnerf.py:203-259,153-201 for the FVR field.
//
// One CTA renders one ray (M = 128 merged samples = one UMMA M tile) at a time; two CTAs are co-resident per SM so
// that one CTA's gather / epilogue overlaps the other's MMAs.  Per ray:
//   rows  (warps 0-3, thread = sample row): PE -> A[:,64:128]; cooperative VM gather of both grids -> two 128x96 bf16
//         tiles;  after each layer: TMEM -> registers -> bias/ReLU -> bf16 -> A operand of the next layer;  finally
//         sigma->alpha compositing with a warp-shuffle transmittance scan.
//   mma   (warp 4, one thread): streams every K=16 weight slice through a 5-stage shared-memory ring with bulk async
//         copies (TMA engine) and issues tcgen05.mma: basis_mat x2 (N=32), sigma_net 128->256->144(=128 geo + sigma),
//         color_net 128(+per-ray view-dir bias)->256->256->16(=rgb).  Accumulators live in 256 TMEM columns.
#include <vector>

#include "common.cuh"
#include "fine_args.cuh"
#include "tc_common.cuh"

namespace edn {
namespace {

using namespace tc;

constexpr int kRowThreads = 128;
constexpr int kThreads = 160;
constexpr int kNst = 5;                 // weight ring stages
constexpr int kStageBytes = 8192;       // one K=16 slice of an N=256 layer
constexpr int kABytes = 65536;          // 128 rows x 256 K bf16
constexpr int kChunkA = 2048;           // bytes of one 8-wide K chunk of a 128-row tile
constexpr uint32_t kTmemCols = 256;
constexpr int kChunksPerRay = 12 + 8 + 16 + 8 + 16 + 16;   // 76 MMA steps per ray
constexpr int kN2 = 144;                // sigma_net.1: 128 geo columns + sigma (col 128) + 15 zero columns
constexpr int kN5 = 16;                 // color_net.2: rgb in columns 0..2

enum : uint8_t { kFirst = 1, kLast = 2 };

struct Chunk {
  uint32_t goff;     // byte offset of the weight slice in the blob
  uint32_t a_off;    // byte offset of the A K-step inside the A buffer
  uint16_t bytes;
  uint16_t n;
  uint16_t d_col;
  uint8_t accum;
  uint8_t flags;
};
__constant__ Chunk c_sched[kChunksPerRay];

struct LayerDef { int K, N; };
// blob order: basis_coarse, basis_fine, sigma0, sigma1, color0, color1, color2
constexpr LayerDef kLayers[7] = {{96, 32}, {96, 32}, {128, 256}, {256, kN2}, {128, 256}, {256, 256}, {256, kN5}};

std::vector<Chunk> build_schedule(int64_t* blob_bytes, int64_t layer_off[7]) {
  std::vector<Chunk> v;
  uint32_t goff = 0;
  const int fine_tile_chunks[6] = {28, 30, 0, 2, 4, 6};
  for (int L = 0; L < 7; ++L) {
    layer_off[L] = goff;
    const int steps = kLayers[L].K / 16, N = kLayers[L].N;
    for (int j = 0; j < steps; ++j) {
      Chunk c{};
      c.goff = goff;
      c.bytes = (uint16_t)(N * 32);
      c.n = (uint16_t)N;
      c.accum = j > 0;
      c.flags = 0;
      if (L == 0) { c.a_off = (16 + 2 * j) * kChunkA; c.d_col = 0; if (j == 0) c.flags |= kFirst; }
      else if (L == 1) { c.a_off = fine_tile_chunks[j] * kChunkA; c.d_col = 32; if (j == steps - 1) c.flags |= kLast; }
      else { c.a_off = 2 * j * kChunkA; c.d_col = 0; if (j == 0) c.flags |= kFirst; if (j == steps - 1) c.flags |= kLast; }
      v.push_back(c);
      goff += c.bytes;
    }
  }
  *blob_bytes = goff;
  return v;
}

struct Misc {
  uint64_t bar_a, bar_acc, full[kNst], empty[kNst];
  uint32_t tmem_base, pad[3];
  float z[kRowThreads];
  float sig[kRowThreads];
  float bias[256];
  float red[4][8];
  float wtot[4];
};
constexpr int kSmemBytes = kABytes + kNst * kStageBytes + (int)sizeof(Misc);

// ---- weight packing -----------------------------------------------------------------------------------------------
// dst element (n, k) of a [K][N] layer -> bf16 index (k/16)*(N*16) + ((k%16)/8)*(N*8) + n*8 + k%8
__global__ void pack_layer_kernel(const float* __restrict__ wt, int ld, int k_valid, int n_valid, const float* __restrict__ vec,
                                  int vec_col, int K, int N, __nv_bfloat16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * N) return;
  const int k = i / N, n = i - k * N;
  float v = 0.f;
  if (k < k_valid && n < n_valid) v = wt[(size_t)k * ld + n];
  else if (vec && n == vec_col && k < k_valid) v = vec[k];
  dst[(size_t)(k / 16) * (N * 16) + ((k % 16) / 8) * (N * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(v);
}

// ---- small device helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fast_sincos(float x, float* s, float* c) {
  // Cody-Waite reduction to [-pi, pi] then MUFU; abs error ~5e-7, far below bf16 resolution of the MMA operand
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.28125f, x);
  r = fmaf(-k, 1.9353071795864769e-3f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}

template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

// 8 channels [c8*8, c8*8+8) of component `comp` at one point: (bilinear plane) * (linear line) -> 16 bytes of bf16
template <typename T>
__device__ __forceinline__ uint4 gather8(const T* __restrict__ plane, const T* __restrict__ line, int C, int c8,
                                         const Taps2& pt, const Taps1& lt) {
  float pv[4][8], lv[2][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) load8<T>(plane + (size_t)pt.off[k] * C + c8 * 8, pv[k]);
#pragma unroll
  for (int k = 0; k < 2; ++k) load8<T>(line + (size_t)lt.off[k] * C + c8 * 8, lv[k]);
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float p = pv[0][i] * pt.w[0];
    p = fmaf(pv[1][i], pt.w[1], p); p = fmaf(pv[2][i], pt.w[2], p); p = fmaf(pv[3][i], pt.w[3], p);
    const float l = fmaf(lv[1][i], lt.w[1], lv[0][i] * lt.w[0]);
    o[i] = p * l;
  }
  return make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}

// Cooperative gather of one grid for the 32 points of this warp: lane (q = lane/8, j = lane%8) serves point 8*gi+j.
template <typename T>
__device__ __forceinline__ void gather_grid(const GridDev& g, bool fine_tile, uint8_t* As, const float* z_s, int warp, int lane,
                                            const float o[3], const float d[3]) {
  const int q = lane >> 3;
#pragma unroll 2
  for (int gi = 0; gi < 4; ++gi) {
    const int pt = warp * 32 + gi * 8 + (lane & 7);
    const float zv = z_s[pt];
    float p[3], n[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], zv));
    normalize_pt(g, p, n);
    uint8_t* row = As + pt * 16;
    {  // component 0: plane (x,y) 64 channels, line z; this lane does channel chunks q and q+4
      Taps2 t2; Taps1 t1;
      plane_taps(n[0], n[1], g.ph[0], g.pw[0], t2);
      line_taps(n[2], g.ll[0], t1);
      const T* pl = reinterpret_cast<const T*>(g.plane[0]);
      const T* ln = reinterpret_cast<const T*>(g.line[0]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = 4 * h + q;
        const uint4 v = gather8<T>(pl, ln, 64, cc, t2, t1);
        const int chunk = fine_tile ? (cc < 4 ? 28 + cc : cc - 4) : 16 + cc;
        st_shared_v4(row + chunk * kChunkA, v.x, v.y, v.z, v.w);
      }
    }
    {  // components 1 (plane (x,z), line y) and 2 (plane (y,z), line x): 16 channels each = 2 chunks each
      const int comp = 1 + (q >> 1), c8 = q & 1;
      Taps2 t2; Taps1 t1;
      plane_taps(comp == 1 ? n[0] : n[1], n[2], g.ph[comp], g.pw[comp], t2);
      line_taps(comp == 1 ? n[1] : n[0], g.ll[comp], t1);
      const uint4 v = gather8<T>(reinterpret_cast<const T*>(g.plane[comp]), reinterpret_cast<const T*>(g.line[comp]), 16, c8, t2, t1);
      const int cc = 8 + q;
      const int chunk = fine_tile ? cc - 4 : 16 + cc;
      st_shared_v4(row + chunk * kChunkA, v.x, v.y, v.z, v.w);
    }
  }
}

// TMEM columns [col0, col0+32) of this thread's row -> (+bias) -> (ReLU) -> bf16 -> A chunks col0/8 .. col0/8+3
template <bool RELU>
__device__ __forceinline__ void epilogue32(uint32_t taddr_row, int col0, uint8_t* a_row, const float* bias_s,
                                           const float* __restrict__ bias_g, float* __restrict__ gout) {
  uint32_t v[32];
  tmem_ld32(taddr_row + col0, v);
  tmem_ld_wait();
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    f[i] = __uint_as_float(v[i]);
    if (bias_s) f[i] += bias_s[col0 + i];
    if (bias_g) f[i] += __ldg(bias_g + col0 + i);
  }
  if (gout) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(gout + col0 + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float x0 = f[8 * j + 2 * e], x1 = f[8 * j + 2 * e + 1];
      if (RELU) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
      pk[e] = pack_bf16x2(x0, x1);
    }
    st_shared_v4(a_row + (col0 / 8 + j) * kChunkA, pk[0], pk[1], pk[2], pk[3]);
  }
}

__device__ __forceinline__ void rows_signal_a(uint64_t* bar_a) {
  fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();          // order our tcgen05.ld's before the MMAs that will overwrite those TMEM columns
  mbar_arrive(bar_a);
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2) fine_fwd_tc_kernel(const FineArgs a, const uint8_t* __restrict__ blob) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;
  uint8_t* Ws = smem + kABytes;
  Misc* m = reinterpret_cast<Misc*>(smem + kABytes + kNst * kStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&m->bar_a, kRowThreads);
    mbar_init(&m->bar_acc, 1);
    for (int s = 0; s < kNst; ++s) { mbar_init(&m->full[s], 1); mbar_init(&m->empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&m->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = m->tmem_base;
  const int64_t n_my = (a.n_rays > (int64_t)blockIdx.x) ? (a.n_rays - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int S = a.S;

  if (warp == 4) {
    // =================================== MMA + weight-stream warp =====================================================
    if (lane == 0) {
      const uint32_t a_base = smem_u32(As), w_base = smem_u32(Ws);
      const uint64_t total = (uint64_t)n_my * kChunksPerRay;
      uint64_t g = 0, g_issue = 0;
      uint32_t pa = 0;
      for (int64_t it = 0; it < n_my; ++it) {
        for (int c = 0; c < kChunksPerRay; ++c, ++g) {
          while (g_issue < total && g_issue < g + kNst) {     // keep the ring full without blocking ahead of need
            const int s = (int)(g_issue % kNst);
            const uint64_t use = g_issue / kNst;
            if (use > 0) {
              const uint32_t par = (uint32_t)((use - 1) & 1);
              if (g_issue == g) mbar_wait(&m->empty[s], par);
              else if (!mbar_test_wait(&m->empty[s], par)) break;
            }
            const Chunk& ci = c_sched[g_issue % kChunksPerRay];
            mbar_expect_tx(&m->full[s], ci.bytes);
            bulk_g2s(Ws + s * kStageBytes, blob + ci.goff, ci.bytes, &m->full[s]);
            ++g_issue;
          }
          const Chunk& ch = c_sched[c];
          if (ch.flags & kFirst) { mbar_wait(&m->bar_a, pa); pa ^= 1; }
          const int s = (int)(g % kNst);
          mbar_wait(&m->full[s], (uint32_t)((g / kNst) & 1));
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(a_base + ch.a_off, kChunkA, 128);
          const uint64_t bdesc = make_smem_desc(w_base + s * kStageBytes, (uint32_t)ch.n * 16u, 128);
          mma_bf16_ss(tmem + ch.d_col, adesc, bdesc, make_idesc_bf16(128, ch.n), ch.accum);
          mma_commit(&m->empty[s]);
          if (ch.flags & kLast) mma_commit(&m->bar_acc);
        }
      }
    }
    __syncwarp();
  } else {
    // =================================== row warps: thread = sample row ===============================================
    const int r = tid;
    uint8_t* a_row = As + r * 16;
    const uint32_t taddr_row = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t pacc = 0;
    const bool mask_near = !(a.flags & EDN_FLAG_TRAIN) && a.rmnearplane > 0.f;
    const float near_thr = a.rmnearplane / 128.0f;
    for (int64_t it = 0; it < n_my; ++it) {
      const int64_t ray = (int64_t)blockIdx.x + it * gridDim.x;
      const float* rb = a.ray_batch + ray * 11;
      const float o[3] = {__ldg(rb + 0), __ldg(rb + 1), __ldg(rb + 2)};
      const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
      const float zv = a.z_vals[ray * S + min(r, S - 1)];
      m->z[r] = zv;
      {  // ---- PE(pts) -> A columns 64..127 (chunks 8..15); column 127 is the zero pad of K = 127 -> 128 -------------
        float pe[64];
#pragma unroll
        for (int i = 0; i < 3; ++i) pe[i] = __fadd_rn(o[i], __fmul_rn(d[i], zv));
#pragma unroll
        for (int f = 0; f < kPeFreqPts; ++f) {
#pragma unroll
          for (int i = 0; i < 3; ++i) fast_sincos(pe[i] * (float)(1 << f), &pe[3 + 6 * f + i], &pe[6 + 6 * f + i]);
        }
        pe[63] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(a_row + (8 + j) * kChunkA, pack_bf16x2(pe[8 * j], pe[8 * j + 1]), pack_bf16x2(pe[8 * j + 2], pe[8 * j + 3]),
                       pack_bf16x2(pe[8 * j + 4], pe[8 * j + 5]), pack_bf16x2(pe[8 * j + 6], pe[8 * j + 7]));
      }
      {  // ---- per-ray bias of color_net.0: b0 + W0[:, 128:155] . PE(viewdir) (fp32, exact sincos) --------------------
        const float vd[3] = {__ldg(rb + 8), __ldg(rb + 9), __ldg(rb + 10)};
        float ped[kPeDir];
#pragma unroll
        for (int i = 0; i < 3; ++i) ped[i] = vd[i];
#pragma unroll
        for (int f = 0; f < kPeFreqDir; ++f) {
#pragma unroll
          for (int i = 0; i < 3; ++i) sincosf(vd[i] * (float)(1 << f), &ped[3 + 6 * f + i], &ped[6 + 6 * f + i]);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col = r + half * kRowThreads;
          float b = a.mlp.color0_b ? __ldg(a.mlp.color0_b + col) : 0.f;
          const float* w = a.mlp.color0_t + (size_t)128 * 256 + col;
#pragma unroll
          for (int j = 0; j < kPeDir; ++j) b = fmaf(__ldg(w + j * 256), ped[j], b);
          m->bias[col] = b;
        }
      }
      named_bar_sync(1, kRowThreads);      // z[] (and bias[]) visible to all row warps
      // ---- VM gather of both grids -> two 128 x 96 bf16 tiles ----------------------------------------------------------
      gather_grid<T>(a.gc, false, As, m->z, warp, lane, o, d);
      gather_grid<T>(a.gf, true, As, m->z, warp, lane, o, d);
      rows_signal_a(&m->bar_a);
      // ---- basis_mat outputs (coarse 32 | fine 32) -> A columns 0..63 ------------------------------------------------
      mbar_wait(&m->bar_acc, pacc); pacc ^= 1; tc_fence_after();
      epilogue32<false>(taddr_row, 0, a_row, nullptr, nullptr, nullptr);
      epilogue32<false>(taddr_row, 32, a_row, nullptr, nullptr, nullptr);
      rows_signal_a(&m->bar_a);
      // ---- sigma_net.0 -> ReLU ---------------------------------------------------------------------------------------------
      mbar_wait(&m->bar_acc, pacc); pacc ^= 1; tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) epilogue32<true>(taddr_row, c0, a_row, nullptr, nullptr, nullptr);
      rows_signal_a(&m->bar_a);
      // ---- sigma_net.1 -> geo (128, linear) + sigma ---------------------------------------------------------------------
      mbar_wait(&m->bar_acc, pacc); pacc ^= 1; tc_fence_after();
      {
        float* gout = (a.feat && r < S) ? a.feat + ((size_t)ray * S + r) * 128 : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) epilogue32<false>(taddr_row, c0, a_row, nullptr, nullptr, gout);
        uint32_t v[16];
        tmem_ld16(taddr_row + 128, v);
        tmem_ld_wait();
        m->sig[r] = __uint_as_float(v[0]);
      }
      rows_signal_a(&m->bar_a);
      // ---- color_net.0 (+ per-ray view-dir bias) -> ReLU ------------------------------------------------------------------
      mbar_wait(&m->bar_acc, pacc); pacc ^= 1; tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) epilogue32<true>(taddr_row, c0, a_row, m->bias, nullptr, nullptr);
      rows_signal_a(&m->bar_a);
      // ---- color_net.1 -> ReLU -------------------------------------------------------------------------------------------------
      mbar_wait(&m->bar_acc, pacc); pacc ^= 1; tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) epilogue32<true>(taddr_row, c0, a_row, nullptr, a.mlp.color1_b, nullptr);
      rows_signal_a(&m->bar_a);
      // ---- color_net.2 -> sigmoid, then compositing (voxnerf.py:153-201) ---------------------------------------------------
      mbar_wait(&m->bar_acc, pacc); pacc ^= 1; tc_fence_after();
      float col[3];
      {
        uint32_t v[16];
        tmem_ld16(taddr_row, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 3; ++i) col[i] = sigmoidf_(__uint_as_float(v[i]) + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + i) : 0.f));
      }
      float alpha = 0.f;
      if (r < S - 1) {
        const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
        const float znext = m->z[r + 1];
        const float dist = __fmul_rn(znext - zv, dnorm);
        float sg = m->sig[r];
        if (a.noise) sg += __ldg(a.noise + ray * (S - 1) + r);
        sg = fmaxf(sg, 0.f);
        if (mask_near && !(znext > near_thr)) sg = 0.f;
        alpha = 1.0f - expf(-__fmul_rn(sg, dist));
      } else if (r == S - 1) {
        alpha = 1.0f;
      }
      float t = 1.0f - alpha;                 // inclusive product scan of (1 - alpha) over the warp
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, t, dlt);
        if (lane >= dlt) t *= y;
      }
      float Tr = __shfl_up_sync(0xffffffffu, t, 1);
      if (lane == 0) Tr = 1.0f;
      if (lane == 31) m->wtot[warp] = t;
      named_bar_sync(1, kRowThreads);
      for (int w2 = 0; w2 < warp; ++w2) Tr *= m->wtot[w2];
      const float wgt = alpha * Tr;
      if (r < S) a.weights[ray * S + r] = wgt;
      float red[5] = {wgt * col[0], wgt * col[1], wgt * col[2], wgt * zv, wgt};
#pragma unroll
      for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) red[i] += __shfl_xor_sync(0xffffffffu, red[i], off);
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) m->red[warp][i] = red[i];
      }
      named_bar_sync(1, kRowThreads);
      if (r < 5) {
        const float v = m->red[0][r] + m->red[1][r] + m->red[2][r] + m->red[3][r];
        if (r < 3) a.rgb[ray * 3 + r] = v; else if (r == 3) a.depth[ray] = v; else a.acc[ray] = v;
      }
      named_bar_sync(1, kRowThreads);      // red[] / wtot[] / z[] free for the next ray
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

int ensure_schedule(int64_t* blob_bytes, int64_t layer_off[7]) {
  static bool uploaded = false;
  static int64_t bytes = 0, offs[7];
  if (!uploaded) {
    std::vector<Chunk> v = build_schedule(&bytes, offs);
    if ((int)v.size() != kChunksPerRay) { set_error("fine_tc: schedule size mismatch"); return EDN_E_INVALID; }
    EDN_CUDA_OK(cudaMemcpyToSymbol(c_sched, v.data(), sizeof(Chunk) * v.size()));
    uploaded = true;
  }
  if (blob_bytes) *blob_bytes = bytes;
  if (layer_off) for (int i = 0; i < 7; ++i) layer_off[i] = offs[i];
  return 0;
}

}  // namespace

int launch_fine_tc(const FineArgs& a, int grid_dtype, cudaStream_t st) {
  EDN_REQUIRE(a.S >= 2 && a.S <= kRowThreads, "edn_render_fine_fwd(bf16): n_samples must be in [2,128], got %d", a.S);
  EDN_REQUIRE(a.mlp.tc_blob != nullptr, "edn_render_fine_fwd(bf16): edn_field_mlp.tc_blob is NULL (call edn_pack_fine_tc)");
  int rc = ensure_schedule(nullptr, nullptr);
  if (rc) return rc;
  const int64_t max_ctas = 2 * (int64_t)num_sms();
  const unsigned gx = (unsigned)(a.n_rays < max_ctas ? a.n_rays : max_ctas);
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(a.mlp.tc_blob);
  if (grid_dtype == EDN_BF16) {
    EDN_CUDA_OK(cudaFuncSetAttribute(fine_fwd_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    fine_fwd_tc_kernel<__nv_bfloat16><<<gx, kThreads, kSmemBytes, st>>>(a, blob);
  } else {
    EDN_CUDA_OK(cudaFuncSetAttribute(fine_fwd_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    fine_fwd_tc_kernel<float><<<gx, kThreads, kSmemBytes, st>>>(a, blob);
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

}  // namespace edn

extern "C" int64_t edn_fine_tc_blob_bytes(void) {
  int64_t bytes = 0, offs[7];
  edn::build_schedule(&bytes, offs);
  return bytes;
}

extern "C" int edn_pack_fine_tc(const edn_field_mlp* mlp, const float* basis_t_coarse, const float* basis_t_fine, void* blob,
                                void* stream) {
  using namespace edn;
  EDN_REQUIRE(mlp && basis_t_coarse && basis_t_fine && blob, "edn_pack_fine_tc: null pointer");
  EDN_REQUIRE(mlp->hidden == 256 && mlp->geo_feat == 128 && mlp->sigma0_t && mlp->sigma1_t && mlp->sigma1_v && mlp->color0_t &&
              mlp->color1_t && mlp->color2_t, "edn_pack_fine_tc: needs the fine field (hidden=256, geo_feat=128)");
  int64_t bytes, off[7];
  int rc = ensure_schedule(&bytes, off);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* b = reinterpret_cast<uint8_t*>(blob);
  struct Src { const float* wt; int ld, kv, nv; const float* vec; int vec_col; };
  const Src src[7] = {{basis_t_coarse, 32, 96, 32, nullptr, -1}, {basis_t_fine, 32, 96, 32, nullptr, -1},
                      {mlp->sigma0_t, 256, 128, 256, nullptr, -1}, {mlp->sigma1_t, 128, 256, 128, mlp->sigma1_v, 128},
                      {mlp->color0_t, 256, 128, 256, nullptr, -1}, {mlp->color1_t, 256, 256, 256, nullptr, -1},
                      {mlp->color2_t, 4, 256, 3, nullptr, -1}};
  for (int L = 0; L < 7; ++L) {
    const int K = kLayers[L].K, N = kLayers[L].N, total = K * N;
    pack_layer_kernel<<<(total + 255) / 256, 256, 0, st>>>(src[L].wt, src[L].ld, src[L].kv, src[L].nv, src[L].vec, src[L].vec_col, K, N,
                                                           reinterpret_cast<__nv_bfloat16*>(b + off[L]));
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

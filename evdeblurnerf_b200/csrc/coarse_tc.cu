// Coarse pass of render_rays on the tensor cores (tcgen05 + TMEM), bf16 operands / fp32 accumulation.
// Replaces networks/renderer.py:157-188 + networks/pdrf/voxnerf.py:203-259,153-201 for the CRR coarse field.
//
// One persistent CTA per SM; two row groups of 128 threads, each group renders one 128-row tile = floor(128 / Nc) rays
// x Nc coarse samples at a time (2 rays for Nc = 64).  All weights (30 KB of bf16 UMMA slices) are resident in shared
// memory; the small CTA footprint leaves ~90 KB of L1 for the VM line tables.  Per tile:
//   rows: bit-exact sample placement, PE -> A[:,32:96], cooperative VM gather -> 128x96 tile, layer epilogues
//         (TMEM -> regs -> bias/ReLU -> bf16 -> next A operand), fp32 rgb head, sequential per-ray compositing;
//   mma (one thread per group): basis_mat (N=32), sigma_net 96->64->16, color_net 16(+view-dir bias)->64->64.
#include <cstddef>

#include "coarse_args.cuh"
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_rows.cuh"

namespace edn {
namespace {

using namespace tc;

constexpr int kGroupThreads = 128;
constexpr int kRowWarps = 8;
constexpr int kThreads = kRowWarps * 32 + 64;     // + 2 MMA issuer warps
constexpr int kAChunks = 24;                      // per group: layer operand chunks 0..11, gather tile chunks 12..23
constexpr int kABytes = kAChunks * kChunkA;       // 48 KB
constexpr uint32_t kTmemCols = 128;               // 2 groups x 64 columns
// resident weight blob (bytes): basis 6x1K | sigma0 6x2K | sigma1 4x512 | color0(geo) 1x2K | color1 4x2K
constexpr int kOffBasis = 0, kOffS0 = 6144, kOffS1 = 18432, kOffC0 = 20480, kOffC1 = 22528, kWBytes = 30720;
constexpr int kMaxRpt = 4;                        // rays per tile (Nc >= 32)

struct alignas(16) GroupMisc {
  float z[kGroupThreads];
  float sig[kGroupThreads];
  float w[kGroupThreads];
  float rgb[kGroupThreads * 3];
  alignas(16) float bias[kMaxRpt][64];
};
struct Misc {
  uint64_t bar_a[2], bar_acc[2], bar_w;
  GridDev grid;
  uint32_t tmem_base, pad[3];
  alignas(16) float wdir[kPeDir][64];   // color_net.0 rows 15..41 (view-direction part), fp32
  alignas(16) float b0[64];
  alignas(16) float b1[64];
  alignas(16) float wrgb[64][4];
  GroupMisc grp[2];
};
constexpr int kSmemBytes = 2 * kABytes + kWBytes + (int)sizeof(Misc);
static_assert(offsetof(Misc, wdir) % 16 == 0 && offsetof(Misc, grp) % 16 == 0 && offsetof(GroupMisc, bias) % 16 == 0, "alignment");

// 32 accumulator columns -> (+bias) -> (ReLU) -> bf16 -> A chunks col0/8.. ; returns the fp32 values in f[]
__device__ __forceinline__ void epi32(const uint32_t (&v)[32], int col0, bool relu, const float* bias_s, uint8_t* a_row,
                                      float (&f)[32], bool store) {
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  if (bias_s) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + col0 + i);
      f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
    }
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
  }
  if (store) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      st_shared_v4(a_row + (col0 / 8 + j) * kChunkA, pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                   pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
  }
}

__device__ __forceinline__ void rows_signal(uint64_t* bar_a) {
  fence_proxy_async_smem();
  tc_fence_before();
  mbar_arrive(bar_a);
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) coarse_fwd_tc_kernel(const CoarseArgs a, const uint8_t* __restrict__ blob) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;
  uint8_t* Wsm = smem + 2 * kABytes;
  Misc* m = reinterpret_cast<Misc*>(Wsm + kWBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.n_samples;
  const int rpt = kGroupThreads / S;                       // rays per tile (host guarantees 32 <= S <= 128)
  const int64_t n_tiles = (a.n_rays + rpt - 1) / rpt;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const int64_t n_my = (n_pairs > (int64_t)blockIdx.x) ? (n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int q = 0; q < 2; ++q) { mbar_init(&m->bar_a[q], kGroupThreads); mbar_init(&m->bar_acc[q], 1); }
    mbar_init(&m->bar_w, 1);
    fence_barrier_init();
    m->grid = a.grid;
  }
  if (warp == kRowWarps) tmem_alloc(&m->tmem_base, kTmemCols);
  for (int i = tid; i < kPeDir * 64; i += kThreads) m->wdir[i / 64][i % 64] = __ldg(a.mlp.color0_t + (15 + i / 64) * 64 + i % 64);
  for (int i = tid; i < 64; i += kThreads) {
    m->b0[i] = a.mlp.color0_b ? __ldg(a.mlp.color0_b + i) : 0.f;
    m->b1[i] = a.mlp.color1_b ? __ldg(a.mlp.color1_b + i) : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) m->wrgb[i][j] = __ldg(a.mlp.color2_t + i * 4 + j);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = m->tmem_base;

  if (warp >= kRowWarps) {
    // =================================== MMA issuer warp of group q (one thread) =======================================
    const int q = warp - kRowWarps;
    if (lane == 0 && n_my > 0) {
      if (q == 0) { mbar_expect_tx(&m->bar_w, kWBytes); bulk_g2s(Wsm, blob, kWBytes, &m->bar_w); }
      const uint32_t aq = smem_u32(As) + q * kABytes, wb = smem_u32(Wsm);
      const uint32_t d_tmem = tmem + q * 64;
      uint32_t pa = 0;
      mbar_wait(&m->bar_w, 0);
      for (int64_t it = 0; it < n_my; ++it) {
        // basis_mat: gather tile (chunks 12..23) x [96 -> 32]
        mbar_wait(&m->bar_a[q], pa); pa ^= 1; tc_fence_after();
#pragma unroll
        for (int j = 0; j < 6; ++j)
          mma_bf16_ss(d_tmem, make_smem_desc(aq + (12 + 2 * j) * kChunkA, kChunkA, 128),
                      make_smem_desc(wb + kOffBasis + j * 1024, 32 * 16, 128), make_idesc_bf16(128, 32), j > 0);
        mma_commit(&m->bar_acc[q]);
        // sigma_net.0: [ft 32 | PE 63 | 0] -> 64
        mbar_wait(&m->bar_a[q], pa); pa ^= 1; tc_fence_after();
#pragma unroll
        for (int j = 0; j < 6; ++j)
          mma_bf16_ss(d_tmem, make_smem_desc(aq + 2 * j * kChunkA, kChunkA, 128),
                      make_smem_desc(wb + kOffS0 + j * 2048, 64 * 16, 128), make_idesc_bf16(128, 64), j > 0);
        mma_commit(&m->bar_acc[q]);
        // sigma_net.1: 64 -> 16 (columns 0..14 geo, 15 sigma)
        mbar_wait(&m->bar_a[q], pa); pa ^= 1; tc_fence_after();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma_bf16_ss(d_tmem, make_smem_desc(aq + 2 * j * kChunkA, kChunkA, 128),
                      make_smem_desc(wb + kOffS1 + j * 512, 16 * 16, 128), make_idesc_bf16(128, 16), j > 0);
        mma_commit(&m->bar_acc[q]);
        // color_net.0, geo part: 16 (15 + zero row) -> 64
        mbar_wait(&m->bar_a[q], pa); pa ^= 1; tc_fence_after();
        mma_bf16_ss(d_tmem, make_smem_desc(aq, kChunkA, 128), make_smem_desc(wb + kOffC0, 64 * 16, 128), make_idesc_bf16(128, 64), 0);
        mma_commit(&m->bar_acc[q]);
        // color_net.1: 64 -> 64
        mbar_wait(&m->bar_a[q], pa); pa ^= 1; tc_fence_after();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma_bf16_ss(d_tmem, make_smem_desc(aq + 2 * j * kChunkA, kChunkA, 128),
                      make_smem_desc(wb + kOffC1 + j * 2048, 64 * 16, 128), make_idesc_bf16(128, 64), j > 0);
        mma_commit(&m->bar_acc[q]);
      }
    }
    __syncwarp();
  } else {
    // =================================== row warps: thread = (ray in tile, sample) =====================================
    const int q = warp >> 2, gwarp = warp & 3;
    const int r = tid & (kGroupThreads - 1);
    uint8_t* Aq = As + q * kABytes;
    uint8_t* a_row = Aq + r * 16;
    GroupMisc* gm = &m->grp[q];
    const uint32_t taddr_row = tmem + ((uint32_t)(gwarp * 32) << 16) + q * 64;
    uint32_t pacc = 0;
    const int bar_id = 1 + q;
    const int lr = min(r / S, rpt - 1), s = (r / S < rpt) ? r - lr * S : S - 1;   // padding rows replay the last sample
    const bool row_valid = r < rpt * S;
    const bool mask_near = !(a.flags & EDN_FLAG_TRAIN) && a.rmnearplane > 0.f;
    for (int64_t it = 0; it < n_my; ++it) {
      const int64_t tile = 2 * ((int64_t)blockIdx.x + it * gridDim.x) + q;
      const int64_t ray_raw = tile * rpt + lr;
      const bool live = row_valid && ray_raw < a.n_rays;
      const int64_t ray = ray_raw < a.n_rays ? ray_raw : a.n_rays - 1;
      const float* rb = a.ray_batch + ray * 11;
      const float o[3] = {__ldg(rb + 0), __ldg(rb + 1), __ldg(rb + 2)};
      const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
      const float zv = place_sample(a, ray, s, __ldg(rb + 6), __ldg(rb + 7));
      gm->z[r] = zv;
      {  // ---- PE(pts) -> A columns 32..95 (chunks 4..11); column 95 is the zero pad of K = 95 -> 96 ---------------------
        float pe[64];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          pe[i] = __fadd_rn(o[i], __fmul_rn(d[i], zv));
          fast_sincos(pe[i], &pe[3 + i], &pe[6 + i]);
        }
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
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(a_row + (4 + j) * kChunkA, pack_bf16x2(pe[8 * j], pe[8 * j + 1]), pack_bf16x2(pe[8 * j + 2], pe[8 * j + 3]),
                       pack_bf16x2(pe[8 * j + 4], pe[8 * j + 5]), pack_bf16x2(pe[8 * j + 6], pe[8 * j + 7]));
      }
      // ---- per-ray bias of color_net.0: b0 + W0[:, 15:42] . PE(viewdir), fp32 --------------------------------------------
#pragma unroll 1
      for (int idx = r; idx < rpt * 64; idx += kGroupThreads) {
        const int br = idx >> 6, col = idx & 63;
        const int64_t bray = min(tile * rpt + br, a.n_rays - 1);
        const float* rb2 = a.ray_batch + bray * 11;
        float ped[kPeDir];
#pragma unroll
        for (int i = 0; i < 3; ++i) { ped[i] = __ldg(rb2 + 8 + i); fast_sincos(ped[i], &ped[3 + i], &ped[6 + i]); }
#pragma unroll
        for (int f = 1; f < kPeFreqDir; ++f) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float sp = ped[3 + 6 * (f - 1) + i], cp = ped[6 + 6 * (f - 1) + i];
            ped[3 + 6 * f + i] = 2.0f * sp * cp;
            ped[6 + 6 * f + i] = fmaf(-2.0f * sp, sp, 1.0f);
          }
        }
        float b = m->b0[col];
#pragma unroll
        for (int j = 0; j < kPeDir; ++j) b = fmaf(m->wdir[j][col], ped[j], b);
        gm->bias[br][col] = b;
      }
      named_bar_sync(bar_id, kGroupThreads);
      // ---- VM gather of the coarse grid -> 128 x 96 bf16 tile (chunks 12..23) ----------------------------------------------
      {
        const GridDev& g = m->grid;
        const int qq = lane >> 3;
#pragma unroll 1
        for (int gi = 0; gi < 4; ++gi) {
          const int pt = gwarp * 32 + gi * 8 + (lane & 7);
          const int plr = min(pt / S, rpt - 1);
          // the point's own ray may differ from this lane's ray: recompute it from the ray batch (rays of a tile are adjacent)
          const int64_t pray = min(tile * rpt + plr, a.n_rays - 1);
          const float* rbp = a.ray_batch + pray * 11;
          const float zp = gm->z[pt];
          float p[3], n[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(__ldg(rbp + i), __fmul_rn(__ldg(rbp + 3 + i), zp));
          normalize_pt(g, p, n);
          uint8_t* row = Aq + pt * 16;
          GatherTask<T> t0, t1, t2;
          {
            Taps2 pt2; Taps1 lt1;
            plane_taps(n[0], n[1], g.ph[0], g.pw[0], pt2);
            line_taps(n[2], g.ll[0], lt1);
            const T* pl = reinterpret_cast<const T*>(g.plane[0]);
            const T* ln = reinterpret_cast<const T*>(g.line[0]);
            t0.issue(pl, ln, 64, qq, pt2, lt1);
            t1.issue(pl, ln, 64, qq + 4, pt2, lt1);
          }
          {
            const int comp = 1 + (qq >> 1);
            Taps2 pt2; Taps1 lt1;
            plane_taps(comp == 1 ? n[0] : n[1], n[2], g.ph[comp], g.pw[comp], pt2);
            line_taps(comp == 1 ? n[1] : n[0], g.ll[comp], lt1);
            t2.issue(reinterpret_cast<const T*>(g.plane[comp]), reinterpret_cast<const T*>(g.line[comp]), 16, qq & 1, pt2, lt1);
          }
          t0.finish(row + (12 + qq) * kChunkA);
          t1.finish(row + (16 + qq) * kChunkA);
          t2.finish(row + (20 + qq) * kChunkA);
        }
      }
      rows_signal(&m->bar_a[q]);
      uint32_t v[32];
      float f[32];
      // ---- basis_mat -> ft (32) -> A columns 0..31 -----------------------------------------------------------------------------
      mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
      tmem_ld32(taddr_row, v); tmem_ld_wait();
      epi32(v, 0, false, nullptr, a_row, f, true);
      rows_signal(&m->bar_a[q]);
      // ---- sigma_net.0 -> ReLU -> A columns 0..63 ---------------------------------------------------------------------------------
      mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) { tmem_ld32(taddr_row + c0, v); tmem_ld_wait(); epi32(v, c0, true, nullptr, a_row, f, true); }
      rows_signal(&m->bar_a[q]);
      // ---- sigma_net.1 -> geo (cols 0..14) + sigma (col 15); A columns 0..15 (col 15 meets a zero weight row) ---------------------
      mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
      float sig_raw;
      {
        uint32_t v16[16];
        tmem_ld16(taddr_row, v16); tmem_ld_wait();
        float g16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) g16[i] = __uint_as_float(v16[i]);
        sig_raw = g16[15];
        if (a.feat && live) {
          float* fo = a.feat + (ray * S + s) * 15;
#pragma unroll
          for (int j = 0; j < 15; ++j) fo[j] = g16[j];
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
          st_shared_v4(a_row + j * kChunkA, pack_bf16x2(g16[8 * j], g16[8 * j + 1]), pack_bf16x2(g16[8 * j + 2], g16[8 * j + 3]),
                       pack_bf16x2(g16[8 * j + 4], g16[8 * j + 5]), pack_bf16x2(g16[8 * j + 6], g16[8 * j + 7]));
      }
      rows_signal(&m->bar_a[q]);
      // ---- color_net.0 (+ per-ray view-dir bias) -> ReLU -> A columns 0..63 --------------------------------------------------------
      mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) { tmem_ld32(taddr_row + c0, v); tmem_ld_wait(); epi32(v, c0, true, gm->bias[lr], a_row, f, true); }
      rows_signal(&m->bar_a[q]);
      // ---- color_net.1 -> ReLU -> rgb head (fp32) -> sigmoid -------------------------------------------------------------------------
      mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
      float col[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        tmem_ld32(taddr_row + c0, v); tmem_ld_wait();
        epi32(v, c0, true, m->b1, a_row, f, false);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 w = *reinterpret_cast<const float4*>(&m->wrgb[c0 + i][0]);
          col[0] = fmaf(f[i], w.x, col[0]); col[1] = fmaf(f[i], w.y, col[1]); col[2] = fmaf(f[i], w.z, col[2]);
        }
      }
      {  // ---- compositing (voxnerf.py:153-201): alpha per row in parallel, then one thread per ray in sample order ---------
        const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
        const bool is_last = (s == S - 1);
        const float z_next = gm->z[min(r + 1, kGroupThreads - 1)];
        const float nz = (a.noise && !is_last) ? __ldg(a.noise + ray * (S - 1) + s) : 0.f;
        gm->sig[r] = alpha_of_sample(sig_raw, zv, z_next, nz, dnorm, mask_near, a.rmnearplane / 128.0f, is_last);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) gm->rgb[3 * r + i] = sigmoidf_(col[i] + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + i) : 0.f));
      named_bar_sync(bar_id, kGroupThreads);
      if (r < rpt) {
        const int64_t r2 = tile * rpt + r;
        if (r2 < a.n_rays) {
          float out[5];
          composite_from_alpha(gm->sig + r * S, gm->rgb + 3 * r * S, gm->z + r * S, S, (a.flags & EDN_FLAG_RELU_RGB) != 0, gm->w + r * S, out);
          a.rgb[r2 * 3 + 0] = out[0]; a.rgb[r2 * 3 + 1] = out[1]; a.rgb[r2 * 3 + 2] = out[2];
          a.depth[r2] = out[3];
          a.acc[r2] = out[4];
        }
      }
      named_bar_sync(bar_id, kGroupThreads);
      if (live) {
        a.z_vals[ray * S + s] = zv;
        a.weights[ray * S + s] = gm->w[r];
      }
      named_bar_sync(bar_id, kGroupThreads);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kRowWarps) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

int launch_coarse_tc(const CoarseArgs& a, int grid_dtype, cudaStream_t st) {
  EDN_REQUIRE(a.n_samples >= 32 && a.n_samples <= kGroupThreads,
              "edn_render_coarse_fwd(bf16): n_samples must be in [32,128], got %d", a.n_samples);
  EDN_REQUIRE(a.mlp.tc_blob != nullptr, "edn_render_coarse_fwd(bf16): edn_field_mlp.tc_blob is NULL (call edn_pack_coarse_tc)");
  const int rpt = kGroupThreads / a.n_samples;
  const int64_t n_tiles = (a.n_rays + rpt - 1) / rpt, n_pairs = (n_tiles + 1) / 2;
  const unsigned gx = (unsigned)(n_pairs < (int64_t)num_sms() ? n_pairs : (int64_t)num_sms());
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(a.mlp.tc_blob);
  if (grid_dtype == EDN_BF16) {
    EDN_CUDA_OK(cudaFuncSetAttribute(coarse_fwd_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    coarse_fwd_tc_kernel<__nv_bfloat16><<<gx, kThreads, kSmemBytes, st>>>(a, blob);
  } else {
    EDN_CUDA_OK(cudaFuncSetAttribute(coarse_fwd_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    coarse_fwd_tc_kernel<float><<<gx, kThreads, kSmemBytes, st>>>(a, blob);
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

}  // namespace edn

extern "C" int64_t edn_coarse_tc_blob_bytes(void) { return edn::kWBytes; }

extern "C" int edn_pack_coarse_tc(const edn_field_mlp* mlp, const float* basis_t, void* blob, void* stream) {
  using namespace edn;
  EDN_REQUIRE(mlp && basis_t && blob, "edn_pack_coarse_tc: null pointer");
  EDN_REQUIRE(mlp->hidden == 64 && mlp->geo_feat == 15 && mlp->sigma0_t && mlp->sigma1_t && mlp->color0_t && mlp->color1_t &&
              mlp->color2_t, "edn_pack_coarse_tc: needs the coarse field (hidden=64, geo_feat=15)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* b = reinterpret_cast<uint8_t*>(blob);
  struct Src { const float* wt; int ld, kv, nv, K, N, rot, off; };
  const Src src[5] = {{basis_t, 32, 96, 32, 96, 32, 0, kOffBasis},
                      {mlp->sigma0_t, 64, 96, 64, 96, 64, 0, kOffS0},
                      {mlp->sigma1_t, 16, 64, 16, 64, 16, 1, kOffS1},      // output col j <- sigma_net.1 row (j+1)%16: geo first, sigma last
                      {mlp->color0_t, 64, 15, 64, 16, 64, 0, kOffC0},      // geo rows only; K row 15 (the sigma column) is zero
                      {mlp->color1_t, 64, 64, 64, 64, 64, 0, kOffC1}};
  for (int L = 0; L < 5; ++L) {
    const int total = src[L].K * src[L].N;
    tc::pack_layer_kernel<<<(total + 255) / 256, 256, 0, st>>>(src[L].wt, src[L].ld, src[L].kv, src[L].nv, src[L].K, src[L].N, src[L].rot,
                                                               reinterpret_cast<__nv_bfloat16*>(b + src[L].off));
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

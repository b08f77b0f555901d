// Coarse pass of render_rays on the tensor cores (tcgen05 + TMEM), bf16 operands / fp32 accumulation.
// Replaces networks/renderer.py:157-188 + networks/pdrf/voxnerf.py:203-259,153-201 for the CRR coarse field.
// Three kernels share this file's layouts and helpers:
//   coarse_fwd_tc_kernel  (round 1; today: S not a multiple of 32, the full / feature_map schedule, EDN_COARSE_V1=1) -- described next;
//   coarse_fwd_tc2_kernel (default, lean schedule): producer warps decoupled from the MLP chain, see its own header further down;
//   coarse_fwd_tc3_kernel (precision EDN_TC32): the bf16 x 3 split-operand parity kernel.
//
// Round-1 kernel: one persistent CTA per SM; two row groups, each renders one 128-row tile = floor(128 / Nc) rays x Nc coarse samples at a
// time (2 rays for Nc = 64) with TWO threads per row (16 row warps: the halves split the gather and the epilogue
// columns).  All weights are resident in shared memory; the small CTA footprint leaves L1 for the VM line tables.
//   rows: bit-exact sample placement, PE -> A, cooperative VM gather -> 128x96 tile, layer epilogues
//         (TMEM -> regs -> bias/ReLU -> bf16 -> next A operand), fp32 sigma / rgb heads, per-ray compositing in sample order;
//   mma (one thread per group), two schedules:
//     lean (default): basis_mat folded into sigma_net.0 ([g 96 | PE 64] = 160 -> 64), sigma_net.1's geo columns folded
//                     into color_net.0 (64 -> 64, + view-dir bias), color_net.1 (64 -> 64): 3 stages;
//     full (feature_map requested): basis_mat (96 -> 32), sigma_net 96 -> 64 -> 16, color_net 16 -> 64 -> 64: 5 stages.
#include <cstddef>
#include <cstdio>
#include <cstdlib>

#include "coarse_args.cuh"
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_rows.cuh"

namespace edn {
namespace {

using namespace tc;

constexpr int kRows = 128;
constexpr int kGroupThreads = 256;                // two threads per row
constexpr int kRowWarps = 16;
constexpr int kThreads = kRowWarps * 32 + 64;     // + 2 MMA issuer warps
constexpr int kAChunks = 24;                      // per group; full: operands 0..11, gather tile 12..23; lean: tile 0..11, PE 12..19
constexpr int kABytes = kAChunks * kChunkA;       // 48 KB
constexpr uint32_t kTmemCols = 128;               // 2 groups x 64 columns
// resident weight blob (bytes)
//   full: basis 6x1K | sigma0 6x2K | sigma1 4x512 | color0(geo) 1x2K | color1 4x2K
//   lean: fold1 10x2K | fold2 4x2K | color1 4x2K
constexpr int kOffBasis = 0, kOffS0 = 6144, kOffS1 = 18432, kOffC0 = 20480, kOffC1 = 22528, kFullBytes = 30720;
constexpr int kOffL1 = kFullBytes, kOffL2 = kOffL1 + 20480, kOffL3 = kOffL2 + 8192, kWBytes = kOffL3 + 8192;
constexpr int kMaxRpt = 4;                        // rays per tile (Nc >= 32)

struct alignas(16) GroupMisc {
  float z[kRows];
  float sig[kRows];       // per-row alpha (after the head sums)
  float w[kRows];
  float rgb[kRows * 3];
  alignas(16) float bias[kMaxRpt][64];
  alignas(16) float headp[2][kRows][4];   // per column-half partial rgb (xyz) / sigma (w) heads
  float red[4][8];        // per-warp partial ray sums (parallel compositing)
  float wtot[4];          // per-warp products of (1 - alpha)
};
struct Misc {
  uint64_t bar_a[2], bar_acc[2], bar_w;
  GridDev grid;
  uint32_t tmem_base, pad[3];
  alignas(16) float wdir[kPeDir][64];   // color_net.0 rows 15..41 (view-direction part), fp32
  alignas(16) float b0[64];
  alignas(16) float b1[64];
  alignas(16) float wsig[64];           // sigma_net.1 row 0 (lean sigma head)
  alignas(16) float wrgb[64][4];
  GroupMisc grp[2];
};
constexpr int kSmemBytes = 2 * kABytes + kWBytes + (int)sizeof(Misc);
static_assert(offsetof(Misc, wdir) % 16 == 0 && offsetof(Misc, grp) % 16 == 0 && offsetof(GroupMisc, bias) % 16 == 0 &&
              offsetof(GroupMisc, headp) % 16 == 0, "alignment");

// NC (16 or 32) accumulator columns [col0, col0 + NC) -> (+bias) -> (ReLU) -> f[] ; optionally bf16 -> A chunks col0/8..
template <int NC>
__device__ __forceinline__ void epi_cols(uint32_t taddr, int col0, bool relu, const float* bias_s, uint8_t* a_row, float (&f)[NC], bool store) {
  if constexpr (NC == 32) {
    uint32_t v[32];
    tmem_ld32(taddr + col0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  } else {
    uint32_t v[16];
    tmem_ld16(taddr + col0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
  }
  if (bias_s) {
#pragma unroll
    for (int i = 0; i < NC; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + col0 + i);
      f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
    }
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < NC; ++i) f[i] = fmaxf(f[i], 0.f);
  }
  if (store) {
#pragma unroll
    for (int j = 0; j < NC / 8; ++j)
      st_shared_v4(a_row + (col0 / 8 + j) * kChunkA, pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                   pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
  }
}

__device__ __forceinline__ void rows_signal(uint64_t* bar_a) {
  fence_proxy_async_smem();
  tc_fence_before();
  mbar_arrive(bar_a);
}

__device__ __forceinline__ void issue_layer(uint32_t d_tmem, uint32_t a_base, uint32_t b_base, int ksteps, int n, uint32_t b_step_bytes) {
  const uint32_t idesc = make_idesc_bf16(128, n);
  for (int j = 0; j < ksteps; ++j)
    mma_bf16_ss(d_tmem, make_smem_desc(a_base + 2 * j * kChunkA, kChunkA, 128), make_smem_desc(b_base + j * b_step_bytes, (uint32_t)n * 16u, 128),
                idesc, j > 0);
}

template <typename T, bool LEAN>
__global__ void __launch_bounds__(kThreads, 1) coarse_fwd_tc_kernel(const CoarseArgs a, const uint8_t* __restrict__ blob) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;
  uint8_t* Wsm = smem + 2 * kABytes;
  Misc* m = reinterpret_cast<Misc*>(Wsm + kWBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.n_samples;
  const int rpt = kRows / S;                               // rays per tile (host guarantees 32 <= S <= 128)
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
    m->wsig[i] = __ldg(a.mlp.sigma1_t + i * 16);           // sigma_net.1 row 0 = column 0 of the transposed weight
#pragma unroll
    for (int j = 0; j < 4; ++j) m->wrgb[i][j] = __ldg(a.mlp.color2_t + i * 4 + j);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = m->tmem_base;

  if (warp >= kRowWarps) {
    // =================================== MMA issuer warp of group q (one thread) =======================================
    // (converged warp, one elected lane issues: uniform descriptors, no per-lane retry loop around UTCHMMA)
    const int q = warp - kRowWarps;
    if (n_my > 0) {
      if (q == 0 && elect_one()) { mbar_expect_tx(&m->bar_w, kWBytes); bulk_g2s(Wsm, blob, kWBytes, &m->bar_w); }
      __syncwarp();
      const uint32_t aq = smem_u32(As) + q * kABytes, wb = smem_u32(Wsm);
      const uint32_t d_tmem = tmem + q * 64;
      uint32_t pa = 0;
      mbar_wait(&m->bar_w, 0);
      auto stage = [&](uint32_t a_base, uint32_t b_off, int ksteps, int n, uint32_t b_step) {
        mbar_wait(&m->bar_a[q], pa); pa ^= 1; tc_fence_after();
        if (elect_one()) {
          issue_layer(d_tmem, a_base, wb + b_off, ksteps, n, b_step);
          mma_commit(&m->bar_acc[q]);
        }
        __syncwarp();
      };
      for (int64_t it = 0; it < n_my; ++it) {
        if (LEAN) {
          stage(aq, kOffL1, 10, 64, 2048);                  // [g 96 | PE 64] -> 64   (basis_mat folded in)
          stage(aq, kOffL2, 4, 64, 2048);                   // 64 -> 64               (sigma_net.1 geo columns folded into color_net.0)
          stage(aq, kOffL3, 4, 64, 2048);                   // color_net.1
        } else {
          stage(aq + 12 * kChunkA, kOffBasis, 6, 32, 1024); // basis_mat: gather tile (chunks 12..23) x [96 -> 32]
          stage(aq, kOffS0, 6, 64, 2048);                   // sigma_net.0: [ft 32 | PE 63 | 0] -> 64
          stage(aq, kOffS1, 4, 16, 512);                    // sigma_net.1: 64 -> 16 (columns 0..14 geo, 15 sigma)
          stage(aq, kOffC0, 1, 64, 2048);                   // color_net.0, geo part: 16 (15 + zero row) -> 64
          stage(aq, kOffC1, 4, 64, 2048);                   // color_net.1: 64 -> 64
        }
      }
    }
    __syncwarp();
  } else {
    // =================================== row warps: TWO threads per (ray in tile, sample) row ==========================
    const int q = warp >> 3, half = (warp >> 2) & 1, gwarp = warp & 3;
    const int r = gwarp * 32 + lane;
    uint8_t* Aq = As + q * kABytes;
    uint8_t* a_row = Aq + r * 16;
    GroupMisc* gm = &m->grp[q];
    const uint32_t taddr_row = tmem + ((uint32_t)(gwarp * 32) << 16) + q * 64;
    uint32_t pacc = 0;
    const int bar_id = 1 + q, bar_half = 3 + q;
    const int lr = min(r / S, rpt - 1), s = (r / S < rpt) ? r - lr * S : S - 1;   // padding rows replay the last sample
    const bool row_valid = r < rpt * S;
    const bool mask_near = !(a.flags & EDN_FLAG_TRAIN) && a.rmnearplane > 0.f;
    constexpr int kTileChunk = LEAN ? 0 : 12, kPeChunk = LEAN ? 12 : 4;
    for (int64_t it = 0; it < n_my; ++it) {
      const int64_t tile = 2 * ((int64_t)blockIdx.x + it * gridDim.x) + q;
      const int64_t ray_raw = tile * rpt + lr;
      const bool live = row_valid && ray_raw < a.n_rays;
      const int64_t ray = ray_raw < a.n_rays ? ray_raw : a.n_rays - 1;
      const float* rb = a.ray_batch + ray * 11;
      const float o[3] = {__ldg(rb + 0), __ldg(rb + 1), __ldg(rb + 2)};
      const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
      const float zv = place_sample(a, ray, s, __ldg(rb + 6), __ldg(rb + 7));
      if (half == 0) {
        gm->z[r] = zv;
        // ---- PE(pts) -> 64 A columns (chunks kPeChunk..+7); the last column is the zero pad of K = 63 -> 64 ---------------
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
          st_shared_v4(a_row + (kPeChunk + j) * kChunkA, pack_bf16x2(pe[8 * j], pe[8 * j + 1]), pack_bf16x2(pe[8 * j + 2], pe[8 * j + 3]),
                       pack_bf16x2(pe[8 * j + 4], pe[8 * j + 5]), pack_bf16x2(pe[8 * j + 6], pe[8 * j + 7]));
      } else {
        // ---- per-ray bias of color_net.0: b0 + W0[:, 15:42] . PE(viewdir), fp32 ------------------------------------------
#pragma unroll 1
        for (int idx = r; idx < rpt * 64; idx += kRows) {
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
      }
      named_bar_sync(bar_id, kGroupThreads);
      // ---- VM gather of the coarse grid -> 128 x 96 bf16 tile; half h handles point groups gi = 2h, 2h + 1 ----------------
      {
        const float* z_s = gm->z;
        gather_points<T>(m->grid, Aq, kTileChunk, gwarp, lane, 2 * half, 2 * half + 2, [&](int pt, float (&p)[3]) {
          const int plr = min(pt / S, rpt - 1);
          const int64_t pray = min(tile * rpt + plr, a.n_rays - 1);   // the point's own ray (rays of a tile are adjacent)
          const float* rbp = a.ray_batch + pray * 11;
          const float zp = z_s[pt];
#pragma unroll
          for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(__ldg(rbp + i), __fmul_rn(__ldg(rbp + 3 + i), zp));
        });
      }
      rows_signal(&m->bar_a[q]);
      float sig_raw = 0.f;
      if (LEAN) {
        // ---- [g | PE] -> 64 (basis_mat folded in), ReLU; sigma head = fp32 dot with sigma_net.1 row 0 (partial per half) ----
        mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
        {
          float f[32];
          epi_cols<32>(taddr_row, 32 * half, true, nullptr, a_row, f, true);
          float sg = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) sg = fmaf(f[i], m->wsig[32 * half + i], sg);
          gm->headp[half][r][3] = sg;
        }
        rows_signal(&m->bar_a[q]);
      } else {
        // ---- basis_mat -> ft (32) -> A columns 0..31 -----------------------------------------------------------------------
        mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
        { float f[16]; epi_cols<16>(taddr_row, 16 * half, false, nullptr, a_row, f, true); }
        rows_signal(&m->bar_a[q]);
        // ---- sigma_net.0 -> ReLU -> A columns 0..63 ---------------------------------------------------------------------------
        mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
        { float f[32]; epi_cols<32>(taddr_row, 32 * half, true, nullptr, a_row, f, true); }
        rows_signal(&m->bar_a[q]);
        // ---- sigma_net.1 -> geo (cols 0..14) + sigma (col 15); A columns 0..15 (col 15 meets a zero weight row) -----------------
        mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
        if (half == 0) {
          float g16[16];
          epi_cols<16>(taddr_row, 0, false, nullptr, a_row, g16, true);
          sig_raw = g16[15];
          if (a.feat && live) {
            float* fo = a.feat + (ray * S + s) * 15;
#pragma unroll
            for (int j = 0; j < 15; ++j) fo[j] = g16[j];
          }
        }
        rows_signal(&m->bar_a[q]);
      }
      // ---- color_net.0 (+ per-ray view-dir bias) -> ReLU -> A columns 0..63 ------------------------------------------------------
      mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
      { float f[32]; epi_cols<32>(taddr_row, 32 * half, true, gm->bias[lr], a_row, f, true); }
      rows_signal(&m->bar_a[q]);
      // ---- color_net.1 -> ReLU -> rgb head (fp32, partial per half) ----------------------------------------------------------------
      mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
      {
        float f[32];
        epi_cols<32>(taddr_row, 32 * half, true, m->b1, a_row, f, false);
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 w = *reinterpret_cast<const float4*>(&m->wrgb[32 * half + i][0]);
          c0 = fmaf(f[i], w.x, c0); c1 = fmaf(f[i], w.y, c1); c2 = fmaf(f[i], w.z, c2);
        }
        gm->headp[half][r][0] = c0; gm->headp[half][r][1] = c1; gm->headp[half][r][2] = c2;
        if (!LEAN) gm->headp[half][r][3] = (half == 0) ? sig_raw : 0.f;
      }
      tc_fence_before();
      named_bar_sync(bar_id, kGroupThreads);
      if (half == 0) {
        {  // ---- compositing (voxnerf.py:153-201): alpha per row in parallel, then one thread per ray in sample order ---------
          const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
          const bool is_last = (s == S - 1);
          const float z_next = gm->z[min(r + 1, kRows - 1)];
          const float nz = (a.noise && !is_last) ? __ldg(a.noise + ray * (S - 1) + s) : 0.f;
          const float sg = gm->headp[0][r][3] + gm->headp[1][r][3];
          gm->sig[r] = alpha_of_sample(sg, zv, z_next, nz, dnorm, mask_near, a.rmnearplane / 128.0f, is_last);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            gm->rgb[3 * r + i] = sigmoidf_(gm->headp[0][r][i] + gm->headp[1][r][i] + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + i) : 0.f));
        }
        if ((S & 31) == 0) {
          // rays are whole warps: transmittance by a warp-shuffle product scan + the per-warp totals of the ray's earlier warps
          const float alpha = (r < rpt * S) ? gm->sig[r] : 0.f;
          float t = 1.0f - alpha;
#pragma unroll
          for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const float y = __shfl_up_sync(0xffffffffu, t, dlt);
            if (lane >= dlt) t *= y;
          }
          float Tr = __shfl_up_sync(0xffffffffu, t, 1);
          if (lane == 0) Tr = 1.0f;
          if (lane == 31) gm->wtot[gwarp] = t;
          named_bar_sync(bar_half, kRows);
          const int wpr = S >> 5, first = (gwarp / wpr) * wpr;       // warps per ray, first warp of this row's ray
          for (int w2 = first; w2 < gwarp; ++w2) Tr *= gm->wtot[w2];
          const float wgt = alpha * Tr;
          gm->w[r] = wgt;
          float cr = gm->rgb[3 * r], cg = gm->rgb[3 * r + 1], cb = gm->rgb[3 * r + 2];
          if (a.flags & EDN_FLAG_RELU_RGB) { cr = fmaxf(cr, 0.f); cg = fmaxf(cg, 0.f); cb = fmaxf(cb, 0.f); }
          float red[5] = {wgt * cr, wgt * cg, wgt * cb, wgt * zv, wgt};
#pragma unroll
          for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) red[i] += __shfl_xor_sync(0xffffffffu, red[i], off);
          }
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 5; ++i) gm->red[gwarp][i] = red[i];
          }
          named_bar_sync(bar_half, kRows);
          if (r < rpt * 5) {
            const int rr = r / 5, i = r % 5;
            const int64_t r2 = tile * rpt + rr;
            if (r2 < a.n_rays) {
              float sum = 0.f;
              for (int w2 = rr * wpr; w2 < (rr + 1) * wpr; ++w2) sum += gm->red[w2][i];
              if (i < 3) a.rgb[r2 * 3 + i] = sum; else if (i == 3) a.depth[r2] = sum; else a.acc[r2] = sum;
            }
          }
        } else {
          named_bar_sync(bar_half, kRows);
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
          named_bar_sync(bar_half, kRows);
        }
        if (live) {
          a.z_vals[ray * S + s] = zv;
          a.weights[ray * S + s] = gm->w[r];
        }
      }
      named_bar_sync(bar_id, kGroupThreads);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kRowWarps) tmem_dealloc(tmem, kTmemCols);
}

// ======================================================================================================================================
// Second-generation coarse kernel (lean schedule, S a multiple of 32): the VM gather is DECOUPLED from the MLP chain.
//   warps 0-15  producer (setmaxnreg 88): sample placement + PE + per-ray view-direction bias + VM gather of the NEXT tile of either
//               group straight into that group's A tile -- as soon as the tile's first MMA stage has retired (a_free), i.e. while the
//               group's epilogue warps are still two stages and a compositing away from finishing the previous tile;
//   warps 16-23 epilogue (64 registers): two groups x 4 warps, ONE thread per row: TMEM -> ReLU -> bf16 -> the group's 16 KB
//               activation buffer (next stage's A operand), fp32 sigma / rgb heads, warp-shuffle compositing;
//   warps 24-25 MMA issuers (one per group, converged warp + elect.sync), warps 26-27 idle donors (24 registers).
// TMEM: per group TWO 64-column accumulators (tile parity), so stage 1 of the next tile runs under the last epilogue of this one.
// Round 1's kernel kept every row warp on gather -> wait -> epilogue -> wait ...; its top stall was the long scoreboard of the gather.
// Bounded wait until a shared-memory tile counter reaches `need` (a protocol bug traps instead of hanging).  Used where a 1-bit
// mbarrier phase parity would be ambiguous: the epilogue may finish tile k - 1 before the producer asks for tile k - 2 (a slow producer,
// e.g. perturbed sample placement), and an mbarrier two phases ahead looks exactly like one that has not completed.
__device__ __forceinline__ void wait_tile_count(const uint32_t* ctr, uint32_t need) {
  uint32_t spins = 0;
  while (*reinterpret_cast<const volatile uint32_t*>(ctr) < need) {
    if (++spins > (1u << 24)) __trap();
  }
  __threadfence_block();
}

constexpr int kV2GatherWarps = 16, kV2EpiWarps = 8;
constexpr int kV2WarpMma = kV2GatherWarps + kV2EpiWarps;
constexpr int kV2Threads = (kV2GatherWarps + kV2EpiWarps + 4) * 32;     // 896
constexpr int kV2ProdThreads = kV2GatherWarps * 32;
constexpr int kV2ABytes = 20 * kChunkA;              // [g 96 | PE 64] = 20 chunks of 8 columns
constexpr int kV2ActBytes = 8 * kChunkA;             // 64 hidden columns
constexpr int kV2WBytes = kWBytes - kOffL1;          // the lean section of the blob
constexpr uint32_t kV2TmemCols = 256;                // 2 groups x 2 tile parities x 64 columns

struct alignas(16) GroupMisc2 {
  float z[2][kRows];                      // tile parity
  alignas(16) float bias[2][kMaxRpt][64];
  float sig[kRows];
  float w[kRows];
  float rgb[kRows * 3];
  float red[4][8];
  float wtot[4];
};
struct Misc2 {
  uint64_t a_full[2], a_free[2], act_full[2], acc[2][2], bar_w;      // acc[group][tile parity]
  GridDev grid;
  uint32_t tmem_base, tiles_done[2], pad[1];                            // tiles_done[group]: tiles whose epilogue has finished
  alignas(16) float wdir[kPeDir][64];
  alignas(16) float b0[64];
  alignas(16) float b1[64];
  alignas(16) float wsig[64];
  alignas(16) float wrgb[64][4];
  GroupMisc2 grp[2];
};
constexpr int kV2SmemBytes = 2 * kV2ABytes + 2 * kV2ActBytes + kV2WBytes + (int)sizeof(Misc2);
static_assert(kV2SmemBytes <= 232448, "shared memory budget");
static_assert(offsetof(Misc2, wdir) % 16 == 0 && offsetof(Misc2, grp) % 16 == 0 && offsetof(GroupMisc2, bias) % 16 == 0, "alignment");

template <typename T>
__global__ void __launch_bounds__(kV2Threads, 1) coarse_fwd_tc2_kernel(const CoarseArgs a, const uint8_t* __restrict__ blob, const int ablate, long long* __restrict__ trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;                                   // [2 groups][40 KB]
  uint8_t* Act = smem + 2 * kV2ABytes;                  // [2 groups][16 KB]
  uint8_t* Wsm = Act + 2 * kV2ActBytes;
  Misc2* m = reinterpret_cast<Misc2*>(Wsm + kV2WBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.n_samples;
  const int rpt = kRows / S;
  const int64_t n_tiles = (a.n_rays + rpt - 1) / rpt;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const int64_t n_my = (n_pairs > (int64_t)blockIdx.x) ? (n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int q = 0; q < 2; ++q) {
      mbar_init(&m->a_full[q], kV2ProdThreads); mbar_init(&m->a_free[q], 1); mbar_init(&m->act_full[q], kRows);
      mbar_init(&m->acc[q][0], 1); mbar_init(&m->acc[q][1], 1);
    }
    mbar_init(&m->bar_w, 1);
    fence_barrier_init();
    m->grid = a.grid;
    m->tiles_done[0] = 0; m->tiles_done[1] = 0;
  }
  if (warp == kV2WarpMma) tmem_alloc(&m->tmem_base, kV2TmemCols);
  if (ablate) {      // dev ablations leave parts of the operand tiles unwritten: start from zeros
    for (int i = tid; i < 2 * kV2ABytes / 16; i += kV2Threads) st_shared_v4(As + i * 16, 0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  for (int i = tid; i < kPeDir * 64; i += kV2Threads) m->wdir[i / 64][i % 64] = __ldg(a.mlp.color0_t + (15 + i / 64) * 64 + i % 64);
  for (int i = tid; i < 64; i += kV2Threads) {
    m->b0[i] = a.mlp.color0_b ? __ldg(a.mlp.color0_b + i) : 0.f;
    m->b1[i] = a.mlp.color1_b ? __ldg(a.mlp.color1_b + i) : 0.f;
    m->wsig[i] = __ldg(a.mlp.sigma1_t + i * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) m->wrgb[i][j] = __ldg(a.mlp.color2_t + i * 4 + j);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = m->tmem_base;

  if (warp >= kV2WarpMma) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    const int q = warp - kV2WarpMma;
    if (q < 2 && n_my > 0) {
      // =================================== MMA issuer of group q ==========================================================
      if (q == 0 && elect_one()) { mbar_expect_tx(&m->bar_w, kV2WBytes); bulk_g2s(Wsm, blob + kOffL1, kV2WBytes, &m->bar_w); }
      __syncwarp();
      const uint32_t aq = smem_u32(As) + q * kV2ABytes, actq = smem_u32(Act) + q * kV2ActBytes, wb = smem_u32(Wsm);
      mbar_wait(&m->bar_w, 0);
      for (int64_t it = 0; it < n_my; ++it) {
        const uint32_t d_tmem = tmem + q * 128 + (uint32_t)(it & 1) * 64;
        mbar_wait(&m->a_full[q], (uint32_t)it & 1); tc_fence_after();
        if (elect_one()) {
          issue_layer(d_tmem, aq, wb, 10, 64, 2048);                       // [g 96 | PE 64] -> 64
          mma_commit(&m->acc[q][it & 1]);
          mma_commit(&m->a_free[q]);                                        // the producer may refill this group's A tile
        }
        __syncwarp();
        mbar_wait(&m->act_full[q], 0); tc_fence_after();
        if (elect_one()) { issue_layer(d_tmem, actq, wb + (kOffL2 - kOffL1), 4, 64, 2048); mma_commit(&m->acc[q][it & 1]); }
        __syncwarp();
        mbar_wait(&m->act_full[q], 1); tc_fence_after();
        if (elect_one()) { issue_layer(d_tmem, actq, wb + (kOffL3 - kOffL1), 4, 64, 2048); mma_commit(&m->acc[q][it & 1]); }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp < kV2GatherWarps) {
    // =================================== producer warps ======================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int gwarp = warp & 3, gi = warp >> 2;
    for (int64_t it = 0; it < n_my; ++it) {
      const int par = (int)(it & 1);
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        auto stamp = [&](int k) { if (trace && blockIdx.x == 0 && tid == 0 && it >= 8 && it < 12) trace[((it - 8) * 2 + q) * 8 + k] = clock64(); };
        stamp(0);
        if (it > 0) {      // one polling warp, the rest blocks in a named barrier
          if (warp == 0) {
            mbar_wait(&m->a_free[q], (uint32_t)(it - 1) & 1);
            if (it >= 2) wait_tile_count(&m->tiles_done[q], (uint32_t)(it - 1));      // tiles 0 .. it - 2 of this group are composited
          }
          named_bar_sync(5, kV2ProdThreads);
        }
        stamp(1);
        const int64_t tile = 2 * ((int64_t)blockIdx.x + it * gridDim.x) + q;
        uint8_t* Aq = As + q * kV2ABytes;
        GroupMisc2* gm = &m->grp[q];
        if (ablate & 2) {
        } else if (warp < 4) {
          // ---- sample depth + PE(pts) of row r -> A chunks 12..19 ------------------------------------------------------------
          const int r = warp * 32 + lane;
          const int lr = min(r / S, rpt - 1), sm = (r / S < rpt) ? r - lr * S : S - 1;
          const int64_t ray = min(tile * rpt + lr, a.n_rays - 1);
          const float* rb = a.ray_batch + ray * 11;
          const float zv = place_sample(a, ray, sm, __ldg(rb + 6), __ldg(rb + 7));
          gm->z[par][r] = zv;
          float pe[64];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            pe[i] = __fadd_rn(__ldg(rb + i), __fmul_rn(__ldg(rb + 3 + i), zv));
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
          uint8_t* a_row = Aq + r * 16;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(a_row + (12 + j) * kChunkA, pack_bf16x2(pe[8 * j], pe[8 * j + 1]), pack_bf16x2(pe[8 * j + 2], pe[8 * j + 3]),
                         pack_bf16x2(pe[8 * j + 4], pe[8 * j + 5]), pack_bf16x2(pe[8 * j + 6], pe[8 * j + 7]));
        } else if (warp < 8) {
          // ---- per-ray bias of color_net.0: b0 + W0[:, 15:42] . PE(viewdir), fp32 ---------------------------------------------
#pragma unroll 1
          for (int idx = (warp - 4) * 32 + lane; idx < rpt * 64; idx += kRows) {
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
            gm->bias[par][br][col] = b;
          }
        }
        stamp(2);
        // ---- VM gather: warp (gwarp, gi) handles points 32 gwarp + 8 gi .. + 7; the depth is recomputed (same function, same bits) ----
        if (!(ablate & 1)) gather_points<T>(m->grid, Aq, 0, gwarp, lane, gi, gi + 1, [&](int pt, float (&p)[3]) {
          const int plr = min(pt / S, rpt - 1), ps = (pt / S < rpt) ? pt - plr * S : S - 1;
          const int64_t pray = min(tile * rpt + plr, a.n_rays - 1);
          const float* rbp = a.ray_batch + pray * 11;
          const float zp = place_sample(a, pray, ps, __ldg(rbp + 6), __ldg(rbp + 7));
#pragma unroll
          for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(__ldg(rbp + i), __fmul_rn(__ldg(rbp + 3 + i), zp));
        });
        stamp(3);
        fence_proxy_async_smem();
        mbar_arrive(&m->a_full[q]);
      }
    }
  } else {
    // =================================== epilogue warps: group q, ONE thread per row ============================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    const int ew = warp - kV2GatherWarps;
    const int q = ew >> 2, gwarp = ew & 3;
    const int r = gwarp * 32 + lane;
    GroupMisc2* gm = &m->grp[q];
    uint8_t* act_row = Act + q * kV2ActBytes + r * 16;
    const int bar_id = 1 + q;
    const int lr = min(r / S, rpt - 1), s = (r / S < rpt) ? r - lr * S : S - 1;
    const bool row_valid = r < rpt * S;
    const bool mask_near = !(a.flags & EDN_FLAG_TRAIN) && a.rmnearplane > 0.f;
    uint32_t pbits = 0u;                   // bit p: phase parity of acc[q][p] (three completions per tile of parity p)
    int cur_par = 0;
    auto wait_acc = [&]() {
      if (gwarp == 0) mbar_wait(&m->acc[q][cur_par], (pbits >> cur_par) & 1u);
      pbits ^= 1u << cur_par;
      named_bar_sync(bar_id, kRows);
      tc_fence_after();
    };
    for (int64_t it = 0; it < n_my; ++it) {
      const int par = (int)(it & 1);
      cur_par = par;
      const int64_t tile = 2 * ((int64_t)blockIdx.x + it * gridDim.x) + q;
      const int64_t ray_raw = tile * rpt + lr;
      const bool live = row_valid && ray_raw < a.n_rays;
      const int64_t ray = ray_raw < a.n_rays ? ray_raw : a.n_rays - 1;
      const float* rb = a.ray_batch + ray * 11;
      const uint32_t taddr_row = tmem + ((uint32_t)(gwarp * 32) << 16) + q * 128 + (uint32_t)par * 64;
      // global operands of the compositing, loaded a whole MLP chain ahead of their use
      const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
      const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
      const bool is_last = (s == S - 1);
      const float nz = (a.noise && !is_last) ? __ldg(a.noise + ray * (S - 1) + s) : 0.f;
      auto estamp = [&](int k) { if (trace && blockIdx.x == 0 && r == 0 && it >= 8 && it < 12) trace[64 + ((it - 8) * 2 + q) * 8 + k] = clock64(); };
      estamp(0);
      // ---- stage 1: [g | PE] -> 64, ReLU; sigma head = fp32 dot with sigma_net.1 row 0 -----------------------------------------
      wait_acc();
      estamp(1);
      float sg = 0.f;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float f[32];
        epi_cols<32>(taddr_row, 32 * h, true, nullptr, act_row, f, true);
#pragma unroll
        for (int i = 0; i < 32; ++i) sg = fmaf(f[i], m->wsig[32 * h + i], sg);
      }
      rows_signal(&m->act_full[q]);
      estamp(2);
      // ---- stage 2: color_net.0 (+ per-ray view-direction bias), ReLU ----------------------------------------------------------------
      wait_acc();
      estamp(3);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float f[32];
        epi_cols<32>(taddr_row, 32 * h, true, gm->bias[par][lr], act_row, f, true);
      }
      rows_signal(&m->act_full[q]);
      estamp(4);
      // ---- stage 3: color_net.1, ReLU -> rgb head (fp32) ---------------------------------------------------------------------------
      wait_acc();
      estamp(5);
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float f[32];
        epi_cols<32>(taddr_row, 32 * h, true, m->b1, act_row, f, false);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 w = *reinterpret_cast<const float4*>(&m->wrgb[32 * h + i][0]);
          c0 = fmaf(f[i], w.x, c0); c1 = fmaf(f[i], w.y, c1); c2 = fmaf(f[i], w.z, c2);
        }
      }
      tc_fence_before();
      estamp(6);
      // ---- compositing (voxnerf.py:153-201): per-row alpha, warp-shuffle transmittance scan, per-ray sums -----------------------------
      const float zv = gm->z[par][r];
      const float z_next = gm->z[par][min(r + 1, kRows - 1)];
      const float alpha = row_valid ? alpha_of_sample(sg, zv, z_next, nz, dnorm, mask_near, a.rmnearplane / 128.0f, is_last) : 0.f;
      float cr = sigmoidf_(c0 + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + 0) : 0.f));
      float cg = sigmoidf_(c1 + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + 1) : 0.f));
      float cb = sigmoidf_(c2 + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + 2) : 0.f));
      float t = 1.0f - alpha;
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, t, dlt);
        if (lane >= dlt) t *= y;
      }
      float Tr = __shfl_up_sync(0xffffffffu, t, 1);
      if (lane == 0) Tr = 1.0f;
      if (lane == 31) gm->wtot[gwarp] = t;
      named_bar_sync(bar_id, kRows);
      const int wpr = S >> 5, first = (gwarp / wpr) * wpr;       // warps per ray, first warp of this row's ray
      for (int w2 = first; w2 < gwarp; ++w2) Tr *= gm->wtot[w2];
      const float wgt = alpha * Tr;
      if (a.flags & EDN_FLAG_RELU_RGB) { cr = fmaxf(cr, 0.f); cg = fmaxf(cg, 0.f); cb = fmaxf(cb, 0.f); }
      float red[5] = {wgt * cr, wgt * cg, wgt * cb, wgt * zv, wgt};
#pragma unroll
      for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) red[i] += __shfl_xor_sync(0xffffffffu, red[i], off);
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) gm->red[gwarp][i] = red[i];
      }
      named_bar_sync(bar_id, kRows);
      if (r < rpt * 5) {
        const int rr = r / 5, i = r % 5;
        const int64_t r2 = tile * rpt + rr;
        if (r2 < a.n_rays) {
          float sum = 0.f;
          for (int w2 = rr * wpr; w2 < (rr + 1) * wpr; ++w2) sum += gm->red[w2][i];
          if (i < 3) a.rgb[r2 * 3 + i] = sum; else if (i == 3) a.depth[r2] = sum; else a.acc[r2] = sum;
        }
      }
      if (live) {
        a.z_vals[ray * S + s] = zv;
        a.weights[ray * S + s] = wgt;
      }
      named_bar_sync(bar_id, kRows);          // red[] / wtot[] are rewritten by the next tile; every thread is done with z[par] / TMEM
      if (r == 0) { __threadfence_block(); atomicAdd(&m->tiles_done[q], 1u); }
      estamp(7);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kV2WarpMma) tmem_dealloc(tmem, kV2TmemCols);
}

// ======================================================================================================================================
// Tensor-core PARITY coarse kernel (precision EDN_TC32, lean schedule): fine_tc3.cu's "bf16 x 3" recipe on the coarse pass.  Every GEMM
// operand (gathered products, PE, activations AND weights) is split into bf16 hi + bf16 lo, every K-step issues three MMAs (hi.hi +
// lo.hi + hi.lo) into one fp32 TMEM accumulator; everything around the GEMMs is the fp32 SIMT kernel's arithmetic (fp32 grids and
// taps, sincosf positional encodings, fp32 heads, bit-exact sample placement).  One group per CTA (two A tiles + two activation
// buffers + 72 KB of split weights fill the shared memory), the v2 roles: 16 producer warps, 4 epilogue warps (one thread per row),
// 1 MMA warp; two 64-column TMEM accumulators by tile parity.
constexpr int kT3EpiWarps = 4;
constexpr int kT3WarpMma = kV2GatherWarps + kT3EpiWarps;                  // 20
constexpr int kT3Threads = (kV2GatherWarps + kT3EpiWarps + 4) * 32;       // 768 (a whole warpgroup for the MMA warp: setmaxnreg)
constexpr int kT3WBytes = 2 * kV2WBytes;             // [L1 hi | L1 lo | L2 hi | L2 lo | L3 hi | L3 lo]
constexpr int kT3OffL1 = 0, kT3OffL2 = 2 * 20480, kT3OffL3 = kT3OffL2 + 2 * 8192;
constexpr uint32_t kT3TmemCols = 128;

struct Misc3 {
  uint64_t a_full, a_free, act_full, acc[2], bar_w;                    // acc[tile parity]
  GridDev grid;
  uint32_t tmem_base, tiles_done, pad[2];
  alignas(16) float wdir[kPeDir][64];
  alignas(16) float b0[64];
  alignas(16) float b1[64];
  alignas(16) float wsig[64];
  alignas(16) float wrgb[64][4];
  GroupMisc2 grp;
};
constexpr int kT3SmemBytes = 2 * kV2ABytes + 2 * kV2ActBytes + kT3WBytes + (int)sizeof(Misc3);
static_assert(kT3SmemBytes <= 232448, "shared memory budget");
static_assert(offsetof(Misc3, wdir) % 16 == 0 && offsetof(Misc3, grp) % 16 == 0, "alignment");

// 32 accumulator columns -> (+bias) -> ReLU -> f[]; optionally the hi / lo split -> the two activation tiles
__device__ __forceinline__ void epi_cols_split(uint32_t taddr, int col0, const float* bias_s, uint8_t* hi_row, uint8_t* lo_row, float (&f)[32],
                                               bool store) {
  uint32_t v[32];
  tmem_ld32(taddr + col0, v);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  if (bias_s) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + col0 + i);
      f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
  if (store) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_bf16x2(f[8 * j + 2 * i], f[8 * j + 2 * i + 1], hi[i], lo[i]);
      st_shared_v4(hi_row + (col0 / 8 + j) * kChunkA, hi[0], hi[1], hi[2], hi[3]);
      st_shared_v4(lo_row + (col0 / 8 + j) * kChunkA, lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// hi.hi + lo.hi + hi.lo per K-step (N = 64)
__device__ __forceinline__ void issue_layer3(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int ksteps) {
  const uint32_t idesc = make_idesc_bf16(128, 64);
  for (int j = 0; j < ksteps; ++j) {
    const uint64_t ah = make_smem_desc(a_hi + 2 * j * kChunkA, kChunkA, 128), al = make_smem_desc(a_lo + 2 * j * kChunkA, kChunkA, 128);
    const uint64_t bh = make_smem_desc(b_hi + j * 2048, 64u * 16u, 128), bl = make_smem_desc(b_lo + j * 2048, 64u * 16u, 128);
    mma_bf16_ss(d_tmem, ah, bh, idesc, j > 0);
    mma_bf16_ss(d_tmem, al, bh, idesc, 1);
    mma_bf16_ss(d_tmem, ah, bl, idesc, 1);
  }
}

template <typename T>
__global__ void __launch_bounds__(kT3Threads, 1) coarse_fwd_tc3_kernel(const CoarseArgs a, const uint8_t* __restrict__ blob3) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* A_hi = smem;
  uint8_t* A_lo = smem + kV2ABytes;
  uint8_t* Act_hi = smem + 2 * kV2ABytes;
  uint8_t* Act_lo = Act_hi + kV2ActBytes;
  uint8_t* Wsm = Act_lo + kV2ActBytes;
  Misc3* m = reinterpret_cast<Misc3*>(Wsm + kT3WBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.n_samples;
  const int rpt = kRows / S;
  const int64_t n_tiles = (a.n_rays + rpt - 1) / rpt;
  const int64_t n_my = (n_tiles > (int64_t)blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    mbar_init(&m->a_full, kV2ProdThreads); mbar_init(&m->a_free, 1); mbar_init(&m->act_full, kRows);
    mbar_init(&m->acc[0], 1); mbar_init(&m->acc[1], 1); mbar_init(&m->bar_w, 1);
    fence_barrier_init();
    m->grid = a.grid;
    m->tiles_done = 0;
  }
  if (warp == kT3WarpMma) tmem_alloc(&m->tmem_base, kT3TmemCols);
  for (int i = tid; i < kPeDir * 64; i += kT3Threads) m->wdir[i / 64][i % 64] = __ldg(a.mlp.color0_t + (15 + i / 64) * 64 + i % 64);
  for (int i = tid; i < 64; i += kT3Threads) {
    m->b0[i] = a.mlp.color0_b ? __ldg(a.mlp.color0_b + i) : 0.f;
    m->b1[i] = a.mlp.color1_b ? __ldg(a.mlp.color1_b + i) : 0.f;
    m->wsig[i] = __ldg(a.mlp.sigma1_t + i * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) m->wrgb[i][j] = __ldg(a.mlp.color2_t + i * 4 + j);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = m->tmem_base;
  GroupMisc2* gm = &m->grp;

  if (warp >= kT3WarpMma) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == kT3WarpMma && n_my > 0) {
      // =================================== MMA issuer ==========================================================================
      if (elect_one()) { mbar_expect_tx(&m->bar_w, kT3WBytes); bulk_g2s(Wsm, blob3, kT3WBytes, &m->bar_w); }
      __syncwarp();
      const uint32_t ah = smem_u32(A_hi), al = smem_u32(A_lo), ch = smem_u32(Act_hi), cl = smem_u32(Act_lo), wb = smem_u32(Wsm);
      mbar_wait(&m->bar_w, 0);
      for (int64_t it = 0; it < n_my; ++it) {
        const uint32_t d_tmem = tmem + (uint32_t)(it & 1) * 64;
        mbar_wait(&m->a_full, (uint32_t)it & 1); tc_fence_after();
        if (elect_one()) {
          issue_layer3(d_tmem, ah, al, wb + kT3OffL1, wb + kT3OffL1 + 20480, 10);
          mma_commit(&m->acc[it & 1]);
          mma_commit(&m->a_free);
        }
        __syncwarp();
        mbar_wait(&m->act_full, 0); tc_fence_after();
        if (elect_one()) { issue_layer3(d_tmem, ch, cl, wb + kT3OffL2, wb + kT3OffL2 + 8192, 4); mma_commit(&m->acc[it & 1]); }
        __syncwarp();
        mbar_wait(&m->act_full, 1); tc_fence_after();
        if (elect_one()) { issue_layer3(d_tmem, ch, cl, wb + kT3OffL3, wb + kT3OffL3 + 8192, 4); mma_commit(&m->acc[it & 1]); }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp < kV2GatherWarps) {
    // =================================== producer warps ======================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");      // pool: 768 x 80 = 61440 >= 512 x 88 + 128 x 80 + 128 x 24
    const int gwarp = warp & 3, gi = warp >> 2;
    for (int64_t it = 0; it < n_my; ++it) {
      const int par = (int)(it & 1);
      if (it > 0) {
        if (warp == 0) {
          mbar_wait(&m->a_free, (uint32_t)(it - 1) & 1);
          if (it >= 2) wait_tile_count(&m->tiles_done, (uint32_t)(it - 1));
        }
        named_bar_sync(5, kV2ProdThreads);
      }
      const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
      if (warp < 4) {
        // ---- sample depth + PE(pts) of row r (the fp32 kernel's arithmetic: sincosf(p 2^f)) -> A chunks 12..19, hi / lo -----------
        const int r = warp * 32 + lane;
        const int lr = min(r / S, rpt - 1), sm = (r / S < rpt) ? r - lr * S : S - 1;
        const int64_t ray = min(tile * rpt + lr, a.n_rays - 1);
        const float* rb = a.ray_batch + ray * 11;
        const float zv = place_sample(a, ray, sm, __ldg(rb + 6), __ldg(rb + 7));
        gm->z[par][r] = zv;
        float pe[64];
#pragma unroll
        for (int i = 0; i < 3; ++i) pe[i] = __fadd_rn(__ldg(rb + i), __fmul_rn(__ldg(rb + 3 + i), zv));
#pragma unroll
        for (int f = 0; f < kPeFreqPts; ++f) {
          const float fr = (float)(1 << f);
#pragma unroll
          for (int i = 0; i < 3; ++i) sincosf(pe[i] * fr, &pe[3 + 6 * f + i], &pe[6 + 6 * f + i]);
        }
        pe[63] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split_bf16x2(pe[8 * j + 2 * i], pe[8 * j + 2 * i + 1], hi[i], lo[i]);
          st_shared_v4(A_hi + r * 16 + (12 + j) * kChunkA, hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(A_lo + r * 16 + (12 + j) * kChunkA, lo[0], lo[1], lo[2], lo[3]);
        }
      } else if (warp < 8) {
        // ---- per-ray bias of color_net.0: b0 + W0[:, 15:42] . PE(viewdir), the fp32 kernel's arithmetic ------------------------------
#pragma unroll 1
        for (int idx = (warp - 4) * 32 + lane; idx < rpt * 64; idx += kRows) {
          const int br = idx >> 6, col = idx & 63;
          const int64_t bray = min(tile * rpt + br, a.n_rays - 1);
          const float* rb2 = a.ray_batch + bray * 11;
          const float vd[3] = {__ldg(rb2 + 8), __ldg(rb2 + 9), __ldg(rb2 + 10)};
          float b = m->b0[col];
#pragma unroll
          for (int i = 0; i < 3; ++i) b = fmaf(m->wdir[i][col], vd[i], b);
#pragma unroll 1
          for (int f = 0; f < kPeFreqDir; ++f) {
            const float fr = (float)(1 << f);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              float sn, cs;
              sincosf(vd[i] * fr, &sn, &cs);
              b = fmaf(m->wdir[3 + 6 * f + i][col], sn, b);
              b = fmaf(m->wdir[6 + 6 * f + i][col], cs, b);
            }
          }
          gm->bias[par][br][col] = b;
        }
      }
      gather_points_split<T>(m->grid, A_hi, A_lo, 0, gwarp, lane, gi, gi + 1, [&](int pt, float (&p)[3]) {
        const int plr = min(pt / S, rpt - 1), ps = (pt / S < rpt) ? pt - plr * S : S - 1;
        const int64_t pray = min(tile * rpt + plr, a.n_rays - 1);
        const float* rbp = a.ray_batch + pray * 11;
        const float zp = place_sample(a, pray, ps, __ldg(rbp + 6), __ldg(rbp + 7));
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(__ldg(rbp + i), __fmul_rn(__ldg(rbp + 3 + i), zp));
      });
      fence_proxy_async_smem();
      mbar_arrive(&m->a_full);
    }
  } else {
    // =================================== epilogue warps: ONE thread per row =====================================================
    const int gwarp = warp - kV2GatherWarps;      // (stays at the launch allocation of 80 registers)
    const int r = gwarp * 32 + lane;
    uint8_t* hi_row = Act_hi + r * 16;
    uint8_t* lo_row = Act_lo + r * 16;
    uint32_t pbits = 0u;                   // bit p: phase parity of acc[p]
    int cur_par = 0;
    const int lr = min(r / S, rpt - 1), s = (r / S < rpt) ? r - lr * S : S - 1;
    const bool row_valid = r < rpt * S;
    const bool mask_near = !(a.flags & EDN_FLAG_TRAIN) && a.rmnearplane > 0.f;
    auto wait_acc = [&]() {
      if (gwarp == 0) mbar_wait(&m->acc[cur_par], (pbits >> cur_par) & 1u);
      pbits ^= 1u << cur_par;
      named_bar_sync(1, kRows);
      tc_fence_after();
    };
    for (int64_t it = 0; it < n_my; ++it) {
      const int par = (int)(it & 1);
      cur_par = par;
      const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
      const int64_t ray_raw = tile * rpt + lr;
      const bool live = row_valid && ray_raw < a.n_rays;
      const int64_t ray = ray_raw < a.n_rays ? ray_raw : a.n_rays - 1;
      const float* rb = a.ray_batch + ray * 11;
      const uint32_t taddr_row = tmem + ((uint32_t)(gwarp * 32) << 16) + (uint32_t)par * 64;
      const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
      const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
      const bool is_last = (s == S - 1);
      const float nz = (a.noise && !is_last) ? __ldg(a.noise + ray * (S - 1) + s) : 0.f;
      wait_acc();
      float sg = 0.f;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float f[32];
        epi_cols_split(taddr_row, 32 * h, nullptr, hi_row, lo_row, f, true);
#pragma unroll
        for (int i = 0; i < 32; ++i) sg = fmaf(f[i], m->wsig[32 * h + i], sg);
      }
      rows_signal(&m->act_full);
      wait_acc();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float f[32];
        epi_cols_split(taddr_row, 32 * h, gm->bias[par][lr], hi_row, lo_row, f, true);
      }
      rows_signal(&m->act_full);
      wait_acc();
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float f[32];
        epi_cols_split(taddr_row, 32 * h, m->b1, hi_row, lo_row, f, false);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 w = *reinterpret_cast<const float4*>(&m->wrgb[32 * h + i][0]);
          c0 = fmaf(f[i], w.x, c0); c1 = fmaf(f[i], w.y, c1); c2 = fmaf(f[i], w.z, c2);
        }
      }
      tc_fence_before();
      const float zv = gm->z[par][r];
      const float z_next = gm->z[par][min(r + 1, kRows - 1)];
      gm->sig[r] = row_valid ? alpha_of_sample(sg, zv, z_next, nz, dnorm, mask_near, a.rmnearplane / 128.0f, is_last) : 0.f;
      gm->rgb[3 * r + 0] = sigmoidf_(c0 + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + 0) : 0.f));
      gm->rgb[3 * r + 1] = sigmoidf_(c1 + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + 1) : 0.f));
      gm->rgb[3 * r + 2] = sigmoidf_(c2 + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + 2) : 0.f));
      named_bar_sync(1, kRows);
      // one thread per ray composites in sample order: the fp32 kernel's summation order (parity mode)
      if (r < rpt) {
        const int64_t r2 = tile * rpt + r;
        if (r2 < a.n_rays) {
          float out[5];
          composite_from_alpha(gm->sig + r * S, gm->rgb + 3 * r * S, gm->z[par] + r * S, S, (a.flags & EDN_FLAG_RELU_RGB) != 0, gm->w + r * S, out);
          a.rgb[r2 * 3 + 0] = out[0]; a.rgb[r2 * 3 + 1] = out[1]; a.rgb[r2 * 3 + 2] = out[2];
          a.depth[r2] = out[3];
          a.acc[r2] = out[4];
        }
      }
      named_bar_sync(1, kRows);
      if (live) {
        a.z_vals[ray * S + s] = zv;
        a.weights[ray * S + s] = gm->w[r];
      }
      named_bar_sync(1, kRows);
      if (r == 0) { __threadfence_block(); atomicAdd(&m->tiles_done, 1u); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kT3WarpMma) tmem_dealloc(tmem, kT3TmemCols);
}

// [K][N] fp32 layer -> bf16 hi (split = 0) or lo = bf16(w - hi) (split = 1) in the UMMA K-major layout of tc::pack_layer_kernel
__global__ void pack_layer_split_generic_kernel(const float* __restrict__ wt, int ld, int K, int N, int split, __nv_bfloat16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * N) return;
  const int k = i / N, n = i - k * N;
  const float w = wt[(size_t)k * ld + n];
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  dst[(size_t)(k / 16) * (N * 16) + ((k % 16) / 8) * (N * 8) + n * 8 + (k % 8)] = split ? __float2bfloat16_rn(w - __bfloat162float(hi)) : hi;
}

// C[M][N] = A[M][K] . B[K][N] (row-major fp32): weight folding at pack time only
__global__ void fold_matmul_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ Cm, int ldc,
                                   int M, int K, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int mi = i / N, ni = i - mi * N;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(A[(size_t)mi * lda + k], B[(size_t)k * ldb + ni], acc);
  Cm[(size_t)mi * ldc + ni] = acc;
}

template <typename T, bool LEAN>
int launch_variant(const CoarseArgs& a, const uint8_t* blob, unsigned gx, cudaStream_t st) {
  EDN_CUDA_OK(cudaFuncSetAttribute(coarse_fwd_tc_kernel<T, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  coarse_fwd_tc_kernel<T, LEAN><<<gx, kThreads, kSmemBytes, st>>>(a, blob);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

}  // namespace

int launch_coarse_tc3(const CoarseArgs& a, int grid_dtype, cudaStream_t st) {
  EDN_REQUIRE(a.n_samples >= 32 && a.n_samples <= kRows, "edn_render_coarse_fwd(tc32): n_samples must be in [32,128], got %d", a.n_samples);
  EDN_REQUIRE(a.mlp.tc_blob != nullptr, "edn_render_coarse_fwd(tc32): edn_field_mlp.tc_blob is NULL (call edn_pack_coarse_tc)");
  EDN_REQUIRE(a.feat == nullptr, "edn_render_coarse_fwd(tc32): feature_map is emitted by the fp32 path only");
  const int rpt = kRows / a.n_samples;
  const int64_t n_tiles = (a.n_rays + rpt - 1) / rpt;
  const unsigned gx = (unsigned)(n_tiles < (int64_t)num_sms() ? n_tiles : (int64_t)num_sms());
  const uint8_t* blob3 = reinterpret_cast<const uint8_t*>(a.mlp.tc_blob) + kWBytes;
  if (grid_dtype == EDN_BF16) {
    EDN_CUDA_OK(cudaFuncSetAttribute(coarse_fwd_tc3_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT3SmemBytes));
    coarse_fwd_tc3_kernel<__nv_bfloat16><<<gx, kT3Threads, kT3SmemBytes, st>>>(a, blob3);
  } else {
    EDN_CUDA_OK(cudaFuncSetAttribute(coarse_fwd_tc3_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT3SmemBytes));
    coarse_fwd_tc3_kernel<float><<<gx, kT3Threads, kT3SmemBytes, st>>>(a, blob3);
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

int launch_coarse_tc(const CoarseArgs& a, int grid_dtype, cudaStream_t st) {
  EDN_REQUIRE(a.n_samples >= 32 && a.n_samples <= kRows, "edn_render_coarse_fwd(bf16): n_samples must be in [32,128], got %d", a.n_samples);
  EDN_REQUIRE(a.mlp.tc_blob != nullptr, "edn_render_coarse_fwd(bf16): edn_field_mlp.tc_blob is NULL (call edn_pack_coarse_tc)");
  const int rpt = kRows / a.n_samples;
  const int64_t n_tiles = (a.n_rays + rpt - 1) / rpt, n_pairs = (n_tiles + 1) / 2;
  const unsigned gx = (unsigned)(n_pairs < (int64_t)num_sms() ? n_pairs : (int64_t)num_sms());
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(a.mlp.tc_blob);
  const bool lean = a.feat == nullptr;            // the geo feature_map only exists in the full schedule
  static const bool force_v1 = getenv("EDN_COARSE_V1") != nullptr;      // dev switch: round 1's coupled kernel
  static const int ablate = getenv("EDN_COARSE_ABLATE") ? atoi(getenv("EDN_COARSE_ABLATE")) : 0;   // dev timing ablations: 1 no gather, 2 no PE / bias
  if (lean && (a.n_samples & 31) == 0 && !force_v1) {
    static const bool want_trace = getenv("EDN_COARSE_TRACE") != nullptr;      // dev: clock64 time line of CTA 0, tiles 8..11 of each group
    long long* trace = nullptr;
    if (want_trace) { EDN_CUDA_OK(cudaMalloc(&trace, 128 * sizeof(long long))); EDN_CUDA_OK(cudaMemsetAsync(trace, 0, 128 * sizeof(long long), st)); }
    if (grid_dtype == EDN_BF16) {
      EDN_CUDA_OK(cudaFuncSetAttribute(coarse_fwd_tc2_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kV2SmemBytes));
      coarse_fwd_tc2_kernel<__nv_bfloat16><<<gx, kV2Threads, kV2SmemBytes, st>>>(a, blob, ablate, trace);
    } else {
      EDN_CUDA_OK(cudaFuncSetAttribute(coarse_fwd_tc2_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kV2SmemBytes));
      coarse_fwd_tc2_kernel<float><<<gx, kV2Threads, kV2SmemBytes, st>>>(a, blob, ablate, trace);
    }
    EDN_CUDA_OK(cudaGetLastError());
    if (trace) {
      long long h[128];
      EDN_CUDA_OK(cudaStreamSynchronize(st));
      EDN_CUDA_OK(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
      cudaFree(trace);
      const long long t0 = h[0];
      for (int role = 0; role < 2; ++role)
        for (int i = 0; i < 8; ++i) {
          printf("[coarse2 %s it=%d q=%d]", role ? "epi " : "prod", 8 + i / 2, i & 1);
          for (int k = 0; k < (role ? 8 : 4); ++k) printf(" s%d=%lld", k, h[role * 64 + i * 8 + k] - t0);
          printf("\n");
        }
    }
    return EDN_OK;
  }
  if (grid_dtype == EDN_BF16)
    return lean ? launch_variant<__nv_bfloat16, true>(a, blob, gx, st) : launch_variant<__nv_bfloat16, false>(a, blob, gx, st);
  return lean ? launch_variant<float, true>(a, blob, gx, st) : launch_variant<float, false>(a, blob, gx, st);
}

}  // namespace edn

extern "C" int64_t edn_coarse_tc_blob_bytes(void) { return edn::kWBytes + edn::kT3WBytes; }      // bf16 schedules + the bf16 x 3 (tc32) section
extern "C" int64_t edn_coarse_tc_pack_workspace_floats(void) { return 160 * 64 + 64 * 64; }

extern "C" int edn_pack_coarse_tc(const edn_field_mlp* mlp, const float* basis_t, float* workspace, void* blob, void* stream) {
  using namespace edn;
  EDN_REQUIRE(mlp && basis_t && blob && workspace, "edn_pack_coarse_tc: null pointer");
  EDN_REQUIRE(mlp->hidden == 64 && mlp->geo_feat == 15 && mlp->sigma0_t && mlp->sigma1_t && mlp->color0_t && mlp->color1_t &&
              mlp->color2_t, "edn_pack_coarse_tc: needs the coarse field (hidden=64, geo_feat=15)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* b = reinterpret_cast<uint8_t*>(blob);
  // folded layers of the lean schedule (fp32 products, rounded to bf16 once by the packer)
  float* f1 = workspace;              // [160][64]: rows 0..95 basis_t [96][32] . sigma0_t[0:32] [32][64]; rows 96..159 sigma0_t[32:96]
  float* f2 = workspace + 160 * 64;   // [64][64] = sigma1_t[:, 1:16] [64][15] . color0_t[0:15] [15][64]
  fold_matmul_kernel<<<(96 * 64 + 255) / 256, 256, 0, st>>>(basis_t, 32, mlp->sigma0_t, 64, f1, 64, 96, 32, 64);
  EDN_CUDA_OK(cudaMemcpyAsync(f1 + 96 * 64, mlp->sigma0_t + 32 * 64, sizeof(float) * 64 * 64, cudaMemcpyDeviceToDevice, st));
  fold_matmul_kernel<<<(64 * 64 + 255) / 256, 256, 0, st>>>(mlp->sigma1_t + 1, 16, mlp->color0_t, 64, f2, 64, 64, 15, 64);
  struct Src { const float* wt; int ld, kv, nv, K, N, rot, off; };
  const Src src[8] = {{basis_t, 32, 96, 32, 96, 32, 0, kOffBasis},
                      {mlp->sigma0_t, 64, 96, 64, 96, 64, 0, kOffS0},
                      {mlp->sigma1_t, 16, 64, 16, 64, 16, 1, kOffS1},      // output col j <- sigma_net.1 row (j+1)%16: geo first, sigma last
                      {mlp->color0_t, 64, 15, 64, 16, 64, 0, kOffC0},      // geo rows only; K row 15 (the sigma column) is zero
                      {mlp->color1_t, 64, 64, 64, 64, 64, 0, kOffC1},
                      {f1, 64, 160, 64, 160, 64, 0, kOffL1}, {f2, 64, 64, 64, 64, 64, 0, kOffL2}, {mlp->color1_t, 64, 64, 64, 64, 64, 0, kOffL3}};
  for (int L = 0; L < 8; ++L) {
    const int total = src[L].K * src[L].N;
    tc::pack_layer_kernel<<<(total + 255) / 256, 256, 0, st>>>(src[L].wt, src[L].ld, src[L].kv, src[L].nv, src[L].K, src[L].N, src[L].rot,
                                                               reinterpret_cast<__nv_bfloat16*>(b + src[L].off));
  }
  // bf16 x 3 section (precision EDN_TC32): the three lean layers as hi / lo pairs, right behind the bf16 schedules
  {
    const float* lw[3] = {f1, f2, mlp->color1_t};
    const int lk[3] = {160, 64, 64};
    const int loff[3] = {kT3OffL1, kT3OffL2, kT3OffL3};
    for (int L = 0; L < 3; ++L)
      for (int split = 0; split < 2; ++split)
        pack_layer_split_generic_kernel<<<(lk[L] * 64 + 255) / 256, 256, 0, st>>>(
            lw[L], 64, lk[L], 64, split, reinterpret_cast<__nv_bfloat16*>(b + kWBytes + loff[L] + split * lk[L] * 64 * 2));
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

// Fine pass of render_rays on tcgen05 / TMEM at fp32-grade accuracy: the "bf16 x 3" tensor-core PARITY mode (precision EDN_TC32).
// Replaces networks/renderer.py:190-217 + networks/pdrf/voxnerf.py:203-259,153-201 for the FVR field (lean schedule, see fine_tc.cu).
//
// Every GEMM operand x (activations AND weights) is split into two bf16 numbers, hi = bf16(x) and lo = bf16(x - hi), and every
// K-step issues THREE tcgen05.mma into the same fp32 TMEM accumulator: hi.hi + lo.hi + hi.lo (the dropped lo.lo term is ~2^-18
// relative).  Products of bf16 pairs are exact in the tensor core's fp32 accumulation, so each layer is accurate to ~2^-17 of
// sum |a||w| -- fp32 SIMT grade -- while the contraction stays on the tensor cores.  Everything around the GEMMs is the fp32
// parity path's arithmetic: fp32 VM planes, fp32 bilinear taps, sincosf positional encoding, fp32 heads / compositing.
// Structure = fine_tc2.cu without the overlap tricks (this mode is judged on accuracy; it runs ~3x the MMAs):
//   gather warps (16): PE + view bias + VM gather of both grids (work-stealing rounds) -> A_hi / A_lo (2 x 64 KB shared memory, single
//                     buffered: the gather of ray k+1 overlaps layers 2, 3 of ray k);
//   MMA warp:         layer 1 SS from shared memory, layers 2 / 3 TS from TMEM ([0,128) A_hi, [128,256) A_lo, [256,512) accumulator);
//   epilogue warps (8): TMEM -> fp32 bias / ReLU / sigma, rgb heads -> hi / lo split -> TMEM; compositing;
//   weight stream:    [layer][K-step pair][hi 16 KB | lo 16 KB] through a 4 x 16 KB ring (cp.async.bulk).
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "fine_args.cuh"
#include "tc_common.cuh"
#include "tc_rows.cuh"

namespace edn {
namespace {

using namespace tc;

constexpr int kRows = 128;
constexpr int kGatherWarps = 16, kEpiWarps = 8;
constexpr int kWarpMma = kGatherWarps + kEpiWarps, kWarpLoad = kWarpMma + 1;
constexpr int kThreads = (kGatherWarps + kEpiWarps + 4) * 32;      // 896 = 7 warpgroups (setmaxnreg: gather 88, epilogue 64, rest 24)
constexpr int kGatherThreads = kGatherWarps * 32;                   // 512
constexpr int kRoleThreads = 256;                                   // epilogue threads
constexpr int kNst = 4;
constexpr int kStageBytes = 16384;         // 2 K-steps x (256 rows x 16 x 2 B) of the hi OR the lo weights
constexpr int kKstepBytes = 8192;
constexpr int kABytes = 65536;
constexpr int kStagesPerRay = 3 * 8 * 2;   // layers x K-step pairs x (hi, lo)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColAhi = 0, kColAlo = 128, kColAcc = 256;

struct alignas(16) RaySlot {
  float z[kRows];
  alignas(16) float bias[256];
};
struct Misc {
  uint64_t a_full, a_empty, slot_free[2];
  uint64_t w_full[kNst], w_empty[kNst];
  uint64_t acc_full, act;
  GridDev grids[2];
  uint32_t tmem_base, round_ctr[2], pad[1];     // round_ctr[ray parity]: the one being reset is never the one being pulled from
  alignas(16) float wsig[256];
  alignas(16) float wrgb[3][256];
  alignas(16) float bias1[256];
  RaySlot slot[2];
  alignas(16) float headp[2][kRows][4];
  float red[4][8];
  float wtot[4];
};
constexpr int kSmemBytes = 2 * kABytes + kNst * kStageBytes + (int)sizeof(Misc);
static_assert(kSmemBytes <= 232448, "shared memory budget");

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) fine_fwd_tc3_kernel(const FineArgs a, const uint8_t* __restrict__ wblob) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* A_hi = smem;
  uint8_t* A_lo = smem + kABytes;
  uint8_t* Ws = smem + 2 * kABytes;
  Misc* m = reinterpret_cast<Misc*>(Ws + kNst * kStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&m->a_full, kGatherThreads); mbar_init(&m->a_empty, 1);
    for (int b = 0; b < 2; ++b) mbar_init(&m->slot_free[b], kRoleThreads);
    for (int s = 0; s < kNst; ++s) { mbar_init(&m->w_full[s], 1); mbar_init(&m->w_empty[s], 1); }
    mbar_init(&m->acc_full, 1); mbar_init(&m->act, kRoleThreads);
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc(&m->tmem_base, kTmemCols);
  for (int i = tid; i < 256; i += kThreads) {
    m->wsig[i] = __ldg(a.mlp.sigma1_v + i);
    m->bias1[i] = a.mlp.color1_b ? __ldg(a.mlp.color1_b + i) : 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) m->wrgb[j][i] = __ldg(a.mlp.color2_t + i * 4 + j);
  }
  if (tid == 32) { m->grids[0] = a.gc; m->grids[1] = a.gf; m->round_ctr[0] = 0; m->round_ctr[1] = 0; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = m->tmem_base;
  const int S = a.S;
  const int64_t n_my = (a.n_rays > (int64_t)blockIdx.x) ? (a.n_rays - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp > kWarpLoad) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");      // idle warps of the donor warpgroup
  } else if (warp == kWarpLoad) {
    // =================================== weight stream =====================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (n_my > 0) {
      const uint32_t total = (uint32_t)n_my * kStagesPerRay;
      for (uint32_t g = 0; g < total; ++g) {
        const int s = g % kNst;
        const uint32_t use = g / kNst;
        if (use > 0) mbar_wait(&m->w_empty[s], (use - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(&m->w_full[s], kStageBytes);
          bulk_g2s(Ws + s * kStageBytes, wblob + (size_t)(g % kStagesPerRay) * kStageBytes, kStageBytes, &m->w_full[s]);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp == kWarpMma) {
    // =================================== MMA issuer (converged warp, one elected lane issues) ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (n_my > 0) {
      const uint32_t idesc = make_idesc_bf16(128, 256);
      const uint32_t w_base = smem_u32(Ws), a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo);
      const uint32_t d_tmem = tmem + kColAcc;
      uint32_t g = 0, n_layer = 0;
      for (int64_t it = 0; it < n_my; ++it) {
#pragma unroll 1
        for (int L = 0; L < 3; ++L) {
          if (n_layer > 0) mbar_wait(&m->act, (n_layer - 1) & 1);      // accumulator drained (and, for L > 0, A_hi / A_lo rewritten)
          if (L == 0) mbar_wait(&m->a_full, (uint32_t)it & 1);
#pragma unroll 1
          for (int p = 0; p < 8; ++p) {
            const int s_hi = g % kNst, s_lo = (g + 1) % kNst;
            mbar_wait(&m->w_full[s_hi], (g / kNst) & 1);
            mbar_wait(&m->w_full[s_lo], ((g + 1) / kNst) & 1);
            tc_fence_after();
            const uint64_t bh0 = make_smem_desc(w_base + s_hi * kStageBytes, 256 * 16, 128);
            const uint64_t bl0 = make_smem_desc(w_base + s_lo * kStageBytes, 256 * 16, 128);
            if (elect_one()) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int ks = 2 * p + i;
                const uint64_t bh = bh0 + (uint64_t)((i * kKstepBytes) >> 4), bl = bl0 + (uint64_t)((i * kKstepBytes) >> 4);
                if (L == 0) {
                  const uint64_t ah = make_smem_desc(a_hi + ks * 2 * kChunkA, kChunkA, 128), al = make_smem_desc(a_lo + ks * 2 * kChunkA, kChunkA, 128);
                  mma_bf16_ss(d_tmem, ah, bh, idesc, ks > 0);      // hi . hi
                  mma_bf16_ss(d_tmem, al, bh, idesc, 1);           // lo . hi
                  mma_bf16_ss(d_tmem, ah, bl, idesc, 1);           // hi . lo
                } else {
                  mma_bf16_ts(d_tmem, tmem + kColAhi + ks * 8, bh, idesc, ks > 0);
                  mma_bf16_ts(d_tmem, tmem + kColAlo + ks * 8, bh, idesc, 1);
                  mma_bf16_ts(d_tmem, tmem + kColAhi + ks * 8, bl, idesc, 1);
                }
              }
              mma_commit(&m->w_empty[s_hi]);
              mma_commit(&m->w_empty[s_lo]);
            }
            __syncwarp();
            g += 2;
          }
          if (elect_one()) {
            mma_commit(&m->acc_full);
            if (L == 0) mma_commit(&m->a_empty);
          }
          __syncwarp();
          ++n_layer;
        }
      }
    }
    __syncwarp();
  } else if (warp < kGatherWarps) {
    // =================================== gather warps ==========================================================================
    // warps 0-3: depth + PE of the 128 rows, warps 4-7: the per-ray view-direction bias; then ALL 16 warps pull the 32 gather rounds
    // (grid x row group x 8-point group) of the ray from a shared counter (the depths come from global memory: no barrier needed).
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int r = (warp & 3) * 32 + lane;
    for (int64_t it = 0; it < n_my; ++it) {
      const int buf = (int)(it & 1);
      if (it > 0) {
        if (warp == 0) {
          mbar_wait(&m->a_empty, (uint32_t)(it - 1) & 1);
          if (it > 1) mbar_wait(&m->slot_free[buf], (uint32_t)((it >> 1) - 1) & 1);
          if (it > 1 && lane == 0) m->round_ctr[buf] = 0;      // last pulled from two rays ago: every warp is past it
        }
        named_bar_sync(5, kGatherThreads);
      }
      RaySlot* slot = &m->slot[buf];
      const int64_t ray = (int64_t)blockIdx.x + it * gridDim.x;
      const float* rb = a.ray_batch + ray * 11;
      const float o[3] = {__ldg(rb + 0), __ldg(rb + 1), __ldg(rb + 2)};
      const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
      if (warp < 4) {
        const float zv = a.z_vals[ray * S + min(r, S - 1)];
        slot->z[r] = zv;
        // PE(pts) (embedding.py:92-98; the fp32 path's arithmetic: sincosf(p * 2^f)) -> A chunks 24..31, hi / lo split
        float pe[64];
#pragma unroll
        for (int i = 0; i < 3; ++i) pe[i] = __fadd_rn(o[i], __fmul_rn(d[i], zv));
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
          st_shared_v4(A_hi + r * 16 + (24 + j) * kChunkA, hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(A_lo + r * 16 + (24 + j) * kChunkA, lo[0], lo[1], lo[2], lo[3]);
        }
      } else if (warp < 8) {
        // per-ray bias of color_net.0: b0 + W0[:, 128:155] . PE(viewdir), the fp32 path's arithmetic (fine_f32.cu)
        const float vd[3] = {__ldg(rb + 8), __ldg(rb + 9), __ldg(rb + 10)};
#pragma unroll 1
        for (int hc = 0; hc < 2; ++hc) {
          const int col = r + hc * kRows;
          float b = a.mlp.color0_b ? __ldg(a.mlp.color0_b + col) : 0.f;
          const float* w = a.mlp.color0_t + (size_t)128 * 256 + col;
#pragma unroll
          for (int i = 0; i < 3; ++i) b = fmaf(__ldg(w + i * 256), vd[i], b);
          for (int f = 0; f < kPeFreqDir; ++f) {
            const float fr = (float)(1 << f);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              float sn, cs;
              sincosf(vd[i] * fr, &sn, &cs);
              b = fmaf(__ldg(w + (3 + 6 * f + i) * 256), sn, b);
              b = fmaf(__ldg(w + (6 + 6 * f + i) * 256), cs, b);
            }
          }
          slot->bias[col] = b;
        }
      }
      {
        const float* z_g = a.z_vals + ray * S;
        for (;;) {
          int k = 0;
          if (lane == 0) k = (int)atomicAdd(&m->round_ctr[buf], 1u);
          k = __shfl_sync(0xffffffffu, k, 0);
          if (k >= 32) break;
          const int hk = k >> 4, gk = (k >> 2) & 3, gik = k & 3;
          gather_points_split<T>(m->grids[hk], A_hi, A_lo, hk ? 12 : 0, gk, lane, gik, gik + 1, [&](int pt, float (&p)[3]) {
            const float zv = __ldg(z_g + min(pt, S - 1));
#pragma unroll
            for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], zv));
          });
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&m->a_full);
    }
  } else {
    // =================================== epilogue warps: TWO threads per sample row ============================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    const int ew = warp - kGatherWarps;
    const int half = ew >> 2, gwarp = ew & 3;
    const int r = gwarp * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(gwarp * 32) << 16);
    const float4* s_wsig = reinterpret_cast<const float4*>(m->wsig);
    const float4* s_wrgb = reinterpret_cast<const float4*>(&m->wrgb[0][0]);
    const float4* s_bias1 = reinterpret_cast<const float4*>(m->bias1);
    const bool mask_near = !(a.flags & EDN_FLAG_TRAIN) && a.rmnearplane > 0.f;
    const float near_thr = a.rmnearplane / 128.0f;
    const bool has_bias1 = a.mlp.color1_b != nullptr;
    uint32_t n_use = 0;
    for (int64_t it = 0; it < n_my; ++it) {
      const int buf = (int)(it & 1);
      RaySlot* slot = &m->slot[buf];
      const int64_t ray = (int64_t)blockIdx.x + it * gridDim.x;
      const float4* s_bias = reinterpret_cast<const float4*>(slot->bias);
      float sig_part = 0.f, rr = 0.f, rg_ = 0.f, rbl = 0.f;
      // layer L and the column chunk are compile-time constants: straight-line epilogue code (this kernel's epilogue is serial with
      // its MMAs, so every instruction saved here is kernel time)
      auto layer = [&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        if (ew == 0) mbar_wait(&m->acc_full, n_use & 1);
        named_bar_sync(4, kRoleThreads);
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col0 = q * 64 + half * 32;
          uint32_t v[32];
          tmem_ld32(lane_base + kColAcc + col0, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (L == 1 || (L == 2 && has_bias1)) {
            const float4* bs = (L == 1) ? s_bias : s_bias1;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = bs[(col0 + i) >> 2];
              f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          if (L == 0) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 w = s_wsig[(col0 + i) >> 2];
              s0 = fmaf(f[i], w.x, s0); s1 = fmaf(f[i + 1], w.y, s1); s0 = fmaf(f[i + 2], w.z, s0); s1 = fmaf(f[i + 3], w.w, s1);
            }
            sig_part += s0 + s1;
          }
          if (L < 2) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) split_bf16x2(f[2 * i], f[2 * i + 1], hi[i], lo[i]);
            tmem_st16(lane_base + kColAhi + (col0 >> 1), hi);
            tmem_st16(lane_base + kColAlo + (col0 >> 1), lo);
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 wr = s_wrgb[(col0 + i) >> 2], wg = s_wrgb[(256 + col0 + i) >> 2], wb = s_wrgb[(512 + col0 + i) >> 2];
              rr = fmaf(f[i], wr.x, rr); rr = fmaf(f[i + 1], wr.y, rr); rr = fmaf(f[i + 2], wr.z, rr); rr = fmaf(f[i + 3], wr.w, rr);
              rg_ = fmaf(f[i], wg.x, rg_); rg_ = fmaf(f[i + 1], wg.y, rg_); rg_ = fmaf(f[i + 2], wg.z, rg_); rg_ = fmaf(f[i + 3], wg.w, rg_);
              rbl = fmaf(f[i], wb.x, rbl); rbl = fmaf(f[i + 1], wb.y, rbl); rbl = fmaf(f[i + 2], wb.z, rbl); rbl = fmaf(f[i + 3], wb.w, rbl);
            }
          }
        }
        if (L < 2) tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&m->act);
        ++n_use;
      };
      layer(std::integral_constant<int, 0>{});
      layer(std::integral_constant<int, 1>{});
      layer(std::integral_constant<int, 2>{});
      m->headp[half][r][0] = rr; m->headp[half][r][1] = rg_; m->headp[half][r][2] = rbl; m->headp[half][r][3] = sig_part;
      named_bar_sync(2, kRoleThreads);
      if (half == 0) {
        // ---- compositing (voxnerf.py:153-201), one thread per sample row ---------------------------------------------------
        const float* rb = a.ray_batch + ray * 11;
        const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
        const float zv = slot->z[r];
        float col[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          col[i] = sigmoidf_(m->headp[0][r][i] + m->headp[1][r][i] + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + i) : 0.f));
        const float sig_raw = m->headp[0][r][3] + m->headp[1][r][3];
        float alpha = 0.f;
        if (r < S - 1) {
          const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
          const float znext = slot->z[r + 1];
          const float dist = __fmul_rn(znext - zv, dnorm);
          float sg = sig_raw;
          if (a.noise) sg += __ldg(a.noise + ray * (S - 1) + r);
          sg = fmaxf(sg, 0.f);
          if (mask_near && !(znext > near_thr)) sg = 0.f;
          alpha = 1.0f - expf(-__fmul_rn(sg, dist));
        } else if (r == S - 1) {
          alpha = 1.0f;
        }
        float t = 1.0f - alpha;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
          const float y = __shfl_up_sync(0xffffffffu, t, dlt);
          if (lane >= dlt) t *= y;
        }
        float Tr = __shfl_up_sync(0xffffffffu, t, 1);
        if (lane == 0) Tr = 1.0f;
        if (lane == 31) m->wtot[gwarp] = t;
        named_bar_sync(3, kRows);
        for (int w2 = 0; w2 < gwarp; ++w2) Tr *= m->wtot[w2];
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
          for (int i = 0; i < 5; ++i) m->red[gwarp][i] = red[i];
        }
        named_bar_sync(3, kRows);
        if (r < 5) {
          const float tot = m->red[0][r] + m->red[1][r] + m->red[2][r] + m->red[3][r];
          if (r < 3) a.rgb[ray * 3 + r] = tot; else if (r == 3) a.depth[ray] = tot; else a.acc[ray] = tot;
        }
      }
      named_bar_sync(2, kRoleThreads);
      mbar_arrive(&m->slot_free[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) tmem_dealloc(tmem, kTmemCols);
}

// [K = 256][N = 256] fp32 layer -> bf16 hi (split = 0) or lo = bf16(w - hi) (split = 1) in the UMMA K-major layout of the lean stream:
// element (n, k) -> (k / 16) * 4096 + ((k % 16) / 8) * 2048 + n * 8 + k % 8
__global__ void pack_layer_split_kernel(const float* __restrict__ wt, int ld, int split, __nv_bfloat16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 256 * 256) return;
  const int k = i / 256, n = i - k * 256;
  const float w = wt[(size_t)k * ld + n];
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 v = split ? __float2bfloat16_rn(w - __bfloat162float(hi)) : hi;
  dst[(size_t)(k / 16) * 4096 + ((k % 16) / 8) * 2048 + n * 8 + (k % 8)] = v;
}

}  // namespace

// Section appended to the fine tensor-core blob: [layer][K-step pair][hi 16 KB | lo 16 KB] of f1 | f23 | color1_t (768 KB)
int pack_fine_tc3_section(const float* f1, const float* f23, const float* color1_t, uint8_t* scratch, uint8_t* dst, cudaStream_t st) {
  const float* src[3] = {f1, f23, color1_t};
  for (int L = 0; L < 3; ++L)
    for (int split = 0; split < 2; ++split) {
      // pack the whole layer contiguously ([K-step][256 x 16], 128 KB) into scratch, then interleave 16 KB stages
      pack_layer_split_kernel<<<(256 * 256 + 255) / 256, 256, 0, st>>>(src[L], 256, split, reinterpret_cast<__nv_bfloat16*>(scratch));
      EDN_CUDA_OK(cudaMemcpy2DAsync(dst + (size_t)L * 262144 + (size_t)split * kStageBytes, 2 * kStageBytes, scratch, kStageBytes, kStageBytes, 8,
                                    cudaMemcpyDeviceToDevice, st));
    }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

int launch_fine_tc3(const FineArgs& a, int grid_dtype, const uint8_t* wblob, cudaStream_t st) {
  EDN_REQUIRE(a.S >= 2 && a.S <= kRows, "edn_render_fine_fwd(tc32): 2 <= n_samples <= 128, got %d", a.S);
  EDN_REQUIRE(a.feat == nullptr, "edn_render_fine_fwd(tc32): depth_feature is emitted by the fp32 path only");
  const unsigned gx = (unsigned)(a.n_rays < (int64_t)num_sms() ? a.n_rays : (int64_t)num_sms());
  if (grid_dtype == EDN_BF16) {
    EDN_CUDA_OK(cudaFuncSetAttribute(fine_fwd_tc3_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    fine_fwd_tc3_kernel<__nv_bfloat16><<<gx, kThreads, kSmemBytes, st>>>(a, wblob);
  } else {
    EDN_CUDA_OK(cudaFuncSetAttribute(fine_fwd_tc3_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    fine_fwd_tc3_kernel<float><<<gx, kThreads, kSmemBytes, st>>>(a, wblob);
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

}  // namespace edn

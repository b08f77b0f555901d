// Fine pass of render_rays on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands / fp32 accumulation.
// Replaces networks/renderer.py:190-217 + networks/pdrf/voxnerf.py:203-259,153-201 for the FVR field.
//
// One persistent CTA per SM renders TWO rays at a time (two row groups of 128 threads; one ray = 128 merged samples
// = one UMMA M tile) so that one ray's epilogue / gather overlaps the other ray's MMAs:
//   rows  (warps 0-15, TWO threads per sample row, group = warp / 8): PE -> A; cooperative VM gather (half 0: coarse grid,
//         half 1: fine grid) -> two 128x96 bf16 tiles;  after each layer: TMEM -> registers -> bias/ReLU -> bf16 -> A
//         operand of the next layer, each half of a row handling half of the columns; sigma (256->1) and rgb (256->3) are
//         fp32 dot products folded into the epilogues;  finally sigma->alpha compositing with a warp-shuffle scan.
//         The row code is instruction-latency bound per warp, hence 16 row warps (640 threads x 96 registers).
//   load  (warps 18 and 19, one thread each, one per ray): stream the layer weights (320 KB per ray, 16 KB stages, L2
//         resident) through the ray group's private 2-stage shared-memory ring with bulk async copies (TMA engine).
//         Private rings cost 2x the L2->smem weight traffic of a shared ring but let the two groups drift apart, so one
//         group's gather / epilogue overlaps the other's MMAs (a shared ring forces lock-step: measured in r1).
//   mma   (warps 16 and 17, one thread each, one per ray): wait for the ray's A operand and the ring stage, issue
//         tcgen05.mma: basis_mat x2 (N=32, resident B), sigma_net 128->256->128(geo), color_net 128(+view-dir
//         bias)->256->256, and commit to the stage-release / accumulator-ready mbarriers.  2 x 256 TMEM columns.
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "fine_args.cuh"
#include "tc_common.cuh"
#include "tc_rows.cuh"

#ifndef EDN_TC_ABLATE_BUILD
#define EDN_TC_ABLATE_BUILD 0      // dev builds only (tools/fine_tc_ablate.sh): compile the EDN_TC_ABLATE timing ablations in
#endif

namespace edn {
namespace {

using namespace tc;

constexpr int kRows = 128;               // sample rows of one ray tile = one UMMA M tile
constexpr int kGroupThreads = 256;       // TWO threads per sample row (halves h = 0 / 1 split the gather grids and the columns)
constexpr int kRowWarps = 16;            // 2 ray groups x 8 warps
constexpr int kThreads = kRowWarps * 32 + 128;  // + 2 MMA issuer warps + 2 weight-stream warps = 640
constexpr int kNst = 2;                  // weight ring stages PER RAY GROUP (private rings: the groups must not run in lock-step)
constexpr int kStageBytes = 16384;
constexpr int kABytes = 65536;           // per ray: 128 rows x 256 K bf16
constexpr int kBasisBytes = 2 * 6 * 1024;  // two basis_mat's, 6 K-steps x (N=32 x 16 x 2 B)
constexpr uint32_t kTmemCols = 512;
// Two layer schedules:
//   full (depth_feature requested, AWP): basis x2 | sigma0 128->256 | sigma1 256->128 (geo) | color0 128->256 | color1 256->256
//   lean (default): basis_mat folded into sigma_net.0 (K = 96+96+64 = 256) and sigma_net.1(geo) folded into color_net.0
//        (both are linear maps without an activation in between): 3 layers of 256->256, 3 sync stages instead of 5.
constexpr int kRingFull = 20, kRingLean = 24;            // 16 KB ring stages per ray
constexpr int kStepsFull = 2 + kRingFull, kStepsLean = kRingLean;
constexpr int64_t kOffFull = kBasisBytes, kOffLean = kOffFull + (int64_t)kRingFull * kStageBytes;
// pair schedule (cta_group::2, lean layers only): each CTA of a pair streams its N half of every layer: [layer][rank][64 KB],
// 4 stages of 4 K-steps (N/2 = 128 columns x 16 x 2 B = 4 KB each) per layer
constexpr int kRingPair = 12;
constexpr int64_t kOffPair = kOffLean + (int64_t)kRingLean * kStageBytes;
constexpr int64_t kOffTc3 = kOffPair + 3 * 131072;     // fine_tc3.cu: hi / lo split weights of the three lean layers (768 KB)
constexpr int64_t kBlobBytes = kOffTc3 + 3 * 262144;

struct StepDesc {
  uint8_t kind;      // 0 = basis coarse tile, 1 = basis fine tile, 2 = ring stage
  uint8_t n_mmas;
  uint16_t n;        // UMMA N
  uint16_t a_k0;     // first A K-step (16 columns each) of this stage
  uint8_t first;     // first step of a layer: wait for the rows' A operand
  uint8_t last;      // last step of a layer: commit the accumulator barrier
};
__constant__ StepDesc c_steps[kStepsFull];
__constant__ StepDesc c_steps_lean[kStepsLean];

std::vector<StepDesc> build_steps(bool lean) {
  std::vector<StepDesc> v;
  if (!lean) {
    v.push_back({0, 6, 32, 0, 1, 0});
    v.push_back({1, 6, 32, 0, 0, 1});
  }
  auto layer = [&](int K, int N) {
    const int kstep_bytes = N * 32, per_stage = kStageBytes / kstep_bytes, stages = (K / 16) / per_stage;
    for (int s = 0; s < stages; ++s)
      v.push_back({2, (uint8_t)per_stage, (uint16_t)N, (uint16_t)(s * per_stage), (uint8_t)(s == 0), (uint8_t)(s == stages - 1)});
  };
  if (lean) {
    layer(256, 256);   // [g_coarse | g_fine | PE] -> sigma_net.0 with basis_mat folded in
    layer(256, 256);   // color_net.0(geo part) o sigma_net.1(geo columns)
    layer(256, 256);   // color_net.1
  } else {
    layer(128, 256);   // sigma_net.0
    layer(256, 128);   // sigma_net.1 (geo columns)
    layer(128, 256);   // color_net.0 (geo part)
    layer(256, 256);   // color_net.1
  }
  return v;
}

struct alignas(16) GroupMisc {
  float tcarry, pad_[3];
  float z[kRows];
  alignas(16) float bias[256];
  alignas(16) float headp[2][kRows][4];   // per column-half partial sigma (x) / rgb (xyz) heads
  float red[4][8];
  float wtot[4];
};
struct Misc {
  uint64_t bar_a[2], bar_acc[2], full[2][kNst], empty[2][kNst], bar_basis;
  uint64_t peer_full[2][kNst];    // pair mode, rank 0: the peer CTA's ring stage has landed
  GridDev grids[2];          // [0] coarse, [1] fine
  uint32_t tmem_base, skew_flag, pad[2];
  alignas(16) float wsig[256];           // sigma_net.1 row 0 (sigma head)
  alignas(16) float wrgb[256][4];        // color_net.2 transposed (rgb head)
  alignas(16) float bias1[256];          // color_net.1 bias (zeros without --rgb_add_bias)
  GroupMisc grp[2];
};
static_assert(offsetof(Misc, wsig) % 16 == 0 && offsetof(Misc, wrgb) % 16 == 0 && offsetof(Misc, bias1) % 16 == 0 &&
              offsetof(Misc, grp) % 16 == 0 && offsetof(GroupMisc, bias) % 16 == 0, "float4 alignment");
constexpr int kSmemBytes = 2 * kABytes + 2 * kNst * kStageBytes + kBasisBytes + (int)sizeof(Misc);
static_assert(kSmemBytes <= 232448, "shared memory budget");

// Cooperative gather of ONE grid (fine_tile = 0 coarse / 1 fine) for the 32 points of this warp.  Lane (q = lane/8,
// j = lane%8) serves point 8*gi+j and reads 16-byte channel chunks, so every warp-wide load covers whole 32 B sectors of
// 8 texels.  Per gi iteration a lane runs 3 tasks (component 0 chunks q and q+4 with 12 loads in flight, then chunk q&1
// of component 1 + q/2).
template <typename T>
__device__ __forceinline__ void gather_tiles(const GridDev* grids_s, uint8_t* As, const float* z_s, int gwarp, int lane,
                                             const float o[3], const float d[3], const bool lean, const int fine_tile) {
  const int q = lane >> 3;
#pragma unroll 1
  for (int gi = 0; gi < 4; ++gi) {
    const GridDev& g = grids_s[fine_tile];
    const int pt = gwarp * 32 + gi * 8 + (lane & 7);
    const float zv = z_s[pt];
    float p[3], n[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], zv));
    normalize_pt(g, p, n);
    uint8_t* row = As + pt * 16;
    // A-buffer chunk of channel chunk cc: lean: coarse cc, fine 12 + cc; full: coarse 16 + cc, fine 28..31,0..7
    const int base = lean ? (fine_tile ? 12 : 0) : (fine_tile ? 28 : 16);
    {  // component 0: plane (x,y) 64 channels, line z: channel chunks q and q + 4, 12 loads in flight
      GatherTask<T> t0, t1;
      Taps2 pt2; Taps1 lt1;
      plane_taps(n[0], n[1], g.ph[0], g.pw[0], pt2);
      line_taps(n[2], g.ll[0], lt1);
      const T* pl = reinterpret_cast<const T*>(g.plane[0]);
      const T* ln = reinterpret_cast<const T*>(g.line[0]);
      t0.issue(pl, ln, 64, q, pt2, lt1);
      t1.issue(pl, ln, 64, q + 4, pt2, lt1);
      t0.finish(row + ((base + q) & 31) * kChunkA);
      t1.finish(row + ((base + 4 + q) & 31) * kChunkA);
    }
    {  // components 1 (plane (x,z), line y) and 2 (plane (y,z), line x): 16 channels each = 2 chunks each
      GatherTask<T> t2;
      const int comp = 1 + (q >> 1);
      Taps2 pt2; Taps1 lt1;
      plane_taps(comp == 1 ? n[0] : n[1], n[2], g.ph[comp], g.pw[comp], pt2);
      line_taps(comp == 1 ? n[1] : n[0], g.ll[comp], lt1);
      t2.issue(reinterpret_cast<const T*>(g.plane[comp]), reinterpret_cast<const T*>(g.line[comp]), 16, q & 1, pt2, lt1);
      t2.finish(row + ((base + 8 + q) & 31) * kChunkA);
    }
  }
}

enum EpiMode { kEpiPlain = 0, kEpiRelu = 1, kEpiReluSigma = 2, kEpiReluRgb = 3 };

// 32 accumulator columns of this thread's row:  x = acc (+ bias);  optional fp32 copy to global;  ReLU;  fp32 dot
// products with the sigma / rgb head;  bf16 -> A operand chunks (except kEpiReluRgb: only the rgb head is produced).
__device__ __forceinline__ void epilogue_block(const uint32_t (&v)[32], int col0, int mode, uint8_t* a_row,
                                               uint32_t bias_s, float* __restrict__ gout, uint32_t wsig, uint32_t wrgb,
                                               float4& head) {   // bias_s / wsig / wrgb: shared-space addresses (0 = none)
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  if (bias_s) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = ld_shared_f4(bias_s + (col0 + i) * 4);
      f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
    }
  }
  if (gout) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(gout + col0 + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
  }
  if (mode == kEpiRelu) {          // ReLU fused into the bf16 conversion
#pragma unroll
    for (int j = 0; j < 4; ++j)
      st_shared_v4(a_row + (col0 / 8 + j) * kChunkA, pack_relu_bf16x2(f[8 * j], f[8 * j + 1]), pack_relu_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                   pack_relu_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_relu_bf16x2(f[8 * j + 6], f[8 * j + 7]));
    return;
  }
  if (mode != kEpiPlain) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
  }
  if (mode == kEpiReluSigma) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 w = ld_shared_f4(wsig + (col0 + i) * 4);
      s0 = fmaf(f[i], w.x, s0); s1 = fmaf(f[i + 1], w.y, s1); s0 = fmaf(f[i + 2], w.z, s0); s1 = fmaf(f[i + 3], w.w, s1);
    }
    head.x += s0 + s1;
  }
  if (mode == kEpiReluRgb) {
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float4 w = ld_shared_f4(wrgb + (col0 + i) * 16);
      r0 = fmaf(f[i], w.x, r0); r1 = fmaf(f[i], w.y, r1); r2 = fmaf(f[i], w.z, r2);
    }
    head.x += r0; head.y += r1; head.z += r2;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      st_shared_v4(a_row + (col0 / 8 + j) * kChunkA, pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                   pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
  }
}

// Layer epilogue over accumulator columns [col_begin, col_begin + ncols) (multiples of 32) of this thread's row; ONE
// copy of the code for all layers (the kernel is instruction-cache sensitive).
__device__ __noinline__ float4 layer_epilogue(uint32_t taddr_row, uint8_t* a_row, int col_begin, int ncols, int mode, uint32_t bias_s,
                                              float* __restrict__ gout, uint32_t wsig, uint32_t wrgb) {
  float4 head = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int col0 = col_begin; col0 < col_begin + ncols; col0 += 32) {
    uint32_t v[32];
    tmem_ld32(taddr_row + col0, v);
    tmem_ld_wait();
    epilogue_block(v, col0, mode, a_row, bias_s, gout, wsig, wrgb, head);
  }
  return head;
}

__device__ __forceinline__ void rows_signal_a(uint64_t* bar_a) {
  fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();          // order our tcgen05.ld's before the MMAs that will overwrite those TMEM columns
  mbar_arrive(bar_a);
}
// pair mode: every row thread of BOTH CTAs arrives on the rank-0 CTA's barrier (its MMA issuer drives the pair)
__device__ __forceinline__ void rows_signal_a_pair(uint32_t bar_a_rank0) {
  asm volatile("fence.proxy.async;" ::: "memory");      // the peer CTA's tensor core reads this operand too: full async-proxy fence
  tc_fence_before();
  mbar_arrive_cluster(bar_a_rank0);
}

// PAIR: two CTAs of a cluster (one TPC) drive cta_group::2 MMAs: M = 256 = both CTAs' 128-row tiles of ray group q, every CTA
// streams only its N half of the weights (half the L2 -> smem weight traffic and twice the K extent per ring stage, which
// is what the single-CTA kernel is latency-bound on).  Lean schedule only.
template <typename T, bool LEAN, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) fine_fwd_tc_kernel(const FineArgs a, const uint8_t* __restrict__ blob) {
  static_assert(!PAIR || LEAN, "the pair variant implements the lean schedule only");
  constexpr int kRingStagesPerRay = PAIR ? kRingPair : (LEAN ? kRingLean : kRingFull);
  constexpr int kStepsPerRay = LEAN ? kStepsLean : kStepsFull;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;                                   // [2][64 KB]
  uint8_t* Ws = smem + 2 * kABytes;                     // ring
  uint8_t* Bs = Ws + 2 * kNst * kStageBytes;            // resident basis_mat operands
  Misc* m = reinterpret_cast<Misc*>(Bs + kBasisBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;

  if (tid == 0) {
    for (int q = 0; q < 2; ++q) { mbar_init(&m->bar_a[q], PAIR ? 2 * kGroupThreads : kGroupThreads); mbar_init(&m->bar_acc[q], 1); }
    for (int q = 0; q < 2; ++q)
      for (int s = 0; s < kNst; ++s) { mbar_init(&m->full[q][s], 1); mbar_init(&m->empty[q][s], 1); mbar_init(&m->peer_full[q][s], 1); }
    mbar_init(&m->bar_basis, 1);
    fence_barrier_init();
    m->skew_flag = 0;
  }
  if (warp == kRowWarps) { if (PAIR) tmem_alloc2(&m->tmem_base, kTmemCols); else tmem_alloc(&m->tmem_base, kTmemCols); }
  for (int i = tid; i < 256; i += kThreads) {
    m->wsig[i] = __ldg(a.mlp.sigma1_v + i);
    m->bias1[i] = a.mlp.color1_b ? __ldg(a.mlp.color1_b + i) : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) m->wrgb[i][j] = __ldg(a.mlp.color2_t + i * 4 + j);
  }
  if (tid == 32) { m->grids[0] = a.gc; m->grids[1] = a.gf; }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem = m->tmem_base;
  const int ablate_bits = EDN_TC_ABLATE_BUILD ? a.ablate : 0;     // compile-time gate: the shipped kernel carries no ablation branches
  const int64_t n_pairs_total = (a.n_rays + 1) / 2;
  // pair mode: both CTAs of a cluster run the same number of iterations (the rank-1 CTA may replay a masked duplicate ray)
  const int64_t first_cta = PAIR ? (int64_t)(blockIdx.x & ~1u) : (int64_t)blockIdx.x;
  const int64_t n_my = (n_pairs_total > first_cta) ? (n_pairs_total - first_cta + gridDim.x - 1) / gridDim.x : 0;
  const int S = a.S;
  const int tpr = (S + kRows - 1) / kRows;                   // 128-row tiles per ray (rays longer than 128 samples span several)
  const int64_t n_it = n_my * tpr;

  if (warp >= kRowWarps + 2) {
    // =================================== weight-stream producer warp of ray group q ==================================
    const int q = warp - (kRowWarps + 2);
    if (lane == 0 && n_my > 0) {
      const uint8_t* stream = blob + (PAIR ? kOffPair : (LEAN ? kOffLean : kOffFull));
      if (q == 0 && !LEAN) { mbar_expect_tx(&m->bar_basis, kBasisBytes); bulk_g2s(Bs, blob, kBasisBytes, &m->bar_basis); }
      uint8_t* ring = Ws + q * kNst * kStageBytes;
      const uint32_t total_ring = (uint32_t)n_it * kRingStagesPerRay;
      for (uint32_t g = 0; g < total_ring; ++g) {
        const int s = g % kNst;
        const uint32_t use = g / kNst;
        if (use > 0) mbar_wait(&m->empty[q][s], (use - 1) & 1);   // the MMAs on the previous tenant completed
        if (ablate_bits & 4) { mbar_expect_tx(&m->full[q][s], 0); continue; }
        mbar_expect_tx(&m->full[q][s], kStageBytes);
        const uint32_t gs = g % kRingStagesPerRay;
        const size_t src = PAIR ? (size_t)(gs >> 2) * 131072 + (size_t)rank * 65536 + (size_t)(gs & 3) * kStageBytes : (size_t)gs * kStageBytes;
        bulk_g2s(ring + s * kStageBytes, stream + src, kStageBytes, &m->full[q][s]);
      }
    }
    __syncwarp();
  } else if (warp >= kRowWarps) {
    // =================================== MMA issuer warp of ray group q (one thread) ===================================
    const int q = warp - kRowWarps;
    if (PAIR && lane == 0 && n_my > 0) {
      const uint32_t total_ring = (uint32_t)n_it * kRingPair;
      if (rank == 1) {
        // ---- peer CTA: forward "my ring stage has landed" to the pair's MMA issuer (rank 0) ------------------------------
        for (uint32_t g = 0; g < total_ring; ++g) {
          const int s = g % kNst;
          mbar_wait(&m->full[q][s], (g / kNst) & 1);
          mbar_arrive_cluster(mapa_u32(smem_u32(&m->peer_full[q][s]), 0));
        }
      } else {
        // ---- rank 0: issue the pair's MMAs: M = 256 (128 rows per CTA), N = 256 (128 columns of B per CTA), K = 16 ----------
        const uint32_t aq = smem_u32(As) + q * kABytes, w_base = smem_u32(Ws) + q * kNst * kStageBytes;
        const uint32_t d_tmem = tmem + q * 256;
        const uint32_t idesc = make_idesc_bf16(256, 256);
        uint32_t pa = 0;
        for (uint32_t g = 0; g < total_ring; ++g) {
          const int s = g % kNst;
          const uint32_t st4 = g & 3;                            // stage within the layer (4 stages of 4 K-steps)
          mbar_wait(&m->full[q][s], (g / kNst) & 1);
          mbar_wait_cluster(&m->peer_full[q][s], (g / kNst) & 1);
          if (st4 == 0) { mbar_wait_cluster(&m->bar_a[q], pa); pa ^= 1; }
          tc_fence_after();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t kst = st4 * 4 + i;
            const uint64_t adesc = make_smem_desc(aq + kst * 2 * kChunkA, kChunkA, 128);
            const uint64_t bdesc = make_smem_desc(w_base + s * kStageBytes + i * 4096, 128 * 16, 128);
            mma_bf16_ss2(d_tmem, adesc, bdesc, idesc, kst > 0);
          }
          mma_commit2(&m->empty[q][s]);
          if (st4 == 3) mma_commit2(&m->bar_acc[q]);
        }
      }
    } else if (!PAIR && lane == 0 && n_my > 0) {
      const uint32_t aq = smem_u32(As) + q * kABytes, w_base = smem_u32(Ws) + q * kNst * kStageBytes, b_base = smem_u32(Bs);
      const uint32_t d_tmem = tmem + q * 256;
      const int fine_tile_chunks[6] = {28, 30, 0, 2, 4, 6};
      uint32_t pa = 0, g = 0;
      if (!LEAN) mbar_wait(&m->bar_basis, 0);
      for (int64_t it = 0; it < n_it; ++it) {
#pragma unroll 1
        for (int step = 0; step < kStepsPerRay; ++step) {
          const StepDesc sd = LEAN ? c_steps_lean[step] : c_steps[step];
          if (sd.kind == 2) {
            const int s = g % kNst;
            mbar_wait(&m->full[q][s], (g / kNst) & 1);       // weights first: they landed long before the rows' A operand
            if (sd.first) { mbar_wait(&m->bar_a[q], pa); pa ^= 1; }
            tc_fence_after();
            const uint32_t idesc = make_idesc_bf16(128, sd.n);
            const uint32_t kstep_bytes = (uint32_t)sd.n * 32u;
            for (int i = 0; i < sd.n_mmas; ++i) {
              const uint32_t kst = sd.a_k0 + i;
              const uint64_t adesc = make_smem_desc(aq + kst * 2 * kChunkA, kChunkA, 128);
              const uint64_t bdesc = make_smem_desc(w_base + s * kStageBytes + i * kstep_bytes, (uint32_t)sd.n * 16u, 128);
              mma_bf16_ss(d_tmem, adesc, bdesc, idesc, kst > 0);
            }
            mma_commit(&m->empty[q][s]);
            ++g;
          } else {
            if (sd.first) { mbar_wait(&m->bar_a[q], pa); pa ^= 1; }
            tc_fence_after();
            const uint32_t idesc = make_idesc_bf16(128, 32);
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const int chunk = sd.kind == 0 ? 16 + 2 * j : fine_tile_chunks[j];
              const uint64_t adesc = make_smem_desc(aq + chunk * kChunkA, kChunkA, 128);
              const uint64_t bdesc = make_smem_desc(b_base + (sd.kind * 6 + j) * 1024, 32 * 16, 128);
              mma_bf16_ss(d_tmem + sd.kind * 32, adesc, bdesc, idesc, j > 0);
            }
          }
          if (sd.last) mma_commit(&m->bar_acc[q]);
        }
      }
    }
    __syncwarp();
  } else {
    // =================================== row warps: TWO threads per sample row ===========================================
    // group q = warp / 8; half = (warp / 4) & 1 selects the VM grid to gather (0 coarse, 1 fine) and the column half of
    // every epilogue; both halves of a row read the same TMEM lane (lane quarter = warp % 4).
    const int q = warp >> 3, half = (warp >> 2) & 1, gwarp = warp & 3;
    const int r = gwarp * 32 + lane;
    uint8_t* Aq = As + q * kABytes;
    uint8_t* a_row = Aq + r * 16;
    GroupMisc* gm = &m->grp[q];
    const uint32_t taddr_row = tmem + ((uint32_t)(gwarp * 32) << 16) + q * 256;
    uint32_t pacc = 0;
    const bool mask_near = !(a.flags & EDN_FLAG_TRAIN) && a.rmnearplane > 0.f;
    const float near_thr = a.rmnearplane / 128.0f;
    const int bar_id = 1 + q, bar_half = 3 + q;
    const uint32_t s_wsig = smem_u32(m->wsig), s_wrgb = smem_u32(m->wrgb);
    if (q == 1 && n_it > 0) {   // phase-shift the second ray group by ~half a ray so that its MMAs fall into the first group's gather / epilogues
      uint32_t spins = 0;
      while (*reinterpret_cast<volatile uint32_t*>(&m->skew_flag) == 0) { if (++spins > (1u << 22)) break; }   // a scheduling hint only: never a correctness dependency
    }
    auto epi = [&](uint32_t taddr, uint8_t* arow, int col_begin, int ncols, int mode, uint32_t bias_s, float* gout, uint32_t wsig,
                   uint32_t wrgb) {
      return (ablate_bits & 2) ? make_float4(0.f, 0.f, 0.f, 0.f) : layer_epilogue(taddr, arow, col_begin, ncols, mode, bias_s, gout, wsig, wrgb);
    };
    auto stamp = [&](int64_t it, int k) {
      if (a.trace && blockIdx.x == 0 && r == 0 && half == 0 && it >= 8 && it < 12) a.trace[((it - 8) * 2 + q) * 16 + k] = clock64();
    };
    float run_sum = 0.f;                    // half 0, thread r < 5: running ray sum (rgb, depth, acc) across the tiles of a ray
    int tile = -1;
    int64_t ray_raw = 2 * (int64_t)blockIdx.x + q - 2 * (int64_t)gridDim.x;
    const uint32_t bar_a_rank0 = PAIR ? mapa_u32(smem_u32(&m->bar_a[q]), 0) : 0u;
    auto signal_a = [&]() { if (PAIR) rows_signal_a_pair(bar_a_rank0); else rows_signal_a(&m->bar_a[q]); };
    for (int64_t it = 0; it < n_it; ++it) {
      stamp(it, 0);
      if (++tile == tpr || it == 0) { tile = 0; ray_raw += 2 * (int64_t)gridDim.x; }
      const int rg = tile * kRows + r;                  // sample index of this row within the ray
      const bool live = ray_raw < a.n_rays;             // odd ray count: the last pair's second ray is a masked duplicate
      const int64_t ray = live ? ray_raw : a.n_rays - 1;
      const float* rb = a.ray_batch + ray * 11;
      const float o[3] = {__ldg(rb + 0), __ldg(rb + 1), __ldg(rb + 2)};
      const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
      const float zv = a.z_vals[ray * S + min(rg, S - 1)];
      if (ablate_bits & 8) {
        if (half == 0) gm->z[r] = zv;
      } else if (half == 0) {
        gm->z[r] = zv;
        if (tile == 0 && r == 0) gm->tcarry = 1.0f;
        // ---- PE(pts) -> 64 A columns (full: chunks 8..15, lean: 24..31); the last column is the zero pad of K = 127 -> 128.
        // sin/cos of the base frequency by range-reduced MUFU, higher octaves by the double-angle recurrence
        // (abs error <= 2^9 * 1e-7, far below the bf16 resolution of the MMA operand)
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
          st_shared_v4(a_row + ((LEAN ? 24 : 8) + j) * kChunkA, pack_bf16x2(pe[8 * j], pe[8 * j + 1]), pack_bf16x2(pe[8 * j + 2], pe[8 * j + 3]),
                       pack_bf16x2(pe[8 * j + 4], pe[8 * j + 5]), pack_bf16x2(pe[8 * j + 6], pe[8 * j + 7]));
      } else if (tile == 0) {
        // ---- per-ray bias of color_net.0: b0 + W0[:, 128:155] . PE(viewdir), fp32; thread r -> columns r and r + 128 ----
        const float vd[3] = {__ldg(rb + 8), __ldg(rb + 9), __ldg(rb + 10)};
        float ped[kPeDir];
#pragma unroll
        for (int i = 0; i < 3; ++i) { ped[i] = vd[i]; fast_sincos(vd[i], &ped[3 + i], &ped[6 + i]); }
#pragma unroll
        for (int f = 1; f < kPeFreqDir; ++f) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float sp = ped[3 + 6 * (f - 1) + i], cp = ped[6 + 6 * (f - 1) + i];
            ped[3 + 6 * f + i] = 2.0f * sp * cp;
            ped[6 + 6 * f + i] = fmaf(-2.0f * sp, sp, 1.0f);
          }
        }
#pragma unroll 1
        for (int hc = 0; hc < 2; ++hc) {
          const int col = r + hc * kRows;
          float b = a.mlp.color0_b ? __ldg(a.mlp.color0_b + col) : 0.f;
          const float* w = a.mlp.color0_t + (size_t)128 * 256 + col;
#pragma unroll
          for (int j = 0; j < kPeDir; ++j) b = fmaf(__ldg(w + j * 256), ped[j], b);
          gm->bias[col] = b;
        }
      }
      named_bar_sync(bar_id, kGroupThreads);      // z[] (and bias[]) visible to the whole group
      stamp(it, 1);
      // ---- VM gather: half 0 gathers the coarse grid, half 1 the fine grid -> two 128 x 96 bf16 tiles ------------------
      if (!(ablate_bits & 1)) gather_tiles<T>(m->grids, Aq, gm->z, gwarp, lane, o, d, LEAN, half);
      signal_a();
      stamp(it, 2);
      if (!LEAN) {
        // ---- basis_mat outputs (coarse 32 | fine 32) -> A columns 0..63 ----------------------------------------------
        mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
        stamp(it, 3);
        epi(taddr_row, a_row, 32 * half, 32, kEpiPlain, 0u, nullptr, s_wsig, s_wrgb);
        signal_a();
        stamp(it, 4);
      }
      // ---- sigma_net.0 -> ReLU (+ sigma head: fp32 dot with sigma_net.1 row 0; partial per column half) -----------------
      mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
      stamp(it, 5);
      {
        const float4 hd = epi(taddr_row, a_row, 128 * half, 128, kEpiReluSigma, 0u, nullptr, s_wsig, s_wrgb);
        gm->headp[half][r][3] = hd.x;
      }
      signal_a();
      stamp(it, 6);
      if (!LEAN) {
        // ---- sigma_net.1 -> geo (128, linear) ----------------------------------------------------------------------------
        mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
        stamp(it, 7);
        epi(taddr_row, a_row, 64 * half, 64, kEpiPlain, 0u,
                       (a.feat && live && rg < S) ? a.feat + ((size_t)ray * S + rg) * 128 : nullptr, s_wsig, s_wrgb);
        signal_a();
      }
      if (q == 0 && it == 0 && r == 0 && half == 0) *reinterpret_cast<volatile uint32_t*>(&m->skew_flag) = 1;
      stamp(it, 8);
      // ---- color_net.0 (+ per-ray view-dir bias) -> ReLU ------------------------------------------------------------------
      mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
      stamp(it, 9);
      epi(taddr_row, a_row, 128 * half, 128, kEpiRelu, smem_u32(gm->bias), nullptr, s_wsig, s_wrgb);
      signal_a();
      stamp(it, 10);
      // ---- color_net.1 -> ReLU -> rgb head (fp32 dot with color_net.2; partial per column half) --------------------------
      mbar_wait(&m->bar_acc[q], pacc); pacc ^= 1; tc_fence_after();
      stamp(it, 11);
      {
        const float4 hd = epi(taddr_row, a_row, 128 * half, 128, kEpiReluRgb, a.mlp.color1_b ? smem_u32(m->bias1) : 0u, nullptr,
                                         s_wsig, s_wrgb);
        gm->headp[half][r][0] = hd.x; gm->headp[half][r][1] = hd.y; gm->headp[half][r][2] = hd.z;
      }
      tc_fence_before();
      named_bar_sync(bar_id, kGroupThreads);      // both halves' head partials visible; all TMEM reads of this tile done
      stamp(it, 12);
      if (half == 0) {
        // ---- compositing (voxnerf.py:153-201), one thread per sample row --------------------------------------------------
        float col[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          col[i] = sigmoidf_(gm->headp[0][r][i] + gm->headp[1][r][i] + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + i) : 0.f));
        const float sig_raw = gm->headp[0][r][3] + gm->headp[1][r][3];
        float alpha = 0.f;
        if (rg < S - 1) {
          const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
          const float znext = (r + 1 < kRows) ? gm->z[r + 1] : a.z_vals[ray * S + rg + 1];
          const float dist = __fmul_rn(znext - zv, dnorm);
          float sg = sig_raw;
          if (a.noise) sg += __ldg(a.noise + ray * (S - 1) + rg);
          sg = fmaxf(sg, 0.f);
          if (mask_near && !(znext > near_thr)) sg = 0.f;
          alpha = 1.0f - expf(-__fmul_rn(sg, dist));
        } else if (rg == S - 1) {
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
        if (lane == 31) gm->wtot[gwarp] = t;
        named_bar_sync(bar_half, kRows);
        Tr *= gm->tcarry;                       // transmittance accumulated over the previous tiles of this ray
        for (int w2 = 0; w2 < gwarp; ++w2) Tr *= gm->wtot[w2];
        const float wgt = alpha * Tr;
        if (live && rg < S) a.weights[ray * S + rg] = wgt;
        float red[5] = {wgt * col[0], wgt * col[1], wgt * col[2], wgt * zv, wgt};
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
        if (r < 5) {
          if (tile == 0) run_sum = 0.f;
          run_sum += gm->red[0][r] + gm->red[1][r] + gm->red[2][r] + gm->red[3][r];
          if (live && tile == tpr - 1) {
            if (r < 3) a.rgb[ray * 3 + r] = run_sum; else if (r == 3) a.depth[ray] = run_sum; else a.acc[ray] = run_sum;
          }
        }
        if (r == 5) gm->tcarry = gm->tcarry * gm->wtot[0] * gm->wtot[1] * gm->wtot[2] * gm->wtot[3];
      }
      named_bar_sync(bar_id, kGroupThreads);      // red[] / wtot[] / z[] / headp[] free for the next ray
      stamp(it, 13);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // the peer may still read this CTA's operands / signal its barriers
  if (warp == kRowWarps) { if (PAIR) tmem_dealloc2(tmem, kTmemCols); else tmem_dealloc(tmem, kTmemCols); }
}

int ensure_schedule() {
  static bool uploaded = false;
  if (!uploaded) {
    std::vector<StepDesc> v = build_steps(false), vl = build_steps(true);
    if ((int)v.size() != kStepsFull || (int)vl.size() != kStepsLean) { set_error("fine_tc: schedule size mismatch"); return EDN_E_INVALID; }
    EDN_CUDA_OK(cudaMemcpyToSymbol(c_steps, v.data(), sizeof(StepDesc) * v.size()));
    EDN_CUDA_OK(cudaMemcpyToSymbol(c_steps_lean, vl.data(), sizeof(StepDesc) * vl.size()));
    uploaded = true;
  }
  return 0;
}

// C[M][N] = A[M][K] . B[K][N] (row-major fp32, leading dimensions lda / ldb / ldc): weight folding at pack time only
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
int launch_variant(const FineArgs& a, const uint8_t* blob, unsigned gx, cudaStream_t st) {
  EDN_CUDA_OK(cudaFuncSetAttribute(fine_fwd_tc_kernel<T, LEAN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  fine_fwd_tc_kernel<T, LEAN, false><<<gx, kThreads, kSmemBytes, st>>>(a, blob);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

// cta_group::2 variant: clusters of 2 CTAs (one TPC), an even grid
template <typename T>
int launch_pair(const FineArgs& a, const uint8_t* blob, unsigned gx, cudaStream_t st) {
  auto kern = fine_fwd_tc_kernel<T, true, true>;
  EDN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx & ~1u);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EDN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a, blob));
  return EDN_OK;
}

}  // namespace

int launch_fine_tc(const FineArgs& a_in, int grid_dtype, cudaStream_t st) {
  FineArgs a = a_in;
  EDN_REQUIRE(a.S >= 2, "edn_render_fine_fwd(bf16): n_samples must be >= 2, got %d", a.S);
  EDN_REQUIRE(a.mlp.tc_blob != nullptr, "edn_render_fine_fwd(bf16): edn_field_mlp.tc_blob is NULL (call edn_pack_fine_tc)");
  int rc = ensure_schedule();
  if (rc) return rc;
  const int64_t n_pairs = (a.n_rays + 1) / 2;
  const unsigned gx = (unsigned)(n_pairs < (int64_t)num_sms() ? n_pairs : (int64_t)num_sms());
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(a.mlp.tc_blob);
  const char* fv = getenv("EDN_TC_FULL");                       // dev switch: force the unfolded schedule
  const bool lean = (a.feat == nullptr) && !(fv && fv[0] == '1');   // depth_feature (geo) only exists in the full schedule
  static const int ablate = [] { const char* e = getenv("EDN_TC_ABLATE"); return e ? atoi(e) : 0; }();
  a.ablate = EDN_TC_ABLATE_BUILD ? ablate : 0;
  if (ablate) {
    static bool told = false;
    if (!told) {
      fprintf(stderr, EDN_TC_ABLATE_BUILD ? "[evdeblur_b200] EDN_TC_ABLATE=%d: fine tensor-core kernel runs a TIMING-ONLY ablation, outputs are invalid\n"
                                          : "[evdeblur_b200] EDN_TC_ABLATE=%d ignored: library built without -DEDN_TC_ABLATE_BUILD=1\n", ablate);
      told = true;
    }
  }
  static const bool force_v1 = [] { const char* e = getenv("EDN_TC_V1"); return e && e[0] == '1'; }();   // dev switch: round-1 kernel
  // fine_tc2: the lean weight stream; with depth_feature requested it adds the geo layer (sigma_net.1's geo columns: the second
  // 64 KB layer of the full section, [K-step][128 x 16])
  if (a.S <= kRows && !force_v1 && !ablate && !(fv && fv[0] == '1'))
    return launch_fine_tc2(a, grid_dtype, blob + kOffLean, blob + kOffFull + 65536, st);
  const char* tr = getenv("EDN_TC_TRACE");
  if (tr && tr[0] == '1') {   // dev tooling: print the phase time line of CTA 0 (synchronises!)
    FineArgs b = a;
    long long* buf = nullptr;
    EDN_CUDA_OK(cudaMallocManaged(&buf, 8 * 16 * sizeof(long long)));
    memset(buf, 0, 8 * 16 * sizeof(long long));
    b.trace = buf;
    const char* pv2 = getenv("EDN_TC_PAIR");
    if (lean && pv2 && pv2[0] == '1') rc = launch_pair<__nv_bfloat16>(b, blob, (unsigned)num_sms(), st);
    else rc = lean ? launch_variant<__nv_bfloat16, true>(b, blob, gx, st) : launch_variant<__nv_bfloat16, false>(b, blob, gx, st);
    if (rc) return rc;
    EDN_CUDA_OK(cudaStreamSynchronize(st));
    static const char* names[14] = {"start", "pe+bias", "gather", "w.basis", "e.ft", "w.L1", "e.L1", "w.L2", "e.L2", "w.L3", "e.L3", "w.L4", "e.L4", "composite"};
    for (int i = 0; i < 8; ++i) {
      fprintf(stderr, "[trace %s it=%d q=%d] t0=%lld :", lean ? "lean" : "full", 8 + i / 2, i % 2, buf[i * 16] - buf[0]);
      long long prev = buf[i * 16];
      for (int k = 1; k < 14; ++k) {
        if (buf[i * 16 + k] == 0) continue;
        fprintf(stderr, " %s=%lld", names[k], buf[i * 16 + k] - prev);
        prev = buf[i * 16 + k];
      }
      fprintf(stderr, "\n");
    }
    cudaFree(buf);
    return EDN_OK;
  }
  const char* pv = getenv("EDN_TC_PAIR");                       // dev switch: CTA-pair (cta_group::2) variant of the lean schedule
  if (lean && pv && pv[0] == '1' && gx >= 2 && num_sms() % 2 == 0) {
    const unsigned gp = (unsigned)(n_pairs < (int64_t)num_sms() ? ((n_pairs + 1) & ~1ll) : (int64_t)num_sms());
    static bool said = false;
    if (!said) { fprintf(stderr, "[edn] fine pass: CTA-pair (cta_group::2) variant, grid %u\n", gp); said = true; }
    return grid_dtype == EDN_BF16 ? launch_pair<__nv_bfloat16>(a, blob, gp, st) : launch_pair<float>(a, blob, gp, st);
  }
  if (grid_dtype == EDN_BF16)
    return lean ? launch_variant<__nv_bfloat16, true>(a, blob, gx, st) : launch_variant<__nv_bfloat16, false>(a, blob, gx, st);
  return lean ? launch_variant<float, true>(a, blob, gx, st) : launch_variant<float, false>(a, blob, gx, st);
}

}  // namespace edn

extern "C" int64_t edn_fine_tc_blob_bytes(void) { return edn::kBlobBytes; }

extern "C" int64_t edn_fine_tc_pack_workspace_floats(void) { return 2 * 256 * 256 + 32768; }
namespace edn { int64_t fine_tc3_blob_offset() { return kOffTc3; } }

extern "C" int edn_pack_fine_tc(const edn_field_mlp* mlp, const float* basis_t_coarse, const float* basis_t_fine, float* workspace,
                                void* blob, void* stream) {
  using namespace edn;
  EDN_REQUIRE(mlp && basis_t_coarse && basis_t_fine && blob && workspace, "edn_pack_fine_tc: null pointer");
  EDN_REQUIRE(mlp->hidden == 256 && mlp->geo_feat == 128 && mlp->sigma0_t && mlp->sigma1_t && mlp->sigma1_v && mlp->color0_t &&
              mlp->color1_t && mlp->color2_t, "edn_pack_fine_tc: needs the fine field (hidden=256, geo_feat=128)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* b = reinterpret_cast<uint8_t*>(blob);
  // ---- folded layers of the lean schedule (fp32 products, rounded to bf16 once by the packer) -------------------------
  float* f1 = workspace;              // [256][256]: rows 0..95 basis_c . sigma0[0:32], 96..191 basis_f . sigma0[32:64], 192..255 sigma0[64:128]
  float* f23 = workspace + 256 * 256; // [256][256] = sigma1(geo) [256][128] . color0(geo rows) [128][256]
  fold_matmul_kernel<<<(96 * 256 + 255) / 256, 256, 0, st>>>(basis_t_coarse, 32, mlp->sigma0_t, 256, f1, 256, 96, 32, 256);
  fold_matmul_kernel<<<(96 * 256 + 255) / 256, 256, 0, st>>>(basis_t_fine, 32, mlp->sigma0_t + 32 * 256, 256, f1 + 96 * 256, 256, 96, 32, 256);
  EDN_CUDA_OK(cudaMemcpyAsync(f1 + 192 * 256, mlp->sigma0_t + 64 * 256, sizeof(float) * 64 * 256, cudaMemcpyDeviceToDevice, st));
  fold_matmul_kernel<<<(256 * 256 + 255) / 256, 256, 0, st>>>(mlp->sigma1_t, 128, mlp->color0_t, 256, f23, 256, 256, 128, 256);
  // blob = [basis coarse 6 KB][basis fine 6 KB] | full: [sigma0 64 KB][sigma1(geo) 64 KB][color0(geo rows) 64 KB][color1 128 KB]
  //        | lean: [fold1 128 KB][fold23 128 KB][color1 128 KB]
  struct Src { const float* wt; int ld, kv, nv, K, N; };
  const Src src[9] = {{basis_t_coarse, 32, 96, 32, 96, 32}, {basis_t_fine, 32, 96, 32, 96, 32},
                      {mlp->sigma0_t, 256, 128, 256, 128, 256}, {mlp->sigma1_t, 128, 256, 128, 256, 128},
                      {mlp->color0_t, 256, 128, 256, 128, 256}, {mlp->color1_t, 256, 256, 256, 256, 256},
                      {f1, 256, 256, 256, 256, 256}, {f23, 256, 256, 256, 256, 256}, {mlp->color1_t, 256, 256, 256, 256, 256}};
  size_t off = 0;
  for (int L = 0; L < 9; ++L) {
    const int total = src[L].K * src[L].N;
    tc::pack_layer_kernel<<<(total + 255) / 256, 256, 0, st>>>(src[L].wt, src[L].ld, src[L].kv, src[L].nv, src[L].K, src[L].N, 0,
                                                               reinterpret_cast<__nv_bfloat16*>(b + off));
    off += (size_t)total * 2;
  }
  // pair section: [layer][rank][K = 256 x N/2 = 128] halves of the three lean layers
  const float* lean_src[3] = {f1, f23, mlp->color1_t};
  for (int L = 0; L < 3; ++L)
    for (int c = 0; c < 2; ++c) {
      tc::pack_layer_kernel<<<(256 * 128 + 255) / 256, 256, 0, st>>>(lean_src[L] + 128 * c, 256, 256, 128, 256, 128, 0,
                                                                     reinterpret_cast<__nv_bfloat16*>(b + off));
      off += 256 * 128 * 2;
    }
  EDN_REQUIRE((int64_t)off == kOffTc3, "edn_pack_fine_tc: blob size mismatch");
  {
    int rc3 = pack_fine_tc3_section(f1, f23, mlp->color1_t, reinterpret_cast<uint8_t*>(workspace + 2 * 256 * 256), b + kOffTc3, st);
    if (rc3) return rc3;
  }
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

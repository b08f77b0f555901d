// Fine pass of render_rays on tcgen05 / TMEM, second generation ("v2"): decoupled gather, activations chained through TMEM.
// Replaces networks/renderer.py:190-217 + networks/pdrf/voxnerf.py:203-259,153-201 for the FVR field (lean schedule: basis_mat
// folded into sigma_net.0, sigma_net.1(geo) folded into color_net.0 -> three 256 -> 256 layers; see fine_tc.cu).
//
// What round 1's kernel (fine_tc.cu) was bound by: a ray group's VM gather, its three layer epilogues and its MMAs form ONE serial
// chain, and shared memory (2 x 64 KB A operands + rings) / TMEM (2 x 256 accumulator columns) capped it at two such chains per SM.
// Here the chain is cut in two and the hidden activations never touch shared memory:
//   gather warps (16):  ray k+1: PE + view-direction bias + cooperative VM gather of both grids -> A_in[(k+1) & 1] (64 KB, UMMA
//                       K-major layout) WHILE ray k is in the MLP: the gather is off the critical chain, double buffered; the 32
//                       gather rounds of a ray are pulled from a shared counter (work stealing).
//   MMA issuer (1 warp, one elected lane): layer 1 reads A_in from shared memory (SS), layers 2 / 3 read their A operand from
//                       TENSOR MEMORY (TS): the epilogue writes relu(acc) as packed bf16 straight back into TMEM (tcgen05.st), so
//                       shared memory holds only layer-1 inputs and the weight ring.  Every layer is issued as two independent
//                       N = 128 chains into two of three rotating 128-column accumulators; the next layer's K-steps 4i..4i+3 only
//                       wait for quarter i of this layer's epilogue (hand[4]).  The per-ray bias rides on one extra K-step.
//   epilogue warps (8): two threads per sample row (two 64-column quarters of every layer each); per 32-column chunk: tcgen05.ld ->
//                       ReLU -> sigma / rgb head partials (constant-bank weights) -> bf16x2 -> tcgen05.st into the next layer's
//                       A operand, everything indexed at compile time; finally sigma -> alpha compositing (warp-shuffle scan).
//   weight stream (1 warp): the three layers' weights ([layer][K-step pair][256 x 16 x 2]) through a 3 x 16 KB ring with
//                       cp.async.bulk (TMA engine); stages are released by tcgen05.commit.
// TMEM map (512 columns): [0,128) the A operand of layers 2 / 3 (rewritten in place), [128,512) three 128-column accumulator slots.
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <type_traits>
#include "common.cuh"
#include "fine_args.cuh"
#include "tc_common.cuh"
#include "tc_rows.cuh"

namespace edn {
namespace {

using namespace tc;

constexpr int kRows = 128;
// Warp roles (7 warpgroups = 896 threads, launched at 72 registers per thread, then re-balanced with setmaxnreg):
//   warps 0-15  gather   (4 warpgroups, raised to 88 registers: two 6-load tap tasks in flight per lane)
//   warps 16-23 epilogue (2 warpgroups, lowered to 64 registers): TWO threads per sample row
//   warp 24 MMA issuer, warp 25 weight stream, warps 26-27 idle (the warpgroup drops to 24 registers and donates the rest)
constexpr int kGatherWarps = 16, kEpiWarps = 8;
constexpr int kWarpMma = kGatherWarps + kEpiWarps, kWarpLoad = kWarpMma + 1;
constexpr int kThreads = (kGatherWarps + kEpiWarps + 4) * 32;      // 896
constexpr int kRoleThreads = kGatherWarps * 32;                     // gather threads (512)
constexpr int kEpiThreads = kEpiWarps * 32;                         // 256
constexpr int kQuarterThreads = 128;                                // epilogue threads serving one 64-column quarter of a layer
constexpr int kNst = 3;                    // weight ring stages (the stream is far from limiting: removing it changes nothing, ablation bit 2)
constexpr int kSlots = 3;                  // per-ray slots (z, bias operand): one more than A_in buffers, so that the gather of ray k
                                           // only waits for layer 1 of ray k - 2 (A_in free), not for its compositing
constexpr int kStageBytes = 16384;         // 2 K-steps of a layer (N = 256 rows x 16 x 2 B = 8 KB per K-step): the lean section of the blob
constexpr int kKstepBytes = 8192;
constexpr int kABytes = 65536;
constexpr int kStagesPerRay = 3 * 8;       // layers x (16 K-steps / 2)
constexpr uint32_t kTmemCols = 512;
// TMEM: [0,128) A operand of layers 2 / 3 (rewritten in place), [128,512) THREE 128-column accumulator slots.  A layer accumulates
// into two slots (output columns 0..127 | 128..255, issued K-major as two independent chains); layer n uses slots (2n % 3,
// (2n + 1) % 3), so the next layer's first slot is the spare one and its second slot is this layer's first, which the epilogue
// drains first: the next layer's K-steps 0..7 (they only read the first half of the new A operand) overlap the second half of
// this layer's epilogue.
constexpr uint32_t kColA = 0, kColAcc = 128;
#ifndef EDN_TC2_PACKED
#define EDN_TC2_PACKED 1                   // fp32x2 packed arithmetic in the gather / epilogue (same results as scalar)
#endif
constexpr int kNQ = 2;                     // N = 128 per instruction: 86 (TS) / 119 (SS) cycles each, measured (fine_tc2 ablate 16 / 48)

// Head weights (sigma_net.1 row 0: [0,256); color_net.2 channel-major: [256,1024)) in CONSTANT memory: every lane of an epilogue warp
// reads the same element, so the FFMAs take them as constant-bank operands -- no load instruction, no shared-memory latency (the
// epilogue's top stall was the short scoreboard behind LDS under the gather's LSU traffic).  A few slots, rotated per launch, so
// that launches in flight on different streams / models do not overwrite each other's copy.
constexpr int kHeadSlots = 8, kHeadFloats = 1024;
__constant__ float c_heads[kHeadSlots][kHeadFloats];

struct alignas(16) RaySlot {              // per-ray data produced by the gather warps, read by the epilogue warps
  float z[kRows];
};
struct Misc {
  uint64_t a_full[2], a_empty[2], slot_free[kSlots];
  uint64_t w_full[kNst], w_empty[kNst];
  uint64_t acc_full, hand[4];           // hand[q]: the epilogue finished quarter q of a layer (A' K 64q..64q+63 written; q = 1 / 3: slot x / y drained)
  GridDev grids[2];
  uint32_t tmem_base, round_ctr[2], pad[1];      // round_ctr[buf]: next gather round of the ray going into A_in[buf] (work stealing)
  alignas(16) float bias1[256];
  RaySlot slot[kSlots];
  alignas(16) float headp[4][kRows][4];   // per column quarter: partial rgb (xyz) / sigma (w) heads
  float red[4][8];
  float wtot[4];
};
static_assert(offsetof(Misc, bias1) % 16 == 0 &&
              offsetof(Misc, slot) % 16 == 0 && offsetof(Misc, headp) % 16 == 0, "float4 alignment");
// The per-ray bias of color_net.0 (b0 + W0[:, 128:155] . PE(viewdir), the same for all 128 rows) is added BY THE TENSOR CORE: one
// extra K-step whose A operand is a constant tile with 1.0 in K columns 0, 1 and whose B operand holds (bf16 hi, bf16 lo) of the
// bias in those two K columns -- exact to 2^-17, and the epilogue of that layer needs no shared-memory loads at all.
constexpr int kOneBytes = 4096;            // A: 128 rows x 16 K bf16
constexpr int kBiasBBytes = 8192;          // B: 256 rows x 16 K bf16, one per ray slot
constexpr int kSmemBytes = 2 * kABytes + kNst * kStageBytes + kOneBytes + kSlots * kBiasBBytes + (int)sizeof(Misc);
static_assert(kSmemBytes <= 232448, "shared memory budget");

// ---- TMEM store: 32 lanes x 16 consecutive 32-bit columns <- 16 registers per thread ------------------------------------
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 16 step (A: 128 lanes x 8 packed-bf16x2 columns), issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}

// FEAT: depth_feature (the geo output of sigma_net.1, voxnerf.py:221) is emitted as well -- an extra N = 128 layer on the SAME A operand
// as the folded second layer (geo = W_s1geo . relu(h1)), issued between layers 1 and 2 into ONE accumulator slot; its epilogue
// streams the 128 fp32 columns to a.feat and leaves the A operand untouched.  wgeo: sigma_net.1's geo columns, [K-step][128 x 16] bf16.
template <typename T, bool FEAT>
__global__ void __launch_bounds__(kThreads, 1) fine_fwd_tc2_kernel(const FineArgs a, const uint8_t* __restrict__ wblob,
                                                                   const uint8_t* __restrict__ wgeo, const int hslot) {
  constexpr int kRayStages = FEAT ? kStagesPerRay + 4 : kStagesPerRay;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;                                   // [2][64 KB] layer-1 operands (double buffered)
  uint8_t* Ws = smem + 2 * kABytes;                     // weight ring
  uint8_t* A_one = Ws + kNst * kStageBytes;
  uint8_t* B_bias = A_one + kOneBytes;                  // [kSlots][8 KB]
  Misc* m = reinterpret_cast<Misc*>(B_bias + kSlots * kBiasBBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (kOneBytes + kSlots * kBiasBBytes) / 16; i += kThreads) {      // zero, then 1.0 | 1.0 in K columns 0, 1 of every A row
    const bool one = i < kRows;                          // 16-byte piece i < 128 of A_one = (row i, K 0..7)
    st_shared_v4(A_one + i * 16, one ? 0x3F803F80u : 0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();                              // these tiles are read by the tensor core (async proxy)

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(&m->a_full[b], kRoleThreads); mbar_init(&m->a_empty[b], 1); }
    for (int b = 0; b < kSlots; ++b) mbar_init(&m->slot_free[b], kEpiThreads);
    for (int s = 0; s < kNst; ++s) { mbar_init(&m->w_full[s], 1); mbar_init(&m->w_empty[s], 1); }
    mbar_init(&m->acc_full, 1);
    for (int q = 0; q < 4; ++q) mbar_init(&m->hand[q], kQuarterThreads);
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc(&m->tmem_base, kTmemCols);
  for (int i = tid; i < 256; i += kThreads) {
    m->bias1[i] = a.mlp.color1_b ? __ldg(a.mlp.color1_b + i) : 0.f;
  }
  if (tid == 32) { m->grids[0] = a.gc; m->grids[1] = a.gf; m->round_ctr[0] = 0; m->round_ctr[1] = 0; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = m->tmem_base;
  const int S = a.S;
  // register re-balancing (first statement of every role branch): the MMA / loader warpgroup donates, the gather warpgroups take
  // dev tooling (EDN_TC_TRACE=1): clock64 stamps of CTA 0 for iterations 8..11: trace[role][it - 8][slot], role 0 gather, 1 MMA, 2 epilogue
  auto stamp = [&](int role, int64_t it, int k) {
    if (a.trace && blockIdx.x == 0 && it >= 8 && it < 12) a.trace[(role * 4 + (it - 8)) * 16 + k] = clock64();
  };
  const int64_t n_my = (a.n_rays > (int64_t)blockIdx.x) ? (a.n_rays - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (a.ablate & 16) {      // dev microbenchmark: ONLY the MMA warp runs, no waits at all -> pure tcgen05.mma issue / execution rate
    if (warp == kWarpMma && n_my > 0) {
      const uint32_t idesc = make_idesc_bf16(128, 256 / kNQ);
      const uint32_t w_base = smem_u32(Ws), a_in = smem_u32(As);
      const bool ts = (a.ablate & 32) != 0;
      for (int64_t it = 0; it < n_my * 3; ++it) {
        for (int st8 = 0; st8 < 8; ++st8) {
          const uint64_t bd0 = make_smem_desc(w_base + (st8 & 3) * kStageBytes, 256 * 16, 128);
          const uint64_t ad0 = make_smem_desc(a_in + st8 * 2 * 2 * kChunkA, kChunkA, 128);
          if (elect_one()) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int q = 0; q < kNQ; ++q) {
                if (ts) mma_bf16_ts(tmem + kColAcc + q * (256 / kNQ), tmem + kColA + (st8 * 2 + i) * 8, bd0 + (uint64_t)((i * kKstepBytes + q * (256 / kNQ) * 16) >> 4), idesc, 1);
                else mma_bf16_ss(tmem + kColAcc + q * (256 / kNQ), ad0 + (uint64_t)(i * (2 * kChunkA >> 4)), bd0 + (uint64_t)((i * kKstepBytes + q * (256 / kNQ) * 16) >> 4), idesc, 1);
              }
            if (a.ablate & 64) mma_commit(&m->w_empty[st8 & 3]);
          }
          __syncwarp();
        }
      }
      if (elect_one()) mma_commit(&m->acc_full);
      __syncwarp();
      mbar_wait(&m->acc_full, 0);
    }
  } else if (warp == kWarpLoad) {
    // =================================== weight stream =====================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (n_my > 0) {      // converged warp, one elected lane issues (uniform operands: no per-lane retry loop around UBLKCP)
      const uint32_t total = (uint32_t)n_my * kRayStages;
      for (uint32_t g = 0; g < total; ++g) {
        const int s = g % kNst;
        const uint32_t use = g / kNst;
        if (use > 0) mbar_wait(&m->w_empty[s], (use - 1) & 1);
        const uint32_t j = g % kRayStages;      // stage within the ray: lean layer 1 | (FEAT: geo, 4 stages) | lean layers 2, 3
        const uint8_t* src = (!FEAT || j < 8) ? wblob + (size_t)j * kStageBytes
                           : (j < 12 ? wgeo + (size_t)(j - 8) * kStageBytes : wblob + (size_t)(j - 4) * kStageBytes);
        if (elect_one()) {
          if (a.ablate & 2) {
            mbar_expect_tx(&m->w_full[s], 0);
          } else {
            mbar_expect_tx(&m->w_full[s], kStageBytes);
            bulk_g2s(Ws + s * kStageBytes, src, kStageBytes, &m->w_full[s]);
          }
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp > kWarpLoad) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");      // idle warps of the donor warpgroup
  } else if (warp == kWarpMma) {
    // =================================== MMA issuer (converged warp, one elected lane issues) ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (n_my > 0) {
      // K-major issue with kNQ INDEPENDENT accumulator chains per K-step: back-to-back MMAs into the same accumulator serialise on
      // the tensor pipe's accumulate latency (~140 cycles, measured: N = 64 MMAs issued quarter by quarter ran at 140 cycles each
      // instead of 32), so consecutive instructions always target different TMEM columns.
      constexpr int kNsub = 256 / kNQ;
      const uint32_t idesc = make_idesc_bf16(128, kNsub);
      const uint32_t w_base = smem_u32(Ws);
      uint32_t g = 0;                       // ring stage counter
      uint32_t n_layer = 0;                 // layers issued so far (3 per ray; 4 with FEAT)
      uint32_t cs = 0;                      // accumulator slot counter: a layer takes slots cs % 3, (cs + 1) % 3 (the geo layer: one)
      for (int64_t it = 0; it < n_my; ++it) {
        const int buf = (int)(it & 1);
        const uint32_t a_in = smem_u32(As) + buf * kABytes;
        if (lane == 0) stamp(1, it, 0);
#pragma unroll 1
        for (int L = 0; L < 3; ++L) {
          const uint32_t dx = tmem + kColAcc + (cs % 3) * 128, dy = tmem + kColAcc + ((cs + 1) % 3) * 128;
          cs += 2;
          const uint32_t a_tm = tmem + kColA;
          // Issue order of a layer (chain x -> slot dx = the spare one, chain y -> slot dy = the previous layer's x slot):
          //   after hand[0] (A' K 0..63):            x.K0-3
          //   after hand[1] (dy drained, K 64..127): y.K0-3, then x / y alternating K4-7
          //   after hand[2] / hand[3]:               K8-11 / K12-15
          // so only the first quarter of the previous epilogue and the last quarter of this layer's MMAs are exposed.
          const uint32_t hp = (n_layer - 1) & 1;
          if (n_layer > 0) mbar_wait(&m->hand[0], hp);
          if (L == 0) mbar_wait(&m->a_full[buf], (uint32_t)(it >> 1) & 1);
          const uint32_t g0 = g;
          auto issue = [&](int st8, bool do_x, bool do_y) {      // the two K-steps of ring stage g0 + st8
            const int s = (g0 + st8) % kNst;
            const uint64_t bd0 = make_smem_desc(w_base + s * kStageBytes, 256 * 16, 128);
            if (L == 0) {
              const uint64_t ad0 = make_smem_desc(a_in + st8 * 2 * 2 * kChunkA, kChunkA, 128);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                if (do_x) mma_bf16_ss(dx, ad0 + (uint64_t)(i * (2 * kChunkA >> 4)), bd0 + (uint64_t)((i * kKstepBytes) >> 4), idesc, (st8 | i) > 0);
                if (do_y) mma_bf16_ss(dy, ad0 + (uint64_t)(i * (2 * kChunkA >> 4)), bd0 + (uint64_t)((i * kKstepBytes + 128 * 16) >> 4), idesc, (st8 | i) > 0);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                if (do_x) mma_bf16_ts(dx, a_tm + (st8 * 2 + i) * 8, bd0 + (uint64_t)((i * kKstepBytes) >> 4), idesc, (st8 | i) > 0);
                if (do_y) mma_bf16_ts(dy, a_tm + (st8 * 2 + i) * 8, bd0 + (uint64_t)((i * kKstepBytes + 128 * 16) >> 4), idesc, (st8 | i) > 0);
              }
            }
          };
          auto wait_stage = [&](int st8) { mbar_wait(&m->w_full[(g0 + st8) % kNst], ((g0 + st8) / kNst) & 1); };
          wait_stage(0); wait_stage(1);
          tc_fence_after();
          if (elect_one()) { issue(0, true, false); issue(1, true, false); }
          __syncwarp();
          if (n_layer > 0) { mbar_wait(&m->hand[1], hp); tc_fence_after(); }
          if (elect_one()) {
            issue(0, false, true); mma_commit(&m->w_empty[(g0 + 0) % kNst]);
            issue(1, false, true); mma_commit(&m->w_empty[(g0 + 1) % kNst]);
          }
          __syncwarp();
#pragma unroll 1
          for (int st8 = 2; st8 < 8; ++st8) {
            if (n_layer > 0 && (st8 == 4 || st8 == 6)) mbar_wait(&m->hand[st8 >> 1], hp);
            wait_stage(st8);
            tc_fence_after();
            if (elect_one()) { issue(st8, true, true); mma_commit(&m->w_empty[(g0 + st8) % kNst]); }
            __syncwarp();
          }
          g += 8;
          if (L == 1) {      // + per-ray bias: [1 1 0 ..] x [hi lo 0 ..]^T into both chains
            if (elect_one()) {
              const uint64_t ad = make_smem_desc(smem_u32(A_one), kChunkA, 128);
              const uint64_t bd = make_smem_desc(smem_u32(B_bias) + (uint32_t)(it % kSlots) * kBiasBBytes, 256 * 16, 128);
              mma_bf16_ss(dx, ad, bd, idesc, 1);
              mma_bf16_ss(dy, ad, bd + (uint64_t)((128 * 16) >> 4), idesc, 1);
            }
            __syncwarp();
          }
          if (elect_one()) {
            mma_commit(&m->acc_full);
            if (L == 0) mma_commit(&m->a_empty[buf]);      // layer 1 has consumed A_in[buf]: the gather warps may refill it
          }
          __syncwarp();
          ++n_layer;
          if (lane == 0) stamp(1, it, 1 + L);
          if (FEAT && L == 0) {
            // ---- geo layer: [128 x 256] A (relu(h1), just written by the epilogue) x [256 -> 128], one chain, one slot -----------
            const uint32_t dg = tmem + kColAcc + (cs % 3) * 128;
            cs += 1;
            const uint32_t hp2 = (n_layer - 1) & 1;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) mbar_wait(&m->hand[q], hp2);        // the whole A operand is written
#pragma unroll 1
            for (int st4 = 0; st4 < 4; ++st4) {
              const int s = g % kNst;
              mbar_wait(&m->w_full[s], (g / kNst) & 1);
              tc_fence_after();
              const uint64_t bd0 = make_smem_desc(w_base + s * kStageBytes, 128 * 16, 128);
              if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  mma_bf16_ts(dg, a_tm + (st4 * 4 + i) * 8, bd0 + (uint64_t)((i * 4096) >> 4), idesc, (st4 | i) > 0);
                mma_commit(&m->w_empty[s]);
              }
              __syncwarp();
              ++g;
            }
            if (elect_one()) mma_commit(&m->acc_full);
            __syncwarp();
            ++n_layer;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < kGatherWarps) {
    // =================================== gather warps: TWO threads per sample row ===========================================
    // half 0: PE + coarse grid; half 1: view-direction bias + fine grid.  Runs one ray ahead of the MLP.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    // warp = (grid 1 bit, row group 2 bits, gi half 1 bit): FOUR threads per sample row
    const int half = warp >> 3, gwarp = (warp >> 1) & 3, gsub = warp & 1;
    const int r = gwarp * 32 + lane;
    for (int64_t it = 0; it < n_my; ++it) {
      const int buf = (int)(it & 1);
      const uint32_t use = (uint32_t)(it >> 1);
      if (tid == 0) stamp(0, it, 0);
      const int sl = (int)(it % kSlots);
      if (use > 0) {      // one polling warp, the rest of the gather group blocks in a named barrier (see the epilogue warps)
        if (warp == 0) {
          mbar_wait(&m->a_empty[buf], (use - 1) & 1);
          if (it >= kSlots) mbar_wait(&m->slot_free[sl], (uint32_t)(it / kSlots - 1) & 1);
          if (lane == 0) m->round_ctr[buf] = 0;      // every warp is past its last pull from this counter (two rays ago)
        }
        named_bar_sync(5, kRoleThreads);
      }
      if (tid == 0) stamp(0, it, 1);
      uint8_t* Aq = As + buf * kABytes;
      uint8_t* a_row = Aq + r * 16;
      RaySlot* slot = &m->slot[sl];
      const int64_t ray = (int64_t)blockIdx.x + it * gridDim.x;
      const float* rb = a.ray_batch + ray * 11;
      const float o[3] = {__ldg(rb + 0), __ldg(rb + 1), __ldg(rb + 2)};
      const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
      if (gsub != 0) {
        // (the second warp of a (grid, row group) only gathers)
      } else if (half == 0) {
        const float zv = a.z_vals[ray * S + min(r, S - 1)];
        slot->z[r] = zv;
        // PE(pts) -> A chunks 24..31 (64 columns, the last one is the zero pad of K = 127 -> 128)
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
          st_shared_v4(a_row + (24 + j) * kChunkA, pack_bf16x2(pe[8 * j], pe[8 * j + 1]), pack_bf16x2(pe[8 * j + 2], pe[8 * j + 3]),
                       pack_bf16x2(pe[8 * j + 4], pe[8 * j + 5]), pack_bf16x2(pe[8 * j + 6], pe[8 * j + 7]));
      } else {
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
          uint32_t hi, lo;
          split_bf16x2(b, 0.f, hi, lo);                   // low halves: bf16 hi / lo of b
          *reinterpret_cast<uint32_t*>(B_bias + sl * kBiasBBytes + col * 16) = (hi & 0xffffu) | (lo << 16);     // K column 0: hi, 1: lo
        }
      }
      if (tid == 0) stamp(0, it, 2);
      if (!(a.ablate & 1)) {      // lean A layout: coarse-grid products in chunks 0..11, fine-grid products in chunks 12..23
        // 32 rounds per ray (grid x row group x 8-point group), PULLED from a shared counter: the warps that just did the PE / bias
        // work take fewer rounds than the others (a fixed 2 rounds per warp left half of the warps 3 k cycles short every ray).
        // The depths come straight from global memory (the z[] copy in shared memory is for the epilogue), so no barrier is needed.
        const float* z_g = a.z_vals + ray * S;
        for (;;) {
          int k = 0;
          if (lane == 0) k = (int)atomicAdd(&m->round_ctr[buf], 1u);
          k = __shfl_sync(0xffffffffu, k, 0);
          if (k >= 32) break;
          const int hk = k >> 4, gk = (k >> 2) & 3, gik = k & 3;
          gather_points<T>(m->grids[hk], Aq, hk ? 12 : 0, gk, lane, gik, gik + 1, [&](int pt, float (&p)[3]) {
            const float zv = __ldg(z_g + min(pt, S - 1));
#pragma unroll
            for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], zv));
          });
        }
      }
      fence_proxy_async_smem();                   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(&m->a_full[buf]);
      if (tid == 0) stamp(0, it, 3);
    }
  } else {
    // =================================== epilogue warps: TWO threads per sample row ============================================
    // (the warp index through a shuffle: provably warp-uniform, so that tq-derived constant-bank addresses stay in uniform registers)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    const int ew = __shfl_sync(0xffffffffu, warp, 0) - kGatherWarps;
    const int th = ew >> 2, gwarp = ew & 3;       // th: this thread serves the 64-column quarters th and th + 2 of every layer
    const float* __restrict__ cw = c_heads[hslot];
    const int r = gwarp * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(gwarp * 32) << 16);
    // plain shared-memory loads (the compiler may batch them; a `volatile` asm load per float4 serialises on its 29-cycle latency)
    const float4* s_bias1 = reinterpret_cast<const float4*>(m->bias1);
    const bool mask_near = !(a.flags & EDN_FLAG_TRAIN) && a.rmnearplane > 0.f;
    const float near_thr = a.rmnearplane / 128.0f;
    const bool has_bias1 = a.mlp.color1_b != nullptr;
    uint32_t n_use = 0;                            // completed accumulator uses (3 per ray; 4 with FEAT)
    uint32_t cs = 0;                               // accumulator slot counter, mirrors the MMA issuer's
    for (int64_t it = 0; it < n_my; ++it) {
      const int sl = (int)(it % kSlots);
      RaySlot* slot = &m->slot[sl];
      const int64_t ray = (int64_t)blockIdx.x + it * gridDim.x;
      float sig_part = 0.f, rr = 0.f, rg_ = 0.f, rbl = 0.f;
      const bool st0 = (ew == 0 && lane == 0);
      if (st0) stamp(2, it, 0);
      // One layer's epilogue; L, the pass (accumulator slot x / y) and the 32-column chunk are COMPILE-TIME constants, so the head
      // weights' constant-bank offsets are immediates off one uniform base and there is no per-chunk branching (the run-time-L
      // version spent ~60 % of the epilogue's issue slots on uniform address arithmetic, branches and dead ablation code).
      auto layer = [&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        // ONE warp polls the mbarrier, the others block in a named barrier: a polling warp re-issues its try_wait loop every ~35
        // cycles, and sixteen of them took ~40 % of the SM's issue slots from the warps doing the work (ncu source page, round 2)
        if (ew == 0) mbar_wait(&m->acc_full, n_use & 1);
        named_bar_sync(4, kEpiThreads);
        if (st0) stamp(2, it, 1 + 2 * L);
        tc_fence_after();
        const float* __restrict__ cwq = cw + th * 64;          // this thread's first quarter; the second one is 128 columns on
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {       // quarters th (slot x) then th + 2 (slot y)
          const int tq = th + 2 * pass;
          const uint32_t acc_q = lane_base + kColAcc + ((cs + pass) % 3) * 128 + th * 64;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int cc = pass * 128 + c * 32;       // compile-time offset of the chunk relative to column 64 th
            uint32_t v[32];
            tmem_ld32(acc_q + c * 32, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
            if (L == 2 && has_bias1) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b4 = s_bias1[(th * 64 + cc + i) >> 2];
                f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
              }
            }
            if (L == 0) {        // sigma head: fp32 dot of relu(h1) with sigma_net.1 row 0 (constant-bank operands, 4 accumulators)
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                s0 = fmaf(fmaxf(f[i], 0.f), cwq[cc + i], s0); s1 = fmaf(fmaxf(f[i + 1], 0.f), cwq[cc + i + 1], s1);
                s2 = fmaf(fmaxf(f[i + 2], 0.f), cwq[cc + i + 2], s2); s3 = fmaf(fmaxf(f[i + 3], 0.f), cwq[cc + i + 3], s3);
              }
              sig_part += (s0 + s1) + (s2 + s3);
            }
            if (L < 2) {
              uint32_t pk[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = pack_relu_bf16x2(f[2 * i], f[2 * i + 1]);
              if (!(a.ablate & 8)) tmem_st16(lane_base + kColA + ((th * 64 + cc) >> 1), pk);   // in place: the layer's MMAs are complete
            } else {             // rgb head: fp32 dot of relu(h3) with color_net.2 (constant-bank operands, 2 accumulators per channel)
              float r0 = 0.f, r1 = 0.f, g0 = 0.f, g1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float x0 = fmaxf(f[i], 0.f), x1 = fmaxf(f[i + 1], 0.f);
                r0 = fmaf(x0, cwq[256 + cc + i], r0); r1 = fmaf(x1, cwq[256 + cc + i + 1], r1);
                g0 = fmaf(x0, cwq[512 + cc + i], g0); g1 = fmaf(x1, cwq[512 + cc + i + 1], g1);
                b0 = fmaf(x0, cwq[768 + cc + i], b0); b1 = fmaf(x1, cwq[768 + cc + i + 1], b1);
              }
              rr += r0 + r1; rg_ += g0 + g1; rbl += b0 + b1;
            }
          }
          // quarter tq of the layer is done: its A' columns are written (and with quarter tq ^ 1 its accumulator slot is drained)
          if (L < 2) tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&m->hand[tq]);
        }
        ++n_use;
        cs += 2;
        if (st0) stamp(2, it, 2 + 2 * L);
      };
      layer(std::integral_constant<int, 0>{});
      if (FEAT) {
        // ---- geo layer: 128 fp32 columns of this row -> depth_feature [ray][r][128]; thread th streams columns 64 th .. 64 th + 63
        if (ew == 0) mbar_wait(&m->acc_full, n_use & 1);
        named_bar_sync(4, kEpiThreads);
        tc_fence_after();
        const uint32_t acc_g = lane_base + kColAcc + (cs % 3) * 128 + th * 64;
        float* gout = (r < S) ? a.feat + ((size_t)ray * S + r) * 128 + th * 64 : nullptr;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(acc_g + c * 32, v);
          tmem_ld_wait();
          if (gout) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(gout + c * 32 + i) =
                  make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          }
        }
        tc_fence_before();
        mbar_arrive(&m->hand[th]);               // the slot is drained; the A operand was not touched: both of this thread's quarters
        mbar_arrive(&m->hand[th + 2]);
        ++n_use;
        cs += 1;
      }
      layer(std::integral_constant<int, 1>{});
      layer(std::integral_constant<int, 2>{});
      m->headp[th][r][0] = rr; m->headp[th][r][1] = rg_; m->headp[th][r][2] = rbl; m->headp[th][r][3] = sig_part;
      named_bar_sync(2, kEpiThreads);               // all four quarters' head partials visible
      if (th == 0) {
        // ---- compositing (voxnerf.py:153-201), one thread per sample row ---------------------------------------------------
        const float* rb = a.ray_batch + ray * 11;
        const float d[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
        const float zv = slot->z[r];
        float col[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          col[i] = sigmoidf_((m->headp[0][r][i] + m->headp[1][r][i]) + (a.mlp.color2_b ? __ldg(a.mlp.color2_b + i) : 0.f));
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
        float t = 1.0f - alpha;                 // inclusive product scan of (1 - alpha) over the warp
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
      named_bar_sync(2, kEpiThreads);               // headp / red / wtot free for the next ray
      if (st0) stamp(2, it, 7);
      mbar_arrive(&m->slot_free[sl]);               // slot[sl] (z, bias operand) may be rewritten by the gather warps
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

// sigma / rgb head weights of the model -> [wsig 256 | r 256 | g 256 | b 256] fp32 (staged in global memory, then copied to a constant slot)
static __global__ void heads_stage_kernel(const float* __restrict__ sigma1_v, const float* __restrict__ color2_t, float* __restrict__ dst) {
  const int i = threadIdx.x;
  dst[i] = sigma1_v[i];
#pragma unroll
  for (int j = 0; j < 3; ++j) dst[256 + j * 256 + i] = color2_t[i * 4 + j];
}

int launch_fine_tc2(const FineArgs& a_in, int grid_dtype, const uint8_t* wblob, const uint8_t* wgeo, cudaStream_t st) {
  FineArgs a = a_in;
  if (int rc = bind_device("edn_render_fine_fwd")) return rc;      // `stage` below lives on the bound device
  static float* stage = nullptr;
  static unsigned launch_no = 0;
  if (!stage) EDN_CUDA_OK(cudaMalloc(&stage, sizeof(float) * kHeadSlots * kHeadFloats));
  const int hslot = (int)(launch_no++ % kHeadSlots);
  heads_stage_kernel<<<1, 256, 0, st>>>(a.mlp.sigma1_v, a.mlp.color2_t, stage + hslot * kHeadFloats);
  EDN_CUDA_OK(cudaMemcpyToSymbolAsync(c_heads, stage + hslot * kHeadFloats, sizeof(float) * kHeadFloats, sizeof(float) * kHeadFloats * hslot,
                                      cudaMemcpyDeviceToDevice, st));
  static const int ablate = [] { const char* e = getenv("EDN_TC2_ABLATE"); return e ? atoi(e) : 0; }();   // dev: TIMING-ONLY ablations (outputs invalid)
  a.ablate = ablate;
  const unsigned gx = (unsigned)(a.n_rays < (int64_t)num_sms() ? a.n_rays : (int64_t)num_sms());
  const char* tr = getenv("EDN_TC_TRACE");
  if (tr && tr[0] == '1' && grid_dtype == EDN_BF16) {   // dev tooling: print the role time lines of CTA 0 (synchronises!)
    long long* buf = nullptr;
    EDN_CUDA_OK(cudaMallocManaged(&buf, 4 * 4 * 16 * sizeof(long long)));
    memset(buf, 0, 4 * 4 * 16 * sizeof(long long));
    a.trace = buf;
    EDN_CUDA_OK(cudaFuncSetAttribute(fine_fwd_tc2_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    fine_fwd_tc2_kernel<__nv_bfloat16, false><<<gx, kThreads, kSmemBytes, st>>>(a, wblob, wgeo, hslot);
    EDN_CUDA_OK(cudaStreamSynchronize(st));
    const long long t0 = buf[16];   // MMA role, it = 8, slot 0
    static const char* role[3] = {"gather", "mma", "epi"};
    for (int r = 0; r < 3; ++r)
      for (int i = 0; i < 4; ++i) {
        fprintf(stderr, "[trace2 %s it=%d]", role[r], 8 + i);
        for (int k = 0; k < 8; ++k) if (buf[(r * 4 + i) * 16 + k]) fprintf(stderr, " s%d=%lld", k, buf[(r * 4 + i) * 16 + k] - t0);
        fprintf(stderr, "\n");
      }
    cudaFree(buf);
    return EDN_OK;
  }
  auto launch = [&](auto kern) -> int {
    EDN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    kern<<<gx, kThreads, kSmemBytes, st>>>(a, wblob, wgeo, hslot);
    return EDN_OK;
  };
  int rc;
  if (grid_dtype == EDN_BF16) rc = a.feat ? launch(fine_fwd_tc2_kernel<__nv_bfloat16, true>) : launch(fine_fwd_tc2_kernel<__nv_bfloat16, false>);
  else rc = a.feat ? launch(fine_fwd_tc2_kernel<float, true>) : launch(fine_fwd_tc2_kernel<float, false>);
  if (rc) return rc;
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

}  // namespace edn

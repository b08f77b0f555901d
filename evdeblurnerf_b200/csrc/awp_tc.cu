// AWP per-sample MLP on the tensor cores (tcgen05 + TMEM): sample_feature_embed_layer (128 -> 64 -> 64 -> 64 -> 64, ReLU) and
// MAM.linear (64 -> 32) of networks/dpnerf/awp.py:96-98 + mam.py:75 over all N*E*S samples -- the EDN_BF16 precision of edn_awp_fwd.
// Round 1 ran these five contractions as cuBLAS TF32 GEMMs with fp32 activations bouncing through HBM between them (1.4 ms on the
// headline batch); here one persistent kernel keeps a 128-sample tile's activations on chip:
//   row warps (4, one thread per sample row): coalesced load of the tile's depth_feature [128 x 128] fp32 -> bf16 UMMA A operand;
//       per layer: TMEM -> registers -> + bias, ReLU -> (fp32 copy to the workspace where the backward / the integration kernel
//       need it) -> bf16 -> A operand of the next layer;
//   MMA warp: all weights (44 KB bf16, UMMA K-major layout) resident in shared memory, M = 128, N = 64 / 32, fp32 accumulation.
// HBM-bound by construction (64 KB in, 48 KB out per tile; + 96 KB when the backward keeps every layer); two CTAs per SM overlap
// one tile's loads / epilogues with the other's MMAs.
#include <cstddef>

#include "awp_layout.cuh"
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_rows.cuh"

namespace edn {
namespace {

using namespace tc;

constexpr int kRows = 128;
constexpr int kThreads = 160;                        // 4 row warps + the MMA warp
constexpr int kABytes = 16 * kChunkA;                // K = 128
constexpr int kW0 = 0, kW1 = 16384, kW2 = 24576, kW3 = 32768, kWm = 40960, kWBytes = 45056;
constexpr uint32_t kTmemCols = 64;
constexpr int kBlobSlots = 8;

struct Misc {
  uint64_t a_full, acc_full, bar_w;
  uint32_t tmem_base, pad[3];
  alignas(16) float bias[4][64];
  alignas(16) float bias_m[32];
};
constexpr int kSmemBytes = kABytes + kWBytes + (int)sizeof(Misc);

struct TcArgs {
  const float* feat;      // [M][128]
  float* act[4];          // [M][64] each (act[0..2] only when keep)
  float* xl;              // [M][32]
  const float* b[4];
  const float* bm;
  int64_t M;
  int keep;
};

// fp32 [K][N] (in-major, out contiguous) -> bf16 UMMA K-major: (k / 16) * (N * 16) + ((k % 16) / 8) * (N * 8) + n * 8 + k % 8
__global__ void awp_pack_kernel(const float* w0, const float* w1, const float* w2, const float* w3, const float* wm, __nv_bfloat16* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* src; int K, N, off, j = i;
  if (j < 128 * 64) { src = w0; K = 128; N = 64; off = kW0; }
  else if ((j -= 128 * 64) < 64 * 64) { src = w1; K = 64; N = 64; off = kW1; }
  else if ((j -= 64 * 64) < 64 * 64) { src = w2; K = 64; N = 64; off = kW2; }
  else if ((j -= 64 * 64) < 64 * 64) { src = w3; K = 64; N = 64; off = kW3; }
  else if ((j -= 64 * 64) < 64 * 32) { src = wm; K = 64; N = 32; off = kWm; }
  else return;
  (void)K;
  const int k = j / N, n = j - k * N;
  dst[off / 2 + (k / 16) * (N * 16) + ((k % 16) / 8) * (N * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(src[(size_t)k * N + n]);
}

__global__ void __launch_bounds__(kThreads, 2) awp_sample_tc_kernel(const TcArgs a, const uint8_t* __restrict__ wblob) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;
  uint8_t* Wsm = smem + kABytes;
  Misc* m = reinterpret_cast<Misc*>(Wsm + kWBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (a.M + kRows - 1) / kRows;
  const int64_t n_my = (n_tiles > (int64_t)blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    mbar_init(&m->a_full, kRows); mbar_init(&m->acc_full, 1); mbar_init(&m->bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&m->tmem_base, kTmemCols);
  for (int i = tid; i < 4 * 64; i += kThreads) m->bias[i >> 6][i & 63] = __ldg(a.b[i >> 6] + (i & 63));
  if (tid < 32) m->bias_m[tid] = __ldg(a.bm + tid);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = m->tmem_base;

  if (warp == 4) {
    // =================================== MMA issuer (converged warp, one elected lane) =====================================
    if (n_my > 0) {
      if (elect_one()) { mbar_expect_tx(&m->bar_w, kWBytes); bulk_g2s(Wsm, wblob, kWBytes, &m->bar_w); }
      __syncwarp();
      mbar_wait(&m->bar_w, 0);
      const uint32_t a0 = smem_u32(As), wb = smem_u32(Wsm);
      uint32_t ph = 0;
      auto stage = [&](uint32_t w_off, int ksteps, int n) {
        mbar_wait(&m->a_full, ph); ph ^= 1;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc = make_idesc_bf16(128, n);
          for (int j = 0; j < ksteps; ++j)
            mma_bf16_ss(tmem, make_smem_desc(a0 + 2 * j * kChunkA, kChunkA, 128), make_smem_desc(wb + w_off + j * n * 32, (uint32_t)n * 16u, 128), idesc, j > 0);
          mma_commit(&m->acc_full);
        }
        __syncwarp();
      };
      for (int64_t it = 0; it < n_my; ++it) {
        stage(kW0, 8, 64);
        stage(kW1, 4, 64);
        stage(kW2, 4, 64);
        stage(kW3, 4, 64);
        stage(kWm, 4, 32);
      }
    }
    __syncwarp();
  } else {
    // =================================== row warps: one thread per sample row ==============================================
    const int r = warp * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t pacc = 0;
    for (int64_t it = 0; it < n_my; ++it) {
      const int64_t row0 = ((int64_t)blockIdx.x + it * gridDim.x) * kRows;
      // ---- depth_feature tile -> bf16 A operand: a warp reads one 512-byte row per instruction (coalesced) --------------------
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const int row = warp * 32 + rr;
        const int64_t gr = row0 + row;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < a.M) v = __ldg(reinterpret_cast<const float4*>(a.feat + gr * 128) + lane);
        const uint32_t p0 = pack_bf16x2(v.x, v.y), p1 = pack_bf16x2(v.z, v.w);
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(smem_u32(As + (lane >> 1) * kChunkA + row * 16 + (lane & 1) * 8)), "r"(p0), "r"(p1) : "memory");
      }
      fence_proxy_async_smem();
      mbar_arrive(&m->a_full);
      const int64_t gr = row0 + r;
      const bool live = gr < a.M;
#pragma unroll 1
      for (int L = 0; L < 4; ++L) {
        mbar_wait(&m->acc_full, pacc); pacc ^= 1;
        tc_fence_after();
        float* out = ((L == 3 || a.keep) && live) ? a.act[L] + gr * 64 : nullptr;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(&m->bias[L][c * 32 + i]);
            f[i] = fmaxf(__uint_as_float(v[i]) + b4.x, 0.f); f[i + 1] = fmaxf(__uint_as_float(v[i + 1]) + b4.y, 0.f);
            f[i + 2] = fmaxf(__uint_as_float(v[i + 2]) + b4.z, 0.f); f[i + 3] = fmaxf(__uint_as_float(v[i + 3]) + b4.w, 0.f);
          }
          if (out) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(out + c * 32 + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(As + (c * 4 + j) * kChunkA + r * 16, pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                         pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&m->a_full);
      }
      // ---- MAM.linear: 64 -> 32, + bias, no activation -> xl ----------------------------------------------------------------
      mbar_wait(&m->acc_full, pacc); pacc ^= 1;
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        if (live) {
          float* out = a.xl + gr * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(&m->bias_m[i]);
            *reinterpret_cast<float4*>(out + i) = make_float4(__uint_as_float(v[i]) + b4.x, __uint_as_float(v[i + 1]) + b4.y,
                                                              __uint_as_float(v[i + 2]) + b4.z, __uint_as_float(v[i + 3]) + b4.w);
          }
        }
      }
      tc_fence_before();       // the next tile's first MMA overwrites these TMEM columns only after a_full, which follows this load
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

// The five per-sample contractions of the AWP forward on tcgen05 (bf16 operands, fp32 accumulation): fills ws.act[3], ws.xl
// (bias included) and, with keep != 0, ws.act[0..2] for the backward.
int awp_sample_mlp_tc(const edn_awp_params* p, const float* depth_feature, int64_t M, const AwpWs& ws, int keep, cudaStream_t st) {
  if (int rc = bind_device("edn_awp_fwd")) return rc;               // `blobs` below lives on the bound device
  static uint8_t* blobs = nullptr;
  static unsigned launch_no = 0;
  if (!blobs) EDN_CUDA_OK(cudaMalloc(&blobs, (size_t)kBlobSlots * kWBytes));
  uint8_t* blob = blobs + (size_t)(launch_no++ % kBlobSlots) * kWBytes;       // rotating slots: launches in flight keep their own weights
  awp_pack_kernel<<<(128 * 64 + 3 * 64 * 64 + 64 * 32 + 255) / 256, 256, 0, st>>>(p->sample_t[0], p->sample_t[1], p->sample_t[2], p->sample_t[3],
                                                                                 p->mam_linear_t, reinterpret_cast<__nv_bfloat16*>(blob));
  TcArgs a{};
  a.feat = depth_feature;
  for (int l = 0; l < 4; ++l) { a.act[l] = ws.act[l]; a.b[l] = p->sample_b[l]; }
  a.xl = ws.xl; a.bm = p->mam_linear_b; a.M = M; a.keep = keep;
  const int64_t n_tiles = (M + kRows - 1) / kRows;
  const int64_t cap = 2 * (int64_t)num_sms();
  const unsigned gx = (unsigned)(n_tiles < cap ? n_tiles : cap);
  EDN_CUDA_OK(cudaFuncSetAttribute(awp_sample_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  awp_sample_tc_kernel<<<gx, kThreads, kSmemBytes, st>>>(a, blob);
  EDN_CUDA_OK(cudaGetLastError());
  return EDN_OK;
}

}  // namespace edn

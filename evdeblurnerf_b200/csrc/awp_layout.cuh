// Workspace layout shared by the AWP forward (awp.cu) and backward (awp_bwd.cu).
#pragma once
#include <stdint.h>

#include "../../include/evdeblur_b200.h"

namespace edn {

struct AwpWs {
  float* gint;    // [NE][64]      integrated sample features (awp.py:49-77)
  float* inter;   // [NE][32]      softmax-over-samples pooled curves (mam.py:34)
  float* xl;      // [NE][S][32]   MAM.linear output ("curves")
  float* att;     // [NE][S]       attention logits
  float* x;       // [N][E][32]    motion features
  float* y;       // [N][E][32]    convd output before BatchNorm
  double* stats;  // [32][2] + 2   BatchNorm batch sums; stats[64] = rows behind them (all-reduced together with the sums)
  double* bn_part; // [64][64]     per-block partial sums of the two-level (deterministic) reduction behind `stats`
  float* act[4];  // [NE*S][64]    post-ReLU activations of sample_feature_embed_layer.l (GEMM path only)
};

inline int64_t awp_ws_floats(int64_t N, int E, int S, bool gemm) {
  const int64_t NE = N * E;
  return NE * 64 + NE * 32 + NE * S * 32 + NE * S + 2 * NE * 32 + 2 * 66 + 2 * 64 * 64 + 8 /* alignment slack */ + (gemm ? 4 * NE * S * 64 : 0);
}

inline AwpWs awp_ws_carve(float* w, int64_t N, int E, int S, bool gemm) {
  const int64_t NE = N * E;
  AwpWs a{};
  a.gint = w; w += NE * 64;
  a.inter = w; w += NE * 32;
  a.xl = w; w += NE * S * 32;
  a.att = w; w += (NE * S + 3) / 4 * 4;
  a.x = w; w += NE * 32;
  a.y = w; w += NE * 32;
  a.stats = reinterpret_cast<double*>(w + (((uintptr_t)w & 7) ? 1 : 0));
  a.bn_part = a.stats + 66;
  w += 2 * 66 + 2 + 2 * 64 * 64;
  w += (4 - ((uintptr_t)w / 4) % 4) % 4;            // 16-byte align the GEMM operands
  if (gemm) for (int l = 0; l < 4; ++l) { a.act[l] = w; w += NE * S * 64; }
  return a;
}

// awp.cu: the forward into `workspace` (gemm_path = materialised per-sample MLP, needed by the backward; tf32 = tensor-core
// GEMMs; phase / bn_rows as in edn_awp_options)
int awp_forward(const edn_awp_params* p, const float* depth_feature, const float* z_vals, const float* rays_d, int32_t rays_d_stride,
                const float* view_feature, int64_t n_rays, int32_t n_exposure, int32_t n_samples, float bn_eps, bool gemm_path,
                bool tf32, int phase, int64_t bn_rows_total, float* workspace, float* ccw, void* stream, bool keep_all = true);

// awp_tc.cu: the per-sample MLP + MAM.linear on tcgen05 (bf16 operands): fills ws.act[3], ws.xl and, with keep != 0, ws.act[0..2]
int awp_sample_mlp_tc(const edn_awp_params* p, const float* depth_feature, int64_t M, const AwpWs& ws, int keep, cudaStream_t st);

// Backward scratch that precedes the large buffers (awp_bwd.cu): ccw_tmp [NE], d_yn [NE][32], d_x [NE][32], bn_sums [64] doubles.
struct AwpBwdHead { float* ccw_tmp; float* d_yn; float* d_x; double* bn_sums; float* next; };
inline AwpBwdHead awp_bwd_head(float* workspace, int64_t N, int E, int S) {
  const int64_t NE = N * E;
  float* base = workspace + awp_ws_floats(N, E, S, true);
  base += (4 - ((uintptr_t)base / 4) % 4) % 4;
  auto take = [&](int64_t n) { float* q = base; base += (n + 3) / 4 * 4; return q; };
  AwpBwdHead h{};
  h.ccw_tmp = take(NE);
  h.d_yn = take(NE * 32);
  h.d_x = take(NE * 32);
  h.bn_sums = reinterpret_cast<double*>(take(2 * 64));
  h.next = base;
  return h;
}

}  // namespace edn

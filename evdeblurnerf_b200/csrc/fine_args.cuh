// Launch arguments shared by the fp32 (fine_f32.cu) and tcgen05 (fine_tc.cu) fine-pass kernels.
#pragma once
#include "common.cuh"

namespace edn {

struct FineArgs {
  GridDev gc, gf;
  edn_field_mlp mlp;
  const float* ray_batch;
  const float* z_vals;
  const float* noise;
  int64_t n_rays;
  int S;
  int flags;
  float rmnearplane;
  float* weights;
  float* rgb;
  float* depth;
  float* acc;
  float* feat;
  long long* trace;   // dev tooling (EDN_TC_TRACE=1): per-phase clock64 stamps of CTA 0, NULL otherwise
  int ablate;         // dev tooling (EDN_TC_ABLATE=bits): TIMING-ONLY ablations, results are garbage: 1 no VM gather, 2 no layer
                      // epilogues, 4 no weight stream (the MMAs read stale shared memory), 8 no PE / view-bias phase
};

// fine_tc2.cu: second-generation tensor-core fine kernel (lean schedule, rays of <= 128 samples); wblob = the lean section of the
// fine tensor-core blob ([layer][K-step][256 x 16] bf16, edn_pack_fine_tc)
int launch_fine_tc2(const FineArgs& a, int grid_dtype, const uint8_t* wblob, const uint8_t* wgeo, cudaStream_t st);

// fine_tc3.cu: bf16 x 3 tensor-core parity mode (EDN_TC32) and its blob section packer (scratch: 128 KB)
int launch_fine_tc3(const FineArgs& a, int grid_dtype, const uint8_t* wblob, cudaStream_t st);
int pack_fine_tc3_section(const float* f1, const float* f23, const float* color1_t, uint8_t* scratch, uint8_t* dst, cudaStream_t st);
int64_t fine_tc3_blob_offset();

}  // namespace edn

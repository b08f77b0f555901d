// tcgen05 tensor-core fine pass (placeholder until the UMMA pipeline lands; fails loudly, never falls back).
#include "common.cuh"
namespace edn {
struct FineArgs;
int launch_fine_tc(const FineArgs&, int, cudaStream_t) {
  set_error("edn_render_fine_fwd: EDN_BF16 (tcgen05) precision is not built in this library");
  return EDN_E_UNSUPPORTED;
}
}  // namespace edn

// K1 tensor-core modes (placeholder until the tcgen05 kernel lands in this file).
#include "rcb_common.cuh"
namespace rcb {
size_t build_tc_workspace_bytes(int, int, int, int, int) { return 0; }
int launch_build_tc(const float*, const float*, void* const*, const rcb_pyramid_layout&, int, int, int, int, int,
                    void*, size_t, cudaStream_t) {
  return RCB_ERR_UNSUPPORTED;
}
}  // namespace rcb

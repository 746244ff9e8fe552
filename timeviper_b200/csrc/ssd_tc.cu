// tcgen05 / TMEM / TMA SSD kernel family -- placeholder until the fused kernel lands.
#include "common.cuh"
#include "ssd.h"

namespace tv {
bool tc_supported(const tv_ssd_params&) { return false; }
size_t tc_workspace_bytes(const tv_ssd_params&) { return 0; }
int ssd_tc_forward(const tv_ssd_params&, void*, cudaStream_t) {
  set_error("ssd(tcgen05): not built");
  return TV_ERR_UNSUPPORTED;
}
}  // namespace tv

// Internal interface between the SSD dispatcher (api.cu) and its two kernel families.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/timeviper_b200.h"

namespace tv {

struct SimtWorkspace {
  size_t dt_off, cs_off, states_off, cb_off, total;
};
SimtWorkspace simt_workspace_layout(const tv_ssd_params& p);
int simt_supported(const tv_ssd_params& p);
int ssd_simt_forward(const tv_ssd_params& p, void* workspace, cudaStream_t s);
// stage (i): dt activation + per-chunk cumsum into (b, nchunks, H, Q) fp32 arrays (shared by both families)
int launch_dt_cumsum(const tv_ssd_params& p, float* dt_act, float* cs, cudaStream_t s);

// tcgen05 / TMEM / TMA family (ssd_tc.cu): bf16, P=80, N=128, Q=128
bool tc_supported(const tv_ssd_params& p);
size_t tc_workspace_bytes(const tv_ssd_params& p);
int ssd_tc_forward(const tv_ssd_params& p, void* workspace, cudaStream_t s);
void set_trace_buffer(void* p);
void set_ablate(int mask);

constexpr size_t kMaxDynSmem = 232448;  // 227 KB per CTA on sm_100

}  // namespace tv

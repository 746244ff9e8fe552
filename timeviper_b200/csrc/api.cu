// C-ABI plumbing: error reporting, SSD argument validation and kernel-family dispatch.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "ssd.h"

namespace tv {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int validate_ssd(const tv_ssd_params* p) {
  TV_CHECK_ARG(p != nullptr, "ssd: null params");
  TV_CHECK_ARG(p->dtype == TV_F32 || p->dtype == TV_BF16, "ssd: dtype %d", p->dtype);
  TV_CHECK_ARG(p->mode == TV_SSD_FULL || p->mode == TV_SSD_STATE_ONLY || p->mode == TV_SSD_DT_ONLY, "ssd: mode %d",
               p->mode);
  TV_CHECK_ARG(p->batch > 0 && p->seqlen > 0 && p->nheads > 0 && p->headdim > 0 && p->ngroups > 0 &&
                   p->dstate > 0 && p->chunk_size > 0,
               "ssd: empty problem (b=%d L=%d H=%d P=%d G=%d N=%d Q=%d)", p->batch, p->seqlen, p->nheads,
               p->headdim, p->ngroups, p->dstate, p->chunk_size);
  TV_CHECK_ARG(p->nheads % p->ngroups == 0, "ssd: nheads %d %% ngroups %d != 0", p->nheads, p->ngroups);
  TV_CHECK_ARG(p->x && p->dt && p->A && p->B, "ssd: x, dt, A, B must be non-null");
  if (p->mode == TV_SSD_FULL) TV_CHECK_ARG(p->C && p->out, "ssd: C and out must be non-null");
  else if (p->mode == TV_SSD_STATE_ONLY)
    TV_CHECK_ARG(p->final_states != nullptr, "ssd: state-only mode needs final_states");
  return TV_OK;
}

}  // namespace tv

extern "C" int tv_abi_version(void) { return TV_ABI_VERSION; }
extern "C" const char* tv_last_error(void) { return tv::g_err; }

extern "C" int tv_ssd_kernel_family(const tv_ssd_params* p) {
  if (p == nullptr) return 0;
  return (!p->force_simt && tv::tc_supported(*p)) ? 1 : 0;
}

extern "C" size_t tv_ssd_workspace_bytes(const tv_ssd_params* p) {
  if (p == nullptr || p->chunk_size <= 0) return 0;
  if (tv_ssd_kernel_family(p) == 1) return tv::tc_workspace_bytes(*p);
  return tv::simt_workspace_layout(*p).total;
}

extern "C" int tv_ssd_chunk_scan_fwd(const tv_ssd_params* p, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  using namespace tv;
  int rc = validate_ssd(p);
  if (rc != TV_OK) return rc;
  const size_t need = tv_ssd_workspace_bytes(p);
  if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
    set_error("ssd: workspace %zu bytes < required %zu", workspace_bytes, need);
    return TV_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (tv_ssd_kernel_family(p) == 1) return ssd_tc_forward(*p, workspace, s);
  if (p->mode == TV_SSD_DT_ONLY) return TV_OK;   // the CUDA-core family recomputes dt/cumsum in every call
  rc = simt_supported(*p);
  if (rc != TV_OK) return rc;
  return ssd_simt_forward(*p, workspace, s);
}

// Debug hook (not part of the reference-facing ABI): per-chunk clock64 stamps of CTA (0,0) of the fused SSD kernel.
extern "C" void tv_debug_set_trace(void* device_buffer) { tv::set_trace_buffer(device_buffer); }
extern "C" void tv_debug_set_ablate(int mask) { tv::set_ablate(mask); }
// Number of kernels this library has enqueued so far in this process (all entry points, all streams).
extern "C" unsigned long long tv_debug_launch_count(void) { return tv::launches_so_far(); }

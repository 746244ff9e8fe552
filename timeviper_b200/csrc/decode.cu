// Single-token decode step of the Mamba-2 mixer for sm_100a.
//
// Replaces causal_conv1d_update and selective_state_update at the reference call sites
// timeviper/model/llm/llm_repo/nano/modeling_nano.py:495-501 and :528-539; they consume the conv state (b, conv_dim, K)
// and the fp32 SSM state (b, H, P, N) that the prefill path leaves in the cache.  Both are latency/launch-bound at
// batch 1 (the SSM state is 5.24 MB read + written per token and layer); the kernels are plain coalesced streaming code.
#include "common.cuh"

namespace tv {

// one thread per (batch, channel): the state row (state_len elements) is read, shifted and written back
template <typename T, bool SILU>
__global__ void __launch_bounds__(256)
conv1d_update_kernel(const T* __restrict__ x, T* __restrict__ state, const T* __restrict__ weight,
                     const T* __restrict__ bias, T* __restrict__ out, int dim, int width, int state_len, int64_t xbs,
                     int64_t obs, int64_t sbs, int64_t sds) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= dim) return;
  T* row = state + (int64_t)b * sbs + (int64_t)c * sds;
  const T xn = x[(int64_t)b * xbs + c];
  float acc = bias != nullptr ? to_f32<T>(bias[c]) : 0.f;
  for (int j = 0; j < state_len; ++j) {
    const T v = j + 1 < state_len ? row[j + 1] : xn;
    row[j] = v;
    const int k = j - (state_len - width);
    if (k >= 0) acc = fmaf(to_f32<T>(weight[(int64_t)c * width + k]), to_f32<T>(v), acc);
  }
  if (SILU) acc = acc / (1.0f + expf(-acc));
  out[(int64_t)b * obs + c] = from_f32<T>(acc);
}

// one warp per (batch, head, p): lanes stride over the state dim, so state / B / C accesses are coalesced
template <typename T, typename S>
__global__ void __launch_bounds__(256)
ssu_kernel(const tv_ssu_params p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t rows = (int64_t)p.batch * p.nheads * p.headdim;
  if (row >= rows) return;
  const int pd = (int)(row % p.headdim);
  const int h = (int)((row / p.headdim) % p.nheads);
  const int b = (int)(row / ((int64_t)p.headdim * p.nheads));
  const int g = h / (p.nheads / p.ngroups);
  const T* x = (const T*)p.x; const T* dtp = (const T*)p.dt; const T* Bp = (const T*)p.B; const T* Cp = (const T*)p.C;
  const T* zp = (const T*)p.z;
  S* st = (S*)p.state + row * p.dstate;
  const float xv = to_f32<T>(x[b * p.x_batch_stride + h * p.x_head_stride + pd * p.x_dim_stride]);
  float dt = to_f32<T>(dtp[b * p.dt_batch_stride + h * p.dt_head_stride + pd * p.dt_dim_stride]);
  if (p.dt_bias != nullptr) dt += p.dt_bias[h * p.bias_head_stride + pd * p.bias_dim_stride];
  if (p.dt_softplus && dt <= 20.f) dt = log1pf(expf(dt));
  dt = fminf(fmaxf(dt, p.dt_min), p.dt_max);
  const float* Arow = p.A + h * p.a_head_stride + pd * p.a_dim_stride;
  const T* Brow = Bp + b * p.b_batch_stride + g * p.b_group_stride;
  const T* Crow = Cp + b * p.c_batch_stride + g * p.c_group_stride;
  float acc = 0.f;
  for (int n = lane; n < p.dstate; n += 32) {
    const float dA = expf(dt * Arow[n * p.a_state_stride]);
    const float s = fmaf(to_f32<S>(st[n]), dA, dt * to_f32<T>(Brow[n]) * xv);
    st[n] = from_f32<S>(s);
    acc = fmaf(s, to_f32<T>(Crow[n]), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (p.D != nullptr) acc = fmaf(xv, p.D[h * p.d_head_stride + pd * p.d_dim_stride], acc);
    if (zp != nullptr) {
      const float zv = to_f32<T>(zp[b * p.z_batch_stride + h * p.z_head_stride + pd * p.z_dim_stride]);
      acc *= zv / (1.0f + expf(-zv));
    }
    ((T*)p.out)[row] = from_f32<T>(acc);
  }
}

template <typename T>
static int launch_conv_update(const tv_conv1d_update_params& p, cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(p.dim, 256), p.batch);
  auto k = p.silu ? conv1d_update_kernel<T, true> : conv1d_update_kernel<T, false>;
  k<<<grid, 256, 0, s>>>((const T*)p.x, (T*)p.conv_state, (const T*)p.weight, (const T*)p.bias, (T*)p.out, p.dim, p.width,
                         p.state_len, p.x_batch_stride, p.out_batch_stride, p.state_batch_stride, p.state_dim_stride);
  TV_LAUNCH_OK();
  return TV_OK;
}

}  // namespace tv

extern "C" int tv_causal_conv1d_update(const tv_conv1d_update_params* p, void* stream) {
  using namespace tv;
  TV_CHECK_ARG(p != nullptr, "causal_conv1d_update: null params");
  TV_CHECK_ARG(p->x && p->conv_state && p->weight && p->out, "causal_conv1d_update: x, conv_state, weight, out must be non-null");
  TV_CHECK_ARG(p->batch > 0 && p->dim > 0 && p->width > 0 && p->state_len >= p->width,
               "causal_conv1d_update: bad sizes (b=%d dim=%d width=%d state_len=%d)", p->batch, p->dim, p->width, p->state_len);
  TV_CHECK_ARG(p->dtype == TV_F32 || p->dtype == TV_BF16, "causal_conv1d_update: dtype %d", p->dtype);
  cudaStream_t s = (cudaStream_t)stream;
  return p->dtype == TV_BF16 ? launch_conv_update<__nv_bfloat16>(*p, s) : launch_conv_update<float>(*p, s);
}

extern "C" int tv_selective_state_update(const tv_ssu_params* p, void* stream) {
  using namespace tv;
  TV_CHECK_ARG(p != nullptr, "selective_state_update: null params");
  TV_CHECK_ARG(p->state && p->x && p->dt && p->A && p->B && p->C && p->out, "selective_state_update: null tensor");
  TV_CHECK_ARG(p->batch > 0 && p->nheads > 0 && p->headdim > 0 && p->ngroups > 0 && p->dstate > 0 &&
                   p->nheads % p->ngroups == 0,
               "selective_state_update: bad sizes (b=%d H=%d P=%d G=%d N=%d)", p->batch, p->nheads, p->headdim, p->ngroups, p->dstate);
  TV_CHECK_ARG((p->dtype == TV_F32 || p->dtype == TV_BF16) && (p->state_dtype == TV_F32 || p->state_dtype == TV_BF16),
               "selective_state_update: dtype %d / state dtype %d", p->dtype, p->state_dtype);
  const int64_t rows = (int64_t)p->batch * p->nheads * p->headdim;
  dim3 grid((unsigned)ceil_div(rows, 8));
  cudaStream_t s = (cudaStream_t)stream;
  if (p->dtype == TV_BF16) {
    if (p->state_dtype == TV_F32) ssu_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>(*p);
    else ssu_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(*p);
  } else {
    if (p->state_dtype == TV_F32) ssu_kernel<float, float><<<grid, 256, 0, s>>>(*p);
    else ssu_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>(*p);
  }
  TV_LAUNCH_OK();
  return TV_OK;
}

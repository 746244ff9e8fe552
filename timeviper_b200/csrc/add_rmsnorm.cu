// Residual add + RMSNorm of the hybrid stack's blocks for sm_100a.
//
// Replaces, per block of the reference's layer loop, `residual + hidden_states` (NemotronHBlock.forward,
// timeviper/model/llm/llm_repo/nano/modeling_nano.py:965) followed by the NEXT block's pre-norm NemotronHRMSNorm
// (:888-904, called at :941):  s = bf16(residual + delta);  out = dtype(w * (s * rsqrt(mean(s^2) + eps))),
// statistics and the weight multiply in fp32 -- the same rounding points as the eager torch code (the sum is rounded to
// the activation dtype before it is normalised, because the reference materialises it in that dtype).
//
// Roofline: HBM streaming.  Eager torch runs the norm as ~7 elementwise passes over an fp32 copy (~14 GB per call at
// 81,920 tokens x 4480) plus a 3-tensor pass for the add; this kernel reads delta and residual once and writes the sum and
// the normed row once (4 x 0.73 GB), or 2 x 0.73 GB without a residual.
// One CTA of 128 threads per row; the row stays in registers between the sum of squares and the scale (16-byte pieces,
// up to 10 per thread: d <= 10240 in bf16, 5120 in fp32).
#include "common.cuh"

namespace tv {

constexpr int ARN_THREADS = 128, ARN_MAXV = 10;

template <typename T, bool HAS_RES>
__global__ void __launch_bounds__(ARN_THREADS)
add_rmsnorm_kernel(const T* __restrict__ x, const T* __restrict__ res, const T* __restrict__ w, T* __restrict__ sum_out,
                   T* __restrict__ out, int d, int64_t xrs, int64_t rrs, int64_t srs, int64_t ors, float eps) {
  constexpr int V = Vec16<T>::N;
  const int64_t row = blockIdx.x;
  const int nvec = d / V;
  const T* xp = x + row * xrs;
  const T* rp = HAS_RES ? res + row * rrs : nullptr;
  float held[ARN_MAXV][V];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < ARN_MAXV; ++i) {
    const int vi = threadIdx.x + ARN_THREADS * i;
    if (vi < nvec) {
      load16<T>(xp + vi * V, held[i]);
      if (HAS_RES) {
        float r[V];
        load16<T>(rp + vi * V, r);
#pragma unroll
        for (int j = 0; j < V; ++j) held[i][j] += r[j];
        if (sizeof(T) == 2) {                        // the reference holds the sum in the activation dtype
#pragma unroll
          for (int j = 0; j < V; ++j) held[i][j] = __bfloat162float(__float2bfloat16_rn(held[i][j]));
        }
        if (sum_out != nullptr) store16<T>(sum_out + row * srs + vi * V, held[i]);
      }
#pragma unroll
      for (int j = 0; j < V; ++j) ss = fmaf(held[i][j], held[i][j], ss);
    }
  }
  __shared__ float part[ARN_THREADS / 32];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
  __syncthreads();
  ss = part[0] + part[1] + part[2] + part[3];
  const float rstd = rsqrtf(ss / (float)d + eps);
#pragma unroll
  for (int i = 0; i < ARN_MAXV; ++i) {
    const int vi = threadIdx.x + ARN_THREADS * i;
    if (vi < nvec) {
      float ww[V], o[V];
      load16<T>(w + vi * V, ww);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = ww[j] * (held[i][j] * rstd);
      store16<T>(out + row * ors + vi * V, o);
    }
  }
}

template <typename T>
static int launch_add_rmsnorm(const tv_add_rmsnorm_params& p, cudaStream_t s) {
  dim3 grid((unsigned)p.rows);
  if (p.residual != nullptr)
    add_rmsnorm_kernel<T, true><<<grid, ARN_THREADS, 0, s>>>((const T*)p.x, (const T*)p.residual, (const T*)p.weight,
                                                             (T*)p.sum_out, (T*)p.out, p.d, p.x_row_stride, p.res_row_stride,
                                                             p.sum_row_stride, p.out_row_stride, p.eps);
  else
    add_rmsnorm_kernel<T, false><<<grid, ARN_THREADS, 0, s>>>((const T*)p.x, nullptr, (const T*)p.weight, nullptr, (T*)p.out, p.d,
                                                              p.x_row_stride, 0, 0, p.out_row_stride, p.eps);
  TV_LAUNCH_OK();
  return TV_OK;
}

}  // namespace tv

extern "C" int tv_add_rmsnorm_fwd(const tv_add_rmsnorm_params* p, void* stream) {
  using namespace tv;
  TV_CHECK_ARG(p != nullptr, "add_rmsnorm: null params");
  TV_CHECK_ARG(p->x && p->weight && p->out, "add_rmsnorm: x, weight and out must be non-null");
  TV_CHECK_ARG(p->rows > 0 && p->d > 0 && p->rows < (1ll << 31), "add_rmsnorm: empty or oversized problem (rows=%lld d=%d)",
               (long long)p->rows, p->d);
  TV_CHECK_ARG(p->dtype == TV_F32 || p->dtype == TV_BF16, "add_rmsnorm: dtype %d", p->dtype);
  const int V = p->dtype == TV_BF16 ? 8 : 4;
  TV_CHECK_ARG(p->d % V == 0, "add_rmsnorm: d %d must be a multiple of %d", p->d, V);
  if (p->d > ARN_THREADS * ARN_MAXV * V) {
    set_error("add_rmsnorm: d %d > %d unsupported", p->d, ARN_THREADS * ARN_MAXV * V);
    return TV_ERR_UNSUPPORTED;
  }
  TV_CHECK_ARG(p->x_row_stride % V == 0 && p->out_row_stride % V == 0 &&
                   (p->residual == nullptr || p->res_row_stride % V == 0) && (p->sum_out == nullptr || p->sum_row_stride % V == 0),
               "add_rmsnorm: row strides must be multiples of %d elements (16 bytes)", V);
  TV_CHECK_ARG(((uintptr_t)p->x % 16 == 0) && ((uintptr_t)p->out % 16 == 0) && ((uintptr_t)p->weight % 16 == 0) &&
                   ((uintptr_t)p->residual % 16 == 0) && ((uintptr_t)p->sum_out % 16 == 0),
               "add_rmsnorm: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  return p->dtype == TV_BF16 ? launch_add_rmsnorm<__nv_bfloat16>(*p, s) : launch_add_rmsnorm<float>(*p, s);
}

// Depthwise causal conv1d (+bias, +SiLU), channel-last, for sm_100a.
//
// Replaces causal_conv1d_fn (causal_conv1d 1.5.2) at the reference call site
// timeviper/model/llm/llm_repo/nano/modeling_nano.py:619-624.  The input there is a strided channel-last
// VIEW into the in_proj output (row stride 22656 elements, base offset 10240): the kernel honours
// arbitrary batch/row strides and only requires unit channel stride and 16-byte alignment.
//
// Roofline: pure HBM streaming, 2 * dim * sizeof(T) bytes per token (49,152 B at dim 12288, bf16).
// Layout/tiling: one thread owns 16 bytes of channels (8 bf16 / 4 fp32) and walks TOK consecutive tokens with the
// K-1 previous rows kept in registers, so every x element is read once per CTA (halo re-read = (K-1)/tok, served
// from L2); a warp reads/writes 512 contiguous bytes per token row.  Input rows are prefetched with cp.async (16 bytes
// per thread) into a per-thread ring in shared memory, so the bytes in flight are not bounded by registers.
#include "common.cuh"

namespace tv {

#ifndef TV_CONV_THREADS
#define TV_CONV_THREADS 32     // one warp per CTA: 16 CTAs/SM by registers; measured 81.5 % vs 80.2 % with 128-thread CTAs
#endif
constexpr int CONV_THREADS = TV_CONV_THREADS;
#ifndef TV_CONV_RING
#define TV_CONV_RING 16         // depth of the per-thread cp.async row ring in shared memory
#endif
#ifndef TV_CONV_MINWARPS
#define TV_CONV_MINWARPS 16     // resident warps per SM the register allocation must allow
#endif
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// Tokens per CTA along the sequence.  Measured at 128K tokens with the cp.async ring (% of the HBM copy peak):
// 32 -> 94.0, 48 -> 99.5, 64 -> 98.4, 128 -> 97.1, 256 -> 96.1, 512 -> 94.1; at 16K tokens 48 and 64 tie (87.5).
#ifndef TV_CONV_TOK
#define TV_CONV_TOK 48
#endif
constexpr int CONV_TOK = TV_CONV_TOK;
// channels per thread = one 16-byte access: 8 (bf16) / 4 (fp32); a warp covers 512 contiguous bytes per token row

template <typename T> struct Raw4 {   // register image of one 16-byte access
  uint4 r;
  __device__ __forceinline__ void load(const T* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float (&v)[Vec16<T>::N]) const { unpack16<T>(r, v); }
  static __device__ __forceinline__ void store(T* p, const float (&v)[Vec16<T>::N]) { store16<T>(p, v); }
};

// 16 resident warps per SM, each thread keeping TV_CONV_RING = 16 row loads in flight through its shared-memory ring:
// 16 warps * 16 rows * 512 B = 128 KB in flight (and of shared memory) per SM.  Measured at 128K tokens: 97 % of the HBM
// peak (the register-prefetch version with 8 rows in flight per thread reached 81 %; ring depth 8 -> 89 %, 12 -> 96 %).
template <typename T, int K, bool SILU>
__global__ void __launch_bounds__(CONV_THREADS, TV_CONV_MINWARPS * 32 / CONV_THREADS)
conv1d_fwd_kernel(const T* __restrict__ x, const T* __restrict__ weight, const T* __restrict__ bias,
                  const T* __restrict__ init, T* __restrict__ out, T* __restrict__ fin,
                  int dim, int L, int64_t xbs, int64_t xss, int64_t obs, int64_t oss) {
  constexpr int V = Vec16<T>::N;
  constexpr bool FAST = sizeof(T) == 2;
  const int c0 = (blockIdx.x * CONV_THREADS + threadIdx.x) * V;
  if (c0 >= dim) return;
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * CONV_TOK;
  const int t1 = min(t0 + CONV_TOK, L);
  x += (int64_t)b * xbs + c0;
  out += (int64_t)b * obs + c0;

  // All arithmetic is done on float2 pairs (FFMA2 / FMUL2 / FADD2, sm_100): the kernel is issue-bound next to the
  // memory system (two MUFU + ~10 scalar instructions per element), and the packed forms halve the FP issue slots.
  constexpr int V2 = V / 2;
  // weights (dim, K) row-major: this thread's V*K values are contiguous
  float2 w[K][V2], bv[V2];
#pragma unroll
  for (int v = 0; v < V2; ++v) {
#pragma unroll
    for (int k = 0; k < K; ++k)
      w[k][v] = make_float2(to_f32<T>(weight[(int64_t)(c0 + 2 * v) * K + k]), to_f32<T>(weight[(int64_t)(c0 + 2 * v + 1) * K + k]));
    bv[v] = bias != nullptr ? make_float2(to_f32<T>(bias[c0 + 2 * v]), to_f32<T>(bias[c0 + 2 * v + 1])) : make_float2(0.f, 0.f);
  }

  // the K-1 rows preceding t0
  float2 win[K - 1][V2];
#pragma unroll
  for (int j = 0; j < K - 1; ++j) {
    const int t = t0 - (K - 1) + j;
    float tmp[V];
    if (t >= 0) {
      Raw4<T> r; r.load(x + (int64_t)t * xss); r.unpack(tmp);
    } else if (init != nullptr) {  // (b, dim, K-1): column t+(K-1) of the carried-in state
#pragma unroll
      for (int v = 0; v < V; ++v) tmp[v] = to_f32<T>(init[((int64_t)b * dim + c0 + v) * (K - 1) + (t + K - 1)]);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) tmp[v] = 0.f;
    }
#pragma unroll
    for (int v = 0; v < V2; ++v) win[j][v] = make_float2(tmp[2 * v], tmp[2 * v + 1]);
  }

  // Row prefetch through shared memory: every thread keeps D rows (16 bytes each) in flight with cp.async into its own
  // ring slots (a warp's slots of one depth are 512 contiguous bytes), so the bytes in flight are not bounded by registers.
  constexpr int D = TV_CONV_RING;
  extern __shared__ uint4 ring_[];
  uint4* my = ring_ + ((threadIdx.x >> 5) * D) * 32 + (threadIdx.x & 31);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (t0 + d < t1) cp_async16(my + d * 32, x + (int64_t)(t0 + d) * xss);
    cp_async_commit();
  }
  int slot = 0;
#pragma unroll 4
  for (int tt = t0; tt < t1; ++tt) {
    cp_async_wait<D - 1>();
    const uint4 raw = my[slot * 32];
    float xf[V], o[V];
    unpack16<T>(raw, xf);
#pragma unroll
    for (int v = 0; v < V2; ++v) {
      const float2 xv = make_float2(xf[2 * v], xf[2 * v + 1]);
      float2 acc = bv[v];
#pragma unroll
      for (int k = 0; k < K - 1; ++k) acc = __ffma2_rn(w[k][v], win[k][v], acc);
      acc = __ffma2_rn(w[K - 1][v], xv, acc);
      if (SILU) {
        if (FAST) {
          const float2 tneg = __fmul2_rn(acc, make_float2(-1.4426950408889634f, -1.4426950408889634f));
          const float2 d = __fadd2_rn(make_float2(ex2_approx_f(tneg.x), ex2_approx_f(tneg.y)), make_float2(1.f, 1.f));
          acc = __fmul2_rn(acc, make_float2(rcp_approx_f(d.x), rcp_approx_f(d.y)));
        } else {
          acc = make_float2(silu<false>(acc.x), silu<false>(acc.y));
        }
      }
      o[2 * v] = acc.x; o[2 * v + 1] = acc.y;
#pragma unroll
      for (int k = 0; k < K - 2; ++k) win[k][v] = win[k + 1][v];
      win[K - 2][v] = xv;
    }
    if (tt + D < t1) cp_async16(my + slot * 32, x + (int64_t)(tt + D) * xss);   // refill the slot just consumed
    cp_async_commit();
    slot = slot + 1 == D ? 0 : slot + 1;
    Raw4<T>::store(out + (int64_t)tt * oss, o);
  }

  // carried-out state = the last K-1 input columns (includes carried-in columns when L < K-1)
  if (fin != nullptr && t1 == L) {
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int j = 0; j < K - 1; ++j)
        fin[((int64_t)b * dim + c0 + v) * (K - 1) + j] = from_f32<T>((v & 1) ? win[j][v >> 1].y : win[j][v >> 1].x);
  }
}

template <typename T, int K>
static int launch_conv(const tv_conv1d_params& p, cudaStream_t s) {
  constexpr int V = Vec16<T>::N;
  dim3 grid((unsigned)ceil_div(p.dim / V, CONV_THREADS), (unsigned)ceil_div(p.seqlen, CONV_TOK), p.batch);
  auto kern = p.silu ? conv1d_fwd_kernel<T, K, true> : conv1d_fwd_kernel<T, K, false>;
  kern<<<grid, CONV_THREADS, (size_t)(CONV_THREADS / 32) * TV_CONV_RING * 512, s>>>((const T*)p.x, (const T*)p.weight, (const T*)p.bias,
                                     (const T*)p.initial_states, (T*)p.out, (T*)p.final_states, p.dim,
                                     p.seqlen, p.x_batch_stride, p.x_seq_stride, p.out_batch_stride,
                                     p.out_seq_stride);
  TV_LAUNCH_OK();
  return TV_OK;
}

template <typename T>
static int dispatch_width(const tv_conv1d_params& p, cudaStream_t s) {
  switch (p.width) {
    case 2: return launch_conv<T, 2>(p, s);
    case 3: return launch_conv<T, 3>(p, s);
    case 4: return launch_conv<T, 4>(p, s);
    default:
      set_error("causal_conv1d: width %d unsupported (2..4)", p.width);
      return TV_ERR_UNSUPPORTED;
  }
}

}  // namespace tv

extern "C" int tv_causal_conv1d_fwd(const tv_conv1d_params* p, void* stream) {
  using namespace tv;
  TV_CHECK_ARG(p != nullptr, "causal_conv1d: null params");
  TV_CHECK_ARG(p->x && p->weight && p->out, "causal_conv1d: x, weight and out must be non-null");
  TV_CHECK_ARG(p->batch > 0 && p->dim > 0 && p->seqlen > 0, "causal_conv1d: empty problem (b=%d dim=%d L=%d)",
               p->batch, p->dim, p->seqlen);
  TV_CHECK_ARG(p->dtype == TV_F32 || p->dtype == TV_BF16, "causal_conv1d: dtype %d", p->dtype);
  const int V = 8;   // documented contract: 16-byte granularity for bf16 (the kernel itself needs 4 channels)
  const int esz = p->dtype == TV_BF16 ? 2 : 4;
  TV_CHECK_ARG(p->dim % (p->dtype == TV_BF16 ? 8 : 4) == 0, "causal_conv1d: dim %d must be a multiple of %d", p->dim,
               p->dtype == TV_BF16 ? 8 : 4);
  const int SV = p->dtype == TV_BF16 ? 8 : 4;
  TV_CHECK_ARG(p->x_seq_stride % SV == 0 && p->x_batch_stride % SV == 0 && p->out_seq_stride % SV == 0 &&
                   p->out_batch_stride % SV == 0,
               "causal_conv1d: strides must be multiples of %d elements (16 bytes)", SV);
  (void)V;
  TV_CHECK_ARG(((uintptr_t)p->x % 16 == 0) && ((uintptr_t)p->out % 16 == 0) && ((uintptr_t)p->weight % 16 == 0) &&
                   (p->bias == nullptr || (uintptr_t)p->bias % 16 == 0),
               "causal_conv1d: x/out/weight/bias must be 16-byte aligned");
  (void)esz;
  cudaStream_t s = (cudaStream_t)stream;
  return p->dtype == TV_BF16 ? dispatch_width<__nv_bfloat16>(*p, s) : dispatch_width<float>(*p, s);
}

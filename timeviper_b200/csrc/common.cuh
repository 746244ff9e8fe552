// Shared helpers for the timeviper_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/timeviper_b200.h"

namespace tv {

void set_error(const char* fmt, ...);
// bookkeeping for tv_debug_launch_count(): every kernel this library enqueues is counted (bench.py reports it)
void count_launches(int n);

#define TV_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::tv::set_error(__VA_ARGS__);         \
      return TV_ERR_INVALID;                \
    }                                       \
  } while (0)

#define TV_CUDA_OK(expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::tv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return TV_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

// after a kernel launch: count it (tv_debug_launch_count) and surface a launch error
#define TV_LAUNCH_OK()            \
  do {                            \
    ::tv::count_launches(1);      \
    TV_CUDA_OK(cudaGetLastError()); \
  } while (0)

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <typename T> struct Vec16;  // 16-byte vector of T
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte global load/store to/from N floats
template <typename T> __device__ __forceinline__ void load16(const T* p, float (&v)[Vec16<T>::N]);
template <> __device__ __forceinline__ void load16<float>(const float* p, float (&v)[4]) {
  float4 r = *reinterpret_cast<const float4*>(p);
  v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <> __device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // bf16 -> f32 is a 16-bit shift
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
// raw 16-byte register image -> N floats
template <typename T> __device__ __forceinline__ void unpack16(const uint4& r, float (&v)[Vec16<T>::N]);
template <> __device__ __forceinline__ void unpack16<float>(const uint4& r, float (&v)[4]) {
  v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
}
template <> __device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4& r, float (&v)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <typename T> __device__ __forceinline__ void store16(T* p, const float (&v)[Vec16<T>::N]);
template <> __device__ __forceinline__ void store16<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
template <> __device__ __forceinline__ void store16<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]); r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]); r.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = r;
}

// SiLU. FAST uses the SFU exp2/rcp approximations (bf16 I/O); otherwise IEEE-ish expf and division (fp32 I/O).
template <bool FAST> __device__ __forceinline__ float silu(float x) {
  if (FAST) return __fdividef(x, 1.0f + __expf(-x));
  return x / (1.0f + expf(-x));
}

__device__ __forceinline__ float ex2_approx_f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx_f(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace tv

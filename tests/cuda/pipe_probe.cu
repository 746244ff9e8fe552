// Single-warp issue-rate probe for the scalar pipes the M build of ssd_tc.cu leans on (sm_100a):
// MUFU.EX2, F2FP.BF16 pack, FFMA, integer rounding pack.  cycles per warp-instruction, ILP 16.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(long long* out, float seed, int nwarps_active) {
  if ((threadIdx.x >> 5) >= nwarps_active) return;
  float v[16];
  uint32_t u[16];
  for (int i = 0; i < 16; ++i) { v[i] = seed + i * 0.01f + threadIdx.x * 1e-3f; u[i] = __float_as_uint(v[i]); }
  long long t0 = clock64();
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) v[i] = ex2(v[i]);
      if (MODE == 1) { __nv_bfloat162 h = __floats2bfloat162_rn(v[i], v[(i + 1) & 15]); u[i] ^= *reinterpret_cast<uint32_t*>(&h); v[i] = __uint_as_float(u[i]); }
      if (MODE == 2) v[i] = fmaf(v[i], 1.0001f, 0.5f);
      if (MODE == 3) { uint32_t a = __float_as_uint(v[i]) + 0x8000u, b = __float_as_uint(v[(i + 1) & 15]) + 0x8000u; u[i] ^= __byte_perm(a, b, 0x7632); v[i] = __uint_as_float(u[i]); }
      if (MODE == 4) {  // software exp2 on the FMA/ALU pipes: 2^x = 2^floor(x) * p(frac), cubic
        float x = fmaxf(v[i], -126.f);
        float fl = floorf(x);
        float f = x - fl;
        float p = fmaf(fmaf(fmaf(0.0555041f, f, 0.2402265f), f, 0.6931472f), f, 1.0f);
        v[i] = __uint_as_float(__float_as_uint(p) + ((int)fl << 23)) * 1e-3f;
      }
    }
  }
  long long t1 = clock64();
  float acc = 0; for (int i = 0; i < 16; ++i) acc += v[i] + __uint_as_float(u[i]);
  if (threadIdx.x == 0) { out[MODE] = (t1 - t0); out[8] = (long long)acc; }
}

int main() {
  long long* d; cudaMalloc(&d, 128);
  const char* names[5] = {"MUFU.EX2", "F2FP.BF16.PACK_AB (cvt.rn.bf16x2.f32)", "FFMA", "IADD+IADD+PRMT round-half-up pack", "software exp2 (floor+cubic+shift)"};
  for (int nw = 1; nw <= 4; nw *= 4) {
    k<0><<<1, 128>>>(d, 0.5f, nw); k<1><<<1, 128>>>(d, 0.5f, nw); k<2><<<1, 128>>>(d, 0.5f, nw); k<3><<<1, 128>>>(d, 0.5f, nw); k<4><<<1, 128>>>(d, -3.5f, nw);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
    for (int m = 0; m < 5; ++m) printf("%d warp(s): %-42s %6.2f cycles per warp-op (1024 ops)\n", nw, names[m], h[m] / 1024.0);
  }
  return 0;
}

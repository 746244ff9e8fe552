// Times the M-build block bodies of ssd_tc.cu in isolation (1 warp per SMSP, then 3 warps per SMSP).
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
#include "../../timeviper_b200/csrc/sm100.cuh"
using namespace tv::sm100;
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&t); }

template <int ID>
__device__ __noinline__ void m_block(uint32_t t_src, uint32_t t_dst, const float* __restrict__ sFk, float Em, int lane, float Dh, bool diag) {
  asm volatile("// copy %0" :: "n"(ID));
  uint32_t r[32];
  tmem_ld32(t_src, r);
  float e[32];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 f4 = *reinterpret_cast<const float4*>(sFk + j);
    e[j] = Em + f4.x; e[j + 1] = Em + f4.y; e[j + 2] = Em + f4.z; e[j + 3] = Em + f4.w;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) e[j] = ex2_approx(e[j]);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    e[j] *= __uint_as_float(r[j]);
    if (diag && j > lane) e[j] = 0.f;
    if (diag && j == lane) e[j] += Dh;
  }
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(e[2 * j], e[2 * j + 1]);
  tmem_st16(t_dst, pk);
}
__device__ __forceinline__ void m_block_offdiag(uint32_t t_src, uint32_t t_dst, const float* __restrict__ sVk, float um) {
  uint32_t r[32];
  tmem_ld32(t_src, r);
  float e[32];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 v4 = *reinterpret_cast<const float4*>(sVk + j);
    e[j] = v4.x * um; e[j + 1] = v4.y * um; e[j + 2] = v4.z * um; e[j + 3] = v4.w * um;
  }
  tmem_ld_wait();
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(e[2 * j] * __uint_as_float(r[2 * j]), e[2 * j + 1] * __uint_as_float(r[2 * j + 1]));
  tmem_st16(t_dst, pk);
}

__global__ void __launch_bounds__(384) k(long long* out, int mode, int nthreads_active) {
  __shared__ float sF[512];
  __shared__ uint32_t tmem_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sF[i] = -0.01f * i;
  if (warp == 0) tmem_alloc<512>(&tmem_s);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (threadIdx.x < nthreads_active) {
    const uint32_t tcb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
      if (mode == 0) m_block<0>(tcb + (it & 3) * 32, tcb + (it & 3) * 16, sF + (it & 3) * 32, -0.3f * lane, lane, 1.0f, true);
      else if (mode == 2) {   // every warp runs its own copy of the code (12 distinct copies: 3 per SMSP)
        switch (warp) {
#define CASE(W) case W: m_block<W + 1>(tcb + (it & 3) * 32, tcb + (it & 3) * 16, sF + (it & 3) * 32, -0.3f * lane, lane, 1.0f, true); break;
          CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11)
#undef CASE
        }
      }
      else m_block_offdiag(tcb + (it & 3) * 32, tcb + (it & 3) * 16, sF + (it & 3) * 32, 0.5f);
    }
    tmem_st_wait();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[mode * 4 + (nthreads_active > 128 ? 1 : 0)] = (t1 - t0) / 64;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
int main() {
  long long* d; cudaMalloc(&d, 128); cudaMemset(d, 0, 128);
  for (int mode = 0; mode < 3; ++mode) { k<<<1, 384>>>(d, mode, 128); k<<<1, 384>>>(d, mode, 384); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  printf("diag block   : %lld cycles (1 warp/SMSP), %lld cycles (3 warps/SMSP)\n", h[0], h[1]);
  printf("diag block, distinct code copy per warp (noinline): %lld cycles (1 warp/SMSP), %lld cycles (3 warps/SMSP)\n", h[8], h[9]);
  printf("offdiag block: %lld cycles (1 warp/SMSP), %lld cycles (3 warps/SMSP)\n", h[4], h[5]);
  return 0;
}

// GPU probe: tcgen05.ld / tcgen05.st throughput per SM versus the number of warps, alone and next to a running
// tcgen05.mma stream.  Not part of the product: it answers "how many TMEM->register bytes per cycle can the roles of
// ssd_tc.cu draw in total", which bounds the M build, the state decay and the epilogue drain together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
#include "../../timeviper_b200/csrc/sm100.cuh"
using namespace tv::sm100;

constexpr int ITERS = 256;

// mode bit0..1: 0 = ld x32 + wait, 1 = two ld x32 in flight + wait, 2 = two st x16 + wait::st, 3 = ld x32, math, st x16 (M-build shape)
// mode bit2: an extra thread keeps the tensor pipe busy with 128x128x16 SS MMAs into columns 256..383
__global__ void __launch_bounds__(576, 1) k(long long* out, int nwarps, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); done = 0; }
  if (warp == 17) tmem_alloc<512>(&tmem_slot);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int op = mode & 3;
  if (warp == 17) {
    if ((mode & 4) && elect_one()) {
      constexpr uint32_t ID = umma_idesc_bf16(128, 128, false, false);
      const uint64_t da = umma_smem_desc(smem_u32(smem), 16, 1024, SWZ_128B);
      const uint64_t db = umma_smem_desc(smem_u32(smem) + 32768, 16, 1024, SWZ_128B);
      int n = 0;
      while (*(volatile int*)&done < nwarps) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t o = (uint32_t)(((j >> 2) * 16384 + (j & 3) * 32) >> 4);
          umma_ss(tmem + 256, umma_desc_advance(da, o), umma_desc_advance(db, o), ID, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, n & 1);
        ++n;
      }
      out[blockIdx.x * 32 + 31] = n;
    }
  } else if (warp < nwarps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t acc = 0;
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
      if (op == 0) {
        uint32_t r[32];
        tmem_ld32(base + (it & 1) * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
      } else if (op == 1) {
        uint32_t r[32], q[32];
        tmem_ld32(base, r);
        tmem_ld32(base + 32, q);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j] + q[j];
      } else if (op == 2) {
        uint32_t r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = acc + j + it;
        tmem_st16(base, r);
        tmem_st16(base + 16, r);
        tmem_st_wait();
      } else {
        uint32_t r[32], pk[16];
        tmem_ld32(base, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[2 * j]) * 1.5f, __uint_as_float(r[2 * j + 1]) * 1.5f);
          pk[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        tmem_st16(base + 32, pk);
        tmem_st_wait();
        acc ^= pk[3];
      }
    }
    const long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x * 32 + warp] = t1 - t0; atomicAdd(&done, 1); }
    if (acc == 0x12345678u) out[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d;
  cudaMalloc(&d, 32 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024 + 1024);
  const char* names[4] = {"ld x32 + wait", "2 ld x32 + wait", "2 st x16 + wait", "ld x32, pack, st x16"};
  const int bytes[4] = {4096, 8192, 4096, 4096 + 2048};
  for (int mma = 0; mma < 2; ++mma)
    for (int op = 0; op < 4; ++op)
      for (int nw : {1, 4, 8, 16}) {
        long long h[32] = {0};
        cudaMemset(d, 0, 32 * 8);
        k<<<1, 576, 66 * 1024 + 1024>>>(d, nw, op | (mma << 2));
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 32 * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
        printf("%-22s mma=%d warps=%2d: %7.1f cycles/iter/warp  -> %6.1f B/cycle/SM%s\n", names[op], mma, nw,
               (double)mx / ITERS, (double)bytes[op] * nw * ITERS / mx, mma ? "" : "");
        if (mma) printf("    (%lld MMA groups of 8 completed meanwhile: %.0f cycles per group)\n", h[31], h[31] ? (double)mx / h[31] : 0.0);
      }
  return 0;
}

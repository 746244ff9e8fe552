// GPU probe: latency and throughput of plain global loads (L2-resident data) as a function of how much shared memory
// the CTA holds, i.e. of how much of the 228 KB unified array is left for L1.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o l1_probe l1_probe.cu && ./l1_probe
#include <stdio.h>
#include <stdint.h>

template <int MODE>   // 0: ld.global.nc (LDG.CONSTANT), 1: ld.global.cg, 2: ld.global (default)
__device__ __forceinline__ uint4 ld(const uint4* p) {
  uint4 v;
  if (MODE == 0) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if (MODE == 1) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// every warp streams its own region: per iteration NLD coalesced LDG.128 (512 B per warp each), then waits for all of them
template <int MODE, int NLD>
__global__ void __launch_bounds__(512, 1) k(const uint4* __restrict__ src, long long* out, int iters, size_t span) {
  extern __shared__ uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4* p = src + ((size_t)(blockIdx.x * 16 + warp) * 4096 + lane);
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint4 v[NLD];
#pragma unroll
    for (int j = 0; j < NLD; ++j) v[j] = ld<MODE>(p + (size_t)((it * NLD + j) % 120) * 32);
#pragma unroll
    for (int j = 0; j < NLD; ++j) acc += v[j].x ^ v[j].w;
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 0x1234567u) smem[threadIdx.x] = 1;
}

int main() {
  const size_t bytes = 148ull * 16 * 4096 * 16;    // 155 MB?  no: 148*16*4096 uint4 = 155 MB
  uint4* d;
  cudaMalloc(&d, bytes);
  cudaMemset(d, 0, bytes);
  long long* out;
  cudaMalloc(&out, 148 * 16 * 8);
  const int iters = 200;
  auto run = [&](auto kern, const char* name, int nld) {
    for (int smem_kb : {8, 100, 180, 210, 224}) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
      for (int nw : {1, 8, 16}) {
        long long h[148 * 16];
        for (int rep = 0; rep < 2; ++rep) kern<<<128, nw * 32, smem_kb * 1024>>>(d, out, iters, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
        cudaMemcpy(h, out, 148 * 16 * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int b = 0; b < 128; ++b) for (int w = 0; w < nw; ++w) avg += (double)h[b * 16 + w] / (128 * nw);
        printf("%-14s x%d  smem %3d KB  warps %2d: %7.0f cycles per batch  -> %6.1f B/cycle/SM\n", name, nld, smem_kb, nw,
               avg / iters, 512.0 * nld * nw * iters / avg);
      }
    }
  };
  run(k<0, 1>, "ld.global.nc", 1);
  run(k<0, 8>, "ld.global.nc", 8);
  run(k<1, 8>, "ld.global.cg", 8);
  return 0;
}

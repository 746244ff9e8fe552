// GPU probe: TMA load rate of the x tile of ssd_tc.cu (128 tokens x 80 bf16 of one head, token stride 24,576 B)
// as a function of the box shape and of the ring depth, with one CTA per head streaming its own head like the
// fused kernel does.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu -lcuda && ./tma_probe
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
#include "../../timeviper_b200/csrc/sm100.cuh"
#include "../../timeviper_b200/csrc/tmap.h"
using namespace tv::sm100;

struct Maps { CUtensorMap m16, m64, m80, b64; };
constexpr int MAXD = 8;

// mode 0: 5 boxes of 16 elements (SW32).  1: one box of 64 (SW128) + one of 16 (SW32).  2: one box of 80, no swizzle.
// mode 3: mode 0 + a B-like and a C-like tile (2 x 2 boxes of 64 x 128 rows, SW128) per chunk, as in the fused kernel.
// mode 4: mode 1 + the same B/C tiles.
__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ Maps maps, int nchunks, int depth, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[MAXD], empty[MAXD];
  const int h = blockIdx.x;
  const bool bc = mode >= 3;
  const int xm = mode >= 3 ? mode - 3 : mode;
  const uint32_t stage = 20480 + (bc ? 65536 : 0);
  const uint32_t bytes = 20480 + (bc ? 65536 : 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % depth, u = c / depth;
      if (c >= depth) mbar_wait(&empty[s], (u - 1) & 1);
      uint8_t* dst = smem + s * stage;
      mbar_arrive_expect_tx(&full[s], bytes);
      if (xm == 0) {
        for (int i = 0; i < 5; ++i) tma_load_4d(dst + i * 4096, &maps.m16, &full[s], 16 * i, h, c * 128, 0);
      } else if (xm == 1) {
        tma_load_4d(dst, &maps.m64, &full[s], 0, h, c * 128, 0);
        tma_load_4d(dst + 16384, &maps.m16, &full[s], 64, h, c * 128, 0);
      } else {
        tma_load_4d(dst, &maps.m80, &full[s], 0, h, c * 128, 0);
      }
      if (bc) {
        const int g = h / 16;
        for (int i = 0; i < 4; ++i)
          tma_load_4d(dst + 20480 + i * 16384, &maps.b64, &full[s], 64 * (i & 1), (i >> 1) ? 8 + g : g, c * 128, 0);
      }
    }
  } else if (threadIdx.x == 32) {
    const long long t0 = clock64();
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % depth, u = c / depth;
      mbar_wait(&full[s], u & 1);
      mbar_arrive(&empty[s]);
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int T = 16384, W = 12288, H = 128;
  __nv_bfloat16* d;
  cudaMalloc(&d, (size_t)T * W * 2);
  cudaMemset(d, 0, (size_t)T * W * 2);
  long long* out;
  cudaMalloc(&out, 148 * 8);
  Maps maps;
  const uint64_t dx[4] = {80, (uint64_t)H, (uint64_t)T, 1}, sx[3] = {160, (uint64_t)W * 2, (uint64_t)T * W * 2};
  const uint32_t b16[4] = {16, 1, 128, 1}, b64[4] = {64, 1, 128, 1}, b80[4] = {80, 1, 128, 1};
  bool ok = tv::encode_bf16_tmap(&maps.m16, d, 4, dx, sx, b16, CU_TENSOR_MAP_SWIZZLE_32B);
  ok &= tv::encode_bf16_tmap(&maps.m64, d, 4, dx, sx, b64, CU_TENSOR_MAP_SWIZZLE_128B);
  ok &= tv::encode_bf16_tmap(&maps.m80, d, 4, dx, sx, b80, CU_TENSOR_MAP_SWIZZLE_NONE);
  const uint64_t db[4] = {128, 16, (uint64_t)T, 1}, sb[3] = {256, (uint64_t)W * 2, (uint64_t)T * W * 2};
  ok &= tv::encode_bf16_tmap(&maps.b64, d + 10240, 4, db, sb, b64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!ok) { printf("tensor map encode failed\n"); return 1; }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[5] = {"x: 5 boxes of 16 (SW32)", "x: box 64 (SW128) + box 16 (SW32)", "x: one box of 80 (no swizzle)",
                          "x 5x16 + B,C tiles", "x 64+16 + B,C tiles"};
  const int nchunks = T / 128;
  for (int mode = 0; mode < 5; ++mode)
    for (int depth : {2, 3, 4, 8}) {
      if (mode >= 3 && depth > 2) continue;
      long long h[148];
      for (int rep = 0; rep < 2; ++rep) {
        k<<<H, 64, 200 * 1024>>>(maps, nchunks, depth, mode, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, out, H * 8, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < H; ++i) avg += (double)h[i] / H;
      const double bytes = 20480.0 + (mode >= 3 ? 65536.0 : 0.0);
      printf("%-36s depth %d: %7.0f cycles/chunk  %5.1f B/cycle/SM\n", names[mode], depth, avg / nchunks, bytes * nchunks / avg);
    }
  return 0;
}

// GPU probe for the UMMA operand layouts ssd_tc.cu depends on.  Not part of the product: it exists to pin the
// descriptor conventions (K-major SW128, MN-major SW128 / SW32, A-from-TMEM packing) on real sm_100a hardware.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../timeviper_b200/csrc/sm100.cuh"
#include "../../timeviper_b200/csrc/tmap.h"

using namespace tv::sm100;
typedef __nv_bfloat16 bf16;

constexpr int M = 128, K = 128, NB = 128, P = 80;

struct Maps { CUtensorMap a, bk, x32, x128, bmn; };

// smem byte offsets of the canonical layouts (tile bases are 1024-byte aligned)
__device__ __forceinline__ uint32_t off_mn_sw32(int k, int p) {   // [k rows][p], atoms of 16 p, 32-byte rows
  return (p >> 4) * 4096 + k * 32 + ((((p & 15) >> 3) ^ ((k >> 2) & 1)) << 4) + (p & 7) * 2;
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ Maps maps, const bf16* __restrict__ Ag,
                                             const bf16* __restrict__ Xg, float* __restrict__ out, int test) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 32 KB  A  [m][k]  K-major SW128 (2 boxes of 64 k)
  uint8_t* sB = smem + 32768;         // 32 KB  Bk [n][k] K-major SW128, or Bssm [k][n] MN-major SW128
  uint8_t* sX = smem + 65536;         // 32 KB  X  [k][p]  MN-major SW32 (5 x 4 KB) or SW128 (2 x 16 KB)
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD = tmem;            // accumulator, up to 128 cols
  const uint32_t tA = tmem + 128;      // A operand in TMEM (64 cols of packed bf16)

  if (threadIdx.x == 0) {
    uint32_t bytes = 32768;
    tma_load_2d(sA, &maps.a, &bar_tma, 0, 0);
    tma_load_2d(sA + 16384, &maps.a, &bar_tma, 64, 0);
    if (test == 1) {
      tma_load_2d(sB, &maps.bk, &bar_tma, 0, 0); tma_load_2d(sB + 16384, &maps.bk, &bar_tma, 64, 0); bytes += 32768;
    } else if (test == 4) {
      tma_load_2d(sB, &maps.bmn, &bar_tma, 0, 0); tma_load_2d(sB + 16384, &maps.bmn, &bar_tma, 64, 0); bytes += 32768;
    }
    if (test == 2 || test == 3 || test == 4) {
      for (int i = 0; i < 5; ++i) tma_load_2d(sX + i * 4096, &maps.x32, &bar_tma, 16 * i, 0);
      bytes += 5 * 4096;
    } else if (test == 6) {
      tma_load_2d(sX, &maps.x128, &bar_tma, 0, 0); tma_load_2d(sX + 16384, &maps.x128, &bar_tma, 64, 0); bytes += 32768;
    }
    mbar_arrive_expect_tx(&bar_tma, bytes);
  }
  if (test == 5) {  // X written by threads (thread k owns row k), scaled by 2: validates the manual SW32 layout
    const int k = threadIdx.x;
    for (int p = 0; p < P; p += 8) {
      uint4 v = *reinterpret_cast<const uint4*>(Xg + k * P + p);
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
      for (int i = 0; i < 4; ++i) h[i] = __hmul2(h[i], __floats2bfloat162_rn(2.f, 2.f));
      *reinterpret_cast<uint4*>(sX + off_mn_sw32(k, p)) = v;
    }
    fence_proxy_async();
  }
  if (test == 3) {  // A into TMEM: lane m holds row m, column j packs (A[m][2j], A[m][2j+1])
    const int m = threadIdx.x;
    for (int c = 0; c < 64; c += 16) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) r[j] = *reinterpret_cast<const uint32_t*>(Ag + m * K + 2 * (c + j));
      tmem_st16(tA + ((uint32_t)(warp * 32) << 16) + c, r);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  mbar_wait(&bar_tma, 0);
  tc_fence_after();

  int ncols = 0;
  if (threadIdx.x == 0) {
    if (test == 1) {  // D[m][n] = sum_k A[m][k] Bk[n][k]
      const uint32_t idesc = umma_idesc_bf16(128, NB, false, false);
      for (int j = 0; j < 8; ++j) {
        const uint32_t o = (j >> 2) * 16384 + (j & 3) * 32;
        umma_ss(tD, umma_smem_desc(smem_u32(sA) + o, 16, 1024, SWZ_128B),
                umma_smem_desc(smem_u32(sB) + o, 16, 1024, SWZ_128B), idesc, j > 0);
      }
    } else if (test == 2 || test == 5) {  // D[m][p] = sum_k A[m][k] X[k][p], X MN-major SW32
      const uint32_t idesc = umma_idesc_bf16(128, P, false, true);
      for (int j = 0; j < 8; ++j) {
        const uint32_t oa = (j >> 2) * 16384 + (j & 3) * 32;
        umma_ss(tD, umma_smem_desc(smem_u32(sA) + oa, 16, 1024, SWZ_128B),
                umma_smem_desc(smem_u32(sX) + j * 512, 4096, 256, SWZ_32B), idesc, j > 0);
      }
    } else if (test == 3) {  // A from TMEM
      const uint32_t idesc = umma_idesc_bf16(128, P, false, true);
      for (int j = 0; j < 8; ++j)
        umma_ts(tD, tA + j * 8, umma_smem_desc(smem_u32(sX) + j * 512, 4096, 256, SWZ_32B), idesc, j > 0);
    } else if (test == 4) {  // D[n][p] = sum_k Bssm[k][n] X[k][p]: A MN-major SW128, B MN-major SW32
      const uint32_t idesc = umma_idesc_bf16(128, P, true, true);
      for (int j = 0; j < 8; ++j)
        umma_ss(tD, umma_smem_desc(smem_u32(sB) + j * 2048, 16384, 1024, SWZ_128B),
                umma_smem_desc(smem_u32(sX) + j * 512, 4096, 256, SWZ_32B), idesc, j > 0);
    } else if (test == 6) {  // X MN-major SW128, N = 80 spans 1.25 atoms
      const uint32_t idesc = umma_idesc_bf16(128, P, false, true);
      for (int j = 0; j < 8; ++j) {
        const uint32_t oa = (j >> 2) * 16384 + (j & 3) * 32;
        umma_ss(tD, umma_smem_desc(smem_u32(sA) + oa, 16, 1024, SWZ_128B),
                umma_smem_desc(smem_u32(sX) + j * 2048, 16384, 1024, SWZ_128B), idesc, j > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  ncols = (test == 1) ? NB : P;
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = threadIdx.x;
  for (int c = 0; c < ncols; c += 16) {
    uint32_t r[16];
    tmem_ld16(tD + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[row * 128 + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
  (void)lane;
}

static float bf(bf16 v) { return __bfloat162float(v); }

int main() {
  std::vector<bf16> A(M * K), Bk(NB * K), X(K * P), Bmn(K * NB);
  srand(1);
  auto rnd = []() { return (float)(rand() % 2001 - 1000) / 1000.f; };
  for (auto& v : A) v = __float2bfloat16(rnd());
  for (auto& v : Bk) v = __float2bfloat16(rnd());
  for (auto& v : X) v = __float2bfloat16(rnd());
  for (auto& v : Bmn) v = __float2bfloat16(rnd());
  bf16 *dA, *dBk, *dX, *dBmn; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dBk, Bk.size() * 2); cudaMalloc(&dX, X.size() * 2 + 4096);
  cudaMalloc(&dBmn, Bmn.size() * 2); cudaMalloc(&dO, 128 * 128 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBk, Bk.data(), Bk.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBmn, Bmn.data(), Bmn.size() * 2, cudaMemcpyHostToDevice);
  Maps maps;
  {
    uint64_t d[2] = {K, M}, s[1] = {K * 2}; uint32_t b[2] = {64, 128};
    bool ok = tv::encode_bf16_tmap(&maps.a, dA, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B);
    ok &= tv::encode_bf16_tmap(&maps.bk, dBk, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B);
    uint64_t dn[2] = {NB, K}, sn[1] = {NB * 2};
    ok &= tv::encode_bf16_tmap(&maps.bmn, dBmn, 2, dn, sn, b, CU_TENSOR_MAP_SWIZZLE_128B);
    uint64_t dx[2] = {P, K}, sx[1] = {P * 2}; uint32_t bx32[2] = {16, 128};
    ok &= tv::encode_bf16_tmap(&maps.x32, dX, 2, dx, sx, bx32, CU_TENSOR_MAP_SWIZZLE_32B);
    ok &= tv::encode_bf16_tmap(&maps.x128, dX, 2, dx, sx, b, CU_TENSOR_MAP_SWIZZLE_128B);
    if (!ok) { printf("tensor map encode failed\n"); return 2; }
  }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[7] = {"", "SS A=Kmaj-SW128 B=Kmaj-SW128 (C.B^T)", "SS B=MNmaj-SW32 via TMA (C.S / M.x)",
                          "TS A=TMEM packed bf16, B=MNmaj-SW32", "SS A=MNmaj-SW128 B=MNmaj-SW32 (state)",
                          "SS B=MNmaj-SW32 written by threads (x2)", "SS B=MNmaj-SW128 N=80 over 1.25 atoms"};
  int fails = 0;
  for (int t = 1; t <= 6; ++t) {
    cudaMemset(dO, 0, 128 * 128 * 4);
    probe<<<1, 128, 100 * 1024>>>(maps, dA, dX, dO, t);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("test %d: CUDA error %s\n", t, cudaGetErrorString(e)); return 3; }
    std::vector<float> O(128 * 128);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    const int ncols = t == 1 ? NB : P;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < ncols; ++j) {
        double ref = 0;
        for (int k = 0; k < K; ++k) {
          if (t == 1) ref += (double)bf(A[i * K + k]) * bf(Bk[j * K + k]);
          else if (t == 4) ref += (double)bf(Bmn[k * NB + i]) * bf(X[k * P + j]);
          else ref += (double)bf(A[i * K + k]) * bf(X[k * P + j]) * (t == 5 ? 2.0 : 1.0);
        }
        maxerr = fmax(maxerr, fabs(ref - O[i * 128 + j])); maxref = fmax(maxref, fabs(ref));
      }
    const bool pass = maxerr < 1e-3 * maxref + 1e-4;
    fails += !pass;
    printf("test %d [%s]: max err %.3e (max ref %.3f) %s\n", t, names[t], maxerr, maxref, pass ? "PASS" : "FAIL");
  }
  return fails ? 1 : 0;
}

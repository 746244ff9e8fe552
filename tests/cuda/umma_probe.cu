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


// ---------------------------------------------------------------------------------------------------------
// Timing: cycles for a group of 8 back-to-back MMAs (K = 128) of each operand-layout combination, and for the
// TMEM read pattern of the epilogue.  Operand contents are whatever is in smem (timing only).
__global__ void __launch_bounds__(128) probe_time(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem; uint8_t* sB = smem + 32768; uint8_t* sX = smem + 65536;
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a = smem_u32(sA), b = smem_u32(sB), x = smem_u32(sX);
  uint32_t parity = 0;
  if (threadIdx.x == 0) {
    const uint64_t aK0 = umma_smem_desc(a, 16, 1024, SWZ_128B), bK0 = umma_smem_desc(b, 16, 1024, SWZ_128B);
    const uint64_t aMN0 = umma_smem_desc(b, 16384, 1024, SWZ_128B);
    const uint64_t x320 = umma_smem_desc(x, 4096, 256, SWZ_32B), x1280 = umma_smem_desc(x, 16384, 1024, SWZ_128B);
#define TIME_GROUP(TYPE, BODY)                                                  \
    for (int rep = 0; rep < 2; ++rep) {                                        \
      long long t0 = clock64();                                                \
      _Pragma("unroll") for (int r = 0; r < 8; ++r) {                          \
        _Pragma("unroll") for (int j = 0; j < 8; ++j) {                        \
          const uint64_t ok = (uint64_t)(((j >> 2) * 16384 + (j & 3) * 32) >> 4); \
          (void)ok; BODY;                                                      \
        }                                                                      \
      }                                                                        \
      umma_commit(&bar_mma);                                                   \
      long long t1 = clock64();                                                \
      mbar_wait(&bar_mma, parity); parity ^= 1;                                \
      long long t2 = clock64();                                                \
      if (rep == 1) { out[TYPE] = (t2 - t0) / 8; out[16 + TYPE] = (t1 - t0) / 8; } \
    }
    TIME_GROUP(0, umma_ss(tmem, aK0 + ok, bK0 + ok, umma_idesc_bf16(128, 128, false, false), j > 0))
    TIME_GROUP(1, umma_ss(tmem + 128, aK0 + ok, x320 + (uint64_t)(j * 32), umma_idesc_bf16(128, 80, false, true), j > 0))
    TIME_GROUP(2, umma_ts(tmem + 128, tmem + 256 + j * 8, x320 + (uint64_t)(j * 32), umma_idesc_bf16(128, 80, false, true), j > 0))
    TIME_GROUP(3, umma_ss(tmem + 128, aMN0 + (uint64_t)(j * 128), x320 + (uint64_t)(j * 32), umma_idesc_bf16(128, 80, true, true), j > 0))
    TIME_GROUP(4, umma_ss(tmem + 128, aK0 + ok, x1280 + (uint64_t)(j * 128), umma_idesc_bf16(128, 80, false, true), j > 0))
    TIME_GROUP(5, umma_ts(tmem + 128, tmem + 256 + j * 8, x1280 + (uint64_t)(j * 128), umma_idesc_bf16(128, 80, false, true), j > 0))
    TIME_GROUP(6, umma_ss(tmem + 128, aMN0 + (uint64_t)(j * 128), x1280 + (uint64_t)(j * 128), umma_idesc_bf16(128, 80, true, true), j > 0))
  }
  __syncthreads();
  tc_fence_after();
  {  // epilogue-like TMEM read: 10 x (ld16 + wait) per warp, all 4 warps
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int rep = 0; rep < 8; ++rep)
      for (int c = 0; c < 160; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) acc ^= r[j];
      }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[8] = (t1 - t0) / 8; out[9] = acc; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------------------
// Contention probe: one thread issues the per-chunk MMA mix of ssd_tc.cu (S, D, G, O groups) R times while the
// other warps generate one kind of background traffic.  Reports cycles per (S+D+G+O) round.
__global__ void __launch_bounds__(256) probe_contention(long long* out, const uint8_t* gsrc, int noise) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem; uint8_t* sB = smem + 32768; uint8_t* sX = smem + 65536; uint8_t* sN = smem + 98304;  // noise area 64 KB
  __shared__ uint64_t bar_mma, bar_never, bar_tma;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 163840 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_never, 1); mbar_init(&bar_tma, 1); stop = 0; fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a = smem_u32(sA), b = smem_u32(sB), x = smem_u32(sX);
  if (threadIdx.x == 0) {
    const uint64_t aK0 = umma_smem_desc(a, 16, 1024, SWZ_128B), bK0 = umma_smem_desc(b, 16, 1024, SWZ_128B);
    const uint64_t aMN0 = umma_smem_desc(b, 16384, 1024, SWZ_128B);
    const uint64_t x320 = umma_smem_desc(x, 4096, 256, SWZ_32B);
    uint32_t parity = 0;
    long long t0 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      for (int r = 0; r < 16; ++r) {
#pragma unroll
        for (int j = 0; j < 8; ++j)   // S
          umma_ss(tmem + 416, aMN0 + (uint64_t)(j * 128), x320 + (uint64_t)(j * 32), umma_idesc_bf16(128, 80, true, true), 1u);
#pragma unroll
        for (int j = 0; j < 8; ++j)   // D
          umma_ts(tmem + 256, tmem + j * 8, x320 + (uint64_t)(j * 32), umma_idesc_bf16(128, 80, false, true), j > 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) { // G
          const uint64_t ok = (uint64_t)(((j >> 2) * 16384 + (j & 3) * 32) >> 4);
          umma_ss(tmem + 128, aK0 + ok, bK0 + ok, umma_idesc_bf16(128, 128, false, false), j > 0);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { // O
          const uint64_t ok = (uint64_t)(((j >> 2) * 16384 + (j & 3) * 32) >> 4);
          umma_ss(tmem + 336, aK0 + ok, x320 + (uint64_t)(j * 32), umma_idesc_bf16(128, 80, false, true), j > 0);
        }
      }
      umma_commit(&bar_mma);
      mbar_wait(&bar_mma, parity); parity ^= 1;
    }
    out[noise] = (clock64() - t0) / 16;
    stop = 1;
  } else if (warp >= 4) {        // warps 4..7: background traffic (warps 1-3 idle so TMEM lane quarters stay free)
    uint32_t acc = 0;
    uint32_t tparity = 0;
    while (!stop) {
      if (noise == 1) {          // shared-memory loads + stores, 16 bytes per lane
        for (int i = 0; i < 16; ++i) {
          uint4 v = *reinterpret_cast<const uint4*>(sN + ((warp - 4) * 16384 + ((i * 32 + lane) * 16) % 16384));
          v.x += acc; acc ^= v.y;
          *reinterpret_cast<uint4*>(sN + ((warp - 4) * 16384 + ((i * 32 + lane) * 16) % 16384)) = v;
        }
      } else if (noise == 2) {   // TMEM loads
        for (int c = 0; c < 64; c += 16) {
          uint32_t r[16];
          tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c, r);
          tmem_ld_wait();
          acc ^= r[3];
        }
      } else if (noise == 3) {   // mbarrier polling
        for (int i = 0; i < 8; ++i) acc += mbar_try_wait(&bar_never, 0) ? 1 : 0;
      } else if (noise == 4) {   // MUFU
        float f = __uint_as_float(0x3f800000u + acc);
        for (int i = 0; i < 16; ++i) f = ex2_approx(f * 0.5f);
        acc += __float_as_uint(f) & 1;
      } else if (noise == 5) {   // bulk global->shared copies (TMA engine writes into smem), 16 KB each, warp 4 only
        if (warp == 4 && lane == 0) {
          mbar_arrive_expect_tx(&bar_tma, 16384);
          bulk_load(sN, gsrc + (acc & 15) * 16384, 16384, &bar_tma);
          mbar_wait(&bar_tma, tparity); tparity ^= 1; acc++;
        }
      } else if (noise == 6) {   // TMEM stores
        for (int c = 0; c < 32; c += 16) {
          uint32_t r[16];
          for (int j = 0; j < 16; ++j) r[j] = acc + j;
          tmem_st16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 496 - 32 + c, r);
          tmem_st_wait();
        }
      }
    }
    if (acc == 0xdeadbeef) out[31] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
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
  {
    long long* dT; cudaMalloc(&dT, 32 * 8); cudaMemset(dT, 0, 32 * 8);
    cudaFuncSetAttribute(probe_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    probe_time<<<1, 128, 100 * 1024>>>(dT);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("timing: CUDA error %s\n", cudaGetErrorString(e)); return 3; }
    long long T[32]; cudaMemcpy(T, dT, sizeof(T), cudaMemcpyDeviceToHost);
    const char* tn[7] = {"G  SS K-SW128 x K-SW128 N=128", "O  SS K-SW128 x MN-SW32 N=80", "D  TS TMEM x MN-SW32 N=80",
                         "S  SS MN-SW128 x MN-SW32 N=80", "O' SS K-SW128 x MN-SW128 N=80", "D' TS TMEM x MN-SW128 N=80",
                         "S' SS MN-SW128 x MN-SW128 N=80"};
    for (int i = 0; i < 7; ++i) printf("cycles per 8-MMA group [%s]: %lld (issue only: %lld)\n", tn[i], T[i], T[16 + i]);
    printf("cycles for 10 x (tcgen05.ld x16 + wait), 4 warps: %lld\n", T[8]);
  }
  {
    long long* dT; cudaMalloc(&dT, 32 * 8); cudaMemset(dT, 0, 32 * 8);
    uint8_t* dsrc; cudaMalloc(&dsrc, 16 * 16384); cudaMemset(dsrc, 1, 16 * 16384);
    cudaFuncSetAttribute(probe_contention, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    const char* nn[7] = {"none", "LDS/STS 16B", "tcgen05.ld", "mbarrier.try_wait polling", "MUFU", "bulk copy -> smem", "tcgen05.st"};
    for (int noise = 0; noise < 7; ++noise) {
      probe_contention<<<1, 256, 170 * 1024>>>(dT, dsrc, noise);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("contention %d: CUDA error %s\n", noise, cudaGetErrorString(e)); return 3; }
    }
    long long T[32]; cudaMemcpy(T, dT, sizeof(T), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 7; ++i) printf("cycles per S+D+G+O round with background [%s]: %lld\n", nn[i], T[i]);
  }
  return fails ? 1 : 0;
}

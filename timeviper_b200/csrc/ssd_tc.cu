// Fused SSD chunked scan on tcgen05 / TMEM / TMA (sm_100a) for the Nanov2-9B Mamba-2 geometry:
// bf16, headdim P = 80, dstate N = 128, chunk Q = 128 (any nheads / ngroups / batch / seqlen).
//
// Replaces the five mamba_ssm Triton kernels behind mamba_chunk_scan_combined
// (visualize/nano/my_ssd_combined.py:795-826) with ONE persistent kernel.  One CTA owns one (batch, head)
// and walks the chunks in order with the 128x80 fp32 running state resident in TMEM, so the per-chunk
// states / C.B^T tiles that the reference materialises in HBM (~123 KB per token) never leave the SM:
// HBM traffic is the algorithmic 45,312 B/token (x, B, C, dt in; y out).
//
// Per chunk c (m, k: tokens in the chunk; n: state; p: head dim), with cs = inclusive cumsum of dt*A:
//   G: CB[m,k]   = sum_n C[m,n] B[k,n]                                   tcgen05 SS, 128x128x128 -> TMEM
//      M[m,k]    = CB[m,k] * exp(cs_m - cs_k) * dt_k  (k <= m)           WG_A: TMEM -> regs -> bf16 -> TMEM
//   D: Yd[m,p]   = sum_k M[m,k] x[k,p]                                   tcgen05 TS (A = M in TMEM), N = 80
//   O: Yo[m,p]   = sum_n C[m,n] S_c[n,p]          (S_c = state entering the chunk, bf16 copy in smem)
//   S: S_{c+1}   = exp(cs_last) * S_c + sum_k B[k,n] * (dt_k exp(cs_last - cs_k) x[k,p])
//                  decay: WG_B TMEM -> regs -> TMEM; the sum: tcgen05 SS accumulating onto it
//      y[m,p]    = Yd + exp(cs_m) * Yo + D * x[m,p]   [* silu(z)]         WG_B epilogue -> global
// Decay, mask and cumsum stay in fp32 registers; only the four contractions touch the tensor cores.
//
// Warp roles (320 threads): warps 0-3 = WG_A (builds M), warps 4-7 = WG_B (x scaling, state decay, bf16 state
// copy, epilogue), warp 8 = TMA producer, warp 9 = MMA issuer + TMEM owner.  Tensor-pipe order per
// iteration is S(c), G(c+1), O(c), D(c): the state recurrence (the only loop-carried dependency) is issued
// first, and the look-ahead C.B^T hides the M build and the epilogue of the previous chunk.
#include "common.cuh"
#include "sm100.cuh"
#include "ssd.h"
#include "tmap.h"

namespace tv {
using namespace sm100;

namespace tc {
constexpr int Q = 128, P = 80, N = 128;
constexpr int THREADS = 320;
constexpr uint32_t TILE_BC = Q * N * 2;          // 32768: two 16 KB halves (n 0..63 | 64..127), SW128
constexpr uint32_t TILE_X = 5 * 4096;            // 20480: five 16-wide p atoms, SW32
constexpr uint32_t OFF_B = 0, OFF_C = TILE_BC, OFF_X = 2 * TILE_BC, OFF_CS = OFF_X + TILE_X, OFF_DT = OFF_CS + 512;
constexpr uint32_t STAGE = OFF_DT + 512;          // 87040
constexpr uint32_t OFF_XS = 2 * STAGE, OFF_S = OFF_XS + TILE_X, OFF_F = OFF_S + TILE_X;   // F: 2 x 512 B
constexpr uint32_t OFF_D = OFF_F + 1024;          // 80 floats (D row), padded to 512
constexpr uint32_t OFF_BAR = OFF_D + 512;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;   // + barriers + alignment slack
static_assert(STAGE % 1024 == 0, "stage must keep 1024-byte alignment");
static_assert(SMEM_BYTES <= kMaxDynSmem, "smem budget");
// TMEM columns
constexpr uint32_t T_CB0 = 0, T_CB1 = 128, T_YD = 256, T_YO = 336, T_ST = 416;

enum Bar { FULL0 = 0, FULL1, EMPTY0, EMPTY1, CBFULL0, CBFULL1, MFULL0, MFULL1, XSFULL, SDECAY, SFULL, STDONE,
           YOFFDONE, YFULL, YEMPTY, NBAR };

struct Maps { CUtensorMap x, b, c; };

struct Args {
  const float* dt_act; const float* cs;            // (b, nchunks, H, Q) fp32
  const float* D; const __nv_bfloat16* z; const float* init;
  __nv_bfloat16* out; float* fin; float* logdecay;
  int L, H, G, nchunks, d_has_hdim;
  int64_t zbs, zss, zhs;
};

__device__ __forceinline__ uint32_t off_sw32(int r, int q) {  // row r, 16-byte chunk q (8 p each) of a [128][80] tile
  return (uint32_t)(q >> 1) * 4096u + (uint32_t)r * 32u + (uint32_t)(((q & 1) ^ ((r >> 2) & 1)) << 4);
}
}  // namespace tc

template <bool FULL>
__global__ void __launch_bounds__(tc::THREADS, 1)
ssd_fused_kernel(const __grid_constant__ tc::Maps maps, const tc::Args a) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBAR * 8);
  const int warp = threadIdx.x >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  const int g = h / (a.H / a.G);
  const int n = a.nchunks;

  if (threadIdx.x == 0) {
    mbar_init(&bars[FULL0], 1); mbar_init(&bars[FULL1], 1);
    mbar_init(&bars[EMPTY0], FULL ? 257 : 129); mbar_init(&bars[EMPTY1], FULL ? 257 : 129);
    mbar_init(&bars[CBFULL0], 1); mbar_init(&bars[CBFULL1], 1);
    mbar_init(&bars[MFULL0], 128); mbar_init(&bars[MFULL1], 128);
    mbar_init(&bars[XSFULL], 128); mbar_init(&bars[SDECAY], 128); mbar_init(&bars[SFULL], 128);
    mbar_init(&bars[STDONE], 1); mbar_init(&bars[YOFFDONE], 1); mbar_init(&bars[YFULL], 1);
    mbar_init(&bars[YEMPTY], 128);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  if (threadIdx.x < P) {
    float* sD = reinterpret_cast<float*>(smem + OFF_D);
    sD[threadIdx.x] = a.D == nullptr ? 0.f : (a.d_has_hdim ? a.D[h * P + threadIdx.x] : a.D[h]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t row0 = ((int64_t)b * n) * a.H + h;     // (b, c, h) row of dt_act / cs is row0 + c*H

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      prefetch_tmap(&maps.x); prefetch_tmap(&maps.b);
      if (FULL) prefetch_tmap(&maps.c);
      for (int c = 0; c < n; ++c) {
        const int s = c & 1;
        uint8_t* st = smem + s * STAGE;
        if (c >= 2) mbar_wait(&bars[EMPTY0 + s], ((c >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&bars[FULL0 + s], (FULL ? 2 * TILE_BC : TILE_BC) + TILE_X + 1024);
        const int t0 = c * Q;
        tma_load_4d(st + OFF_B, &maps.b, &bars[FULL0 + s], 0, g, t0, b);
        tma_load_4d(st + OFF_B + 16384, &maps.b, &bars[FULL0 + s], 64, g, t0, b);
        if (FULL) {
          tma_load_4d(st + OFF_C, &maps.c, &bars[FULL0 + s], 0, g, t0, b);
          tma_load_4d(st + OFF_C + 16384, &maps.c, &bars[FULL0 + s], 64, g, t0, b);
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) tma_load_4d(st + OFF_X + i * 4096, &maps.x, &bars[FULL0 + s], 16 * i, h, t0, b);
        bulk_load(st + OFF_CS, a.cs + (row0 + (int64_t)c * a.H) * Q, 512, &bars[FULL0 + s]);
        bulk_load(st + OFF_DT, a.dt_act + (row0 + (int64_t)c * a.H) * Q, 512, &bars[FULL0 + s]);
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      constexpr uint32_t ID_CB = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t ID_Y = umma_idesc_bf16(128, P, false, true);
      constexpr uint32_t ID_ST = umma_idesc_bf16(128, P, true, true);
      const uint32_t sbase = smem_u32(smem);
      auto issue_cb = [&](int c) {   // G(c): CB = C . B^T
        const int s = c & 1;
        mbar_wait(&bars[FULL0 + s], (c >> 1) & 1);
        tc_fence_after();
        const uint32_t sc = sbase + s * STAGE + OFF_C, sb = sbase + s * STAGE + OFF_B;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t o = (j >> 2) * 16384 + (j & 3) * 32;
          umma_ss(tmem + (s ? T_CB1 : T_CB0), umma_smem_desc(sc + o, 16, 1024, SWZ_128B),
                  umma_smem_desc(sb + o, 16, 1024, SWZ_128B), ID_CB, j > 0);
        }
        umma_commit(&bars[CBFULL0 + s]);
      };
      if (FULL) issue_cb(0);
      for (int c = 0; c < n; ++c) {
        const int s = c & 1;
        const uint32_t st = sbase + s * STAGE;
        // ---- S(c): state += B^T . xs
        mbar_wait(&bars[FULL0 + s], (c >> 1) & 1);
        mbar_wait(&bars[SDECAY], c & 1);
        mbar_wait(&bars[XSFULL], c & 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_ss(tmem + T_ST, umma_smem_desc(st + OFF_B + j * 2048, 16384, 1024, SWZ_128B),
                  umma_smem_desc(sbase + OFF_XS + j * 512, 4096, 256, SWZ_32B), ID_ST, 1u);
        umma_commit(&bars[STDONE]);
        if (FULL) {
          // ---- G(c+1)
          if (c + 1 < n) issue_cb(c + 1);
          // ---- O(c): Yo = C . S_c
          mbar_wait(&bars[SFULL], c & 1);
          if (c > 0) mbar_wait(&bars[YEMPTY], (c - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t o = (j >> 2) * 16384 + (j & 3) * 32;
            umma_ss(tmem + T_YO, umma_smem_desc(st + OFF_C + o, 16, 1024, SWZ_128B),
                    umma_smem_desc(sbase + OFF_S + j * 512, 4096, 256, SWZ_32B), ID_Y, j > 0);
          }
          umma_commit(&bars[YOFFDONE]);
          // ---- D(c): Yd = M . x   (A = M, packed bf16 in the first 64 columns of this chunk's CB buffer)
          mbar_wait(&bars[MFULL0 + s], (c >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_ts(tmem + T_YD, tmem + (s ? T_CB1 : T_CB0) + j * 8,
                    umma_smem_desc(st + OFF_X + j * 512, 4096, 256, SWZ_32B), ID_Y, j > 0);
          umma_commit(&bars[YFULL]);
        }
        umma_commit(&bars[EMPTY0 + s]);   // every MMA that reads stage s has been issued
      }
    }
  } else if (warp < 4) {
    // =========================== WG_A: M = CB (.) decay, in place in TMEM ===========================
    if (FULL) {
      const int m = threadIdx.x;
      const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
      constexpr float LOG2E = 1.4426950408889634f;
      for (int c = 0; c < n; ++c) {
        const int s = c & 1;
        const float* sCS = reinterpret_cast<const float*>(smem + s * STAGE + OFF_CS);
        const float* sDT = reinterpret_cast<const float*>(smem + s * STAGE + OFF_DT);
        float* sF = reinterpret_cast<float*>(smem + OFF_F + s * 512);
        mbar_wait(&bars[FULL0 + s], (c >> 1) & 1);
        const float Em = sCS[m] * LOG2E;
        sF[m] = __log2f(sDT[m]) - Em;            // F_k = log2(dt_k) - cs_k*log2e  (dt = 0 -> -inf -> weight 0)
        named_bar_sync(1, 128);
        mbar_arrive(&bars[EMPTY0 + s]);          // cs/dt of this stage are consumed (F lives in its own buffer)
        mbar_wait(&bars[CBFULL0 + s], (c >> 1) & 1);
        tc_fence_after();
        const uint32_t tcb = tmem + (s ? T_CB1 : T_CB0) + lane_base;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          uint32_t pk[16];
          if (kb > warp) {                       // whole 32x32 block above the diagonal
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          } else {
            uint32_t r[32];
            tmem_ld32(tcb + kb * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 f4 = *reinterpret_cast<const float4*>(sF + kb * 32 + j);
              const float f[4] = {f4.x, f4.y, f4.z, f4.w};
              float v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int k = kb * 32 + j + i;
                const float e = exp2f(Em + f[i]);
                v[i] = (k <= m) ? __uint_as_float(r[j + i]) * e : 0.f;
              }
              pk[(j >> 1)] = pack_bf16x2(v[0], v[1]);
              pk[(j >> 1) + 1] = pack_bf16x2(v[2], v[3]);
            }
          }
          tmem_st16(tcb + kb * 16, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[MFULL0 + s]);
      }
    }
  } else {
    // =========================== WG_B: xs, state decay / bf16 copy, epilogue ===========================
    const int r = threadIdx.x - 128;             // token row (x, y) and state row n
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float* sD = reinterpret_cast<const float*>(smem + OFF_D);
    uint32_t xkeep[40];                           // x row of the previous chunk (bf16x2), for D*x in its epilogue
    float e_keep = 0.f;
    float logsum = 0.f;
#pragma unroll
    for (int i = 0; i < 40; ++i) xkeep[i] = 0u;

    auto epilogue = [&](int c) {                  // y rows of chunk c
      mbar_wait(&bars[YFULL], c & 1);
      tc_fence_after();
      uint32_t yo_pk[40];
      const int t = c * Q + r;
#pragma unroll
      for (int pc = 0; pc < 5; ++pc) {
        uint32_t yd[16], yo[16];
        tmem_ld16(tmem + T_YD + lane_base + pc * 16, yd);
        tmem_ld16(tmem + T_YO + lane_base + pc * 16, yo);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const uint32_t xp = xkeep[pc * 8 + (j >> 1)];
          const float x0 = __uint_as_float(xp << 16), x1 = __uint_as_float(xp & 0xffff0000u);
          float y0 = fmaf(e_keep, __uint_as_float(yo[j]), __uint_as_float(yd[j]));
          float y1 = fmaf(e_keep, __uint_as_float(yo[j + 1]), __uint_as_float(yd[j + 1]));
          y0 = fmaf(sD[pc * 16 + j], x0, y0);
          y1 = fmaf(sD[pc * 16 + j + 1], x1, y1);
          if (a.z != nullptr && t < a.L) {
            const __nv_bfloat16* zp = a.z + b * a.zbs + (int64_t)t * a.zss + (int64_t)h * a.zhs + pc * 16 + j;
            y0 *= silu<true>(__bfloat162float(zp[0]));
            y1 *= silu<true>(__bfloat162float(zp[1]));
          }
          yo_pk[pc * 8 + (j >> 1)] = pack_bf16x2(y0, y1);
        }
      }
      tc_fence_before();
      mbar_arrive(&bars[YEMPTY]);                 // Yd / Yo may be overwritten
      if (t < a.L) {
        uint4* op = reinterpret_cast<uint4*>(a.out + (((int64_t)b * a.L + t) * a.H + h) * P);
#pragma unroll
        for (int q = 0; q < 10; ++q) op[q] = make_uint4(yo_pk[4 * q], yo_pk[4 * q + 1], yo_pk[4 * q + 2], yo_pk[4 * q + 3]);
      }
    };

    for (int c = 0; c < n; ++c) {
      const int s = c & 1;
      const uint8_t* st = smem + s * STAGE;
      const float* sCS = reinterpret_cast<const float*>(st + OFF_CS);
      const float* sDT = reinterpret_cast<const float*>(st + OFF_DT);
      mbar_wait(&bars[FULL0 + s], (c >> 1) & 1);
      const float cs_last = sCS[Q - 1], cs_r = sCS[r];
      const float w_r = sDT[r] * __expf(cs_last - cs_r);
      const float e_r = __expf(cs_r), a_c = __expf(cs_last);
      logsum += cs_last;
      // ---- x row -> registers; xs = x * w_r
      uint32_t xcur[40], xs[40];
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const uint4 v = *reinterpret_cast<const uint4*>(st + OFF_X + off_sw32(r, q));
        xcur[4 * q] = v.x; xcur[4 * q + 1] = v.y; xcur[4 * q + 2] = v.z; xcur[4 * q + 3] = v.w;
      }
      mbar_arrive(&bars[EMPTY0 + s]);             // smem of this stage consumed by WG_B
#pragma unroll
      for (int i = 0; i < 40; ++i)
        xs[i] = pack_bf16x2(__uint_as_float(xcur[i] << 16) * w_r, __uint_as_float(xcur[i] & 0xffff0000u) * w_r);
      // ---- previous state MMA done: xs buffer is free and S_c (entering state) is complete in TMEM
      if (c > 0) { mbar_wait(&bars[STDONE], (c - 1) & 1); tc_fence_after(); }
#pragma unroll
      for (int q = 0; q < 10; ++q)
        *reinterpret_cast<uint4*>(smem + OFF_XS + off_sw32(r, q)) = make_uint4(xs[4 * q], xs[4 * q + 1], xs[4 * q + 2], xs[4 * q + 3]);
      fence_proxy_async();
      mbar_arrive(&bars[XSFULL]);
      // ---- state row n = r: decay in place (critical path), keep a bf16 copy of the un-decayed entering state
      uint32_t spk[40];
#pragma unroll
      for (int pc = 0; pc < 5; ++pc) {
        uint32_t v[16];
        if (c == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = a.init == nullptr ? 0u
                                     : __float_as_uint(a.init[(((int64_t)b * a.H + h) * P + pc * 16 + j) * N + r]);
        } else {
          tmem_ld16(tmem + T_ST + lane_base + pc * 16, v);
          tmem_ld_wait();
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2)
          spk[pc * 8 + (j >> 1)] = pack_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * a_c);
        tmem_st16(tmem + T_ST + lane_base + pc * 16, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars[SDECAY]);
      if (FULL) {
        if (c > 0) mbar_wait(&bars[YOFFDONE], (c - 1) & 1);    // O(c-1) finished reading the bf16 state copy
#pragma unroll
        for (int q = 0; q < 10; ++q)
          *reinterpret_cast<uint4*>(smem + OFF_S + off_sw32(r, q)) = make_uint4(spk[4 * q], spk[4 * q + 1], spk[4 * q + 2], spk[4 * q + 3]);
        fence_proxy_async();
        mbar_arrive(&bars[SFULL]);
        if (c > 0) epilogue(c - 1);
#pragma unroll
        for (int i = 0; i < 40; ++i) xkeep[i] = xcur[i];
        e_keep = e_r;
      }
    }
    if (FULL) epilogue(n - 1);
    // ---- final state = state after the last chunk
    mbar_wait(&bars[STDONE], (n - 1) & 1);
    tc_fence_after();
    if (a.fin != nullptr) {
#pragma unroll
      for (int pc = 0; pc < 5; ++pc) {
        uint32_t v[16];
        tmem_ld16(tmem + T_ST + lane_base + pc * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          a.fin[(((int64_t)b * a.H + h) * P + pc * 16 + j) * N + r] = __uint_as_float(v[j]);
      }
    }
    if (a.logdecay != nullptr && r == 0) a.logdecay[(int64_t)b * a.H + h] = logsum;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------ host
bool tc_supported(const tv_ssd_params& p) {
  if (p.dtype != TV_BF16 || p.headdim != tc::P || p.dstate != tc::N || p.chunk_size != tc::Q) return false;
  if (p.nheads % p.ngroups != 0) return false;
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  if (!al16(p.x) || !al16(p.B) || (p.mode == TV_SSD_FULL && (!al16(p.C) || !al16(p.out)))) return false;
  auto m8 = [](int64_t v) { return v % 8 == 0; };
  if (!m8(p.x_head_stride) || !m8(p.x_seq_stride) || !m8(p.x_batch_stride) || !m8(p.b_group_stride) ||
      !m8(p.b_seq_stride) || !m8(p.b_batch_stride))
    return false;
  if (p.mode == TV_SSD_FULL && (!m8(p.c_group_stride) || !m8(p.c_seq_stride) || !m8(p.c_batch_stride))) return false;
  return get_encode_tiled() != nullptr;
}

size_t tc_workspace_bytes(const tv_ssd_params& p) {
  const int64_t nchunks = ceil_div(p.seqlen, p.chunk_size);
  const size_t per = (size_t)p.batch * nchunks * p.nheads * p.chunk_size * sizeof(float);
  return 2 * ((per + 255) & ~(size_t)255);
}

int ssd_tc_forward(const tv_ssd_params& p, void* workspace, cudaStream_t s) {
  using namespace tc;
  const int nchunks = (int)ceil_div(p.seqlen, Q);
  const size_t per = (((size_t)p.batch * nchunks * p.nheads * Q * sizeof(float)) + 255) & ~(size_t)255;
  float* dt_act = (float*)workspace;
  float* cs = (float*)((char*)workspace + per);
  int rc = launch_dt_cumsum(p, dt_act, cs, s);
  if (rc != TV_OK) return rc;

  Maps maps;
  const uint64_t L = (uint64_t)p.seqlen;
  auto bstride = [&](int64_t bs, int64_t ss) { return (uint64_t)(p.batch == 1 ? ss * (int64_t)L : bs) * 2; };
  {
    const uint64_t d[4] = {(uint64_t)P, (uint64_t)p.nheads, L, (uint64_t)p.batch};
    const uint64_t st[3] = {(uint64_t)p.x_head_stride * 2, (uint64_t)p.x_seq_stride * 2,
                            bstride(p.x_batch_stride, p.x_seq_stride)};
    const uint32_t box[4] = {16, 1, (uint32_t)Q, 1};
    if (!encode_bf16_tmap(&maps.x, p.x, 4, d, st, box, CU_TENSOR_MAP_SWIZZLE_32B)) {
      set_error("ssd(tcgen05): cuTensorMapEncodeTiled(x) failed");
      return TV_ERR_CUDA;
    }
  }
  {
    const uint64_t d[4] = {(uint64_t)N, (uint64_t)p.ngroups, L, (uint64_t)p.batch};
    const uint32_t box[4] = {64, 1, (uint32_t)Q, 1};
    const uint64_t sb[3] = {(uint64_t)p.b_group_stride * 2, (uint64_t)p.b_seq_stride * 2,
                            bstride(p.b_batch_stride, p.b_seq_stride)};
    if (!encode_bf16_tmap(&maps.b, p.B, 4, d, sb, box, CU_TENSOR_MAP_SWIZZLE_128B)) {
      set_error("ssd(tcgen05): cuTensorMapEncodeTiled(B) failed");
      return TV_ERR_CUDA;
    }
    if (p.mode == TV_SSD_FULL) {
      const uint64_t sc[3] = {(uint64_t)p.c_group_stride * 2, (uint64_t)p.c_seq_stride * 2,
                              bstride(p.c_batch_stride, p.c_seq_stride)};
      if (!encode_bf16_tmap(&maps.c, p.C, 4, d, sc, box, CU_TENSOR_MAP_SWIZZLE_128B)) {
        set_error("ssd(tcgen05): cuTensorMapEncodeTiled(C) failed");
        return TV_ERR_CUDA;
      }
    } else {
      maps.c = maps.b;
    }
  }
  Args a;
  a.dt_act = dt_act; a.cs = cs; a.D = p.D; a.z = (const __nv_bfloat16*)p.z; a.init = p.initial_states;
  a.out = (__nv_bfloat16*)p.out; a.fin = p.final_states; a.logdecay = p.logdecay_sum;
  a.L = p.seqlen; a.H = p.nheads; a.G = p.ngroups; a.nchunks = nchunks; a.d_has_hdim = p.d_has_hdim;
  a.zbs = p.z_batch_stride; a.zss = p.z_seq_stride; a.zhs = p.z_head_stride;
  dim3 grid(p.nheads, p.batch);
  if (p.mode == TV_SSD_FULL) {
    TV_CUDA_OK(cudaFuncSetAttribute(ssd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    ssd_fused_kernel<true><<<grid, THREADS, SMEM_BYTES, s>>>(maps, a);
  } else {
    TV_CUDA_OK(cudaFuncSetAttribute(ssd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    ssd_fused_kernel<false><<<grid, THREADS, SMEM_BYTES, s>>>(maps, a);
  }
  TV_CUDA_OK(cudaGetLastError());
  return TV_OK;
}

}  // namespace tv

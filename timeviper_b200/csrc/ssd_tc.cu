// Fused SSD chunked scan on tcgen05 / TMEM / TMA (sm_100a) for the Nanov2-9B Mamba-2 geometry:
// bf16, headdim P = 80, dstate N = 128, chunk Q = 128 (any nheads / ngroups / batch / seqlen).
//
// Replaces the five mamba_ssm Triton kernels behind mamba_chunk_scan_combined
// (visualize/nano/my_ssd_combined.py:795-826) with ONE persistent kernel.  One CTA owns one (batch, head)
// and walks the chunks in order; the per-chunk states / C.B^T tiles that the reference materialises in HBM
// (~123 KB per token) never leave the SM: HBM traffic is the algorithmic 45,312 B/token (x, B, C, dt in; y out).
//
// Per chunk c (m, k: tokens in the chunk; n: state; p: head dim), with cs = inclusive cumsum of dt*A:
//   G: CB[m,k]   = sum_n C[m,n] B[k,n]                                   tcgen05 SS, 128x128x128 -> TMEM (2 buffers)
//      M[m,k]    = CB[m,k] * exp(cs_m - cs_k) * dt_k  (k <= m)           WG_A + helper: TMEM -> regs -> bf16 -> TMEM,
//                                                                         IN PLACE over the first 64 columns of CB
//   D: Yd[m,p]   = sum_k M[m,k] x[k,p]                                   tcgen05 TS (A = M in TMEM), N = 80
//   O: Yo[m,p]   = sum_n C[m,n] S_c[n,p]          (S_c = state entering the chunk, bf16 copy in smem)
//   S: dS[n,p]   = sum_k B[k,n] * (dt_k exp(cs_last - cs_k) x[k,p])      tcgen05 SS into a FRESH accumulator
//      S_{c+1}   = exp(cs_last) * S_c + dS                               WG_S: the running state lives in REGISTERS
//      y[m,p]    = Yd + exp(cs_m) * Yo + D * x[m,p]   [* silu(z)]         WG_C: TMEM -> regs -> bf16, each warp through
//                                                                         its own staging area -> its own TMA stores
// Decay, mask and cumsum stay in fp32 registers; only the four contractions touch the tensor cores.
//
// Round-2 dataflow (why it looks like this: profiles/r02_ssd_restructure.md, profiles/r02_ssd_v5.md).  The round-1 kernel
// kept the state in a TMEM accumulator that the S MMA accumulated onto, so every chunk paid MMA -> commit -> TMEM load /
// decay / store -> MMA as a serial chain, and one in-order issuing thread coupled that chain to the C.B^T -> M -> D chain.
//   * The S MMA never accumulates across chunks: the recurrence is 80 FMAs per thread in WG_S's registers, everything
//     else is feed-forward.
//   * C.B^T is double-buffered in TMEM and M overwrites it in place, so G / M run ahead of the state path; the M builders
//     take cs / dt from L2 one chunk ahead (not from the x stage), so M(c) is ready when x(c) lands.
//   * ONE MMA-issuing thread and one TMA thread, both POLLING (mbarrier.test_wait) their streams / rings instead of
//     waiting in program order; a group of 8 MMAs is always issued back to back.
//   * The x stage goes back to the TMA producer with the D commit: y is staged per epilogue warp in its own 3 KB.
//   * Register budgets per role via setmaxnreg (WG_S holds the 128x80 fp32 state: 80 registers per thread).
//
// Warp roles (640 threads): warps 0-3 = WG_A (diagonal M blocks + one off-diagonal block), warps 4-7 = WG_X (x scaling
// for the S MMA), warps 8-11 = WG_S (state), warps 12-15 = WG_C (epilogue), warp 16 = TMA producer, warp 17 = MMA issuer
// + TMEM owner, warps 18 / 19 = M helpers (off-diagonal blocks 0 / 1 of row quarters 2 / 3).
#include "common.cuh"
#include "sm100.cuh"
#include "ssd.h"
#include "tmap.h"

#include <algorithm>
#include <mutex>
#include <unordered_map>

namespace tv {
using namespace sm100;

namespace tc {
constexpr int Q = 128, P = 80, N = 128;
// Back-off of the two polling threads after a pass that found nothing to do (ns; 0 = spin).  Every probe is a shared-memory
// operation, and the shared-memory pipe is what bounds this kernel.
#ifndef TV_POLL_NS_PROD
#define TV_POLL_NS_PROD 0
#endif
#ifndef TV_POLL_NS_MMA
#define TV_POLL_NS_MMA 0
#endif
constexpr int THREADS = 640;
constexpr int W_A = 0, W_X = 4, W_S = 8, W_C = 12, W_PROD = 16, W_MMA = 17;   // first warp of each role (18 / 19: M helpers)
// Registers per thread after setmaxnreg.  The pool is what the CTA was LAUNCHED with (640 threads x 96 registers, the
// most __launch_bounds__(640) allows), not the SM's register file: the five warpgroups must sum to 5 x 96 = 480.
constexpr int REG_LAUNCH = 96;
constexpr int REG_A = 104, REG_X = 64, REG_S = 136, REG_C = 96, REG_MISC = 80;
static_assert(REG_A + REG_X + REG_S + REG_C + REG_MISC <= 5 * REG_LAUNCH, "register pool of the CTA");
template <int R> __device__ __forceinline__ void reg_set() {      // executed by all four warps of a warpgroup
  if (R > REG_LAUNCH) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R));
  if (R < REG_LAUNCH) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R));
}
constexpr uint32_t TILE_BC = Q * N * 2;          // 32768: two 16 KB halves (n 0..63 | 64..127), SW128
constexpr uint32_t TILE_X = 5 * 4096;            // 20480: five 16-wide p atoms, SW32
constexpr uint32_t XSTAGE = TILE_X + 1024;       // x tile | cs[128] f32 | dt[128] f32
constexpr uint32_t OFF_B = 0, OFF_C = 2 * TILE_BC, OFF_X = 4 * TILE_BC;          // two buffers of each
constexpr uint32_t OFF_XS = OFF_X + 2 * XSTAGE, OFF_S = OFF_XS + TILE_X, OFF_Y = OFF_S + TILE_X;
constexpr uint32_t YSTG = 3 * 1024;              // per epilogue warp: three 32-row x 16-column SW32 atoms of staged y
constexpr uint32_t OFF_SCR = OFF_Y + 4 * YSTG;
constexpr uint32_t SCR = 512;                    // per-warp fp32 scratch of the 6 M-building warps (decay factors)
constexpr uint32_t OFF_D = OFF_SCR + 6 * SCR;    // 80 floats (D row), padded to 512
constexpr uint32_t OFF_BAR = OFF_D + 512;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 512 + 1024;   // + barriers + alignment slack
static_assert(OFF_XS % 1024 == 0 && OFF_S % 1024 == 0 && XSTAGE % 256 == 0, "tile alignment");
static_assert(SMEM_BYTES <= kMaxDynSmem, "smem budget");
// TMEM columns: C.B^T / M (two buffers; M = 64 columns of packed bf16 written over columns 0..63), Yd, Yo, dS
constexpr uint32_t T_CB0 = 0, T_CB1 = 128, T_YD = 256, T_YO = 336, T_DS = 416;

enum Bar { FULLB0 = 0, FULLB1, EMPTYB0, EMPTYB1, FULLC0, FULLC1, EMPTYC0, EMPTYC1, FULLX0, FULLX1, EMPTYX0, EMPTYX1,
           CBFULL0, CBFULL1, MFULL0, MFULL1, HREAD0, HREAD1, XSFULL, STDONE, DSFREE, SFULL, YOFULL, YDFULL, YDFREE, YOFREE,
           NBAR };
static_assert(NBAR * 8 + 8 <= 512, "barrier area");

struct Maps { CUtensorMap x, b, c, y; };

struct Args {
  const float* dt_act; const float* cs;            // (b, nchunks, H, Q) fp32
  const float* D; const __nv_bfloat16* z; const float* init;
  __nv_bfloat16* out; float* fin; float* logdecay;
  int L, H, G, nchunks, d_has_hdim;
  int64_t zbs, zss, zhs;
  const __nv_bfloat16 *xp, *bp, *cp;               // raw pointers / element strides of x, B, C: L2 prefetch warp only
  int64_t xbs, xss, xhs, bbs, bss, bgs, cbs, css, cgs;
  long long* trace;                                // optional: per-chunk clock64 stamps of CTA (0,0), 16 per chunk
  int ablate;                                      // profiling only (TV_ENABLE_TRACE builds): bitmask of work to skip
};

#ifdef TV_ENABLE_TRACE   // profiling builds only (python timeviper_b200/build.py --trace): keeps the hot loops small
#define TV_TRACE(ev, c)                                                                    \
  do {                                                                                      \
    if (a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) a.trace[(int64_t)(c) * 16 + (ev)] = clock64(); \
  } while (0)
#else
#define TV_TRACE(ev, c) do { } while (0)
#endif
// Ablation switches (variant builds only, tools/build_variants.py): the kernel with parts of the work skipped, to find
// the critical path; in normal builds TV_ABLATE() is the constant 0 and the branches vanish.
#if defined(TV_ABL)      // compile-time mask (tools/build_variants.py): no register or branch cost, unlike the trace build
#define TV_ABLATE(bit) (((TV_ABL) >> (bit)) & 1)
#else
#define TV_ABLATE(bit) 0
#endif
__device__ __forceinline__ uint32_t off_sw32(int r, int q) {  // row r, 16-byte chunk q (8 p each) of a [128][80] tile
  return (uint32_t)(q >> 1) * 4096u + (uint32_t)r * 32u + (uint32_t)(((q & 1) ^ ((r >> 2) & 1)) << 4);
}
}  // namespace tc

// One diagonal 32x32 block of M for this warp's 32 rows: M[m,k] = CB[m,k] * 2^(Em + F_k), masked to k <= m.  The packed
// bf16 result stays in registers (the caller stores it once the columns it overwrites have been read by everybody).
// Written stage by stage (TMEM load, 32 exponent arguments, 32 MUFU.EX2, 32 FMUL, 16 packs) so that the exp2 stream is
// issue-bound on the XU pipe (8 cycles per warp instruction) instead of latency-bound.
template <bool DFOLD>
__device__ __forceinline__ void m_diag(uint32_t t_src, uint32_t (&pk)[16], const float* __restrict__ sFk, float Em,
                                       int lane, float Dh) {
  uint32_t r[32];
  tmem_ld32(t_src, r);
  float2 e[16];                                  // packed pairs: FADD2 / FMUL2 halve the FP issue slots
  const float2 Em2 = make_float2(Em, Em);
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float4 f4 = *reinterpret_cast<const float4*>(sFk + 2 * j);
    e[j] = __fadd2_rn(Em2, make_float2(f4.x, f4.y));
    e[j + 1] = __fadd2_rn(Em2, make_float2(f4.z, f4.w));
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) e[j] = make_float2(ex2_approx(e[j].x), ex2_approx(e[j].y));
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float2 v = __fmul2_rn(e[j], make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])));
    if (2 * j > lane) v.x = 0.f;                   // causal mask
    if (2 * j + 1 > lane) v.y = 0.f;
    if (DFOLD && 2 * j == lane) v.x += Dh;         // D skip folded into the diagonal
    if (DFOLD && 2 * j + 1 == lane) v.y += Dh;
    pk[j] = pack_bf16x2(v.x, v.y);
  }
}

// Diagonal block without transcendentals: exp(cs_m - cs_k) dt_k = u_m * v_k around ref = cs just before the block's rows,
// u_m = exp(cs_m - ref) <= 1, v_k = dt_k exp(ref - cs_k) >= dt_k.  v_k grows with the decay inside the block, so the caller
// takes this path only while the block's total decay keeps exp(ref - cs_k) far from overflow (masked entries k > m are
// finite and dropped by a select, never multiplied by zero); otherwise m_diag with its 32 exponentials per row.
template <bool DFOLD>
__device__ __forceinline__ void m_diag_fast(uint32_t t_src, uint32_t (&pk)[16], const float* __restrict__ sVk, float um,
                                            int lane, float Dh) {
  uint32_t r[32];
  tmem_ld32(t_src, r);
  const float2 um2 = make_float2(um, um);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float4 v4 = *reinterpret_cast<const float4*>(sVk + 2 * j);
    const float2 e0 = __fmul2_rn(make_float2(v4.x, v4.y), um2), e1 = __fmul2_rn(make_float2(v4.z, v4.w), um2);
    float2 a0 = __fmul2_rn(e0, make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])));
    float2 a1 = __fmul2_rn(e1, make_float2(__uint_as_float(r[2 * j + 2]), __uint_as_float(r[2 * j + 3])));
    if (2 * j > lane) a0.x = 0.f;                  // causal mask
    if (2 * j + 1 > lane) a0.y = 0.f;
    if (2 * j + 2 > lane) a1.x = 0.f;
    if (2 * j + 3 > lane) a1.y = 0.f;
    if (DFOLD && 2 * j == lane) a0.x += Dh;        // D skip folded into the diagonal
    if (DFOLD && 2 * j + 1 == lane) a0.y += Dh;
    if (DFOLD && 2 * j + 2 == lane) a1.x += Dh;
    if (DFOLD && 2 * j + 3 == lane) a1.y += Dh;
    pk[j] = pack_bf16x2(a0.x, a0.y);
    pk[j + 1] = pack_bf16x2(a1.x, a1.y);
  }
}

// Off-diagonal 32x32 block (every k of the block precedes every row of this warp): the decay factorises without
// overflow around ref = cs at the last token before the warp's rows,
//   exp(cs_m - cs_k) dt_k = u_m * v_k,   u_m = exp(cs_m - ref) <= 1,   v_k = dt_k exp(ref - cs_k) <= dt_k,
// so the block costs two FMULs per element and NO transcendental (u_m: one exp per row, v_k: one per column).
__device__ __forceinline__ void m_offdiag(uint32_t t_src, uint32_t (&pk)[16], const float* __restrict__ sVk, float um) {
  uint32_t r[32];
  tmem_ld32(t_src, r);
  const float2 um2 = make_float2(um, um);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float4 v4 = *reinterpret_cast<const float4*>(sVk + 2 * j);
    const float2 e0 = __fmul2_rn(make_float2(v4.x, v4.y), um2), e1 = __fmul2_rn(make_float2(v4.z, v4.w), um2);
    const float2 a0 = __fmul2_rn(e0, make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])));
    const float2 a1 = __fmul2_rn(e1, make_float2(__uint_as_float(r[2 * j + 2]), __uint_as_float(r[2 * j + 3])));
    pk[j] = pack_bf16x2(a0.x, a0.y);
    pk[j + 1] = pack_bf16x2(a1.x, a1.y);
  }
}

// DFOLD: D is a per-head scalar and is added to the diagonal of M (bf16 A operand), so that Yd = (M + D I) x and
// the epilogue never touches x.  For bf16 parameters D*x is exact in the fp32 accumulator; the only extra rounding
// is bf16(M_mm + D), of the order of the bf16 rounding of y itself.  D of shape (H, P) takes the explicit path.
//
// mbarrier phases: a barrier that is used once per chunk is never more than one phase ahead of any of its waiters
// (every producer of phase c+1 depends, directly or not, on each waiter of phase c having passed its wait); the
// per-buffer barriers (..0 / ..1) advance once per two chunks under the same rule.  The orders that make this true:
//   issuer SO:  S(c) before O(c)   -- YOFULL(c) therefore also says that xs(c) was consumed, i.e. WG_X has finished
//                                     reading the raw x tile of chunk c, which the epilogue then overwrites with y(c);
//   issuer GD:  G(c+1) before D(c) -- G(c+1) waits for D(c-1) (YDFULL), whose A operand M(c-1) lives in the buffer it
//                                     overwrites.
template <bool HAS_Z, bool DFOLD>
__global__ void __launch_bounds__(tc::THREADS, 1)
ssd_fused_kernel(const __grid_constant__ tc::Maps maps, const tc::Args a) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBAR * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int hpg = a.H / a.G;
  const int g = h / hpg;
  const int n = a.nchunks;
  constexpr float LOG2E = 1.4426950408889634f;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[FULLB0 + i], 1); mbar_init(&bars[FULLC0 + i], 1); mbar_init(&bars[FULLX0 + i], 1);
      mbar_init(&bars[EMPTYB0 + i], 2); mbar_init(&bars[EMPTYC0 + i], 2);   // one commit from each issuer
      // one lane per WG_X warp + the D commit (+ WG_C for the explicit D*x path)
      mbar_init(&bars[EMPTYX0 + i], DFOLD ? 5 : 9);
      mbar_init(&bars[CBFULL0 + i], 1);
      mbar_init(&bars[MFULL0 + i], 6);               // 4 WG_A warps + 2 helper warps
      mbar_init(&bars[HREAD0 + i], 2);
    }
    mbar_init(&bars[XSFULL], 4); mbar_init(&bars[STDONE], 1); mbar_init(&bars[DSFREE], 4); mbar_init(&bars[SFULL], 4);
    mbar_init(&bars[YOFULL], 1); mbar_init(&bars[YDFULL], 1); mbar_init(&bars[YDFREE], 4); mbar_init(&bars[YOFREE], 4);
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc<512>(tmem_slot);
  if (threadIdx.x < P) {
    float* sD = reinterpret_cast<float*>(smem + OFF_D);
    sD[threadIdx.x] = a.D == nullptr ? 0.f : (a.d_has_hdim ? a.D[h * P + threadIdx.x] : a.D[h]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t row0 = ((int64_t)b * n) * a.H + h;     // (b, c, h) row of dt_act / cs is row0 + c*H

  if (warp >= W_PROD) {
    reg_set<REG_MISC>();
    if (warp == W_PROD) {
      // =========================== TMA producer: three rings, polled ===========================
      if (elect_one()) {
        prefetch_tmap(&maps.x); prefetch_tmap(&maps.b); prefetch_tmap(&maps.c); prefetch_tmap(&maps.y);
        int cb = 0, cc = 0, cx = 0;
        while (cb < n || cc < n || cx < n) {
          const int before = cb + cc + cx;
          if (cx < n) {
            const int s = cx & 1, u = cx >> 1;
            if (cx < 2 || mbar_test_wait(&bars[EMPTYX0 + s], (u - 1) & 1)) {
              TV_TRACE(0, cx);
#ifdef TV_ENABLE_TRACE
              if (a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
                unsigned long long gt;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                a.trace[(int64_t)cx * 16 + 15] = (long long)gt;
              }
#endif
              uint8_t* xs_ = smem + OFF_X + s * XSTAGE;
              const int t0 = cx * Q;
              mbar_arrive_expect_tx(&bars[FULLX0 + s], TV_ABLATE(8) ? 1024 : (TV_ABLATE(14) ? 2 * 4096 + 1024 : XSTAGE));
              if (!TV_ABLATE(8))
#pragma unroll
              for (int i = 0; i < (TV_ABLATE(14) ? 2 : 5); ++i) tma_load_4d(xs_ + i * 4096, &maps.x, &bars[FULLX0 + s], 16 * i, h, t0, b);
              bulk_load(xs_ + TILE_X, a.cs + (row0 + (int64_t)cx * a.H) * Q, 512, &bars[FULLX0 + s]);
              bulk_load(xs_ + TILE_X + 512, a.dt_act + (row0 + (int64_t)cx * a.H) * Q, 512, &bars[FULLX0 + s]);
              ++cx;
            }
          }
          if (cb < n) {
            const int s = cb & 1, u = cb >> 1;
            if (cb < 2 || mbar_test_wait(&bars[EMPTYB0 + s], (u - 1) & 1)) {
              const int t0 = cb * Q;
              if (TV_ABLATE(9)) mbar_arrive(&bars[FULLB0 + s]); else {
              mbar_arrive_expect_tx(&bars[FULLB0 + s], TILE_BC);
              tma_load_4d(smem + OFF_B + s * TILE_BC, &maps.b, &bars[FULLB0 + s], 0, g, t0, b);
              tma_load_4d(smem + OFF_B + s * TILE_BC + 16384, &maps.b, &bars[FULLB0 + s], 64, g, t0, b);
              }
              ++cb;
            }
          }
          if (cc < n) {
            const int s = cc & 1, u = cc >> 1;
            if (cc < 2 || mbar_test_wait(&bars[EMPTYC0 + s], (u - 1) & 1)) {
              const int t0 = cc * Q;
              if (TV_ABLATE(9)) mbar_arrive(&bars[FULLC0 + s]); else {
              mbar_arrive_expect_tx(&bars[FULLC0 + s], TILE_BC);
              tma_load_4d(smem + OFF_C + s * TILE_BC, &maps.c, &bars[FULLC0 + s], 0, g, t0, b);
              tma_load_4d(smem + OFF_C + s * TILE_BC + 16384, &maps.c, &bars[FULLC0 + s], 64, g, t0, b);
              }
              ++cc;
            }
          }
          if (TV_POLL_NS_PROD > 0 && cb + cc + cx == before) __nanosleep(TV_POLL_NS_PROD);
        }
      }
    } else if (warp == W_MMA) {
      // =========================== MMA issuer: four streams (S, O, D, G), polled ===========================
      // One thread issues every MMA, a whole group of 8 at a time (groups of different shapes issued concurrently from
      // two threads interleave on the tensor pipe), but it never WAITS for a group in program order: each pass over
      // the four streams issues whichever next group has all its inputs, state path (S, O) first.
      if (elect_one()) {
        constexpr uint32_t ID_CB = umma_idesc_bf16(128, 128, false, false);
        constexpr uint32_t ID_Y = umma_idesc_bf16(128, P, false, true);
        constexpr uint32_t ID_ST = umma_idesc_bf16(128, P, true, true);
        const uint32_t sbase = smem_u32(smem);
        // descriptor templates; per-MMA offsets are added to the 14-bit start-address field (units of 16 bytes)
        const uint64_t dB_k = umma_smem_desc(sbase + OFF_B, 16, 1024, SWZ_128B);        // B as K-major operand (G)
        const uint64_t dB_mn = umma_smem_desc(sbase + OFF_B, 16384, 1024, SWZ_128B);    // B as MN-major A operand (S)
        const uint64_t dC_k = umma_smem_desc(sbase + OFF_C, 16, 1024, SWZ_128B);
        const uint64_t dX = umma_smem_desc(sbase + OFF_X, 4096, 256, SWZ_32B);
        const uint64_t dXS = umma_smem_desc(sbase + OFF_XS, 4096, 256, SWZ_32B);
        const uint64_t dS = umma_smem_desc(sbase + OFF_S, 4096, 256, SWZ_32B);
        // The issue loops are fully unrolled with compile-time descriptor offsets: a rolled loop costs ~110 cycles per
        // MMA on the single issuing thread (descriptor arithmetic + R2UR), which made issue the bottleneck.
        int cS = 0, cO = 0, cD = 0, cG = 0;          // next chunk of each stream
        int dDone = 0;                               // D groups known to be complete (YDFULL phases observed)
#ifdef TV_ENABLE_TRACE
        long long passes = 0;
#endif
#pragma unroll 1
        while (cO < n || cD < n) {
          const int before = cS + cO + cD + cG;
#ifdef TV_ENABLE_TRACE
          ++passes;
#endif
          // ---- S(c): dS = B^T . xs  (fresh accumulator; WG_S has drained dS(c-1))
          if (cS < n) {
            const int c = cS, s = c & 1, u = c >> 1;
            if (mbar_test_wait(&bars[XSFULL], c & 1) && mbar_test_wait(&bars[FULLB0 + s], u & 1) &&
                (c == 0 || mbar_test_wait(&bars[DSFREE], (c - 1) & 1))) {
              tc_fence_after();
              TV_TRACE(2, c);
#ifdef TV_ENABLE_TRACE
              if (a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) a.trace[(int64_t)c * 16 + 10] = passes;
#endif
              const uint64_t db = umma_desc_advance(dB_mn, (uint32_t)(s * (TILE_BC >> 4)));
              if (!TV_ABLATE(7))
#pragma unroll
              for (int j = 0; j < 8; ++j)
                umma_ss(tmem + T_DS, umma_desc_advance(db, j * 128), umma_desc_advance(dXS, j * 32), ID_ST, j > 0);
              umma_commit(&bars[STDONE]);
              umma_commit(&bars[EMPTYB0 + s]);
              ++cS;
            }
          }
          // ---- O(c): Yo = C . S_c   (only after S(c): see the note on YOFULL above the kernel)
          if (cO < cS) {
            const int c = cO, s = c & 1, u = c >> 1;
            if (mbar_test_wait(&bars[SFULL], c & 1) && mbar_test_wait(&bars[FULLC0 + s], u & 1) &&
                (c == 0 || mbar_test_wait(&bars[YOFREE], (c - 1) & 1))) {
              tc_fence_after();
              TV_TRACE(3, c);
              const uint64_t dc = umma_desc_advance(dC_k, (uint32_t)(s * (TILE_BC >> 4)));
              if (!TV_ABLATE(5))
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t o = (uint32_t)(((j >> 2) * 16384 + (j & 3) * 32) >> 4);
                umma_ss(tmem + T_YO, umma_desc_advance(dc, o), umma_desc_advance(dS, j * 32), ID_Y, j > 0);
              }
              umma_commit(&bars[YOFULL]);
              umma_commit(&bars[EMPTYC0 + s]);
              ++cO;
            }
          }
          // ---- D(c): Yd = M . x   (A = M(c), packed bf16, columns 0..63 of its C.B^T buffer)
          if (cD < cG) {
            const int c = cD, s = c & 1, u = c >> 1;
            if (mbar_test_wait(&bars[MFULL0 + s], u & 1) && mbar_test_wait(&bars[FULLX0 + s], u & 1) &&
                (c == 0 || mbar_test_wait(&bars[YDFREE], (c - 1) & 1))) {
              while (dDone < c) { mbar_wait(&bars[YDFULL], dDone & 1); ++dDone; }   // complete already (YDFREE(c-1))
              tc_fence_after();
              TV_TRACE(4, c);
              const uint64_t dx = umma_desc_advance(dX, (uint32_t)(s * (XSTAGE >> 4)));
              const uint32_t tmA = tmem + (s ? T_CB1 : T_CB0);
              if (!TV_ABLATE(6))
#pragma unroll
              for (int j = 0; j < 8; ++j) umma_ts(tmem + T_YD, tmA + j * 8, umma_desc_advance(dx, j * 32), ID_Y, j > 0);
              umma_commit(&bars[YDFULL]);
              umma_commit(&bars[EMPTYX0 + s]);
              ++cD;
            }
          }
          // ---- G(c): C.B^T, up to two chunks ahead of D; overwrites the buffer that held M(c-2)
          if (cG < n && cG <= cD + 1) {
            const int c = cG, s = c & 1, u = c >> 1;
            if (dDone < c - 1 && mbar_test_wait(&bars[YDFULL], dDone & 1)) ++dDone;
            if (dDone >= c - 1 && mbar_test_wait(&bars[FULLB0 + s], u & 1) && mbar_test_wait(&bars[FULLC0 + s], u & 1)) {
              tc_fence_after();
              TV_TRACE(1, c);
              const uint32_t so = (uint32_t)(s * (TILE_BC >> 4));
              const uint64_t dc = umma_desc_advance(dC_k, so), db = umma_desc_advance(dB_k, so);
              const uint32_t tcb = tmem + (s ? T_CB1 : T_CB0);
              if (!TV_ABLATE(4))
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t o = (uint32_t)(((j >> 2) * 16384 + (j & 3) * 32) >> 4);
                umma_ss(tcb, umma_desc_advance(dc, o), umma_desc_advance(db, o), ID_CB, j > 0);
              }
              umma_commit(&bars[CBFULL0 + s]);
              umma_commit(&bars[EMPTYB0 + s]);
              umma_commit(&bars[EMPTYC0 + s]);
              ++cG;
            }
          }
          if (TV_POLL_NS_MMA > 0 && cS + cO + cD + cG == before) __nanosleep(TV_POLL_NS_MMA);
        }
      }
    } else {
      // =========================== helper warps 18 / 19: off-diagonal blocks 0 and 1 of row quarters 2 / 3 ===========================
      // WG_A's warp of the same quarter writes its packed blocks over C.B^T block 1 and may do so once HREAD says that
      // this warp's load of block 1 has landed.
      const int q = warp & 3;
      const int m = q * 32 + lane;
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
      float* scr = reinterpret_cast<float*>(smem + OFF_SCR + (4 + q - 2) * SCR);      // V[64]
      // cs / dt of the next chunk are fetched from global memory (L2) one chunk ahead, NOT from the x stage: M(c) must be
      // ready when x(c) lands, so that D(c) can run at once and hand the x stage back to the TMA producer.
      const float* gcs = a.cs + row0 * Q;
      const float* gdt = a.dt_act + row0 * Q;
      const int64_t cstride = (int64_t)a.H * Q;
      float n_csm = gcs[m], n_ref = gcs[32 * q - 1], n_cs0 = gcs[lane], n_cs1 = gcs[32 + lane], n_dt0 = gdt[lane], n_dt1 = gdt[32 + lane];
      for (int c = 0; c < n; ++c) {
        const int s = c & 1, u = c >> 1;
        const float ref = n_ref;
        const float um = ex2_approx((n_csm - ref) * LOG2E);                // u_m = exp(cs_m - ref) <= 1
        __syncwarp();
        scr[lane] = n_dt0 * ex2_approx((ref - n_cs0) * LOG2E);             // v_k = dt_k exp(ref - cs_k) <= dt_k
        scr[32 + lane] = n_dt1 * ex2_approx((ref - n_cs1) * LOG2E);
        __syncwarp();
        if (c + 1 < n) {
          const float* ncs = gcs + (c + 1) * cstride;
          const float* ndt = gdt + (c + 1) * cstride;
          n_csm = ncs[m]; n_ref = ncs[32 * q - 1]; n_cs0 = ncs[lane]; n_cs1 = ncs[32 + lane]; n_dt0 = ndt[lane]; n_dt1 = ndt[32 + lane];
        }
        mbar_wait(&bars[CBFULL0 + s], u & 1);
        tc_fence_after();
        const uint32_t tcb = tmem + (s ? T_CB1 : T_CB0) + lane_base;
        uint32_t pk[16];
        if (!TV_ABLATE(1)) {
          m_offdiag(tcb, pk, scr, um);
          tmem_st16(tcb, pk);                        // over C.B^T block 0, which only this warp reads
          m_offdiag(tcb + 32, pk, scr + 32, um);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[HREAD0 + s]);
        if (!TV_ABLATE(1)) {
          tmem_st16(tcb + 16, pk);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[MFULL0 + s]);
      }
    }
  } else if (warp < W_X) {
    // =========================== WG_A: M = C.B^T (.) decay, in place ===========================
    // Row quarter q (TMEM lanes 32q..32q+31, reachable only from warps with id % 4 == q) owns q+1 blocks of 32 columns of
    // C.B^T; packed M block kb goes to columns 16kb..16kb+15 of the same buffer, i.e. over C.B^T block kb/2.  Ten blocks
    // over six warps, at most two each: this warp takes the diagonal block plus block 0 (q = 1) or block 2 (q = 3); the
    // helper warps 18 / 19 take blocks 0 and 1 of quarters 2 / 3, and this warp stores over C.B^T block 1 only after
    // HREAD (their loads have landed).  The blocks above the diagonal are zeroed every chunk (the buffer is overwritten
    // by the next C.B^T).  Nothing on the state path (x scaling, state fold) waits for these warps.
    reg_set<REG_A>();
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float Dh = (DFOLD && a.D != nullptr) ? a.D[h] : 0.f;
    float* scr = reinterpret_cast<float*>(smem + OFF_SCR + warp * SCR);      // V[32] (off-diagonal block) | F[32] (diagonal)
    const int kbo = q > 0 ? q - 1 : 0;           // the off-diagonal block of quarters 1 and 3 (its last cs is `ref`)
    uint32_t zz[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) zz[j] = 0u;
    // cs / dt of the next chunk are fetched from global memory (L2) one chunk ahead, NOT from the x stage: M(c) must be
    // ready when x(c) lands, so that D(c) can run at once and hand the x stage back to the TMA producer.
    const float* gcs = a.cs + row0 * Q;
    const float* gdt = a.dt_act + row0 * Q;
    const int64_t cstride = (int64_t)a.H * Q;
    float n_csm = gcs[m], n_dtm = gdt[m], n_cso = gcs[kbo * 32 + lane], n_dto = gdt[kbo * 32 + lane];
    for (int c = 0; c < n; ++c) {
      const int s = c & 1, u = c >> 1;
      const float cs_m = n_csm, dt_m = n_dtm, cs_o = n_cso, dt_o = n_dto;
      if (c + 1 < n) {
        const float* ncs = gcs + (c + 1) * cstride;
        const float* ndt = gdt + (c + 1) * cstride;
        n_csm = ncs[m]; n_dtm = ndt[m]; n_cso = ncs[kbo * 32 + lane]; n_dto = ndt[kbo * 32 + lane];
      }
      const float Em = cs_m * LOG2E;
      const float ref = q > 0 ? __shfl_sync(0xffffffffu, cs_o, 31) : 0.f;   // cs just before this warp's rows
      const bool fast = ref - __shfl_sync(0xffffffffu, cs_m, 31) < 60.f;    // decay inside the diagonal block (warp-uniform)
      const float um = ex2_approx((cs_m - ref) * LOG2E);                    // u_m = exp(cs_m - ref) <= 1
      __syncwarp();                              // every lane is done with the previous chunk's factors
      // diagonal block: v_k = dt_k exp(ref - cs_k) (fast path) or F_k = log2(dt_k) - cs_k*log2e (dt = 0 -> -inf -> weight 0)
      scr[32 + lane] = fast ? dt_m * ex2_approx((ref - cs_m) * LOG2E) : __log2f(dt_m) - Em;
      if (q & 1) scr[lane] = dt_o * ex2_approx((ref - cs_o) * LOG2E);
      __syncwarp();
      mbar_wait(&bars[CBFULL0 + s], u & 1);
      tc_fence_after();
      if (threadIdx.x == 96) TV_TRACE(5, c);
      const uint32_t tcb = tmem + (s ? T_CB1 : T_CB0) + lane_base;
      uint32_t pko[16], pkd[16];
      if (!TV_ABLATE(1)) {
        if (q & 1) m_offdiag(tcb + kbo * 32, pko, scr, um);
        if (fast) m_diag_fast<DFOLD>(tcb + q * 32, pkd, scr + 32, um, lane, Dh);
        else m_diag<DFOLD>(tcb + q * 32, pkd, scr + 32, Em, lane, Dh);
      }
      if (q >= 2) { mbar_wait(&bars[HREAD0 + s], u & 1); tc_fence_after(); }
      if (!TV_ABLATE(1)) {
        if (q & 1) tmem_st16(tcb + kbo * 16, pko);
        tmem_st16(tcb + q * 16, pkd);
#pragma unroll
        for (int j = 1; j < 4; ++j)
          if (j > q) tmem_st16(tcb + j * 16, zz);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[MFULL0 + s]);
      if (threadIdx.x == 96) TV_TRACE(6, c);
    }
  } else if (warp < W_S) {
    // =========================== WG_X: xs = w.x for the S MMA ===========================
    reg_set<REG_X>();
    const int r = threadIdx.x - W_X * 32;        // token row of x
    for (int c = 0; c < n; ++c) {
      const int s = c & 1, u = c >> 1;
      const uint8_t* xst = smem + OFF_X + s * XSTAGE;
      const float* sCS = reinterpret_cast<const float*>(xst + TILE_X);
      const float* sDT = sCS + 128;
      mbar_wait(&bars[FULLX0 + s], u & 1);
      if (r == 0) TV_TRACE(7, c);
      const float cs_last = sCS[Q - 1];
      const float w_r = sDT[r] * __expf(cs_last - sCS[r]);
      const __nv_bfloat162 w2 = __float2bfloat162_rn(w_r);
      // ---- xs = w_r * x into registers now; stored once S(c-1) has consumed the previous xs
      uint32_t xs[40];
      if (!TV_ABLATE(13))
#pragma unroll
      for (int qq = 0; qq < 10; ++qq) {
        uint4 v = *reinterpret_cast<const uint4*>(xst + off_sw32(r, qq));
        __nv_bfloat162* hv = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) hv[i] = __hmul2(hv[i], w2);
        xs[4 * qq] = v.x; xs[4 * qq + 1] = v.y; xs[4 * qq + 2] = v.z; xs[4 * qq + 3] = v.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[EMPTYX0 + s]);     // the raw x tile and cs / dt have been read
      if (c > 0) mbar_wait(&bars[STDONE], (c - 1) & 1);
      if (r == 0) TV_TRACE(8, c);
      if (!TV_ABLATE(13))
#pragma unroll
      for (int qq = 0; qq < 10; ++qq)
        *reinterpret_cast<uint4*>(smem + OFF_XS + off_sw32(r, qq)) = make_uint4(xs[4 * qq], xs[4 * qq + 1], xs[4 * qq + 2], xs[4 * qq + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[XSFULL]);
      if (r == 96) TV_TRACE(9, c);
    }
  } else if (warp < W_C) {
    // =========================== WG_S: the running state, in registers ===========================
    // Thread r holds row n = r of the 128x80 fp32 state.  Per chunk: S_{c+1} = exp(cs_last) S_c + dS(c), then the bf16
    // copy of S_{c+1} that O(c+1) reads goes to shared memory once O(c) has finished with the previous copy.
    reg_set<REG_S>();
    const int r = threadIdx.x - W_S * 32;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float st[P];
#pragma unroll
    for (int j = 0; j < P; ++j) st[j] = a.init == nullptr ? 0.f : a.init[(((int64_t)b * a.H + h) * P + j) * N + r];
    auto write_copy = [&]() {
#pragma unroll
      for (int qq = 0; qq < 10; ++qq)
        *reinterpret_cast<uint4*>(smem + OFF_S + off_sw32(r, qq)) =
            make_uint4(pack_bf16x2(st[8 * qq], st[8 * qq + 1]), pack_bf16x2(st[8 * qq + 2], st[8 * qq + 3]),
                       pack_bf16x2(st[8 * qq + 4], st[8 * qq + 5]), pack_bf16x2(st[8 * qq + 6], st[8 * qq + 7]));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[SFULL]);
    };
    write_copy();
    float logsum = 0.f;
    float cs_next = a.cs[row0 * Q + (Q - 1)];
    for (int c = 0; c < n; ++c) {
      const float cs_last = cs_next;
      if (c + 1 < n) cs_next = a.cs[(row0 + (int64_t)(c + 1) * a.H) * Q + (Q - 1)];
      const float a_c = __expf(cs_last);
      logsum += cs_last;
      mbar_wait(&bars[STDONE], c & 1);
      tc_fence_after();
      if (!TV_ABLATE(2)) {
        uint32_t v[48];
#pragma unroll
        for (int pc = 0; pc < 3; ++pc) tmem_ld16(tmem + T_DS + lane_base + pc * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[pc * 16]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 48; ++j) st[j] = fmaf(a_c, st[j], __uint_as_float(v[j]));
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) tmem_ld16(tmem + T_DS + lane_base + 48 + pc * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[pc * 16]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[DSFREE]);
#pragma unroll
        for (int j = 0; j < 32; ++j) st[48 + j] = fmaf(a_c, st[48 + j], __uint_as_float(v[j]));
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[DSFREE]);
      }
      if (c + 1 < n) {
        mbar_wait(&bars[YOFULL], c & 1);           // O(c) has read the copy of S_c
        write_copy();
      }
      if (r == 0) TV_TRACE(14, c);
    }
    if (a.fin != nullptr) {
#pragma unroll
      for (int j = 0; j < P; ++j) a.fin[(((int64_t)b * a.H + h) * P + j) * N + r] = st[j];
    }
    if (a.logdecay != nullptr && r == 0) a.logdecay[(int64_t)b * a.H + h] = logsum;
  } else {
    // =========================== WG_C: epilogue  y = Yd + exp(cs_m) Yo + D x  [* silu(z)] ===========================
    // Each warp owns 32 token rows end to end: drain, combine, stage its bf16 piece in its own 3 KB staging area and write
    // it with its own TMA stores, in two passes (columns 0..47, then 48..79 over the first two atoms once TMA has read them).
    reg_set<REG_C>();
    const int r = threadIdx.x - W_C * 32;          // token row m
    const int wq = warp & 3;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const float* sD = reinterpret_cast<const float*>(smem + OFF_D);
    uint8_t* ystg = smem + OFF_Y + wq * YSTG;
    for (int c = 0; c < n; ++c) {
      const int s = c & 1, u = c >> 1;
      const uint8_t* xst = smem + OFF_X + s * XSTAGE;
      const int t = c * Q + r;
      const float e_r = __expf(a.cs[(row0 + (int64_t)c * a.H) * Q + r]);   // one coalesced 512-byte row per chunk (L2 hit)
      // z gate: the row's 160 bytes are fetched now, ahead of the waits, so that no round of the drain sees a global load
      uint4 zr[HAS_Z ? 10 : 1];
      if (HAS_Z) {
        const __nv_bfloat16* zrow = a.z + b * a.zbs + (int64_t)t * a.zss + (int64_t)h * a.zhs;
#pragma unroll
        for (int i = 0; i < 10; ++i)
          zr[HAS_Z ? i : 0] = t < a.L ? *reinterpret_cast<const uint4*>(zrow + i * 8) : make_uint4(0, 0, 0, 0);
      }
      const bool rows_live = c * Q + wq * 32 < a.L;
      if (!DFOLD) mbar_wait(&bars[FULLX0 + s], u & 1);
      mbar_wait(&bars[YDFULL], c & 1);
      mbar_wait(&bars[YOFULL], c & 1);
      tc_fence_after();
      if (r == 0) TV_TRACE(11, c);
      // Software-pipelined drain: the TMEM loads of round pc+1 are in flight while round pc is combined and stored;
      // the accumulators are handed back as soon as the last load has landed, before the last round's math.
      uint32_t ydA[16], yoA[16], ydB[16], yoB[16];
      tmem_ld16(tmem + T_YD + lane_base, ydA);
      tmem_ld16(tmem + T_YO + lane_base, yoA);
      if (lane == 0) tma_store_wait_read();        // the staging atoms of the previous chunk's second pass have been read
      __syncwarp();
      tmem_ld_wait();
      auto finish_round = [&](int pc, const uint32_t (&yd)[16], const uint32_t (&yo)[16]) {
        float xv[16];
        if (!DFOLD) {                              // explicit D*x path ((H,P)-shaped D): x row from the x stage
          const uint4 xa = *reinterpret_cast<const uint4*>(xst + off_sw32(r, 2 * pc));
          const uint4 xb = *reinterpret_cast<const uint4*>(xst + off_sw32(r, 2 * pc + 1));
          const uint32_t xw[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            xv[2 * j] = sD[pc * 16 + 2 * j] * __uint_as_float(xw[j] << 16);
            xv[2 * j + 1] = sD[pc * 16 + 2 * j + 1] * __uint_as_float(xw[j] & 0xffff0000u);
          }
        }
        float zv[16];
        if (HAS_Z) {
          const uint4 za = zr[HAS_Z ? 2 * pc : 0], zb = zr[HAS_Z ? 2 * pc + 1 : 0];
          const uint32_t zw[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            zv[2 * j] = silu<true>(__uint_as_float(zw[j] << 16));
            zv[2 * j + 1] = silu<true>(__uint_as_float(zw[j] & 0xffff0000u));
          }
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 y2 = __ffma2_rn(make_float2(e_r, e_r), make_float2(__uint_as_float(yo[2 * j]), __uint_as_float(yo[2 * j + 1])),
                                       make_float2(__uint_as_float(yd[2 * j]), __uint_as_float(yd[2 * j + 1])));
          float y0 = y2.x, y1 = y2.y;
          if (!DFOLD) { y0 += xv[2 * j]; y1 += xv[2 * j + 1]; }
          if (HAS_Z) { y0 *= zv[2 * j]; y1 *= zv[2 * j + 1]; }
          pk[j] = pack_bf16x2(y0, y1);
        }
        // stage the 32-byte row piece: atom pc (first pass) or pc-3 (second pass), SW32 layout; TMA clips the ragged tail
        if (!TV_ABLATE(11)) {
          uint8_t* at = ystg + (pc < 3 ? pc : pc - 3) * 1024 + lane * 32;
          const uint32_t sw = (uint32_t)((lane >> 2) & 1) << 4;
          *reinterpret_cast<uint4*>(at + sw) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(at + (sw ^ 16u)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      };
      auto store_pass = [&](int first, int count) {   // atoms first..first+count-1 of y -> global
        fence_proxy_async();                       // the staged piece becomes visible to the async proxy (TMA)
        __syncwarp();
        if (lane == 0) {
          if (!TV_ABLATE(11) && rows_live) {
            for (int i = 0; i < (TV_ABLATE(15) ? 1 : count); ++i)
              tma_store_4d(&maps.y, ystg + i * 1024, 16 * (first + i), h, c * Q + wq * 32, b);
          }
          tma_store_commit();
        }
      };
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&bars[YDFREE]); mbar_arrive(&bars[YOFREE]); }
        if (r == 0) TV_TRACE(12, c);
      };
      if (TV_ABLATE(0)) {
        release_acc();
      } else {
        // rounds: 0 (A) 1 (B) 2 (A) | 3 (B) 4 (A)
        tmem_ld16(tmem + T_YD + lane_base + 16, ydB);
        tmem_ld16(tmem + T_YO + lane_base + 16, yoB);
        finish_round(0, ydA, yoA);
        tmem_ld_wait();
        tmem_ld16(tmem + T_YD + lane_base + 32, ydA);
        tmem_ld16(tmem + T_YO + lane_base + 32, yoA);
        finish_round(1, ydB, yoB);
        tmem_ld_wait();
        tmem_ld16(tmem + T_YD + lane_base + 48, ydB);
        tmem_ld16(tmem + T_YO + lane_base + 48, yoB);
        finish_round(2, ydA, yoA);
        tmem_ld16(tmem + T_YD + lane_base + 64, ydA);
        tmem_ld16(tmem + T_YO + lane_base + 64, yoA);
        store_pass(0, 3);
        tmem_ld_wait();
        release_acc();                             // the last loads have landed: hand Yd / Yo back
        if (lane == 0) tma_store_wait_read();      // first pass read: atoms 0 and 1 are free again
        __syncwarp();
        finish_round(3, ydB, yoB);
        finish_round(4, ydA, yoA);
        store_pass(3, 2);
      }
      if (!DFOLD) { __syncwarp(); if (lane == 0) mbar_arrive(&bars[EMPTYX0 + s]); }   // x rows read
      if (r == 0) TV_TRACE(13, c);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<512>(tmem);
}


// =================================================================================================================
// Shard summary (pass 1 of the sequence-sharded prefill): final state from a zero initial state + sum of dt*A.
//
// No recurrence is needed for this: with T_c = sum of the chunk totals of all LATER chunks (<= 0),
//   S_final[n,p] = sum_c sum_k B[k,n] * x[k,p] * dt_k * exp(cs_last(c) - cs_k + T_c)
// is ONE long-K GEMM accumulated in TMEM (every weight is <= dt_k, so nothing overflows), the chunks can be taken
// in any order, and a chunk whose T_c has underflowed fp32 contributes exactly zero.  The kernel therefore walks
// the chunks BACKWARDS from the end of the shard and stops at first_chunk[b,h], the first chunk whose suffix decay is
// still representable (ssd_suffix_scan_kernel).  For fast-decaying heads that is 1-3 chunks, whatever the shard length.
// =================================================================================================================
constexpr float kLogUnderflow = -104.f;   // exp(x) == 0 in fp32 (incl. denormals) for x < -103.98

// one warp per (b, h): logdecay_sum = sum_c cs_last(c);  first_chunk = smallest c with T_c >= kLogUnderflow
__global__ void __launch_bounds__(128)
ssd_suffix_scan_kernel(const float* __restrict__ cs, float* __restrict__ logdecay, int* __restrict__ first_chunk,
                       int BH, int H, int nchunks, int Q) {
  const int bh = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (bh >= BH) return;
  const int b = bh / H, h = bh - b * H;
  const float* base = cs + (((int64_t)b * nchunks) * H + h) * Q + (Q - 1);     // chunk c at + c*H*Q
  float carry = 0.f;            // sum of the totals of chunks after the current block of 32
  int first = 0;
  bool found = false;
  for (int hi = nchunks; hi > 0; hi -= 32) {          // blocks of 32 chunks, from the end
    const int c = hi - 1 - lane;                      // lane 0 = last chunk of the block
    const float v = c >= 0 ? base[(int64_t)c * H * Q] : 0.f;
    float incl = v;                                   // inclusive scan over lanes = suffix sum over chunks
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const float Tc = carry + incl - v;                // sum over chunks strictly after c
    const unsigned dead = __ballot_sync(0xffffffffu, c >= 0 && Tc < kLogUnderflow);
    if (!found && dead != 0u) {                       // lowest dead lane = latest dead chunk: everything before is dead too
      first = hi - 1 - (__ffs(dead) - 1) + 1;
      found = true;
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) {
    if (logdecay != nullptr) logdecay[bh] = carry;
    first_chunk[bh] = found ? first : 0;
  }
}

namespace st {
constexpr int THREADS = 192;                     // warps 0-3: x scaling, warp 4: TMA producer, warp 5: MMA issuer
// Two stages of (B tile | x tile): 109 KB, so that TWO CTAs fit on an SM -- the walk over the chunks of one head is
// a latency chain (TMA -> scale -> MMA), and a second resident CTA (the other half of the same head's chunks, see
// `split`) fills its bubbles.  w*x is written back IN PLACE over the x tile (nothing else reads raw x here).
constexpr int NST = 2;
constexpr uint32_t OFF_B = 0, OFF_X = NST * tc::TILE_BC;
constexpr uint32_t OFF_BAR = OFF_X + NST * tc::XSTAGE;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 128 + 1024;
enum Bar { FULL0 = 0, EMPTY0 = FULL0 + NST, SCALED0 = EMPTY0 + NST, DONE = SCALED0 + NST, NBAR };
static_assert(OFF_X % 1024 == 0 && 2 * SMEM_BYTES <= 226 * 1024, "state kernel smem: two CTAs per SM");
}  // namespace st

constexpr int kMaxStateSplit = 8;
// grid (H, batch, split).  Part z of `split` handles the visited-chunk indices [total*z/split, total*(z+1)/split) of the
// backwards walk n-1, n-2, ..., c_first; because the decay is taken relative to the END of the shard the partial states
// simply add up.  With split > 1 every part with work writes its partial state to `partial` [(b,h)][part][P][N] and
// ssd_state_reduce_kernel adds the parts in a fixed order (no atomics: the summary is bit-reproducible); a long walk
// (a slowly decaying head) is then `split` short chains on different SMs instead of one chain of nchunks steps.
__global__ void __launch_bounds__(st::THREADS, 2)
ssd_state_kernel(const __grid_constant__ tc::Maps maps, const tc::Args a, const int* __restrict__ first_chunk,
                 const int split, float* __restrict__ partial) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + st::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + st::OFF_BAR + st::NBAR * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y, part = blockIdx.z;
  const int g = h / (a.H / a.G);
  const int n = a.nchunks;
  const int c_first = first_chunk[(int64_t)b * a.H + h];
  const int total = n - c_first;                       // chunks n-1, n-2, ..., c_first
  const int i_begin = (int)((int64_t)total * part / split), i_end = (int)((int64_t)total * (part + 1) / split);
  const int count = i_end - i_begin;
  if (count == 0 && split > 1) return;                 // no partial state: the reduction skips this part
  if (threadIdx.x == 0) {
    for (int i = 0; i < st::NST; ++i) {
      mbar_init(&bars[st::FULL0 + i], 1); mbar_init(&bars[st::EMPTY0 + i], 1); mbar_init(&bars[st::SCALED0 + i], 4);
    }
    mbar_init(&bars[st::DONE], 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t row0 = ((int64_t)b * n) * a.H + h;

  if (warp == 4) {
    if (elect_one()) {
      prefetch_tmap(&maps.x); prefetch_tmap(&maps.b);
      for (int j = 0; j < count; ++j) {
        const int c = n - 1 - (i_begin + j), s = j % st::NST, u = j / st::NST;
        if (j >= st::NST) mbar_wait(&bars[st::EMPTY0 + s], (u - 1) & 1);
        uint8_t* sb = smem + st::OFF_B + s * TILE_BC;
        uint8_t* sx = smem + st::OFF_X + s * XSTAGE;
        mbar_arrive_expect_tx(&bars[st::FULL0 + s], TILE_BC + XSTAGE);
        const int t0 = c * Q;
        tma_load_4d(sb, &maps.b, &bars[st::FULL0 + s], 0, g, t0, b);
        tma_load_4d(sb + 16384, &maps.b, &bars[st::FULL0 + s], 64, g, t0, b);
#pragma unroll
        for (int k = 0; k < 5; ++k) tma_load_4d(sx + k * 4096, &maps.x, &bars[st::FULL0 + s], 16 * k, h, t0, b);
        bulk_load(sx + TILE_X, a.cs + (row0 + (int64_t)c * a.H) * Q, 512, &bars[st::FULL0 + s]);
        bulk_load(sx + TILE_X + 512, a.dt_act + (row0 + (int64_t)c * a.H) * Q, 512, &bars[st::FULL0 + s]);
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t ID_ST = umma_idesc_bf16(128, P, true, true);
      const uint32_t sbase = smem_u32(smem);
      const uint64_t dB = umma_smem_desc(sbase + st::OFF_B, 16384, 1024, SWZ_128B);
      const uint64_t dXS = umma_smem_desc(sbase + st::OFF_X, 4096, 256, SWZ_32B);
#pragma unroll 1
      for (int j = 0; j < count; ++j) {
        const int s = j % st::NST, u = j / st::NST;
        mbar_wait(&bars[st::SCALED0 + s], u & 1);       // implies FULL: the scaling warps waited for the tiles
        tc_fence_after();
        const uint64_t db = umma_desc_advance(dB, (uint32_t)(s * (TILE_BC >> 4)));
        const uint64_t dx = umma_desc_advance(dXS, (uint32_t)(s * (XSTAGE >> 4)));
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem, umma_desc_advance(db, k * 128), umma_desc_advance(dx, k * 32), ID_ST, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&bars[st::EMPTY0 + s]);
      }
      umma_commit(&bars[st::DONE]);
    }
  } else {
    // ---- x scaling in place: x[k,p] *= dt_k * exp(cs_last - cs_k + T_c), T_c = totals of the chunks already visited
    const int r = threadIdx.x;
    float T = 0.f;
    if (i_begin > 0) {                                   // totals of the chunks the earlier part(s) visit
      for (int i = lane; i < i_begin; i += 32) T += a.cs[(row0 + (int64_t)(n - 1 - i) * a.H) * Q + (Q - 1)];
      T = warp_sum(T);
    }
    for (int j = 0; j < count; ++j) {
      const int s = j % st::NST, u = j / st::NST;
      uint8_t* xst = smem + st::OFF_X + s * XSTAGE;
      const float* sCS = reinterpret_cast<const float*>(xst + TILE_X);
      const float* sDT = sCS + 128;
      mbar_wait(&bars[st::FULL0 + s], u & 1);
      const float cs_last = sCS[Q - 1];
      const float w_r = sDT[r] * __expf(cs_last - sCS[r] + T);
      T += cs_last;
      const __nv_bfloat162 w2 = __float2bfloat162_rn(w_r);
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        uint4 v = *reinterpret_cast<const uint4*>(xst + off_sw32(r, q));
        __nv_bfloat162* hv = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) hv[k] = __hmul2(hv[k], w2);
        *reinterpret_cast<uint4*>(xst + off_sw32(r, q)) = v;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[st::SCALED0 + s]);
    }
    mbar_wait(&bars[st::DONE], 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
    for (int pc = 0; pc < 5; ++pc) {
      uint32_t v[16];
      if (count > 0) {
        tmem_ld16(tmem + lane_base + pc * 16, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = 0u;       // split == 1 and no live chunk: the summary is exactly zero
      }
      float* dst = split > 1 ? partial + ((((int64_t)b * a.H + h) * split + part) * P + pc * 16) * N + r
                             : a.fin + (((int64_t)b * a.H + h) * P + pc * 16) * N + r;
#pragma unroll
      for (int k = 0; k < 16; ++k) dst[(int64_t)k * N] = __uint_as_float(v[k]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<128>(tmem);
}

// out[b,h] = sum over the parts that had work, in part order (the same partition arithmetic as ssd_state_kernel)
__global__ void __launch_bounds__(256)
ssd_state_reduce_kernel(const float* __restrict__ partial, const int* __restrict__ first_chunk, float* __restrict__ out,
                        int nchunks, int split, int PN4) {
  const int64_t bh = blockIdx.y;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= PN4) return;
  const int total = nchunks - first_chunk[bh];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int part = 0; part < split; ++part) {
    const int i_begin = (int)((int64_t)total * part / split), i_end = (int)((int64_t)total * (part + 1) / split);
    if (i_end > i_begin) {
      const float4 v = reinterpret_cast<const float4*>(partial)[(bh * split + part) * PN4 + e];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  reinterpret_cast<float4*>(out)[bh * PN4 + e] = acc;
}

// ------------------------------------------------------------------------------------------------ host
static int g_ablate = 0;
void set_ablate(int m) { g_ablate = m; }
static void* g_trace_ptr = nullptr;   // debug: device buffer of nchunks*16 int64 (tv_debug_set_trace)
void set_trace_buffer(void* p) { g_trace_ptr = p; }
// what the dt / cumsum arrays in a workspace were computed from (see reuse_dt_cumsum below)
struct DtTag {
  const void *dt, *A, *bias;
  int batch, seqlen, nheads;
  int64_t sb, ss, sh;
  int softplus;
  float lo, hi;
  bool operator==(const DtTag& o) const {
    // A and dt_bias are often fp32 temporaries of bf16 parameters (a new address per call): compared by presence only
    return dt == o.dt && (A == nullptr) == (o.A == nullptr) && (bias == nullptr) == (o.bias == nullptr) &&
           batch == o.batch && seqlen == o.seqlen && nheads == o.nheads && sb == o.sb && ss == o.ss && sh == o.sh &&
           softplus == o.softplus && lo == o.lo && hi == o.hi;
  }
};
static std::mutex g_dt_tags_mu;
static std::unordered_map<const void*, DtTag> g_dt_tags;

bool tc_supported(const tv_ssd_params& p) {
  // chunk_size: any multiple of 128 (the reference class default is 256, configuration_nano.py:137-175).  The chunk size only
  // sets where the recurrence is re-anchored -- the result is the same function of the inputs, up to rounding -- so the kernel
  // always walks chunks of 128 tokens.
  if (p.dtype != TV_BF16 || p.headdim != tc::P || p.dstate != tc::N || p.chunk_size <= 0 || p.chunk_size % tc::Q != 0) return false;
  if (p.nheads % p.ngroups != 0) return false;
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  if (!al16(p.x) || !al16(p.B) || (p.mode == TV_SSD_FULL && (!al16(p.C) || ((uintptr_t)p.out & 31) != 0))) return false;
  if (p.z != nullptr && (!al16(p.z) || p.z_head_stride % 8 || p.z_seq_stride % 8 || p.z_batch_stride % 8)) return false;
  auto m8 = [](int64_t v) { return v % 8 == 0; };
  if (!m8(p.x_head_stride) || !m8(p.x_seq_stride) || !m8(p.x_batch_stride) || !m8(p.b_group_stride) ||
      !m8(p.b_seq_stride) || !m8(p.b_batch_stride))
    return false;
  if (p.mode == TV_SSD_FULL && (!m8(p.c_group_stride) || !m8(p.c_seq_stride) || !m8(p.c_batch_stride))) return false;
  return get_encode_tiled() != nullptr;
}

size_t tc_workspace_bytes(const tv_ssd_params& p) {
  const int64_t nchunks = ceil_div(p.seqlen, tc::Q);
  const size_t per = (size_t)p.batch * nchunks * p.nheads * tc::Q * sizeof(float);
  // dt | cumsum | first_chunk | partial shard summaries.  The size does not depend on `mode`: the sharded path calls
  // DT_ONLY, STATE_ONLY and FULL on ONE workspace (reuse_dt_cumsum), which must not be re-allocated in between.
  return 2 * ((per + 255) & ~(size_t)255) + ((((size_t)p.batch * p.nheads * sizeof(int)) + 255) & ~(size_t)255) +
         (size_t)p.batch * p.nheads * kMaxStateSplit * tc::P * tc::N * sizeof(float);
}

int ssd_tc_forward(const tv_ssd_params& p_in, void* workspace, cudaStream_t s) {
  using namespace tc;
  tv_ssd_params p = p_in;
  p.chunk_size = Q;                              // internal chunks of 128 tokens, whatever multiple the caller asked for
  const int nchunks = (int)ceil_div(p.seqlen, Q);
  const size_t per = (((size_t)p.batch * nchunks * p.nheads * Q * sizeof(float)) + 255) & ~(size_t)255;
  float* dt_act = (float*)workspace;
  float* cs = (float*)((char*)workspace + per);
  // `reuse_dt_cumsum` trusts that `workspace` still holds dt / cumsum of a previous call: remember which inputs the
  // arrays in a given workspace were computed from and refuse a reuse that does not match (another dt tensor, other dims
  // or strides, another limit, or a workspace this library has never filled).  Contents changed in place behind the same
  // pointer, or other A / dt_bias values, cannot be seen from here.
  {
    const DtTag tag{p.dt, p.A, p.dt_bias, p.batch, p.seqlen, p.nheads, p.dt_batch_stride, p.dt_seq_stride,
                    p.dt_head_stride, p.dt_softplus, p.dt_min, p.dt_max};
    std::lock_guard<std::mutex> lock(g_dt_tags_mu);
    if (!p.reuse_dt_cumsum) {
      g_dt_tags[workspace] = tag;
    } else {
      auto it = g_dt_tags.find(workspace);
      if (it == g_dt_tags.end() || !(it->second == tag)) {
        set_error("ssd: reuse_dt_cumsum, but this workspace does not hold the dt/cumsum arrays of these inputs "
                  "(another dt tensor, other dims / strides / dt limits, or a reallocated workspace)");
        return TV_ERR_INVALID;
      }
    }
  }
  if (!p.reuse_dt_cumsum) {
    int rc = launch_dt_cumsum(p, dt_act, cs, s);
    if (rc != TV_OK) return rc;
  }
  if (p.mode == TV_SSD_DT_ONLY) return TV_OK;

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
      const uint64_t dy[4] = {(uint64_t)P, (uint64_t)p.nheads, L, (uint64_t)p.batch};
      const uint64_t sy[3] = {(uint64_t)P * 2, (uint64_t)p.nheads * P * 2, (uint64_t)L * p.nheads * P * 2};
      const uint32_t by[4] = {16, 1, 32, 1};          // one epilogue warp's rows per store
      if (!encode_bf16_tmap(&maps.y, p.out, 4, dy, sy, by, CU_TENSOR_MAP_SWIZZLE_32B)) {
        set_error("ssd(tcgen05): cuTensorMapEncodeTiled(out) failed");
        return TV_ERR_CUDA;
      }
    } else {
      maps.c = maps.b;
      maps.y = maps.b;
    }
  }
  Args a;
  a.dt_act = dt_act; a.cs = cs; a.D = p.D; a.z = (const __nv_bfloat16*)p.z; a.init = p.initial_states;
  a.out = (__nv_bfloat16*)p.out; a.fin = p.final_states; a.logdecay = p.logdecay_sum;
  a.L = p.seqlen; a.H = p.nheads; a.G = p.ngroups; a.nchunks = nchunks; a.d_has_hdim = p.d_has_hdim;
  a.trace = (long long*)g_trace_ptr;
  a.ablate = g_ablate;
  a.zbs = p.z_batch_stride; a.zss = p.z_seq_stride; a.zhs = p.z_head_stride;
  a.xp = (const __nv_bfloat16*)p.x; a.bp = (const __nv_bfloat16*)p.B; a.cp = (const __nv_bfloat16*)p.C;
  a.xbs = p.x_batch_stride; a.xss = p.x_seq_stride; a.xhs = p.x_head_stride;
  a.bbs = p.b_batch_stride; a.bss = p.b_seq_stride; a.bgs = p.b_group_stride;
  a.cbs = p.c_batch_stride; a.css = p.c_seq_stride; a.cgs = p.c_group_stride;
  dim3 grid(p.nheads, p.batch);
  auto launch = [&](auto kern) -> int {
    TV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    kern<<<grid, THREADS, SMEM_BYTES, s>>>(maps, a);
    return TV_OK;
  };
  int lrc;
  const bool dfold = !p.d_has_hdim;          // scalar-per-head D (or no D): fold into M's diagonal
  if (p.mode != TV_SSD_FULL) {
    int* first_chunk = (int*)((char*)workspace + 2 * per);
    const int BH = p.batch * p.nheads;
    ssd_suffix_scan_kernel<<<(BH + 3) / 4, 128, 0, s>>>(cs, p.logdecay_sum, first_chunk, BH, p.nheads, nchunks, Q);
    TV_LAUNCH_OK();
    TV_CUDA_OK(cudaFuncSetAttribute(ssd_state_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st::SMEM_BYTES));
    // up to 8 CTAs per head, >= 8 chunks of the walk each
    const int split = std::max(1, std::min(kMaxStateSplit, nchunks / 8));
    float* partial = (float*)((char*)workspace + 2 * per + ((((size_t)BH * sizeof(int)) + 255) & ~(size_t)255));
    ssd_state_kernel<<<dim3(p.nheads, p.batch, split), st::THREADS, st::SMEM_BYTES, s>>>(maps, a, first_chunk, split, partial);
    if (split > 1) {
      TV_LAUNCH_OK();
      const int PN4 = P * N / 4;
      ssd_state_reduce_kernel<<<dim3((unsigned)ceil_div(PN4, 256), (unsigned)BH), 256, 0, s>>>(partial, first_chunk,
                                                                                            p.final_states, nchunks, split, PN4);
    }
    lrc = TV_OK;
  }
  else if (p.z != nullptr) lrc = dfold ? launch(ssd_fused_kernel<true, true>) : launch(ssd_fused_kernel<true, false>);
  else lrc = dfold ? launch(ssd_fused_kernel<false, true>) : launch(ssd_fused_kernel<false, false>);
  if (lrc != TV_OK) return lrc;
  TV_LAUNCH_OK();
  return TV_OK;
}

}  // namespace tv

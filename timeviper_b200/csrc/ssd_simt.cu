// SSD chunked scan on CUDA cores with fp32 arithmetic (sm_100a).
//
// This is the fp32-accuracy implementation of mamba_chunk_scan_combined (stages of
// visualize/nano/my_ssd_combined.py:743-843): it serves fp32 inputs (north_star tolerance 1e-4, which a
// TF32/BF16 tensor-core contraction cannot meet) and every shape the tcgen05 kernel is not specialised
// for.  The bf16 Nanov2-9B shape (P=80, N=128, Q=128) runs on ssd_tc.cu instead.
//
// Stages (one kernel each, same decomposition as the reference so intermediates can be compared):
//   (i)   dt = clamp(softplus(dt + bias)); dA = dt*A; per-chunk inclusive cumsum        [ssd_dt_cumsum]
//   (ii)  S_c[p,n] = sum_k x[k,p] * dt_k * exp(cs_last - cs_k) * B[k,n]                 [ssd_chunk_state]
//   (iii) s <- exp(cs_last(c)) * s + S_c ; stores the state ENTERING each chunk          [ssd_state_passing]
//   (iv)  CB[m,k] = sum_n C[m,n] B[k,n]   (lower block-triangle only)                   [ssd_bmm_chunk]
//   (v)   y[m,p] = sum_{k<=m} CB[m,k] exp(cs_m - cs_k) dt_k x[k,p] + exp(cs_m) C_m . s_c + D x[m,p]  [ssd_chunk_scan]
// Tokens past seqlen behave as zero padding: dt = 0 (no decay, no contribution), x = 0.
#include "common.cuh"
#include "ssd.h"

namespace tv {

// ---------------------------------------------------------------------------------------------
// (i) dt activation + per-chunk cumulative sum.  grid (nchunks, ceil(H/32), batch), 128 threads.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
ssd_dt_cumsum_kernel(const T* __restrict__ dt, const float* __restrict__ A, const float* __restrict__ dt_bias,
                     float* __restrict__ dt_out, float* __restrict__ cs_out, int L, int H, int Q, int nchunks,
                     int64_t dbs, int64_t dss, int64_t dhs, int softplus, float dt_min, float dt_max) {
  extern __shared__ float sm[];  // [Q][33]
  const int c = blockIdx.x, h0 = blockIdx.y * 32, b = blockIdx.z;
  const int t0 = c * Q;
  for (int i = threadIdx.x; i < Q * 32; i += blockDim.x) {
    const int q = i >> 5, hh = i & 31;
    float v = 0.f;
    bool valid = (t0 + q < L) && (h0 + hh < H);
    if (valid) {
      v = to_f32<T>(dt[b * dbs + (int64_t)(t0 + q) * dss + (int64_t)(h0 + hh) * dhs]);
      if (dt_bias != nullptr) v += dt_bias[h0 + hh];
      if (softplus && v <= 20.f) v = log1pf(expf(v));
      v = fminf(fmaxf(v, dt_min), dt_max);
    }
    sm[q * 33 + hh] = valid ? v : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = Q / 32;  // Q is a multiple of 32
  for (int hh = warp; hh < 32 && h0 + hh < H; hh += 4) {
    const float a = A[h0 + hh];
    float run = 0.f;
    float loc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < per) { run += sm[(lane * per + j) * 33 + hh] * a; loc[j] = run; }
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    const float excl = incl - run;
    const int64_t base = (((int64_t)b * nchunks + c) * H + h0 + hh) * Q + lane * per;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < per) {
        dt_out[base + j] = sm[(lane * per + j) * 33 + hh];
        cs_out[base + j] = excl + loc[j];
      }
  }
}

// bf16 fast path of (i) for unit head stride: a warp reads 64 heads of one token as ONE 128-byte request (two heads
// per lane), the activated dt goes through shared memory as two [Q][33] planes (even / odd heads, conflict-free both
// ways), and every head's chunk is scanned 32 tokens at a time so that dt_out / cs_out are written as full 128-byte
// lines.  grid (nchunks, ceil(H/64), batch), 256 threads.  168 MB of traffic at 128K tokens x 128 heads.
__global__ void __launch_bounds__(256)
ssd_dt_cumsum_bf16x2_kernel(const __nv_bfloat16* __restrict__ dt, const float* __restrict__ A,
                            const float* __restrict__ dt_bias, float* __restrict__ dt_out, float* __restrict__ cs_out,
                            int L, int H, int Q, int nchunks, int64_t dbs, int64_t dss, int softplus, float dt_min,
                            float dt_max) {
  extern __shared__ float sm[];  // [2][Q][33]
  const int c = blockIdx.x, h0 = blockIdx.y * 64, b = blockIdx.z;
  const int t0 = c * Q;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sm1 = sm + Q * 33;
  {
    const int h = h0 + 2 * lane;                      // H is even on this path
    const bool hv = h < H;
    const float b0 = (dt_bias != nullptr && hv) ? dt_bias[h] : 0.f, b1 = (dt_bias != nullptr && hv) ? dt_bias[h + 1] : 0.f;
    const __nv_bfloat16* src = dt + b * dbs + h;
    for (int q = warp; q < Q; q += 8) {
      float v0 = 0.f, v1 = 0.f;
      if (hv && t0 + q < L) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(src + (int64_t)(t0 + q) * dss);
        v0 = __uint_as_float(w << 16) + b0;
        v1 = __uint_as_float(w & 0xffff0000u) + b1;
        if (softplus && v0 <= 20.f) v0 = log1pf(expf(v0));
        if (softplus && v1 <= 20.f) v1 = log1pf(expf(v1));
        v0 = fminf(fmaxf(v0, dt_min), dt_max);
        v1 = fminf(fmaxf(v1, dt_min), dt_max);
      }
      sm[q * 33 + lane] = v0;
      sm1[q * 33 + lane] = v1;
    }
  }
  __syncthreads();
  for (int hh = warp; hh < 64 && h0 + hh < H; hh += 8) {
    const float a = A[h0 + hh];
    const float* plane = (hh & 1) ? sm1 : sm;
    const int col = hh >> 1;
    const int64_t base = (((int64_t)b * nchunks + c) * H + h0 + hh) * Q;
    float carry = 0.f;
    for (int q0 = 0; q0 < Q; q0 += 32) {
      const float v = plane[(q0 + lane) * 33 + col];
      float incl = v * a;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      dt_out[base + q0 + lane] = v;
      cs_out[base + q0 + lane] = carry + incl;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Register-tiled smem GEMM: acc[i][j] += sum_k A[k][ty+16i] * B[k][tx+16j]   (256 threads = 16 x 16)
// ---------------------------------------------------------------------------------------------
template <int TI, int TJ>
__device__ __forceinline__ void tile_fma(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                         int ldb, int K, int ty, int tx, float (&acc)[TI][TJ]) {
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    float a[TI], bb[TJ];
#pragma unroll
    for (int i = 0; i < TI; ++i) a[i] = A[k * lda + ty + 16 * i];
#pragma unroll
    for (int j = 0; j < TJ; ++j) bb[j] = B[k * ldb + tx + 16 * j];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
      for (int j = 0; j < TJ; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
  }
}

// ---------------------------------------------------------------------------------------------
// (ii) chunk states.  grid (nchunks, H, batch), 256 threads.  smem: xs[Q][16*TP] | Bs[Q][128] | w[Q]
// ---------------------------------------------------------------------------------------------
template <typename T, int TP>
__global__ void __launch_bounds__(256)
ssd_chunk_state_kernel(const T* __restrict__ x, const T* __restrict__ Bm, const float* __restrict__ dt_act,
                       const float* __restrict__ cs, float* __restrict__ states, int L, int H, int P, int G,
                       int N, int Q, int nchunks, int64_t xbs, int64_t xss, int64_t xhs, int64_t bbs,
                       int64_t bss, int64_t bgs) {
  constexpr int LDP = 16 * TP, LDN = 128;
  extern __shared__ float sm[];
  float* xs = sm;
  float* Bs = xs + Q * LDP;
  float* w = Bs + Q * LDN;
  const int c = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int g = h / (H / G);
  const int t0 = c * Q;
  const int64_t sc = (((int64_t)b * nchunks + c) * H + h) * Q;
  const float cs_last = cs[sc + Q - 1];
  for (int k = threadIdx.x; k < Q; k += 256) w[k] = dt_act[sc + k] * expf(cs_last - cs[sc + k]);
  __syncthreads();
  for (int i = threadIdx.x; i < Q * LDP; i += 256) {
    const int k = i / LDP, p = i - k * LDP;
    float v = 0.f;
    if (p < P && t0 + k < L) v = to_f32<T>(x[b * xbs + (int64_t)(t0 + k) * xss + h * xhs + p]) * w[k];
    xs[i] = v;
  }
  for (int i = threadIdx.x; i < Q * LDN; i += 256) {
    const int k = i >> 7, n = i & 127;
    float v = 0.f;
    if (n < N && t0 + k < L) v = to_f32<T>(Bm[b * bbs + (int64_t)(t0 + k) * bss + g * bgs + n]);
    Bs[i] = v;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[TP][8];
#pragma unroll
  for (int i = 0; i < TP; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  tile_fma<TP, 8>(xs, LDP, Bs, LDN, Q, ty, tx, acc);
  float* so = states + (((int64_t)b * nchunks + c) * H + h) * (int64_t)P * N;
#pragma unroll
  for (int i = 0; i < TP; ++i) {
    const int p = ty + 16 * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = tx + 16 * j;
      if (p < P && n < N) so[(int64_t)p * N + n] = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (iii) inter-chunk recurrence.  grid (ceil(P*N/256), H, batch).  In place: states[c] <- state entering c.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ssd_state_passing_kernel(float* __restrict__ states, const float* __restrict__ cs, const float* __restrict__ init,
                         float* __restrict__ fin, float* __restrict__ logdecay_sum, int H, int PN, int Q,
                         int nchunks, int write_entering) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  if (e >= PN) return;
  float s = init != nullptr ? init[((int64_t)b * H + h) * PN + e] : 0.f;
  float logsum = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const int64_t bc = ((int64_t)b * nchunks + c) * H + h;
    const float dA = cs[bc * Q + Q - 1];
    float* sp = states + bc * PN + e;
    const float sc = *sp;
    if (write_entering) *sp = s;
    s = fmaf(expf(dA), s, sc);
    logsum += dA;
  }
  if (fin != nullptr) fin[((int64_t)b * H + h) * PN + e] = s;
  if (logdecay_sum != nullptr && e == 0) logdecay_sum[(int64_t)b * H + h] = logsum;
}

// ---------------------------------------------------------------------------------------------
// (iv) CB = C B^T per (chunk, group), 64x64 output blocks on or below the diagonal.
// grid (nblk*(nblk+1)/2, nchunks*G, batch), 256 threads.  smem: Ct[128][65] | Bt[128][65]
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
ssd_bmm_chunk_kernel(const T* __restrict__ Cm, const T* __restrict__ Bm, float* __restrict__ CB, int L, int G,
                     int N, int Q, int nchunks, int64_t cbs, int64_t css, int64_t cgs, int64_t bbs, int64_t bss,
                     int64_t bgs) {
  constexpr int LD = 65;
  extern __shared__ float sm[];
  float* Ct = sm;
  float* Bt = sm + 128 * LD;
  int mb = 0, kb = blockIdx.x;  // triangular index -> (mb, kb), kb <= mb
  while (kb > mb) { kb -= mb + 1; ++mb; }
  const int c = blockIdx.y / G, g = blockIdx.y % G, b = blockIdx.z;
  const int t0 = c * Q;
  for (int i = threadIdx.x; i < 64 * 128; i += 256) {
    const int r = i >> 7, n = i & 127;
    const int tm = t0 + mb * 64 + r, tk = t0 + kb * 64 + r;
    float vc = 0.f, vb = 0.f;
    if (n < N && tm < L) vc = to_f32<T>(Cm[b * cbs + (int64_t)tm * css + g * cgs + n]);
    if (n < N && tk < L) vb = to_f32<T>(Bm[b * bbs + (int64_t)tk * bss + g * bgs + n]);
    Ct[n * LD + r] = vc;
    Bt[n * LD + r] = vb;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  tile_fma<4, 4>(Ct, LD, Bt, LD, N, ty, tx, acc);
  float* o = CB + (((int64_t)b * nchunks + c) * G + g) * (int64_t)Q * Q;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[(int64_t)(mb * 64 + ty + 16 * i) * Q + kb * 64 + tx + 16 * j] = acc[i][j];
}

// ---------------------------------------------------------------------------------------------
// (v) chunk output.  grid (nchunks * Q/64, H, batch), 256 threads; one CTA = 64 output rows of one head.
// smem: Mt[Q][65] | xs[Q][16*TP] | Ct[128][65] | St[128][16*TP+1] | cs[Q] | dt[Q]
// ---------------------------------------------------------------------------------------------
template <typename T, int TP>
__global__ void __launch_bounds__(256)
ssd_chunk_scan_kernel(const T* __restrict__ x, const T* __restrict__ Cm, const T* __restrict__ z,
                      const float* __restrict__ dt_act, const float* __restrict__ cs_g,
                      const float* __restrict__ CB, const float* __restrict__ states, const float* __restrict__ D,
                      T* __restrict__ out, int L, int H, int P, int G, int N, int Q, int nchunks, int64_t xbs,
                      int64_t xss, int64_t xhs, int64_t cbs, int64_t css, int64_t cgs, int64_t zbs, int64_t zss,
                      int64_t zhs, int d_has_hdim) {
  constexpr int LDM = 65, LDP = 16 * TP, LDS = 16 * TP + 1;
  constexpr bool FAST = sizeof(T) == 2;
  extern __shared__ float sm[];
  float* Mt = sm;
  float* xs = Mt + Q * LDM;
  float* Ct = xs + Q * LDP;
  float* St = Ct + 128 * LDM;
  float* cs = St + 128 * LDS;
  float* dts = cs + Q;
  const int nmb = Q / 64;
  const int c = blockIdx.x / nmb, mb = blockIdx.x % nmb, h = blockIdx.y, b = blockIdx.z;
  const int g = h / (H / G);
  const int t0 = c * Q, m0 = mb * 64;
  const int kmax = m0 + 64;  // causal: rows of this block only see k < kmax
  const int64_t sc = (((int64_t)b * nchunks + c) * H + h) * Q;
  for (int k = threadIdx.x; k < Q; k += 256) { cs[k] = cs_g[sc + k]; dts[k] = dt_act[sc + k]; }
  __syncthreads();
  // Mt[k][mi] = CB[m][k] * exp(cs_m - cs_k) * dt_k  for k <= m, else 0
  const float* cb = CB + (((int64_t)b * nchunks + c) * G + g) * (int64_t)Q * Q;
  for (int i = threadIdx.x; i < 64 * kmax; i += 256) {
    const int mi = i / kmax, k = i - mi * kmax;
    const int m = m0 + mi;
    float v = 0.f;
    if (k <= m) v = cb[(int64_t)m * Q + k] * expf(cs[m] - cs[k]) * dts[k];
    Mt[k * LDM + mi] = v;
  }
  for (int i = threadIdx.x; i < kmax * LDP; i += 256) {
    const int k = i / LDP, p = i - k * LDP;
    float v = 0.f;
    if (p < P && t0 + k < L) v = to_f32<T>(x[b * xbs + (int64_t)(t0 + k) * xss + h * xhs + p]);
    xs[i] = v;
  }
  // Ct[n][mi] = C[m][n] * exp(cs_m)
  for (int i = threadIdx.x; i < 64 * 128; i += 256) {
    const int mi = i >> 7, n = i & 127;
    float v = 0.f;
    if (n < N && t0 + m0 + mi < L)
      v = to_f32<T>(Cm[b * cbs + (int64_t)(t0 + m0 + mi) * css + g * cgs + n]) * expf(cs[m0 + mi]);
    Ct[n * LDM + mi] = v;
  }
  // St[n][p] = entering state [p][n]
  const float* sp = states + (((int64_t)b * nchunks + c) * H + h) * (int64_t)P * N;
  for (int i = threadIdx.x; i < 128 * LDP; i += 256) {
    const int p = i >> 7, n = i & 127;  // n fastest: coalesced global read
    float v = 0.f;
    if (p < P && n < N) v = sp[(int64_t)p * N + n];
    St[n * LDS + p] = v;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][TP];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TP; ++j) acc[i][j] = 0.f;
  tile_fma<4, TP>(Mt, LDM, xs, LDP, kmax, ty, tx, acc);
  tile_fma<4, TP>(Ct, LDM, St, LDS, N, ty, tx, acc);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 16 * i;
    const int t = t0 + m;
    if (t >= L) continue;
#pragma unroll
    for (int j = 0; j < TP; ++j) {
      const int p = tx + 16 * j;
      if (p >= P) continue;
      float y = acc[i][j];
      if (D != nullptr) y = fmaf(d_has_hdim ? D[h * P + p] : D[h], xs[m * LDP + p], y);
      if (z != nullptr) y *= silu<FAST>(to_f32<T>(z[b * zbs + (int64_t)t * zss + h * zhs + p]));
      out[(((int64_t)b * L + t) * H + h) * P + p] = from_f32<T>(y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

SimtWorkspace simt_workspace_layout(const tv_ssd_params& p) {
  SimtWorkspace w{};
  const int64_t nchunks = ceil_div(p.seqlen, p.chunk_size);
  const size_t per = (size_t)p.batch * nchunks * p.nheads * p.chunk_size * sizeof(float);
  w.dt_off = 0;
  w.cs_off = align256(per);
  w.states_off = w.cs_off + align256(per);
  const size_t st = (size_t)p.batch * nchunks * p.nheads * p.headdim * p.dstate * sizeof(float);
  w.cb_off = w.states_off + align256(st);
  const size_t cb = p.mode == TV_SSD_STATE_ONLY
                        ? 0
                        : (size_t)p.batch * nchunks * p.ngroups * p.chunk_size * p.chunk_size * sizeof(float);
  w.total = w.cb_off + align256(cb);
  return w;
}

int simt_supported(const tv_ssd_params& p) {
  if (p.headdim > 128 || p.dstate > 128) {
    set_error("ssd(simt): headdim %d / dstate %d > 128 unsupported", p.headdim, p.dstate);
    return TV_ERR_UNSUPPORTED;
  }
  if (p.chunk_size != 64 && p.chunk_size != 128 && p.chunk_size != 256) {
    set_error("ssd(simt): chunk_size %d unsupported (64, 128, 256)", p.chunk_size);
    return TV_ERR_UNSUPPORTED;
  }
  return TV_OK;
}

int launch_dt_cumsum(const tv_ssd_params& p, float* dt_act, float* cs, cudaStream_t s) {
  const int Q = p.chunk_size, H = p.nheads, L = p.seqlen;
  const int nchunks = (int)ceil_div(L, Q);
  dim3 grid(nchunks, (unsigned)ceil_div(H, 32), p.batch);
  if (p.dtype == TV_BF16 && p.dt_head_stride == 1 && H % 2 == 0 && Q % 32 == 0 && p.dt_seq_stride % 2 == 0 &&
      p.dt_batch_stride % 2 == 0 && ((uintptr_t)p.dt & 3) == 0) {
    dim3 grid2(nchunks, (unsigned)ceil_div(H, 64), p.batch);
    if (2 * Q * 33 * sizeof(float) > 48 * 1024)
      TV_CUDA_OK(cudaFuncSetAttribute(ssd_dt_cumsum_bf16x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(2 * Q * 33 * sizeof(float))));
    ssd_dt_cumsum_bf16x2_kernel<<<grid2, 256, 2 * Q * 33 * sizeof(float), s>>>(
        (const __nv_bfloat16*)p.dt, p.A, p.dt_bias, dt_act, cs, L, H, Q, nchunks, p.dt_batch_stride, p.dt_seq_stride,
        p.dt_softplus, p.dt_min, p.dt_max);
  } else if (p.dtype == TV_BF16)
    ssd_dt_cumsum_kernel<__nv_bfloat16><<<grid, 128, Q * 33 * sizeof(float), s>>>(
        (const __nv_bfloat16*)p.dt, p.A, p.dt_bias, dt_act, cs, L, H, Q, nchunks, p.dt_batch_stride,
        p.dt_seq_stride, p.dt_head_stride, p.dt_softplus, p.dt_min, p.dt_max);
  else
    ssd_dt_cumsum_kernel<float><<<grid, 128, Q * 33 * sizeof(float), s>>>(
        (const float*)p.dt, p.A, p.dt_bias, dt_act, cs, L, H, Q, nchunks, p.dt_batch_stride, p.dt_seq_stride,
        p.dt_head_stride, p.dt_softplus, p.dt_min, p.dt_max);
  TV_LAUNCH_OK();
  return TV_OK;
}

template <typename T, int TP>
static int run_simt(const tv_ssd_params& p, char* ws, cudaStream_t s) {
  const SimtWorkspace w = simt_workspace_layout(p);
  const int Q = p.chunk_size, H = p.nheads, P = p.headdim, N = p.dstate, G = p.ngroups, L = p.seqlen;
  const int nchunks = (int)ceil_div(L, Q);
  float* dt_act = (float*)(ws + w.dt_off);
  float* cs = (float*)(ws + w.cs_off);
  float* states = (float*)(ws + w.states_off);
  float* CB = (float*)(ws + w.cb_off);
  {
    const int rc = launch_dt_cumsum(p, dt_act, cs, s);
    if (rc != TV_OK) return rc;
  }
  {
    const size_t smem = ((size_t)Q * 16 * TP + (size_t)Q * 128 + Q) * sizeof(float);
    if (smem > kMaxDynSmem) {
      set_error("ssd(simt): chunk_size %d with headdim %d needs %zu B of shared memory", Q, P, smem);
      return TV_ERR_UNSUPPORTED;
    }
    auto k = ssd_chunk_state_kernel<T, TP>;
    TV_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nchunks, H, p.batch);
    k<<<grid, 256, smem, s>>>((const T*)p.x, (const T*)p.B, dt_act, cs, states, L, H, P, G, N, Q, nchunks,
                              p.x_batch_stride, p.x_seq_stride, p.x_head_stride, p.b_batch_stride,
                              p.b_seq_stride, p.b_group_stride);
    TV_LAUNCH_OK();
  }
  {
    const int PN = P * N;
    dim3 grid((unsigned)ceil_div(PN, 256), H, p.batch);
    ssd_state_passing_kernel<<<grid, 256, 0, s>>>(states, cs, p.initial_states, p.final_states, p.logdecay_sum,
                                                  H, PN, Q, nchunks, p.mode == TV_SSD_FULL ? 1 : 0);
    TV_LAUNCH_OK();
  }
  if (p.mode == TV_SSD_STATE_ONLY) return TV_OK;
  {
    const int nblk = Q / 64;
    const size_t smem = 2 * 128 * 65 * sizeof(float);
    auto k = ssd_bmm_chunk_kernel<T>;
    TV_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nblk * (nblk + 1) / 2, nchunks * G, p.batch);
    k<<<grid, 256, smem, s>>>((const T*)p.C, (const T*)p.B, CB, L, G, N, Q, nchunks, p.c_batch_stride,
                              p.c_seq_stride, p.c_group_stride, p.b_batch_stride, p.b_seq_stride,
                              p.b_group_stride);
    TV_LAUNCH_OK();
  }
  {
    const size_t smem =
        ((size_t)Q * 65 + (size_t)Q * 16 * TP + 128 * 65 + 128 * (16 * TP + 1) + 2 * Q) * sizeof(float);
    if (smem > kMaxDynSmem) {
      set_error("ssd(simt): chunk_size %d with headdim %d needs %zu B of shared memory", Q, P, smem);
      return TV_ERR_UNSUPPORTED;
    }
    auto k = ssd_chunk_scan_kernel<T, TP>;
    TV_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nchunks * (Q / 64), H, p.batch);
    k<<<grid, 256, smem, s>>>((const T*)p.x, (const T*)p.C, (const T*)p.z, dt_act, cs, CB, states, p.D,
                              (T*)p.out, L, H, P, G, N, Q, nchunks, p.x_batch_stride, p.x_seq_stride,
                              p.x_head_stride, p.c_batch_stride, p.c_seq_stride, p.c_group_stride,
                              p.z_batch_stride, p.z_seq_stride, p.z_head_stride, p.d_has_hdim);
    TV_LAUNCH_OK();
  }
  return TV_OK;
}

int ssd_simt_forward(const tv_ssd_params& p, void* workspace, cudaStream_t s) {
  char* ws = (char*)workspace;
  const bool small_p = p.headdim <= 80;
  if (p.dtype == TV_BF16)
    return small_p ? run_simt<__nv_bfloat16, 5>(p, ws, s) : run_simt<__nv_bfloat16, 8>(p, ws, s);
  return small_p ? run_simt<float, 5>(p, ws, s) : run_simt<float, 8>(p, ws, s);
}

// ---------------------------------------------------------------------------------------------
// boundary-state fold for the sequence-sharded path
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fold_boundary_kernel(const float* __restrict__ states, const float* __restrict__ logdecay,
                     const float* __restrict__ init, float* __restrict__ out, int rank, int64_t srs, int64_t lrs,
                     int PN4) {
  // one thread = 4 consecutive state elements (PN % 4 == 0); the decay chain is re-evaluated per thread (rank <= 7)
  const int64_t bh = blockIdx.y;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= PN4) return;
  const int64_t off = bh * PN4 + e;
  float4 s = init != nullptr ? reinterpret_cast<const float4*>(init)[off] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < rank; ++r) {
    const float d = expf(logdecay[(int64_t)r * lrs + bh]);
    const float4 v = *reinterpret_cast<const float4*>(states + (int64_t)r * srs + off * 4);
    s.x = fmaf(d, s.x, v.x); s.y = fmaf(d, s.y, v.y); s.z = fmaf(d, s.z, v.z); s.w = fmaf(d, s.w, v.w);
  }
  reinterpret_cast<float4*>(out)[off] = s;
}

// The same fold reading each rank's summary THROUGH ITS OWN POINTER -- peer memory of the other GPUs (NVLink / NVSwitch
// loads), no gathered copy: the boundary-state exchange and the fold are one kernel.  The decay chain is walked from
// rank-1 downwards first: once the accumulated log-decay is below the fp32 underflow point the earlier summaries would be
// multiplied by exactly 0, so they are never fetched -- with realistic decay rates a rank reads little more than its
// left neighbour's 5 MB instead of all rank*5 MB.
struct FoldPeers {
  const float* s[16];
  const float* lp[16];
};

// one block per (b, h); FOLD_E float4 per thread are in flight at a time (peer loads have ~2 us of latency: the kernel is
// a copy with as many independent loads outstanding as the registers allow, not a chain of dependent round trips)
constexpr int FOLD_E = 10;
__global__ void __launch_bounds__(256)
fold_boundary_p2p_kernel(const FoldPeers peers, const float* __restrict__ init, float* __restrict__ out, int rank, int PN4) {
  const int64_t bh = blockIdx.x;
  // all log-decays of the earlier ranks at once (independent loads, one round trip)
  float lp[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) lp[r] = r < rank ? __ldg(peers.lp[r] + bh) : 0.f;
  // first summary whose weight exp(sum of logdecay of the ranks after it) is still representable
  int j0 = rank;                       // rank: "start from init"
  float acc_log = 0.f;
#pragma unroll
  for (int r = 15; r >= 0; --r)
    if (r < rank && acc_log >= -104.f) { j0 = r; acc_log += lp[r]; }
  const bool init_live = acc_log >= -104.f && init != nullptr;
  const float4* ini = reinterpret_cast<const float4*>(init) + bh * PN4;
  float4* dst = reinterpret_cast<float4*>(out) + bh * PN4;
  for (int e0 = threadIdx.x; e0 < PN4; e0 += 256 * FOLD_E) {
    float4 s[FOLD_E];
#pragma unroll
    for (int i = 0; i < FOLD_E; ++i) {
      const int e = e0 + i * 256;
      s[i] = (init_live && e < PN4) ? ini[e] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int r = j0; r < rank; ++r) {
      const float d = expf(lp[r]);
      const float4* src = reinterpret_cast<const float4*>(peers.s[r]) + bh * PN4;
      float4 v[FOLD_E];
#pragma unroll
      for (int i = 0; i < FOLD_E; ++i) {
        const int e = e0 + i * 256;
        v[i] = e < PN4 ? __ldg(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < FOLD_E; ++i) {
        s[i].x = fmaf(d, s[i].x, v[i].x); s[i].y = fmaf(d, s[i].y, v[i].y);
        s[i].z = fmaf(d, s[i].z, v[i].z); s[i].w = fmaf(d, s[i].w, v[i].w);
      }
    }
#pragma unroll
    for (int i = 0; i < FOLD_E; ++i) {
      const int e = e0 + i * 256;
      if (e < PN4) dst[e] = s[i];
    }
  }
}

}  // namespace tv

extern "C" int tv_ssd_fold_boundary_states_p2p(const void* const* state_ptrs, const void* const* logdecay_ptrs,
                                               const float* initial, float* out, int32_t rank, int32_t batch,
                                               int32_t nheads, int32_t headdim, int32_t dstate, void* stream) {
  using namespace tv;
  TV_CHECK_ARG(out != nullptr && rank >= 0 && rank <= 16 && batch > 0 && nheads > 0 && headdim > 0 && dstate > 0,
               "fold_boundary_states_p2p: bad arguments (at most 16 ranks)");
  TV_CHECK_ARG(rank == 0 || (state_ptrs != nullptr && logdecay_ptrs != nullptr), "fold_boundary_states_p2p: null pointer tables");
  const int64_t PN = (int64_t)headdim * dstate, BH = (int64_t)batch * nheads;
  FoldPeers peers;
  for (int r = 0; r < 16; ++r) {
    peers.s[r] = r < rank ? (const float*)state_ptrs[r] : nullptr;
    peers.lp[r] = r < rank ? (const float*)logdecay_ptrs[r] : nullptr;
    if (r < rank)
      TV_CHECK_ARG(peers.s[r] != nullptr && peers.lp[r] != nullptr && (uintptr_t)peers.s[r] % 16 == 0,
                   "fold_boundary_states_p2p: summary pointer of rank %d is null or not 16-byte aligned", r);
  }
  TV_CHECK_ARG(PN % 4 == 0 && (uintptr_t)out % 16 == 0 && (initial == nullptr || (uintptr_t)initial % 16 == 0),
               "fold_boundary_states_p2p: headdim*dstate must be a multiple of 4, out/initial 16-byte aligned");
  fold_boundary_p2p_kernel<<<(unsigned)BH, 256, 0, (cudaStream_t)stream>>>(peers, initial, out, rank, (int)(PN / 4));
  TV_LAUNCH_OK();
  return TV_OK;
}

extern "C" int tv_ssd_fold_boundary_states(const float* states, const float* logdecay, const float* initial,
                                           float* out, int32_t rank, int32_t batch, int32_t nheads,
                                           int32_t headdim, int32_t dstate, int64_t states_rank_stride,
                                           int64_t logdecay_rank_stride, void* stream) {
  using namespace tv;
  TV_CHECK_ARG(out != nullptr && rank >= 0 && batch > 0 && nheads > 0 && headdim > 0 && dstate > 0,
               "fold_boundary_states: bad arguments");
  TV_CHECK_ARG(rank == 0 || (states != nullptr && logdecay != nullptr), "fold_boundary_states: null summaries");
  const int64_t PN = (int64_t)headdim * dstate, BH = (int64_t)batch * nheads;
  const int64_t srs = states_rank_stride > 0 ? states_rank_stride : BH * PN;
  const int64_t lrs = logdecay_rank_stride > 0 ? logdecay_rank_stride : BH;
  TV_CHECK_ARG(PN % 4 == 0 && srs % 4 == 0 && ((uintptr_t)states % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                   (initial == nullptr || (uintptr_t)initial % 16 == 0),
               "fold_boundary_states: headdim*dstate and the rank stride must be multiples of 4, pointers 16-byte aligned");
  dim3 grid((unsigned)ceil_div(PN / 4, 256), (unsigned)BH);
  fold_boundary_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(states, logdecay, initial, out, rank, srs, lrs,
                                                              (int)(PN / 4));
  TV_LAUNCH_OK();
  return TV_OK;
}

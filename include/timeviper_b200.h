/*
 * timeviper_b200 -- C ABI of the B200-native Mamba-2 mixer prefill path.
 *
 * The reference (xiaomi-research/timeviper) has no native code: its mixer calls module-level Python names
 * bound to third-party wheels (timeviper/model/llm/llm_repo/nano/modeling_nano.py:60-97).  Each entry
 * point below is what a binding for one of those names calls; timeviper_b200/ops.py is that binding
 * (ctypes) and INTEGRATION.md shows how the reference rebinds its globals to it.
 *
 * Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the library never allocates or frees device memory, never synchronises, never changes device:
 *     outputs and workspaces are caller-allocated, work is enqueued on `stream` (a cudaStream_t);
 *   - strides are in ELEMENTS of the tensor's dtype;
 *   - return 0 on success, a negative tv_status otherwise; tv_last_error() gives the message
 *     (thread-local).  There is no CPU fallback.
 */
#ifndef TIMEVIPER_B200_H
#define TIMEVIPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TV_ABI_VERSION 4

typedef enum { TV_F32 = 0, TV_BF16 = 1 } tv_dtype;

typedef enum {
  TV_OK = 0,
  TV_ERR_INVALID = -1,     /* bad shape / stride / alignment / dtype (Python raises ValueError)      */
  TV_ERR_UNSUPPORTED = -2, /* valid request this build has no kernel for (NotImplementedError)        */
  TV_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed (RuntimeError)                        */
  TV_ERR_WORKSPACE = -4    /* workspace too small                                                      */
} tv_status;

int tv_abi_version(void);
const char* tv_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * causal_conv1d_fn  (causal_conv1d 1.5.2; call site modeling_nano.py:619-624; positional form
 * visualize/nano/my_ssd_combined.py:1606-1614).
 *   y[b,c,t] = act(bias[c] + sum_k w[c,k] * x[b,c,t-(K-1)+k]),  x[t<0] = initial_states or 0.
 * x and out are CHANNEL-LAST: element (b,c,t) at  b*batch_stride + t*seq_stride + c.
 * initial_states / final_states: (batch, dim, width-1), contiguous, same dtype as x; may be NULL.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  const void* weight;          /* (dim, width) contiguous, dtype of x                           */
  const void* bias;            /* (dim,) or NULL                                                */
  const void* initial_states;  /* (batch, dim, width-1) or NULL                                 */
  void* out;
  void* final_states;          /* (batch, dim, width-1) or NULL                                 */
  int32_t batch, dim, seqlen, width;
  int64_t x_batch_stride, x_seq_stride;
  int64_t out_batch_stride, out_seq_stride;
  int32_t silu;                /* 1: SiLU ("silu"/"swish"), 0: identity                          */
  int32_t dtype;               /* tv_dtype of x / weight / bias / out                            */
} tv_conv1d_params;

int tv_causal_conv1d_fwd(const tv_conv1d_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------
 * rmsnorm_fn (mamba_ssm.ops.triton.layernorm_gated; call site modeling_nano.py:372-380).
 *   norm_before_gate=0:  u = x*silu(z);  out = u * rsqrt(mean_group(u^2)+eps) * w (+bias)
 *   norm_before_gate=1:  out = (x * rsqrt(mean_group(x^2)+eps) * w (+bias)) * silu(z)
 * x, z, out: (rows, d) with unit inner stride and the given row strides; z may be NULL.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  const void* z;
  const void* weight;          /* (d,)                                                           */
  const void* bias;            /* (d,) or NULL                                                   */
  void* out;
  int64_t rows;
  int32_t d, group_size;
  int64_t x_row_stride, z_row_stride, out_row_stride;
  float eps;
  int32_t norm_before_gate;
  int32_t dtype;               /* tv_dtype of x / z / weight / bias / out                        */
} tv_rmsnorm_params;

int tv_gated_rmsnorm_fwd(const tv_rmsnorm_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Residual add + NemotronHRMSNorm of the hybrid layer loop: the end of one block,
 * `residual + hidden_states` (modeling_nano.py:965), and the pre-norm of the next (:888-904, call :941) in one pass:
 *   s = dtype(x + residual)   (written to sum_out when given);   out = dtype(w * (s * rsqrt(mean_d(s^2) + eps)))
 * residual and sum_out may be NULL (plain RMSNorm of x).  x, residual, sum_out, out: (rows, d), unit inner stride.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  const void* residual;        /* (rows, d) or NULL                                              */
  const void* weight;          /* (d,)                                                           */
  void* sum_out;               /* (rows, d) or NULL: x + residual, rounded to dtype              */
  void* out;
  int64_t rows;
  int32_t d;
  int32_t dtype;               /* tv_dtype of every tensor                                       */
  int64_t x_row_stride, res_row_stride, sum_row_stride, out_row_stride;
  float eps;
  int32_t reserved;
} tv_add_rmsnorm_params;

int tv_add_rmsnorm_fwd(const tv_add_rmsnorm_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------
 * mamba_chunk_scan_combined forward (mamba_ssm 2.2.5 ssd_combined; signature
 * visualize/nano/my_ssd_combined.py:1270-1306, stages :743-843; call site modeling_nano.py:639-653).
 *   x (b,L,H,P)  dt (b,L,H)  A (H,) f32  B,C (b,L,G,N)  D (H,) or (H,P) f32  z (b,L,H,P)
 *   dt_bias (H,) f32   initial_states (b,H,P,N) f32   ->   out (b,L,H,P) x-dtype, final_states (b,H,P,N) f32
 * Head h reads group h / (H/G).  x/z/B/C have unit stride on their last dim; out is contiguous.
 * `mode`: TV_SSD_FULL computes out (+final_states); TV_SSD_STATE_ONLY computes only final_states and
 * chunk_logdecay_sum (the shard summary of the sequence-sharded path) and touches neither out, C, D nor z.
 * TV_SSD_DT_ONLY runs only the dt activation + per-chunk cumsum into `workspace` (it reads dt, A, dt_bias and no
 * other tensor -- x/B are only inspected for the family choice), so that a caller can overlap it with the conv
 * that produces x/B/C and pass reuse_dt_cumsum = 1 to the calls that follow.  The library remembers which dt tensor,
 * dims, strides and limits each workspace was filled from and returns TV_ERR_INVALID for a reuse that does not match.
 * ------------------------------------------------------------------------------------------- */
typedef enum { TV_SSD_FULL = 0, TV_SSD_STATE_ONLY = 1, TV_SSD_DT_ONLY = 2 } tv_ssd_mode;

typedef struct {
  const void* x; const void* dt; const float* A; const void* B; const void* C;
  const float* D;              /* NULL, (H,) or (H,P) -- see d_has_hdim                          */
  const void* z;               /* NULL or (b,L,H,P)                                              */
  const float* dt_bias;        /* NULL or (H,)                                                   */
  const float* initial_states; /* NULL or (b,H,P,N) contiguous                                   */
  void* out;                   /* (b,L,H,P) contiguous                                           */
  float* final_states;         /* NULL or (b,H,P,N) contiguous                                   */
  float* logdecay_sum;         /* NULL or (b,H): sum over the sequence of dt*A (log of the decay) */
  int32_t batch, seqlen, nheads, headdim, ngroups, dstate, chunk_size;
  int64_t x_batch_stride, x_seq_stride, x_head_stride;
  int64_t dt_batch_stride, dt_seq_stride, dt_head_stride;
  int64_t b_batch_stride, b_seq_stride, b_group_stride;
  int64_t c_batch_stride, c_seq_stride, c_group_stride;
  int64_t z_batch_stride, z_seq_stride, z_head_stride;
  int32_t d_has_hdim;          /* 1 if D is (H,P)                                                */
  int32_t dt_softplus;
  float dt_min, dt_max;        /* dt_limit; (0, +inf) is a no-op clamp                           */
  int32_t dtype;               /* tv_dtype of x / dt / B / C / z / out                           */
  int32_t mode;                /* tv_ssd_mode                                                    */
  int32_t force_simt;          /* 1: use the fp32 CUDA-core kernels even where a tcgen05 kernel exists */
  int32_t reuse_dt_cumsum;     /* 1: `workspace` still holds dt/cumsum of the previous call with the same dt, A, dt_bias,
                                  dims and kernel family (pass 2 of the sharded path right after pass 1): skip stage (i) */
} tv_ssd_params;

/* Bytes of caller-allocated scratch tv_ssd_chunk_scan_fwd needs for these dims (0 is possible). */
size_t tv_ssd_workspace_bytes(const tv_ssd_params* p);
int tv_ssd_chunk_scan_fwd(const tv_ssd_params* p, void* workspace, size_t workspace_bytes, void* stream);
/* Which kernel family tv_ssd_chunk_scan_fwd would run: 0 = fp32 CUDA-core, 1 = tcgen05/TMEM/TMA. */
int tv_ssd_kernel_family(const tv_ssd_params* p);

/* ---------------------------------------------------------------------------------------------
 * Sequence-sharded prefill: fold the gathered per-shard summaries into the state entering shard `rank`.
 *   S_in(0) = initial (or 0);  S_in(r+1) = exp(logdecay[r]) * S_in(r) + states[r]
 * states: (world, b, H, P, N) f32, logdecay: (world, b, H) f32 (as gathered), out: (b, H, P, N) f32.
 * *_rank_stride: elements between consecutive ranks' summaries (0 = densely packed), so that one flat
 * all-gather buffer [S_r | logdecay_r] per rank can be folded in place.
 * New work (the reference has no sequence parallelism); the hook it plugs into is
 * the `initial_states=` argument of mamba_chunk_scan_combined (my_ssd_combined.py:1280,1300).
 * ------------------------------------------------------------------------------------------- */
int tv_ssd_fold_boundary_states(const float* states, const float* logdecay, const float* initial,
                                float* out, int32_t rank, int32_t batch, int32_t nheads,
                                int32_t headdim, int32_t dstate, int64_t states_rank_stride,
                                int64_t logdecay_rank_stride, void* stream);

/* The same fold with every rank's summary read through ITS OWN pointer: state_ptrs[r] / logdecay_ptrs[r] (HOST arrays of
 * `rank` device pointers, r < rank) may be peer memory of other GPUs of the node (CUDA IPC / symmetric memory mapped
 * into this process), so the boundary-state exchange and the fold are one kernel over NVLink.  Summaries whose weight
 * has underflowed fp32 are not fetched.  New work; at most 16 ranks. */
int tv_ssd_fold_boundary_states_p2p(const void* const* state_ptrs, const void* const* logdecay_ptrs, const float* initial,
                                    float* out, int32_t rank, int32_t batch, int32_t nheads, int32_t headdim,
                                    int32_t dstate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Single-token decode step: the two operators the reference's cached branch calls,
 * modeling_nano.py:495-501 and :528-539, consuming the states the prefill path leaves in the cache.
 *
 * causal_conv1d_update (causal_conv1d 1.5.2): conv_state (batch, dim, state_len >= width) is shifted left by one column
 * in place, x (batch, dim) becomes its last column, out[b,c] = act(bias[c] + sum_k w[c,k] * state[b,c,state_len-width+k]).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;               /* (batch, dim), unit channel stride                              */
  void* conv_state;            /* (batch, dim, state_len), unit stride on the last dim, in place */
  const void* weight;          /* (dim, width) contiguous                                        */
  const void* bias;            /* (dim,) or NULL                                                 */
  void* out;                   /* (batch, dim), unit channel stride                              */
  int32_t batch, dim, width, state_len;
  int64_t x_batch_stride, out_batch_stride, state_batch_stride, state_dim_stride;
  int32_t silu;
  int32_t dtype;               /* tv_dtype of x / conv_state / weight / bias / out               */
} tv_conv1d_update_params;

int tv_causal_conv1d_update(const tv_conv1d_update_params* p, void* stream);

/* selective_state_update (mamba_ssm.ops.triton.selective_state_update), one token:
 *   dt' = clamp(softplus(dt + dt_bias));  state = state * exp(dt' * A) + dt' * B[g] * x;  out = sum_n state * C[g] + D * x
 *   [out *= silu(z)],  g = h / (H/G).  state (batch, H, P, N) contiguous, updated in place.
 * x, dt, z: (batch, H, P); A: (H, P, N); D, dt_bias: (H, P) -- all addressed through element strides, so the expanded
 * (stride-0) views the reference builds at modeling_nano.py:515-522 are passed as they are.  A, D, dt_bias are f32. */
typedef struct {
  void* state;
  const void* x; const void* dt; const float* A; const void* B; const void* C;
  const float* D; const void* z; const float* dt_bias;
  void* out;                   /* (batch, H, P) contiguous                                       */
  int32_t batch, nheads, headdim, ngroups, dstate;
  int64_t x_batch_stride, x_head_stride, x_dim_stride;
  int64_t dt_batch_stride, dt_head_stride, dt_dim_stride;
  int64_t a_head_stride, a_dim_stride, a_state_stride;
  int64_t b_batch_stride, b_group_stride;            /* unit stride over the state dim */
  int64_t c_batch_stride, c_group_stride;
  int64_t d_head_stride, d_dim_stride;
  int64_t z_batch_stride, z_head_stride, z_dim_stride;
  int64_t bias_head_stride, bias_dim_stride;
  int32_t dt_softplus;
  float dt_min, dt_max;        /* clamp after softplus; (0, +inf) = none                          */
  int32_t dtype;               /* tv_dtype of x / dt / B / C / z / out                            */
  int32_t state_dtype;         /* tv_dtype of state                                               */
} tv_ssu_params;

int tv_selective_state_update(const tv_ssu_params* p, void* stream);

/* Debug hook (profiling only): device buffer of nchunks*16 int64 that CTA (0,0) of the fused SSD kernel fills
 * with clock64() stamps of its pipeline events; NULL (default) disables it. */
void tv_debug_set_trace(void* device_buffer);
/* Profiling builds only (python -m timeviper_b200.build --trace): bitmask of per-role work the fused SSD kernel skips,
 * to find the critical path by ablation (results are then wrong by construction).  No effect in normal builds. */
void tv_debug_set_ablate(int mask);
/* Number of kernels this library has enqueued so far in this process (all entry points, all streams): what bench.py
 * reports as gpu_launches. */
unsigned long long tv_debug_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TIMEVIPER_B200_H */

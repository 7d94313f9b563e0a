/*
 * smmd.h -- C ABI of libsmmd.so: the B200-native (sm_100a) replacement for the data-parallel hot path
 * of playHing/Scaled-MMD-GAN: the pairwise-kernel MMD^2 loss (+ feature gradients) and the
 * cubic-polynomial KID score.
 *
 * The reference has no FFI / plugin registry: its boundary is a set of Python call signatures
 * (SURVEY.md section 8b).  Every entry point below names the reference interface it replaces
 * (file:line relative to the reference repo root).  The Python drop-in with the reference's own
 * names/kwargs lives in scaled-mmd-gan_b200/smmd/{mmd,compute_scores}.py and binds this header via
 * ctypes.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers owned by the caller; the library never allocates or
 *     frees user memory and only uses the caller-supplied workspace.  Matrices are row-major.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises
 *     the host, and is re-entrant as long as workspaces are distinct.
 *   - Returns 0 (SMMD_OK) or a negative smmd_status; never throws, never aborts.  Shape / dtype /
 *     architecture problems are reported before anything is launched.
 *   - There is no CPU fallback: on a device that is not sm_100 the calls return SMMD_EARCH.
 */
#ifndef SMMD_H_
#define SMMD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SMMD_API __attribute__((visibility("default")))
#else
#define SMMD_API
#endif

#define SMMD_VERSION 100 /* 0.1.0 */
#define SMMD_MAX_PARAMS 8
#define SMMD_NUM_SCALARS 16

typedef enum smmd_status {
  SMMD_OK = 0,
  SMMD_EINVAL = -1,       /* null pointer / bad enum / bad parameter value                      */
  SMMD_ESHAPE = -2,       /* m, n, d, ld* or subset shape not acceptable                        */
  SMMD_EDTYPE = -3,       /* dtype not supported by the selected precision path                 */
  SMMD_EARCH = -4,        /* current device is not sm_100 (B200)                                */
  SMMD_EWORKSPACE = -5,   /* workspace NULL, misaligned or smaller than *_workspace_bytes()     */
  SMMD_ECUDA = -6,        /* a CUDA runtime/driver call failed (see smmd_last_cuda_error())     */
  SMMD_EUNSUPPORTED = -7  /* combination not implemented on the requested precision path        */
} smmd_status;

typedef enum smmd_dtype { SMMD_F32 = 0, SMMD_BF16 = 1 } smmd_dtype;

/* Kernel families = the `_<name>_kernel` zoo of gan/core/mmd.py:18-188 (selected by name string in
 * gan/core/model.py:314 and gan/core/smmd.py:11).  The mix_rq_*dot wrappers (mmd.py:119-136) are
 * SMMD_K_MIX_RQ with add_dot set; SMMD_K_RBF is SMMD_K_MIX_RBF with one sigma (mmd.py:55-82). */
typedef enum smmd_kernel_id {
  SMMD_K_DISTANCE = 0,      /* mmd.py:18-37   */
  SMMD_K_TANH_DISTANCE = 1, /* mmd.py:40-41   */
  SMMD_K_DOT = 2,           /* mmd.py:44-52   */
  SMMD_K_RBF = 3,           /* mmd.py:55-82   params[0]=sigma, wts[0]=wt          */
  SMMD_K_MIX_RBF = 4,       /* mmd.py:85-116  params=sigmas, wts                  */
  SMMD_K_MIX_RQ = 5,        /* mmd.py:143-188 params=alphas, wts, add_dot         */
  SMMD_K_TANH_MIX_RQ = 6,   /* mmd.py:139-140 */
  SMMD_K_POLY = 7           /* compute_scores.py:232-244 / sklearn polynomial_kernel:
                               params[0]=gamma (<=0 -> 1/d), params[1]=coef0, degree in `degree` */
} smmd_kernel_id;

/* Arithmetic path.  FP32: exact fp32 SIMT kernels (tolerance tier rel 1e-5 vs the oracle).
 * BF16: tcgen05 tensor cores, bf16 operands / fp32 accumulate (tier rel 1e-3).
 * BF16X3: tcgen05 with a 3-term split-bf16 Gram (hi*hi + lo*hi + hi*lo), forward-only paths
 * (KID, value-only MMD^2); ~fp32-accurate Gram at 3x the tensor work.
 * FP16: the BF16 gradient paths with IEEE-half operands (features and the weight matrix W; W is carried times a power of two
 * derived from its bound): 11 significant bits instead of 8 -- gradients within 3e-4 of max|g| instead of 1e-3..2e-3 (bf16),
 * same speed -- for features inside half's range (6e-5 <= |z| <= 6e4; a larger |z| becomes inf and raises
 * SMMD_S_NONFINITE).  Explicit request only (AUTO never picks it); forward-only calls return SMMD_EUNSUPPORTED.
 * AUTO: FP32 for small problems (m + n < 1024 or d < 32) and for combinations the tensor-core kernels do not
 * cover (dot kernel), BF16 otherwise.  The exact gradient kernel keeps a feature row in registers and supports
 * d <= 2048: with gradients and a wider d, AUTO takes BF16 whatever the size, and an explicit FP32 request returns
 * SMMD_EUNSUPPORTED.  An explicit BF16 / BF16X3 request is honoured or refused, never silently downgraded. */
typedef enum smmd_precision {
  SMMD_PREC_FP32 = 0,
  SMMD_PREC_BF16 = 1,
  SMMD_PREC_BF16X3 = 2,
  SMMD_PREC_AUTO = 3,
  SMMD_PREC_FP16 = 4
} smmd_precision;

/* One MMD^2 problem: X = generated/fake features [m,d], Y = real features [n,d] -- the reference's
 * argument order kernel(G, images) (model.py:315). */
typedef struct smmd_problem {
  int64_t m, n, d;                 /* rows of X, rows of Y, feature dim (all >= 1; m,n >= 2 if !biased) */
  int64_t ldx, ldy;                /* row strides of X / Y in elements (>= d)                  */
  int32_t dtype;                   /* smmd_dtype of X and Y                                    */
  int32_t kernel_id;               /* smmd_kernel_id                                           */
  int32_t nparams;                 /* number of sigmas / alphas (<= SMMD_MAX_PARAMS)           */
  float params[SMMD_MAX_PARAMS];   /* sigmas (rbf), alphas (rq), {gamma, coef0} (poly)         */
  float wts[SMMD_MAX_PARAMS];      /* mixture weights                                          */
  float add_dot;                   /* mix_rq add_dot (mmd.py:143,170-171,183-185)              */
  int32_t degree;                  /* polynomial degree (SMMD_K_POLY)                          */
  int32_t biased;                  /* mmd2(K, biased) (mmd.py:194)                             */
  int32_t precision;               /* smmd_precision                                           */
  /* Row sharding for one process per GPU: X and Y passed to the call are the FULL (all-gathered)
   * matrices on every rank; rank r owns X rows [m*r/world, m*(r+1)/world) and Y rows
   * [n*r/world, n*(r+1)/world), computes only those rows of the Gram against all columns and
   * returns partial sums (+ complete gradients for its rows).  world=1 -> single GPU. */
  int32_t rank, world;
} smmd_problem;

/* scalars[] layout (device, double[SMMD_NUM_SCALARS]) written by smmd_mmd2_fwd_bwd: */
enum {
  SMMD_S_MMD2 = 0,     /* final MMD^2 (only meaningful when world == 1)                       */
  SMMD_S_SUM_XX = 1,   /* sum of K_XX over owned rows, diagonal excluded                      */
  SMMD_S_SUM_YY = 2,   /* sum of K_YY over owned rows, diagonal excluded                      */
  SMMD_S_SUM_XY = 3,   /* sum of K_XY over owned X rows x all Y columns                       */
  SMMD_S_SUM_YX = 4,   /* sum of K_YX over owned Y rows x all X columns                       */
  SMMD_S_DIAG_X = 5,   /* sum_i K_XX[i,i] over owned rows (analytic diagonal)                 */
  SMMD_S_DIAG_Y = 6,   /* sum_i K_YY[i,i] over owned rows                                     */
  SMMD_S_NONFINITE = 7,/* != 0 when a NaN/Inf reached the sums (model.py:537 asserts on it)   */
  SMMD_S_VAR = 8,      /* variance estimate   (smmd_mmd2_and_ratio only)                      */
  SMMD_S_RATIO = 9     /* mmd2/sqrt(max(var,min_var_est)) (smmd_mmd2_and_ratio only)          */
};

SMMD_API int smmd_version(void);
SMMD_API const char* smmd_strerror(int status);
/* Text of the last CUDA error seen by this thread inside the library ("" if none). */
SMMD_API const char* smmd_last_cuda_error(void);
/* 1 if the current device can run the library (compute capability 10.x), else 0. */
SMMD_API int smmd_device_supported(void);

/* Replaces: mmd2(kernel(X, Y, ...), biased) + tf.gradients of it w.r.t. X and Y
 *   gan/core/mmd.py:18-220 (kernels + _mmd2), called from gan/core/model.py:313-319 and
 *   gan/core/smmd.py:10-19; autodiff at gan/core/model.py:446,452.
 * scalars: device double[SMMD_NUM_SCALARS].  dX [owned_m, d] / dY [owned_n, d] fp32 row-major
 * (ld = d) receive dMMD2/dX, dMMD2/dY for the owned rows; both NULL = forward only.
 * smmd_mmd2_workspace_bytes is exact for the problem AS GIVEN, including its rank / world row shard (the
 * work split of a shard is not bounded by the unsharded one): query it with the struct you will pass. */
SMMD_API size_t smmd_mmd2_workspace_bytes(const smmd_problem* p, int want_grad);
SMMD_API int smmd_mmd2_fwd_bwd(const smmd_problem* p, const void* X, const void* Y, double* scalars,
                      float* dX, float* dY, void* workspace, size_t workspace_bytes, void* stream);

/* One-process-per-GPU variant of the above (new capability, no reference counterpart: the reference never
 * computes the loss across towers, gan/core/model.py:186-218).  `gathered` is the all_gather of every rank's
 * [X_local ; Y_local] block: `world` blocks of (m/world + n/world) rows with row pitch `ld`, element type
 * p->dtype (all-gathering bf16 halves the NVLink bytes).  X_owned / Y_owned (nullable, fp32, pitch ld_owned) are
 * this rank's own rows at full precision; they are used for the r_i z_i term of the gradient. */
SMMD_API int smmd_mmd2_fwd_bwd_gathered(const smmd_problem* p, const void* gathered, int64_t ld, const float* X_owned,
                                        const float* Y_owned, int64_t ld_owned, double* scalars, float* dX,
                                        float* dY, void* workspace, size_t workspace_bytes, void* stream);

/* Multi-GPU: after all-reducing scalars[1..7] over ranks, turn the summed partials into MMD^2
 * (same arithmetic as the single-GPU finalisation).  sums/out are device pointers; out[0] = MMD^2. */
SMMD_API int smmd_mmd2_combine(const smmd_problem* p, const double* sums, double* out, void* stream);

/* Peer-memory variant of the sharded loss (one process per GPU, all GPUs of one NVLink / NVSwitch domain): the exchange
 * steps of the path run INSIDE the library's kernels over peer-mapped memory instead of as NCCL collectives around them.
 *   - every rank owns an exchange buffer of smmd_peer_buffer_bytes() bytes, zero-filled once, mapped into every other
 *     rank's address space by the caller (CUDA IPC / VMM / torch symmetric memory); peers->base[r] is rank r's buffer as
 *     seen from THIS process (base[rank] is the own buffer);
 *   - a call publishes the local rows into the own buffer and raises a flag in every peer's buffer (release, system
 *     scope); the operand-preparation kernel pulls each peer's rows over NVLink as soon as that peer's flag is up
 *     (the all_gather, fused into the copy / convert / norm pass that had to read those rows anyway); the partial sums
 *     are written into every peer's buffer the same way and each rank combines them in rank order (the all_reduce +
 *     smmd_mmd2_combine, fused into one kernel).  Latency-bound shapes (<= 1024 global rows, d <= 64, fp32 tier) run
 *     publish + pull + loss + gradients + sum exchange + combine as ONE launch.
 * `step` is the collective's sequence number: 1 for the first call after the buffers were zeroed, +1 per call, the same
 * on every rank; step = 0 lets the kernels count on the device (last completed step of this rank + 1, kept in the own
 * buffer) -- the call sequence then contains no host-side state and can be captured into a CUDA graph and replayed (calls
 * must be stream-ordered in that mode) (two slots and two sets of flags alternate: a rank may be one call ahead of its slowest peer, and two
 * consecutive calls may be in flight on two streams; calls of the same parity must be stream-ordered).  p describes the GLOBAL problem (m, n = total rows, p->rank / p->world = this shard; m and n multiples of
 * world); X_local / Y_local are this rank's fp32 rows (pitch ld_local); scalars receives the COMBINED result (identical on
 * every rank), dX / dY the gradients of the local rows.  Every wait is bounded (a missing peer traps after 4 s).
 * New capability without a reference counterpart (the reference's towers never exchange features,
 * gan/core/model.py:186-218); parity target = the single-device result on the concatenated batch. */
#define SMMD_MAX_PEERS 16
typedef struct smmd_peer_table {
  int32_t world, rank;
  void* base[SMMD_MAX_PEERS];
} smmd_peer_table;
/* Optional: a cudaEvent_t (NULL to clear; per calling thread) that the following smmd_mmd2_fwd_bwd_peers calls record on
 * their stream as soon as the peers' rows have been pulled (after the operand preparation; at the end of the call for
 * the one-launch path).  Lets a pipelined caller order host<->device copies behind the NVLink phase of a step: on the
 * measured 8-GPU boxes a PCIe device->host copy running next to the pull stalls the NVLink traffic for its whole
 * duration (0.1 ms -> 1.4 ms per step), while the same copy next to the kernels costs nothing. */
SMMD_API int smmd_peer_set_pull_event(void* cuda_event);
SMMD_API size_t smmd_peer_buffer_bytes(int64_t rows_local, int64_t d);
SMMD_API int smmd_mmd2_fwd_bwd_peers(const smmd_problem* p, const smmd_peer_table* peers, uint64_t step,
                                     const float* X_local, const float* Y_local, int64_t ld_local, double* scalars,
                                     float* dX, float* dY, void* workspace, size_t workspace_bytes, void* stream);

/* Host-side helper (no GPU work): the subset draw of polynomial_mmd_averages, gan/compute_scores.py:219-222 --
 *   for each subset: np.random.choice(len(codes_g), subset_size, replace=False), then the same for codes_r --
 * reproduced bit for bit from numpy's legacy global RNG state: choice(n, m, replace=False) is permutation(n)[:m], i.e. a
 * Fisher-Yates shuffle of arange(n) from the top with one rejection-sampled `random_interval(i)` (masked 32-bit MT19937
 * outputs) per position.  key[624] / pos are np.random.get_state()[1:3]; they are advanced in place, so that
 * np.random.set_state afterwards leaves the global stream exactly where the reference's loop would have left it.
 * idx_g / idx_r: host int32 [n_subsets][subset_size].  The stream is walked twice: a sequential scan that only counts
 * what each shuffle consumes (16 outputs classified at a time) and remembers the generator state it starts from, then the
 * shuffles themselves, branch-free, on up to 16 threads (SMMD_DRAW_THREADS overrides; 1 = one thread, no scan).  The
 * numpy loop takes 107-136 ms at 100 x 1000 of 50k on the bench box's host (most of a host-codes KID call), this 11.8 ms. */
SMMD_API int smmd_draw_subsets_mt19937(uint32_t* key, int32_t* pos, int64_t len_g, int64_t len_r, int32_t n_subsets,
                                       int32_t subset_size, int32_t* idx_g, int32_t* idx_r);

/* Replaces: mmd2_and_ratio(K, biased, min_var_est) -- gan/core/mmd.py:223-293 (+ ops.sq_sum / ops.dot,
 * gan/core/ops.py:209-225).  Requires m == n (mmd.py:237).  Reproduces the reference's unbiased
 * branch that keeps the diagonal (mmd.py:273-276).  scalars[MMD2, VAR, RATIO] are written. */
SMMD_API int smmd_mmd2_and_ratio(const smmd_problem* p, const void* X, const void* Y, double min_var_est,
                        double* scalars, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces: kernel(X, Y, K_XY_only=True) -- e.g. gan/core/mmd.py:31-32,72-73,106-107,172-173, the
 * witness term of the gradient penalty (gan/core/model.py:336).  K [m, n] fp32 row-major, ld = ldk. */
SMMD_API int smmd_kernel_xy(const smmd_problem* p, const void* X, const void* Y, float* K, int64_t ldk,
                   void* stream);
/* VJP of the above: given dK [m,n] returns dX [m,d], dY [n,d] (fp32, ld = d). */
SMMD_API int smmd_kernel_xy_bwd(const smmd_problem* p, const void* X, const void* Y, const float* dK,
                       int64_t lddk, float* dX, float* dY, void* stream);

/* Second-order VJP of the witness block: the backward of smmd_kernel_xy_bwd, needed because the gradient
 * penalty differentiates |d witness / d x_hat| again w.r.t. the critic (gan/core/model.py:336-341,
 * `tf.gradients(witness, [x_hat_data])` inside the loss that `tf.gradients(d_loss, d_vars)` then differentiates).
 * Inputs: X [m,d], Y [n,d], dK [m,n] as in smmd_kernel_xy_bwd; VX [m,d] / VY [n,d] fp32 contiguous cotangents of
 * that call's (dX, dY) (either may be NULL = zero).  Outputs (fp32, contiguous): ddK [m,n] = dL/d(dK),
 * gX [m,d] = dL/dX, gY [n,d] = dL/dY.  Not available for the poly / tanh_* kernel ids (SMMD_EUNSUPPORTED): tanh
 * is applied to the features by the caller. */
SMMD_API int smmd_kernel_xy_bwd2(const smmd_problem* p, const void* X, const void* Y, const float* dK, int64_t lddk,
                        const float* VX, const float* VY, float* ddK, float* gX, float* gY, void* stream);

/* KID: polynomial_mmd_averages over subsets given by explicit row indices.
 *   Replaces gan/compute_scores.py:211-229 (subset loop + fancy-index gather), :232-244
 *   (polynomial_mmd = 3x sklearn polynomial_kernel) and :252-335 (_mmd2_and_variance), as called
 *   from gan/utils/scorer.py:103-109 and the CLI (compute_scores.py:489-493).
 * The reference draws the subsets from numpy's global RNG (compute_scores.py:221-222); the library
 * has no hidden RNG: the caller passes the indices (int32, [n_subsets, subset_size]). */
typedef enum smmd_mmd_est { SMMD_EST_UNBIASED = 0, SMMD_EST_BIASED = 1, SMMD_EST_USTAT = 2 } smmd_mmd_est;

typedef struct smmd_kid_problem {
  int64_t n_g, n_r, d;             /* rows of codes_g, rows of codes_r, feature dim           */
  int64_t ldg, ldr;                /* row strides (elements)                                  */
  int32_t dtype;                   /* smmd_dtype of the codes (F32)                           */
  int32_t n_subsets, subset_size;  /* S, m                                                    */
  int32_t degree;                  /* 3                                                       */
  float gamma;                     /* <= 0 -> 1/d (sklearn gamma=None)                        */
  float coef0;                     /* 1                                                       */
  int64_t var_at_m;                /* compute_scores.py:213,223 (min(len_g,len_r)); <=0 -> m  */
  int32_t mmd_est;                 /* smmd_mmd_est (compute_scores.py:290-300)                */
  int32_t ret_var;                 /* also compute the variance estimate (:305-333)           */
  int32_t precision;               /* smmd_precision (AUTO -> BF16X3)                         */
  int32_t first_subset, n_local;   /* subset shard of this rank: [first, first+n_local); n_local<=0 -> all */
} smmd_kid_problem;

SMMD_API size_t smmd_kid_workspace_bytes(const smmd_kid_problem* p);
/* mmd2_out / var_out: device double[n_subsets] (entries of the local shard are written);
 * var_out may be NULL when ret_var == 0. */
SMMD_API int smmd_kid_subsets(const smmd_kid_problem* p, const void* codes_g, const void* codes_r,
                     const int32_t* idx_g, const int32_t* idx_r, double* mmd2_out, double* var_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* 3-sample test of the scorer (gan/utils/scorer.py:119-163): the "Y related sums" that
 * np_diff_polynomial_mmd2_and_ratio_with_saving (gan/core/mmd.py:429-444) builds with _np_get_sums
 * (gan/core/mmd.py:515-539) from K_XY = (X Y^T / d + 1)^3 and K_YY, without materialising either block.
 * X, Y: device [m, d] (equal sizes, mmd.py:516); the problem struct is smmd_kid_problem with
 * n_subsets = 1 and n_g = n_r = subset_size = m (degree / gamma / coef0 / precision as for KID).
 * sums_out: device double[3 m + 2] =
 *   [0, m)    Kt_YY_sums    row sums of K_YY without the diagonal
 *   [m, 2m)   K_XY_sums_0   sum over x of K(x, y_j)       (K_XY.sum(axis=0))
 *   [2m, 3m)  K_XY_sums_1   sum over y of K(x_i, y)       (K_XY.sum(axis=1))
 *   [3m]      Kt_YY_2_sum   sum of K_YY^2 without the diagonal
 *   [3m + 1]  K_XY_2_sum    sum of K_XY^2
 * The estimator arithmetic on these vectors (mmd.py:447-512) is host-side in smmd/mmd.py. */
SMMD_API size_t smmd_poly_sums_workspace_bytes(const smmd_kid_problem* p);
SMMD_API int smmd_poly_sums(const smmd_kid_problem* p, const void* X, const void* Y, double* sums_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* The estimator arithmetic alone, on row statistics the CALLER has reduced from dense kernel blocks -- the dense-block
 * compatibility entry points `_mmd2_and_variance(K_XX, K_XY, K_YY, ...)` of gan/compute_scores.py:252-335 and
 * gan/core/mmd.py:236-293 (callers that already hold K matrices).  stats: device double[2 m][6], X rows then Y rows, per row
 *   [0] sum_j k(z_i, z_j) over its own set without the diagonal   [1] the same over the other set
 *   [2] sum of squares of [0]'s terms                              [3] sum of squares of [1]'s terms
 *   [4] k(z_i, z_i)                                                [5] k(x_i, y_i) (X rows; u-statistic)
 * smmd_kid_from_row_stats writes mmd2_out[0] (and var_out[0] when ret_var); smmd_ratio_from_row_stats writes
 * scalars[SMMD_S_MMD2 / _VAR / _RATIO] with the reference's quirks (const-diagonal kernels subtract the constant,
 * the unbiased branch keeps the diagonal). */
SMMD_API int smmd_kid_from_row_stats(const double* stats, int64_t m, int mmd_est, int ret_var, int64_t var_at_m,
                                     double* mmd2_out, double* var_out, void* stream);
SMMD_API int smmd_ratio_from_row_stats(const double* stats, int64_t m, int biased, int has_const_diagonal,
                                       double const_diagonal, double min_var_est, double* scalars, void* stream);

/* Roofline instrumentation (bench.py): when enabled on this thread, the library records a CUDA-event
 * pair on the launching stream around the DOMINANT kernel of each call (the fused tcgen05 Gram kernel /
 * the K-streaming KID kernel / the SIMT row kernel).  smmd_profile_last_ms() synchronises on the stop
 * event and returns that launch's device time in milliseconds (< 0 if nothing was recorded). */
SMMD_API void smmd_profile_enable(int on);
SMMD_API float smmd_profile_last_ms(void);
/* Two-kernel paths (symmetric MMD^2: W generation, then O = W Z) also record the boundary between the two launches:
 * device time of the first and of the second kernel of the last recorded call.  SMMD_EINVAL if that call was a
 * single-kernel path. */
SMMD_API int smmd_profile_last_split_ms(float* first_ms, float* second_ms);

/* Path-selection options (process-wide; set them before concurrent use).  Values are validated and clamped to what
 * the kernels support.  Names: "sym" (0/1: symmetric two-pass path), "sym_min_rows" (stacked rows from which it is
 * used; 0 = measured default), "symf" / "symf_min_rows" (fused variant for d <= 256), "sym_max_w_mb", "fused_pair", "fused_lockstep", "fused_ksplit", "wz_pair",
 * "wz_min_d" (clamped to [0, 256]), "wz_panel_mb", "disable_small".  Tests use them to run a given code path on a
 * small shape; there is no reference counterpart.  Returns SMMD_EINVAL for an unknown name. */
SMMD_API int smmd_set_option(const char* name, long long value);

/* Introspection for tests/bench: number of kernels the last call on this thread launched, and the
 * name of the code path it took ("simt_fp32", "tc_bf16_fused", "tc_bf16_fwd", ...). */
SMMD_API int smmd_last_launch_count(void);
SMMD_API const char* smmd_last_path(void);

#ifdef __cplusplus
}
#endif
#endif /* SMMD_H_ */

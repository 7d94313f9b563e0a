// smmd_internal.h -- structures shared between the C-ABI dispatcher and the CUDA translation units.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/smmd.h"
#include "smmd_peer.cuh"

namespace smmd {

enum Family : int { FAM_DISTANCE = 0, FAM_DOT = 1, FAM_RBF = 2, FAM_RQ = 3, FAM_POLY = 4 };

constexpr float kEps = 1.0e-5f;  // `_eps`, gan/core/mmd.py:6

// Pair-kernel description passed by value to kernels.
struct KernelFn {
  int family;
  int np;
  float p0[SMMD_MAX_PARAMS];  // rbf: gamma_k = 1/(2 sigma_k^2)   rq: 1/(2 alpha_k)
  float p1[SMMD_MAX_PARAMS];  // rbf: -gamma_k * log2(e)          rq: alpha_k
  float w[SMMD_MAX_PARAMS];   // mixture weights
  float add_dot;              // rq only
  float poly_gamma, poly_coef0;
  int degree;
  int tanh_features;          // tanh applied to the features in the prep pass
  int true_distance;          // distance kernel: keep the sqrt(|x|^2+eps) terms (needed outside mmd2)
  float const_diag;           // sum of weights (rbf / rq), reference's 4th tuple element
  int has_const_diag;         // 0 for distance / dot / poly ("False" in the reference)
};

// Geometry of one stacked problem Z = [X ; Y].
struct Geometry {
  int64_t m, n, d;
  int64_t x0, x1, y0, y1;  // owned row ranges (row shard); world==1 -> [0,m), [0,n)
  int biased;
};

// Where the feature rows come from.  Plain: X [m][ldx], Y [n][ldy].  Gathered (one process per GPU): both
// pointers address the all_gather of every rank's [X_local ; Y_local] block, i.e. `world` blocks of
// (blk_x + blk_y) rows; X row i then lives at row (i / blk_x) * (blk_x + blk_y) + i % blk_x, Y row j at
// (j / blk_y) * (blk_x + blk_y) + blk_x + j % blk_y.  Xo / Yo: optional fp32 copies of the OWNED rows (row 0 =
// first owned row) used where full input precision matters (the r_i z_i term of the gradient).
struct SrcLayout {
  const void* X;
  const void* Y;
  int dtype;
  int64_t ldx, ldy;
  int64_t blk_x, blk_y;   // 0 = plain layout
  const float* Xo;
  const float* Yo;
  int64_t ldo;
  const PeerSrc* peers;   // non-null: the blocks are the peers' published rows (smmd_peer.cuh), pulled over NVLink
};
__host__ __device__ inline int64_t src_row(int64_t i, bool in_x, int64_t blk_x, int64_t blk_y) {
  if (blk_x <= 0) return i;
  const int64_t blk = in_x ? blk_x : blk_y;
  return (i / blk) * (blk_x + blk_y) + (in_x ? 0 : blk_x) + i % blk;
}

// Coefficients of the MMD^2 bilinear form (per ORDERED pair of the stacked Gram).
struct Coefs {
  double a_xx, a_yy, a_xy;  // a_xy = -1/(m n)
  int diag_in_sum;          // 1 if K_ii enters the estimator (biased, or unbiased + const-diag quirk)
  int f16;                  // tensor-core gradient paths: fp16 operands (SMMD_PREC_FP16) instead of bf16
};

inline Coefs make_coefs(const Geometry& g, const KernelFn& k) {
  Coefs c;
  c.f16 = 0;
  double m = (double)g.m, n = (double)g.n;
  if (g.biased) {
    c.a_xx = 1.0 / (m * m);
    c.a_yy = 1.0 / (n * n);
  } else {
    c.a_xx = 1.0 / (m * (m - 1.0));
    c.a_yy = 1.0 / (n * (n - 1.0));
  }
  c.a_xy = -1.0 / (m * n);
  c.diag_in_sum = (g.biased || k.has_const_diag) ? 1 : 0;
  return c;
}

// Row statistics emitted by the row kernels (one record per owned stacked row), all double.
enum RowStat : int {
  RS_SAME = 0,   // sum_{j same set, j != i} k_ij
  RS_CROSS = 1,  // sum_{j other set} k_ij
  RS_SQ_SAME = 2,
  RS_SQ_CROSS = 3,
  RS_DIAG = 4,   // k_ii (analytic)
  RS_PAIR = 5,   // k(x_i, y_i) (the XY diagonal; u-statistic)
  RS_COUNT = 6
};

// ---- roofline instrumentation (smmd_capi.cu): event pair around the dominant kernel of a call ----------
void prof_begin(cudaStream_t s);
void prof_end(cudaStream_t s);
void prof_mark(cudaStream_t s);   // boundary between the two kernels of a two-pass path

// ---- peer-memory exchange (smmd_peer.cu) -----------------------------------------------------------------
void peer_after_pull(cudaStream_t s);   // smmd_capi.cu: records the caller's event once the peers' rows have been pulled
cudaError_t launch_peer_publish(const float* X, const float* Y, int64_t ld, int64_t blk_x, int64_t blk_y, int64_t d,
                                int to_bf16, const smmd_peer_table& pt, uint64_t step, cudaStream_t s);
PeerSrc make_peer_src(const smmd_peer_table& pt, int64_t rows_local, int64_t d, uint64_t step);
cudaError_t launch_peer_combine(const KernelFn& kf, const Geometry& g, double* scalars, const smmd_peer_table& pt,
                                uint64_t step, cudaStream_t s);
bool peer_small_eligible(const KernelFn& kf, const Geometry& g);
cudaError_t launch_peer_small_mmd2(const KernelFn& kf, const Geometry& g, const Coefs& c, const float* X, const float* Y,
                                   int64_t ld, float* dX, float* dY, double* partials, unsigned int* counters,
                                   double* scalars, const smmd_peer_table& pt, uint64_t step, cudaStream_t s);

// ---- launches implemented in smmd_simt.cu -------------------------------------------------------
struct SimtPlan {
  int64_t dpitch;      // fp32 feature pitch (multiple of 8)
  int64_t M;           // m + n
  size_t off_Z, off_norm, off_stats, off_end;
};
SimtPlan simt_plan(int64_t m, int64_t n, int64_t d, int64_t batch);

cudaError_t launch_prep_f32(const SrcLayout& src, int64_t m, int64_t n, int64_t d, int tanh_features, float* Z,
                            float* norms, int64_t dpitch, cudaStream_t s);
cudaError_t launch_gather_f32(const void* G, const void* R, int dtype, int64_t ldg, int64_t ldr, int64_t d,
                              const int32_t* idx_g, const int32_t* idx_r, int64_t first, int64_t nsub, int64_t msub,
                              float* Z, float* norms, int64_t dpitch, cudaStream_t s);
cudaError_t launch_simt_rows(const KernelFn& kf, const Geometry& g, const Coefs& c, const float* Z,
                             const float* norms, int64_t dpitch, int64_t batch, double* stats, float* dX, float* dY,
                             int want_stats2, cudaStream_t s);
cudaError_t launch_finalize_mmd2(const KernelFn& kf, const Geometry& g, const double* stats, const float* norms,
                                 double* scalars, cudaStream_t s);
// second-stage reduction: partials[nblocks][6] = (sxx, syy, sxy, syx, dgx, dgy) per CTA, summed in fixed order
cudaError_t launch_finalize_partials(const KernelFn& kf, const Geometry& g, const double* partials, int64_t nblocks,
                                     double* scalars, cudaStream_t s);
cudaError_t launch_combine_mmd2(const KernelFn& kf, const Geometry& g, const double* sums, double* out,
                                cudaStream_t s);
cudaError_t launch_finalize_ratio(const KernelFn& kf, const Geometry& g, const double* stats, double min_var_est,
                                  double* scalars, cudaStream_t s);
bool small_mmd2_eligible(const KernelFn& kf, const Geometry& g, const SrcLayout& src);
cudaError_t launch_small_mmd2(const KernelFn& kf, const Geometry& g, const Coefs& c, const SrcLayout& src, float* dX,
                              float* dY, double* partials, unsigned int* counter, double* scalars, cudaStream_t s);
cudaError_t launch_poly_sums(const double* stats, int64_t m, double* out, cudaStream_t s);
cudaError_t launch_finalize_kid(const double* stats, int64_t nsub, int64_t msub, int64_t first, int est,
                                int ret_var, int64_t var_at_m, double* mmd2_out, double* var_out, cudaStream_t s);
cudaError_t launch_kernel_xy(const KernelFn& kf, const float* Z, const float* norms, int64_t dpitch, int64_t m,
                             int64_t n, int64_t d, float* K, int64_t ldk, cudaStream_t s);
cudaError_t launch_kernel_xy_bwd(const KernelFn& kf, const float* Z, const float* norms, int64_t dpitch, int64_t m,
                                 int64_t n, int64_t d, const float* dK, int64_t lddk, float* dX, float* dY,
                                 cudaStream_t s);
cudaError_t launch_kernel_xy_bwd2(const KernelFn& kf, const float* Z, const float* norms, int64_t dpitch, int64_t m,
                                  int64_t n, int64_t d, const float* dK, int64_t lddk, const float* VX, const float* VY,
                                  float* ddK, float* A, float* B, float* gX, float* gY, cudaStream_t s);

}  // namespace smmd

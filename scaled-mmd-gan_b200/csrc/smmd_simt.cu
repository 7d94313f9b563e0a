// smmd_simt.cu -- exact-fp32 SIMT path (tolerance tier rel 1e-5) + the finalisation kernels shared with
// the tensor-core path.
//
// Replaces, for small / exactness-critical problems, the whole chain of TF ops behind
//   mmd2(kernel(X, Y))           gan/core/mmd.py:18-220   (+ autodiff, gan/core/model.py:446,452)
//   mmd2_and_ratio(K)            gan/core/mmd.py:223-293
//   kernel(X, Y, K_XY_only=True) gan/core/mmd.py:31-32,72-73,106-107,172-173
//   polynomial_mmd(_averages)    gan/compute_scores.py:211-335 (exact variant)
// with: one prep/gather pass (stack Z=[X;Y], tanh, squared norms), ONE row kernel that never
// materialises an N x N matrix (warp = row, lanes = 32 columns; Gram entry, kernel transform, block
// sums and the gradient row  g_i = sum_j 4 a_ij k'(D_ij)(z_i - z_j) + 2 a_ij dk/dG z_j  in one sweep),
// and a tiny fp64 finalisation.
#include <cuda_bf16.h>
#include "smmd_kfun.cuh"

namespace smmd {

constexpr int kRowsPerCta = 8;
constexpr int kChunk = 256;  // feature chunk staged in shared memory

// ------------------------------------------------------------------------------------------------
// prep / gather: Z[b][r][:] = f(src row), norms[b][r] = |z|^2 (fp32)
// ------------------------------------------------------------------------------------------------
struct PrepArgs {
  const void* A;  // X or codes_g
  const void* B;  // Y or codes_r
  int dtype;
  int64_t lda, ldb, ma, mb, d, dpitch;
  const int32_t* idxA;  // optional [batch][ma] row indices into A (KID subsets)
  const int32_t* idxB;
  int64_t first_batch;
  int tanh_features;
  float* Z;
  float* norms;
  int64_t blk_a, blk_b;  // gathered block layout (0 = plain), see SrcLayout
};

__global__ void __launch_bounds__(256) prep_rows_kernel(PrepArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t M = a.ma + a.mb;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  const int64_t b = blockIdx.y;
  if (r >= M) return;
  const bool inA = r < a.ma;
  int64_t src = inA ? r : r - a.ma;
  if (a.idxA) src = inA ? a.idxA[(a.first_batch + b) * a.ma + src] : a.idxB[(a.first_batch + b) * a.mb + src];
  else src = src_row(src, inA, a.blk_a, a.blk_b);
  const int64_t ld = inA ? a.lda : a.ldb;
  const void* base = inA ? a.A : a.B;
  float* zrow = a.Z + (b * M + r) * a.dpitch;
  float acc = 0.f;
  for (int64_t c = lane; c < a.dpitch; c += 32) {
    float v = 0.f;
    if (c < a.d) {
      v = a.dtype == SMMD_F32 ? reinterpret_cast<const float*>(base)[src * ld + c]
                              : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[src * ld + c]);
      if (a.tanh_features) v = tanhf(v);
    }
    zrow[c] = v;
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) a.norms[b * M + r] = acc;
}

cudaError_t launch_prep_f32(const SrcLayout& src, int64_t m, int64_t n, int64_t d, int tanh_features, float* Z,
                            float* norms, int64_t dpitch, cudaStream_t s) {
  PrepArgs a{src.X, src.Y, src.dtype, src.ldx, src.ldy, m, n, d, dpitch, nullptr, nullptr, 0, tanh_features, Z, norms,
             src.blk_x, src.blk_y};
  dim3 grid((unsigned)((m + n + 7) / 8), 1);
  prep_rows_kernel<<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_gather_f32(const void* G, const void* R, int dtype, int64_t ldg, int64_t ldr, int64_t d,
                              const int32_t* idx_g, const int32_t* idx_r, int64_t first, int64_t nsub, int64_t msub,
                              float* Z, float* norms, int64_t dpitch, cudaStream_t s) {
  PrepArgs a{G, R, dtype, ldg, ldr, msub, msub, d, dpitch, idx_g, idx_r, first, 0, Z, norms, 0, 0};
  dim3 grid((unsigned)((2 * msub + 7) / 8), (unsigned)nsub);
  prep_rows_kernel<<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}

SimtPlan simt_plan(int64_t m, int64_t n, int64_t d, int64_t batch) {
  SimtPlan p;
  p.dpitch = (d + 7) / 8 * 8;
  p.M = m + n;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t o = 0;
  p.off_Z = o;
  o = up(o + (size_t)batch * p.M * p.dpitch * sizeof(float));
  p.off_norm = o;
  o = up(o + (size_t)batch * p.M * sizeof(float));
  p.off_stats = o;
  o = up(o + (size_t)batch * p.M * RS_COUNT * sizeof(double));
  p.off_end = o;
  return p;
}

// ------------------------------------------------------------------------------------------------
// the row kernel
// ------------------------------------------------------------------------------------------------
struct SimtArgs {
  KernelFn kf;
  const float* Z;
  const float* norms;
  int64_t dpitch, d, m, n, M;
  int64_t x0, ox, y0, oy;  // owned X rows [x0, x0+ox), owned Y rows [y0, y0+oy)
  float a_xx, a_yy, a_xy;
  int diag_in_sum;
  double* stats;  // [batch][ox+oy][RS_COUNT] (may be null in witness mode)
  float* dX;      // [ox][d]
  float* dY;      // [oy][d]
  // witness / VJP mode (K_XY_only backward): pair weight = dK[i][j] on cross pairs, 0 on same-set pairs
  const float* dK;
  int64_t lddk;
};

template <int NCH, bool GRAD>
__global__ void __launch_bounds__(256) simt_rows_kernel(SimtArgs a) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dch = a.dpitch < kChunk ? (int)a.dpitch : kChunk;
  const int cpitch = dch + 4;
  float* colT = sm;                 // [32][cpitch]
  float* rowT = sm + 32 * cpitch;   // [kRowsPerCta][dch]
  const int nch = (int)((a.dpitch + kChunk - 1) / kChunk);

  const int64_t b = blockIdx.y;
  const float* Z = a.Z + b * a.M * a.dpitch;
  const float* norms = a.norms + b * a.M;
  const int64_t owned = a.ox + a.oy;
  const int64_t lr = (int64_t)blockIdx.x * kRowsPerCta + warp;
  const bool row_valid = lr < owned;
  const int64_t ig = row_valid ? (lr < a.ox ? a.x0 + lr : a.m + a.y0 + (lr - a.ox)) : 0;
  const bool rowX = ig < a.m;
  const float ni = norms[ig];
  const bool witness = a.dK != nullptr;

  float4 zi[NCH][2], acc[NCH][2];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      acc[ch][t] = make_float4(0.f, 0.f, 0.f, 0.f);
      zi[ch][t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (GRAD) {
        int64_t c = (int64_t)ch * kChunk + 128 * t + lane * 4;
        if (c < a.dpitch) zi[ch][t] = *reinterpret_cast<const float4*>(Z + ig * a.dpitch + c);
      }
    }

  double s_same = 0, s_cross = 0, q_same = 0, q_cross = 0, pairv = 0, wsum = 0;
  // column range: witness mode only visits the other set
  int64_t jbeg = 0, jend = a.M;
  if (witness) {
    jbeg = rowX ? a.m : 0;
    jend = rowX ? a.M : a.m;
  }
  // all warps of the CTA must walk the same tiles (shared colT): take the union over the CTA's rows.
  // Rows of a CTA can straddle the X/Y boundary only in witness mode with mixed row types; handle by
  // visiting [0, M) in that (rare) case.
  {
    __shared__ int mixed;
    if (tid == 0) mixed = 0;
    __syncthreads();
    if (witness && row_valid) {
      int64_t first_ig = ((int64_t)blockIdx.x * kRowsPerCta < a.ox) ? a.x0 : a.m;  // type of row 0 of this CTA
      bool firstX = first_ig < a.m;
      if (firstX != rowX) mixed = 1;
    }
    __syncthreads();
    if (witness && mixed) {
      jbeg = 0;
      jend = a.M;
    }
    if (witness && !row_valid) {  // idle warps follow row 0 of the CTA
      bool firstX = ((int64_t)blockIdx.x * kRowsPerCta < a.ox);
      jbeg = mixed ? 0 : (firstX ? a.m : 0);
      jend = mixed ? a.M : (firstX ? a.M : a.m);
    }
  }
  const int64_t jt0 = jbeg / 32, jt1 = (jend + 31) / 32;

  // which pair quantities the family needs: G = <a,b> and/or D = |a-b|^2 (difference form)
  const bool needG = a.kf.family == FAM_DOT || a.kf.family == FAM_POLY || a.kf.add_dot > 0.f;
  const bool needD = a.kf.family == FAM_RBF || a.kf.family == FAM_RQ || a.kf.family == FAM_DISTANCE;

  for (int64_t jt = jt0; jt < jt1; ++jt) {
    float S = 0.f, Dd = 0.f;
    for (int ch = 0; ch < nch; ++ch) {
      __syncthreads();
      const int q4 = dch >> 2;
      for (int idx = tid; idx < 32 * q4; idx += 256) {
        int r = idx / q4, c4 = idx - r * q4;
        int64_t jg = jt * 32 + r, c = (int64_t)ch * kChunk + c4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (jg < a.M && c < a.dpitch) v = *reinterpret_cast<const float4*>(Z + jg * a.dpitch + c);
        *reinterpret_cast<float4*>(colT + r * cpitch + c4 * 4) = v;
      }
      if (nch > 1 || jt == jt0) {
        for (int idx = tid; idx < kRowsPerCta * q4; idx += 256) {
          int r = idx / q4, c4 = idx - r * q4;
          int64_t l2 = (int64_t)blockIdx.x * kRowsPerCta + r, c = (int64_t)ch * kChunk + c4 * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (l2 < owned && c < a.dpitch) {
            int64_t g2 = l2 < a.ox ? a.x0 + l2 : a.m + a.y0 + (l2 - a.ox);
            v = *reinterpret_cast<const float4*>(Z + g2 * a.dpitch + c);
          }
          *reinterpret_cast<float4*>(rowT + r * dch + c4 * 4) = v;
        }
      }
      __syncthreads();
      const float4* rp = reinterpret_cast<const float4*>(rowT + warp * dch);
      const float4* cp = reinterpret_cast<const float4*>(colT + lane * cpitch);
      if (needG) {
#pragma unroll 4
        for (int c4 = 0; c4 < q4; ++c4) {
          float4 x = rp[c4], y = cp[c4];
          S = fmaf(x.x, y.x, S);
          S = fmaf(x.y, y.y, S);
          S = fmaf(x.z, y.z, S);
          S = fmaf(x.w, y.w, S);
        }
      }
      if (needD) {
#pragma unroll 4
        for (int c4 = 0; c4 < q4; ++c4) {
          float4 x = rp[c4], y = cp[c4];
          float e0 = x.x - y.x, e1 = x.y - y.y, e2 = x.z - y.z, e3 = x.w - y.w;
          Dd = fmaf(e0, e0, Dd);
          Dd = fmaf(e1, e1, Dd);
          Dd = fmaf(e2, e2, Dd);
          Dd = fmaf(e3, e3, Dd);
        }
      }
    }
    // ---- pair epilogue: lane <-> column jg ----
    const int64_t jg = jt * 32 + lane;
    const bool valid = row_valid && jg < a.M && jg >= jbeg && jg < jend;
    float wd = 0.f, wg = 0.f;
    if (valid) {
      const float nj = norms[jg];
      const bool colX = jg < a.m;
      const bool same = (colX == rowX);
      const bool isdiag = (jg == ig);
      PairVal pv = eval_exact(a.kf, S, Dd, ni, nj);
      if (witness) {
        if (!same) {
          float w = rowX ? a.dK[ig * a.lddk + (jg - a.m)] : a.dK[jg * a.lddk + (ig - a.m)];
          wd = 2.f * w * pv.kd;
          wg = w * pv.kg;
          wsum += (double)w;
        }
      } else {
        const float aco = same ? (rowX ? a.a_xx : a.a_yy) : a.a_xy;
        if (isdiag) {
          wg = a.diag_in_sum ? 2.f * aco * pv.kg : 0.f;
        } else {
          wd = 4.f * aco * pv.kd;
          wg = 2.f * aco * pv.kg;
          const double k = (double)pv.k;
          if (same) {
            s_same += k;
            q_same += k * k;
          } else {
            s_cross += k;
            q_cross += k * k;
            if (rowX && (jg - a.m) == ig) pairv = k;
          }
        }
      }
    }
    if (GRAD) {
      for (int ch = 0; ch < nch; ++ch) {
        if (nch > 1) {  // bring chunk `ch` of the column tile back
          __syncthreads();
          const int q4 = dch >> 2;
          for (int idx = tid; idx < 32 * q4; idx += 256) {
            int r = idx / q4, c4 = idx - r * q4;
            int64_t j2 = jt * 32 + r, c = (int64_t)ch * kChunk + c4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j2 < a.M && c < a.dpitch) v = *reinterpret_cast<const float4*>(Z + j2 * a.dpitch + c);
            *reinterpret_cast<float4*>(colT + r * cpitch + c4 * 4) = v;
          }
          __syncthreads();
        }
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          if (cc != ch) continue;
          for (int jj = 0; jj < 32; ++jj) {
            const float wdj = __shfl_sync(0xffffffffu, wd, jj);
            const float wgj = __shfl_sync(0xffffffffu, wg, jj);
            if (wdj == 0.f && wgj == 0.f) continue;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int c = 128 * t + lane * 4;
              if (c < dch) {
                const float4 zj = *reinterpret_cast<const float4*>(colT + jj * cpitch + c);
                float4& A = acc[cc][t];
                const float4 z = zi[cc][t];
                A.x = fmaf(wdj, z.x - zj.x, fmaf(wgj, zj.x, A.x));
                A.y = fmaf(wdj, z.y - zj.y, fmaf(wgj, zj.y, A.y));
                A.z = fmaf(wdj, z.z - zj.z, fmaf(wgj, zj.z, A.z));
                A.w = fmaf(wdj, z.w - zj.w, fmaf(wgj, zj.w, A.w));
              }
            }
          }
        }
      }
    }
  }

  // ---- per-row outputs ----
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_same += __shfl_xor_sync(0xffffffffu, s_same, o);
    s_cross += __shfl_xor_sync(0xffffffffu, s_cross, o);
    q_same += __shfl_xor_sync(0xffffffffu, q_same, o);
    q_cross += __shfl_xor_sync(0xffffffffu, q_cross, o);
    pairv += __shfl_xor_sync(0xffffffffu, pairv, o);
    wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  }
  if (!row_valid) return;
  if (a.stats && lane == 0) {
    double* st = a.stats + (b * owned + lr) * RS_COUNT;
    st[RS_SAME] = s_same;
    st[RS_CROSS] = s_cross;
    st[RS_SQ_SAME] = q_same;
    st[RS_SQ_CROSS] = q_cross;
    double dg;
    if (a.kf.family == FAM_RQ) dg = (double)a.kf.const_diag + (double)a.kf.add_dot * (double)ni;
    else dg = (double)diag_value(a.kf, ni);
    st[RS_DIAG] = dg;
    st[RS_PAIR] = pairv;
  }
  if (GRAD) {
    float* out = rowX ? a.dX + (ig - a.x0) * a.d : a.dY + (ig - a.m - a.y0) * a.d;
    // distance kernel with its sqrt(|x|^2+eps) terms (witness mode only): d/dz_i sqrt(n_i+eps) = z_i/sqrt(n_i+eps)
    float nterm = 0.f;
    if (witness && a.kf.family == FAM_DISTANCE && a.kf.true_distance) nterm = (float)wsum * rsqrtf(ni + kEps);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int64_t c = (int64_t)ch * kChunk + 128 * t + lane * 4;
        const float g[4] = {acc[ch][t].x, acc[ch][t].y, acc[ch][t].z, acc[ch][t].w};
        const float z[4] = {zi[ch][t].x, zi[ch][t].y, zi[ch][t].z, zi[ch][t].w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c + e < a.d && (128 * t + lane * 4) < dch) {
            float v = g[e] + nterm * z[e];
            if (a.kf.tanh_features) v *= (1.f - z[e] * z[e]);
            out[c + e] = v;
          }
      }
  }
}

template <bool GRAD>
static cudaError_t launch_rows_t(const SimtArgs& a, int64_t batch, cudaStream_t s) {
  const int dch = a.dpitch < kChunk ? (int)a.dpitch : kChunk;
  const size_t smem = (size_t)(32 * (dch + 4) + kRowsPerCta * dch) * sizeof(float);
  const int64_t owned = a.ox + a.oy;
  dim3 grid((unsigned)((owned + kRowsPerCta - 1) / kRowsPerCta), (unsigned)batch);
  const int nch = (int)((a.dpitch + kChunk - 1) / kChunk);
  if (nch <= 1) simt_rows_kernel<1, GRAD><<<grid, 256, smem, s>>>(a);
  else if (nch <= 2) simt_rows_kernel<2, GRAD><<<grid, 256, smem, s>>>(a);
  else if (nch <= 4) simt_rows_kernel<4, GRAD><<<grid, 256, smem, s>>>(a);
  else if (nch <= 8) simt_rows_kernel<(GRAD ? 8 : 1), GRAD><<<grid, 256, smem, s>>>(a);
  else if (!GRAD) simt_rows_kernel<1, GRAD><<<grid, 256, smem, s>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_simt_rows(const KernelFn& kf, const Geometry& g, const Coefs& c, const float* Z,
                             const float* norms, int64_t dpitch, int64_t batch, double* stats, float* dX, float* dY,
                             int /*want_stats2*/, cudaStream_t s) {
  SimtArgs a;
  a.kf = kf;
  a.Z = Z;
  a.norms = norms;
  a.dpitch = dpitch;
  a.d = g.d;
  a.m = g.m;
  a.n = g.n;
  a.M = g.m + g.n;
  a.x0 = g.x0;
  a.ox = g.x1 - g.x0;
  a.y0 = g.y0;
  a.oy = g.y1 - g.y0;
  a.a_xx = (float)c.a_xx;
  a.a_yy = (float)c.a_yy;
  a.a_xy = (float)c.a_xy;
  a.diag_in_sum = c.diag_in_sum;
  a.stats = stats;
  a.dX = dX;
  a.dY = dY;
  a.dK = nullptr;
  a.lddk = 0;
  if (dX != nullptr || dY != nullptr) return launch_rows_t<true>(a, batch, s);
  return launch_rows_t<false>(a, batch, s);
}

// ------------------------------------------------------------------------------------------------
// block reduction helper for the finalisers (fixed order -> deterministic)
// ------------------------------------------------------------------------------------------------
template <int NQ>
__device__ void block_reduce(double (&q)[NQ], double* sh /* [NQ][256] */) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NQ; ++i) sh[i * 256 + tid] = q[i];
  __syncthreads();
  for (int stride = 128; stride > 0; stride >>= 1) {
    if (tid < stride) {
#pragma unroll
      for (int i = 0; i < NQ; ++i) sh[i * 256 + tid] += sh[i * 256 + tid + stride];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < NQ; ++i) q[i] = sh[i * 256];
}

// same tree as block_reduce, for the first 256 threads of a larger block (named barrier 1)
template <int NQ>
__device__ void block_reduce256(double (&q)[NQ], double* sh /* [NQ][256] */) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NQ; ++i) sh[i * 256 + tid] = q[i];
  asm volatile("bar.sync 1, 256;" ::: "memory");
  for (int stride = 128; stride > 0; stride >>= 1) {
    if (tid < stride) {
#pragma unroll
      for (int i = 0; i < NQ; ++i) sh[i * 256 + tid] += sh[i * 256 + tid + stride];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
#pragma unroll
  for (int i = 0; i < NQ; ++i) q[i] = sh[i * 256];
}

struct FinArgs {
  KernelFn kf;
  int64_t m, n, ox, oy;
  int biased;
  int full;  // world == 1: write the final MMD^2
  const double* stats;
  double* scalars;
};

__global__ void __launch_bounds__(512) finalize_mmd2_kernel(FinArgs a) {
  __shared__ double sh[6 * 256];
  __shared__ double sh2[2][6 * 256];
  double q[6] = {0, 0, 0, 0, 0, 0};  // sxx, syy, sxy, syx, dgx, dgy
  // 512 threads, two row streams per reduction lane, 4 rows in flight per thread; partial sums are folded
  // in a fixed order (deterministic).
  const int lane256 = threadIdx.x & 255, stream = threadIdx.x >> 8;
  const int64_t rows = a.ox + a.oy;
  for (int64_t r0 = threadIdx.x; r0 < rows; r0 += 2048) {
    double vs[4], vc[4], vd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = r0 + 512 * u;
      const double* st = a.stats + (r < rows ? r : 0) * RS_COUNT;
      const bool ok = r < rows;
      vs[u] = ok ? st[RS_SAME] : 0.0;
      vc[u] = ok ? st[RS_CROSS] : 0.0;
      vd[u] = ok ? st[RS_DIAG] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = r0 + 512 * u;
      if (r < a.ox) {
        q[0] += vs[u];
        q[2] += vc[u];
        q[4] += vd[u];
      } else {
        q[1] += vs[u];
        q[3] += vc[u];
        q[5] += vd[u];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) sh2[stream][i * 256 + lane256] = q[i];
  __syncthreads();
  if (threadIdx.x >= 256) return;
#pragma unroll
  for (int i = 0; i < 6; ++i) q[i] = sh2[0][i * 256 + lane256] + sh2[1][i * 256 + lane256];
  block_reduce256<6>(q, sh);
  if (threadIdx.x == 0) {
    double* o = a.scalars;
    for (int i = 0; i < SMMD_NUM_SCALARS; ++i) o[i] = 0.0;
    o[SMMD_S_SUM_XX] = q[0];
    o[SMMD_S_SUM_YY] = q[1];
    o[SMMD_S_SUM_XY] = q[2];
    o[SMMD_S_SUM_YX] = q[3];
    o[SMMD_S_DIAG_X] = q[4];
    o[SMMD_S_DIAG_Y] = q[5];
    double v = mmd2_from_sums(a.kf, (double)a.m, (double)a.n, a.biased, q[0], q[1], q[2], q[3], q[4], q[5]);
    o[SMMD_S_MMD2] = a.full ? v : 0.0;
    bool bad = false;
    for (int i = 0; i < 6; ++i) bad = bad || !isfinite(q[i]);
    o[SMMD_S_NONFINITE] = bad ? 1.0 : 0.0;
  }
}

cudaError_t launch_finalize_mmd2(const KernelFn& kf, const Geometry& g, const double* stats, const float*,
                                 double* scalars, cudaStream_t s) {
  FinArgs a{kf, g.m, g.n, g.x1 - g.x0, g.y1 - g.y0, g.biased,
            (g.x0 == 0 && g.x1 == g.m && g.y0 == 0 && g.y1 == g.n) ? 1 : 0, stats, scalars};
  finalize_mmd2_kernel<<<1, 512, 0, s>>>(a);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) finalize_partials_kernel(KernelFn kf, int64_t m, int64_t n, int biased, int full,
                                                                const double* partials, int64_t nblocks,
                                                                double* scalars) {
  __shared__ double sh[6 * 256];
  double q[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t b = threadIdx.x; b < nblocks; b += 256) {
#pragma unroll
    for (int i = 0; i < 6; ++i) q[i] += partials[b * 6 + i];
  }
  block_reduce<6>(q, sh);
  if (threadIdx.x == 0) {
    for (int i = 0; i < SMMD_NUM_SCALARS; ++i) scalars[i] = 0.0;
    scalars[SMMD_S_SUM_XX] = q[0];
    scalars[SMMD_S_SUM_YY] = q[1];
    scalars[SMMD_S_SUM_XY] = q[2];
    scalars[SMMD_S_SUM_YX] = q[3];
    scalars[SMMD_S_DIAG_X] = q[4];
    scalars[SMMD_S_DIAG_Y] = q[5];
    const double v = mmd2_from_sums(kf, (double)m, (double)n, biased, q[0], q[1], q[2], q[3], q[4], q[5]);
    scalars[SMMD_S_MMD2] = full ? v : 0.0;
    bool bad = false;
    for (int i = 0; i < 6; ++i) bad = bad || !isfinite(q[i]);
    scalars[SMMD_S_NONFINITE] = bad ? 1.0 : 0.0;
  }
}

cudaError_t launch_finalize_partials(const KernelFn& kf, const Geometry& g, const double* partials, int64_t nblocks,
                                     double* scalars, cudaStream_t s) {
  const int full = (g.x0 == 0 && g.x1 == g.m && g.y0 == 0 && g.y1 == g.n) ? 1 : 0;
  finalize_partials_kernel<<<1, 256, 0, s>>>(kf, g.m, g.n, g.biased, full, partials, nblocks, scalars);
  return cudaGetLastError();
}

__global__ void combine_mmd2_kernel(KernelFn kf, int64_t m, int64_t n, int biased, const double* sums, double* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    out[0] = mmd2_from_sums(kf, (double)m, (double)n, biased, sums[SMMD_S_SUM_XX], sums[SMMD_S_SUM_YY],
                            sums[SMMD_S_SUM_XY], sums[SMMD_S_SUM_YX], sums[SMMD_S_DIAG_X], sums[SMMD_S_DIAG_Y]);
}
cudaError_t launch_combine_mmd2(const KernelFn& kf, const Geometry& g, const double* sums, double* out,
                                cudaStream_t s) {
  combine_mmd2_kernel<<<1, 32, 0, s>>>(kf, g.m, g.n, g.biased, sums, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// second-order statistics: the 18 sums both variance formulas are built from
// ------------------------------------------------------------------------------------------------
enum {
  Q_SX, Q_SY, Q_CX, Q_CY, Q_DX, Q_DY, Q_D2X, Q_D2Y, Q_SX2, Q_SY2, Q_CX2, Q_CY2, Q_QX, Q_QY, Q_QXY, Q_DOTX, Q_DOTY,
  Q_TRXY, Q_N
};

// stats: [2m][RS_COUNT] (X rows then Y rows).  excess_const > -1: const-diagonal kernels, where the
// reference subtracts `const` instead of the true diagonal (quirk kept: mmd.py:239-243,256-257).
__device__ void gather_second_order(const double* stats, int64_t m, bool const_diag, double cd, double (&q)[Q_N],
                                    double* sh) {
  for (int i = 0; i < Q_N; ++i) q[i] = 0.0;
  for (int64_t i = threadIdx.x; i < m; i += 256) {
    const double* sx = stats + i * RS_COUNT;
    const double* sy = stats + (m + i) * RS_COUNT;
    double rx = sx[RS_SAME], ry = sy[RS_SAME];
    double qx = sx[RS_SQ_SAME], qy = sy[RS_SQ_SAME];
    double dx = sx[RS_DIAG], dy = sy[RS_DIAG];
    if (const_diag) {  // row sums keep (true diag - const); squares keep true_diag^2 - const^2
      rx += dx - cd;
      ry += dy - cd;
      qx += dx * dx - cd * cd;
      qy += dy * dy - cd * cd;
      dx = cd;
      dy = cd;
    }
    q[Q_SX] += rx;
    q[Q_SY] += ry;
    q[Q_CX] += sx[RS_CROSS];
    q[Q_CY] += sy[RS_CROSS];
    q[Q_DX] += dx;
    q[Q_DY] += dy;
    q[Q_D2X] += dx * dx;
    q[Q_D2Y] += dy * dy;
    q[Q_SX2] += rx * rx;
    q[Q_SY2] += ry * ry;
    q[Q_CX2] += sx[RS_CROSS] * sx[RS_CROSS];
    q[Q_CY2] += sy[RS_CROSS] * sy[RS_CROSS];
    q[Q_QX] += qx;
    q[Q_QY] += qy;
    q[Q_QXY] += sx[RS_SQ_CROSS];
    q[Q_DOTX] += rx * sx[RS_CROSS];
    q[Q_DOTY] += ry * sy[RS_CROSS];
    q[Q_TRXY] += sx[RS_PAIR];
  }
  block_reduce<Q_N>(q, sh);
}

// gan/core/mmd.py:234-293
__global__ void __launch_bounds__(256) finalize_ratio_kernel(KernelFn kf, int64_t mm, int biased, const double* stats,
                                                             double min_var_est, double* scalars) {
  extern __shared__ double shd[];
  double q[Q_N];
  gather_second_order(stats, mm, kf.has_const_diag != 0, (double)kf.const_diag, q, shd);
  if (threadIdx.x != 0) return;
  const double m = (double)mm;
  const double sX = q[Q_SX], sY = q[Q_SY], sXY = q[Q_CX];
  double val;
  if (biased) val = (sX + q[Q_DX]) / (m * m) + (sY + q[Q_DY]) / (m * m) - 2 * sXY / (m * m);
  else val = (sX + q[Q_DX]) / (m * (m - 1)) + (sY + q[Q_DY]) / (m * (m - 1)) - 2 * sXY / (m * m);
  const double var = 2 / (m * m * (m - 1) * (m - 1)) * (2 * q[Q_SX2] - q[Q_QX] + 2 * q[Q_SY2] - q[Q_QY]) -
                     (4 * m - 6) / (m * m * m * (m - 1) * (m - 1) * (m - 1)) * (sX * sX + sY * sY) +
                     4 * (m - 2) / (m * m * m * (m - 1) * (m - 1)) * (q[Q_CX2] + q[Q_CY2]) -
                     4 * (m - 3) / (m * m * m * (m - 1) * (m - 1)) * q[Q_QXY] -
                     (8 * m - 12) / (m * m * m * m * m * (m - 1)) * sXY * sXY +
                     8 / (m * m * m * (m - 1)) * (1 / m * (sX + sY) * sXY - q[Q_DOTX] - q[Q_DOTY]);
  for (int i = 0; i < SMMD_NUM_SCALARS; ++i) scalars[i] = 0.0;
  scalars[SMMD_S_MMD2] = val;
  scalars[SMMD_S_VAR] = var;
  scalars[SMMD_S_RATIO] = val / sqrt(fmax(var, min_var_est));
  scalars[SMMD_S_NONFINITE] = (isfinite(val) && isfinite(var)) ? 0.0 : 1.0;
}

cudaError_t launch_finalize_ratio(const KernelFn& kf, const Geometry& g, const double* stats, double min_var_est,
                                  double* scalars, cudaStream_t s) {
  finalize_ratio_kernel<<<1, 256, Q_N * 256 * sizeof(double), s>>>(kf, g.m, g.biased, stats, min_var_est, scalars);
  return cudaGetLastError();
}

// gan/compute_scores.py:252-335, one CTA per subset
__global__ void __launch_bounds__(256) finalize_kid_kernel(const double* stats, int64_t msub, int64_t first, int est,
                                                           int ret_var, int64_t var_at_m, double* mmd2_out,
                                                           double* var_out) {
  extern __shared__ double shd[];
  const int64_t sub = blockIdx.x;
  double q[Q_N];
  gather_second_order(stats + sub * 2 * msub * RS_COUNT, msub, false, 0.0, q, shd);
  if (threadIdx.x != 0) return;
  const double m = (double)msub, m1 = m - 1, m2 = m - 2;
  const double sX = q[Q_SX], sY = q[Q_SY], sXY = q[Q_CX];
  double val;
  if (est == SMMD_EST_BIASED) val = (sX + q[Q_DX]) / (m * m) + (sY + q[Q_DY]) / (m * m) - 2 * sXY / (m * m);
  else {
    val = (sX + sY) / (m * m1);
    if (est == SMMD_EST_UNBIASED) val -= 2 * sXY / (m * m);
    else val -= 2 * (sXY - q[Q_TRXY]) / (m * m1);
  }
  mmd2_out[first + sub] = val;
  if (!ret_var || var_out == nullptr) return;
  const double dots = q[Q_DOTX] + q[Q_DOTY];
  const double zeta1 = 1 / (m * m1 * m2) * (q[Q_SX2] - q[Q_QX] + q[Q_SY2] - q[Q_QY]) -
                       1 / ((m * m1) * (m * m1)) * (sX * sX + sY * sY) +
                       1 / (m * m * m1) * (q[Q_CX2] + q[Q_CY2] - 2 * q[Q_QXY]) - 2 / (m * m * m * m) * sXY * sXY -
                       2 / (m * m * m1) * dots + 2 / (m * m * m * m1) * (sX + sY) * sXY;
  const double zeta2 = 1 / (m * m1) * (q[Q_QX] + q[Q_QY]) - 1 / ((m * m1) * (m * m1)) * (sX * sX + sY * sY) +
                       2 / (m * m) * q[Q_QXY] - 2 / (m * m * m * m) * sXY * sXY - 4 / (m * m * m1) * dots +
                       4 / (m * m * m * m1) * (sX + sY) * sXY;
  const double vm = (double)var_at_m;
  var_out[first + sub] = 4 * (vm - 2) / (vm * (vm - 1)) * zeta1 + 2 / (vm * (vm - 1)) * zeta2;
}

cudaError_t launch_finalize_kid(const double* stats, int64_t nsub, int64_t msub, int64_t first, int est,
                                int ret_var, int64_t var_at_m, double* mmd2_out, double* var_out, cudaStream_t s) {
  finalize_kid_kernel<<<(unsigned)nsub, 256, Q_N * 256 * sizeof(double), s>>>(stats, msub, first, est, ret_var,
                                                                            var_at_m, mmd2_out, var_out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// latency-bound shapes: ONE launch (every shipped YAML runs batch 64 with dof_dim 1..16)
// ------------------------------------------------------------------------------------------------
// Warp = row i, lanes = columns j (no shared column tiles, no block barriers in the pair loop): z_i and z_j live in
// registers (DMAX features, template), S / D / |z_j|^2 straight from the fp32 inputs (8 KB in total, L1 resident),
// eval_exact, fp64 row sums, and the gradient contribution of the pair is accumulated right there from the same
// registers; the 32 per-lane partial rows are folded through shared memory in fixed order.  (A first version with
// lanes = features in a second pass over the weights was a 5800-instruction dependent chain per warp: 28 us.)
// The last CTA to finish
// (ticket counter, zeroed by a memset node in front of the launch) folds the per-CTA partial sums in fixed order
// and writes the scalars -- prep, row kernel and finalize of the general exact path in a single kernel.
constexpr int kSmallMaxD = 64;
constexpr int kSmallMaxM = 1024;

struct SmallArgs {
  KernelFn kf;
  const float* X;
  const float* Y;
  int64_t ldx, ldy;
  int m, n, d;
  float a_xx, a_yy, a_xy;
  int diag_in_sum, biased;
  float* dX;
  float* dY;
  double* partials;        // [gridDim.x][6]
  unsigned int* counter;
  double* scalars;
};

template <int DMAX>
__global__ void __launch_bounds__(256) small_mmd2_kernel(SmallArgs a) {
  extern __shared__ float sm[];
  const int M = a.m + a.n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* accS = sm + (warp * 32 + lane) * (DMAX + 1);   // [8][32][DMAX + 1]: per-lane partial gradient rows
  __shared__ double red[kRowsPerCta][6];
  __shared__ int is_last;
  const int ig = blockIdx.x * kRowsPerCta + warp;
  const bool row_valid = ig < M;
  const bool rowX = ig < a.m;
  double q[6] = {0, 0, 0, 0, 0, 0};   // sxx, syy, sxy, syx, dgx, dgy of this row
  if (row_valid) {
    const float* zi_p = rowX ? a.X + (int64_t)ig * a.ldx : a.Y + (int64_t)(ig - a.m) * a.ldy;
    float zi[DMAX], acc[DMAX];
    float ni = 0.f;
#pragma unroll
    for (int c = 0; c < DMAX; ++c) {
      zi[c] = c < a.d ? zi_p[c] : 0.f;     // same address in all lanes: one broadcast load per feature
      ni = fmaf(zi[c], zi[c], ni);
      acc[c] = 0.f;
    }
    double s_same = 0.0, s_cross = 0.0;
    for (int j = lane; j < M; j += 32) {
      const bool colX = j < a.m;
      const float* zj_p = colX ? a.X + (int64_t)j * a.ldx : a.Y + (int64_t)(j - a.m) * a.ldy;
      float zj[DMAX];
      float S = 0.f, Dd = 0.f, nj = 0.f;
#pragma unroll
      for (int c = 0; c < DMAX; ++c) {
        zj[c] = c < a.d ? zj_p[c] : 0.f;
        const float e = zi[c] - zj[c];
        S = fmaf(zi[c], zj[c], S);
        Dd = fmaf(e, e, Dd);
        nj = fmaf(zj[c], zj[c], nj);
      }
      const PairVal pv = eval_exact(a.kf, S, Dd, ni, nj);
      const bool same = (colX == rowX);
      const float aco = same ? (rowX ? a.a_xx : a.a_yy) : a.a_xy;
      float wd = 0.f, wg = 0.f;
      if (j == ig) {
        wg = a.diag_in_sum ? 2.f * aco * pv.kg : 0.f;
      } else {
        wd = 4.f * aco * pv.kd;
        wg = 2.f * aco * pv.kg;
        if (same) s_same += (double)pv.k;
        else s_cross += (double)pv.k;
      }
      if (a.dX) {
#pragma unroll
        for (int c = 0; c < DMAX; ++c) acc[c] = fmaf(wd, zi[c] - zj[c], fmaf(wg, zj[c], acc[c]));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_same += __shfl_xor_sync(0xffffffffu, s_same, o);
      s_cross += __shfl_xor_sync(0xffffffffu, s_cross, o);
    }
    const double dg = a.kf.family == FAM_RQ ? (double)a.kf.const_diag + (double)a.kf.add_dot * (double)ni
                                            : (double)diag_value(a.kf, ni);
    q[rowX ? 0 : 1] = s_same;
    q[rowX ? 2 : 3] = s_cross;
    q[rowX ? 4 : 5] = dg;
    if (a.dX) {
      // fold the 32 per-lane partial rows: lane c sums feature c over the lanes in fixed order
#pragma unroll
      for (int c = 0; c < DMAX; ++c) accS[c] = acc[c];
      __syncwarp();
      const float* wbase = sm + warp * 32 * (DMAX + 1);
      float* out = rowX ? a.dX + (int64_t)ig * a.d : a.dY + (int64_t)(ig - a.m) * a.d;
      for (int c = lane; c < a.d; c += 32) {
        float t = 0.f;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) t += wbase[l * (DMAX + 1) + c];
        out[c] = t;
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) red[warp][i] = q[i];
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) {
      double t = 0.0;
      for (int w = 0; w < kRowsPerCta; ++w) t += red[w][i];   // fixed order
      a.partials[(int64_t)blockIdx.x * 6 + i] = t;
    }
    __threadfence();
    is_last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid < 6) {
    double t = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(a.partials + (int64_t)b * 6 + tid);   // fixed order
    red[0][tid] = t;
  }
  __syncthreads();
  if (tid == 0) {
    double* o = a.scalars;
    const double* t = red[0];
    for (int i = 0; i < SMMD_NUM_SCALARS; ++i) o[i] = 0.0;
    o[SMMD_S_SUM_XX] = t[0];
    o[SMMD_S_SUM_YY] = t[1];
    o[SMMD_S_SUM_XY] = t[2];
    o[SMMD_S_SUM_YX] = t[3];
    o[SMMD_S_DIAG_X] = t[4];
    o[SMMD_S_DIAG_Y] = t[5];
    o[SMMD_S_MMD2] = mmd2_from_sums(a.kf, (double)a.m, (double)a.n, a.biased, t[0], t[1], t[2], t[3], t[4], t[5]);
    bool bad = false;
    for (int i = 0; i < 6; ++i) bad = bad || !isfinite(t[i]);
    o[SMMD_S_NONFINITE] = bad ? 1.0 : 0.0;
  }
}

bool small_mmd2_eligible(const KernelFn& kf, const Geometry& g, const SrcLayout& src) {
  return src.dtype == SMMD_F32 && src.blk_x == 0 && src.blk_y == 0 && src.Xo == nullptr && !kf.tanh_features &&
         g.d <= kSmallMaxD && g.m + g.n <= kSmallMaxM && g.x0 == 0 && g.x1 == g.m && g.y0 == 0 && g.y1 == g.n &&
         kf.family != FAM_POLY;
}

cudaError_t launch_small_mmd2(const KernelFn& kf, const Geometry& g, const Coefs& c, const SrcLayout& src, float* dX,
                              float* dY, double* partials, unsigned int* counter, double* scalars, cudaStream_t s) {
  SmallArgs a;
  a.kf = kf;
  a.X = static_cast<const float*>(src.X);
  a.Y = static_cast<const float*>(src.Y);
  a.ldx = src.ldx;
  a.ldy = src.ldy;
  a.m = (int)g.m;
  a.n = (int)g.n;
  a.d = (int)g.d;
  a.a_xx = (float)c.a_xx;
  a.a_yy = (float)c.a_yy;
  a.a_xy = (float)c.a_xy;
  a.diag_in_sum = c.diag_in_sum;
  a.biased = g.biased;
  a.dX = dX;
  a.dY = dY;
  a.partials = partials;
  a.counter = counter;
  a.scalars = scalars;
  const int M = a.m + a.n;
  cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), s);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((M + kRowsPerCta - 1) / kRowsPerCta);
  auto go = [&](auto kern, int dmax) -> cudaError_t {
    const size_t smem = (size_t)kRowsPerCta * 32 * (dmax + 1) * sizeof(float);
    if (smem > 48 * 1024) {
      cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e2 != cudaSuccess) return e2;
    }
    kern<<<grid, 256, smem, s>>>(a);
    return cudaGetLastError();
  };
  if (a.d <= 4) return go(small_mmd2_kernel<4>, 4);
  if (a.d <= 16) return go(small_mmd2_kernel<16>, 16);
  if (a.d <= 32) return go(small_mmd2_kernel<32>, 32);
  return go(small_mmd2_kernel<64>, 64);
}

// gan/core/mmd.py:515-539 (_np_get_sums): per-row statistics -> the five "Y related sums" of the 3-sample test.
// stats: [2m][RS_COUNT] (X rows then Y rows); out: [0,m) Kt_YY_sums, [m,2m) K_XY_sums_0 (per y_j), [2m,3m)
// K_XY_sums_1 (per x_i), [3m] Kt_YY_2_sum, [3m+1] K_XY_2_sum.  One CTA, fixed reduction order.
__global__ void __launch_bounds__(256) poly_sums_kernel(const double* stats, int64_t m, double* out) {
  extern __shared__ double shd[];
  double q[2] = {0.0, 0.0};
  for (int64_t i = threadIdx.x; i < m; i += 256) {
    const double* sx = stats + i * RS_COUNT;
    const double* sy = stats + (m + i) * RS_COUNT;
    out[i] = sy[RS_SAME];            // off-diagonal row sum of K_YY
    out[m + i] = sy[RS_CROSS];       // sum_x K(x, y_i)
    out[2 * m + i] = sx[RS_CROSS];   // sum_y K(x_i, y)
    q[0] += sy[RS_SQ_SAME];
    q[1] += sx[RS_SQ_CROSS];
  }
  block_reduce<2>(q, shd);
  if (threadIdx.x == 0) {
    out[3 * m] = q[0];
    out[3 * m + 1] = q[1];
  }
}

cudaError_t launch_poly_sums(const double* stats, int64_t m, double* out, cudaStream_t s) {
  poly_sums_kernel<<<1, 256, 2 * 256 * sizeof(double), s>>>(stats, m, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K_XY_only: dense witness block (the one place an m x n matrix is written, because the caller asks
// for it) and its VJP through the row kernel in witness mode.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kernel_xy_kernel(KernelFn kf, const float* Z, const float* norms,
                                                        int64_t dpitch, int64_t m, int64_t n, float* K, int64_t ldk) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.y * 8 + w, j = (int64_t)blockIdx.x * 32 + lane;
  if (i >= m || j >= n) return;
  const float4* a = reinterpret_cast<const float4*>(Z + i * dpitch);
  const float4* b = reinterpret_cast<const float4*>(Z + (m + j) * dpitch);
  float S = 0.f, Dd = 0.f;
  for (int64_t c = 0; c < dpitch / 4; ++c) {
    float4 x = a[c], y = b[c];
    S = fmaf(x.x, y.x, S);
    S = fmaf(x.y, y.y, S);
    S = fmaf(x.z, y.z, S);
    S = fmaf(x.w, y.w, S);
    float e0 = x.x - y.x, e1 = x.y - y.y, e2 = x.z - y.z, e3 = x.w - y.w;
    Dd = fmaf(e0, e0, Dd);
    Dd = fmaf(e1, e1, Dd);
    Dd = fmaf(e2, e2, Dd);
    Dd = fmaf(e3, e3, Dd);
  }
  K[i * ldk + j] = eval_exact(kf, S, Dd, norms[i], norms[m + j]).k;
}

cudaError_t launch_kernel_xy(const KernelFn& kf, const float* Z, const float* norms, int64_t dpitch, int64_t m,
                             int64_t n, int64_t, float* K, int64_t ldk, cudaStream_t s) {
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((m + 7) / 8));
  kernel_xy_kernel<<<grid, 256, 0, s>>>(kf, Z, norms, dpitch, m, n, K, ldk);
  return cudaGetLastError();
}

cudaError_t launch_kernel_xy_bwd(const KernelFn& kf, const float* Z, const float* norms, int64_t dpitch, int64_t m,
                                 int64_t n, int64_t d, const float* dK, int64_t lddk, float* dX, float* dY,
                                 cudaStream_t s) {
  SimtArgs a;
  a.kf = kf;
  a.Z = Z;
  a.norms = norms;
  a.dpitch = dpitch;
  a.d = d;
  a.m = m;
  a.n = n;
  a.M = m + n;
  a.x0 = 0;
  a.ox = m;
  a.y0 = 0;
  a.oy = n;
  a.a_xx = a.a_yy = a.a_xy = 0.f;
  a.diag_in_sum = 0;
  a.stats = nullptr;
  a.dX = dX;
  a.dY = dY;
  a.dK = dK;
  a.lddk = lddk;
  return launch_rows_t<true>(a, 1, s);
}

// ------------------------------------------------------------------------------------------------
// second-order VJP of the witness block (gradient penalty: the gradient of |d witness / d x_hat| flows back
// into the critic through K_XY's first backward, gan/core/model.py:336-341).
// First backward:  gX_i = sum_j dK_ij [2 k'(D_ij)(x_i - y_j) + kg y_j + 2 f'(|x_i|^2) x_i]   (gY_j alike).
// With cotangents VX, VY of (gX, gY):  L = sum_ij dK_ij [2 k' u_ij + kg p_ij + 2 f'(n_i) s_i + 2 f'(n_j) t_j],
//   u_ij = <VX_i - VY_j, x_i - y_j>,  p_ij = <VX_i, y_j> + <VY_j, x_i>,  s_i = <VX_i, x_i>,  t_j = <VY_j, y_j>.
// Kernel 1 (per pair): dL/ddK_ij and the pair weights A_ij = 4 dK_ij k'' u_ij, B_ij = 2 dK_ij k'.
// Kernel 2 (per row):  dL/dx_i = sum_j [A_ij (x_i - y_j) + B_ij (VX_i - VY_j) + kg dK_ij VY_j]
//                                + r_i [4 f''(n_i) s_i x_i + 2 f'(n_i) VX_i],   r_i = sum_j dK_ij   (dL/dy_j alike).
// f(n) = sqrt(n + eps) only for the distance kernel (mmd.py:29); zero otherwise.
// ------------------------------------------------------------------------------------------------
struct Kxy2Args {
  KernelFn kf;
  const float* Z;       // [m + n][dpitch] stacked fp32 rows (prep output)
  const float* norms;
  int64_t dpitch, m, n, d;
  const float* dK;
  int64_t lddk;
  const float* VX;      // [m][d] or null
  const float* VY;      // [n][d] or null
  float* ddK;           // [m][n]
  float* A;             // [m][n] scratch
  float* B;             // [m][n] scratch
  float* gX;            // [m][d]
  float* gY;            // [n][d]
  float kg;             // coefficient of the Gram (dot) part of the kernel
};

__device__ __forceinline__ float dist_f1(float n) { return 0.5f * rsqrtf(n + kEps); }              // f'
__device__ __forceinline__ float dist_f2(float n) { return -0.25f * rsqrtf(n + kEps) / (n + kEps); }  // f''

__global__ void __launch_bounds__(256) kxy_bwd2_pair_kernel(Kxy2Args a) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.y * 8 + w, j = (int64_t)blockIdx.x * 32 + lane;
  if (i >= a.m || j >= a.n) return;
  const float* x = a.Z + i * a.dpitch;
  const float* y = a.Z + (a.m + j) * a.dpitch;
  const float* vx = a.VX ? a.VX + i * a.d : nullptr;
  const float* vy = a.VY ? a.VY + j * a.d : nullptr;
  float Dd = 0.f, u = 0.f, pd = 0.f, sx = 0.f, ty = 0.f;
  for (int64_t c = 0; c < a.d; ++c) {
    const float xc = x[c], yc = y[c];
    const float vxc = vx ? vx[c] : 0.f, vyc = vy ? vy[c] : 0.f;
    const float e = xc - yc;
    Dd = fmaf(e, e, Dd);
    u = fmaf(vxc - vyc, e, u);
    pd = fmaf(vxc, yc, pd);
    pd = fmaf(vyc, xc, pd);
    sx = fmaf(vxc, xc, sx);
    ty = fmaf(vyc, yc, ty);
  }
  const PairD2 dd = eval_second(a.kf, Dd);
  float g = 2.f * dd.kd * u + a.kg * pd;
  if (a.kf.family == FAM_DISTANCE && a.kf.true_distance)
    g += 2.f * dist_f1(a.norms[i]) * sx + 2.f * dist_f1(a.norms[a.m + j]) * ty;
  const float dk = a.dK[i * a.lddk + j];
  a.ddK[i * a.n + j] = g;
  a.A[i * a.n + j] = 4.f * dk * dd.kdd * u;
  a.B[i * a.n + j] = 2.f * dk * dd.kd;
}

template <bool ROWS_X>
__global__ void __launch_bounds__(128) kxy_bwd2_rows_kernel(Kxy2Args a) {
  __shared__ float red[2][4];
  const int64_t row = blockIdx.x;                        // i (X rows) or j (Y rows)
  const int64_t cnt = ROWS_X ? a.n : a.m;                // partners
  const float* self = a.Z + (ROWS_X ? row : a.m + row) * a.dpitch;
  const float* vself = ROWS_X ? (a.VX ? a.VX + row * a.d : nullptr) : (a.VY ? a.VY + row * a.d : nullptr);
  const float* vother_base = ROWS_X ? a.VY : a.VX;
  float* out = (ROWS_X ? a.gX : a.gY) + row * a.d;
  const bool dist = a.kf.family == FAM_DISTANCE && a.kf.true_distance;
  // r = sum of dK over partners, s = <V_row, z_row> (distance kernel only)
  float r = 0.f, sdot = 0.f;
  if (dist) {
    for (int64_t k = threadIdx.x; k < cnt; k += 128) r += ROWS_X ? a.dK[row * a.lddk + k] : a.dK[k * a.lddk + row];
    if (vself)
      for (int64_t c = threadIdx.x; c < a.d; c += 128) sdot = fmaf(vself[c], self[c], sdot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      r += __shfl_xor_sync(0xffffffffu, r, o);
      sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
    }
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = r;
      red[1][threadIdx.x >> 5] = sdot;
    }
    __syncthreads();
    r = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    sdot = red[1][0] + red[1][1] + red[1][2] + red[1][3];
  }
  const float nself = a.norms[ROWS_X ? row : a.m + row];
  for (int64_t c = threadIdx.x; c < a.d; c += 128) {
    const float zc = self[c];
    const float vc = vself ? vself[c] : 0.f;
    float acc = 0.f;
    for (int64_t k = 0; k < cnt; ++k) {
      const int64_t ij = ROWS_X ? row * a.n + k : k * a.n + row;
      const float Aij = a.A[ij], Bij = a.B[ij];
      const float oc = a.Z[(ROWS_X ? a.m + k : k) * a.dpitch + c];          // partner feature
      const float voc = vother_base ? vother_base[k * a.d + c] : 0.f;      // partner cotangent
      // X rows:  A (x - y) + B (vx - vy) + kg dK vy      Y rows: -A (x - y) - B (vx - vy) + kg dK vx
      //          = A (z - o) + B (v - vo) + ...                  = A (z - o) + B (v - vo) + ...   (sign folds in)
      acc = fmaf(Aij, zc - oc, acc);
      acc = fmaf(Bij, vc - voc, acc);
      if (a.kg != 0.f) acc = fmaf(a.kg * (ROWS_X ? a.dK[row * a.lddk + k] : a.dK[k * a.lddk + row]), voc, acc);
    }
    if (dist) acc += r * (4.f * dist_f2(nself) * sdot * zc + 2.f * dist_f1(nself) * vc);
    out[c] = acc;
  }
}

cudaError_t launch_kernel_xy_bwd2(const KernelFn& kf, const float* Z, const float* norms, int64_t dpitch, int64_t m,
                                  int64_t n, int64_t d, const float* dK, int64_t lddk, const float* VX, const float* VY,
                                  float* ddK, float* A, float* B, float* gX, float* gY, cudaStream_t s) {
  Kxy2Args a;
  a.kf = kf;
  a.Z = Z;
  a.norms = norms;
  a.dpitch = dpitch;
  a.m = m;
  a.n = n;
  a.d = d;
  a.dK = dK;
  a.lddk = lddk;
  a.VX = VX;
  a.VY = VY;
  a.ddK = ddK;
  a.A = A;
  a.B = B;
  a.gX = gX;
  a.gY = gY;
  a.kg = kf.family == FAM_DOT ? 1.f : (kf.family == FAM_RQ ? kf.add_dot : 0.f);
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((m + 7) / 8));
  kxy_bwd2_pair_kernel<<<grid, 256, 0, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  kxy_bwd2_rows_kernel<true><<<(unsigned)m, 128, 0, s>>>(a);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  kxy_bwd2_rows_kernel<false><<<(unsigned)n, 128, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace smmd

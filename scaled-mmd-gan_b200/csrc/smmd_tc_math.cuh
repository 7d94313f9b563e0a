// smmd_tc_math.cuh -- register-resident epilogue math for the tensor-core kernels.
//
// Every variant exposes
//     eval(S, nij, k_raw, kd_raw)      S = <z_i,z_j> from TMEM, nij = |z_i|^2 + |z_j|^2
//     k_scale(), kd_scale()            constants folded in once per tile / per coefficient
// with  k = k_scale * k_raw  (kernel value, enters the block sums) and  dk/dD = kd_scale * kd_raw.
// The epilogue is the limiter of the fused kernel at d = 256: per Gram element the
// tensor pipe needs 1/8 cycle, so every FMA-pipe instruction and above all every MUFU op counts
// (MUFU: 16/clk/SM = 8 FMA-pipe issue slots each).  Hence the specialisations:
//   MathRbf1        one sigma (the shipped yml kernel, mmd.py:55): 1 MUFU
//   MathRbfLadder   gammas in ratio 4 (sigmas {1,2,4,8,16}, BASELINE config 1): 1 MUFU + 2 FMUL / step
//   MathRq3Default  the reference's default alphas (.1, 1, 10), unit weights (mmd.py:143): ONE rcp of the
//                   product of the three bases (instead of three), u^10 by a 4-multiply chain, and a single
//                   lg2/ex2 pair for alpha = .1  -> 3 MUFU instead of 9
//   MathGeneric     any sigmas / alphas / weights (parameters read from shared memory)
#pragma once
#include "smmd_kfun.cuh"

namespace smmd {

enum TcVariant : int {
  TV_RBF1 = 0,
  TV_RBF_LADDER5,
  TV_RBF_GENERIC,
  TV_RQ3_DEFAULT,
  TV_RQ_GENERIC,
  TV_DISTANCE,
  TV_POLY3,
  TV_POLY_GENERIC,
  TV_NONE
};

struct MathRbf1 {
  float c1, w, g;
  __device__ explicit MathRbf1(const KernelFn& f, const float*) : c1(f.p1[0]), w(f.w[0]), g(-f.p0[0] * f.w[0]) {}
  __device__ __forceinline__ float k_scale() const { return w; }
  __device__ __forceinline__ float kd_scale() const { return g; }
  __device__ __forceinline__ void eval(float S, float nij, float& k, float& kd) const {
    const float D = fmaxf(fmaf(-2.f, S, nij), 0.f);
    k = fast_ex2(c1 * D);
    kd = k;
  }
};

// components sorted by increasing gamma with gamma[i+1] = 4 gamma[i]  ->  e[i+1] = e[i]^4
template <int NP>
struct MathRbfLadder {
  float c1, w[NP], g[NP];
  __device__ explicit MathRbfLadder(const KernelFn& f, const float*) : c1(f.p1[0]) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      w[i] = f.w[i];
      g[i] = -f.p0[i] * f.w[i];
    }
  }
  __device__ __forceinline__ float k_scale() const { return 1.f; }
  __device__ __forceinline__ float kd_scale() const { return 1.f; }
  __device__ __forceinline__ void eval(float S, float nij, float& k, float& kd) const {
    const float D = fmaxf(fmaf(-2.f, S, nij), 0.f);
    float e = fast_ex2(c1 * D);
    k = w[0] * e;
    kd = g[0] * e;
#pragma unroll
    for (int i = 1; i < NP; ++i) {
      e *= e;
      e *= e;
      k = fmaf(w[i], e, k);
      kd = fmaf(g[i], e, kd);
    }
  }
};

struct MathRq3Default {
  __device__ explicit MathRq3Default(const KernelFn&, const float*) {}
  __device__ __forceinline__ float k_scale() const { return 1.f; }
  __device__ __forceinline__ float kd_scale() const { return -0.5f; }
  __device__ __forceinline__ void eval(float S, float nij, float& k, float& kd) const {
    // D is capped far above any realistic squared distance so the product of the bases stays finite
    const float D = fminf(fmaxf(fmaf(-2.f, S, nij), 0.f), 1.0e10f);
    const float b1 = fmaf(D, 5.f, 1.f);    // alpha = .1 : 1 + D / (2 * .1)
    const float b2 = fmaf(D, .5f, 1.f);    // alpha = 1
    const float b3 = fmaf(D, .05f, 1.f);   // alpha = 10
    const float p23 = b2 * b3;
    const float R = fast_rcp(b1 * p23);
    const float r1 = R * p23;
    const float t = R * b1;
    const float r2 = t * b3;               // 1 / b2
    const float r3 = t * b2;               // 1 / b3
    const float e1 = fast_ex2(-0.1f * fast_lg2(b1));
    const float q2 = r3 * r3, q4 = q2 * q2, q8 = q4 * q4;
    const float e3 = q8 * q2;              // b3^-10
    k = (e1 + r2) + e3;
    kd = fmaf(e3, r3, fmaf(r2, r2, e1 * r1));   // sum_k e_k / b_k ; kd_scale = -1/2
  }
};

// parameters in shared memory: sp[0][i] = p0, sp[1][i] = p1, sp[2][i] = w
template <int FAM>
struct MathGeneric {
  const float* sp;
  int np;
  __device__ explicit MathGeneric(const KernelFn& f, const float* smem_params) : sp(smem_params), np(f.np) {}
  __device__ __forceinline__ float k_scale() const { return 1.f; }
  __device__ __forceinline__ float kd_scale() const { return 1.f; }
  __device__ __forceinline__ void eval(float S, float nij, float& k, float& kd) const {
    const float D = fmaxf(fmaf(-2.f, S, nij), 0.f);
    k = 0.f;
    kd = 0.f;
#pragma unroll 1
    for (int i = 0; i < np; ++i) {
      const float p0 = sp[i], p1 = sp[8 + i], w = sp[16 + i];
      if constexpr (FAM == FAM_RBF) {
        const float e = w * fast_ex2(p1 * D);
        k += e;
        kd = fmaf(-p0, e, kd);
      } else {
        const float base = fmaf(D, p0, 1.f);
        const float e = w * fast_ex2(-p1 * fast_lg2(base));
        k += e;
        kd = fmaf(-0.5f * e, fast_rcp(base), kd);
      }
    }
  }
};

struct MathDistance {
  __device__ explicit MathDistance(const KernelFn&, const float*) {}
  __device__ __forceinline__ float k_scale() const { return -1.f; }
  __device__ __forceinline__ float kd_scale() const { return -0.5f; }
  __device__ __forceinline__ void eval(float S, float nij, float& k, float& kd) const {
    const float t = fmaf(-2.f, S, nij) + kEps;   // D not clamped before +eps (mmd.py:12,29)
    const float rs = t > 0.f ? fast_rsqrt(t) : 0.f;
    k = t * rs;                                   // sqrt(max(t,0))
    kd = rs;
  }
};

struct MathPoly3 {
  float gamma, c0;
  __device__ explicit MathPoly3(const KernelFn& f, const float*) : gamma(f.poly_gamma), c0(f.poly_coef0) {}
  __device__ __forceinline__ float k_scale() const { return 1.f; }
  __device__ __forceinline__ float kd_scale() const { return 0.f; }
  __device__ __forceinline__ void eval(float S, float, float& k, float& kd) const {
    const float b = fmaf(gamma, S, c0);
    k = b * b * b;
    kd = 0.f;
  }
};

struct MathPolyN {
  float gamma, c0;
  int degree;
  __device__ explicit MathPolyN(const KernelFn& f, const float*) : gamma(f.poly_gamma), c0(f.poly_coef0), degree(f.degree) {}
  __device__ __forceinline__ float k_scale() const { return 1.f; }
  __device__ __forceinline__ float kd_scale() const { return 0.f; }
  __device__ __forceinline__ void eval(float S, float, float& k, float& kd) const {
    const float b = fmaf(gamma, S, c0);
    float p = b;
    for (int i = 1; i < degree; ++i) p *= b;
    k = p;
    kd = 0.f;
  }
};

// Host: pick the variant; may reorder the mixture components of `kf` (sums are order-independent).
inline TcVariant select_tc_variant(KernelFn& kf) {
  auto close = [](double a, double b) { return fabs(a - b) <= 1e-6 * fabs(b); };
  switch (kf.family) {
    case FAM_DISTANCE: return TV_DISTANCE;
    case FAM_POLY: return kf.degree == 3 ? TV_POLY3 : TV_POLY_GENERIC;
    case FAM_RBF: {
      if (kf.np == 1) return TV_RBF1;
      if (kf.np == 5) {
        // sort by increasing gamma, then check the ratio-4 ladder
        KernelFn s = kf;
        for (int i = 0; i < s.np; ++i)
          for (int j = i + 1; j < s.np; ++j)
            if (s.p0[j] < s.p0[i]) {
              std::swap(s.p0[i], s.p0[j]);
              std::swap(s.p1[i], s.p1[j]);
              std::swap(s.w[i], s.w[j]);
            }
        bool ladder = true;
        for (int i = 0; i + 1 < s.np; ++i) ladder = ladder && close(s.p0[i + 1], 4.0 * s.p0[i]);
        if (ladder) {
          kf = s;
          return TV_RBF_LADDER5;
        }
      }
      return TV_RBF_GENERIC;
    }
    case FAM_RQ: {
      if (kf.np == 3) {
        bool used[3] = {false, false, false};
        const double want[3] = {0.1, 1.0, 10.0};
        int hits = 0;
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j)
            if (!used[j] && close(kf.p1[i], want[j]) && kf.w[i] == 1.f) {
              used[j] = true;
              ++hits;
              break;
            }
        if (hits == 3) return TV_RQ3_DEFAULT;
      }
      return TV_RQ_GENERIC;
    }
    default: return TV_NONE;
  }
}

}  // namespace smmd

// smmd_kfun.cuh -- device evaluation of the pair kernels k(G, D) of gan/core/mmd.py and their partials.
//
// For one pair (a, b):  G = <a,b>,  D_raw = |a|^2 + |b|^2 - 2G.
//   value   k            -- what enters the block sums
//   kd      dk/dD_raw    -- flows into the gradient through D  (dD/da = 2(a-b))
//   kg      dk/dG direct -- add_dot / dot / poly part           (dG/da = b)
//
// Two flavours: `eval_exact` (IEEE-ish fp32 libm, used by the SIMT fp32 path; mirrors the reference's
// exp(-alpha*log(.)) formulation) and `eval_fast<FAM>` (ex2/lg2/rcp approximations, used in the
// tensor-core epilogues where operands are bf16 anyway).
#pragma once
#include "smmd_internal.h"

namespace smmd {

struct PairVal {
  float k, kd, kg;
};

// `Draw` is the squared distance.  The exact path evaluates it in difference form sum_c (a_c-b_c)^2,
// which is >= 0 by construction and free of the cancellation the reference's fp32 Gram form
// (-2*XY + |x|^2 + |y|^2, mmd.py:67) suffers for near-by points -- i.e. closer to the fp64 truth.
__device__ __forceinline__ PairVal eval_exact(const KernelFn& f, float G, float Draw, float ni, float nj) {
  PairVal r;
  r.k = 0.f;
  r.kd = 0.f;
  r.kg = 0.f;
  switch (f.family) {
    case FAM_DOT:
      r.k = G;
      r.kg = 1.f;
      break;
    case FAM_POLY: {
      float b = f.poly_gamma * G + f.poly_coef0;
      float p = 1.f;  // b^(degree-1)
      for (int i = 1; i < f.degree; ++i) p *= b;
      r.k = p * b;
      r.kg = (float)f.degree * f.poly_gamma * p;
      break;
    }
    case FAM_DISTANCE: {
      // mysqrt(x) = sqrt(max(x + eps, 0)), D NOT clamped first (mmd.py:12,29)
      float t = Draw + kEps;
      float root = sqrtf(fmaxf(t, 0.f));
      r.k = -root;
      r.kd = t > 0.f ? -0.5f / root : 0.f;
      if (f.true_distance) r.k += sqrtf(fmaxf(ni + kEps, 0.f)) + sqrtf(fmaxf(nj + kEps, 0.f));
      break;
    }
    case FAM_RBF: {
      float D = fmaxf(Draw, 0.f);
      float k = 0.f, kd = 0.f;
      for (int i = 0; i < f.np; ++i) {
        float e = f.w[i] * expf(-f.p0[i] * D);
        k += e;
        kd -= f.p0[i] * e;
      }
      r.k = k;
      r.kd = Draw > 0.f ? kd : 0.f;
      break;
    }
    case FAM_RQ: {
      float D = fmaxf(Draw, 0.f);
      float k = 0.f, kd = 0.f;
      for (int i = 0; i < f.np; ++i) {
        float base = 1.f + D * f.p0[i];          // 1 + D/(2 alpha)
        float e = f.w[i] * expf(-f.p1[i] * logf(base));
        k += e;
        kd -= 0.5f * e / base;
      }
      r.k = k;
      r.kd = Draw > 0.f ? kd : 0.f;
      if (f.add_dot > 0.f) {
        r.k += f.add_dot * G;
        r.kg = f.add_dot;
      }
      break;
    }
  }
  return r;
}

// First and second derivative of the distance part g(D) of a kernel (k = g(D) + kg*G + f(|a|^2) + f(|b|^2)),
// for the double backward of the witness (gradient penalty, gan/core/model.py:327-350).  Same clamping
// conventions as eval_exact: the reference's max(D, 0) has zero derivative on the clamped side.
struct PairD2 {
  float kd, kdd;
};
__device__ __forceinline__ PairD2 eval_second(const KernelFn& f, float Draw) {
  PairD2 r;
  r.kd = 0.f;
  r.kdd = 0.f;
  switch (f.family) {
    case FAM_DISTANCE: {
      float t = Draw + kEps;
      if (t > 0.f) {
        float root = sqrtf(t);
        r.kd = -0.5f / root;
        r.kdd = 0.25f / (root * t);
      }
      break;
    }
    case FAM_RBF: {
      if (Draw > 0.f) {
        for (int i = 0; i < f.np; ++i) {
          float e = f.w[i] * expf(-f.p0[i] * Draw);
          r.kd -= f.p0[i] * e;
          r.kdd += f.p0[i] * f.p0[i] * e;
        }
      }
      break;
    }
    case FAM_RQ: {
      if (Draw > 0.f) {
        for (int i = 0; i < f.np; ++i) {
          float base = 1.f + Draw * f.p0[i];   // 1 + D/(2 alpha), p0 = 1/(2 alpha), p1 = alpha
          float e = f.w[i] * expf(-f.p1[i] * logf(base));
          r.kd -= 0.5f * e / base;
          r.kdd += 0.5f * f.p0[i] * (f.p1[i] + 1.f) * e / (base * base);
        }
      }
      break;
    }
    default: break;   // dot: no distance part
  }
  return r;
}

// Analytic diagonal value k(a, a) (D = 0 exactly in the reference because the norms ARE the Gram diagonal).
__device__ __forceinline__ float diag_value(const KernelFn& f, float ni) {
  switch (f.family) {
    case FAM_DOT: return ni;
    case FAM_POLY: {
      float b = f.poly_gamma * ni + f.poly_coef0, p = 1.f;
      for (int i = 0; i < f.degree; ++i) p *= b;
      return p;
    }
    case FAM_DISTANCE: {
      float v = -sqrtf(kEps);
      if (f.true_distance) v += 2.f * sqrtf(fmaxf(ni + kEps, 0.f));
      return v;
    }
    case FAM_RBF: return f.const_diag;
    case FAM_RQ: return f.const_diag + f.add_dot * ni;
  }
  return 0.f;
}

// ---- fast approximations for the tensor-core epilogues -------------------------------------------
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2): one issue slot for two lanes of work ------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }

// k and dk/dD for the D-dependent part only (dot parts are handled by the caller).  D already clamped
// for rbf/rq; for distance D is the raw value.
template <int FAM>
__device__ __forceinline__ void eval_fast(const KernelFn& f, float D, float& k, float& kd) {
  if constexpr (FAM == FAM_RBF) {
    k = 0.f;
    kd = 0.f;
#pragma unroll
    for (int i = 0; i < SMMD_MAX_PARAMS; ++i) {
      if (i < f.np) {
        float e = f.w[i] * fast_ex2(f.p1[i] * D);  // p1 = -gamma*log2(e)
        k += e;
        kd = fmaf(-f.p0[i], e, kd);
      }
    }
  } else if constexpr (FAM == FAM_RQ) {
    k = 0.f;
    kd = 0.f;
#pragma unroll
    for (int i = 0; i < SMMD_MAX_PARAMS; ++i) {
      if (i < f.np) {
        float base = fmaf(D, f.p0[i], 1.f);
        float e = f.w[i] * fast_ex2(-f.p1[i] * fast_lg2(base));
        k += e;
        kd = fmaf(-0.5f * e, fast_rcp(base), kd);
      }
    }
  } else if constexpr (FAM == FAM_DISTANCE) {
    float t = D + kEps;
    float rs = t > 0.f ? fast_rsqrt(t) : 0.f;
    k = -t * rs;
    kd = -0.5f * rs;
  } else {
    k = 0.f;
    kd = 0.f;
  }
}

// final assembly of MMD^2 from the six block sums (gan/core/mmd.py:194-220), fp64
__device__ __forceinline__ double mmd2_from_sums(const KernelFn& kf, double m, double n, int biased, double sxx,
                                                 double syy, double sxy, double syx, double dgx, double dgy) {
  double a_xx, a_yy;
  const double a_xy = -1.0 / (m * n);
  if (biased) {
    a_xx = 1.0 / (m * m);
    a_yy = 1.0 / (n * n);
    return a_xx * (sxx + dgx) + a_yy * (syy + dgy) + a_xy * (sxy + syx);
  }
  a_xx = 1.0 / (m * (m - 1.0));
  a_yy = 1.0 / (n * (n - 1.0));
  double ex = 0.0, ey = 0.0;
  if (kf.has_const_diag) {  // trace := m * const_diagonal (mmd.py:209-212), whatever the true diagonal is
    ex = dgx - m * (double)kf.const_diag;
    ey = dgy - n * (double)kf.const_diag;
  }
  return a_xx * (sxx + ex) + a_yy * (syy + ey) + a_xy * (sxy + syx);
}


}  // namespace smmd

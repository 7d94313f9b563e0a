// smmd_tc_math.cuh -- register-resident epilogue math for the tensor-core kernels.
//
// Every variant exposes, on PAIRS of adjacent Gram columns (packed fp32x2 -> FFMA2/FMUL2/FADD2, one issue
// slot per two elements; MUFU ops per half),
//     eval2(S, nij, k_raw, kd_raw)     S = <z_i,z_j> from TMEM, nij = |z_i|^2 + |z_j|^2
//     k_scale(), kd_scale()            constants folded in once per tile / per coefficient
// with  k = k_scale * k_raw  (kernel value, enters the block sums) and  dk/dD = kd_scale * kd_raw.
// The epilogue is the limiter of the fused kernel at d = 256: per Gram element the tensor pipe
// needs 1/8 cycle, so every issue slot and above all every MUFU op counts (MUFU: 16/clk/SM = 8 FMA-pipe issue
// slots each).  Hence the specialisations:
//   MathRbf1        one sigma (the shipped yml kernel, mmd.py:55): 1 MUFU
//   MathRbfLadder   gammas in ratio 4 (sigmas {1,2,4,8,16}, BASELINE config 1): 1 MUFU + 2 FMUL / step
//   MathRq3Default  the reference's default alphas (.1, 1, 10), unit weights (mmd.py:143): ONE rcp of the
//                   product of the three bases (instead of three), u^10 by a 4-multiply chain, and a single
//                   lg2/ex2 pair for alpha = .1  -> 3 MUFU instead of 9
//   MathGeneric     any sigmas / alphas / weights (parameters read from shared memory)
// The lower clamp max(D,0) of the reference (mmd.py:67) is a no-op here up to fp32 rounding of D (the norms
// are computed from the same bf16 operands the MMA sees), and a slightly negative D is harmless for every
// transform below, so it is dropped; an upper cap keeps products of bases finite for absurd distances.
#pragma once
#include "sm100_ptx.cuh"
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
  TV_NULL,   // developer ablation: k = kd = S (no transform) -- measures the bare pipeline
  TV_NONE
};

__device__ __forceinline__ float2 dist2(float2 S, float2 nij) { return fma2(S, bc2(-2.f), nij); }
__device__ __forceinline__ float2 ex2_2(float2 x) { return make_float2(fast_ex2(x.x), fast_ex2(x.y)); }
__device__ __forceinline__ float2 lg2_2(float2 x) { return make_float2(fast_lg2(x.x), fast_lg2(x.y)); }
__device__ __forceinline__ float2 rcp_2(float2 x) { return make_float2(fast_rcp(x.x), fast_rcp(x.y)); }

struct MathRbf1 {
  static constexpr bool kHasEvalN = false;
  static constexpr bool kF16 = false;
  static constexpr bool kHasPt = false;
  static constexpr bool kHasSplit = false;
  float c1, w, g;
  __device__ explicit MathRbf1(const KernelFn& f, const float*) : c1(f.p1[0]), w(f.w[0]), g(-f.p0[0] * f.w[0]) {}
  __device__ __forceinline__ float k_scale() const { return w; }
  __device__ __forceinline__ float kd_scale() const { return g; }
  __device__ __forceinline__ void eval2(float2 S, float2 nij, float2& k, float2& kd) const {
    // c1 * D folded into one FFMA2: c1*(nij - 2S) = (-2 c1) S + c1 nij
    k = ex2_2(fma2(S, bc2(-2.f * c1), mul2(nij, bc2(c1))));
    kd = k;
  }
};

// components sorted by increasing gamma with gamma[i+1] = 4 gamma[i]  ->  e[i+1] = e[i]^4
template <int NP>
struct MathRbfLadder {
  static constexpr bool kHasEvalN = false;
  static constexpr bool kF16 = false;
  static constexpr bool kHasPt = false;
  static constexpr bool kHasSplit = false;
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
  __device__ __forceinline__ void eval2(float2 S, float2 nij, float2& k, float2& kd) const {
    float2 e = ex2_2(fma2(S, bc2(-2.f * c1), mul2(nij, bc2(c1))));
    k = mul2(e, bc2(w[0]));
    kd = mul2(e, bc2(g[0]));
#pragma unroll
    for (int i = 1; i < NP; ++i) {
      e = mul2(e, e);
      e = mul2(e, e);
      k = fma2(e, bc2(w[i]), k);
      kd = fma2(e, bc2(g[i]), kd);
    }
  }
};

// "Pair term" hooks (symmetric path): a variant may fold |z_i|^2 + |z_j|^2 into its first affine step,
//     rt = row_term(|z_i|^2)  once per row,   pt = pair_term(rt, |z_j|^2 pair)  one packed op per column pair,
//     evalPt(S, pt, k, kd)    which then needs no separate D = nij - 2 S.
// Variants without the hooks get rt = |z_i|^2, pt = nij and their ordinary eval.
template <class Math>
__device__ __forceinline__ float math_row_term(const Math& m, float ni) {
  if constexpr (Math::kHasPt) return m.row_term(ni);
  else return ni;
}
template <class Math>
__device__ __forceinline__ float2 math_pair_term(const Math& m, float2 rt, float2 nj) {
  if constexpr (Math::kHasPt) return m.pair_term(rt, nj);
  else return add2(rt, nj);
}

// Default interleave helper: variants without a hand-interleaved evalN fall back to eval2 per pair.
template <class Math, int NP>
__device__ __forceinline__ void eval_pairs(const Math& m, const float2 (&S)[NP], const float2 (&nij)[NP],
                                           float2 (&k)[NP], float2 (&kd)[NP]) {
  if constexpr (Math::kHasEvalN) {
    m.template evalN<NP>(S, nij, k, kd);
  } else {
#pragma unroll
    for (int i = 0; i < NP; ++i) m.eval2(S[i], nij[i], k[i], kd[i]);
  }
}
template <class Math, int NP>
__device__ __forceinline__ void eval_pairs_pt(const Math& m, const float2 (&S)[NP], const float2 (&pt)[NP],
                                              float2 (&k)[NP], float2 (&kd)[NP]) {
  if constexpr (Math::kHasPt) m.template evalPt<NP>(S, pt, k, kd);
  else eval_pairs<Math, NP>(m, S, pt, k, kd);
}

struct MathRq3Default {
  static constexpr bool kHasEvalN = true;
  static constexpr bool kF16 = false;
  static constexpr bool kHasPt = true;
#ifdef SMMD_SYM_NOSPLIT
  static constexpr bool kHasSplit = false;
#else
  static constexpr bool kHasSplit = true;
#endif
  // Split form for software pipelining: `front` runs the affine steps and all three MUFU ops of a pair, `back` the
  // FMA-only tail.  The caller issues front(g + 1) next to back(g), so a warp's instruction stream mixes MUFU and
  // FMA work at every point instead of alternating between a MUFU burst and a long FMA tail.
  struct Pair {
    float2 b1, b2, b3, p23, R, e1;
  };
  __device__ __forceinline__ void front(float2 S, float2 pt, Pair& s) const {
    s.b1 = fma2(S, bc2(-10.f), pt);
    s.b2 = fma2(s.b1, bc2(.1f), bc2(.9f));
    s.b3 = fma2(s.b1, bc2(.01f), bc2(.99f));
    const float2 L = lg2_2(s.b1);
    s.p23 = mul2(s.b2, s.b3);
    s.R = rcp_2(mul2(s.b1, s.p23));
    s.e1 = ex2_2(mul2(L, bc2(-0.1f)));
  }
  __device__ __forceinline__ void back(const Pair& s, float2& k, float2& kd) const {
    const float2 r1 = mul2(s.R, s.p23);
    const float2 t = mul2(s.R, s.b1);
    const float2 r2 = mul2(t, s.b3);
    const float2 r3 = mul2(t, s.b2);
    const float2 q2 = mul2(r3, r3), q4 = mul2(q2, q2), q8 = mul2(q4, q4);
    const float2 e3 = mul2(q8, q2);
    k = add2(add2(s.e1, r2), e3);
    kd = fma2(e3, r3, fma2(r2, r2, mul2(s.e1, r1)));
  }
  // b1 = 1 + 5 D = (1 + 5 |z_i|^2) + 5 |z_j|^2 - 10 S;  b2 = 1 + D/2 = .1 b1 + .9;  b3 = 1 + D/20 = .01 b1 + .99
  __device__ __forceinline__ float row_term(float ni) const { return fmaf(5.f, ni, 1.f); }
  __device__ __forceinline__ float2 pair_term(float2 rt, float2 nj) const { return fma2(nj, bc2(5.f), rt); }
  // No upper clamp here: the product of the bases stays finite up to D ~ 1e12 and beyond that rcp(inf) = 0 zeroes
  // every term; only |z|^2 > 1e18 could produce 0 * inf, and such rows are reported as non-finite by the caller.
  template <int NP>
  __device__ __forceinline__ void evalPt(const float2 (&S)[NP], const float2 (&pt)[NP], float2 (&k)[NP],
                                         float2 (&kd)[NP]) const {
    float2 b1[NP], b2[NP], b3[NP], p23[NP], R[NP], L[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      b1[i] = fma2(S[i], bc2(-10.f), pt[i]);
      b2[i] = fma2(b1[i], bc2(.1f), bc2(.9f));
      b3[i] = fma2(b1[i], bc2(.01f), bc2(.99f));
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) L[i] = lg2_2(b1[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      p23[i] = mul2(b2[i], b3[i]);
      R[i] = mul2(b1[i], p23[i]);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) R[i] = rcp_2(R[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) L[i] = mul2(L[i], bc2(-0.1f));
#pragma unroll
    for (int i = 0; i < NP; ++i) L[i] = ex2_2(L[i]);           // e1 = b1^-0.1
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const float2 r1 = mul2(R[i], p23[i]);
      const float2 t = mul2(R[i], b1[i]);
      const float2 r2 = mul2(t, b3[i]);
      const float2 r3 = mul2(t, b2[i]);
      const float2 q2 = mul2(r3, r3), q4 = mul2(q2, q2), q8 = mul2(q4, q4);
      const float2 e3 = mul2(q8, q2);
      k[i] = add2(add2(L[i], r2), e3);
      kd[i] = fma2(e3, r3, fma2(r2, r2, mul2(L[i], r1)));
    }
  }
  __device__ explicit MathRq3Default(const KernelFn&, const float*) {}
  __device__ __forceinline__ float k_scale() const { return 1.f; }
  __device__ __forceinline__ float kd_scale() const { return -0.5f; }
  // NP pairs in lock-step: every stage is issued for all pairs before the next dependent stage, so the
  // FMA pipe works on pair i+1..i+3 while the MUFU results of pair i are in flight (the per-pair chain is
  // FMA -> rcp/lg2 -> FMA -> ex2 -> FMA, three MUFU latencies deep).
  template <int NP>
  __device__ __forceinline__ void evalN(const float2 (&S)[NP], const float2 (&nij)[NP], float2 (&k)[NP],
                                        float2 (&kd)[NP]) const {
    float2 b1[NP], b2[NP], b3[NP], p23[NP], R[NP], L[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      float2 D = dist2(S[i], nij[i]);
      D = make_float2(fminf(D.x, 1.0e10f), fminf(D.y, 1.0e10f));
      b1[i] = fma2(D, bc2(5.f), bc2(1.f));
      b2[i] = fma2(D, bc2(.5f), bc2(1.f));
      b3[i] = fma2(D, bc2(.05f), bc2(1.f));
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) L[i] = lg2_2(b1[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      p23[i] = mul2(b2[i], b3[i]);
      R[i] = mul2(b1[i], p23[i]);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) R[i] = rcp_2(R[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) L[i] = mul2(L[i], bc2(-0.1f));
#pragma unroll
    for (int i = 0; i < NP; ++i) L[i] = ex2_2(L[i]);           // e1 = b1^-0.1
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const float2 r1 = mul2(R[i], p23[i]);
      const float2 t = mul2(R[i], b1[i]);
      const float2 r2 = mul2(t, b3[i]);
      const float2 r3 = mul2(t, b2[i]);
      const float2 q2 = mul2(r3, r3), q4 = mul2(q2, q2), q8 = mul2(q4, q4);
      const float2 e3 = mul2(q8, q2);
      k[i] = add2(add2(L[i], r2), e3);
      kd[i] = fma2(e3, r3, fma2(r2, r2, mul2(L[i], r1)));
    }
  }
  __device__ __forceinline__ void eval2(float2 S, float2 nij, float2& k, float2& kd) const {
    float2 D = dist2(S, nij);
    D = make_float2(fminf(D.x, 1.0e10f), fminf(D.y, 1.0e10f));  // keeps b1*b2*b3 finite
    const float2 b1 = fma2(D, bc2(5.f), bc2(1.f));     // alpha = .1 : 1 + D / (2 * .1)
    const float2 b2 = fma2(D, bc2(.5f), bc2(1.f));     // alpha = 1
    const float2 b3 = fma2(D, bc2(.05f), bc2(1.f));    // alpha = 10
    const float2 p23 = mul2(b2, b3);
    const float2 R = rcp_2(mul2(b1, p23));
    const float2 r1 = mul2(R, p23);
    const float2 t = mul2(R, b1);
    const float2 r2 = mul2(t, b3);                     // 1 / b2
    const float2 r3 = mul2(t, b2);                     // 1 / b3
    const float2 e1 = ex2_2(mul2(lg2_2(b1), bc2(-0.1f)));
    const float2 q2 = mul2(r3, r3), q4 = mul2(q2, q2), q8 = mul2(q4, q4);
    const float2 e3 = mul2(q8, q2);                    // b3^-10
    k = add2(add2(e1, r2), e3);
    kd = fma2(e3, r3, fma2(r2, r2, mul2(e1, r1)));     // sum_k e_k / b_k ; kd_scale = -1/2
  }
};

// parameters in shared memory: sp[0..7] = p0, sp[8..15] = p1, sp[16..23] = w
template <int FAM>
struct MathGeneric {
  static constexpr bool kHasEvalN = false;
  static constexpr bool kF16 = false;
  static constexpr bool kHasPt = false;
  static constexpr bool kHasSplit = false;
  const float* sp;
  int np;
  __device__ explicit MathGeneric(const KernelFn& f, const float* smem_params) : sp(smem_params), np(f.np) {}
  __device__ __forceinline__ float k_scale() const { return 1.f; }
  __device__ __forceinline__ float kd_scale() const { return 1.f; }
  __device__ __forceinline__ void eval2(float2 S, float2 nij, float2& k, float2& kd) const {
    float2 D = dist2(S, nij);
    D = make_float2(fminf(fmaxf(D.x, 0.f), 1.0e30f), fminf(fmaxf(D.y, 0.f), 1.0e30f));
    k = bc2(0.f);
    kd = bc2(0.f);
#pragma unroll 1
    for (int i = 0; i < np; ++i) {
      const float p0 = sp[i], p1 = sp[8 + i], w = sp[16 + i];
      if constexpr (FAM == FAM_RBF) {
        const float2 e = mul2(ex2_2(mul2(D, bc2(p1))), bc2(w));
        k = add2(k, e);
        kd = fma2(e, bc2(-p0), kd);
      } else {
        const float2 base = fma2(D, bc2(p0), bc2(1.f));
        const float2 e = mul2(ex2_2(mul2(lg2_2(base), bc2(-p1))), bc2(w));
        k = add2(k, e);
        kd = fma2(mul2(e, bc2(-0.5f)), rcp_2(base), kd);
      }
    }
  }
};

struct MathDistance {
  static constexpr bool kHasEvalN = false;
  static constexpr bool kF16 = false;
  static constexpr bool kHasPt = false;
  static constexpr bool kHasSplit = false;
  __device__ explicit MathDistance(const KernelFn&, const float*) {}
  __device__ __forceinline__ float k_scale() const { return -1.f; }
  __device__ __forceinline__ float kd_scale() const { return -0.5f; }
  __device__ __forceinline__ void eval2(float2 S, float2 nij, float2& k, float2& kd) const {
    const float2 t = add2(dist2(S, nij), bc2(kEps));   // D not clamped before +eps (mmd.py:12,29)
    kd = make_float2(t.x > 0.f ? fast_rsqrt(t.x) : 0.f, t.y > 0.f ? fast_rsqrt(t.y) : 0.f);
    k = mul2(t, kd);                                   // sqrt(max(t,0))
  }
};

struct MathNull {
  static constexpr bool kHasEvalN = false;
  static constexpr bool kF16 = false;
  static constexpr bool kHasPt = false;
  static constexpr bool kHasSplit = false;
  __device__ explicit MathNull(const KernelFn&, const float*) {}
  __device__ __forceinline__ float k_scale() const { return 1.f; }
  __device__ __forceinline__ float kd_scale() const { return 1.f; }
  __device__ __forceinline__ void eval2(float2 S, float2, float2& k, float2& kd) const {
    k = S;
    kd = S;
  }
};

struct MathPoly3 {
  static constexpr bool kHasEvalN = false;
  static constexpr bool kF16 = false;
  static constexpr bool kHasPt = false;
  static constexpr bool kHasSplit = false;
  float gamma, c0;
  __device__ explicit MathPoly3(const KernelFn& f, const float*) : gamma(f.poly_gamma), c0(f.poly_coef0) {}
  __device__ __forceinline__ float k_scale() const { return 1.f; }
  __device__ __forceinline__ float kd_scale() const { return 0.f; }
  __device__ __forceinline__ void eval2(float2 S, float2, float2& k, float2& kd) const {
    const float2 b = fma2(S, bc2(gamma), bc2(c0));
    k = mul2(mul2(b, b), b);
    kd = bc2(0.f);
  }
};

struct MathPolyN {
  static constexpr bool kHasEvalN = false;
  static constexpr bool kF16 = false;
  static constexpr bool kHasPt = false;
  static constexpr bool kHasSplit = false;
  float gamma, c0;
  int degree;
  __device__ explicit MathPolyN(const KernelFn& f, const float*) : gamma(f.poly_gamma), c0(f.poly_coef0), degree(f.degree) {}
  __device__ __forceinline__ float k_scale() const { return 1.f; }
  __device__ __forceinline__ float kd_scale() const { return 0.f; }
  __device__ __forceinline__ void eval2(float2 S, float2, float2& k, float2& kd) const {
    const float2 b = fma2(S, bc2(gamma), bc2(c0));
    float2 p = b;
    for (int i = 1; i < degree; ++i) p = mul2(p, b);
    k = p;
    kd = bc2(0.f);
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

// ---- fp16 operand tier (SMMD_PREC_FP16): the same epilogue math, W packed to IEEE half instead of bf16 ------------------
// F16Of<Math> only flips the compile-time flag the kernels read (operand format of the UMMA descriptors, W packing,
// unpacking of the staged W block); the kernels stay templated on one Math parameter.
template <class Base>
struct F16Of : Base {
  static constexpr bool kF16 = true;
  using Base::Base;
};
template <bool F16>
__device__ __forceinline__ uint32_t pack_w(float lo, float hi) {
  if constexpr (F16) return sm100::pack_f16x2(lo, hi);
  else return sm100::pack_bf16x2(lo, hi);
}
template <bool F16>
__device__ __forceinline__ float2 unpack_w(uint32_t w) {
  if constexpr (F16) return make_float2(sm100::f16_lo_to_f32(w), sm100::f16_hi_to_f32(w));
  else return make_float2(sm100::bf16_lo_to_f32(w), sm100::bf16_hi_to_f32(w));
}
template <class Math>
__host__ __device__ constexpr uint32_t operand_fmt() { return Math::kF16 ? 0u : 1u; }   // kFmtF16 : kFmtBF16

}  // namespace smmd

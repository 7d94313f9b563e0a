// smmd_tc_common.cuh -- pieces shared by the tensor-core translation units (smmd_tc.cu: operand preparation and
// dispatch; smmd_tc_fused.cu: fused fwd+bwd kernel; smmd_tc_wz.cu: two-pass path for wide features;
// smmd_tc_gram.cu: Gram + reduction kernels for KID / 3-sample sums / value-only MMD^2).
#pragma once
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include "sm100_ptx.cuh"
#include "smmd_kfun.cuh"
#include "smmd_tc.h"
#include "smmd_tc_math.cuh"
#include "tmap_host.h"

namespace smmd {
namespace tc {
using namespace sm100;

constexpr int BM = 128;            // rows per row block (UMMA M)
constexpr int BNF = 64;            // fused kernel: columns per tile
// Kernels with 16 epilogue warps + producer + issuer launch 20 warps (two idle ones complete the fifth warpgroup) at 96
// registers and re-split them with setmaxnreg: 16 x kEpiRegs + 4 x kCtlRegs = 20 x 96 (the CTA's pool is what it
// launched with; asking for more than that fails the launch).
constexpr int kRoleThreads = 640;
constexpr uint32_t kEpiRegs = 104, kCtlRegs = 64;
constexpr int kThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 two epilogue groups
constexpr int kMaxSmem = 232448;   // 227 KB

inline int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }

// Path-selection options.  Defaults are the measured best; every value is validated (an out-of-range request is
// clamped to what the kernels support -- never a silently corrupt configuration).  They can be changed per process
// through smmd_set_option() (tests use it to force a path on small shapes) and, for convenience in the developer
// harness, through SMMD_<NAME> environment variables read once at first use.  Knobs that produce meaningless results
// (null epilogue math, single-pass timing) exist only in developer builds (-DSMMD_DEV_KNOBS, `make DEV=1`).
struct Tuning {
  int fused_ksplit;     // epilogue column slices per tile of the single-tile fused kernel (1 or 2)
  int null_math;        // DEV: results are meaningless
  int fused_pair;       // tile-pair variant of the fused kernel (default 1)
  int fused_lockstep;   // whole row blocks per CTA for large Z (default 1)
  int wz_min_d;         // features above which the fused backward switches to the two-pass path; clamped to [0, 256]
                        // (the fused kernel keeps O[128 x d] in tensor memory: d <= 256)
  int64_t wz_panel_bytes;   // byte budget of one W row panel (>= 1 MB)
  int wz_pair;          // pass 1 as CTA pairs with cta_group::2 UMMAs (default 1)
  int sym;              // symmetric two-pass path for whole problems (default 1)
  int symf;             // d <= 256: fused variant of the symmetric path (direct products inside pass 1; default 1)
  int sym_only;         // DEV timing knob: 1 = pass 1 only, 2 = pass 2 only; results are meaningless
  int64_t sym_min_rows;      // stacked rows from which the symmetric paths are used (0 = the measured default)
  int64_t symf_min_rows;     // stacked rows from which d <= 256 takes the fused variant (0 = the measured default)
  int64_t sym_max_w_bytes;   // largest W (Mp x Mp bf16) the symmetric path may place in the workspace
  int disable_small;    // exact path: skip the one-launch small-problem kernel (tests: force the general kernels)
};
inline int64_t clamp_i64(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }
// returns false for an unknown name
inline bool tuning_set(Tuning& v, const char* name, int64_t x) {
  if (!strcmp(name, "fused_ksplit")) v.fused_ksplit = x == 1 ? 1 : 2;
  else if (!strcmp(name, "fused_pair")) v.fused_pair = x != 0;
  else if (!strcmp(name, "fused_lockstep")) v.fused_lockstep = x != 0;
  else if (!strcmp(name, "wz_min_d")) v.wz_min_d = (int)clamp_i64(x, 0, 256);
  else if (!strcmp(name, "wz_panel_mb")) v.wz_panel_bytes = clamp_i64(x, 1, (int64_t)1 << 20) << 20;
  else if (!strcmp(name, "wz_pair")) v.wz_pair = x != 0;
  else if (!strcmp(name, "sym")) v.sym = x != 0;
  else if (!strcmp(name, "symf")) v.symf = x != 0;
  else if (!strcmp(name, "sym_min_rows")) v.sym_min_rows = clamp_i64(x, 0, (int64_t)1 << 40);
  else if (!strcmp(name, "symf_min_rows")) v.symf_min_rows = clamp_i64(x, 0, (int64_t)1 << 40);
  else if (!strcmp(name, "sym_max_w_mb")) v.sym_max_w_bytes = clamp_i64(x, 0, (int64_t)1 << 24) << 20;
  else if (!strcmp(name, "disable_small")) v.disable_small = x != 0;
#ifdef SMMD_DEV_KNOBS
  else if (!strcmp(name, "debug_nullmath")) v.null_math = x != 0;
  else if (!strcmp(name, "sym_only")) v.sym_only = (int)clamp_i64(x, 0, 2);
#endif
  else return false;
  return true;
}
inline Tuning& tuning_mut() {
  static Tuning t = [] {
    Tuning v;
    v.fused_ksplit = 2;
    v.null_math = 0;
    v.fused_pair = 1;
    v.fused_lockstep = 1;
    v.wz_min_d = 256;
    v.wz_pair = 1;
    v.wz_panel_bytes = (int64_t)6 << 30;
    v.sym = 1;
    v.symf = 1;
    v.sym_only = 0;
    v.sym_min_rows = 0;
    v.symf_min_rows = 0;
    v.sym_max_w_bytes = (int64_t)48 << 30;
    v.disable_small = 0;
    static const char* const names[] = {"fused_ksplit", "fused_pair", "fused_lockstep", "wz_min_d", "wz_panel_mb", "wz_pair",
                                        "sym", "symf", "sym_min_rows", "symf_min_rows", "sym_max_w_mb", "disable_small", "debug_nullmath", "sym_only"};
    for (const char* nm : names) {
      char env[64] = "SMMD_";
      size_t k = 5;
      for (const char* c = nm; *c && k + 1 < sizeof(env); ++c) env[k++] = (char)(*c >= 'a' && *c <= 'z' ? *c - 32 : *c);
      env[k] = 0;
      if (const char* e = getenv(env)) tuning_set(v, nm, atoll(e));
    }
    return v;
  }();
  return t;
}
inline const Tuning& tuning() { return tuning_mut(); }

inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

// ------------------------------------------------------------------------------------------------
// prep: fp32/bf16 rows -> padded bf16 operand matrix (+ optional lo part), squared norms of exactly the
// values the tensor core will see, optional gather (KID subsets), optional tanh, stats initialisation.
// Layout per problem b: rows [0,mp) = X (valid < m), rows [mp, mp+np) = Y (valid < n); pad rows are zero.
// ------------------------------------------------------------------------------------------------
struct PrepTcArgs {
  const void* A;
  const void* B;
  int dtype;
  int64_t lda, ldb, m, n, mp, np, d, dp, dpz;
  const int32_t* idxA;
  const int32_t* idxB;
  int64_t first_batch;
  int tanh_features, split;
  __nv_bfloat16* Z;
  float* norms;
  double* stats;  // optional [batch][m+n][RS_COUNT]: zeroed, RS_DIAG set analytically
  KernelFn kf;
  int64_t blk_a, blk_b;  // gathered block layout (0 = plain), see SrcLayout
  int f16;               // write IEEE half instead of bf16 (fp16 operand tier; not with `split`)
  PeerSrc peer;          // peer.on: block r of A / B is pulled from rank r's exchange buffer once its flag is up
};
// (SrcLayout -> PrepTcArgs: the peer description rides along when the call came through smmd_mmd2_fwd_bwd_peers)
inline void prep_set_peers(PrepTcArgs& pa, const SrcLayout& src) {
  if (src.peers) pa.peer = *src.peers;
  else pa.peer.on = 0;
}

// launch wrappers (kernels live in smmd_tc.cu): grid = (ceil(rows / 8), batch)
cudaError_t launch_prep_tc(const PrepTcArgs& a, int64_t rows, unsigned batch, cudaStream_t s);
cudaError_t launch_colsum_tc(const __nv_bfloat16* Z, int64_t dpz, int64_t dp, int64_t m, int64_t mp, int64_t n,
                             double* csum, int f16, cudaStream_t s);

// copy the mixture parameters to shared memory for the generic math variants: sp[0..7]=p0, [8..15]=p1, [16..23]=w
__device__ __forceinline__ void stage_params(const KernelFn& kf, float* sp) {
  if (threadIdx.x < 8) {
    sp[threadIdx.x] = kf.p0[threadIdx.x];
    sp[8 + threadIdx.x] = kf.p1[threadIdx.x];
    sp[16 + threadIdx.x] = kf.w[threadIdx.x];
  }
}

// 16 columns of one row of a fused-epilogue tile: kernel transform, tile sum, row sum of W, and W packed to
// bf16x2.  SPECIAL tiles (diagonal inside / padded columns) mask per element; interior tiles run the
// unmasked instruction stream.  Kept small and called from a ROLLED loop: the whole hot loop must fit the
// instruction caches (a fully unrolled 64-column epilogue stalled ~50% on instruction fetch).
template <class Math, bool SPECIAL>
__device__ __forceinline__ void fused_chunk16(const Math& math, const uint32_t (&v)[16], const float* __restrict__ nj,
                                              float ni, float2 cw, int col0, int lim, int gi, float2& tsum,
                                              float2& rsum, uint32_t (&wpk)[8]) {
  const float2 ni2 = bc2(ni);
#pragma unroll
  for (int c = 0; c < 16; c += 8) {   // 4 pairs (8 columns) evaluated in lock-step
    const float4 na = *reinterpret_cast<const float4*>(nj + c);
    const float4 nb = *reinterpret_cast<const float4*>(nj + c + 4);
    float2 S[4], nij[4], k[4], kd[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) S[e] = make_float2(__uint_as_float(v[c + 2 * e]), __uint_as_float(v[c + 2 * e + 1]));
    nij[0] = add2(ni2, make_float2(na.x, na.y));
    nij[1] = add2(ni2, make_float2(na.z, na.w));
    nij[2] = add2(ni2, make_float2(nb.x, nb.y));
    nij[3] = add2(ni2, make_float2(nb.z, nb.w));
    eval_pairs<Math, 4>(math, S, nij, k, kd);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (SPECIAL) {
        const int col = col0 + c + 2 * e;
        const bool ok0 = (col < lim) && (col != gi), ok1 = (col + 1 < lim) && (col + 1 != gi);
        k[e] = make_float2(ok0 ? k[e].x : 0.f, ok1 ? k[e].y : 0.f);
        kd[e] = make_float2(ok0 ? kd[e].x : 0.f, ok1 ? kd[e].y : 0.f);
      }
      tsum = add2(tsum, k[e]);
      const float2 ww = mul2(kd[e], cw);
      rsum = add2(rsum, ww);
      wpk[(c >> 1) + e] = pack_w<Math::kF16>(ww.x, ww.y);
    }
  }
}

// bound of |dk/dD| over D >= 0 (sizes the fixed-point accumulator of the row sums of W and the fp16 scale of W)
inline double kd_abs_bound(const KernelFn& kf) {
  double b = 0.0;
  switch (kf.family) {
    case FAM_RBF:
      for (int i = 0; i < kf.np; ++i) b += fabs((double)kf.w[i]) * (double)kf.p0[i];
      break;
    case FAM_RQ:
      for (int i = 0; i < kf.np; ++i) b += 0.5 * fabs((double)kf.w[i]);
      break;
    case FAM_DISTANCE: b = 0.5 / sqrt(1.0e-7); break;
    default: b = 1.0;
  }
  return b > 0.0 ? b : 1.0;
}
// fp16 operand tier: W = 4 a k'(D) is ~1/N^2 and would underflow IEEE half, so it is carried times a power of two chosen
// from its bound (largest |W| lands near 2^12; half keeps 11 significant bits down to 2^-14, i.e. over 26 binades below
// the largest weight) and the gradient rows are multiplied back in the finalize kernels.  bf16 tier: scale 1.
inline double w_scale_for(const Coefs& c, const KernelFn& kf) {
  if (!c.f16) return 1.0;
  const double cmax = 4.0 * std::max(std::max(fabs(c.a_xx), fabs(c.a_yy)), fabs(c.a_xy));
  int ex = 0;
  frexp(cmax * kd_abs_bound(kf), &ex);   // bound < 2^ex
  return ldexp(1.0, std::max(-120, std::min(120, 12 - ex)));
}

// ---- per-path entry points (one translation unit each); `variant` = select_tc_variant(kf) ----
cudaError_t tc_run_fused(const KernelFn& kf, TcVariant variant, const Geometry& g, const Coefs& c, const SrcLayout& src, double* scalars,
                         float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches, const char** path);
size_t tc_fused_workspace_bytes(const Geometry& g);
cudaError_t tc_run_wz(const KernelFn& kf, TcVariant variant, const Geometry& g, const Coefs& c, const SrcLayout& src, double* scalars,
                      float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches, const char** path);
size_t tc_wz_workspace_bytes(const Geometry& g);
// symmetric two-pass path (smmd_tc_sym.cu): whole problem on this GPU, every unordered pair evaluated once
bool tc_sym_eligible(const Geometry& g);
cudaError_t tc_run_sym(const KernelFn& kf, TcVariant variant, const Geometry& g, const Coefs& c, const SrcLayout& src, double* scalars,
                       float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches, const char** path);
size_t tc_sym_workspace_bytes(const Geometry& g);
cudaError_t tc_run_value_only(const KernelFn& kf, TcVariant variant, const Geometry& g, const Coefs& c, const SrcLayout& src, int precision, double* scalars,
                              float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches, const char** path);
size_t tc_value_only_workspace_bytes(int64_t m, int64_t n, int64_t d, int precision);

}  // namespace tc
}  // namespace smmd

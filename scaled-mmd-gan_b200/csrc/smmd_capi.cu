// smmd_capi.cu -- the extern "C" boundary declared in include/smmd.h: argument validation, path
// selection (exact-fp32 SIMT vs tcgen05 tensor-core kernels), workspace carving and launches.
// No CPU fallback exists: unsupported devices get SMMD_EARCH.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "smmd_internal.h"
#include "smmd_tc.h"

using namespace smmd;

namespace {
thread_local char g_cuda_err[256] = "";
thread_local int g_launches = 0;
thread_local const char* g_path = "none";

thread_local cudaEvent_t g_pull_event = nullptr;   // smmd_peer_set_pull_event
thread_local int g_in_peer_call = 0;
thread_local int g_prof_on = 0;
thread_local int g_prof_recorded = 0;
thread_local cudaEvent_t g_prof_ev[3] = {nullptr, nullptr, nullptr};   // begin, end, mid (two-kernel paths)
thread_local int g_prof_mid = 0;

int cuda_fail(cudaError_t e) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return SMMD_ECUDA;
}
#define SMMD_CUDA(x)                        \
  do {                                      \
    cudaError_t e__ = (x);                  \
    if (e__ != cudaSuccess) return cuda_fail(e__); \
    ++g_launches;                           \
  } while (0)

int device_ok() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

// Translate the public problem into the device-side kernel description.  Returns smmd_status.
int build_kernel_fn(int kernel_id, int nparams, const float* params, const float* wts, float add_dot, int degree,
                    int64_t d, KernelFn* out) {
  KernelFn k;
  memset(&k, 0, sizeof(k));
  k.np = 0;
  k.degree = 0;
  switch (kernel_id) {
    case SMMD_K_TANH_DISTANCE: k.tanh_features = 1;  // fallthrough
    case SMMD_K_DISTANCE: k.family = FAM_DISTANCE; break;
    case SMMD_K_DOT: k.family = FAM_DOT; break;
    case SMMD_K_RBF:
    case SMMD_K_MIX_RBF: {
      k.family = FAM_RBF;
      int np = kernel_id == SMMD_K_RBF ? 1 : nparams;
      if (np < 1 || np > SMMD_MAX_PARAMS) return SMMD_EINVAL;
      k.np = np;
      double cd = 0;
      for (int i = 0; i < np; ++i) {
        if (!(params[i] > 0.f) || !std::isfinite(params[i]) || !std::isfinite(wts[i])) return SMMD_EINVAL;
        double gamma = 1.0 / (2.0 * (double)params[i] * (double)params[i]);
        k.p0[i] = (float)gamma;
        k.p1[i] = (float)(-gamma * 1.4426950408889634);
        k.w[i] = wts[i];
        cd += wts[i];
      }
      k.const_diag = (float)cd;
      k.has_const_diag = 1;
      break;
    }
    case SMMD_K_TANH_MIX_RQ: k.tanh_features = 1;  // fallthrough
    case SMMD_K_MIX_RQ: {
      k.family = FAM_RQ;
      if (nparams < 1 || nparams > SMMD_MAX_PARAMS) return SMMD_EINVAL;
      k.np = nparams;
      double cd = 0;
      for (int i = 0; i < nparams; ++i) {
        if (!(params[i] > 0.f) || !std::isfinite(params[i]) || !std::isfinite(wts[i])) return SMMD_EINVAL;
        k.p0[i] = (float)(1.0 / (2.0 * (double)params[i]));
        k.p1[i] = params[i];
        k.w[i] = wts[i];
        cd += wts[i];
      }
      if (!std::isfinite(add_dot) || add_dot < 0.f) return SMMD_EINVAL;
      k.add_dot = add_dot;
      k.const_diag = (float)cd;  // quirk kept: add_dot is NOT part of const_diagonal (mmd.py:186-188)
      k.has_const_diag = 1;
      break;
    }
    case SMMD_K_POLY: {
      k.family = FAM_POLY;
      if (degree < 1 || degree > 8) return SMMD_EINVAL;
      k.degree = degree;
      k.poly_gamma = (nparams >= 1 && params[0] > 0.f) ? params[0] : (float)(1.0 / (double)d);
      k.poly_coef0 = nparams >= 2 ? params[1] : 1.f;
      break;
    }
    default: return SMMD_EINVAL;
  }
  *out = k;
  return SMMD_OK;
}

int validate_problem(const smmd_problem* p) {
  if (!p) return SMMD_EINVAL;
  if (p->m < 1 || p->n < 1 || p->d < 1) return SMMD_ESHAPE;
  if (!p->biased && (p->m < 2 || p->n < 2)) return SMMD_ESHAPE;
  if (p->ldx < p->d || p->ldy < p->d) return SMMD_ESHAPE;
  if (p->m + p->n > (int64_t)1 << 24 || p->d > 1 << 16) return SMMD_ESHAPE;
  if (p->dtype != SMMD_F32 && p->dtype != SMMD_BF16) return SMMD_EDTYPE;
  if (p->world < 1 || p->rank < 0 || p->rank >= p->world) return SMMD_EINVAL;
  if (p->precision < SMMD_PREC_FP32 || p->precision > SMMD_PREC_FP16) return SMMD_EINVAL;
  if (p->kernel_id < 0 || p->kernel_id > SMMD_K_POLY) return SMMD_EINVAL;
  return SMMD_OK;
}

Geometry make_geometry(const smmd_problem* p) {
  Geometry g;
  g.m = p->m;
  g.n = p->n;
  g.d = p->d;
  g.x0 = p->m * p->rank / p->world;
  g.x1 = p->m * (p->rank + 1) / p->world;
  g.y0 = p->n * p->rank / p->world;
  g.y1 = p->n * (p->rank + 1) / p->world;
  g.biased = p->biased;
  return g;
}

// AUTO: tensor cores only pay off once the Gram is big and the contraction deep enough; combinations the
// tensor-core kernels do not cover stay on the exact path.  An explicit BF16/BF16X3 request is honoured or
// refused (SMMD_EUNSUPPORTED), never silently downgraded.
constexpr int64_t kExactGradMaxD = 2048;

int resolve_precision(const smmd_problem* p, int want_grad) {
  int prec = p->precision;
  if (prec == SMMD_PREC_AUTO) {
    KernelFn kf;
    // the exact gradient kernel keeps a row's features in registers: d <= 2048; wider rows take the tensor-core path
    const bool big = ((p->m + p->n) >= 1024 && p->d >= 32) || (want_grad && p->d > kExactGradMaxD);
    const bool ok = big &&
                    build_kernel_fn(p->kernel_id, p->nparams, p->params, p->wts, p->add_dot, p->degree, p->d, &kf) ==
                        SMMD_OK &&
                    tc_mmd2_covers(kf, make_geometry(p), want_grad);
    prec = ok ? SMMD_PREC_BF16 : SMMD_PREC_FP32;
  }
  return prec;
}

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
}  // namespace

namespace smmd {
void peer_after_pull(cudaStream_t s) {
  if (g_in_peer_call && g_pull_event) cudaEventRecord(g_pull_event, s);
}
void prof_begin(cudaStream_t s) {
  if (!g_prof_on) return;
  if (!g_prof_ev[0]) {
    cudaEventCreate(&g_prof_ev[0]);
    cudaEventCreate(&g_prof_ev[1]);
    cudaEventCreate(&g_prof_ev[2]);
  }
  g_prof_mid = 0;
  cudaEventRecord(g_prof_ev[0], s);
}
void prof_mark(cudaStream_t s) {
  if (!g_prof_on || !g_prof_ev[0]) return;
  cudaEventRecord(g_prof_ev[2], s);
  g_prof_mid = 1;
}
void prof_end(cudaStream_t s) {
  if (!g_prof_on || !g_prof_ev[0]) return;
  cudaEventRecord(g_prof_ev[1], s);
  g_prof_recorded = 1;
}
}  // namespace smmd

extern "C" {

void smmd_profile_enable(int on) {
  g_prof_on = on ? 1 : 0;
  g_prof_recorded = 0;
}
float smmd_profile_last_ms(void) {
  if (!g_prof_recorded) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(g_prof_ev[1]) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, g_prof_ev[0], g_prof_ev[1]) != cudaSuccess) return -1.f;
  return ms;
}

int smmd_profile_last_split_ms(float* first_ms, float* second_ms) {
  if (!g_prof_recorded || !g_prof_mid || !first_ms || !second_ms) return SMMD_EINVAL;
  if (cudaEventSynchronize(g_prof_ev[1]) != cudaSuccess) return SMMD_ECUDA;
  if (cudaEventElapsedTime(first_ms, g_prof_ev[0], g_prof_ev[2]) != cudaSuccess) return SMMD_ECUDA;
  if (cudaEventElapsedTime(second_ms, g_prof_ev[2], g_prof_ev[1]) != cudaSuccess) return SMMD_ECUDA;
  return SMMD_OK;
}

int smmd_version(void) { return SMMD_VERSION; }

const char* smmd_strerror(int status) {
  switch (status) {
    case SMMD_OK: return "ok";
    case SMMD_EINVAL: return "invalid argument";
    case SMMD_ESHAPE: return "invalid shape or stride";
    case SMMD_EDTYPE: return "unsupported dtype";
    case SMMD_EARCH: return "device is not sm_100 (B200): no fallback path exists";
    case SMMD_EWORKSPACE: return "workspace missing, misaligned or too small";
    case SMMD_ECUDA: return "CUDA call failed";
    case SMMD_EUNSUPPORTED: return "combination not supported on the requested precision path";
    default: return "unknown status";
  }
}

const char* smmd_last_cuda_error(void) { return g_cuda_err; }
int smmd_device_supported(void) { return device_ok(); }
int smmd_set_option(const char* name, long long value) { return tc_set_option(name, value) ? SMMD_OK : SMMD_EINVAL; }
int smmd_last_launch_count(void) { return g_launches; }
const char* smmd_last_path(void) { return g_path; }

size_t smmd_mmd2_workspace_bytes(const smmd_problem* p, int want_grad) {
  if (validate_problem(p) != SMMD_OK) return 0;
  const int prec = resolve_precision(p, want_grad);
  if (prec == SMMD_PREC_FP32) return simt_plan(p->m, p->n, p->d, 1).off_end;
  return tc_mmd2_workspace_bytes(make_geometry(p), want_grad, prec);   // exact for this rank's row range
}

static int mmd2_fwd_bwd_impl(const smmd_problem* p, const SrcLayout& src, double* scalars, float* dX, float* dY,
                             void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  g_path = "none";
  int st = validate_problem(p);
  if (st != SMMD_OK) return st;
  if (!src.X || !src.Y || !scalars) return SMMD_EINVAL;
  if ((dX == nullptr) != (dY == nullptr)) return SMMD_EINVAL;
  if (p->kernel_id == SMMD_K_POLY && dX) return SMMD_EUNSUPPORTED;
  if (!device_ok()) return SMMD_EARCH;
  const int want_grad = dX != nullptr;
  KernelFn kf;
  st = build_kernel_fn(p->kernel_id, p->nparams, p->params, p->wts, p->add_dot, p->degree, p->d, &kf);
  if (st != SMMD_OK) return st;
  const size_t need = smmd_mmd2_workspace_bytes(p, want_grad);
  if (!workspace || workspace_bytes < need || !aligned(workspace, 256)) return SMMD_EWORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Geometry g = make_geometry(p);
  const Coefs c = make_coefs(g, kf);
  const int prec = resolve_precision(p, want_grad);

  if (prec == SMMD_PREC_FP32) {
    if (want_grad && p->d > kExactGradMaxD) return SMMD_EUNSUPPORTED;   // (dot kernel / explicit fp32 with d > 2048)
    const SimtPlan pl = simt_plan(p->m, p->n, p->d, 1);
    char* ws = static_cast<char*>(workspace);
    if (!tc_small_kernel_disabled() && small_mmd2_eligible(kf, g, src)) {   // latency-bound shapes: one launch (+ a 4-byte memset node)
      g_path = "simt_fp32_small";
      prof_begin(s);
      SMMD_CUDA(launch_small_mmd2(kf, g, c, src, dX, dY, reinterpret_cast<double*>(ws + pl.off_stats),
                                  reinterpret_cast<unsigned int*>(ws + pl.off_norm), scalars, s));
      prof_end(s);
      return SMMD_OK;
    }
    g_path = "simt_fp32";
    float* Z = reinterpret_cast<float*>(ws + pl.off_Z);
    float* norms = reinterpret_cast<float*>(ws + pl.off_norm);
    double* stats = reinterpret_cast<double*>(ws + pl.off_stats);
    SMMD_CUDA(launch_prep_f32(src, p->m, p->n, p->d, kf.tanh_features, Z, norms, pl.dpitch, s));
    prof_begin(s);
    SMMD_CUDA(launch_simt_rows(kf, g, c, Z, norms, pl.dpitch, 1, stats, dX, dY, 0, s));
    prof_end(s);
    SMMD_CUDA(launch_finalize_mmd2(kf, g, stats, norms, scalars, s));
    return SMMD_OK;
  }
  if (prec == SMMD_PREC_BF16X3 && want_grad) return SMMD_EUNSUPPORTED;
  if (prec == SMMD_PREC_FP16 && !want_grad) return SMMD_EUNSUPPORTED;   // the fp16 operand tier exists for the gradient paths
  if (!tc_mmd2_covers(kf, g, want_grad)) return SMMD_EUNSUPPORTED;
  int launches = 0;
  Coefs ct = c;
  ct.f16 = prec == SMMD_PREC_FP16 ? 1 : 0;
  cudaError_t e = tc_mmd2_run(kf, g, ct, src, prec, scalars, dX, dY, workspace, workspace_bytes, s, &launches, &g_path);
  g_launches += launches;
  if (e != cudaSuccess) return cuda_fail(e);
  return SMMD_OK;
}

int smmd_mmd2_fwd_bwd(const smmd_problem* p, const void* X, const void* Y, double* scalars, float* dX, float* dY,
                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!p) return SMMD_EINVAL;
  SrcLayout src{X, Y, p->dtype, p->ldx, p->ldy, 0, 0, nullptr, nullptr, 0};
  return mmd2_fwd_bwd_impl(p, src, scalars, dX, dY, workspace, workspace_bytes, stream);
}

int smmd_mmd2_fwd_bwd_gathered(const smmd_problem* p, const void* gathered, int64_t ld, const float* X_owned,
                               const float* Y_owned, int64_t ld_owned, double* scalars, float* dX, float* dY,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!p || !gathered) return SMMD_EINVAL;
  if (p->world < 1 || p->m % p->world || p->n % p->world || ld < p->d) return SMMD_ESHAPE;
  if ((X_owned == nullptr) != (Y_owned == nullptr) || (X_owned && ld_owned < p->d)) return SMMD_EINVAL;
  SrcLayout src{gathered, gathered, p->dtype, ld, ld, p->m / p->world, p->n / p->world, X_owned, Y_owned, ld_owned};
  return mmd2_fwd_bwd_impl(p, src, scalars, dX, dY, workspace, workspace_bytes, stream);
}

int smmd_peer_set_pull_event(void* cuda_event) {
  g_pull_event = static_cast<cudaEvent_t>(cuda_event);
  return SMMD_OK;
}

size_t smmd_peer_buffer_bytes(int64_t rows_local, int64_t d) {
  if (rows_local <= 0 || d <= 0) return 0;
  return peer_buffer_bytes(rows_local, d);
}

int smmd_mmd2_fwd_bwd_peers(const smmd_problem* p, const smmd_peer_table* peers, uint64_t step, const float* X_local,
                            const float* Y_local, int64_t ld_local, double* scalars, float* dX, float* dY,
                            void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  g_path = "none";
  if (!p || !peers || !X_local || !Y_local || !scalars) return SMMD_EINVAL;
  int st = validate_problem(p);
  if (st != SMMD_OK) return st;
  if (peers->world != p->world || peers->rank != p->rank || p->world > SMMD_MAX_PEERS) return SMMD_EINVAL;
  for (int r = 0; r < p->world; ++r)
    if (!peers->base[r] || !aligned(peers->base[r], 256)) return SMMD_EINVAL;
  if (p->m % p->world || p->n % p->world || ld_local < p->d) return SMMD_ESHAPE;
  if ((dX == nullptr) != (dY == nullptr)) return SMMD_EINVAL;
  if (p->kernel_id == SMMD_K_POLY) return SMMD_EUNSUPPORTED;
  if (!device_ok()) return SMMD_EARCH;
  const int want_grad = dX != nullptr;
  KernelFn kf;
  st = build_kernel_fn(p->kernel_id, p->nparams, p->params, p->wts, p->add_dot, p->degree, p->d, &kf);
  if (st != SMMD_OK) return st;
  const size_t need = smmd_mmd2_workspace_bytes(p, want_grad);
  if (!workspace || workspace_bytes < need || !aligned(workspace, 256)) return SMMD_EWORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Geometry g = make_geometry(p);
  const Coefs c = make_coefs(g, kf);
  const int prec = resolve_precision(p, want_grad);
  const int64_t blk_x = p->m / p->world, blk_y = p->n / p->world;
  char* ws = static_cast<char*>(workspace);

  if (prec == SMMD_PREC_FP32) {
    // latency-bound shapes: one launch does the whole sharded step.  (Mid-size exact problems have no peer variant:
    // the caller uses the collective-based path for those.)
    if (!peer_small_eligible(kf, g)) return SMMD_EUNSUPPORTED;
    const SimtPlan pl = simt_plan(p->m, p->n, p->d, 1);
    g_path = "simt_fp32_small_peer";
    prof_begin(s);
    SMMD_CUDA(launch_peer_small_mmd2(kf, g, c, X_local, Y_local, ld_local, dX, dY, reinterpret_cast<double*>(ws + pl.off_stats),
                                     reinterpret_cast<unsigned int*>(ws + pl.off_norm), scalars, *peers, step, s));
    prof_end(s);
    g_launches = 1;
    if (g_pull_event) cudaEventRecord(g_pull_event, s);
    return SMMD_OK;
  }
  if (prec == SMMD_PREC_BF16X3 && want_grad) return SMMD_EUNSUPPORTED;
  if (prec == SMMD_PREC_FP16 && !want_grad) return SMMD_EUNSUPPORTED;
  if (!tc_mmd2_covers(kf, g, want_grad)) return SMMD_EUNSUPPORTED;
  // 1. publish the local rows in the operand format (bf16; the fp16 tier publishes fp32) and raise the data flags
  const int to_bf16 = prec == SMMD_PREC_FP16 ? 0 : 1;
  SMMD_CUDA(launch_peer_publish(X_local, Y_local, ld_local, blk_x, blk_y, p->d, to_bf16, *peers, step, s));
  // 2. the tensor-core path on the owned rows; its operand preparation pulls every peer's rows over NVLink
  const PeerSrc ps = make_peer_src(*peers, blk_x + blk_y, p->d, step);
  SrcLayout src{peers->base[p->rank], peers->base[p->rank], to_bf16 ? SMMD_BF16 : SMMD_F32, p->d, p->d, blk_x, blk_y,
                X_local, Y_local, ld_local, &ps};   // (X / Y only have to be non-null: the preparation reads the peers' slots)
  int launches = 0;
  Coefs ct = c;
  ct.f16 = prec == SMMD_PREC_FP16 ? 1 : 0;
  g_in_peer_call = 1;
  cudaError_t e = tc_mmd2_run(kf, g, ct, src, prec, scalars, dX, dY, workspace, workspace_bytes, s, &launches, &g_path);
  g_in_peer_call = 0;
  g_launches = launches + (p->world > 1 ? 3 : 2);
  if (e != cudaSuccess) return cuda_fail(e);
  // 3. partial sums -> every peer, wait, add in rank order, MMD^2 (replaces all_reduce + smmd_mmd2_combine)
  SMMD_CUDA(launch_peer_combine(kf, g, scalars, *peers, step, s));
  return SMMD_OK;
}

int smmd_mmd2_combine(const smmd_problem* p, const double* sums, double* out, void* stream) {
  g_launches = 0;
  int st = validate_problem(p);
  if (st != SMMD_OK) return st;
  if (!sums || !out) return SMMD_EINVAL;
  if (!device_ok()) return SMMD_EARCH;
  KernelFn kf;
  st = build_kernel_fn(p->kernel_id, p->nparams, p->params, p->wts, p->add_dot, p->degree, p->d, &kf);
  if (st != SMMD_OK) return st;
  SMMD_CUDA(launch_combine_mmd2(kf, make_geometry(p), sums, out, static_cast<cudaStream_t>(stream)));
  return SMMD_OK;
}

int smmd_mmd2_and_ratio(const smmd_problem* p, const void* X, const void* Y, double min_var_est, double* scalars,
                        void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  g_path = "none";
  int st = validate_problem(p);
  if (st != SMMD_OK) return st;
  if (!X || !Y || !scalars) return SMMD_EINVAL;
  if (p->m != p->n) return SMMD_ESHAPE;  // "Assumes X, Y are same shape" (mmd.py:237)
  if (p->world != 1) return SMMD_EUNSUPPORTED;
  if (!device_ok()) return SMMD_EARCH;
  KernelFn kf;
  st = build_kernel_fn(p->kernel_id, p->nparams, p->params, p->wts, p->add_dot, p->degree, p->d, &kf);
  if (st != SMMD_OK) return st;
  kf.true_distance = 1;  // second-order statistics see the real K entries
  const SimtPlan pl = simt_plan(p->m, p->n, p->d, 1);
  if (!workspace || workspace_bytes < pl.off_end || !aligned(workspace, 256)) return SMMD_EWORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  g_path = "simt_fp32_stats";
  char* ws = static_cast<char*>(workspace);
  float* Z = reinterpret_cast<float*>(ws + pl.off_Z);
  float* norms = reinterpret_cast<float*>(ws + pl.off_norm);
  double* stats = reinterpret_cast<double*>(ws + pl.off_stats);
  Geometry g = make_geometry(p);
  const Coefs c = make_coefs(g, kf);
  SrcLayout src{X, Y, p->dtype, p->ldx, p->ldy, 0, 0, nullptr, nullptr, 0};
  SMMD_CUDA(launch_prep_f32(src, p->m, p->n, p->d, kf.tanh_features, Z, norms, pl.dpitch, s));
  SMMD_CUDA(launch_simt_rows(kf, g, c, Z, norms, pl.dpitch, 1, stats, nullptr, nullptr, 1, s));
  SMMD_CUDA(launch_finalize_ratio(kf, g, stats, min_var_est, scalars, s));
  return SMMD_OK;
}

// K_XY_only needs Z/norms scratch: it is carved from a small internal cache-free scheme -- the caller
// passes no workspace here, so these two entry points use stream-ordered allocation.
static int witness_common(const smmd_problem* p, const void* X, const void* Y, KernelFn* kf, float** Z, float** norms,
                          SimtPlan* pl, cudaStream_t s) {
  int st = validate_problem(p);
  if (st != SMMD_OK) return st;
  if (!X || !Y) return SMMD_EINVAL;
  if (!device_ok()) return SMMD_EARCH;
  st = build_kernel_fn(p->kernel_id, p->nparams, p->params, p->wts, p->add_dot, p->degree, p->d, kf);
  if (st != SMMD_OK) return st;
  kf->true_distance = 1;
  *pl = simt_plan(p->m, p->n, p->d, 1);
  void* buf = nullptr;
  cudaError_t e = cudaMallocAsync(&buf, pl->off_stats, s);
  if (e != cudaSuccess) return cuda_fail(e);
  *Z = reinterpret_cast<float*>(static_cast<char*>(buf) + pl->off_Z);
  *norms = reinterpret_cast<float*>(static_cast<char*>(buf) + pl->off_norm);
  SrcLayout src{X, Y, p->dtype, p->ldx, p->ldy, 0, 0, nullptr, nullptr, 0};
  e = launch_prep_f32(src, p->m, p->n, p->d, kf->tanh_features, *Z, *norms, pl->dpitch, s);
  if (e != cudaSuccess) {
    cudaFreeAsync(buf, s);
    return cuda_fail(e);
  }
  ++g_launches;
  return SMMD_OK;
}

int smmd_kernel_xy(const smmd_problem* p, const void* X, const void* Y, float* K, int64_t ldk, void* stream) {
  g_launches = 0;
  g_path = "simt_fp32_kxy";
  if (!K || (p && ldk < p->n)) return SMMD_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KernelFn kf;
  float *Z, *norms;
  SimtPlan pl;
  int st = witness_common(p, X, Y, &kf, &Z, &norms, &pl, s);
  if (st != SMMD_OK) return st;
  cudaError_t e = launch_kernel_xy(kf, Z, norms, pl.dpitch, p->m, p->n, p->d, K, ldk, s);
  cudaFreeAsync(Z, s);  // Z is the base of the allocation (off_Z == 0)
  if (e != cudaSuccess) return cuda_fail(e);
  ++g_launches;
  return SMMD_OK;
}

int smmd_kernel_xy_bwd(const smmd_problem* p, const void* X, const void* Y, const float* dK, int64_t lddk, float* dX,
                       float* dY, void* stream) {
  g_launches = 0;
  g_path = "simt_fp32_kxy_bwd";
  if (!dK || !dX || !dY || (p && lddk < p->n)) return SMMD_EINVAL;
  if (p && p->d > kExactGradMaxD) return SMMD_EUNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KernelFn kf;
  float *Z, *norms;
  SimtPlan pl;
  int st = witness_common(p, X, Y, &kf, &Z, &norms, &pl, s);
  if (st != SMMD_OK) return st;
  cudaError_t e = launch_kernel_xy_bwd(kf, Z, norms, pl.dpitch, p->m, p->n, p->d, dK, lddk, dX, dY, s);
  cudaFreeAsync(Z, s);
  if (e != cudaSuccess) return cuda_fail(e);
  ++g_launches;
  return SMMD_OK;
}

int smmd_kernel_xy_bwd2(const smmd_problem* p, const void* X, const void* Y, const float* dK, int64_t lddk,
                        const float* VX, const float* VY, float* ddK, float* gX, float* gY, void* stream) {
  g_launches = 0;
  g_path = "simt_fp32_kxy_bwd2";
  if (!dK || !ddK || !gX || !gY || (p && lddk < p->n)) return SMMD_EINVAL;
  if (p && (p->kernel_id == SMMD_K_POLY || p->kernel_id == SMMD_K_TANH_DISTANCE || p->kernel_id == SMMD_K_TANH_MIX_RQ))
    return SMMD_EUNSUPPORTED;   // tanh kernels: the caller applies tanh to the features (smmd/mmd.py kernel_xy)
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KernelFn kf;
  float *Z, *norms;
  SimtPlan pl;
  int st = witness_common(p, X, Y, &kf, &Z, &norms, &pl, s);
  if (st != SMMD_OK) return st;
  float* AB = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&AB), (size_t)2 * p->m * p->n * sizeof(float), s);
  if (e != cudaSuccess) {
    cudaFreeAsync(Z, s);
    return cuda_fail(e);
  }
  e = launch_kernel_xy_bwd2(kf, Z, norms, pl.dpitch, p->m, p->n, p->d, dK, lddk, VX, VY, ddK, AB, AB + p->m * p->n, gX,
                            gY, s);
  cudaFreeAsync(AB, s);
  cudaFreeAsync(Z, s);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 3;
  return SMMD_OK;
}

// ---- KID -----------------------------------------------------------------------------------------
static int validate_kid(const smmd_kid_problem* p) {
  if (!p) return SMMD_EINVAL;
  if (p->n_g < 1 || p->n_r < 1 || p->d < 1 || p->ldg < p->d || p->ldr < p->d) return SMMD_ESHAPE;
  if (p->n_subsets < 1 || p->subset_size < 3) return SMMD_ESHAPE;
  if (p->subset_size > p->n_g || p->subset_size > p->n_r) return SMMD_ESHAPE;  // replace=False (compute_scores.py:221)
  if (p->dtype != SMMD_F32 && p->dtype != SMMD_BF16) return SMMD_EDTYPE;
  if (p->degree < 1 || p->degree > 8) return SMMD_EINVAL;
  if (p->mmd_est < SMMD_EST_UNBIASED || p->mmd_est > SMMD_EST_USTAT) return SMMD_EINVAL;
  if (p->precision < SMMD_PREC_FP32 || p->precision > SMMD_PREC_AUTO) return SMMD_EINVAL;
  if (p->first_subset < 0 || p->first_subset >= p->n_subsets) return SMMD_EINVAL;
  if (p->n_local > 0 && p->first_subset + p->n_local > p->n_subsets) return SMMD_EINVAL;
  return SMMD_OK;
}
static int kid_local(const smmd_kid_problem* p) { return p->n_local > 0 ? p->n_local : p->n_subsets - p->first_subset; }
static int kid_precision(const smmd_kid_problem* p) {
  int prec = p->precision;
  if (prec == SMMD_PREC_AUTO) prec = tc_kid_supported(p->d, p->subset_size, kid_local(p)) ? SMMD_PREC_BF16X3 : SMMD_PREC_FP32;
  return prec;
}

size_t smmd_kid_workspace_bytes(const smmd_kid_problem* p) {
  if (validate_kid(p) != SMMD_OK) return 0;
  if (kid_precision(p) != SMMD_PREC_FP32 && !tc_kid_supported(p->d, p->subset_size, kid_local(p))) return 0;   // refused
  if (kid_precision(p) == SMMD_PREC_FP32) return simt_plan(p->subset_size, p->subset_size, p->d, kid_local(p)).off_end;
  return tc_kid_workspace_bytes(p->subset_size, p->d, kid_local(p), kid_precision(p));
}

// Gram row statistics of the local subsets: stats[nloc][2m][RS_COUNT] inside the workspace.
static int kid_row_stats(const smmd_kid_problem* p, const void* codes_g, const void* codes_r, const int32_t* idx_g,
                         const int32_t* idx_r, int want_second_order, void* workspace, size_t workspace_bytes,
                         cudaStream_t s, double** stats_out) {
  KernelFn kf;
  float prm[2] = {p->gamma, p->coef0};
  int st = build_kernel_fn(SMMD_K_POLY, 2, prm, prm, 0.f, p->degree, p->d, &kf);
  if (st != SMMD_OK) return st;
  const int64_t m = p->subset_size, nloc = kid_local(p);
  const int prec = kid_precision(p);
  char* ws = static_cast<char*>(workspace);
  if (prec == SMMD_PREC_FP32) {
    g_path = "simt_fp32_kid";
    const SimtPlan pl = simt_plan(m, m, p->d, nloc);
    float* Z = reinterpret_cast<float*>(ws + pl.off_Z);
    float* norms = reinterpret_cast<float*>(ws + pl.off_norm);
    *stats_out = reinterpret_cast<double*>(ws + pl.off_stats);
    SMMD_CUDA(launch_gather_f32(codes_g, codes_r, p->dtype, p->ldg, p->ldr, p->d, idx_g, idx_r, p->first_subset, nloc, m,
                                Z, norms, pl.dpitch, s));
    Geometry g{m, m, p->d, 0, m, 0, m, 0};
    Coefs c = make_coefs(g, kf);
    prof_begin(s);
    SMMD_CUDA(launch_simt_rows(kf, g, c, Z, norms, pl.dpitch, nloc, *stats_out, nullptr, nullptr, 1, s));
    prof_end(s);
  } else {
    if (!tc_kid_supported(p->d, m, nloc)) return SMMD_EUNSUPPORTED;
    int launches = 0;
    cudaError_t e = tc_kid_run(kf, codes_g, codes_r, p->dtype, p->ldg, p->ldr, p->d, idx_g, idx_r, p->first_subset, nloc,
                               m, prec, want_second_order, workspace, workspace_bytes, stats_out, s, &launches, &g_path);
    g_launches += launches;
    if (e != cudaSuccess) return cuda_fail(e);
  }
  return SMMD_OK;
}

int smmd_kid_subsets(const smmd_kid_problem* p, const void* codes_g, const void* codes_r, const int32_t* idx_g,
                     const int32_t* idx_r, double* mmd2_out, double* var_out, void* workspace, size_t workspace_bytes,
                     void* stream) {
  g_launches = 0;
  g_path = "none";
  int st = validate_kid(p);
  if (st != SMMD_OK) return st;
  if (!codes_g || !codes_r || !idx_g || !idx_r || !mmd2_out) return SMMD_EINVAL;
  if (p->ret_var && !var_out) return SMMD_EINVAL;
  if (!device_ok()) return SMMD_EARCH;
  if (kid_precision(p) != SMMD_PREC_FP32 && !tc_kid_supported(p->d, p->subset_size, kid_local(p))) return SMMD_EUNSUPPORTED;
  const size_t need = smmd_kid_workspace_bytes(p);
  if (!workspace || workspace_bytes < need || !aligned(workspace, 256)) return SMMD_EWORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t m = p->subset_size, nloc = kid_local(p);
  const int64_t var_at_m = p->var_at_m > 0 ? p->var_at_m : m;
  double* stats = nullptr;
  st = kid_row_stats(p, codes_g, codes_r, idx_g, idx_r, p->ret_var || p->mmd_est == SMMD_EST_USTAT, workspace,
                     workspace_bytes, s, &stats);
  if (st != SMMD_OK) return st;
  SMMD_CUDA(launch_finalize_kid(stats, nloc, m, p->first_subset, p->mmd_est, p->ret_var, var_at_m, mmd2_out, var_out, s));
  return SMMD_OK;
}

// ---- 3-sample test sums (gan/core/mmd.py:429-444, 515-539) -----------------------------------------
static int validate_poly_sums(const smmd_kid_problem* p) {
  int st = validate_kid(p);
  if (st != SMMD_OK) return st;
  // X, Y (and the saved Z) are equally sized samples (mmd.py:516 "Assumes X, Y, Z are same shape")
  if (p->n_subsets != 1 || p->n_g != p->subset_size || p->n_r != p->subset_size) return SMMD_ESHAPE;
  if (p->first_subset != 0 || p->n_local > 1) return SMMD_EINVAL;
  return SMMD_OK;
}

size_t smmd_poly_sums_workspace_bytes(const smmd_kid_problem* p) {
  if (validate_poly_sums(p) != SMMD_OK) return 0;
  return smmd_kid_workspace_bytes(p);
}

// ---- estimators on caller-supplied row statistics (dense-block compatibility entry points of the Python layer) ----
int smmd_kid_from_row_stats(const double* stats, int64_t m, int mmd_est, int ret_var, int64_t var_at_m, double* mmd2_out,
                            double* var_out, void* stream) {
  if (!stats || !mmd2_out || (ret_var && !var_out)) return SMMD_EINVAL;
  if (m < 3) return SMMD_ESHAPE;
  if (mmd_est < SMMD_EST_UNBIASED || mmd_est > SMMD_EST_USTAT) return SMMD_EINVAL;
  if (!device_ok()) return SMMD_EARCH;
  SMMD_CUDA(launch_finalize_kid(stats, 1, m, 0, mmd_est, ret_var, var_at_m > 0 ? var_at_m : m, mmd2_out, var_out,
                                static_cast<cudaStream_t>(stream)));
  return SMMD_OK;
}

int smmd_ratio_from_row_stats(const double* stats, int64_t m, int biased, int has_const_diagonal, double const_diagonal,
                              double min_var_est, double* scalars, void* stream) {
  if (!stats || !scalars) return SMMD_EINVAL;
  if (m < 2) return SMMD_ESHAPE;
  if (!device_ok()) return SMMD_EARCH;
  KernelFn kf;
  memset(&kf, 0, sizeof(kf));
  kf.has_const_diag = has_const_diagonal ? 1 : 0;
  kf.const_diag = (float)const_diagonal;
  Geometry g{m, m, 1, 0, m, 0, m, biased ? 1 : 0};
  SMMD_CUDA(launch_finalize_ratio(kf, g, stats, min_var_est, scalars, static_cast<cudaStream_t>(stream)));
  return SMMD_OK;
}

int smmd_poly_sums(const smmd_kid_problem* p, const void* X, const void* Y, double* sums_out, void* workspace,
                   size_t workspace_bytes, void* stream) {
  g_launches = 0;
  g_path = "none";
  int st = validate_poly_sums(p);
  if (st != SMMD_OK) return st;
  if (!X || !Y || !sums_out) return SMMD_EINVAL;
  if (!device_ok()) return SMMD_EARCH;
  if (kid_precision(p) != SMMD_PREC_FP32 && !tc_kid_supported(p->d, p->subset_size, kid_local(p))) return SMMD_EUNSUPPORTED;
  const size_t need = smmd_poly_sums_workspace_bytes(p);
  if (!workspace || workspace_bytes < need || !aligned(workspace, 256)) return SMMD_EWORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* stats = nullptr;
  st = kid_row_stats(p, X, Y, nullptr, nullptr, 1, workspace, workspace_bytes, s, &stats);
  if (st != SMMD_OK) return st;
  SMMD_CUDA(launch_poly_sums(stats, p->subset_size, sums_out, s));
  return SMMD_OK;
}

}  // extern "C"

// tc_check.cu -- developer harness: runs the tensor-core path of libsmmd.so against the (oracle-validated)
// exact fp32 path through the public C ABI, and times both with CUDA events.
//   tc_check mmd <kernel: rbf|mix_rbf|mix_rq|mix_rq_dot|distance> <m> <n> <d> [reps]
//   tc_check kid <n_codes> <d> <subsets> <m> [reps]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <cuda_runtime.h>
#include "../../../include/smmd.h"

#ifdef SMMD_PIPE_TIMING
namespace smmd { void pipe_timing_dump(bool reset); }
#endif
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)
#define SK(x) do { int st = (x); if (st != 0) { printf("smmd error %d (%s) [%s] at line %d\n", st, smmd_strerror(st), smmd_last_cuda_error(), __LINE__); exit(3); } } while (0)

static void fill_problem(smmd_problem* p, const char* kname, int64_t m, int64_t n, int64_t d) {
  memset(p, 0, sizeof(*p));
  p->m = m; p->n = n; p->d = d; p->ldx = d; p->ldy = d;
  p->dtype = SMMD_F32; p->rank = 0; p->world = 1; p->biased = 0;
  if (!strcmp(kname, "rbf")) { p->kernel_id = SMMD_K_RBF; p->nparams = 1; p->params[0] = 1.f; p->wts[0] = 1.f; }
  else if (!strcmp(kname, "mix_rbf")) {
    p->kernel_id = SMMD_K_MIX_RBF; p->nparams = 5;
    const float s[5] = {1, 2, 4, 8, 16};
    for (int i = 0; i < 5; ++i) { p->params[i] = s[i]; p->wts[i] = 1.f; }
  } else if (!strcmp(kname, "mix_rq") || !strcmp(kname, "mix_rq_dot")) {
    p->kernel_id = SMMD_K_MIX_RQ; p->nparams = 3;
    const float a[3] = {0.1f, 1.f, 10.f};
    for (int i = 0; i < 3; ++i) { p->params[i] = a[i]; p->wts[i] = 1.f; }
    p->add_dot = !strcmp(kname, "mix_rq_dot") ? 0.1f : 0.f;
  } else if (!strcmp(kname, "distance")) { p->kernel_id = SMMD_K_DISTANCE; }
  else { printf("unknown kernel %s\n", kname); exit(1); }
}

struct Result { double sc[SMMD_NUM_SCALARS]; std::vector<float> gx, gy; float ms; const char* path; int launches; };

static Result run_mmd(smmd_problem p, int prec, const float* dX_in, const float* dY_in, int reps, bool grad) {
  p.precision = prec;
  Result r;
  size_t wsb = smmd_mmd2_workspace_bytes(&p, grad);
  void* ws; double* sc; float *gx = nullptr, *gy = nullptr;
  CK(cudaMalloc(&ws, wsb)); CK(cudaMalloc(&sc, sizeof(double) * SMMD_NUM_SCALARS));
  if (grad) { CK(cudaMalloc(&gx, p.m * p.d * 4)); CK(cudaMalloc(&gy, p.n * p.d * 4)); }
  SK(smmd_mmd2_fwd_bwd(&p, dX_in, dY_in, sc, gx, gy, ws, wsb, nullptr));
  r.path = smmd_last_path(); r.launches = smmd_last_launch_count();
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) SK(smmd_mmd2_fwd_bwd(&p, dX_in, dY_in, sc, gx, gy, ws, wsb, nullptr));
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&r.ms, e0, e1)); r.ms /= reps;
  CK(cudaMemcpy(r.sc, sc, sizeof(r.sc), cudaMemcpyDeviceToHost));
  if (grad) {
    r.gx.resize(p.m * p.d); r.gy.resize(p.n * p.d);
    CK(cudaMemcpy(r.gx.data(), gx, r.gx.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r.gy.data(), gy, r.gy.size() * 4, cudaMemcpyDeviceToHost));
    cudaFree(gx); cudaFree(gy);
  }
  cudaFree(ws); cudaFree(sc);
  return r;
}

static void cmp(const std::vector<float>& a, const std::vector<float>& b, const char* name) {
  double mx = 0, mag = 0; size_t at = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    double e = fabs((double)a[i] - b[i]);
    if (!(e <= mx)) { mx = e; at = i; }
    mag = fmax(mag, fabs((double)b[i]));
  }
  printf("   %s: max|err| %.3e  max|ref| %.3e  rel %.3e  (at %zu: %g vs %g)\n", name, mx, mag, mx / fmax(mag, 1e-300), at,
         a.size() ? a[at] : 0.f, b.size() ? b[at] : 0.f);
}

__global__ void spin_kernel(long long cycles, unsigned long long* out) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  long long c0 = clock64();
  while (clock64() - c0 < cycles) {}
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (unsigned long long)(clock64() - c0); }
}
// warm the clocks up (idle B200s sit at 120 MHz) and report the effective SM frequency
static void warm_and_report(const char* tag) {
  unsigned long long* d; CK(cudaMalloc(&d, 16));
  for (int i = 0; i < 3; ++i) spin_kernel<<<148, 128>>>(100000000LL, d);
  unsigned long long h[2];
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  printf("[clock %s] %.0f MHz effective SM clock\n", tag, (double)h[1] / (double)h[0] * 1e3);
  cudaFree(d);
}

// `tc_check suite`: one short call of every tensor-core gradient path in ONE process (for compute-sanitizer runs): the
// fused symmetric path (tc_symf_kernel + mirrored pass), the plain symmetric path (d > 256), the row-stacked fused
// tile-pair kernel and the row-stacked two-pass path, each against the exact fp32 path.
static int run_suite() {
  struct Case { const char* label; const char* k; int64_t m, n, d; long long sym, min_rows; int prec; };
  const Case cases[] = {{"symf", "mix_rq", 1500, 1400, 256, 1, 1, SMMD_PREC_BF16}, {"sym", "mix_rq", 1200, 1300, 320, 1, 1, SMMD_PREC_BF16},
                        {"fused_pair", "mix_rq", 1500, 1400, 256, 0, 0, SMMD_PREC_BF16}, {"wz", "mix_rq", 1200, 1300, 320, 0, 0, SMMD_PREC_BF16},
                        {"symf_rbf_d64", "rbf", 700, 900, 64, 1, 1, SMMD_PREC_BF16},
                        {"symf_f16", "mix_rq", 1500, 1400, 256, 1, 1, SMMD_PREC_FP16}, {"sym_f16", "mix_rq", 1200, 1300, 320, 1, 1, SMMD_PREC_FP16},
                        {"fused_f16", "mix_rq", 1500, 1400, 256, 0, 0, SMMD_PREC_FP16}, {"wz_f16", "mix_rq", 1200, 1300, 320, 0, 0, SMMD_PREC_FP16},
                        {"symf_f16_rbf", "mix_rbf", 700, 900, 64, 1, 1, SMMD_PREC_FP16}, {"sym_f16_dist", "distance", 900, 800, 512, 1, 1, SMMD_PREC_FP16}};
  int bad = 0;
  for (const Case& c : cases) {
    smmd_set_option("sym", c.sym);
    smmd_set_option("sym_min_rows", c.min_rows);
    smmd_set_option("symf_min_rows", c.min_rows);
    std::mt19937 rng(99);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> hX(c.m * c.d), hY(c.n * c.d);
    const float sc = 1.f / sqrtf((float)c.d);
    for (auto& v : hX) v = nd(rng) * sc;
    for (auto& v : hY) v = (1.05f * nd(rng) + 0.1f) * sc;
    float *dXin, *dYin;
    CK(cudaMalloc(&dXin, hX.size() * 4)); CK(cudaMalloc(&dYin, hY.size() * 4));
    CK(cudaMemcpy(dXin, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dYin, hY.data(), hY.size() * 4, cudaMemcpyHostToDevice));
    smmd_problem p; fill_problem(&p, c.k, c.m, c.n, c.d);
    Result tc = run_mmd(p, c.prec, dXin, dYin, 1, true);
    Result ex = run_mmd(p, SMMD_PREC_FP32, dXin, dYin, 1, true);
    double gmax = 0, emax = 0;
    for (size_t i = 0; i < tc.gx.size(); ++i) { gmax = fmax(gmax, fabs((double)ex.gx[i])); emax = fmax(emax, fabs((double)tc.gx[i] - ex.gx[i])); }
    for (size_t i = 0; i < tc.gy.size(); ++i) { gmax = fmax(gmax, fabs((double)ex.gy[i])); emax = fmax(emax, fabs((double)tc.gy[i] - ex.gy[i])); }
    const double vrel = fabs(tc.sc[0] - ex.sc[0]) / fabs(ex.sc[0]);
    const bool ok = vrel <= 1e-3 && emax <= (c.prec == SMMD_PREC_FP16 ? 1e-3 : 4e-3) * gmax;
    printf("[suite %-12s %s %lldx%lldx%lld] path=%s launches=%d  mmd2 rel %.2e  grad err/max %.2e  %s\n", c.label, c.k,
           (long long)c.m, (long long)c.n, (long long)c.d, tc.path, tc.launches, vrel, emax / gmax, ok ? "ok" : "MISMATCH");
    bad += !ok;
    cudaFree(dXin); cudaFree(dYin);
  }
  return bad;
}

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  if (!strcmp(argv[1], "suite")) return run_suite();
  warm_and_report("start");
  if (!strcmp(argv[1], "mmd")) {
    const char* kname = argv[2];
    int64_t m = atoll(argv[3]), n = atoll(argv[4]), d = atoll(argv[5]);
    int reps = argc > 6 ? atoi(argv[6]) : 3;
    int ref = argc > 7 ? atoi(argv[7]) : 1;
    std::mt19937 rng(1234);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> hX(m * d), hY(n * d);
    const float sc = 1.f / sqrtf((float)d);
    for (auto& v : hX) v = nd(rng) * sc;
    for (auto& v : hY) v = (1.05f * nd(rng) + 0.1f) * sc;
    float *dXin, *dYin;
    CK(cudaMalloc(&dXin, hX.size() * 4)); CK(cudaMalloc(&dYin, hY.size() * 4));
    CK(cudaMemcpy(dXin, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dYin, hY.data(), hY.size() * 4, cudaMemcpyHostToDevice));
    smmd_problem p; fill_problem(&p, kname, m, n, d);
    const int tcprec = getenv("TC_CHECK_FP16") ? SMMD_PREC_FP16 : SMMD_PREC_BF16;   // fp16 operand tier
    Result tc = run_mmd(p, tcprec, dXin, dYin, reps, true);
    const double pairs = (double)m * n * d;  // N^2 d (m = n)
    printf("[%s m=%lld n=%lld d=%lld] TC  path=%s launches=%d  %.3f ms  mmd2=%.9g  nonfinite=%g  -> %.3e pairs/s, %.1f TFLOP/s (14 N^2 d)\n",
           kname, (long long)m, (long long)n, (long long)d, tc.path, tc.launches, tc.ms, tc.sc[0], tc.sc[7],
           pairs / (tc.ms * 1e-3), 14.0 * pairs / (tc.ms * 1e-3) / 1e12);
#ifdef SMMD_PIPE_TIMING
    smmd::pipe_timing_dump(true);
#endif
    Result tv = run_mmd(p, SMMD_PREC_BF16, dXin, dYin, reps, false);
    printf("   value-only TC path=%s %.3f ms mmd2=%.9g\n", tv.path, tv.ms, tv.sc[0]);
    if (ref) {
      Result ex = run_mmd(p, SMMD_PREC_FP32, dXin, dYin, 1, true);
      printf("   exact path=%s %.3f ms mmd2=%.9g   rel diff TC vs exact: %.3e (value-only: %.3e)\n", ex.path, ex.ms, ex.sc[0],
             fabs(tc.sc[0] - ex.sc[0]) / fabs(ex.sc[0]), fabs(tv.sc[0] - ex.sc[0]) / fabs(ex.sc[0]));
      for (int i = 1; i <= 6; ++i) printf("   sum[%d] tc %.9g exact %.9g\n", i, tc.sc[i], ex.sc[i]);
      cmp(tc.gx, ex.gx, "dX");
      cmp(tc.gy, ex.gy, "dY");
    }
    return 0;
  }
  if (!strcmp(argv[1], "kid")) {
    int64_t nc = atoll(argv[2]), d = atoll(argv[3]);
    int S = atoi(argv[4]), m = atoi(argv[5]);
    int reps = argc > 6 ? atoi(argv[6]) : 3;
    int ref = argc > 7 ? atoi(argv[7]) : 1;
    std::mt19937 rng(1234);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> hg(nc * d), hr(nc * d);
    for (auto& v : hg) v = fmaxf(nd(rng), 0.f);
    for (auto& v : hr) v = fmaxf(nd(rng) + 0.02f, 0.f);
    std::vector<int32_t> ig((size_t)S * m), ir((size_t)S * m);
    for (int s = 0; s < S; ++s) {  // distinct indices per subset (replace=False)
      std::vector<int32_t> perm(nc);
      for (int64_t i = 0; i < nc; ++i) perm[i] = (int32_t)i;
      std::shuffle(perm.begin(), perm.end(), rng);
      for (int i = 0; i < m; ++i) ig[(size_t)s * m + i] = perm[i];
      std::shuffle(perm.begin(), perm.end(), rng);
      for (int i = 0; i < m; ++i) ir[(size_t)s * m + i] = perm[i];
    }
    float *dg, *dr; int32_t *dig, *dir_; double *dm, *dv;
    CK(cudaMalloc(&dg, hg.size() * 4)); CK(cudaMalloc(&dr, hr.size() * 4));
    CK(cudaMalloc(&dig, ig.size() * 4)); CK(cudaMalloc(&dir_, ir.size() * 4));
    CK(cudaMalloc(&dm, S * 8)); CK(cudaMalloc(&dv, S * 8));
    CK(cudaMemcpy(dg, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dr, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dig, ig.data(), ig.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dir_, ir.data(), ir.size() * 4, cudaMemcpyHostToDevice));
    smmd_kid_problem kp; memset(&kp, 0, sizeof(kp));
    kp.n_g = nc; kp.n_r = nc; kp.d = d; kp.ldg = d; kp.ldr = d; kp.dtype = SMMD_F32;
    kp.n_subsets = S; kp.subset_size = m; kp.degree = 3; kp.gamma = -1.f; kp.coef0 = 1.f; kp.var_at_m = nc;
    kp.mmd_est = SMMD_EST_UNBIASED;
    std::vector<double> ref_m(S), ref_v(S);
    const int precs[3] = {SMMD_PREC_FP32, SMMD_PREC_BF16X3, SMMD_PREC_BF16};
    const char* pn[3] = {"fp32", "bf16x3", "bf16"};
    for (int pi = ref ? 0 : 1; pi < 3; ++pi)
      for (int rv = 1; rv >= 0; --rv) {
        kp.precision = precs[pi]; kp.ret_var = rv;
        size_t wsb = smmd_kid_workspace_bytes(&kp);
        void* ws; CK(cudaMalloc(&ws, wsb));
        SK(smmd_kid_subsets(&kp, dg, dr, dig, dir_, dm, dv, ws, wsb, nullptr));
        const char* path = smmd_last_path();
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        int rr = pi == 0 ? 1 : reps;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < rr; ++i) SK(smmd_kid_subsets(&kp, dg, dr, dig, dir_, dm, dv, ws, wsb, nullptr));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= rr;
        std::vector<double> hm(S), hv(S);
        CK(cudaMemcpy(hm.data(), dm, S * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hv.data(), dv, S * 8, cudaMemcpyDeviceToHost));
        if (pi == 0 && rv) { ref_m = hm; ref_v = hv; }
        double em = 0, ev = 0;
        if (ref) for (int s = 0; s < S; ++s) { em = fmax(em, fabs(hm[s] - ref_m[s]) / fabs(ref_m[s])); if (rv) ev = fmax(ev, fabs(hv[s] - ref_v[s]) / fabs(ref_v[s])); }
        printf("[kid n=%lld d=%lld S=%d m=%d] %-7s ret_var=%d path=%s  %.3f ms (ws %.1f MB)  %.3e rows/s  %.1f TFLOP/s(6m^2d)  mmd[0]=%.9g var[0]=%.6g  max rel err mmd %.2e var %.2e\n",
               (long long)nc, (long long)d, S, m, pn[pi], rv, path, ms, wsb / 1e6, 2.0 * S * m / (ms * 1e-3),
               6.0 * m * m * d * S / (ms * 1e-3) / 1e12, hm[0], hv[0], em, ev);
        cudaFree(ws);
      }
    return 0;
  }
  return 1;
}

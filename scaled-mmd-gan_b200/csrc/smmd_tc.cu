// smmd_tc.cu -- tcgen05 (5th-gen tensor core) path, sm_100a only: operand preparation and dispatch.  The kernels:
//
//  smmd_tc_fused.cu
//  tc_fused_kernel   : MMD^2 forward AND backward in one sweep over Gram tiles (flash-attention shaped):
//                        S  = Z_i Z_j^T          UMMA #1 (SS, bf16, fp32 accum in TMEM)
//                        W  = 4 a k'(D(S))       epilogue warps: TMEM -> regs -> kernel transform; block sums,
//                                                row sums; W written back to TMEM as bf16
//                        O += W Z_j              UMMA #2 (A = W from TMEM, B = the SAME Z_j smem tile MN-major)
//                      so neither the N x N kernel matrix nor its derivative ever reaches HBM.
//                      Replaces gan/core/mmd.py:55-188 (kernels) + :194-220 (mmd2) + the TF autodiff graph
//                      (gan/core/model.py:446,452) for d <= 256.
//  smmd_tc_wz.cu
//  tc_wgen_kernel    : pass 1 of the wide-feature (d > 256) backward: 128 x 256 Gram tiles with K streamed, the same
//  tc_wz_kernel        epilogue math, bf16 W tiles stored per row panel; pass 2 is O = W Z as a 256 x 256 macro-tile GEMM
//  wz_finalize_rows_kernel  (details at the kernels; pass 1 runs as cta_group::2 CTA pairs on large panels).
//  smmd_tc_gram.cu
//  tc_stream_kernel  : K-streaming 128 x 128 Gram tiles + fused reduction epilogue (row stats): value-only MMD^2.
//  tc_macro_kernel   : 256 x 256 macro-tile Gram + reduction epilogue, batched over problems: KID
//                      (gan/compute_scores.py:232-335, all subsets in one launch) and the 3-sample sums
//                      (gan/core/mmd.py:515-539).  bf16 or split-bf16 (hi*hi + lo*hi + hi*lo) operands.
//
// Work distribution of the gradient kernels is "stream-K" style: the flattened (row block, column tile) space is cut
// into equal contiguous chunks, one per persistent CTA (grid = #SMs), partial results land in per-(CTA, slot)
// workspace slabs and are reduced in a fixed order by the finalisation kernel (deterministic).  The Gram + reduction
// kernels deal tiles round-robin (L2 residency) and accumulate row statistics with fp64 atomics.
// Shared pieces (constants, tuning knobs, PrepTcArgs, the 16-column epilogue step): smmd_tc_common.cuh.
#include <cuda_fp16.h>
#include "smmd_tc_common.cuh"

namespace smmd {
namespace tc {

__global__ void __launch_bounds__(256) prep_tc_kernel(PrepTcArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t Mp = a.mp + a.np;
  const int64_t p = (int64_t)blockIdx.x * 8 + warp;
  const int64_t b = blockIdx.y;
  if (p >= Mp) return;
  const bool inA = p < a.mp;
  const int64_t loc = inA ? p : p - a.mp;
  const bool valid = loc < (inA ? a.m : a.n);
  int64_t src = loc;
  if (valid && a.idxA) src = inA ? a.idxA[(a.first_batch + b) * a.m + loc] : a.idxB[(a.first_batch + b) * a.n + loc];
  else src = src_row(src, inA, a.blk_a, a.blk_b);
  int64_t ld = inA ? a.lda : a.ldb;
  const void* base = inA ? a.A : a.B;
  if (a.peer.on) {
    // the all_gather, fused: block r of the global rows is read straight from rank r's exchange buffer over NVLink, as
    // soon as that rank's data flag for this step is up (own block: no wait, it was published earlier on this stream)
    const int64_t blk = inA ? a.blk_a : a.blk_b;
    const int64_t r = valid ? loc / blk : (int64_t)a.peer.self;
    const unsigned long long step = peer_step_of(a.peer.step, a.peer.base[a.peer.self]);
    base = peer_data_slot(a.peer.base[r], step, a.peer.slot_bytes);
    src = (inA ? 0 : a.blk_a) + (valid ? loc - r * blk : 0);
    ld = a.d;
    if (valid && r != a.peer.self) {
      if (lane == 0) peer_wait_flag(peer_flag(a.peer.base[a.peer.self], kPeerOffDataFlag, step, (int)r), step);
      __syncwarp();
    }
  }
  __nv_bfloat16* zrow = a.Z + (b * Mp + p) * a.dpz;
  float acc = 0.f;
  // fast path: fp32 rows, 16-B aligned, no split/tanh tail handling needed per element: 8 features per lane
  // per step (2 x LDG.128 -> 1 x STG.128), several independent loads in flight
  const bool vec = a.dtype == SMMD_F32 && !a.split && (ld % 4 == 0) && (a.d % 8 == 0) &&
                   ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
  if (a.f16) {
    // fp16 operand tier: IEEE half operands (11-bit significand); norms from exactly the rounded values
    for (int64_t c = 2 * lane; c < a.dp; c += 64) {
      float v[2] = {0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (valid && c + e < a.d) {
          v[e] = a.dtype == SMMD_F32 ? reinterpret_cast<const float*>(base)[src * ld + c + e]
                                     : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[src * ld + c + e]);
          if (a.tanh_features) v[e] = tanhf(v[e]);
        }
      }
      const __half2 h2 = __floats2half2_rn(v[0], v[1]);
      const float2 f2 = __half22float2(h2);
      *reinterpret_cast<__half2*>(zrow + c) = h2;
      acc = fmaf(f2.x, f2.x, acc);
      acc = fmaf(f2.y, f2.y, acc);
    }
  } else if (vec) {
    const float* srow = reinterpret_cast<const float*>(base) + src * ld;
    for (int64_t c = 8 * lane; c < a.dp; c += 256) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid && c < a.d) {
        const float4 lo = *reinterpret_cast<const float4*>(srow + c), hi = *reinterpret_cast<const float4*>(srow + c + 4);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
        v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        if (a.tanh_features) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = tanhf(v[e]);
        }
      }
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        pk[e] = *reinterpret_cast<const uint32_t*>(&h2);
        const float f0 = __bfloat162float(h2.x), f1 = __bfloat162float(h2.y);
        acc = fmaf(f0, f0, acc);
        acc = fmaf(f1, f1, acc);
      }
      *reinterpret_cast<uint4*>(zrow + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  } else if (a.dtype == SMMD_BF16 && !a.split && !a.tanh_features && (ld % 8 == 0) && (a.d % 8 == 0) &&
             ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) {
    // bf16 rows (the all-gathered block of the multi-GPU path): straight 16-byte copies + norms of the same values
    const __nv_bfloat16* srow = reinterpret_cast<const __nv_bfloat16*>(base) + src * ld;
    for (int64_t c = 8 * lane; c < a.dp; c += 256) {
      uint4 pk = make_uint4(0u, 0u, 0u, 0u);
      if (valid && c < a.d) pk = *reinterpret_cast<const uint4*>(srow + c);
      const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float f0 = __uint_as_float(w[e] << 16), f1 = __uint_as_float(w[e] & 0xffff0000u);
        acc = fmaf(f0, f0, acc);
        acc = fmaf(f1, f1, acc);
      }
      *reinterpret_cast<uint4*>(zrow + c) = pk;
    }
  } else
  for (int64_t c = 2 * lane; c < a.dp; c += 64) {
    float v[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (valid && c + e < a.d) {
        v[e] = a.dtype == SMMD_F32 ? reinterpret_cast<const float*>(base)[src * ld + c + e]
                                   : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[src * ld + c + e]);
        if (a.tanh_features) v[e] = tanhf(v[e]);
      }
    }
    __nv_bfloat16 h0 = __float2bfloat16_rn(v[0]), h1 = __float2bfloat16_rn(v[1]);
    const float f0 = __bfloat162float(h0), f1 = __bfloat162float(h1);
    *reinterpret_cast<__nv_bfloat162*>(zrow + c) = __nv_bfloat162(h0, h1);
    if (a.split) {
      __nv_bfloat16 l0 = __float2bfloat16_rn(v[0] - f0), l1 = __float2bfloat16_rn(v[1] - f1);
      *reinterpret_cast<__nv_bfloat162*>(zrow + a.dp + c) = __nv_bfloat162(l0, l1);
      const float g0 = __bfloat162float(l0), g1 = __bfloat162float(l1);
      // the 3-term Gram sees hi*hi + 2*hi*lo on the diagonal
      acc = fmaf(f0, f0 + 2.f * g0, acc);
      acc = fmaf(f1, f1 + 2.f * g1, acc);
    } else {
      acc = fmaf(f0, f0, acc);
      acc = fmaf(f1, f1, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    a.norms[b * Mp + p] = acc;
    if (a.stats && valid) {
      double* st = a.stats + (b * (a.m + a.n) + (inA ? loc : a.m + loc)) * RS_COUNT;
      for (int i = 0; i < RS_COUNT; ++i) st[i] = 0.0;
      double dg;
      if (a.kf.family == FAM_RQ) dg = (double)a.kf.const_diag + (double)a.kf.add_dot * (double)acc;
      else if (a.kf.family == FAM_POLY) {
        double bb = (double)a.kf.poly_gamma * (double)acc + (double)a.kf.poly_coef0;
        dg = 1.0;
        for (int i = 0; i < a.kf.degree; ++i) dg *= bb;
      } else dg = (double)diag_value(a.kf, acc);
      st[RS_DIAG] = dg;
    }
  }
}

// column sums of the X block and of the Y block (needed by the add_dot terms): csum[2][dp] double
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* Z, int64_t dpz, int64_t dp, int64_t m,
                                                     int64_t mp, int64_t n, double* csum, int f16) {
  const int64_t c = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int which = blockIdx.y;  // 0 = X, 1 = Y
  const int rl = threadIdx.x >> 5;
  __shared__ double sh[8][33];
  double s = 0.0;
  const int64_t r0 = which ? mp : 0, cnt = which ? n : m;
  if (c < dp)
    for (int64_t r = rl; r < cnt; r += 8)
      s += f16 ? (double)__half2float(reinterpret_cast<const __half*>(Z)[(r0 + r) * dpz + c])
               : (double)__bfloat162float(Z[(r0 + r) * dpz + c]);
  sh[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < dp) {
    for (int i = 1; i < 8; ++i) s += sh[i][threadIdx.x & 31];
    csum[which * dp + c] = s;
  }
}

cudaError_t launch_prep_tc(const PrepTcArgs& a, int64_t rows, unsigned batch, cudaStream_t s) {
  prep_tc_kernel<<<dim3((unsigned)((rows + 7) / 8), batch), 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_colsum_tc(const __nv_bfloat16* Z, int64_t dpz, int64_t dp, int64_t m, int64_t mp, int64_t n,
                             double* csum, int f16, cudaStream_t s) {
  colsum_kernel<<<dim3((unsigned)((dp + 31) / 32), 2), 256, 0, s>>>(Z, dpz, dp, m, mp, n, csum, f16);
  return cudaGetLastError();
}

}  // namespace tc

using namespace tc;

// ------------------------------------------------------------------------------------------------
// public (library-internal) interface
// ------------------------------------------------------------------------------------------------
bool tc_set_option(const char* name, long long value) { return name && tuning_set(tuning_mut(), name, (int64_t)value); }
bool tc_small_kernel_disabled() { return tuning().disable_small != 0; }

bool tc_mmd2_supported(int64_t d, int want_grad) { return d >= 1 && d <= 65536; }

static bool use_two_pass(int64_t d) { return d > tuning().wz_min_d; }

static bool tc_family_ok(const KernelFn& kf) {
  return kf.family == FAM_RBF || kf.family == FAM_RQ || kf.family == FAM_DISTANCE;
}

// Which (family, mode) combinations the tensor-core kernels cover.  dot is closed-form territory
// (sum_ij <x_i,y_j> = <sum x, sum y>) and stays on the exact path.
bool tc_mmd2_covers(const KernelFn& kf, const Geometry& g, int want_grad) {
  if (!tc_family_ok(kf) || !tc_mmd2_supported(g.d, want_grad)) return false;
  if (!want_grad) {
    if (kf.add_dot > 0.f) return false;
    if (!(g.x0 == 0 && g.x1 == g.m && g.y0 == 0 && g.y1 == g.n)) return false;
  }
  return true;
}

size_t tc_mmd2_workspace_bytes(const Geometry& g, int want_grad, int precision) {
  // the plan of exactly this row range (rank / world): slab counts are not monotone in the range, so a full-range
  // plan does not bound a shard's plan
  const size_t grad = !want_grad ? 0
                      : tc_sym_eligible(g) ? tc_sym_workspace_bytes(g)
                                           : (use_two_pass(g.d) ? tc_wz_workspace_bytes(g) : tc_fused_workspace_bytes(g));
  return std::max(grad, tc_value_only_workspace_bytes(g.m, g.n, g.d, precision));
}

cudaError_t tc_mmd2_run(const KernelFn& kf_in, const Geometry& g, const Coefs& c, const SrcLayout& src, int precision,
                        double* scalars, float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches,
                        const char** path) {
  if (!tc_family_ok(kf_in)) return cudaErrorNotSupported;
  KernelFn kf = kf_in;
  TcVariant variant = select_tc_variant(kf);
  if (tuning().null_math) variant = TV_NULL;
  if (dX != nullptr) {
    if (tc_sym_eligible(g)) return tc_run_sym(kf, variant, g, c, src, scalars, dX, dY, ws, ws_bytes, s, launches, path);
    if (use_two_pass(g.d)) return tc_run_wz(kf, variant, g, c, src, scalars, dX, dY, ws, ws_bytes, s, launches, path);
    return tc_run_fused(kf, variant, g, c, src, scalars, dX, dY, ws, ws_bytes, s, launches, path);
  }
  return tc_run_value_only(kf, variant, g, c, src, precision, scalars, dX, dY, ws, ws_bytes, s, launches, path);
}

}  // namespace smmd

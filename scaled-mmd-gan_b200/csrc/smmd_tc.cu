// smmd_tc.cu -- tcgen05 (5th-gen tensor core) path, sm_100a only.
//
//  tc_fused_kernel   : MMD^2 forward AND backward in one sweep over Gram tiles (flash-attention shaped):
//                        S  = Z_i Z_j^T          UMMA #1 (SS, bf16, fp32 accum in TMEM)
//                        W  = 4 a k'(D(S))       epilogue warps: TMEM -> regs -> kernel transform; block sums,
//                                                row sums; W written back to TMEM as bf16
//                        O += W Z_j              UMMA #2 (A = W from TMEM, B = the SAME Z_j smem tile MN-major)
//                      so neither the N x N kernel matrix nor its derivative ever reaches HBM.
//                      Replaces gan/core/mmd.py:55-188 (kernels) + :194-220 (mmd2) + the TF autodiff graph
//                      (gan/core/model.py:446,452) for d <= 256.
//  tc_wgen_kernel    : pass 1 of the wide-feature (d > 256) backward: 128 x 256 Gram tiles with K streamed, the same
//  tc_wz_kernel        epilogue math, bf16 W tiles stored per row panel; pass 2 is O = W Z as a 256 x 256 macro-tile GEMM
//  wz_finalize_rows_kernel  (details at the kernels; pass 1 runs as cta_group::2 CTA pairs on large panels).
//  tc_stream_kernel  : K-streaming 128 x 128 Gram tiles + fused reduction epilogue (row stats): value-only MMD^2.
//  tc_macro_kernel   : 256 x 256 macro-tile Gram + reduction epilogue, batched over problems: KID
//                      (gan/compute_scores.py:232-335, all subsets in one launch) and the 3-sample sums
//                      (gan/core/mmd.py:515-539).  bf16 or split-bf16 (hi*hi + lo*hi + hi*lo) operands.
//
// Work distribution is "stream-K" style: the flattened (row block, column tile) space is cut into equal
// contiguous chunks, one per persistent CTA (grid = #SMs), partial results land in per-(CTA, slot)
// workspace slabs and are reduced in a fixed order by the finalisation kernel (deterministic).
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "sm100_ptx.cuh"
#include "smmd_kfun.cuh"
#include "smmd_tc.h"
#include "smmd_tc_math.cuh"
#include "tmap_host.h"

namespace smmd {
using namespace sm100;

namespace {

constexpr int BM = 128;            // rows per row block (UMMA M)
constexpr int BNF = 64;            // fused kernel: columns per tile
constexpr int kThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 two epilogue groups
constexpr int kMaxSmem = 232448;   // 227 KB

inline int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }

// Developer tuning knobs (environment overrides are read once; defaults are the measured best).
struct Tuning {
  int fused_ksplit;
  int null_math;   // developer ablation (SMMD_DEBUG_NULLMATH=1): results are meaningless
  int fused_lockstep;   // whole row blocks per CTA for large Z (SMMD_FUSED_LOCKSTEP=0 disables)
  int wz_min_d;    // features above which the fused backward switches to the two-pass (W panel + GEMM) path
  int64_t wz_panel_bytes;   // byte budget of one W row panel
  int wz_pair;      // pass 1 as CTA pairs with cta_group::2 UMMAs (SMMD_WZ_PAIR=0 disables)
};
const Tuning& tuning() {
  static Tuning t = [] {
    Tuning v;
    v.fused_ksplit = 2;
    if (const char* e = getenv("SMMD_FUSED_KSPLIT")) v.fused_ksplit = atoi(e) == 1 ? 1 : 2;
    v.null_math = getenv("SMMD_DEBUG_NULLMATH") ? 1 : 0;
    v.fused_lockstep = 1;
    if (const char* e = getenv("SMMD_FUSED_LOCKSTEP")) v.fused_lockstep = atoi(e) != 0;
    v.wz_min_d = 256;
    if (const char* e = getenv("SMMD_WZ_MIN_D")) v.wz_min_d = atoi(e);
    v.wz_pair = 1;
    if (const char* e = getenv("SMMD_WZ_PAIR")) v.wz_pair = atoi(e) != 0;
    v.wz_panel_bytes = (int64_t)6 << 30;
    if (const char* e = getenv("SMMD_WZ_PANEL_MB")) v.wz_panel_bytes = (int64_t)atoll(e) << 20;
    return v;
  }();
  return t;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

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
};

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
  const int64_t ld = inA ? a.lda : a.ldb;
  const void* base = inA ? a.A : a.B;
  __nv_bfloat16* zrow = a.Z + (b * Mp + p) * a.dpz;
  float acc = 0.f;
  // fast path: fp32 rows, 16-B aligned, no split/tanh tail handling needed per element: 8 features per lane
  // per step (2 x LDG.128 -> 1 x STG.128), several independent loads in flight
  const bool vec = a.dtype == SMMD_F32 && !a.split && (ld % 4 == 0) && (a.d % 8 == 0) &&
                   ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
  if (vec) {
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
                                                     int64_t mp, int64_t n, double* csum) {
  const int64_t c = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int which = blockIdx.y;  // 0 = X, 1 = Y
  const int rl = threadIdx.x >> 5;
  __shared__ double sh[8][33];
  double s = 0.0;
  const int64_t r0 = which ? mp : 0, cnt = which ? n : m;
  if (c < dp)
    for (int64_t r = rl; r < cnt; r += 8) s += (double)__bfloat162float(Z[(r0 + r) * dpz + c]);
  sh[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < dp) {
    for (int i = 1; i < 8; ++i) s += sh[i][threadIdx.x & 31];
    csum[which * dp + c] = s;
  }
}

// copy the mixture parameters to shared memory for the generic math variants: sp[0..7]=p0, [8..15]=p1, [16..23]=w
__device__ __forceinline__ void stage_params(const KernelFn& kf, float* sp) {
  if (threadIdx.x < 8) {
    sp[threadIdx.x] = kf.p0[threadIdx.x];
    sp[8 + threadIdx.x] = kf.p1[threadIdx.x];
    sp[16 + threadIdx.x] = kf.w[threadIdx.x];
  }
}

// ---- optional pipeline timing (compile with -DSMMD_PIPE_TIMING; developer builds only) --------------------
#ifdef SMMD_PIPE_TIMING
__device__ unsigned long long g_pipe_dbg[32];
#define PT_DECL(role) const bool pt_on = (blockIdx.x == 0) && (role); long long pt_t = 0
#define PT_BEGIN() do { if (pt_on) pt_t = clock64(); } while (0)
#define PT_END(slot) do { if (pt_on) { long long n_ = clock64(); atomicAdd(&g_pipe_dbg[slot], (unsigned long long)(n_ - pt_t)); pt_t = n_; } } while (0)
#define PT_COUNT(slot) do { if (pt_on) atomicAdd(&g_pipe_dbg[slot], 1ull); } while (0)
#else
#define PT_DECL(role)
#define PT_BEGIN()
#define PT_END(slot)
#define PT_COUNT(slot)
#endif

// ================================================================================================
// fused forward + backward kernel
// ================================================================================================
struct FusedArgs {
  KernelFn kf;
  int64_t m, n, mp, np;
  float c_xx, c_yy, c_xy;      // 4 * a_xx etc. (folded into W)
  const float* norms;          // [Mp]
  int nrb_x, rb_x0, nrb_y, rb_y0;
  int T;                       // column tiles of 64 over the padded stacked matrix
  int dp, npanel, nst;         // padded feature dim, 64-wide panels, Zj ring depth
  int ksplit;                  // epilogue column slices per tile (1 or 2)
  int64_t total_tiles, chunk;
  int slots;
  float* Opart;                // [grid][slots][128][DP]
  float* rpart;                // [grid][slots][2][128]
  double* spart;               // [grid][slots][2][128][2]
};

// TMEM columns: O[dp <= 256] | S0..S2[64 each] | W0,W1[32 each]
constexpr uint32_t TM_O = 0, TM_S = 256, TM_W = 448;
constexpr int kZjRowBytes = BNF * 128;   // one 64-wide panel of a column tile
constexpr int kZiRowBytes = BM * 128;    // one 64-wide panel of the row block

// smem = 1023 B alignment slack + Zi + nst * Zj tile + 512 B (barriers, tmem slot, staged params)
inline int fused_stages(int npanel) {
  int nst = (kMaxSmem - 1024 - 512 - npanel * kZiRowBytes) / (npanel * kZjRowBytes);
  return nst > 8 ? 8 : nst;
}
inline int fused_smem(int npanel, int nst) { return 1024 + npanel * kZiRowBytes + nst * (npanel * kZjRowBytes) + 512; }

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
      wpk[(c >> 1) + e] = pack_bf16x2(ww.x, ww.y);
    }
  }
}

// KSPLIT = column slices per tile: each of the two epilogue groups has 4*KSPLIT warps (TMEM lane quarter x
// column slice).  Warp roles: warps [0, 8*KSPLIT) = epilogue, then the TMA producer, and LAST the UMMA issuer:
// the warp scheduler favours the highest warp id on a sub-partition, and the single issuing thread is on the
// critical path of the whole CTA (measured: as warp 1 it needed ~3100 cycles per tile, most of it waiting for
// issue slots behind the epilogue warps and polling mbarriers at ~100-150 cycles per poll).
//
// mbarriers (all phases tracked with running counters, no div/mod):
//   zj_full[nst]   TMA -> UMMA issuer                      (Zj tile landed)
//   zj_empty[nst]  UMMA #2 commit -> TMA producer AND the epilogue group (its W buffer is drained)
//   s_full[3]      UMMA #1 commit -> epilogue group
//   w_full[2]      epilogue group -> UMMA issuer            (W published; also implies the S buffer is free,
//                                                            because a thread loads S before it writes W)
//   zi_full/zi_empty, o_full/o_empty  per row-block unit
template <class Math, int KSPLIT>
__global__ void __launch_bounds__(64 + 256 * KSPLIT, 1)
tc_fused_kernel(const __grid_constant__ CUtensorMap tmap_zi, const __grid_constant__ CUtensorMap tmap_zj,
                const __grid_constant__ FusedArgs a) {
  constexpr int NPART = 2 * KSPLIT;           // partial-result slices per row (group x column slice)
  constexpr int CH_PER = (BNF / 16) / KSPLIT; // 16-column chunks per thread per tile (4 or 2)
  constexpr int EPI_WARPS = 8 * KSPLIT;
  const int NPANEL = a.npanel, NST = a.nst, DP = a.dp;
  const int ZI_BYTES = NPANEL * kZiRowBytes, ZJ_BYTES = NPANEL * kZjRowBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sZi = smem;
  uint8_t* sZj = smem + ZI_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sZj + NST * ZJ_BYTES);
  uint64_t* zj_full = bars;             // [NST]
  uint64_t* zj_empty = bars + NST;      // [NST]
  uint64_t* s_full = bars + 2 * NST;    // [3]
  uint64_t* w_full = s_full + 3;        // [2]
  uint64_t* zi_full = w_full + 2;
  uint64_t* zi_empty = zi_full + 1;
  uint64_t* o_full = zi_empty + 1;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);
  float* sParams = reinterpret_cast<float*>(tmem_slot + 4);  // [24]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_params(a.kf, sParams);
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&zj_full[i], 1);
      mbar_init(&zj_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&s_full[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&w_full[i], 128 * KSPLIT);
    mbar_init(zi_full, 1);
    mbar_init(zi_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 256 * KSPLIT);
    fence_mbar_init();
  }
  if (warp == EPI_WARPS + 1) tmem_alloc<512>(tmem_slot);
  if (warp == EPI_WARPS && lane == 0) {
    prefetch_tmap(&tmap_zi);
    prefetch_tmap(&tmap_zj);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int64_t pos0 = (int64_t)blockIdx.x * a.chunk;
  const int64_t pos1 = pos0 + a.chunk < a.total_tiles ? pos0 + a.chunk : a.total_tiles;
  auto rb_of = [&](int64_t rbi) -> int { return rbi < a.nrb_x ? a.rb_x0 + (int)rbi : a.rb_y0 + (int)(rbi - a.nrb_x); };

  if (warp == EPI_WARPS) {
    // ===================== TMA producer =====================
    // The WHOLE warp runs this loop convergently and only the issue instructions are predicated on one
    // elected lane: operands then live in uniform registers.  (Issuing from inside `if (lane == 0)` makes the
    // compiler wrap every UTMALDG / UTCHMMA in an ELECT + R2UR "waterfall" loop, ~80 cycles per instruction.)
    {
      uint32_t unit = 0, st = 0, ph = 0;
      int rbi = (int)(pos0 / a.T);
      int t0 = (int)(pos0 - (int64_t)rbi * a.T);
      for (int64_t left = pos1 - pos0; left > 0; ++rbi, t0 = 0, ++unit) {
        const int TU = (int)std::min<int64_t>(a.T - t0, left);
        const int rb = rb_of(rbi);
        mbar_wait(zi_empty, (unit & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(zi_full, ZI_BYTES);
          for (int p = 0; p < NPANEL; ++p) tma_load_2d(sZi + p * (BM * 128), &tmap_zi, zi_full, p * 64, rb * BM);
        }
        __syncwarp();
        for (int t = t0; t < t0 + TU; ++t) {
          mbar_wait(&zj_empty[st], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&zj_full[st], ZJ_BYTES);
            uint8_t* dst = sZj + st * ZJ_BYTES;
            for (int p = 0; p < NPANEL; ++p) tma_load_2d(dst + p * (BNF * 128), &tmap_zj, &zj_full[st], p * 64, t * BNF);
          }
          __syncwarp();
          if (++st == (uint32_t)NST) {
            st = 0;
            ph ^= 1;
          }
        }
        left -= TU;
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ===================== UMMA issuer (warp-convergent loop, one elected lane issues) ========================
    {
      constexpr uint32_t idesc1 = make_idesc(BM, BNF, kFmtBF16, false, false);
      const uint32_t idesc2 = make_idesc(BM, (uint32_t)DP, kFmtBF16, false, true);
      const uint32_t hi = desc_hi_sw128(1024);
      const uint32_t zi_lo = desc_lo(smem_u32(sZi), 16);                   // K-major A: LBO unused (16 B)
      const uint32_t zj_lo1 = desc_lo(smem_u32(sZj), 16);                  // K-major B for UMMA #1
      const uint32_t zj_lo2 = desc_lo(smem_u32(sZj), BNF * 128);           // MN-major B for UMMA #2: LBO = panel stride
      const uint32_t stage_step = (uint32_t)ZJ_BYTES >> 4;                 // descriptor address units are 16 B
      uint32_t unit = 0;
      uint32_t st1 = 0, ph1 = 0, sb1 = 0;             // UMMA #1 stream: Zj stage / phase, S buffer
      uint32_t st2 = 0, wb2 = 0, wph2 = 0;            // UMMA #2 stream: Zj stage, W buffer / phase
      int rbi = (int)(pos0 / a.T);
      int t0 = (int)(pos0 - (int64_t)rbi * a.T);
      PT_DECL(lane == 0);
      for (int64_t left = pos1 - pos0; left > 0; ++rbi, t0 = 0, ++unit) {
        const int TU = (int)std::min<int64_t>(a.T - t0, left);
        mbar_wait(zi_full, unit & 1);
        // static order, UMMA #1 three tiles ahead of UMMA #2; 2 waits + 2 commits per tile
        for (int jj = 0; jj < TU + 3; ++jj) {
          const int b2 = jj - 3;
          if (b2 >= 0) {  // ---- UMMA #2 for local tile b2: O += W * Zj
            PT_BEGIN();
            mbar_wait(&w_full[wb2], wph2);
            if (b2 == 0) mbar_wait(o_empty, (unit & 1) ^ 1);
            tc_fence_after();
            PT_END(0);
            const uint32_t blo = zj_lo2 + st2 * stage_step;
            const uint32_t wad = tmem + TM_W + wb2 * 32;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < BNF / 16; ++kk)
                umma_ts2(tmem + TM_O, wad + kk * 8, blo + kk * (2048 >> 4), hi, idesc2, (b2 > 0 || kk > 0) ? 1u : 0u);
              umma_commit(&zj_empty[st2]);     // frees the Zj stage (producer) and this W buffer (epilogue group)
              if (b2 == TU - 1) umma_commit(o_full);
            }
            __syncwarp();
            if (++st2 == (uint32_t)NST) st2 = 0;
            wph2 ^= wb2;  // phase flips each time the buffer index wraps 1 -> 0
            wb2 ^= 1;
            PT_END(1);
            PT_COUNT(3);
          }
          if (jj < TU) {  // ---- UMMA #1 for local tile jj: S = Zi * Zj^T  (S buffer is free: see w_full above)
            PT_BEGIN();
            mbar_wait(&zj_full[st1], ph1);
            tc_fence_after();
            PT_END(4);
            const uint32_t blo = zj_lo1 + st1 * stage_step;
            const uint32_t sad = tmem + TM_S + sb1 * 64;
            if (elect_one()) {
              for (int p = 0; p < NPANEL; ++p) {
                const uint32_t ap = zi_lo + p * ((BM * 128) >> 4), bp = blo + p * ((BNF * 128) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss2(sad, ap + k * 2, bp + k * 2, hi, idesc1, (p | k) ? 1u : 0u);
              }
              umma_commit(&s_full[sb1]);
              if (jj == TU - 1) umma_commit(zi_empty);
            }
            __syncwarp();
            if (++st1 == (uint32_t)NST) {
              st1 = 0;
              ph1 ^= 1;
            }
            if (++sb1 == 3) sb1 = 0;
            PT_END(2);
          }
        }
        left -= TU;
      }
    }
  } else {
    // ===================== epilogue groups =====================
    const int grp = warp / (4 * KSPLIT);      // 0 / 1
    const int half = (warp % (4 * KSPLIT)) >> 2;  // column slice of the tile handled by this warp
    const int part = grp * KSPLIT + half;
    const int q = warp & 3;                   // TMEM lane quarter this warp may touch
    const int r = q * 32 + lane;              // row inside the row block
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const Math math(a.kf, sParams);
    const float kscale = math.k_scale(), kdscale = math.kd_scale();
    uint32_t unit = 0;
    int slot = 0;
    // running ring state (this group handles every second tile of the CTA's stream)
    uint32_t par = 0;                       // parity of the global tile counter
    uint32_t sb = 0, sph = 0;               // S buffer / phase of the current tile
    uint32_t st = 0, ph = 0;                // Zj stage / phase of the current tile
    uint32_t st_m2 = 0, ph_m2 = 0;          // ... and of the tile two back (whose UMMA #2 drains this group's W buffer)
    uint32_t gcount = 0;                    // global tile counter (only its first two values matter)
    int rbi = (int)(pos0 / a.T);
    int t0 = (int)(pos0 - (int64_t)rbi * a.T);
    const int mp = (int)a.mp, mvalid = (int)a.m, yvalid = (int)(a.mp + a.n);
    PT_DECL(warp == 0 && lane == 0);
    for (int64_t left = pos1 - pos0; left > 0; ++rbi, t0 = 0, ++unit, ++slot) {
      const int TU = (int)std::min<int64_t>(a.T - t0, left);
      const int rb = rb_of(rbi);
      const int gi = rb * BM + r;
      const bool rowX = gi < mp;
      const float ni = a.norms[gi];
      float2 rsum = make_float2(0.f, 0.f);
      double dsame = 0.0, dcross = 0.0;
      for (int lt = 0; lt < TU; ++lt) {
        if ((int)par == grp) {
          PT_BEGIN();
          const int c0 = (t0 + lt) * BNF;
          const bool colX = c0 < mp;
          const bool same = (colX == rowX);
          const float2 cw = bc2((same ? (rowX ? a.c_xx : a.c_yy) : a.c_xy) * kdscale);
          const int lim = colX ? mvalid : yvalid;                       // first invalid column of this region
          const bool special = (c0 + BNF > lim) || ((c0 >> 7) == rb);   // pad columns or diagonal inside
          const float* nj = a.norms + c0 + half * (CH_PER * 16);        // column norms: tiny, L1/L2 resident
          const uint32_t s_addr = tmem + TM_S + sb * 64 + half * (CH_PER * 16) + lane_base;
          const uint32_t w_addr = tmem + TM_W + grp * 32 + half * (CH_PER * 8) + lane_base;
          float2 tsum = make_float2(0.f, 0.f);
          mbar_wait(&s_full[sb], sph);
          tc_fence_after();
          PT_END(9);
          if (!special) {
            // two 16-column chunks per iteration, the next tcgen05.ld in flight while the current chunk is computed
            uint32_t va[16], vb[16], wpk[8];
            tmem_ld_x16(s_addr, va);
#pragma unroll 1
            for (int it = 0; it < CH_PER / 2; ++it) {
              tmem_ld_wait();
              tmem_ld_x16(s_addr + (2 * it + 1) * 16, vb);
              fused_chunk16<Math, false>(math, va, nj + (2 * it) * 16, ni, cw, 0, 0, 0, tsum, rsum, wpk);
              if (it == 0 && gcount >= 2) mbar_wait(&zj_empty[st_m2], ph_m2);   // W buffer drained by UMMA #2 of tile-2
              tmem_st_x8(w_addr + (2 * it) * 8, wpk);
              tmem_ld_wait();
              if (2 * it + 2 < CH_PER) tmem_ld_x16(s_addr + (2 * it + 2) * 16, va);
              fused_chunk16<Math, false>(math, vb, nj + (2 * it + 1) * 16, ni, cw, 0, 0, 0, tsum, rsum, wpk);
              tmem_st_x8(w_addr + (2 * it + 1) * 8, wpk);
            }
          } else {
#pragma unroll 1
            for (int ch = 0; ch < CH_PER; ++ch) {
              uint32_t v[16], wpk[8];
              tmem_ld_x16(s_addr + ch * 16, v);
              tmem_ld_wait();
              fused_chunk16<Math, true>(math, v, nj + ch * 16, ni, cw, c0 + half * (CH_PER * 16) + ch * 16, lim, gi, tsum,
                                        rsum, wpk);
              if (ch == 0 && gcount >= 2) mbar_wait(&zj_empty[st_m2], ph_m2);
              tmem_st_x8(w_addr + ch * 8, wpk);
            }
          }
          PT_END(11);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&w_full[grp]);
          PT_END(13);
          PT_COUNT(14);
          if (same) dsame += (double)((tsum.x + tsum.y) * kscale);
          else dcross += (double)((tsum.x + tsum.y) * kscale);
        }
        // advance the ring state by one tile of the CTA's stream
        par ^= 1;
        if (gcount >= 2) {
          if (++st_m2 == (uint32_t)NST) {
            st_m2 = 0;
            ph_m2 ^= 1;
          }
        }
        ++gcount;
        if (++st == (uint32_t)NST) {
          st = 0;
          ph ^= 1;
        }
        if (++sb == 3) {
          sb = 0;
          sph ^= 1;
        }
      }
      // ---- unit end: drain O (this thread's slice of the feature columns) ----
      mbar_wait(o_full, unit & 1);
      tc_fence_after();
      {
        const int64_t sl = (int64_t)blockIdx.x * a.slots + slot;
        const int seg = DP / NPART;  // feature columns drained by this thread (multiple of 16)
        float* orow = a.Opart + (sl * BM + r) * DP + part * seg;
        for (int c = 0; c < seg; c += 16) {
          uint32_t v[16];
          tmem_ld_x16(tmem + TM_O + part * seg + c + lane_base, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            *reinterpret_cast<float4*>(orow + c + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                   __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
        }
        a.rpart[(sl * NPART + part) * BM + r] = rsum.x + rsum.y;
        double* sp = a.spart + ((sl * NPART + part) * BM + r) * 2;
        sp[0] = dsame;
        sp[1] = dcross;
      }
      tc_fence_before();
      mbar_arrive(o_empty);
      left -= TU;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) tmem_dealloc<512>(tmem);
}

// ---- finalisation of the fused kernel: reduce slabs, form gradients and per-row stats ---------------
struct FinRowsArgs {
  KernelFn kf;
  int64_t m, n, mp, np, d;
  int64_t x0, ox, y0, oy;
  int dp;
  int nrb_x, rb_x0, nrb_y, rb_y0, T;
  int64_t chunk;
  int slots, npart;
  double a_xx, a_yy, a_xy;
  const __nv_bfloat16* Z;
  int64_t dpz;
  SrcLayout src;       // original features: the r_i * z_i term uses the unrounded row (fp32 owned rows if given)
  const float* norms;
  const double* csum;  // [2][dp] or null
  const float* Opart;
  const float* rpart;
  const double* spart;
  float* dX;
  float* dY;
  double* partials;  // [gridDim.x][6] per-CTA block sums (second stage: launch_finalize_partials)
};

constexpr int kFinRowsPerWarp = 4;
constexpr int kFinRowsPerCta = 8 * kFinRowsPerWarp;

// One warp per row (4 rows per warp): reduce the per-(CTA, slot) slabs in fixed order, form the gradient row
// with the fp32 z_i, and fold the row's block sums into per-CTA partials (second stage: finalize_partials).
__global__ void __launch_bounds__(256) tc_finalize_rows_kernel(FinRowsArgs a) {
  constexpr int NT = 2;   // 128-feature groups per row: dp <= 256
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double sh[8][6];
  double q[6] = {0, 0, 0, 0, 0, 0};  // sxx, syy, sxy, syx, dgx, dgy of this warp's rows
  const double a_xy = a.a_xy;
  const bool dot = a.kf.family == FAM_RQ && a.kf.add_dot > 0.f && a.csum != nullptr;
  for (int rr = 0; rr < kFinRowsPerWarp; ++rr) {
    const int64_t lr = ((int64_t)blockIdx.x * 8 + warp) * kFinRowsPerWarp + rr;
    if (lr >= a.ox + a.oy) break;
    const bool rowX = lr < a.ox;
    const int64_t li = rowX ? a.x0 + lr : a.y0 + (lr - a.ox);   // index inside X or Y
    const int64_t gi = rowX ? li : a.mp + li;                   // padded stacked row
    const int rb = (int)(gi / BM), r = (int)(gi % BM);
    const int64_t rbi = rowX ? rb - a.rb_x0 : a.nrb_x + (rb - a.rb_y0);
    const int64_t f0 = rbi * a.T, f1 = f0 + a.T - 1;
    const int64_t g0 = f0 / a.chunk, g1 = f1 / a.chunk;
    float rs = 0.f;
    double ssame = 0.0, scross = 0.0;
    for (int64_t g = g0; g <= g1; ++g) {
      const int64_t sl = g * a.slots + (rbi - (g * a.chunk) / a.T);
      for (int pt = 0; pt < a.npart; ++pt) {
        rs += a.rpart[(sl * a.npart + pt) * BM + r];
        const double* sp = a.spart + ((sl * a.npart + pt) * BM + r) * 2;
        ssame += sp[0];
        scross += sp[1];
      }
    }
    const double a_same = rowX ? a.a_xx : a.a_yy;
    double dsame = 0.0, dcross = 0.0;  // z_i . colsum(same set) / (other set)
    float* out = nullptr;
    if (a.dX) out = rowX ? a.dX + (li - a.x0) * a.d : a.dY + (li - a.y0) * a.d;
    // source of z_i: the fp32 owned rows when the caller supplied them, else the (possibly gathered) inputs
    const bool owned = (rowX ? a.src.Xo : a.src.Yo) != nullptr;
    const void* src = owned ? static_cast<const void*>(rowX ? a.src.Xo : a.src.Yo) : (rowX ? a.src.X : a.src.Y);
    const int64_t ld = owned ? a.src.ldo : (rowX ? a.src.ldx : a.src.ldy);
    const int sdtype = owned ? (int)SMMD_F32 : a.src.dtype;
    const int64_t srow = owned ? (rowX ? li - a.x0 : li - a.y0) : src_row(li, rowX, a.src.blk_x, a.src.blk_y);
    const bool vec = out != nullptr && !dot && sdtype == SMMD_F32 && (a.d % 4 == 0) && (ld % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec) {
      // lane owns features [4 lane + 128 t, +4), t < NT
      float4 oacc[NT], z4[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) oacc[t] = z4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* zsrc = reinterpret_cast<const float*>(src) + srow * ld;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int c = 4 * lane + 128 * t;
        if (c < a.d) z4[t] = *reinterpret_cast<const float4*>(zsrc + c);
      }
      for (int64_t g = g0; g <= g1; ++g) {
        const int64_t sl = g * a.slots + (rbi - (g * a.chunk) / a.T);
        const float* orow = a.Opart + (sl * BM + r) * a.dp;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int c = 4 * lane + 128 * t;
          if (c < a.dp) {
            const float4 o = *reinterpret_cast<const float4*>(orow + c);
            oacc[t].x += o.x;
            oacc[t].y += o.y;
            oacc[t].z += o.z;
            oacc[t].w += o.w;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int c = 4 * lane + 128 * t;
        if (c < a.d) {
          float zz[4] = {z4[t].x, z4[t].y, z4[t].z, z4[t].w};
          const float oo[4] = {oacc[t].x, oacc[t].y, oacc[t].z, oacc[t].w};
          float gv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (a.kf.tanh_features) zz[e] = tanhf(zz[e]);
            gv[e] = rs * zz[e] - oo[e];                      // W already carries the factor 4 a_ij
            if (a.kf.tanh_features) gv[e] *= (1.f - zz[e] * zz[e]);
          }
          *reinterpret_cast<float4*>(out + c) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        }
      }
    } else {
      // general path: lane owns features lane, lane+32, ... (at most 4 NT)
      float oacc[4 * NT];
#pragma unroll
      for (int t = 0; t < 4 * NT; ++t) oacc[t] = 0.f;
      if (out) {
        for (int64_t g = g0; g <= g1; ++g) {
          const int64_t sl = g * a.slots + (rbi - (g * a.chunk) / a.T);
          const float* orow = a.Opart + (sl * BM + r) * a.dp;
#pragma unroll
          for (int t = 0; t < 4 * NT; ++t) {
            const int c = lane + 32 * t;
            if (c < a.dp) oacc[t] += orow[c];
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 4 * NT; ++t) {
        const int c = lane + 32 * t;
        if (c >= a.d) continue;
        // z_i at full input precision: g_i = 4 sum_j W_ij (z_i - z_j) is dominated by r_i z_i, so rounding
        // z_i to bf16 here would put a 2^-9 relative error straight into the gradient
        const int64_t sidx = srow * ld + c;
        float z = sdtype == SMMD_F32 ? reinterpret_cast<const float*>(src)[sidx]
                                      : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[sidx]);
        if (a.kf.tanh_features) z = tanhf(z);
        if (out) {
          float gv = rs * z - oacc[t];
          if (dot) {
            const double cs = a.csum[(rowX ? 0 : 1) * a.dp + c], co = a.csum[(rowX ? 1 : 0) * a.dp + c];
            gv += (float)(2.0 * (double)a.kf.add_dot * (a_same * cs + a_xy * co));
          }
          if (a.kf.tanh_features) gv *= (1.f - z * z);
          out[c] = gv;
        }
        if (dot) {
          dsame += (double)z * a.csum[(rowX ? 0 : 1) * a.dp + c];
          dcross += (double)z * a.csum[(rowX ? 1 : 0) * a.dp + c];
        }
      }
      if (dot) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          dsame += __shfl_xor_sync(0xffffffffu, dsame, o);
          dcross += __shfl_xor_sync(0xffffffffu, dcross, o);
        }
      }
    }
    // row totals (lane-uniform values); the dot part of the kernel is closed form:
    //   sum_{j != i} <z_i, z_j> = <z_i, colsum> - |z_i|^2
    const float ni = a.norms[gi];
    const double v_same = ssame + (dot ? (double)a.kf.add_dot * (dsame - (double)ni) : 0.0);
    const double v_cross = scross + (dot ? (double)a.kf.add_dot * dcross : 0.0);
    const double v_diag = a.kf.family == FAM_RQ ? (double)a.kf.const_diag + (double)a.kf.add_dot * (double)ni
                                                : (double)diag_value(a.kf, ni);
    if (rowX) {
      q[0] += v_same;
      q[2] += v_cross;
      q[4] += v_diag;
    } else {
      q[1] += v_same;
      q[3] += v_cross;
      q[5] += v_diag;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) sh[warp][i] = q[i];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];   // fixed order
    a.partials[(int64_t)blockIdx.x * 6 + threadIdx.x] = t;
  }
}

// ================================================================================================
// K-streaming Gram + reduction epilogue (KID, value-only MMD^2)
// ================================================================================================
constexpr int BNS = 128;
constexpr int kStreamStages = 6;
constexpr int kStreamStageBytes = 2 * BM * 128;  // A panel + B panel
constexpr int kStreamSmem = 1024 + kStreamStages * kStreamStageBytes + 1024;

struct StreamArgs {
  KernelFn kf;
  int64_t m, n, mp, np;   // per problem
  int RB, CT;             // row blocks / column tiles (128) per problem
  int nkp;                // 64-wide k-panels of the operand (dp/64)
  int ncombo;             // 1 (bf16) or 3 (split)
  int64_t dp;
  int64_t total_tiles, chunk;
  const float* norms;     // [batch][Mp]
  double* stats;          // [batch][m+n][RS_COUNT]
  int want_sq;
};

template <class Math>
__global__ void __launch_bounds__(kThreads, 1)
tc_stream_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ StreamArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStreamStages * kStreamStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStreamStages;
  uint64_t* acc_full = empty + kStreamStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sParams = reinterpret_cast<float*>(tmem_slot + 4);  // [24]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_params(a.kf, sParams);
  if (tid == 0) {
    for (int i = 0; i < kStreamStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  if (warp == 0 && lane == 0) prefetch_tmap(&tmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t Mp = a.mp + a.np;
  // tiles are dealt round-robin (tile t -> CTA t mod grid): the CTAs work on ~grid consecutive tiles at any time
  // (one row block, neighbouring column tiles), so the streamed operands stay L2 resident for any problem size
  const int nk = a.nkp * a.ncombo;

  if (warp == 0) {
    {  // whole warp, elected lane issues (see tc_fused_kernel)
      uint32_t st = 0, ph = 0;
      const int dpi = (int)a.dp;
      for (int64_t pos = blockIdx.x; pos < a.total_tiles; pos += gridDim.x) {
        const int ct = (int)(pos % a.CT);
        const int64_t brb = pos / a.CT;
        const int rb = (int)(brb % a.RB);
        const int64_t b = brb / a.RB;
        const int32_t arow = (int32_t)(b * Mp) + rb * BM, brow = (int32_t)(b * Mp) + ct * BNS;
        for (int combo = 0; combo < a.ncombo; ++combo) {
          const int32_t aoff = combo == 1 ? dpi : 0, boff = combo == 2 ? dpi : 0;
          for (int p = 0; p < a.nkp; ++p) {
            mbar_wait(&empty[st], ph ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full[st], kStreamStageBytes);
              uint8_t* sa = smem + st * kStreamStageBytes;
              tma_load_2d(sa, &tmap, &full[st], p * 64 + aoff, arow);
              tma_load_2d(sa + BM * 128, &tmap, &full[st], p * 64 + boff, brow);
            }
            __syncwarp();
            if (++st == kStreamStages) {
              st = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = make_idesc(BM, BNS, kFmtBF16, false, false);
      const uint32_t hi = desc_hi_sw128(1024);
      const uint32_t a_lo0 = desc_lo(smem_u32(smem), 16);
      uint32_t st = 0, ph = 0, ab = 0, aph = 0;
      for (int64_t pos = blockIdx.x; pos < a.total_tiles; pos += gridDim.x) {
        mbar_wait(&acc_empty[ab], aph ^ 1);
        tc_fence_after();
        const uint32_t dad = tmem + ab * BNS;
        for (int kk = 0; kk < nk; ++kk) {
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const uint32_t alo = a_lo0 + st * (kStreamStageBytes >> 4), blo = alo + ((BM * 128) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss2(dad, alo + k * 2, blo + k * 2, hi, idesc, (kk | k) ? 1u : 0u);
            umma_commit(&empty[st]);
          }
          __syncwarp();
          if (++st == kStreamStages) {
            st = 0;
            ph ^= 1;
          }
        }
        if (elect_one()) umma_commit(&acc_full[ab]);
        __syncwarp();
        aph ^= ab;
        ab ^= 1;
      }
    }
  } else {
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const Math math(a.kf, sParams);
    const float kscale = math.k_scale();
    // accumulators of the (problem, row block) currently being swept by this thread
    int64_t cur_brb = -1;
    double s_same = 0, s_cross = 0, q_same = 0, q_cross = 0, pairv = 0;
    float ni = 0.f;
    auto flush = [&]() {
      if (cur_brb < 0) return;
      const int rb = (int)(cur_brb % a.RB);
      const int64_t b = cur_brb / a.RB;
      const int64_t gi = (int64_t)rb * BM + r;
      const bool rowX = gi < a.mp;
      const int64_t loc = rowX ? gi : gi - a.mp;
      if (loc < (rowX ? a.m : a.n)) {
        double* st = a.stats + (b * (a.m + a.n) + (rowX ? loc : a.m + loc)) * RS_COUNT;
        if (s_same != 0.0) atomicAdd(st + RS_SAME, s_same);
        if (s_cross != 0.0) atomicAdd(st + RS_CROSS, s_cross);
        if (a.want_sq) {
          if (q_same != 0.0) atomicAdd(st + RS_SQ_SAME, q_same);
          if (q_cross != 0.0) atomicAdd(st + RS_SQ_CROSS, q_cross);
          if (pairv != 0.0) atomicAdd(st + RS_PAIR, pairv);
        }
      }
      s_same = s_cross = q_same = q_cross = pairv = 0.0;
    };
    uint64_t tc = 0;
    for (int64_t pos = blockIdx.x; pos < a.total_tiles; pos += gridDim.x, ++tc) {
      if ((int)(tc & 1) != grp) continue;
      const int ct = (int)(pos % a.CT);
      const int64_t brb = pos / a.CT;
      const int rb = (int)(brb % a.RB);
      const int64_t b = brb / a.RB;
      if (brb != cur_brb) {
        flush();
        cur_brb = brb;
        ni = a.norms[b * Mp + (int64_t)rb * BM + r];
      }
      const int64_t gi = (int64_t)rb * BM + r;
      const bool rowX = gi < a.mp;
      const int64_t c0 = (int64_t)ct * BNS;
      const bool colX = c0 < a.mp;
      const bool same = (colX == rowX);
      const int64_t lim = colX ? a.m : a.mp + a.n;
      const int64_t pair_col = rowX ? a.mp + gi : -1;  // the (x_i, y_i) element
      const bool special = (c0 + BNS > lim) || (ct == rb) || (rowX && ct == rb + (int)(a.mp / BNS));
      mbar_wait(&acc_full[grp], (uint32_t)((tc >> 1) & 1));
      tc_fence_after();
      const float* nj = a.norms + b * Mp + c0;
      float2 tsum = make_float2(0.f, 0.f), tsq = make_float2(0.f, 0.f);
      const float2 ni2 = bc2(ni), ks2 = bc2(kscale);
#pragma unroll 1
      for (int h = 0; h < BNS / 16; ++h) {
        uint32_t v[16];
        tmem_ld_x16(tmem + grp * BNS + h * 16 + lane_base, v);
        tmem_ld_wait();
        if (h == BNS / 16 - 1) {
          tc_fence_before();
          mbar_arrive(&acc_empty[grp]);
        }
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const float4 n4 = __ldg(reinterpret_cast<const float4*>(nj + h * 16 + c));
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const float2 S = make_float2(__uint_as_float(v[c + e]), __uint_as_float(v[c + e + 1]));
            const float2 nn = e == 0 ? make_float2(n4.x, n4.y) : make_float2(n4.z, n4.w);
            float2 k, kd;
            math.eval2(S, add2(ni2, nn), k, kd);
            k = mul2(k, ks2);
            if (special) {
              const int64_t col = c0 + h * 16 + c + e;
              const bool ok0 = (col < lim) && (col != gi), ok1 = (col + 1 < lim) && (col + 1 != gi);
              k = make_float2(ok0 ? k.x : 0.f, ok1 ? k.y : 0.f);
              if (col == pair_col) pairv = (double)k.x;
              if (col + 1 == pair_col) pairv = (double)k.y;
            }
            tsum = add2(tsum, k);
            tsq = fma2(k, k, tsq);
          }
        }
      }
      if (same) {
        s_same += (double)(tsum.x + tsum.y);
        q_same += (double)(tsq.x + tsq.y);
      } else {
        s_cross += (double)(tsum.x + tsum.y);
        q_cross += (double)(tsq.x + tsq.y);
      }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// ================================================================================================
// 256 x 256 macro-tile Gram kernel for light epilogues (KID's cubic polynomial)
// ================================================================================================
// The K-streaming 128 x 128 kernel above needs 32 KB of operands per 256 tensor cycles per SM (128 B/clk),
// ~3.4x what L2 delivers to 148 SMs at once, so KID (d = 2048, x3 for the split-bf16 Gram) ran L2-bound.
// Here one CTA owns a 256 x 256 block of the stacked Gram: per 64-wide k-panel it loads 2 x 16 KB of rows and
// 32 KB of columns (64 KB per 1024 tensor cycles = 64 B/clk) and issues 8 UMMAs 128x256x16 into two 256-column
// TMEM accumulators (all 512 columns, single-buffered: the cubic epilogue is ~3% of a tile's tensor time).
// When only block totals are needed (ret_var = False, the scorer's default, compute_scores.py:290-300) the
// symmetry of the stacked Gram is used: only macro tiles J >= I are computed, same-set tiles above the
// diagonal count twice, the Y x X mirror of the cross block is skipped.
constexpr int BMAC = 256;
constexpr int kMacStages = 3;
constexpr int kMacStageBytes = 2 * BM * 128 + BMAC * 128;   // 2 row panels + 1 column panel = 64 KB
constexpr int kMacSmem = 1024 + kMacStages * kMacStageBytes + 1024;

struct MacroArgs {
  KernelFn kf;
  int64_t m, n, mp, np;   // per problem; mp, np multiples of 256
  int R, Rx;              // macro tiles per side of the stacked matrix / of the X block
  int sym;                // 1: upper triangle with weights (totals only); 0: all R*R tiles (row statistics)
  int tiles_per_batch;
  int nkp, ncombo;
  int64_t dp;
  int64_t total_tiles, chunk;
  const float* norms;     // [batch][Mp]
  double* stats;          // [batch][m+n][RS_COUNT]
  int want_sq;
};

__device__ __forceinline__ void macro_decode(const MacroArgs& a, int t, int& I, int& J) {
  if (!a.sym) {
    I = t / a.R;
    J = t - I * a.R;
    return;
  }
  int i = 0, cnt = a.R;
  while (t >= cnt) {   // row i of the upper triangle holds R - i tiles; R <= 64
    t -= cnt;
    ++i;
    --cnt;
  }
  I = i;
  J = i + t;
}

template <class Math>
__global__ void __launch_bounds__(kThreads, 1)
tc_macro_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ MacroArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kMacStages * kMacStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMacStages;
  uint64_t* acc_full = empty + kMacStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  float* sParams = reinterpret_cast<float*>(tmem_slot + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_params(a.kf, sParams);
  if (tid == 0) {
    for (int i = 0; i < kMacStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 256);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  if (warp == 8 && lane == 0) prefetch_tmap(&tmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t Mp = a.mp + a.np;
  // Tiles are dealt round-robin (tile t -> CTA t mod grid): at any time the CTAs work on ~grid consecutive tiles,
  // i.e. on a handful of problems whose operands (16.8 MB per KID subset) stay L2 resident.  Contiguous chunks per
  // CTA had every CTA inside a different subset: all operand traffic came from HBM (22 GB per KID call).
  const int nk = a.nkp * a.ncombo;

  if (warp == 8) {
    // ---- TMA producer (whole warp, elected lane issues) ----
    uint32_t st = 0, ph = 0;
    const int dpi = (int)a.dp;
    for (int64_t pos = blockIdx.x; pos < a.total_tiles; pos += gridDim.x) {
      const int64_t b = pos / a.tiles_per_batch;
      int I, J;
      macro_decode(a, (int)(pos - b * a.tiles_per_batch), I, J);
      const int32_t arow = (int32_t)(b * Mp) + I * BMAC, brow = (int32_t)(b * Mp) + J * BMAC;
      for (int combo = 0; combo < a.ncombo; ++combo) {
        const int32_t aoff = combo == 1 ? dpi : 0, boff = combo == 2 ? dpi : 0;
        for (int p = 0; p < a.nkp; ++p) {
          mbar_wait(&empty[st], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[st], kMacStageBytes);
            uint8_t* sa = smem + st * kMacStageBytes;
            tma_load_2d(sa, &tmap, &full[st], p * 64 + aoff, arow);
            tma_load_2d(sa + BM * 128, &tmap, &full[st], p * 64 + aoff, arow + BM);
            tma_load_2d(sa + 2 * BM * 128, &tmap, &full[st], p * 64 + boff, brow);
            tma_load_2d(sa + 3 * BM * 128, &tmap, &full[st], p * 64 + boff, brow + BM);
          }
          __syncwarp();
          if (++st == kMacStages) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 9) {
    // ---- UMMA issuer ----
    constexpr uint32_t idesc = make_idesc(BM, BMAC, kFmtBF16, false, false);
    const uint32_t hi = desc_hi_sw128(1024);
    const uint32_t base_lo = desc_lo(smem_u32(smem), 16);
    uint32_t st = 0, ph = 0, aph = 0;
    for (int64_t pos = blockIdx.x; pos < a.total_tiles; pos += gridDim.x) {
      mbar_wait(acc_empty, aph ^ 1);
      tc_fence_after();
      for (int kk = 0; kk < nk; ++kk) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a0 = base_lo + st * (kMacStageBytes >> 4), a1 = a0 + ((BM * 128) >> 4), bl = a0 + ((2 * BM * 128) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_ss2(tmem, a0 + k * 2, bl + k * 2, hi, idesc, (kk | k) ? 1u : 0u);
            umma_ss2(tmem + BMAC, a1 + k * 2, bl + k * 2, hi, idesc, (kk | k) ? 1u : 0u);
          }
          umma_commit(&empty[st]);
        }
        __syncwarp();
        if (++st == kMacStages) {
          st = 0;
          ph ^= 1;
        }
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
      aph ^= 1;
    }
  } else {
    // ---- epilogue: warp w -> row half w/4 of the macro tile, TMEM lane quarter w%4 ----
    const int hrow = warp >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const Math math(a.kf, sParams);
    const float kscale = math.k_scale();
    int64_t cur_key = -1;   // (batch, macro row) currently accumulated by this thread
    double s_same = 0, s_cross = 0, q_same = 0, q_cross = 0, pairv = 0;
    float ni = 0.f;
    int64_t cur_b = 0;
    int cur_I = 0;
    auto flush = [&]() {
      if (cur_key < 0) return;
      const int64_t gi = (int64_t)cur_I * BMAC + hrow * BM + r;
      const bool rowX = gi < a.mp;
      const int64_t loc = rowX ? gi : gi - a.mp;
      if (loc < (rowX ? a.m : a.n)) {
        double* st = a.stats + (cur_b * (a.m + a.n) + (rowX ? loc : a.m + loc)) * RS_COUNT;
        if (s_same != 0.0) atomicAdd(st + RS_SAME, s_same);
        if (s_cross != 0.0) atomicAdd(st + RS_CROSS, s_cross);
        if (a.want_sq) {
          if (q_same != 0.0) atomicAdd(st + RS_SQ_SAME, q_same);
          if (q_cross != 0.0) atomicAdd(st + RS_SQ_CROSS, q_cross);
          if (pairv != 0.0) atomicAdd(st + RS_PAIR, pairv);
        }
      }
      s_same = s_cross = q_same = q_cross = pairv = 0.0;
    };
    uint32_t fph = 0;
    for (int64_t pos = blockIdx.x; pos < a.total_tiles; pos += gridDim.x) {
      const int64_t b = pos / a.tiles_per_batch;
      int I, J;
      macro_decode(a, (int)(pos - b * a.tiles_per_batch), I, J);
      const int64_t key = b * a.R + I;
      if (key != cur_key) {
        flush();
        cur_key = key;
        cur_b = b;
        cur_I = I;
        ni = a.norms[b * Mp + (int64_t)I * BMAC + hrow * BM + r];
      }
      const int64_t gi = (int64_t)I * BMAC + hrow * BM + r;
      const bool rowX = gi < a.mp;
      const int64_t c0 = (int64_t)J * BMAC;
      const bool colX = c0 < a.mp;
      const bool same = (colX == rowX);
      const int64_t lim = colX ? a.m : a.mp + a.n;
      const int64_t pair_col = rowX ? a.mp + gi : -1;
      const bool special = (c0 + BMAC > lim) || (I == J) || (rowX && J == I + a.Rx);
      const float wgt = (a.sym && same && J > I) ? 2.f : 1.f;
      mbar_wait(acc_full, fph);
      fph ^= 1;
      tc_fence_after();
      const float* nj = a.norms + b * Mp + c0;
      float2 tsum = make_float2(0.f, 0.f), tsq = make_float2(0.f, 0.f);
      const float2 ni2 = bc2(ni), ks2 = bc2(kscale);
      const uint32_t acc = tmem + hrow * BMAC + lane_base;
      uint32_t va[16], vb[16];
      tmem_ld_x16(acc, va);
#pragma unroll 1
      for (int h = 0; h < BMAC / 32; ++h) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld_wait();
          const int ch = 2 * h + half;
          if (ch + 1 < BMAC / 16) tmem_ld_x16(acc + (ch + 1) * 16, half ? va : vb);
          else {   // every column of this row is in registers: hand the accumulators back
            tc_fence_before();
            mbar_arrive(acc_empty);
          }
          const uint32_t(&v)[16] = half ? vb : va;
#pragma unroll
          for (int c = 0; c < 16; c += 4) {
            const float4 n4 = __ldg(reinterpret_cast<const float4*>(nj + ch * 16 + c));
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const float2 S = make_float2(__uint_as_float(v[c + e]), __uint_as_float(v[c + e + 1]));
              const float2 nn = e == 0 ? make_float2(n4.x, n4.y) : make_float2(n4.z, n4.w);
              float2 k, kd;
              math.eval2(S, add2(ni2, nn), k, kd);
              k = mul2(k, ks2);
              if (special) {
                const int64_t col = c0 + ch * 16 + c + e;
                const bool ok0 = (col < lim) && (col != gi), ok1 = (col + 1 < lim) && (col + 1 != gi);
                k = make_float2(ok0 ? k.x : 0.f, ok1 ? k.y : 0.f);
                if (col == pair_col) pairv = (double)k.x;
                if (col + 1 == pair_col) pairv = (double)k.y;
              }
              tsum = add2(tsum, k);
              tsq = fma2(k, k, tsq);
            }
          }
        }
      }
      const double ts = (double)((tsum.x + tsum.y) * wgt), tq = (double)((tsq.x + tsq.y) * wgt);
      if (same) {
        s_same += ts;
        q_same += tq;
      } else {
        s_cross += ts;
        q_cross += tq;
      }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// ================================================================================================
// two-pass path for wide features (d > 256): W = A o k'(D) materialised per row panel, then O = W Z as a GEMM
// ================================================================================================
// The fused kernel keeps O (128 x d fp32) in tensor memory, which caps it at d = 256.  A thread-block-cluster
// variant (feature-sliced O, partial Gram tiles reduce-scattered and W all-gathered over distributed shared
// memory) was built and measured first: correct, but shared memory left only a 2-3 stage TMA ring next to the
// exchange buffers and it reached 23% (d = 512) / 6% (d = 1024) of peak (profiles/r01_cluster_*.log).
// The two-pass path below reaches 45% / 57% and has no upper limit on d:
//   pass 1 (tc_wgen_kernel):  128 x 256 tiles of S = Z_i Z_j^T with K streamed through a 4-stage TMA ring (both
//          operands, 48 KB per 64-deep step: 25% less L2 traffic per flop than 128 x 128), two 256-column TMEM
//          accumulators alternate between tiles, all 16 epilogue warps drain each tile; epilogue = kernel transform
//          -> tile sums, row sums of W, W tile (bf16) stored to a row-panel buffer W[panel rows][Mp] (K-major for
//          pass 2).  The panel is sized by a byte budget (default 6 GB), so the N x N matrix never exists as a
//          whole; panels run back to back on the stream.  Inside a panel the column tiles are walked in windows of
//          ~24 MB of Z_j that all CTAs share at any time (L2 residency for any panel size / row shard).
//   pass 2 (tc_wz_kernel):    O[256 rows x 256 features] = W[256 x Mp] Z[Mp x 256]: 2 x 128-row A panels and one
//          64-row MN-major Z tile per 64-deep K step (64 KB / 1024 tensor cycles, the macro-tile ratio), all 512
//          TMEM columns as accumulators, K split S ways per unit so that units * S fills the SMs; partial tiles go
//          to slabs and are reduced in fixed order by the finalize kernel (deterministic).
//   finalize (wz_finalize_rows_kernel): g_i = r_i z_i - O_i (+ closed-form dot term), block sums.
struct WgenArgs {
  KernelFn kf;
  int64_t m, n, mp, np;
  float c_xx, c_yy, c_xy;
  const float* norms;
  int nrb_x, rb_x0, nrb_y, rb_y0;   // owned row blocks (as FusedArgs)
  int rbi0;                         // first owned row block (flat index) of this panel
  int TX, CT;                       // 256-column tiles of the X columns / per row block (X tiles, then Y tiles)
  int Wc, nwin, nrb_p;              // column window (tiles), windows, row blocks of the panel.  Every CTA walks the
                                    // windows in order and takes `chunk` of the nrb_p * Wc positions (row block,
                                    // tile) of each: at any time all CTAs work inside ONE window of Z_j tiles (L2)
  int spw;                          // result slots per (CTA, window)
  int nkp;                          // 64-feature panels
  int64_t chunk;                    // positions per work unit inside one window
  int slots;
  __nv_bfloat16* W;                 // [panel row blocks * 128][ldw]
  int64_t ldw;
  float* rpart;                     // [grid][slots][4][128]   (part = column quarter of the tile)
  double* spart;                    // [grid][slots][4][128][2]
};

constexpr int BNW = 256;                                  // tile width of pass 1
template <bool PAIR>
struct WgCfg {
  static constexpr int kStages = PAIR ? 6 : 4;
  // one 128-row Z_i panel + the Z_j tile (256 rows; PAIR: this CTA's 128-row half) per 64-deep step
  static constexpr int kStageBytes = BM * 128 + (PAIR ? BM : BNW) * 128;
  static constexpr int kSmem = 1024 + kStages * kStageBytes + 1024;
};

constexpr int kWgEpiWarps = 16;                           // 4 TMEM lane quarters x 4 column quarters
constexpr int kWgThreads = (kWgEpiWarps + 2) * 32;

// PAIR: the kernel runs as clusters of two CTAs that take the two row blocks of a row-block pair through the same
// column tiles with ONE UMMA stream: `tcgen05.mma.cta_group::2` (M = 256: 128 rows per CTA, each CTA supplies half
// of the Z_j tile from its own shared memory).  Per 64-deep step a CTA then receives 16 KB (its Z_i panel) + 16 KB
// (half a Z_j tile) instead of 48 KB -- the SM's operand ingest (~44 B/clk, measured identical for pass 1 and
// pass 2) is what bounds this kernel, and a TMA-multicast variant that still delivered the whole tile to both SMs
// gained only 3%.  Stages shrink to 32 KB, so the ring is 6 deep.  Protocol: both producers load into their own
// shared memory and signal the LEADER's `full` barrier (cta_group::2 TMA); the leader's issuer runs the UMMAs and
// multicasts its commits to both CTAs' `empty` / `acc_full` barriers; both CTAs' epilogue warps release the
// accumulator on the leader's `acc_empty` (remote arrive).
template <class Math, bool PAIR>
__global__ void __launch_bounds__(kWgThreads, 1)
tc_wgen_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ WgenArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kWgStages = WgCfg<PAIR>::kStages, kWgStageBytes = WgCfg<PAIR>::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgStages;
  uint64_t* acc_full = empty + kWgStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sParams = reinterpret_cast<float*>(tmem_slot + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_params(a.kf, sParams);
  if (tid == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], PAIR ? 2 * kWgEpiWarps : kWgEpiWarps);   // one elected arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == kWgEpiWarps + 1) {
    if (PAIR) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  if (warp == kWgEpiWarps && lane == 0) {
    prefetch_tmap(&tmap);
    prefetch_tmap(&tmap_b);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();   // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  auto rb_of = [&](int rbi) -> int { return rbi < a.nrb_x ? a.rb_x0 + rbi : a.rb_y0 + (rbi - a.nrb_x); };
  // tile ct of a row block: first column / end of its column set (X tiles never run into the Y columns)
  auto col0_of = [&](int ct) -> int { return ct < a.TX ? ct * BNW : (int)a.mp + (ct - a.TX) * BNW; };

  // this work unit's positions inside every window: [wp0, wp1) of the (row block [pair], tile) pairs.
  // PAIR: a position's row-block index counts row-block PAIRS; this CTA takes block 2 * index + rank of it.
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int unit_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nrbu = PAIR ? (a.nrb_p + 1) / 2 : a.nrb_p;       // row-block units of the panel
  const int wtot = nrbu * a.Wc;
  const int wp0 = (int)std::min<int64_t>((int64_t)unit_id * a.chunk, wtot);
  const int wp1 = (int)std::min<int64_t>((int64_t)wp0 + a.chunk, wtot);
  const int rbl_first = wp0 / a.Wc, t_first = wp0 - rbl_first * a.Wc;
  // row block of unit index u for this CTA (an odd panel's last pair has a dummy second block: it follows the
  // pipeline with the last real block's data and stores nothing)
  auto rbl_of = [&](int u) -> int { return PAIR ? 2 * u + rank : u; };
  auto rbl_ld = [&](int u) -> int { const int b = rbl_of(u); return b < a.nrb_p ? b : a.nrb_p - 1; };

  if (warp == kWgEpiWarps) {
    // ===================== TMA producer =====================
    uint32_t st = 0, ph = 0;
    for (int w = 0; w < a.nwin; ++w) {
    int rbl = rbl_first, t = t_first;
    for (int pos = wp0; pos < wp1; ++pos) {
      const int ct = w * a.Wc + t;
      const int32_t arow = rb_of(a.rbi0 + rbl_ld(rbl)) * BM, brow = ct < a.CT ? col0_of(ct) : 0;
      for (int p = 0; p < (ct < a.CT ? a.nkp : 0); ++p) {
        mbar_wait(&empty[st], ph ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + st * kWgStageBytes;
          if (PAIR) {   // own Z_i panel + own half of the Z_j tile; the bytes of both CTAs count on the leader's barrier
            const uint32_t lbar = map_to_cta(smem_u32(&full[st]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full[st], 2 * kWgStageBytes);
            tma_load_2d_pair(sa, &tmap, lbar, p * 64, arow);
            tma_load_2d_pair(sa + BM * 128, &tmap, lbar, p * 64, brow + rank * BM);
          } else {
            mbar_arrive_expect_tx(&full[st], kWgStageBytes);
            tma_load_2d(sa, &tmap, &full[st], p * 64, arow);
            tma_load_2d(sa + BM * 128, &tmap_b, &full[st], p * 64, brow);
          }
        }
        __syncwarp();
        if (++st == kWgStages) {
          st = 0;
          ph ^= 1;
        }
      }
      if (++t == a.Wc) {
        t = 0;
        ++rbl;
      }
    }
    }
  } else if (warp == kWgEpiWarps + 1) {
    // ===================== UMMA issuer =====================
    constexpr uint32_t idesc = make_idesc(PAIR ? 2 * BM : BM, BNW, kFmtBF16, false, false);
    const uint32_t hi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem), 16);
    uint32_t st = 0, ph = 0, ab = 0, aph = 0;
#ifdef SMMD_PIPE_TIMING
    long long wg_acc = 0, wg_full = 0, wg_tiles = 0;
    const long long wg_start = clock64();
#endif
    for (int w = 0; w < (PAIR && rank != 0 ? 0 : a.nwin); ++w) {   // PAIR: only the leader issues
    int t = t_first;
    for (int pos = wp0; pos < wp1; ++pos) {
      const bool real = w * a.Wc + t < a.CT;
      if (++t == a.Wc) t = 0;
      if (!real) continue;   // padding position of the last window
#ifdef SMMD_PIPE_TIMING
      const long long wg_t0 = clock64();
#endif
      if (PAIR) mbar_wait_cluster(&acc_empty[ab], aph ^ 1);   // the peer's epilogue arrives remotely
      else mbar_wait(&acc_empty[ab], aph ^ 1);
#ifdef SMMD_PIPE_TIMING
      wg_acc += clock64() - wg_t0;
      ++wg_tiles;
#endif
      tc_fence_after();
      const uint32_t dad = tmem + ab * BNW;
      for (int kk = 0; kk < a.nkp; ++kk) {
#ifdef SMMD_PIPE_TIMING
        const long long wg_t1 = clock64();
#endif
        mbar_wait(&full[st], ph);
#ifdef SMMD_PIPE_TIMING
        wg_full += clock64() - wg_t1;
#endif
        tc_fence_after();
        const uint32_t alo = a_lo0 + st * (kWgStageBytes >> 4), blo = alo + ((BM * 128) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (PAIR) umma_ss2_pair(dad, alo + k * 2, blo + k * 2, hi, idesc, (kk | k) ? 1u : 0u);
            else umma_ss2(dad, alo + k * 2, blo + k * 2, hi, idesc, (kk | k) ? 1u : 0u);
          }
          if (PAIR) umma_commit_pair(&empty[st], 3);
          else umma_commit(&empty[st]);
        }
        __syncwarp();
        if (++st == kWgStages) {
          st = 0;
          ph ^= 1;
        }
      }
      if (elect_one()) {
        if (PAIR) umma_commit_pair(&acc_full[ab], 3);
        else umma_commit(&acc_full[ab]);
      }
      __syncwarp();
      aph ^= ab;
      ab ^= 1;
    }
    }
#ifdef SMMD_PIPE_TIMING
    if (blockIdx.x == 0 && lane == 0 && wg_tiles > 0)
      printf("[wgen issuer] tiles %lld  cycles/tile: total %lld  wait acc_empty %lld  wait full %lld\n", wg_tiles,
             (clock64() - wg_start) / wg_tiles, wg_acc / wg_tiles, wg_full / wg_tiles);
#endif
  } else {
    // ===================== epilogue: all 16 warps on every tile (TMEM lane quarter x column quarter) ==========
    // With only two accumulators the issuer can start tile t+2 as soon as tile t is drained, so the drain latency
    // of ONE tile is what matters: 16 warps on one tile halve it compared with two groups on alternate tiles
    // (measured: tensor pipe 52% -> see profiles/).
    const int part = warp >> 2;         // columns [part * 64, +64) of the tile
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const Math math(a.kf, sParams);
    const float kscale = math.k_scale(), kdscale = math.kd_scale();
    const int mp = (int)a.mp, mvalid = (int)a.m, yvalid = (int)(a.mp + a.n);
    uint32_t tc = 0;   // real tiles of this CTA so far (parity = accumulator buffer)
    for (int w = 0; w < a.nwin; ++w) {
    int rbl = rbl_first, ct0 = t_first;
    int slot = w * a.spw;
    for (int left = wp1 - wp0; left > 0; ct0 = 0, ++slot, ++rbl) {
      const int TU = std::min(a.Wc - ct0, left);   // positions of this (window, row block) unit
      const bool real_rb = rbl_of(rbl) < a.nrb_p;
      const int rb = rb_of(a.rbi0 + rbl_ld(rbl));
      const int gi = rb * BM + r;
      const bool rowX = gi < mp;
      const float ni = a.norms[gi];
      float2 rsum = make_float2(0.f, 0.f);
      double dsame = 0.0, dcross = 0.0;
      __nv_bfloat16* wrow = a.W + ((int64_t)rbl_ld(rbl) * BM + r) * a.ldw;
      for (int lt = 0; lt < TU; ++lt) {
        const int ct = w * a.Wc + ct0 + lt;
        if (ct >= a.CT) break;           // padding positions at the end of the last window
        const int grp = (int)(tc & 1);   // accumulator buffer of this tile
        const int c0 = col0_of(ct);
        const bool colX = ct < a.TX;
        const bool same = (colX == rowX);
        const float2 cw = bc2((same ? (rowX ? a.c_xx : a.c_yy) : a.c_xy) * kdscale);
        const int lim = colX ? mvalid : yvalid;
        const int cend = colX ? mp : (int)(a.mp + a.np);               // end of this column set
        const int nch = (cend - c0 < BNW ? cend - c0 : BNW) / 16;      // chunks that belong to this tile
        const bool special = (c0 + BNW > lim) || (rb * BM >= c0 && rb * BM < c0 + BNW);
        mbar_wait(&acc_full[grp], (tc >> 1) & 1);
        tc_fence_after();
        const float* nj = a.norms + c0;
        float2 tsum = make_float2(0.f, 0.f);
        const int h0 = part * (BNW / 64);
        const int h1 = nch < h0 + BNW / 64 ? nch : h0 + BNW / 64;
        auto release_acc = [&]() {   // whole warp has finished its tcgen05.ld of this tile: one elected arrive
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(map_to_cta(smem_u32(&acc_empty[grp]), 0));
            else mbar_arrive(&acc_empty[grp]);
          }
        };
        if (h1 <= h0) release_acc();   // nothing of this tile in my column quarter
        // software pipeline over the 16-column chunks: the tcgen05.ld of chunk h+1 is in flight during the math of
        // chunk h (two register buffers, as in the fused kernel); one 32-byte store per chunk and row
        auto do_chunk = [&](const uint32_t (&v)[16], int h) {
          uint32_t wpk[8];
          if (!special) fused_chunk16<Math, false>(math, v, nj + h * 16, ni, cw, 0, 0, 0, tsum, rsum, wpk);
          else fused_chunk16<Math, true>(math, v, nj + h * 16, ni, cw, c0 + h * 16, lim, gi, tsum, rsum, wpk);
          if (real_rb) st_global_v8(wrow + c0 + h * 16, wpk);
        };
        uint32_t va[16], vb[16];
        if (h0 < h1) tmem_ld_x16(tmem + grp * BNW + h0 * 16 + lane_base, va);
#pragma unroll 1
        for (int h = h0; h < h1; h += 2) {
          tmem_ld_wait();                                   // va = chunk h
          if (h + 1 < h1) tmem_ld_x16(tmem + grp * BNW + (h + 1) * 16 + lane_base, vb);
          else release_acc();
          do_chunk(va, h);
          if (h + 1 < h1) {
            tmem_ld_wait();                                 // vb = chunk h + 1
            if (h + 2 < h1) tmem_ld_x16(tmem + grp * BNW + (h + 2) * 16 + lane_base, va);
            else release_acc();
            do_chunk(vb, h + 1);
          }
        }
        if (same) dsame += (double)((tsum.x + tsum.y) * kscale);
        else dcross += (double)((tsum.x + tsum.y) * kscale);
        ++tc;
      }
      // slab of (work unit, slot[, rank]): PAIR keeps the two row blocks of a pair side by side
      const int64_t sl = PAIR ? ((int64_t)unit_id * a.slots + slot) * 2 + rank : (int64_t)unit_id * a.slots + slot;
      a.rpart[(sl * 4 + part) * BM + r] = rsum.x + rsum.y;
      double* sp = a.spart + ((sl * 4 + part) * BM + r) * 2;
      sp[0] = dsame;
      sp[1] = dcross;
      left -= TU;
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();   // nobody leaves while the peer may still multicast into its shared memory
  if (warp == kWgEpiWarps + 1) {
    if (PAIR) tmem_dealloc_pair<512>(tmem);
    else tmem_dealloc<512>(tmem);
  }
}

// ---- pass 2: O = W Z ----
constexpr int kWzStages = 3;
constexpr int kWzStageBytes = 2 * BM * 128 + 4 * BNF * 128;   // two 128-row W panels + four 64-feature Z panels = 64 KB
constexpr int kWzSmem = 1024 + kWzStages * kWzStageBytes + 1024;

struct WzArgs {
  int nmb;        // 256-row macro blocks of the panel
  int FB;         // 256-feature blocks
  int dp;         // padded feature count (multiple of 64)
  int KT;         // K steps of 64 (= Mp / 64)
  int S;          // K splits per unit
  int ksteps;     // K steps per piece
  float* Opart;   // [unit = mb * FB + fb][S][256][256]
};

__global__ void __launch_bounds__(kThreads, 1)
tc_wz_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_z,
             const __grid_constant__ WzArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWzStages * kWzStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWzStages;
  uint64_t* acc_full = empty + kWzStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kWzStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_z);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // piece = (split s, unit u), s-major so that one wave of CTAs walks the same rows of Z (L2 reuse)
  const int units = a.nmb * a.FB;
  const int s = (int)blockIdx.x / units;
  const int u = (int)blockIdx.x - s * units;
  const int mb = u / a.FB, fb = u - mb * a.FB;
  const int k0 = s * a.ksteps;
  const int k1 = k0 + a.ksteps < a.KT ? k0 + a.ksteps : a.KT;
  const int nf = a.dp - fb * 256 < 256 ? a.dp - fb * 256 : 256;   // features of this block (multiple of 64)
  const int npan = nf / 64;

  if (warp == 8) {
    uint32_t st = 0, ph = 0;
    for (int ks = k0; ks < k1; ++ks) {
      mbar_wait(&empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[st], 2 * BM * 128 + npan * BNF * 128);
        uint8_t* sa = smem + st * kWzStageBytes;
        tma_load_2d(sa, &tmap_w, &full[st], ks * 64, mb * 256);
        tma_load_2d(sa + BM * 128, &tmap_w, &full[st], ks * 64, mb * 256 + BM);
        uint8_t* sb = sa + 2 * BM * 128;
        for (int p = 0; p < npan; ++p) tma_load_2d(sb + p * (BNF * 128), &tmap_z, &full[st], fb * 256 + p * 64, ks * 64);
      }
      __syncwarp();
      if (++st == kWzStages) {
        st = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 9) {
    const uint32_t idesc = make_idesc(BM, (uint32_t)nf, kFmtBF16, false, true);   // B = Z tile, MN-major
    const uint32_t hi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem), 16);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem + 2 * BM * 128), BNF * 128);
    uint32_t st = 0, ph = 0;
    for (int ks = k0; ks < k1; ++ks) {
      mbar_wait(&full[st], ph);
      tc_fence_after();
      const uint32_t alo = a_lo0 + st * (kWzStageBytes >> 4), blo = b_lo0 + st * (kWzStageBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss2(tmem + h * 256, alo + h * ((BM * 128) >> 4) + k * 2, blo + k * (2048 >> 4), hi, idesc,
                     (ks > k0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[st]);
        if (ks == k1 - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++st == kWzStages) {
        st = 0;
        ph ^= 1;
      }
    }
  } else {
    // drain: warp -> (row half, TMEM lane quarter); each thread stores its row of the partial tile
    const int half = warp >> 2, q = warp & 3;
    const int row = half * BM + q * 32 + lane;
    float* orow = a.Opart + (((int64_t)u * a.S + s) * 256 + row) * 256;
    if (k1 > k0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const uint32_t base = tmem + half * 256 + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c < nf; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(base + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          *reinterpret_cast<float4*>(orow + c + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                 __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
      }
    } else {
      for (int c = 0; c < nf; c += 4) *reinterpret_cast<float4*>(orow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// ---- finalize of the two-pass path (one panel): warp per row, lanes over features ----
struct WzFinArgs {
  KernelFn kf;
  int64_t m, n, mp, np, d;
  int64_t x0, ox, y0, oy;
  int dp;
  int nrb_x, rb_x0, nrb_y, rb_y0;
  int rbi0, nrb_p;           // panel: first owned row block (flat) / count
  int Wc, nwin;              // pass-1 column window (tiles) / number of windows
  int spw;                   // pass-1 slots per (work unit, window)
  int pair;                  // pass 1 ran as CTA pairs (row-block pairs, two slabs per slot)
  int64_t chunk;             // pass-1 positions per work unit inside one window
  int slots;
  int FB, S;                 // pass-2 feature blocks / K splits
  double a_xx, a_yy, a_xy;
  SrcLayout src;
  const float* norms;
  const double* csum;
  const float* Opart;
  const float* rpart;
  const double* spart;
  float* dX;
  float* dY;
  double* partials;          // [gridDim.x][6] of this panel
};

__global__ void __launch_bounds__(256) wz_finalize_rows_kernel(WzFinArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double sh[8][6];
  double q[6] = {0, 0, 0, 0, 0, 0};
  const bool dot = a.kf.family == FAM_RQ && a.kf.add_dot > 0.f && a.csum != nullptr;
  for (int rr = 0; rr < kFinRowsPerWarp; ++rr) {
    const int64_t lr = ((int64_t)blockIdx.x * 8 + warp) * kFinRowsPerWarp + rr;   // row of the panel
    if (lr >= (int64_t)a.nrb_p * BM) break;
    const int rbl = (int)(lr / BM), r = (int)(lr % BM);
    const int rbi = a.rbi0 + rbl;
    const int rb = rbi < a.nrb_x ? a.rb_x0 + rbi : a.rb_y0 + (rbi - a.nrb_x);
    const int64_t gi = (int64_t)rb * BM + r;
    const bool rowX = gi < a.mp;
    const int64_t li = rowX ? gi : gi - a.mp;
    if (rowX ? (li < a.x0 || li >= a.x0 + a.ox) : (li < a.y0 || li >= a.y0 + a.oy)) continue;   // not owned / padding
    // pass-1 slabs of this row block: one (window, row block) unit per window, each cut over <= 2 CTAs
    float rs = 0.f;
    double ssame = 0.0, scross = 0.0;
    for (int w = 0; w < a.nwin; ++w) {
      const int ru = a.pair ? rbl >> 1 : rbl;                      // row-block unit of this row block
      const int64_t f0 = (int64_t)ru * a.Wc, f1 = f0 + a.Wc - 1;   // positions of this unit inside the window
      const int64_t g0 = f0 / a.chunk, g1 = f1 / a.chunk;
      for (int64_t g = g0; g <= g1; ++g) {
        int64_t sl = g * a.slots + (int64_t)w * a.spw + (ru - (g * a.chunk) / a.Wc);
        if (a.pair) sl = sl * 2 + (rbl & 1);
        for (int pt = 0; pt < 4; ++pt) {
          rs += a.rpart[(sl * 4 + pt) * BM + r];
          const double* sp = a.spart + ((sl * 4 + pt) * BM + r) * 2;
          ssame += sp[0];
          scross += sp[1];
        }
      }
    }
    const double a_same = rowX ? a.a_xx : a.a_yy;
    double dsame = 0.0, dcross = 0.0;
    float* out = nullptr;
    if (a.dX) out = rowX ? a.dX + (li - a.x0) * a.d : a.dY + (li - a.y0) * a.d;
    const bool owned = (rowX ? a.src.Xo : a.src.Yo) != nullptr;
    const void* src = owned ? static_cast<const void*>(rowX ? a.src.Xo : a.src.Yo) : (rowX ? a.src.X : a.src.Y);
    const int64_t ld = owned ? a.src.ldo : (rowX ? a.src.ldx : a.src.ldy);
    const int sdtype = owned ? (int)SMMD_F32 : a.src.dtype;
    const int64_t srow = owned ? (rowX ? li - a.x0 : li - a.y0) : src_row(li, rowX, a.src.blk_x, a.src.blk_y);
    const int mbl = rbl >> 1;                 // macro block of the panel, row inside it
    const int row256 = (rbl & 1) * BM + r;
    const float* obase = a.Opart + (((int64_t)mbl * a.FB) * a.S * 256 + row256) * 256;   // + (fb * S + s) * 65536 + cc
    const bool vec = out != nullptr && !dot && sdtype == SMMD_F32 && (a.d % 4 == 0) && (ld % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec) {
      const float* zsrc = reinterpret_cast<const float*>(src) + srow * ld;
      for (int c = 4 * lane; c < a.d; c += 128) {   // 4 consecutive features per lane (never straddle a 256 block)
        const float* op = obase + (int64_t)(c >> 8) * a.S * 65536 + (c & 255);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < a.S; ++s) {   // fixed order
          const float4 t = *reinterpret_cast<const float4*>(op + (int64_t)s * 65536);
          o.x += t.x;
          o.y += t.y;
          o.z += t.z;
          o.w += t.w;
        }
        const float4 z4 = *reinterpret_cast<const float4*>(zsrc + c);
        float zz[4] = {z4.x, z4.y, z4.z, z4.w};
        const float oo[4] = {o.x, o.y, o.z, o.w};
        float gv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (a.kf.tanh_features) zz[e] = tanhf(zz[e]);
          gv[e] = rs * zz[e] - oo[e];
          if (a.kf.tanh_features) gv[e] *= (1.f - zz[e] * zz[e]);
        }
        *reinterpret_cast<float4*>(out + c) = make_float4(gv[0], gv[1], gv[2], gv[3]);
      }
    } else {
      for (int c = lane; c < a.d; c += 32) {
        float o = 0.f;
        if (out) {
          const float* op = obase + (int64_t)(c >> 8) * a.S * 65536 + (c & 255);
          for (int s = 0; s < a.S; ++s) o += op[(int64_t)s * 65536];   // fixed order
        }
        const int64_t sidx = srow * ld + c;
        float z = sdtype == SMMD_F32 ? reinterpret_cast<const float*>(src)[sidx]
                                      : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[sidx]);
        if (a.kf.tanh_features) z = tanhf(z);
        if (out) {
          float gv = rs * z - o;
          if (dot) {
            const double cs = a.csum[(rowX ? 0 : 1) * a.dp + c], co = a.csum[(rowX ? 1 : 0) * a.dp + c];
            gv += (float)(2.0 * (double)a.kf.add_dot * (a_same * cs + a.a_xy * co));
          }
          if (a.kf.tanh_features) gv *= (1.f - z * z);
          out[c] = gv;
        }
        if (dot) {
          dsame += (double)z * a.csum[(rowX ? 0 : 1) * a.dp + c];
          dcross += (double)z * a.csum[(rowX ? 1 : 0) * a.dp + c];
        }
      }
    }
    if (dot) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dsame += __shfl_xor_sync(0xffffffffu, dsame, o);
        dcross += __shfl_xor_sync(0xffffffffu, dcross, o);
      }
    }
    const float ni = a.norms[gi];
    const double v_same = ssame + (dot ? (double)a.kf.add_dot * (dsame - (double)ni) : 0.0);
    const double v_cross = scross + (dot ? (double)a.kf.add_dot * dcross : 0.0);
    const double v_diag = a.kf.family == FAM_RQ ? (double)a.kf.const_diag + (double)a.kf.add_dot * (double)ni
                                                : (double)diag_value(a.kf, ni);
    if (rowX) {
      q[0] += v_same;
      q[2] += v_cross;
      q[4] += v_diag;
    } else {
      q[1] += v_same;
      q[3] += v_cross;
      q[5] += v_diag;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) sh[warp][i] = q[i];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    a.partials[(int64_t)blockIdx.x * 6 + threadIdx.x] = t;
  }
}

struct FusedPlan {
  int64_t mp, np, Mp, dp;
  int nrb_x, rb_x0, nrb_y, rb_y0, T, grid, slots;
  int64_t total, chunk;
  size_t off_Z, off_norm, off_csum, off_O, off_r, off_s, off_stats, off_end;
};

size_t up256(size_t v) { return (v + 255) / 256 * 256; }

FusedPlan fused_plan(int64_t m, int64_t n, int64_t d, int64_t x0, int64_t x1, int64_t y0, int64_t y1) {
  FusedPlan p;
  p.mp = round_up(m, BM);
  p.np = round_up(n, BM);
  p.Mp = p.mp + p.np;
  p.dp = round_up(d, 64);
  p.rb_x0 = (int)(x0 / BM);
  p.nrb_x = x1 > x0 ? (int)((x1 - 1) / BM) - p.rb_x0 + 1 : 0;
  p.rb_y0 = (int)((p.mp + y0) / BM);
  p.nrb_y = y1 > y0 ? (int)((p.mp + y1 - 1) / BM) - p.rb_y0 + 1 : 0;
  p.T = (int)(p.Mp / BNF);
  p.total = (int64_t)(p.nrb_x + p.nrb_y) * p.T;
  p.grid = (int)std::min<int64_t>(sm_count(), p.total);
  if (p.grid < 1) p.grid = 1;
  p.chunk = (p.total + p.grid - 1) / p.grid;
  // Large Z: give every CTA WHOLE row blocks when that costs < 2% balance.  All CTAs then start their sweeps at
  // column tile 0 together and stay in step, so the Z_j tiles in flight are a narrow window instead of all of Z
  // (ncu at N = 65536, d = 256: 2.0 GB of DRAM reads per launch with free-running chunks, 29x the 67 MB of Z;
  // no change in run time -- L2 misses were only 1.4% of it -- but the re-reads are gone).
  if (tuning().fused_lockstep && (int64_t)p.Mp * p.dp * 2 > ((int64_t)32 << 20)) {
    const int64_t nrb = p.nrb_x + p.nrb_y;
    const int64_t aligned = (nrb + p.grid - 1) / p.grid * p.T;
    if (aligned * p.grid * 100 <= p.total * 102) p.chunk = aligned;
  }
  p.grid = (int)((p.total + p.chunk - 1) / p.chunk);
  p.slots = (int)((p.chunk + p.T - 1) / p.T) + 1;
  size_t o = 0;
  p.off_Z = o;
  o = up256(o + (size_t)p.Mp * p.dp * 2);
  p.off_norm = o;
  o = up256(o + (size_t)p.Mp * 4);
  p.off_csum = o;
  o = up256(o + (size_t)2 * p.dp * 8);
  p.off_O = o;
  o = up256(o + (size_t)p.grid * p.slots * BM * p.dp * 4);
  p.off_r = o;
  o = up256(o + (size_t)p.grid * p.slots * 4 * BM * 4);
  p.off_s = o;
  o = up256(o + (size_t)p.grid * p.slots * 4 * BM * 2 * 8);
  p.off_stats = o;
  o = up256(o + (size_t)((x1 - x0) + (y1 - y0)) * RS_COUNT * 8);
  p.off_end = o;
  return p;
}

template <class Math, int KSPLIT>
cudaError_t launch_fused_k(const CUtensorMap& tzi, const CUtensorMap& tzj, const FusedArgs& a, int grid, cudaStream_t s) {
  const int smem = fused_smem(a.npanel, a.nst);
  if (a.nst < 4 || smem > kMaxSmem) return cudaErrorInvalidConfiguration;
  auto kern = tc_fused_kernel<Math, KSPLIT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, 64 + 256 * KSPLIT, smem, s>>>(tzi, tzj, a);
  return cudaGetLastError();
}
template <class Math>
cudaError_t launch_fused_t(const CUtensorMap& tzi, const CUtensorMap& tzj, const FusedArgs& a, int grid, cudaStream_t s) {
  return a.ksplit == 2 ? launch_fused_k<Math, 2>(tzi, tzj, a, grid, s) : launch_fused_k<Math, 1>(tzi, tzj, a, grid, s);
}

cudaError_t launch_fused(TcVariant v, const CUtensorMap& tzi, const CUtensorMap& tzj, const FusedArgs& a, int grid,
                         cudaStream_t s) {
  switch (v) {
    case TV_RBF1: return launch_fused_t<MathRbf1>(tzi, tzj, a, grid, s);
    case TV_RBF_LADDER5: return launch_fused_t<MathRbfLadder<5>>(tzi, tzj, a, grid, s);
    case TV_RBF_GENERIC: return launch_fused_t<MathGeneric<FAM_RBF>>(tzi, tzj, a, grid, s);
    case TV_RQ3_DEFAULT: return launch_fused_t<MathRq3Default>(tzi, tzj, a, grid, s);
    case TV_RQ_GENERIC: return launch_fused_t<MathGeneric<FAM_RQ>>(tzi, tzj, a, grid, s);
    case TV_DISTANCE: return launch_fused_t<MathDistance>(tzi, tzj, a, grid, s);
    case TV_NULL: return launch_fused_t<MathNull>(tzi, tzj, a, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

// ---- two-pass path: plan ----
struct WzPlan {
  int64_t mp, np, Mp, dp;
  int nrb_x, rb_x0, nrb_y, rb_y0, nrb;
  int Wc, nwin;         // pass-1 column window (tiles) / windows per row block
  int TX, CT, KT, FB;   // pass-1 tiles (256 columns) over the X columns / per row block
  int P, npanels;   // row blocks per panel (even) / panels
  size_t off_Z, off_norm, off_csum, off_W, off_r, off_s, off_O, off_stats, off_end;
  int fin_blocks_total;
};
struct WzPanel {
  int rbi0, nrb_p;
  int grid1, slots1, spw1, pair;   // grid1 counts work units (CTAs, or CTA pairs when pair = 1)
  int64_t tiles1, chunk1;
  int nmb, units, S, ksteps, grid2;
  int fin_blocks, fin_block0;
};

// K splits per unit so that units * S fills whole waves of SMs (>= 8 K steps per piece, <= 16 splits)
int wz_choose_split(int units, int KT) {
  const int sm = sm_count();
  int best = 1;
  double best_eff = 0.0;
  for (int S = 1; S <= 16 && KT / S >= 8; ++S) {
    const int64_t pieces = (int64_t)units * S;
    const int64_t waves = (pieces + sm - 1) / sm;
    const double eff = (double)pieces / (double)(waves * sm);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = S;
    }
    if (eff >= 0.92) return S;
  }
  return best;
}

WzPanel wz_panel(const WzPlan& p, int idx) {
  WzPanel q;
  q.rbi0 = idx * p.P;
  q.nrb_p = std::min(p.P, p.nrb - q.rbi0);
  q.pair = tuning().wz_pair && q.nrb_p >= sm_count() ? 1 : 0;   // pairs need >= 1 row-block pair per CTA pair
  const int nrbu = q.pair ? (q.nrb_p + 1) / 2 : q.nrb_p;
  q.tiles1 = (int64_t)nrbu * p.Wc;   // positions of ONE window (every work unit takes chunk1 of them, per window)
  q.grid1 = (int)std::min<int64_t>(q.pair ? sm_count() / 2 : sm_count(), q.tiles1);
  if (q.grid1 < 1) q.grid1 = 1;
  q.chunk1 = (q.tiles1 + q.grid1 - 1) / q.grid1;
  q.grid1 = (int)((q.tiles1 + q.chunk1 - 1) / q.chunk1);
  q.spw1 = (int)((q.chunk1 + p.Wc - 1) / p.Wc) + 1;
  q.slots1 = p.nwin * q.spw1;
  q.nmb = (q.nrb_p + 1) / 2;
  q.units = q.nmb * p.FB;
  q.S = wz_choose_split(q.units, p.KT);
  q.ksteps = (p.KT + q.S - 1) / q.S;
  q.grid2 = q.units * q.S;
  q.fin_blocks = (q.nrb_p * BM + kFinRowsPerCta - 1) / kFinRowsPerCta;
  q.fin_block0 = idx * ((p.P * BM + kFinRowsPerCta - 1) / kFinRowsPerCta);
  return q;
}

WzPlan wz_plan(int64_t m, int64_t n, int64_t d, int64_t x0, int64_t x1, int64_t y0, int64_t y1) {
  WzPlan p;
  p.mp = round_up(m, BM);
  p.np = round_up(n, BM);
  p.Mp = p.mp + p.np;
  p.dp = round_up(d, 64);
  p.rb_x0 = (int)(x0 / BM);
  p.nrb_x = x1 > x0 ? (int)((x1 - 1) / BM) - p.rb_x0 + 1 : 0;
  p.rb_y0 = (int)((p.mp + y0) / BM);
  p.nrb_y = y1 > y0 ? (int)((p.mp + y1 - 1) / BM) - p.rb_y0 + 1 : 0;
  p.nrb = p.nrb_x + p.nrb_y;
  p.TX = (int)((p.mp + BNW - 1) / BNW);
  p.CT = p.TX + (int)((p.np + BNW - 1) / BNW);
  p.KT = (int)(p.Mp / 64);
  p.FB = (int)((p.dp + 255) / 256);
  // column window: ~24 MB of Z_j tiles, so that window + the panel's row blocks stay L2 resident whatever the
  // chunk alignment (whole-Z sweeps per row block thrashed L2 once Z > 126 MB: 86.9 ms instead of 66 at d = 1024)
  p.Wc = (int)std::max<int64_t>(4, std::min<int64_t>(p.CT, ((int64_t)24 << 20) / (BNW * p.dp * 2)));
  if (p.Mp * p.dp * 2 <= ((int64_t)96 << 20)) p.Wc = p.CT;   // Z fits L2: one window (measured 5% faster)
  p.nwin = (p.CT + p.Wc - 1) / p.Wc;
  int64_t P = tuning().wz_panel_bytes / (BM * p.Mp * 2);
  P = std::max<int64_t>(2, P & ~int64_t(1));
  P = std::min<int64_t>(P, (p.nrb + 1) & ~1);
  p.P = (int)std::max<int64_t>(2, P);
  p.npanels = std::max(1, (p.nrb + p.P - 1) / p.P);
  size_t need_r = 0, need_O = 0;
  int fin_total = 0;
  for (int i = 0; i < p.npanels; i += std::max(1, p.npanels - 1)) {   // first (full) and last panel bound all others
    const WzPanel q = wz_panel(p, i);
    need_r = std::max(need_r, (size_t)q.grid1 * q.slots1 * (q.pair ? 2 : 1) * 4 * BM);
    need_O = std::max(need_O, (size_t)q.units * q.S * 256 * 256 * 4);
    if (p.npanels == 1) break;
  }
  {
    const WzPanel last = wz_panel(p, p.npanels - 1);
    fin_total = last.fin_block0 + last.fin_blocks;
  }
  p.fin_blocks_total = fin_total;
  size_t o = 0;
  p.off_Z = o;
  o = up256(o + (size_t)p.Mp * p.dp * 2);
  p.off_norm = o;
  o = up256(o + (size_t)p.Mp * 4);
  p.off_csum = o;
  o = up256(o + (size_t)2 * p.dp * 8);
  p.off_W = o;
  o = up256(o + (size_t)p.P * BM * p.Mp * 2);
  p.off_r = o;
  o = up256(o + need_r * 4);
  p.off_s = o;
  o = up256(o + need_r * 2 * 8);
  p.off_O = o;
  o = up256(o + need_O);
  p.off_stats = o;
  o = up256(o + (size_t)std::max(1, fin_total) * 6 * 8);
  p.off_end = o;
  return p;
}

template <class Math>
cudaError_t launch_wgen_t(const CUtensorMap& tm, const CUtensorMap& tb, const WgenArgs& a, int grid, int pair,
                          cudaStream_t s) {
  if (!pair) {
    auto kern = tc_wgen_kernel<Math, false>;
    constexpr int kWgSmem = WgCfg<false>::kSmem;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kWgThreads, kWgSmem, s>>>(tm, tb, a);
    return cudaGetLastError();
  }
  auto kern = tc_wgen_kernel<Math, true>;
  constexpr int kWgSmem = WgCfg<true>::kSmem;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * grid));
  cfg.blockDim = dim3(kWgThreads);
  cfg.dynamicSmemBytes = (size_t)kWgSmem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tm, tb, a);
}
cudaError_t launch_wgen(TcVariant v, const CUtensorMap& tm, const CUtensorMap& tb, const WgenArgs& a, int grid,
                        int pair, cudaStream_t s) {
  switch (v) {
    case TV_RBF1: return launch_wgen_t<MathRbf1>(tm, tb, a, grid, pair, s);
    case TV_RBF_LADDER5: return launch_wgen_t<MathRbfLadder<5>>(tm, tb, a, grid, pair, s);
    case TV_RBF_GENERIC: return launch_wgen_t<MathGeneric<FAM_RBF>>(tm, tb, a, grid, pair, s);
    case TV_RQ3_DEFAULT: return launch_wgen_t<MathRq3Default>(tm, tb, a, grid, pair, s);
    case TV_RQ_GENERIC: return launch_wgen_t<MathGeneric<FAM_RQ>>(tm, tb, a, grid, pair, s);
    case TV_DISTANCE: return launch_wgen_t<MathDistance>(tm, tb, a, grid, pair, s);
    case TV_NULL: return launch_wgen_t<MathNull>(tm, tb, a, grid, pair, s);
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_wz(const CUtensorMap& tw, const CUtensorMap& tz, const WzArgs& a, int grid, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(tc_wz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWzSmem);
  if (e != cudaSuccess) return e;
  tc_wz_kernel<<<grid, kThreads, kWzSmem, s>>>(tw, tz, a);
  return cudaGetLastError();
}

struct StreamPlan {
  int64_t mp, np, Mp, dp, dpz;
  int RB, CT, grid;
  int64_t total, chunk;
  size_t off_Z, off_norm, off_stats, off_end;
};

StreamPlan stream_plan(int64_t m, int64_t n, int64_t d, int64_t batch, int split) {
  StreamPlan p;
  p.mp = round_up(m, BM);
  p.np = round_up(n, BM);
  p.Mp = p.mp + p.np;
  p.dp = round_up(d, 64);
  p.dpz = split ? 2 * p.dp : p.dp;
  p.RB = (int)(p.Mp / BM);
  p.CT = (int)(p.Mp / BNS);
  p.total = batch * p.RB * p.CT;
  p.grid = (int)std::min<int64_t>(sm_count(), p.total);
  p.chunk = (p.total + p.grid - 1) / p.grid;
  p.grid = (int)((p.total + p.chunk - 1) / p.chunk);
  size_t o = 0;
  p.off_Z = o;
  o = up256(o + (size_t)batch * p.Mp * p.dpz * 2);
  p.off_norm = o;
  o = up256(o + (size_t)batch * p.Mp * 4);
  p.off_stats = o;
  o = up256(o + (size_t)batch * (m + n) * RS_COUNT * 8);
  p.off_end = o;
  return p;
}

struct MacroPlan {
  int64_t mp, np, Mp, dp, dpz;
  int R, Rx, tiles_per_batch, grid;
  int64_t total, chunk;
  size_t off_Z, off_norm, off_stats, off_end;
};

MacroPlan macro_plan(int64_t m, int64_t n, int64_t d, int64_t batch, int split, int sym) {
  MacroPlan p;
  p.mp = round_up(m, BMAC);
  p.np = round_up(n, BMAC);
  p.Mp = p.mp + p.np;
  p.dp = round_up(d, 64);
  p.dpz = split ? 2 * p.dp : p.dp;
  p.R = (int)(p.Mp / BMAC);
  p.Rx = (int)(p.mp / BMAC);
  p.tiles_per_batch = sym ? p.R * (p.R + 1) / 2 : p.R * p.R;
  p.total = batch * p.tiles_per_batch;
  p.grid = (int)std::min<int64_t>(sm_count(), p.total);
  p.chunk = (p.total + p.grid - 1) / p.grid;
  p.grid = (int)((p.total + p.chunk - 1) / p.chunk);
  size_t o = 0;
  p.off_Z = o;
  o = up256(o + (size_t)batch * p.Mp * p.dpz * 2);
  p.off_norm = o;
  o = up256(o + (size_t)batch * p.Mp * 4);
  p.off_stats = o;
  o = up256(o + (size_t)batch * (m + n) * RS_COUNT * 8);
  p.off_end = o;
  return p;
}

template <class Math>
cudaError_t launch_macro_t(const CUtensorMap& tm, const MacroArgs& a, int grid, cudaStream_t s) {
  auto kern = tc_macro_kernel<Math>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMacSmem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, kMacSmem, s>>>(tm, a);
  return cudaGetLastError();
}

template <class Math>
cudaError_t launch_stream_t(const CUtensorMap& tm, const StreamArgs& a, int grid, cudaStream_t s) {
  auto kern = tc_stream_kernel<Math>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, kStreamSmem, s>>>(tm, a);
  return cudaGetLastError();
}

cudaError_t launch_stream(TcVariant v, const CUtensorMap& tm, const StreamArgs& a, int grid, cudaStream_t s) {
  switch (v) {
    case TV_RBF1: return launch_stream_t<MathRbf1>(tm, a, grid, s);
    case TV_RBF_LADDER5: return launch_stream_t<MathRbfLadder<5>>(tm, a, grid, s);
    case TV_RBF_GENERIC: return launch_stream_t<MathGeneric<FAM_RBF>>(tm, a, grid, s);
    case TV_RQ3_DEFAULT: return launch_stream_t<MathRq3Default>(tm, a, grid, s);
    case TV_RQ_GENERIC: return launch_stream_t<MathGeneric<FAM_RQ>>(tm, a, grid, s);
    case TV_DISTANCE: return launch_stream_t<MathDistance>(tm, a, grid, s);
    case TV_POLY3: return launch_stream_t<MathPoly3>(tm, a, grid, s);
    case TV_POLY_GENERIC: return launch_stream_t<MathPolyN>(tm, a, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

#ifdef SMMD_PIPE_TIMING
void pipe_timing_dump(bool reset) {
  unsigned long long h[32];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_pipe_dbg, sizeof(h));
  const double nt2 = (double)(h[3] ? h[3] : 1), nt = (double)(h[14] ? h[14] : 1);
  printf("[pipe timing CTA0] MMA thread per tile: wait_w %.0f  issue#2 %.0f  wait_zj %.0f  issue#1 %.0f   (tiles %llu)\n",
         h[0] / nt2, h[1] / nt2, h[4] / nt2, h[2] / nt2, h[3]);
  printf("[pipe timing CTA0] epilogue warp0 per OWN tile: wait_s %.0f  ld+math+st(+wait W drained) %.0f  st-drain+arrive %.0f"
         "   (tiles %llu)\n", h[9] / nt, h[11] / nt, h[13] / nt, h[14]);
  if (reset) {
    memset(h, 0, sizeof(h));
    cudaMemcpyToSymbol(g_pipe_dbg, h, sizeof(h));
  }
}
#endif

// ------------------------------------------------------------------------------------------------
// public (library-internal) interface
// ------------------------------------------------------------------------------------------------
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

size_t tc_mmd2_workspace_bytes(int64_t m, int64_t n, int64_t d, int want_grad, int precision) {
  // worst case over shards: a full-range plan bounds every rank's plan
  const size_t fused = !want_grad ? 0 : (use_two_pass(d) ? wz_plan(m, n, d, 0, m, 0, n).off_end
                                                          : fused_plan(m, n, d, 0, m, 0, n).off_end);
  const size_t stream = stream_plan(m, n, d, 1, precision == SMMD_PREC_BF16X3).off_end + 4096;
  return std::max(fused, stream);
}

cudaError_t tc_mmd2_run(const KernelFn& kf_in, const Geometry& g, const Coefs& c, const SrcLayout& src, int precision,
                        double* scalars, float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches,
                        const char** path) {
  const void* X = src.X;
  const void* Y = src.Y;
  const int dtype = src.dtype;
  const int64_t ldx = src.ldx, ldy = src.ldy;
  if (!tc_family_ok(kf_in)) return cudaErrorNotSupported;
  KernelFn kf = kf_in;
  TcVariant variant = select_tc_variant(kf);
  if (tuning().null_math) variant = TV_NULL;
  char* w = static_cast<char*>(ws);
  cudaError_t e;
  const bool want_grad = dX != nullptr;
  if (want_grad && use_two_pass(g.d)) {
    *path = "tc_bf16_wz";
    const WzPlan p = wz_plan(g.m, g.n, g.d, g.x0, g.x1, g.y0, g.y1);
    if (p.off_end > ws_bytes) return cudaErrorInvalidValue;
    __nv_bfloat16* Z = reinterpret_cast<__nv_bfloat16*>(w + p.off_Z);
    float* norms = reinterpret_cast<float*>(w + p.off_norm);
    double* csum = reinterpret_cast<double*>(w + p.off_csum);
    __nv_bfloat16* Wb = reinterpret_cast<__nv_bfloat16*>(w + p.off_W);
    PrepTcArgs pa{X, Y, dtype, ldx, ldy, g.m, g.n, p.mp, p.np, g.d, p.dp, p.dp, nullptr, nullptr, 0,
                  kf.tanh_features, 0, Z, norms, nullptr, kf, src.blk_x, src.blk_y};
    prep_tc_kernel<<<dim3((unsigned)((p.Mp + 7) / 8), 1), 256, 0, s>>>(pa);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*launches;
    const bool dot = kf.family == FAM_RQ && kf.add_dot > 0.f;
    if (dot) {
      colsum_kernel<<<dim3((unsigned)((p.dp + 31) / 32), 2), 256, 0, s>>>(Z, p.dp, p.dp, g.m, p.mp, g.n, csum);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      ++*launches;
    }
    CUtensorMap t1, t1b, tz, tw;
    if (!smmd_host::make_tmap_bf16_2d(&t1, Z, p.Mp, p.dp, p.dp, BM)) return cudaErrorUnknown;
    if (!smmd_host::make_tmap_bf16_2d(&t1b, Z, p.Mp, p.dp, p.dp, BNW)) return cudaErrorUnknown;
    if (!smmd_host::make_tmap_bf16_2d(&tz, Z, p.Mp, p.dp, p.dp, BNF)) return cudaErrorUnknown;
    if (!smmd_host::make_tmap_bf16_2d(&tw, Wb, (int64_t)p.P * BM, p.Mp, p.Mp, BM)) return cudaErrorUnknown;
    double* partials = reinterpret_cast<double*>(w + p.off_stats);
    prof_begin(s);
    for (int ip = 0; ip < p.npanels; ++ip) {
      const WzPanel q = wz_panel(p, ip);
      WgenArgs ga;
      ga.kf = kf;
      ga.m = g.m;
      ga.n = g.n;
      ga.mp = p.mp;
      ga.np = p.np;
      ga.c_xx = (float)(4.0 * c.a_xx);
      ga.c_yy = (float)(4.0 * c.a_yy);
      ga.c_xy = (float)(4.0 * c.a_xy);
      ga.norms = norms;
      ga.nrb_x = p.nrb_x;
      ga.rb_x0 = p.rb_x0;
      ga.nrb_y = p.nrb_y;
      ga.rb_y0 = p.rb_y0;
      ga.rbi0 = q.rbi0;
      ga.TX = p.TX;
      ga.CT = p.CT;
      ga.Wc = p.Wc;
      ga.nwin = p.nwin;
      ga.nrb_p = q.nrb_p;
      ga.spw = q.spw1;
      ga.nkp = (int)(p.dp / 64);
      ga.chunk = q.chunk1;
      ga.slots = q.slots1;
      ga.W = Wb;
      ga.ldw = p.Mp;
      ga.rpart = reinterpret_cast<float*>(w + p.off_r);
      ga.spart = reinterpret_cast<double*>(w + p.off_s);
      if (q.pair) *path = "tc_bf16_wz_pair";   // at least one panel ran pass 1 as CTA pairs (cta_group::2)
      if ((e = launch_wgen(variant, t1, t1b, ga, q.grid1, q.pair, s)) != cudaSuccess) return e;
      ++*launches;
      WzArgs za;
      za.nmb = q.nmb;
      za.FB = p.FB;
      za.dp = (int)p.dp;
      za.KT = p.KT;
      za.S = q.S;
      za.ksteps = q.ksteps;
      za.Opart = reinterpret_cast<float*>(w + p.off_O);
      if ((e = launch_wz(tw, tz, za, q.grid2, s)) != cudaSuccess) return e;
      ++*launches;
      WzFinArgs fr;
      fr.kf = kf;
      fr.m = g.m;
      fr.n = g.n;
      fr.mp = p.mp;
      fr.np = p.np;
      fr.d = g.d;
      fr.x0 = g.x0;
      fr.ox = g.x1 - g.x0;
      fr.y0 = g.y0;
      fr.oy = g.y1 - g.y0;
      fr.dp = (int)p.dp;
      fr.nrb_x = p.nrb_x;
      fr.rb_x0 = p.rb_x0;
      fr.nrb_y = p.nrb_y;
      fr.rb_y0 = p.rb_y0;
      fr.rbi0 = q.rbi0;
      fr.nrb_p = q.nrb_p;
      fr.Wc = p.Wc;
      fr.nwin = p.nwin;
      fr.spw = q.spw1;
      fr.pair = q.pair;
      fr.chunk = q.chunk1;
      fr.slots = q.slots1;
      fr.FB = p.FB;
      fr.S = q.S;
      fr.a_xx = c.a_xx;
      fr.a_yy = c.a_yy;
      fr.a_xy = c.a_xy;
      fr.src = src;
      fr.norms = norms;
      fr.csum = dot ? csum : nullptr;
      fr.Opart = za.Opart;
      fr.rpart = ga.rpart;
      fr.spart = ga.spart;
      fr.dX = dX;
      fr.dY = dY;
      fr.partials = partials + (int64_t)q.fin_block0 * 6;
      wz_finalize_rows_kernel<<<(unsigned)q.fin_blocks, 256, 0, s>>>(fr);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      ++*launches;
    }
    prof_end(s);
    e = launch_finalize_partials(kf, g, partials, (unsigned)p.fin_blocks_total, scalars, s);
    if (e != cudaSuccess) return e;
    ++*launches;
    return cudaSuccess;
  }
  if (want_grad) {
    const FusedPlan p = fused_plan(g.m, g.n, g.d, g.x0, g.x1, g.y0, g.y1);
    if (p.off_end > ws_bytes) return cudaErrorInvalidValue;
    *path = "tc_bf16_fused";
    __nv_bfloat16* Z = reinterpret_cast<__nv_bfloat16*>(w + p.off_Z);
    float* norms = reinterpret_cast<float*>(w + p.off_norm);
    double* csum = reinterpret_cast<double*>(w + p.off_csum);
    PrepTcArgs pa{X, Y, dtype, ldx, ldy, g.m, g.n, p.mp, p.np, g.d, p.dp, p.dp, nullptr, nullptr, 0,
                  kf.tanh_features, 0, Z, norms, nullptr, kf, src.blk_x, src.blk_y};
    prep_tc_kernel<<<dim3((unsigned)((p.Mp + 7) / 8), 1), 256, 0, s>>>(pa);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*launches;
    const bool dot = kf.family == FAM_RQ && kf.add_dot > 0.f;
    if (dot) {
      colsum_kernel<<<dim3((unsigned)((p.dp + 31) / 32), 2), 256, 0, s>>>(Z, p.dp, p.dp, g.m, p.mp, g.n, csum);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      ++*launches;
    }
    CUtensorMap tzi, tzj;
    if (!smmd_host::make_tmap_bf16_2d(&tzi, Z, p.Mp, p.dp, p.dp, BM)) return cudaErrorUnknown;
    if (!smmd_host::make_tmap_bf16_2d(&tzj, Z, p.Mp, p.dp, p.dp, BNF)) return cudaErrorUnknown;
    FusedArgs fa;
    fa.kf = kf;
    fa.m = g.m;
    fa.n = g.n;
    fa.mp = p.mp;
    fa.np = p.np;
    fa.c_xx = (float)(4.0 * c.a_xx);
    fa.c_yy = (float)(4.0 * c.a_yy);
    fa.c_xy = (float)(4.0 * c.a_xy);
    fa.norms = norms;
    fa.nrb_x = p.nrb_x;
    fa.rb_x0 = p.rb_x0;
    fa.nrb_y = p.nrb_y;
    fa.rb_y0 = p.rb_y0;
    fa.T = p.T;
    fa.dp = (int)p.dp;
    fa.npanel = (int)(p.dp / 64);
    fa.nst = fused_stages(fa.npanel);
    fa.ksplit = tuning().fused_ksplit;
    fa.total_tiles = p.total;
    fa.chunk = p.chunk;
    fa.slots = p.slots;
    fa.Opart = reinterpret_cast<float*>(w + p.off_O);
    fa.rpart = reinterpret_cast<float*>(w + p.off_r);
    fa.spart = reinterpret_cast<double*>(w + p.off_s);
    prof_begin(s);
    e = launch_fused(variant, tzi, tzj, fa, p.grid, s);
    prof_end(s);
    if (e != cudaSuccess) return e;
    ++*launches;
    FinRowsArgs fr;
    fr.kf = kf;
    fr.m = g.m;
    fr.n = g.n;
    fr.mp = p.mp;
    fr.np = p.np;
    fr.d = g.d;
    fr.x0 = g.x0;
    fr.ox = g.x1 - g.x0;
    fr.y0 = g.y0;
    fr.oy = g.y1 - g.y0;
    fr.dp = (int)p.dp;
    fr.nrb_x = p.nrb_x;
    fr.rb_x0 = p.rb_x0;
    fr.nrb_y = p.nrb_y;
    fr.rb_y0 = p.rb_y0;
    fr.T = p.T;
    fr.chunk = p.chunk;
    fr.slots = p.slots;
    fr.npart = 2 * fa.ksplit;
    fr.a_xx = c.a_xx;
    fr.a_yy = c.a_yy;
    fr.a_xy = c.a_xy;
    fr.Z = Z;
    fr.dpz = p.dp;
    fr.src = src;
    fr.norms = norms;
    fr.csum = dot ? csum : nullptr;
    fr.Opart = fa.Opart;
    fr.rpart = fa.rpart;
    fr.spart = fa.spart;
    fr.dX = dX;
    fr.dY = dY;
    fr.partials = reinterpret_cast<double*>(w + p.off_stats);
    const unsigned fin_blocks = (unsigned)((fr.ox + fr.oy + kFinRowsPerCta - 1) / kFinRowsPerCta);
    tc_finalize_rows_kernel<<<fin_blocks, 256, 0, s>>>(fr);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*launches;
    e = launch_finalize_partials(kf, g, fr.partials, fin_blocks, scalars, s);
    if (e != cudaSuccess) return e;
    ++*launches;
    return cudaSuccess;
  }
  // ---- value only: streaming kernel over the whole stacked Gram (world == 1 only for now) ----
  if (!(g.x0 == 0 && g.x1 == g.m && g.y0 == 0 && g.y1 == g.n)) return cudaErrorNotSupported;
  const int split = precision == SMMD_PREC_BF16X3;
  *path = split ? "tc_bf16x3_stream" : "tc_bf16_stream";
  const StreamPlan p = stream_plan(g.m, g.n, g.d, 1, split);
  if (p.off_end > ws_bytes) return cudaErrorInvalidValue;
  __nv_bfloat16* Z = reinterpret_cast<__nv_bfloat16*>(w + p.off_Z);
  float* norms = reinterpret_cast<float*>(w + p.off_norm);
  double* stats = reinterpret_cast<double*>(w + p.off_stats);
  if (kf.add_dot > 0.f) return cudaErrorNotSupported;  // value-only add_dot goes through the fused/SIMT paths
  PrepTcArgs pa{X, Y, dtype, ldx, ldy, g.m, g.n, p.mp, p.np, g.d, p.dp, p.dpz, nullptr, nullptr, 0,
                kf.tanh_features, split, Z, norms, stats, kf, src.blk_x, src.blk_y};
  prep_tc_kernel<<<dim3((unsigned)((p.Mp + 7) / 8), 1), 256, 0, s>>>(pa);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  ++*launches;
  CUtensorMap tm;
  if (!smmd_host::make_tmap_bf16_2d(&tm, Z, p.Mp, p.dpz, p.dpz, BM)) return cudaErrorUnknown;
  StreamArgs sa;
  sa.kf = kf;
  sa.m = g.m;
  sa.n = g.n;
  sa.mp = p.mp;
  sa.np = p.np;
  sa.RB = p.RB;
  sa.CT = p.CT;
  sa.nkp = (int)(p.dp / 64);
  sa.ncombo = split ? 3 : 1;
  sa.dp = p.dp;
  sa.total_tiles = p.total;
  sa.chunk = p.chunk;
  sa.norms = norms;
  sa.stats = stats;
  sa.want_sq = 0;
  prof_begin(s);
  e = launch_stream(variant, tm, sa, p.grid, s);
  prof_end(s);
  if (e != cudaSuccess) return e;
  ++*launches;
  e = launch_finalize_mmd2(kf, g, stats, norms, scalars, s);
  if (e != cudaSuccess) return e;
  ++*launches;
  return cudaSuccess;
}

bool tc_kid_supported(int64_t d) { return d >= 1 && d <= 65536; }

// KID runs on the 256 x 256 macro-tile kernel (polynomial kernels only reach this entry point)
size_t tc_kid_workspace_bytes(int64_t msub, int64_t d, int64_t nsub, int precision) {
  return macro_plan(msub, msub, d, nsub, precision == SMMD_PREC_BF16X3, 0).off_end;
}

cudaError_t tc_kid_run(const KernelFn& kf_in, const void* G, const void* R, int dtype, int64_t ldg, int64_t ldr, int64_t d,
                       const int32_t* idx_g, const int32_t* idx_r, int64_t first, int64_t nsub, int64_t msub,
                       int precision, int want_second_order, void* ws, size_t ws_bytes, double** stats_out,
                       cudaStream_t s, int* launches, const char** path) {
  KernelFn kf = kf_in;
  const TcVariant variant = select_tc_variant(kf);
  if (variant != TV_POLY3 && variant != TV_POLY_GENERIC) return cudaErrorNotSupported;
  const int split = precision == SMMD_PREC_BF16X3;
  const int sym = want_second_order ? 0 : 1;
  *path = split ? (sym ? "tc_bf16x3_kid_sym" : "tc_bf16x3_kid") : (sym ? "tc_bf16_kid_sym" : "tc_bf16_kid");
  const MacroPlan p = macro_plan(msub, msub, d, nsub, split, sym);
  if (p.off_end > ws_bytes) return cudaErrorInvalidValue;
  if ((int64_t)nsub * p.Mp >= ((int64_t)1 << 31) || p.R > 64) return cudaErrorInvalidValue;
  char* w = static_cast<char*>(ws);
  __nv_bfloat16* Z = reinterpret_cast<__nv_bfloat16*>(w + p.off_Z);
  float* norms = reinterpret_cast<float*>(w + p.off_norm);
  double* stats = reinterpret_cast<double*>(w + p.off_stats);
  cudaError_t e;
  PrepTcArgs pa{G, R, dtype, ldg, ldr, msub, msub, p.mp, p.np, d, p.dp, p.dpz, idx_g, idx_r, first,
                0, split, Z, norms, stats, kf, 0, 0};
  prep_tc_kernel<<<dim3((unsigned)((p.Mp + 7) / 8), (unsigned)nsub), 256, 0, s>>>(pa);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  ++*launches;
  CUtensorMap tm;
  if (!smmd_host::make_tmap_bf16_2d(&tm, Z, (uint64_t)nsub * p.Mp, p.dpz, p.dpz, BM)) return cudaErrorUnknown;
  MacroArgs ma;
  ma.kf = kf;
  ma.m = msub;
  ma.n = msub;
  ma.mp = p.mp;
  ma.np = p.np;
  ma.R = p.R;
  ma.Rx = p.Rx;
  ma.sym = sym;
  ma.tiles_per_batch = p.tiles_per_batch;
  ma.nkp = (int)(p.dp / 64);
  ma.ncombo = split ? 3 : 1;
  ma.dp = p.dp;
  ma.total_tiles = p.total;
  ma.chunk = p.chunk;
  ma.norms = norms;
  ma.stats = stats;
  ma.want_sq = want_second_order;
  prof_begin(s);
  e = variant == TV_POLY3 ? launch_macro_t<MathPoly3>(tm, ma, p.grid, s) : launch_macro_t<MathPolyN>(tm, ma, p.grid, s);
  prof_end(s);
  if (e != cudaSuccess) return e;
  ++*launches;
  *stats_out = stats;
  return cudaSuccess;
}

}  // namespace smmd

// smmd_tc_sym.cu -- symmetric two-pass MMD^2 forward + backward on tcgen05 (any d, whole problem on one GPU).
//
// The stacked Gram of Z = [X ; Y] is symmetric, and so is the weight matrix W = 4 A o k'(D) of the backward pass.
// The epilogue (3 MUFU + ~23 packed fp32 ops per element for the default RQ mixture), not the tensor pipe, bounds
// the loss at d <= 512, so this path evaluates every UNORDERED pair once:
//   pass 1 (tc_sym_wgen_kernel)  upper-triangle 128 x 256 tiles of S = Z_i Z_j^T (K streamed through a TMA ring,
//          two 256-column TMEM accumulators) -> kernel transform -> weighted block sums (off-diagonal blocks count
//          twice), row sums of W, and the bf16 W tile staged in shared memory (SW128 panels) and written with one
//          TMA store per 128 x 64 panel.  The column sums of an off-diagonal block (= its contribution to the row
//          sums r_j of the mirrored block) are taken from the staged bf16 tile.  Row / column sums go to a
//          fixed-point int64 accumulator (atomics whose result does not depend on their order).
//   pass 2 (tc_sym_wz_kernel)    O = W_sym Z: 256 rows x 256 features per CTA; for a column block J >= R the stored
//          tile W[R, J] is the K-major A operand, for J < R the stored tile W[J, R] is read "transposed" as an
//          MN-major A operand (descriptor LBO = 8 KB between the two 64-wide M panels) -- W is stored once and read
//          twice.  B = Z_j tile MN-major, all 512 TMEM columns as accumulators, K split over CTAs.
//   finalize (sym_finalize_rows_kernel)  g_i = r_i z_i - O_i (+ closed-form add_dot term), diagonal values.
// Executed tensor work 12 N^2 d (4 + 8) against 14 N^2 d algorithmic and 16 N^2 d for the row-stacked fused kernel;
// epilogue elements 2 N^2 instead of 4 N^2.  Replaces gan/core/mmd.py:55-188 + :194-220 + the TF autodiff graph
// (gan/core/model.py:446,452).
#include "smmd_tc_common.cuh"

namespace smmd {
namespace tc {
namespace {

constexpr int BNW = 256;                                   // pass-1 tile width
constexpr int kSyEpiWarps = 16;                            // 4 TMEM lane quarters x 4 column quarters
constexpr int kSyThreads = kRoleThreads;
constexpr int kSyStages = 3;
constexpr int kSyStageBytes = BM * 128 + BNW * 128;        // Z_i panel + Z_j tile per 64-deep step = 48 KB
constexpr int kSyStagingBytes = kSyEpiWarps * 4096;        // 64 KB: one 32-row x 64-column bf16 block per epilogue warp
constexpr int kSyNormBufs = 8;                             // column norms of the tiles in flight (1 KB each)
constexpr int kSyNormBytes = kSyNormBufs * BNW * 4;
constexpr int kSySmem = 1024 + kSyStages * kSyStageBytes + kSyStagingBytes + kSyNormBytes + 1024;
static_assert(kSySmem <= kMaxSmem, "pass-1 shared memory");

constexpr int kFinRowsPerWarp = 4;
constexpr int kFinRowsPerCta = 8 * kFinRowsPerWarp;

// ---- work units of pass 1 -----------------------------------------------------------------------------------
// The upper triangle is walked band by band (RB row blocks) and, inside a band, window by window (CW column
// tiles): the Z rows of one band + one window stay L2 resident while all CTAs work inside them.  A unit is a run
// of <= PL consecutive column tiles of ONE row block (row sums stay in registers over the run); units are numbered
// compactly in that order and dealt round-robin to the persistent CTAs.
struct SymGeo {
  int NB, CT;          // row blocks of 128, column tiles of 256 (the last tile may be half)
  int RB, CW, PL, PLs; // band (row blocks), window (tiles), piece length (tiles, a power of two) and its log2
  int nbands, nwin;
};

struct UnitWalk {
  int b, w, rb, rb_end, wst, wend;
  int64_t base;
  __device__ __forceinline__ void set_block(const SymGeo& g) {
    rb = b * g.RB;
    rb_end = rb + g.RB < g.NB ? rb + g.RB : g.NB;
    wst = w * g.CW;
    wend = wst + g.CW < g.CT ? wst + g.CW : g.CT;
  }
  __device__ __forceinline__ void init(const SymGeo& g) {
    b = 0;
    w = 0;
    base = 0;
    set_block(g);
  }
  // unit `idx` (non-decreasing over calls) -> (row block, [ct0, ct1)); false when idx is past the last unit.
  // Units are numbered compactly (rows / windows below the diagonal contribute none), so dealing them round-robin
  // balances the CTAs to within one unit.
  __device__ __forceinline__ bool locate(const SymGeo& g, int64_t idx, int& rb_out, int& ct0, int& ct1) {
    for (;;) {
      if (b >= g.nbands) return false;
      const int lo = wst > (rb >> 1) ? wst : (rb >> 1);   // first tile of row block rb inside this window
      const int nt = wend - lo;
      const int np = nt > 0 ? (nt + g.PL - 1) >> g.PLs : 0;
      if (idx < base + np) {
        const int k = (int)(idx - base);
        ct0 = lo + (k << g.PLs);
        ct1 = ct0 + g.PL < wend ? ct0 + g.PL : wend;
        rb_out = rb;
        return true;
      }
      base += np;
      if (++rb >= rb_end) {
        if (++w >= g.nwin) {
          ++b;
          w = ((b * g.RB) >> 1) / g.CW;   // windows left of it lie below the band's diagonal
        }
        if (b < g.nbands) set_block(g);
      }
    }
  }
};

struct SymWgenArgs {
  KernelFn kf;
  SymGeo geo;
  int64_t m, n, mp, np, Mp;
  float c_xx, c_yy, c_xy;        // 4 a_xx etc.
  const float* norms;            // [Mp]
  int nkp;                       // 64-feature panels
  float rscale, rclamp;          // fixed-point scale (power of two) / clamp of one contribution to the row sums of W
  unsigned long long* racc;      // [Mp] fixed-point row sums of W
  double* partials;              // [gridDim.x][6]: (sxx, syy, sxy, syx, 0, 0) of this CTA
};

__device__ __forceinline__ void tma_store_2d_hint(const void* tmap, const void* smem_src, int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void fixed_add(unsigned long long* p, float v, float scale, float clamp) {
  v = fminf(fmaxf(v, -clamp), clamp);
  atomicAdd(p, (unsigned long long)__float2ll_rn(v * scale));
}

__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// 16 columns of one row (as fused_chunk16), with the column norms read from shared memory and folded into the
// variant's first affine step (pair terms): no global load and no separate D = nij - 2 S in the element stream.
template <class Math, bool SPECIAL>
__device__ __forceinline__ void sym_chunk16(const Math& math, const uint32_t (&v)[16], uint32_t nj_smem, float rt,
                                            float2 cw, int col0, int lim, int gi, float2& tsum, float2& rsum,
                                            uint32_t (&wpk)[8]) {
  const float2 rt2 = bc2(rt);
#pragma unroll
  for (int c = 0; c < 16; c += 8) {   // 4 pairs (8 columns) evaluated in lock-step
    const float4 na = ld_shared_f4(nj_smem + c * 4);
    const float4 nb = ld_shared_f4(nj_smem + c * 4 + 16);
    float2 S[4], pt[4], k[4], kd[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) S[e] = make_float2(__uint_as_float(v[c + 2 * e]), __uint_as_float(v[c + 2 * e + 1]));
    pt[0] = math_pair_term(math, rt2, make_float2(na.x, na.y));
    pt[1] = math_pair_term(math, rt2, make_float2(na.z, na.w));
    pt[2] = math_pair_term(math, rt2, make_float2(nb.x, nb.y));
    pt[3] = math_pair_term(math, rt2, make_float2(nb.z, nb.w));
    eval_pairs_pt<Math, 4>(math, S, pt, k, kd);
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

// 64 columns of one row for a variant with a front/back split, software-pipelined over groups of 2 column pairs:
// front(g + 1) (affine steps + MUFU) is issued next to back(g) (FMA tail), see MathRq3Default::Pair.  TMEM loads are
// double-buffered per 16-column chunk; W goes to the warp's staging block.  `wait_prev` = the previous tile's TMA
// store may still be reading the staging block.
template <class Math, bool SPECIAL>
__device__ __forceinline__ void sym_quarter_split(const Math& math, uint32_t tbase, uint32_t nj_smem, float rt, float2 cw,
                                                  int col0, int lim, int gi, uint32_t srow, uint32_t sw, bool wait_prev,
                                                  int lane, uint64_t* acc_empty_bar, float2& tsum, float2& rsum) {
  using Pair = typename Math::Pair;
  const float2 rt2 = bc2(rt);
  uint32_t v[2][16];
  Pair st[2][2];
  auto run_front = [&](const uint32_t (&vv)[16], int g, int ch, Pair (&dst)[2]) {   // group g (0..3) of chunk ch
    const float4 nn = ld_shared_f4(nj_smem + (ch * 16 + g * 4) * 4);
    math.front(make_float2(__uint_as_float(vv[4 * g]), __uint_as_float(vv[4 * g + 1])),
               math.pair_term(rt2, make_float2(nn.x, nn.y)), dst[0]);
    math.front(make_float2(__uint_as_float(vv[4 * g + 2]), __uint_as_float(vv[4 * g + 3])),
               math.pair_term(rt2, make_float2(nn.z, nn.w)), dst[1]);
  };
  tmem_ld_x16(tbase, v[0]);
  tmem_ld_wait();
  tmem_ld_x16(tbase + 16, v[1]);
  run_front(v[0], 0, 0, st[0]);
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t wpk[8];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int cur = (ch * 4 + g) & 1;
      // front of the next group (next chunk's first group needs that chunk's TMEM load to have landed)
      if (g < 3) {
        run_front(v[ch & 1], g + 1, ch, st[cur ^ 1]);
      } else if (ch < 3) {
        tmem_ld_wait();
        run_front(v[(ch + 1) & 1], 0, ch + 1, st[cur ^ 1]);
      }
      float2 k[2], kd[2];
      math.back(st[cur][0], k[0], kd[0]);
      math.back(st[cur][1], k[1], kd[1]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (SPECIAL) {
          const int col = col0 + ch * 16 + g * 4 + 2 * e;
          const bool ok0 = (col < lim) && (col != gi), ok1 = (col + 1 < lim) && (col + 1 != gi);
          k[e] = make_float2(ok0 ? k[e].x : 0.f, ok1 ? k[e].y : 0.f);
          kd[e] = make_float2(ok0 ? kd[e].x : 0.f, ok1 ? kd[e].y : 0.f);
        }
        tsum = add2(tsum, k[e]);
        const float2 ww = mul2(kd[e], cw);
        rsum = add2(rsum, ww);
        wpk[2 * g + e] = pack_w<Math::kF16>(ww.x, ww.y);
      }
    }
    // chunk ch is consumed: its TMEM buffer takes chunk ch + 2 (the load of chunk ch + 1 was waited for above)
    if (ch + 2 < 4) tmem_ld_x16(tbase + (ch + 2) * 16, v[ch & 1]);
    if (ch == 0 && wait_prev) {
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
    }
    st_shared_v4(srow + (((2 * ch) ^ sw) << 4), wpk[0], wpk[1], wpk[2], wpk[3]);
    st_shared_v4(srow + (((2 * ch + 1) ^ sw) << 4), wpk[4], wpk[5], wpk[6], wpk[7]);
    if (ch == 2) {   // the last TMEM load (chunk 3) was waited for at g == 3 of this chunk: release the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty_bar);
    }
  }
}

template <class Math>
__global__ void __launch_bounds__(kSyThreads, 1)   // 20 warps at 96 registers, re-split by role (kEpiRegs / kCtlRegs)
tc_sym_wgen_kernel(const __grid_constant__ CUtensorMap tmap_zi, const __grid_constant__ CUtensorMap tmap_zj,
                   const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ SymWgenArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kSyStages * kSyStageBytes;
  float* nbuf = reinterpret_cast<float*>(staging + kSyStagingBytes);   // [kSyNormBufs][256] column norms per tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kSyStagingBytes + kSyNormBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kSyStages;
  uint64_t* acc_full = empty + kSyStages;    // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint64_t* nfull = acc_empty + 2;           // [kSyNormBufs] column norms of tile t landed in nbuf[t % 8]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(nfull + kSyNormBufs);
  float* sParams = reinterpret_cast<float*>(tmem_slot + 4);   // [24]
  double* sRed = reinterpret_cast<double*>(sParams + 24);     // [16][3] end-of-kernel reduction of the block sums
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_params(a.kf, sParams);
  if (tid == 0) {
    for (int i = 0; i < kSyStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kSyEpiWarps);   // one elected arrive per epilogue warp
    }
    for (int i = 0; i < kSyNormBufs; ++i) mbar_init(&nfull[i], 1);
    fence_mbar_init();
  }
  if (warp == kSyEpiWarps + 1) tmem_alloc<512>(tmem_slot);
  if (warp == kSyEpiWarps && lane == 0) {
    prefetch_tmap(&tmap_zi);
    prefetch_tmap(&tmap_zj);
    prefetch_tmap(&tmap_w);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const SymGeo& geo = a.geo;

  // (nested on purpose: ptxas allocates registers per setmaxnreg region only when each region is a branch of its own)
  if (warp >= kSyEpiWarps) {
  reg_dec<kCtlRegs>();
  if (warp == kSyEpiWarps) {
    // ===================== TMA producer =====================
    // The column norms of tile t go to nbuf[t % 8] with one 1 KB bulk copy.  Buffer reuse needs no barrier: the
    // producer is at most kSyStages = 3 tiles ahead of the issuer (one stage per tile when d <= 64), the issuer at most
    // 2 tiles (accumulators) ahead of the epilogue's release, and a warp that has released tile s may still be doing
    // the math of s: the oldest tile whose norms can still be read while tile t is loaded is t - 6 > t - 8.
    uint32_t st = 0, ph = 0, tcnt = 0;
    UnitWalk uw;
    uw.init(geo);
    int rb, ct0, ct1;
    for (int64_t u = blockIdx.x; uw.locate(geo, u, rb, ct0, ct1); u += gridDim.x) {
      for (int ct = ct0; ct < ct1; ++ct, ++tcnt) {
        for (int p = 0; p < a.nkp; ++p) {
          mbar_wait_sleep(&empty[st], ph ^ 1, 128);
          if (p == 0 && elect_one()) {
            const uint32_t nb = tcnt & (kSyNormBufs - 1);
            mbar_arrive_expect_tx(&nfull[nb], BNW * 4);
            bulk_load_1d(nbuf + nb * BNW, a.norms + (int64_t)ct * BNW, BNW * 4, &nfull[nb]);
          }
          if (elect_one()) {
            uint8_t* sa = smem + st * kSyStageBytes;
            mbar_arrive_expect_tx(&full[st], kSyStageBytes);
            tma_load_2d(sa, &tmap_zi, &full[st], p * 64, rb * BM);
            tma_load_2d(sa + BM * 128, &tmap_zj, &full[st], p * 64, ct * BNW);   // rows past Mp are zero-filled
          }
          __syncwarp();
          if (++st == kSyStages) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == kSyEpiWarps + 1) {
    // ===================== UMMA issuer =====================
    constexpr uint32_t idesc = make_idesc(BM, BNW, operand_fmt<Math>(), false, false);
    const uint32_t hi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem), 16);
    uint32_t st = 0, ph = 0, ab = 0, aph = 0;
    UnitWalk uw;
    uw.init(geo);
    int rb, ct0, ct1;
    for (int64_t u = blockIdx.x; uw.locate(geo, u, rb, ct0, ct1); u += gridDim.x) {
      for (int ct = ct0; ct < ct1; ++ct) {
        mbar_wait_sleep(&acc_empty[ab], aph ^ 1, 64);
        tc_fence_after();
        const uint32_t dad = tmem + ab * BNW;
        for (int kk = 0; kk < a.nkp; ++kk) {
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const uint32_t alo = a_lo0 + st * (kSyStageBytes >> 4), blo = alo + ((BM * 128) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss2(dad, alo + k * 2, blo + k * 2, hi, idesc, (kk | k) ? 1u : 0u);
            umma_commit(&empty[st]);
          }
          __syncwarp();
          if (++st == kSyStages) {
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
  }
  } else {
    reg_inc<kEpiRegs>();
    // ===================== epilogue: 16 warps = 4 TMEM lane quarters x 4 column quarters (64 columns each) =====
    // Every warp is self-contained: it owns 32 rows x 64 columns of each tile, stages its bf16 W block in its own
    // 4 KB of shared memory (SW128 atoms of 8 rows), stores it with its own TMA store and takes the column sums of
    // its block from the staged copy -- no barrier between warps, so the four warps of a scheduler drift apart and
    // their MUFU / FMA phases overlap (a first version synchronised the four row quarters of a column quarter three
    // times per tile through named barriers: 6850 cycles per 16K elements instead of the fused kernel's 4700).
    const int part = warp >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;                     // row inside the row block
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* stg_ptr = staging + warp * 4096;        // this warp's 32 rows x 128 B
    const uint32_t stg = smem_u32(stg_ptr);
    const Math math(a.kf, sParams);
    const float kscale = math.k_scale(), kdscale = math.kd_scale();
    const int mp = (int)a.mp, Mp = (int)a.Mp, mvalid = (int)a.m, yvalid = (int)(a.mp + a.n);
    double sxx = 0.0, syy = 0.0, sxy = 0.0;
    uint32_t tc = 0;        // tiles of this CTA so far (parity = accumulator buffer)
    bool stored = false;    // this warp has a TMA store in flight that may still read its staging block
    UnitWalk uw;
    uw.init(geo);
    int rb, ct0, ct1;
    for (int64_t u = blockIdx.x; uw.locate(geo, u, rb, ct0, ct1); u += gridDim.x) {
      if (ct0 >= ct1) continue;
      const int gi = rb * BM + r;
      const bool rowX = gi < mp;
      const bool row_ok = rowX ? gi < mvalid : gi < yvalid;
      const bool pad_rows = rowX ? (rb + 1) * BM > mvalid : (rb + 1) * BM > yvalid;   // block holds padding rows
      const float rt = math_row_term(math, a.norms[gi]);
      float2 rsum = make_float2(0.f, 0.f);
      float fsame = 0.f, fcross = 0.f;   // block sums of this unit (fp32 over <= 8 tiles; fp64 across units)
      bool any = false;
      for (int ct = ct0; ct < ct1; ++ct, ++tc) {
        const int grp = (int)(tc & 1);
        const int c0 = ct * BNW + part * 64;          // first column of this warp's quarter
        const int J = c0 >> 7;                        // its 128-column block
        const bool active = (J >= rb) && (c0 < Mp);
        mbar_wait_sleep(&acc_full[grp], (tc >> 1) & 1, 64);
        tc_fence_after();
        auto release_acc = [&]() {   // the whole warp has finished its tcgen05.ld of this tile: one elected arrive
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[grp]);
        };
        if (!active) {
          release_acc();
          continue;
        }
        any = true;
        const bool colX = c0 < mp;
        const bool same = (colX == rowX);
        const bool diag = (J == rb);
        const float2 cw = bc2((same ? (rowX ? a.c_xx : a.c_yy) : a.c_xy) * kdscale);
        const int lim = row_ok ? (colX ? mvalid : yvalid) : 0;   // padding rows: every column masked
        const bool special = diag || pad_rows || (c0 + 64 > (colX ? mvalid : yvalid));
        const uint32_t nb = tc & (kSyNormBufs - 1);
        mbar_wait(&nfull[nb], (tc / kSyNormBufs) & 1);
        const uint32_t nj = smem_u32(nbuf + nb * BNW + part * 64);
        float2 tsum = make_float2(0.f, 0.f);
        const uint32_t srow = stg + lane * 128;
        const uint32_t sw = (uint32_t)(lane & 7);
        if constexpr (Math::kHasSplit) {
          const uint32_t tbase = tmem + grp * BNW + part * 64 + lane_base;
          if (!special)
            sym_quarter_split<Math, false>(math, tbase, nj, rt, cw, 0, 0, 0, srow, sw, stored, lane, &acc_empty[grp], tsum, rsum);
          else
            sym_quarter_split<Math, true>(math, tbase, nj, rt, cw, c0, lim, gi, srow, sw, stored, lane, &acc_empty[grp], tsum,
                                          rsum);
        } else {
        auto do_chunk = [&](const uint32_t (&v)[16], int h) {
            uint32_t wpk[8];
            if (!special) sym_chunk16<Math, false>(math, v, nj + h * 64, rt, cw, 0, 0, 0, tsum, rsum, wpk);
            else sym_chunk16<Math, true>(math, v, nj + h * 64, rt, cw, c0 + h * 16, lim, gi, tsum, rsum, wpk);
            if (h == 0 && stored) {   // the previous tile's TMA store must have read the block before it is overwritten
              if (lane == 0) bulk_wait_read0();
              __syncwarp();
            }
            st_shared_v4(srow + (((2 * h) ^ sw) << 4), wpk[0], wpk[1], wpk[2], wpk[3]);
            st_shared_v4(srow + (((2 * h + 1) ^ sw) << 4), wpk[4], wpk[5], wpk[6], wpk[7]);
          };
          {
            uint32_t va[16], vb[16];
            const uint32_t tbase = tmem + grp * BNW + part * 64 + lane_base;
            tmem_ld_x16(tbase, va);
            tmem_ld_wait();
            tmem_ld_x16(tbase + 16, vb);
            do_chunk(va, 0);
            tmem_ld_wait();
            tmem_ld_x16(tbase + 32, va);
            do_chunk(vb, 1);
            tmem_ld_wait();
            tmem_ld_x16(tbase + 48, vb);
            do_chunk(va, 2);
            tmem_ld_wait();
            release_acc();
            do_chunk(vb, 3);
          }
        }
        // a cross block is seen once (X rows, Y columns) and stands for K_XY and K_YX; a same-set block above the
        // diagonal stands for itself and its mirror
        const float ts = (tsum.x + tsum.y) * kscale;
        if (!same) fcross += ts;
        else fsame += diag ? ts : 2.f * ts;
        // ---- staged block complete: TMA store, and (off-diagonal blocks) its column sums -> r_j of the mirror ----
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d_hint(&tmap_w, stg_ptr, c0, rb * BM + q * 32, kL2EvictFirst);
          bulk_commit();
        }
        stored = true;
        if (!diag) {
          // lane = column pair (2 lane, 2 lane + 1) of the block: sum its 32 rows from the staged bf16 values
          const uint32_t wofs = (uint32_t)(lane & 3) * 4;
          const uint32_t wch = (uint32_t)(lane >> 2);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint32_t p0 = ld_shared_u32(stg + i * 128 + ((wch ^ (uint32_t)(i & 7)) << 4) + wofs);
            const uint32_t p1 = ld_shared_u32(stg + (i + 1) * 128 + ((wch ^ (uint32_t)((i + 1) & 7)) << 4) + wofs);
            acc0 = add2(acc0, unpack_w<Math::kF16>(p0));
            acc1 = add2(acc1, unpack_w<Math::kF16>(p1));
          }
          const float2 t = add2(acc0, acc1);
          fixed_add(a.racc + c0 + 2 * lane, t.x, a.rscale, a.rclamp);
          fixed_add(a.racc + c0 + 2 * lane + 1, t.y, a.rscale, a.rclamp);
        }
      }
      if (any && row_ok) fixed_add(a.racc + gi, rsum.x + rsum.y, a.rscale, a.rclamp);
      sxy += (double)fcross;
      if (rowX) sxx += (double)fsame;
      else syy += (double)fsame;
    }
    if (stored && lane == 0) bulk_wait0();   // all W stores of this warp have completed before the CTA exits
    // ---- block sums of this CTA: fixed-order reduction over the 16 epilogue warps ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sxx += __shfl_xor_sync(0xffffffffu, sxx, o);
      syy += __shfl_xor_sync(0xffffffffu, syy, o);
      sxy += __shfl_xor_sync(0xffffffffu, sxy, o);
    }
    if (lane == 0) {
      sRed[warp * 3 + 0] = sxx;
      sRed[warp * 3 + 1] = syy;
      sRed[warp * 3 + 2] = sxy;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 6) {   // (sxx, syy, sxy, syx = sxy, 0, 0)
    double t = 0.0;
    const int src = tid == 3 ? 2 : tid;
    if (tid < 4)
      for (int w = 0; w < kSyEpiWarps; ++w) t += sRed[w * 3 + src];
    a.partials[(int64_t)blockIdx.x * 6 + tid] = t;
  }
  if (warp == kSyEpiWarps + 1) tmem_dealloc<512>(tmem);
}

// ---- fused pass 1 for d <= 256 (tc_symf_kernel): upper-triangle tiles, O_I += W Z_J in the same sweep -------------
// For d <= 256 the accumulator O_I [128 x d] fits tensor memory next to two 128-column S buffers, exactly as in the
// row-stacked fused kernel (smmd_tc_fused.cu: tile pairs, W written in place over S, UMMA #2 in TS form).  This kernel
// runs that pipeline over the UPPER TRIANGLE only: a unit is one row block I against a run of 128-column blocks
// J in [J0, J1), J >= I.  Per tile it does everything the row-stacked kernel does for the direct product
// (O_I += W[I,J] Z_J on the otherwise ~20%-busy tensor pipe), and additionally stages the bf16 W block in shared
// memory and stores it with TMA, so that a second, tensor-bound pass only has to add the MIRRORED products
// O_J += W[I,J]^T Z_I (tc_sym_wz_kernel with tonly = 1).  Against the plain symmetric path this halves pass 2 (and its
// reads of W); against the row-stacked kernel it halves the epilogue elements.
// Units: "full" units (row block above the window, all WB column blocks of a window) first, window-major -- the CTAs
// that work at the same time stream the same WB * 128 rows of Z_j out of L2 -- then the NB units that start on the
// diagonal, longest first; all dealt round-robin, so CTAs finish within one short unit of each other.
struct SymfGeo {
  int NB;        // row blocks of 128 (= column blocks of 128)
  int WB;        // column blocks per window
  int nwin;      // windows
  int last_len;  // column blocks in the last window (<= WB)
  int64_t nfull, nunits;
};
struct SymfUnit {
  int rb, J0, J1;
};
__host__ __device__ inline int64_t symf_full_base(const SymfGeo& g, int w) { return (int64_t)g.WB * w * (w - 1) / 2; }
__host__ __device__ inline int64_t symf_full_index(const SymfGeo& g, int w, int rb) { return symf_full_base(g, w) + rb; }
__host__ __device__ inline int64_t symf_diag_index(const SymfGeo& g, int w, int k) {   // unit of row block w * WB + k
  const int64_t A = (int64_t)g.last_len * g.nwin;
  return g.nfull + (k < g.last_len ? (int64_t)k * g.nwin + w : A + (int64_t)(k - g.last_len) * (g.nwin - 1) + w);
}
__device__ __forceinline__ SymfUnit symf_locate(const SymfGeo& g, int64_t u) {
  SymfUnit r;
  if (u < g.nfull) {
    int w = 1;
    while (symf_full_base(g, w + 1) <= u) ++w;
    r.rb = (int)(u - symf_full_base(g, w));
    r.J0 = w * g.WB;
    r.J1 = r.J0 + g.WB < g.NB ? r.J0 + g.WB : g.NB;
  } else {
    const int64_t t = u - g.nfull, A = (int64_t)g.last_len * g.nwin;
    int k, w;
    if (t < A) {
      k = (int)(t / g.nwin);
      w = (int)(t - (int64_t)k * g.nwin);
    } else {
      const int64_t v = t - A;
      k = g.last_len + (int)(v / (g.nwin - 1));
      w = (int)(v % (g.nwin - 1));
    }
    r.rb = w * g.WB + k;
    r.J0 = r.rb;
    r.J1 = (w + 1) * g.WB < g.NB ? (w + 1) * g.WB : g.NB;
  }
  return r;
}

struct SymfArgs {
  KernelFn kf;
  SymfGeo geo;
  int64_t m, n, mp, np, Mp;
  float c_xx, c_yy, c_xy;
  const float* norms;
  int dp, npanel;
  float rscale, rclamp;
  unsigned long long* racc;      // [Mp] fixed-point row sums of W (direct rows + mirrored columns)
  float* Opart;                  // [nunits][128][dp] partial O of the direct products
  double* partials;              // [gridDim.x][6]
};

constexpr uint32_t TMF_S = 256;                 // TMEM: O[dp <= 256] | pair buffer 0 [128] | pair buffer 1 [128]
constexpr int kSfStages = 4;
constexpr int kSfZjBytes = BNF * 128;           // one 64-wide k-panel of a 64-row column tile
constexpr int kSfZiBytes = BM * 128;
constexpr int kSfStagingBytes = 16 * 2048;      // one 32-row x 32-column bf16 block per epilogue warp
constexpr int kSfThreads = kRoleThreads;
inline int symf_smem(int npanel) { return 1024 + npanel * kSfZiBytes + kSfStages * npanel * kSfZjBytes + kSfStagingBytes + 1536; }

template <class Math>
__global__ void __launch_bounds__(kSfThreads, 1)   // 20 warps (two idle, they complete the last warpgroup): 96 registers at launch, see reg_inc below
tc_symf_kernel(const __grid_constant__ CUtensorMap tmap_zi, const __grid_constant__ CUtensorMap tmap_zj,
               const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ SymfArgs a) {
  constexpr int KSPLIT = 2, NPART = 4, EPI_WARPS = 16, NST = kSfStages;
  const int NPANEL = a.npanel, DP = a.dp;
  const int ZI_BYTES = NPANEL * kSfZiBytes;
  constexpr int PANEL_STRIDE = NST * kSfZjBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sZi = smem;
  uint8_t* sZj = smem + ZI_BYTES;                     // [NPANEL][NST][64 rows x 128 B]
  uint8_t* staging = sZj + NPANEL * PANEL_STRIDE;     // 2 KB aligned: [16 warps][32 rows][64 B], 64-byte swizzle
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kSfStagingBytes);
  uint64_t* zj_full = bars;             // [NST]
  uint64_t* zj_empty = bars + NST;      // [NST]
  uint64_t* s_full = bars + 2 * NST;    // [2]
  uint64_t* w_full = s_full + 2;        // [2][2]  (pair buffer, group), see tc_fused_pair_kernel
  uint64_t* pb_free = w_full + 4;       // [2]
  uint64_t* zi_full = pb_free + 2;
  uint64_t* zi_empty = zi_full + 1;
  uint64_t* o_full = zi_empty + 1;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);
  float* sParams = reinterpret_cast<float*>(tmem_slot + 4);      // [24]
  // column norms of the tile in ring stage st: copied by the producer together with the tile (same mbarrier), read by the
  // epilogue from shared memory -- a global load at the head of every 16-column chunk cost 25% of the epilogue's time
  float* nbuf = sParams + 24 + 8;                                  // [NST][64] (16-byte aligned)
  double* sRed = reinterpret_cast<double*>(staging);              // [16][3], after the last W store has drained

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_params(a.kf, sParams);
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&zj_full[i], 1);
      mbar_init(&zj_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&w_full[2 * i], 128 * KSPLIT);
      mbar_init(&w_full[2 * i + 1], 128 * KSPLIT);
      mbar_init(&pb_free[i], 1);
    }
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
    prefetch_tmap(&tmap_w);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const SymfGeo& geo = a.geo;

  // Register budget: 20 warps launch with 96 registers each; the control warpgroup (producer, issuer, two idle warps) hands
  // registers back and the four epilogue warpgroups take them (setmaxnreg): the epilogue's per-tile state then stays in
  // registers instead of being re-derived under the 96-register cap.
  if (warp >= EPI_WARPS) {
   reg_dec<kCtlRegs>();
   if (warp == EPI_WARPS) {
    // ===================== TMA producer =====================
    uint32_t unit = 0, st = 0, ph = 0;
    for (int64_t u = blockIdx.x; u < geo.nunits; u += gridDim.x, ++unit) {
      const SymfUnit un = symf_locate(geo, u);
      mbar_wait_sleep(zi_empty, (unit & 1) ^ 1, 128);
      if (elect_one()) {
        mbar_arrive_expect_tx(zi_full, ZI_BYTES);
        for (int p = 0; p < NPANEL; ++p) tma_load_2d(sZi + p * (BM * 128), &tmap_zi, zi_full, p * 64, un.rb * BM);
      }
      __syncwarp();
      for (int t = 2 * un.J0; t < 2 * un.J1; ++t) {
        mbar_wait_sleep(&zj_empty[st], ph ^ 1, 128);
        if (elect_one()) {
          mbar_arrive_expect_tx(&zj_full[st], NPANEL * kSfZjBytes + BNF * 4);
          uint8_t* dst = sZj + st * kSfZjBytes;
          for (int p = 0; p < NPANEL; ++p) tma_load_2d(dst + p * PANEL_STRIDE, &tmap_zj, &zj_full[st], p * 64, t * BNF);
          bulk_load_1d(nbuf + st * BNF, a.norms + (int64_t)t * BNF, BNF * 4, &zj_full[st]);
        }
        __syncwarp();
        if (++st == (uint32_t)NST) {
          st = 0;
          ph ^= 1;
        }
      }
    }
   } else if (warp == EPI_WARPS + 1) {
    // ===================== UMMA issuer (as tc_fused_pair_kernel) =====================
    constexpr uint32_t idesc1 = make_idesc(BM, 2 * BNF, operand_fmt<Math>(), false, false);
    const uint32_t idesc2 = make_idesc(BM, (uint32_t)DP, operand_fmt<Math>(), false, true);
    const uint32_t hi = desc_hi_sw128(1024);
    const uint32_t zi_lo = desc_lo(smem_u32(sZi), 16);
    const uint32_t zj_lo1 = desc_lo(smem_u32(sZj), 16);
    const uint32_t zj_lo2 = desc_lo(smem_u32(sZj), PANEL_STRIDE);
    constexpr uint32_t stage_step = (uint32_t)kSfZjBytes >> 4, panel_step = (uint32_t)PANEL_STRIDE >> 4;
    uint32_t unit = 0;
    uint32_t st1 = 0, ph1 = 0, gp1 = 0;
    uint32_t st2 = 0, gp2 = 0;
    for (int64_t u = blockIdx.x; u < geo.nunits; u += gridDim.x, ++unit) {
      const SymfUnit un = symf_locate(geo, u);
      const int NP = un.J1 - un.J0;
      mbar_wait(zi_full, unit & 1);
      for (int j = 0; j < NP + 2; ++j) {
        const int pm2 = j - 2;
        if (pm2 >= 0) {
          const uint32_t pb = gp2 & 1;
#pragma unroll
          for (int g = 0; g < 2; ++g) {   // ---- UMMA #2 for tile g of pair pm2: O += W * Zj
            mbar_wait(&w_full[pb * 2 + g], (gp2 >> 1) & 1);
            if (pm2 == 0 && g == 0) mbar_wait(o_empty, (unit & 1) ^ 1);
            tc_fence_after();
            const uint32_t blo = zj_lo2 + st2 * stage_step;
            const uint32_t wad = tmem + TMF_S + pb * 128 + g * 64;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < BNF / 16; ++kk)
                umma_ts2(tmem, wad + (kk >> 1) * 32 + (kk & 1) * 8, blo + kk * (2048 >> 4), hi, idesc2,
                         (pm2 > 0 || g > 0 || kk > 0) ? 1u : 0u);
              umma_commit(&zj_empty[st2]);
              if (g == 1) umma_commit(&pb_free[pb]);
              if (pm2 == NP - 1 && g == 1) umma_commit(o_full);
            }
            __syncwarp();
            if (++st2 == (uint32_t)NST) st2 = 0;
          }
          ++gp2;
        }
        if (j < NP) {   // ---- UMMA #1 for pair j: S[128 x 128] = Zi * [Zj(2J); Zj(2J+1)]^T
          const uint32_t pb = gp1 & 1;
          mbar_wait(&zj_full[st1], ph1);
          mbar_wait(&zj_full[st1 + 1], ph1);
          if (gp1 >= 2) mbar_wait(&pb_free[pb], ((gp1 >> 1) - 1) & 1);
          tc_fence_after();
          const uint32_t blo = zj_lo1 + st1 * stage_step;
          const uint32_t sad = tmem + TMF_S + pb * 128;
          if (elect_one()) {
            for (int p = 0; p < NPANEL; ++p) {
              const uint32_t ap = zi_lo + p * ((BM * 128) >> 4), bp = blo + p * panel_step;
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_ss2(sad, ap + k * 2, bp + k * 2, hi, idesc1, (p | k) ? 1u : 0u);
            }
            umma_commit(&s_full[pb]);
            if (j == NP - 1) umma_commit(zi_empty);
          }
          __syncwarp();
          st1 += 2;
          if (st1 == (uint32_t)NST) {
            st1 = 0;
            ph1 ^= 1;
          }
          ++gp1;
        }
      }
    }
   }
  } else {
    reg_inc<kEpiRegs>();
    // ===================== epilogue: group g takes tile g of every pair; warp = 32 rows x 32 columns =====================
    const int grp = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int part = grp * KSPLIT + half;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* stg_ptr = staging + warp * 2048;
    const uint32_t stg = smem_u32(stg_ptr);
    const uint32_t srow = stg + lane * 64;
    const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
    const Math math(a.kf, sParams);
    const float kscale = math.k_scale(), kdscale = math.kd_scale();
    const int mp = (int)a.mp, mvalid = (int)a.m, yvalid = (int)(a.mp + a.n);
    double sxx = 0.0, syy = 0.0, sxy = 0.0;
    uint32_t unit = 0, gp = 0;
    bool stored = false;
    for (int64_t u = blockIdx.x; u < geo.nunits; u += gridDim.x, ++unit) {
      const SymfUnit un = symf_locate(geo, u);
      const int rb = un.rb;
      const int gi = rb * BM + r;
      const bool rowX = gi < mp;
      const bool row_ok = rowX ? gi < mvalid : gi < yvalid;
      const bool pad_rows = rowX ? (rb + 1) * BM > mvalid : (rb + 1) * BM > yvalid;
      const float rt = math_row_term(math, a.norms[gi]);
      float2 rsum = make_float2(0.f, 0.f);
      float fsame = 0.f, fcross = 0.f;
      for (int J = un.J0; J < un.J1; ++J, ++gp) {
        const uint32_t pb = gp & 1;
        // ring stage of this group's tile: the stages advance by two per pair (NST = 4: pair parity picks the half)
        const uint32_t nj = smem_u32(nbuf + (((gp & 1) << 1) + grp) * BNF + half * 32);
        const int c0 = (2 * J + grp) * BNF;            // first column of this group's 64-column tile
        const int cb = c0 + half * 32;                 // first column of this warp's 32-column block
        const bool colX = c0 < mp;
        const bool same = (colX == rowX);
        const bool diag = (J == rb);
        const float2 cw = bc2((same ? (rowX ? a.c_xx : a.c_yy) : a.c_xy) * kdscale);
        const int cvalid = colX ? mvalid : yvalid;
        const int lim = row_ok ? cvalid : 0;           // padding rows: every column masked (W = 0 there)
        const bool special = diag || pad_rows || (c0 + BNF > cvalid);
        const uint32_t s_addr = tmem + TMF_S + pb * 128 + grp * 64 + half * 32 + lane_base;
        float2 tsum = make_float2(0.f, 0.f);
        mbar_wait(&s_full[pb], (gp >> 1) & 1);
        tc_fence_after();
        {
          uint32_t va[16], vb[16], wpk[8];
          tmem_ld_x16(s_addr, va);
          tmem_ld_wait();
          tmem_ld_x16(s_addr + 16, vb);
          if (!special) sym_chunk16<Math, false>(math, va, nj, rt, cw, 0, 0, 0, tsum, rsum, wpk);
          else sym_chunk16<Math, true>(math, va, nj, rt, cw, cb, lim, gi, tsum, rsum, wpk);
          tmem_st_x8(s_addr, wpk);            // W in place: columns [0, 8) of this warp's slice (already read)
          if (stored) {                       // the previous tile's TMA store must have read the staging block
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
          }
          st_shared_v4(srow + ((0u ^ sw) << 4), wpk[0], wpk[1], wpk[2], wpk[3]);
          st_shared_v4(srow + ((1u ^ sw) << 4), wpk[4], wpk[5], wpk[6], wpk[7]);
          tmem_ld_wait();
          if (!special) sym_chunk16<Math, false>(math, vb, nj + 64, rt, cw, 0, 0, 0, tsum, rsum, wpk);
          else sym_chunk16<Math, true>(math, vb, nj + 64, rt, cw, cb + 16, lim, gi, tsum, rsum, wpk);
          tmem_st_x8(s_addr + 8, wpk);
          st_shared_v4(srow + ((2u ^ sw) << 4), wpk[0], wpk[1], wpk[2], wpk[3]);
          st_shared_v4(srow + ((3u ^ sw) << 4), wpk[4], wpk[5], wpk[6], wpk[7]);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&w_full[pb * 2 + grp]);
        const float ts = (tsum.x + tsum.y) * kscale;
        if (!same) fcross += ts;
        else fsame += diag ? ts : 2.f * ts;
        // ---- off the issuer's critical path: W block -> HBM (for the mirrored products of pass 2), column sums ----
        if (!diag) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // W is kept column-group-major for this path: [Mp / 64 column groups][Mp rows][64 columns], so the two
            // 32-column halves of a row quarter interleave into one contiguous 4 KB region and pass 2 reads a transposed
            // 64 x 64 operand box as 8 KB of consecutive memory (row-major W made every 128-byte row segment a separate
            // DRAM page visit: the mirrored pass ran at 66% of the HBM rate)
            tma_store_2d_hint(&tmap_w, stg_ptr, cb & 63, (cb >> 6) * (int32_t)a.Mp + rb * BM + q * 32, kL2EvictFirst);
            bulk_commit();
          }
          stored = true;
          // lanes 0..15 / 16..31 take the even / odd rows; lane & 15 = column pair
          const uint32_t pcol = (uint32_t)(lane & 15), hrow = (uint32_t)(lane >> 4);
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t rr = 2 * i + hrow;
            const uint32_t wv = ld_shared_u32(stg + rr * 64 + ((((pcol >> 2) ^ ((uint32_t)i & 3u))) << 4) + (pcol & 3u) * 4);
            acc = add2(acc, unpack_w<Math::kF16>(wv));
          }
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
          acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
          if (lane < 16) {
            fixed_add(a.racc + cb + 2 * lane, acc.x, a.rscale, a.rclamp);
            fixed_add(a.racc + cb + 2 * lane + 1, acc.y, a.rscale, a.rclamp);
          }
        } else {
          __syncwarp();   // (diagonal blocks are not stored: nothing mirrors them; staging stays owned by this warp)
        }
      }
      // ---- unit end: row sums, block sums, drain O ----
      if (row_ok) {
        fixed_add(a.racc + gi, rsum.x + rsum.y, a.rscale, a.rclamp);
        sxy += (double)fcross;
        if (rowX) sxx += (double)fsame;
        else syy += (double)fsame;
      }
      mbar_wait(o_full, unit & 1);
      tc_fence_after();
      {
        const int seg = DP / NPART;
        float* orow = a.Opart + ((int64_t)u * BM + r) * DP + part * seg;
        for (int c = 0; c < seg; c += 16) {
          uint32_t v[16];
          tmem_ld_x16(tmem + part * seg + c + lane_base, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            *reinterpret_cast<float4*>(orow + c + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                   __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(o_empty);
    }
    if (stored && lane == 0) bulk_wait0();
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sxx += __shfl_xor_sync(0xffffffffu, sxx, o);
      syy += __shfl_xor_sync(0xffffffffu, syy, o);
      sxy += __shfl_xor_sync(0xffffffffu, sxy, o);
    }
    named_bar_sync(1, EPI_WARPS * 32);   // every epilogue warp's stores have drained: the staging area is free
    if (lane == 0) {
      sRed[warp * 3 + 0] = sxx;
      sRed[warp * 3 + 1] = syy;
      sRed[warp * 3 + 2] = sxy;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 6) {
    double t = 0.0;
    const int src = tid == 3 ? 2 : tid;
    if (tid < 4)
      for (int w = 0; w < EPI_WARPS; ++w) t += sRed[w * 3 + src];
    a.partials[(int64_t)blockIdx.x * 6 + tid] = t;
  }
  if (warp == EPI_WARPS + 1) tmem_dealloc<512>(tmem);
}

// ---- pass 2: O = W_sym Z -----------------------------------------------------------------------------------
constexpr int kSzStages = 3;
constexpr int kSzStageBytes = 2 * BM * 128 + 4 * BNF * 128;   // two 128-row W operands + four 64-feature Z panels = 64 KB
constexpr int kSzSmem = 1024 + kSzStages * kSzStageBytes + 1024;

struct SymWzArgs {
  int nmb;        // 256-row macro blocks
  int FB;         // 256-feature blocks
  int dp;         // padded feature count (multiple of 64)
  int KT;         // K steps of 64 (= Mp / 64)
  int S;          // K splits per unit
  int ksteps;     // K steps per piece
  float* Opart;   // [unit = mb * FB + fb][S][256][256]
  // mirrored-only mode (after tc_symf_kernel, which has already added the direct products): macro block mb only takes
  // the stored tiles W[J, R] with J < R, read transposed, i.e. K steps [0, 4 mb + 2) (its second row block reaches two
  // steps further than its first).  Pieces are runs of 4 * ell K steps; macro block mb has mb / ell + 1 of them and
  // pieces are numbered mb-major: piece = (P(mb) + s) * FB + fb, P(mb) = sum_{i < mb} (i / ell + 1).
  int tonly, ell;
  int Mp;         // stacked padded rows (tonly: row stride of a column group of W)
  int f16;        // operands are IEEE half (fp16 tier) instead of bf16
};
__host__ __device__ inline int64_t symt_prefix(int mb, int ell) {   // P(mb)
  const int64_t q = mb / ell, r = mb % ell;
  return mb + (int64_t)ell * q * (q - 1) / 2 + q * r;
}

__global__ void __launch_bounds__(kThreads, 1)
tc_sym_wz_kernel(const __grid_constant__ CUtensorMap tmap_wd, const __grid_constant__ CUtensorMap tmap_wt,
                 const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ SymWzArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSzStages * kSzStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kSzStages;
  uint64_t* acc_full = empty + kSzStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kSzStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmap_wd);
    prefetch_tmap(&tmap_wt);
    prefetch_tmap(&tmap_z);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  int s, u, mb, fb, k0, k1;
  if (!a.tonly) {
    // piece = (split s, unit u), s-major so that one wave of CTAs walks the same rows of Z (L2 reuse)
    const int units = a.nmb * a.FB;
    s = (int)blockIdx.x / units;
    u = (int)blockIdx.x - s * units;
    mb = u / a.FB;
    fb = u - mb * a.FB;
    k0 = s * a.ksteps;
    k1 = k0 + a.ksteps < a.KT ? k0 + a.ksteps : a.KT;
  } else {
    const int64_t pc = (int64_t)blockIdx.x / a.FB;
    fb = (int)((int64_t)blockIdx.x - pc * a.FB);
    int q = 0;                                   // pieces of macro blocks [q ell, (q + 1) ell): (q + 1) each
    while (symt_prefix((q + 1) * a.ell, a.ell) <= pc) ++q;
    const int64_t rem = pc - symt_prefix(q * a.ell, a.ell);
    mb = q * a.ell + (int)(rem / (q + 1));
    s = (int)(rem % (q + 1));
    u = mb * a.FB + fb;
    k0 = s * 4 * a.ell;
    const int kend = 4 * mb + 2 < a.KT ? 4 * mb + 2 : a.KT;
    k1 = k0 + 4 * a.ell < kend ? k0 + 4 * a.ell : kend;
  }
  // mirrored-only mode: row block h of the macro block is active for K steps ks < 4 mb + 2 h
  const int hlim0 = a.tonly ? 4 * mb : a.KT, hlim1 = a.tonly ? 4 * mb + 2 : a.KT;
  const int nf = a.dp - fb * 256 < 256 ? a.dp - fb * 256 : 256;   // features of this block (multiple of 64)
  const int npan = nf / 64;

  if (warp == 8) {
    uint32_t st = 0, ph = 0;
    for (int ks = k0; ks < k1; ++ks) {
      mbar_wait(&empty[st], ph ^ 1);
      if (elect_one()) {
        const int nact = (ks < hlim0 ? 1 : 0) + (ks < hlim1 ? 1 : 0);
        mbar_arrive_expect_tx(&full[st], nact * BM * 128 + npan * BNF * 128);
        uint8_t* sa = smem + st * kSzStageBytes;
        const int J = ks >> 1;   // 128-column block of this K step
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int R = 2 * mb + h;
          uint8_t* dst = sa + h * (BM * 128);
          if (ks >= (h ? hlim1 : hlim0)) continue;
          // W streams through once per direction (evict-first: it must not displace Z, which every CTA re-reads)
          if (!a.tonly && J >= R) {   // stored tile W[R, J]: 128 rows x 64 columns, K-major A
            tma_load_2d_hint(dst, &tmap_wd, &full[st], ks * 64, R * BM, kL2EvictFirst);
          } else if (!a.tonly) {   // stored tile W[J, R] read transposed: 64 K-rows x 128 columns = two 64 x 64 boxes, MN-major A
            tma_load_2d_hint(dst, &tmap_wt, &full[st], R * BM, ks * 64, kL2EvictFirst);
            tma_load_2d_hint(dst + 64 * 128, &tmap_wt, &full[st], R * BM + 64, ks * 64, kL2EvictFirst);
          } else {                 // same operand from the column-group-major W of the fused variant (see tc_symf_kernel)
            tma_load_2d_hint(dst, &tmap_wt, &full[st], 0, (2 * R) * a.Mp + ks * 64, kL2EvictFirst);
            tma_load_2d_hint(dst + 64 * 128, &tmap_wt, &full[st], 0, (2 * R + 1) * a.Mp + ks * 64, kL2EvictFirst);
          }
        }
        uint8_t* sb = sa + 2 * BM * 128;
        for (int p = 0; p < npan; ++p)
          tma_load_2d_hint(sb + p * (BNF * 128), &tmap_z, &full[st], fb * 256 + p * 64, ks * 64, kL2EvictLast);
      }
      __syncwarp();
      if (++st == kSzStages) {
        st = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 9) {
    const uint32_t fmt = a.f16 ? kFmtF16 : kFmtBF16;
    const uint32_t idesc_d = make_idesc(BM, (uint32_t)nf, fmt, false, true);   // A K-major, B = Z tile MN-major
    const uint32_t idesc_t = make_idesc(BM, (uint32_t)nf, fmt, true, true);    // A MN-major (transposed tile)
    const uint32_t hi = desc_hi_sw128(1024);
    const uint32_t ad_lo0 = desc_lo(smem_u32(smem), 16);
    const uint32_t at_lo0 = desc_lo(smem_u32(smem), 64 * 128);   // LBO = distance between the two 64-wide M panels
    const uint32_t b_lo0 = desc_lo(smem_u32(smem + 2 * BM * 128), BNF * 128);
    uint32_t st = 0, ph = 0;
    for (int ks = k0; ks < k1; ++ks) {
      mbar_wait(&full[st], ph);
      tc_fence_after();
      const uint32_t sofs = st * (kSzStageBytes >> 4);
      const uint32_t blo = b_lo0 + sofs;
      const int J = ks >> 1;
      if (elect_one()) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (ks >= (h ? hlim1 : hlim0)) continue;
          const bool direct = !a.tonly && J >= 2 * mb + h;
          const uint32_t hofs = sofs + h * ((BM * 128) >> 4);
          if (direct) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss2(tmem + h * 256, ad_lo0 + hofs + k * 2, blo + k * (2048 >> 4), hi, idesc_d, (ks > k0 || k > 0) ? 1u : 0u);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss2(tmem + h * 256, at_lo0 + hofs + k * (2048 >> 4), blo + k * (2048 >> 4), hi, idesc_t,
                       (ks > k0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[st]);
        if (ks == k1 - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++st == kSzStages) {
        st = 0;
        ph ^= 1;
      }
    }
  } else {
    // drain: warp -> (row half, TMEM lane quarter); each thread stores its row of the partial tile
    const int half = warp >> 2, q = warp & 3;
    const int row = half * BM + q * 32 + lane;
    float* orow = a.tonly ? a.Opart + ((int64_t)blockIdx.x * 256 + row) * 256
                          : a.Opart + (((int64_t)u * a.S + s) * 256 + row) * 256;
    if (k1 > k0 && k0 < (half ? hlim1 : hlim0)) {
      mbar_wait_sleep(acc_full, 0, 1000);   // the drain warps idle for the whole K loop: no polling next to the issuer
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

// ---- finalize: warp per row, lanes over features -------------------------------------------------------------
struct SymFinArgs {
  KernelFn kf;
  int64_t m, n, mp, np, d;
  int dp;
  int FB, S;                 // pass-2 feature blocks / K splits
  double a_xx, a_yy, a_xy;
  double rinv;               // 1 / fixed-point scale of racc
  SrcLayout src;
  const float* norms;
  const double* csum;
  const float* Opart;
  const unsigned long long* racc;
  float* dX;
  float* dY;
  double* partials;          // [gridDim.x][6]: (dot same X, dot same Y, dot cross X, dot cross Y, dgx, dgy)
  // fused pass 1 (tc_symf_kernel): O_i = direct slabs of the row block's units + mirrored pieces of pass 2
  int symf, ell;
  SymfGeo fgeo;
  const float* Odir;         // [nunits][128][dp]
  float gscale;              // 1 / (power-of-two scale W was carried with), see w_scale_for()
};

// sum of the partial O values of one row at features [c, c + 4) (c % 4 == 0, inside one 256-feature block), fixed order
__device__ __forceinline__ float4 sym_row_o4(const SymFinArgs& a, int rbl, int r, int c) {
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  auto add = [&](const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    o.x += t.x;
    o.y += t.y;
    o.z += t.z;
    o.w += t.w;
  };
  const int mbl = rbl >> 1, row256 = (rbl & 1) * BM + r;
  if (!a.symf) {
    const float* op = a.Opart + (((int64_t)mbl * a.FB + (c >> 8)) * a.S * 256 + row256) * 256 + (c & 255);
    for (int s = 0; s < a.S; ++s) add(op + (int64_t)s * 65536);
    return o;
  }
  const int w0 = rbl / a.fgeo.WB, k = rbl - w0 * a.fgeo.WB;
  add(a.Odir + ((int64_t)symf_diag_index(a.fgeo, w0, k) * BM + r) * a.dp + c);
  for (int w = w0 + 1; w < a.fgeo.nwin; ++w) add(a.Odir + ((int64_t)symf_full_index(a.fgeo, w, rbl) * BM + r) * a.dp + c);
  // mirrored pieces of macro block mbl: row block h = rbl & 1 is active in piece s while 4 ell s < 4 mbl + 2 h
  const int64_t P = symt_prefix(mbl, a.ell);
  const int np = mbl / a.ell + 1, klim = 4 * mbl + 2 * (rbl & 1);
  for (int s = 0; s < np && 4 * a.ell * s < klim; ++s) add(a.Opart + (((P + s) * a.FB + (c >> 8)) * 256 + row256) * 256 + (c & 255));
  return o;
}
__device__ __forceinline__ float sym_row_o1(const SymFinArgs& a, int rbl, int r, int c) {
  float o = 0.f;
  const int mbl = rbl >> 1, row256 = (rbl & 1) * BM + r;
  if (!a.symf) {
    const float* op = a.Opart + (((int64_t)mbl * a.FB + (c >> 8)) * a.S * 256 + row256) * 256 + (c & 255);
    for (int s = 0; s < a.S; ++s) o += op[(int64_t)s * 65536];
    return o;
  }
  const int w0 = rbl / a.fgeo.WB, k = rbl - w0 * a.fgeo.WB;
  o += a.Odir[((int64_t)symf_diag_index(a.fgeo, w0, k) * BM + r) * a.dp + c];
  for (int w = w0 + 1; w < a.fgeo.nwin; ++w) o += a.Odir[((int64_t)symf_full_index(a.fgeo, w, rbl) * BM + r) * a.dp + c];
  const int64_t P = symt_prefix(mbl, a.ell);
  const int np = mbl / a.ell + 1, klim = 4 * mbl + 2 * (rbl & 1);
  for (int s = 0; s < np && 4 * a.ell * s < klim; ++s) o += a.Opart[(((P + s) * a.FB + (c >> 8)) * 256 + row256) * 256 + (c & 255)];
  return o;
}

__global__ void __launch_bounds__(256) sym_finalize_rows_kernel(SymFinArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double sh[8][6];
  double q[6] = {0, 0, 0, 0, 0, 0};
  const bool dot = a.kf.family == FAM_RQ && a.kf.add_dot > 0.f && a.csum != nullptr;
  for (int rr = 0; rr < kFinRowsPerWarp; ++rr) {
    const int64_t gi = ((int64_t)blockIdx.x * 8 + warp) * kFinRowsPerWarp + rr;   // padded stacked row
    if (gi >= a.mp + a.np) break;
    const bool rowX = gi < a.mp;
    const int64_t li = rowX ? gi : gi - a.mp;
    if (li >= (rowX ? a.m : a.n)) continue;   // padding row
    const float rs = (float)((double)(long long)a.racc[gi] * a.rinv);
    const double a_same = rowX ? a.a_xx : a.a_yy;
    double dsame = 0.0, dcross = 0.0;
    float* out = nullptr;
    if (a.dX) out = rowX ? a.dX + li * a.d : a.dY + li * a.d;
    const void* src = rowX ? a.src.X : a.src.Y;
    const int64_t ld = rowX ? a.src.ldx : a.src.ldy;
    const int sdtype = a.src.dtype;
    const int64_t srow = src_row(li, rowX, a.src.blk_x, a.src.blk_y);
    const int rbl = (int)(gi / BM), r = (int)(gi % BM);
    const bool vec = out != nullptr && !dot && sdtype == SMMD_F32 && (a.d % 4 == 0) && (ld % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec) {
      const float* zsrc = reinterpret_cast<const float*>(src) + srow * ld;
      for (int c = 4 * lane; c < a.d; c += 128) {   // 4 consecutive features per lane (never straddle a 256 block)
        const float4 o = sym_row_o4(a, rbl, r, c);   // fixed order
        const float4 z4 = *reinterpret_cast<const float4*>(zsrc + c);
        float zz[4] = {z4.x, z4.y, z4.z, z4.w};
        const float oo[4] = {o.x, o.y, o.z, o.w};
        float gv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (a.kf.tanh_features) zz[e] = tanhf(zz[e]);
          gv[e] = (rs * zz[e] - oo[e]) * a.gscale;
          if (a.kf.tanh_features) gv[e] *= (1.f - zz[e] * zz[e]);
        }
        *reinterpret_cast<float4*>(out + c) = make_float4(gv[0], gv[1], gv[2], gv[3]);
      }
    } else {
      for (int c = lane; c < a.d; c += 32) {
        float o = 0.f;
        if (out) o = sym_row_o1(a, rbl, r, c);   // fixed order
        const int64_t sidx = srow * ld + c;
        float z = sdtype == SMMD_F32 ? reinterpret_cast<const float*>(src)[sidx]
                                      : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[sidx]);
        if (a.kf.tanh_features) z = tanhf(z);
        if (out) {
          float gv = (rs * z - o) * a.gscale;
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
    // the dot part of the kernel is closed form: sum_{j != i} <z_i, z_j> = <z_i, colsum> - |z_i|^2
    const float ni = a.norms[gi];
    const double v_same = dot ? (double)a.kf.add_dot * (dsame - (double)ni) : 0.0;
    const double v_cross = dot ? (double)a.kf.add_dot * dcross : 0.0;
    double v_diag = a.kf.family == FAM_RQ ? (double)a.kf.const_diag + (double)a.kf.add_dot * (double)ni
                                          : (double)diag_value(a.kf, ni);
    // the unclamped epilogue math is finite for |z|^2 <= 1e18; beyond that the result is reported as non-finite
    if (!(ni <= 1.0e18f)) v_diag = __longlong_as_double(0x7ff8000000000000LL);
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

// ---- plan --------------------------------------------------------------------------------------------------
struct SymPlan {
  int64_t mp, np, Mp, dp;
  SymGeo geo;
  int grid1;
  int KT, FB, nmb, units, S, ksteps, grid2;
  int fin_blocks;
  size_t off_Z, off_norm, off_csum, off_W, off_racc, off_O, off_stats, off_end;
};

// K splits per unit, chosen by a small cost model (cycles): a piece costs its K steps (8 UMMAs of 128 x 256 x 16 = 1024
// tensor cycles per step) plus a fixed ~14K cycles (TMEM allocation, pipeline fill, 256 KB drain); pieces run in waves of
// #SMs; every piece also adds 512 KB of slab traffic (written here, read by the finalize kernel).  The first version
// maximised wave occupancy alone and cut N = 8192, d = 512 into 1920 pieces of 17 K steps: 273 us for a 140 us GEMM.
int sym_choose_split(int units, int KT) {
  const int sm = sm_count();
  int best = 1;
  double best_cost = 1e300;
  for (int S = 1; S <= 16 && KT / S >= 8; ++S) {
    const int64_t pieces = (int64_t)units * S;
    const int64_t waves = (pieces + sm - 1) / sm;
    const double ksteps = (double)((KT + S - 1) / S);
    const double cost = (double)waves * (ksteps * 1024.0 + 14000.0) + (double)pieces * 153.0;
    if (cost < best_cost) {
      best_cost = cost;
      best = S;
    }
  }
  return best;
}

SymPlan sym_plan(int64_t m, int64_t n, int64_t d) {
  SymPlan p;
  p.mp = round_up(m, BM);
  p.np = round_up(n, BM);
  p.Mp = p.mp + p.np;
  p.dp = round_up(d, 64);
  SymGeo& g = p.geo;
  g.NB = (int)(p.Mp / BM);
  g.CT = (g.NB + 1) / 2;
  // band / window: ~16 MB of Z rows each (both stay L2 resident while every CTA works inside them)
  const int64_t rows16 = std::max<int64_t>(1024, ((int64_t)16 << 20) / (p.dp * 2));
  g.RB = (int)std::min<int64_t>(g.NB, rows16 / BM);
  g.CW = (int)std::min<int64_t>(g.CT, rows16 / BNW);
  if (p.Mp * p.dp * 2 <= ((int64_t)64 << 20)) {   // Z fits L2: one band, one window
    g.RB = g.NB;
    g.CW = g.CT;
  }
  g.PL = 8;
  // small problems: shorter pieces, so that every CTA gets >= 16 units and the round-robin deal leaves < 6% tail
  while (g.PL > 1 && (int64_t)g.NB * g.CT / 2 / g.PL < 16 * sm_count()) g.PL >>= 1;
  g.PLs = 0;
  while ((1 << g.PLs) < g.PL) ++g.PLs;
  g.PL = 1 << g.PLs;
  g.nbands = (g.NB + g.RB - 1) / g.RB;
  g.nwin = (g.CT + g.CW - 1) / g.CW;
  p.grid1 = sm_count();
  p.KT = (int)(p.Mp / 64);
  p.FB = (int)((p.dp + 255) / 256);
  p.nmb = (g.NB + 1) / 2;
  p.units = p.nmb * p.FB;
  p.S = sym_choose_split(p.units, p.KT);
  p.ksteps = (p.KT + p.S - 1) / p.S;
  p.grid2 = p.units * p.S;
  p.fin_blocks = (int)((p.Mp + kFinRowsPerCta - 1) / kFinRowsPerCta);
  size_t o = 0;
  p.off_Z = o;
  o = up256(o + (size_t)p.Mp * p.dp * 2);
  p.off_norm = o;
  o = up256(o + (size_t)(p.Mp + BNW) * 4);   // + one tile: the bulk copy of the last (half) tile's column norms
  p.off_csum = o;
  o = up256(o + (size_t)2 * p.dp * 8);
  p.off_W = o;
  o = up256(o + (size_t)p.Mp * p.Mp * 2);
  p.off_racc = o;
  o = up256(o + (size_t)p.Mp * 8);
  p.off_O = o;
  o = up256(o + (size_t)p.units * p.S * 256 * 256 * 4);
  p.off_stats = o;
  o = up256(o + (size_t)(p.grid1 + p.fin_blocks) * 6 * 8);
  p.off_end = o;
  return p;
}

template <class Math>
cudaError_t launch_sym_wgen_t(const CUtensorMap& tzi, const CUtensorMap& tzj, const CUtensorMap& tw, const SymWgenArgs& a,
                              int grid, cudaStream_t s) {
  auto kern = tc_sym_wgen_kernel<Math>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSySmem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kSyThreads, kSySmem, s>>>(tzi, tzj, tw, a);
  return cudaGetLastError();
}
cudaError_t launch_sym_wgen(TcVariant v, bool f16, const CUtensorMap& tzi, const CUtensorMap& tzj, const CUtensorMap& tw,
                            const SymWgenArgs& a, int grid, cudaStream_t s) {
  if (f16) {
    switch (v) {
      case TV_RBF1: return launch_sym_wgen_t<F16Of<MathRbf1>>(tzi, tzj, tw, a, grid, s);
      case TV_RBF_LADDER5: return launch_sym_wgen_t<F16Of<MathRbfLadder<5>>>(tzi, tzj, tw, a, grid, s);
      case TV_RBF_GENERIC: return launch_sym_wgen_t<F16Of<MathGeneric<FAM_RBF>>>(tzi, tzj, tw, a, grid, s);
      case TV_RQ3_DEFAULT: return launch_sym_wgen_t<F16Of<MathRq3Default>>(tzi, tzj, tw, a, grid, s);
      case TV_RQ_GENERIC: return launch_sym_wgen_t<F16Of<MathGeneric<FAM_RQ>>>(tzi, tzj, tw, a, grid, s);
      case TV_DISTANCE: return launch_sym_wgen_t<F16Of<MathDistance>>(tzi, tzj, tw, a, grid, s);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (v) {
    case TV_RBF1: return launch_sym_wgen_t<MathRbf1>(tzi, tzj, tw, a, grid, s);
    case TV_RBF_LADDER5: return launch_sym_wgen_t<MathRbfLadder<5>>(tzi, tzj, tw, a, grid, s);
    case TV_RBF_GENERIC: return launch_sym_wgen_t<MathGeneric<FAM_RBF>>(tzi, tzj, tw, a, grid, s);
    case TV_RQ3_DEFAULT: return launch_sym_wgen_t<MathRq3Default>(tzi, tzj, tw, a, grid, s);
    case TV_RQ_GENERIC: return launch_sym_wgen_t<MathGeneric<FAM_RQ>>(tzi, tzj, tw, a, grid, s);
    case TV_DISTANCE: return launch_sym_wgen_t<MathDistance>(tzi, tzj, tw, a, grid, s);
    case TV_NULL: return launch_sym_wgen_t<MathNull>(tzi, tzj, tw, a, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

// ---- plan of the fused variant (d <= 256) ---------------------------------------------------------------------
struct SymfPlan {
  int64_t mp, np, Mp, dp;
  SymfGeo geo;
  int grid1, nmb, ell, KT;
  int64_t pieces;
  int fin_blocks;
  size_t off_Z, off_norm, off_csum, off_W, off_racc, off_O1, off_O2, off_stats, off_end;
};

SymfPlan symf_plan(int64_t m, int64_t n, int64_t d) {
  SymfPlan p;
  p.mp = round_up(m, BM);
  p.np = round_up(n, BM);
  p.Mp = p.mp + p.np;
  p.dp = round_up(d, 64);
  SymfGeo& g = p.geo;
  g.NB = (int)(p.Mp / BM);
  // window = the largest power of two of column blocks (<= 256: 16 MB of Z_j at d = 256) that still leaves every CTA
  // >= 20 full units, so that the round-robin deal balances to a few percent
  g.WB = 256;
  auto nfull_of = [&](int wb) {
    const int64_t nw = (g.NB + wb - 1) / wb;
    return (int64_t)wb * nw * (nw - 1) / 2;
  };
  // (every unit ends with a 128 KB drain of O, so units are kept long: >= 4 full units per CTA are enough because the
  // diagonal units, dealt longest first after them, even the CTAs out to within one short unit)
  while (g.WB > 4 && nfull_of(g.WB) < (int64_t)4 * sm_count()) g.WB >>= 1;
  g.nwin = (g.NB + g.WB - 1) / g.WB;
  g.last_len = g.NB - (g.nwin - 1) * g.WB;
  g.nfull = nfull_of(g.WB);
  g.nunits = g.nfull + g.NB;
  p.grid1 = (int)std::min<int64_t>(sm_count(), g.nunits);
  p.nmb = (g.NB + 1) / 2;
  p.KT = (int)(p.Mp / 64);
  // mirrored pass: pieces of 4 * ell K steps (one CTA each, ~equal length): the longest pieces that still fill whole
  // waves of SMs to >= 90% (a piece pays a fixed prologue + a 256 KB drain, so short pieces waste the tensor pipe)
  {
    const int sm = sm_count();
    int best = 2;
    double best_eff = -1.0;
    for (int ell = 64; ell >= 2; ell >>= 1) {
      const int64_t pc = symt_prefix(p.nmb, ell);
      const double eff = (double)pc / (double)((pc + sm - 1) / sm * sm);
      if (pc >= 2 * sm && eff >= 0.9) {
        best = ell;
        break;
      }
      if (eff > best_eff) {
        best_eff = eff;
        best = ell;
      }
    }
    p.ell = best;
  }
  p.pieces = symt_prefix(p.nmb, p.ell);
  p.fin_blocks = (int)((p.Mp + kFinRowsPerCta - 1) / kFinRowsPerCta);
  size_t o = 0;
  p.off_Z = o;
  o = up256(o + (size_t)p.Mp * p.dp * 2);
  p.off_norm = o;
  o = up256(o + (size_t)p.Mp * 4);
  p.off_csum = o;
  o = up256(o + (size_t)2 * p.dp * 8);
  p.off_W = o;
  o = up256(o + (size_t)p.Mp * p.Mp * 2);
  p.off_racc = o;
  o = up256(o + (size_t)p.Mp * 8);
  p.off_O1 = o;
  o = up256(o + (size_t)g.nunits * BM * p.dp * 4);
  p.off_O2 = o;
  o = up256(o + (size_t)p.pieces * 256 * 256 * 4);
  p.off_stats = o;
  o = up256(o + (size_t)(p.grid1 + p.fin_blocks) * 6 * 8);
  p.off_end = o;
  return p;
}

template <class Math>
cudaError_t launch_symf_t(const CUtensorMap& tzi, const CUtensorMap& tzj, const CUtensorMap& tw, const SymfArgs& a, int grid,
                          cudaStream_t s) {
  const int smem = symf_smem(a.npanel);
  if (smem > kMaxSmem) return cudaErrorInvalidConfiguration;
  auto kern = tc_symf_kernel<Math>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kSfThreads, smem, s>>>(tzi, tzj, tw, a);
  return cudaGetLastError();
}
cudaError_t launch_symf(TcVariant v, bool f16, const CUtensorMap& tzi, const CUtensorMap& tzj, const CUtensorMap& tw, const SymfArgs& a,
                        int grid, cudaStream_t s) {
  if (f16) {
    switch (v) {
      case TV_RBF1: return launch_symf_t<F16Of<MathRbf1>>(tzi, tzj, tw, a, grid, s);
      case TV_RBF_LADDER5: return launch_symf_t<F16Of<MathRbfLadder<5>>>(tzi, tzj, tw, a, grid, s);
      case TV_RBF_GENERIC: return launch_symf_t<F16Of<MathGeneric<FAM_RBF>>>(tzi, tzj, tw, a, grid, s);
      case TV_RQ3_DEFAULT: return launch_symf_t<F16Of<MathRq3Default>>(tzi, tzj, tw, a, grid, s);
      case TV_RQ_GENERIC: return launch_symf_t<F16Of<MathGeneric<FAM_RQ>>>(tzi, tzj, tw, a, grid, s);
      case TV_DISTANCE: return launch_symf_t<F16Of<MathDistance>>(tzi, tzj, tw, a, grid, s);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (v) {
    case TV_RBF1: return launch_symf_t<MathRbf1>(tzi, tzj, tw, a, grid, s);
    case TV_RBF_LADDER5: return launch_symf_t<MathRbfLadder<5>>(tzi, tzj, tw, a, grid, s);
    case TV_RBF_GENERIC: return launch_symf_t<MathGeneric<FAM_RBF>>(tzi, tzj, tw, a, grid, s);
    case TV_RQ3_DEFAULT: return launch_symf_t<MathRq3Default>(tzi, tzj, tw, a, grid, s);
    case TV_RQ_GENERIC: return launch_symf_t<MathGeneric<FAM_RQ>>(tzi, tzj, tw, a, grid, s);
    case TV_DISTANCE: return launch_symf_t<MathDistance>(tzi, tzj, tw, a, grid, s);
    case TV_NULL: return launch_symf_t<MathNull>(tzi, tzj, tw, a, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t run_symf(const KernelFn& kf, TcVariant variant, const Geometry& g, const Coefs& c, const SrcLayout& src,
                     double* scalars, float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches,
                     const char** path) {
  char* w = static_cast<char*>(ws);
  cudaError_t e;
  *path = "tc_bf16_symf";
  const SymfPlan p = symf_plan(g.m, g.n, g.d);
  if (p.off_end > ws_bytes) return cudaErrorInvalidValue;
  __nv_bfloat16* Z = reinterpret_cast<__nv_bfloat16*>(w + p.off_Z);
  float* norms = reinterpret_cast<float*>(w + p.off_norm);
  double* csum = reinterpret_cast<double*>(w + p.off_csum);
  __nv_bfloat16* Wb = reinterpret_cast<__nv_bfloat16*>(w + p.off_W);
  unsigned long long* racc = reinterpret_cast<unsigned long long*>(w + p.off_racc);
  double* partials = reinterpret_cast<double*>(w + p.off_stats);
  PrepTcArgs pa{src.X, src.Y, src.dtype, src.ldx, src.ldy, g.m, g.n, p.mp, p.np, g.d, p.dp, p.dp, nullptr, nullptr, 0,
                kf.tanh_features, 0, Z, norms, nullptr, kf, src.blk_x, src.blk_y};
  pa.f16 = c.f16;
  prep_set_peers(pa, src);
  const double wscale = w_scale_for(c, kf);
  if ((e = launch_prep_tc(pa, p.Mp, 1, s)) != cudaSuccess) return e;
  peer_after_pull(s);   // (peer calls: the caller's "rows pulled" event, see smmd_peer_set_pull_event)
  ++*launches;
  if ((e = cudaMemsetAsync(racc, 0, (size_t)p.Mp * 8, s)) != cudaSuccess) return e;
  const bool dot = kf.family == FAM_RQ && kf.add_dot > 0.f;
  if (dot) {
    if ((e = launch_colsum_tc(Z, p.dp, p.dp, g.m, p.mp, g.n, csum, c.f16, s)) != cudaSuccess) return e;
    ++*launches;
  }
  CUtensorMap tzi, tzj, tz, twd, twt, tws;
  if (!smmd_host::make_tmap_bf16_2d(&tzi, Z, p.Mp, p.dp, p.dp, BM)) return cudaErrorUnknown;
  if (!smmd_host::make_tmap_bf16_2d(&tzj, Z, p.Mp, p.dp, p.dp, BNF)) return cudaErrorUnknown;
  if (!smmd_host::make_tmap_bf16_2d(&tz, Z, p.Mp, p.dp, p.dp, BNF)) return cudaErrorUnknown;
  // W column-group-major: a 2-D matrix of (Mp / 64) * Mp rows x 64 columns (row = column group * Mp + row of W)
  const uint64_t wrows = (uint64_t)(p.Mp / 64) * (uint64_t)p.Mp;
  if (!smmd_host::make_tmap_bf16_2d(&twd, Wb, wrows, 64, 64, BM)) return cudaErrorUnknown;   // (unused by the mirrored pass)
  if (!smmd_host::make_tmap_bf16_2d(&twt, Wb, wrows, 64, 64, 64)) return cudaErrorUnknown;
  if (!smmd_host::make_tmap_bf16_2d_box32(&tws, Wb, wrows, 64, 64, 32)) return cudaErrorUnknown;   // pass-1 stores
  const double cmax = 4.0 * wscale * std::max(std::max(fabs(c.a_xx), fabs(c.a_yy)), fabs(c.a_xy));
  const double wb = cmax * kd_abs_bound(kf);
  int ex = 0;
  frexp(wb * (double)p.Mp, &ex);
  const int shift = std::max(-100, std::min(100, 61 - ex));
  SymfArgs fa;
  fa.kf = kf;
  fa.geo = p.geo;
  fa.m = g.m;
  fa.n = g.n;
  fa.mp = p.mp;
  fa.np = p.np;
  fa.Mp = p.Mp;
  fa.c_xx = (float)(4.0 * c.a_xx * wscale);
  fa.c_yy = (float)(4.0 * c.a_yy * wscale);
  fa.c_xy = (float)(4.0 * c.a_xy * wscale);
  fa.norms = norms;
  fa.dp = (int)p.dp;
  fa.npanel = (int)(p.dp / 64);
  fa.rscale = (float)ldexp(1.0, shift);
  fa.rclamp = (float)(wb * (double)p.Mp);
  fa.racc = racc;
  fa.Opart = reinterpret_cast<float*>(w + p.off_O1);
  fa.partials = partials;
  prof_begin(s);
#ifdef SMMD_DEV_KNOBS
  const int only = tuning().sym_only;
#else
  const int only = 0;
#endif
  if (only != 2) {
    if ((e = launch_symf(variant, c.f16 != 0, tzi, tzj, tws, fa, p.grid1, s)) != cudaSuccess) return e;
    ++*launches;
  }
  prof_mark(s);
  SymWzArgs za;
  za.nmb = p.nmb;
  za.FB = 1;
  za.dp = (int)p.dp;
  za.KT = p.KT;
  za.S = 1;
  za.ksteps = p.KT;
  za.Opart = reinterpret_cast<float*>(w + p.off_O2);
  za.tonly = 1;
  za.ell = p.ell;
  za.Mp = (int)p.Mp;
  za.f16 = c.f16;
  if ((e = cudaFuncSetAttribute(tc_sym_wz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSzSmem)) != cudaSuccess)
    return e;
  if (only != 1) {
    tc_sym_wz_kernel<<<(unsigned)p.pieces, kThreads, kSzSmem, s>>>(twd, twt, tz, za);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*launches;
  }
  prof_end(s);
  SymFinArgs fr;
  fr.kf = kf;
  fr.m = g.m;
  fr.n = g.n;
  fr.mp = p.mp;
  fr.np = p.np;
  fr.d = g.d;
  fr.dp = (int)p.dp;
  fr.FB = 1;
  fr.S = 1;
  fr.a_xx = c.a_xx;
  fr.a_yy = c.a_yy;
  fr.a_xy = c.a_xy;
  fr.rinv = ldexp(1.0, -shift);
  fr.src = src;
  fr.norms = norms;
  fr.csum = dot ? csum : nullptr;
  fr.Opart = za.Opart;
  fr.racc = racc;
  fr.dX = dX;
  fr.dY = dY;
  fr.partials = partials + (int64_t)p.grid1 * 6;
  fr.symf = 1;
  fr.ell = p.ell;
  fr.fgeo = p.geo;
  fr.Odir = fa.Opart;
  fr.gscale = (float)(1.0 / wscale);
  sym_finalize_rows_kernel<<<(unsigned)p.fin_blocks, 256, 0, s>>>(fr);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  ++*launches;
  e = launch_finalize_partials(kf, g, partials, (int64_t)p.grid1 + p.fin_blocks, scalars, s);
  if (e != cudaSuccess) return e;
  ++*launches;
  return cudaSuccess;
}

}  // namespace

bool tc_sym_eligible(const Geometry& g) {
  if (!tuning().sym) return false;
  if (!(g.x0 == 0 && g.x1 == g.m && g.y0 == 0 && g.y1 == g.n)) return false;   // whole problem on this GPU
  const int64_t Mp = round_up(g.m, BM) + round_up(g.n, BM);
  // measured cross-over against the row-stacked kernels (warm single calls, mix_rq): d = 256: N = 2048 per side 0.073 vs
  // 0.079 ms, 8192: 0.331 vs 0.346, 16384: 1.055 vs 1.165; d = 512: N = 2048 0.085 vs 0.099, N = 1024 0.066 vs 0.071
  const int64_t min_rows = tuning().sym_min_rows > 0 ? tuning().sym_min_rows : 4096;
  if (Mp < min_rows) return false;
  return Mp * Mp * 2 <= tuning().sym_max_w_bytes;
}

// d <= 256: the fused variant (direct products inside pass 1, mirrored products in pass 2)
// (from ~20K rows per side on: N = 16384 1.139 vs 1.055 ms for the unfused pair, 24576: 2.29 vs 2.39, 65536: 14.7 vs 16.7;
// below that the 128 KB O drain per unit and the short mirrored pieces cost more than the halved second pass saves)
static bool use_symf(const Geometry& g) {
  if (!tuning().symf || round_up(g.d, 64) > 256) return false;
  const int64_t Mp = round_up(g.m, BM) + round_up(g.n, BM);
  return Mp >= (tuning().symf_min_rows > 0 ? tuning().symf_min_rows : 40960);
}

size_t tc_sym_workspace_bytes(const Geometry& g) {
  return use_symf(g) ? symf_plan(g.m, g.n, g.d).off_end : sym_plan(g.m, g.n, g.d).off_end;
}

cudaError_t tc_run_sym(const KernelFn& kf, TcVariant variant, const Geometry& g, const Coefs& c, const SrcLayout& src,
                       double* scalars, float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches,
                       const char** path) {
  if (use_symf(g)) return run_symf(kf, variant, g, c, src, scalars, dX, dY, ws, ws_bytes, s, launches, path);
  char* w = static_cast<char*>(ws);
  cudaError_t e;
  *path = "tc_bf16_sym";
  const SymPlan p = sym_plan(g.m, g.n, g.d);
  if (p.off_end > ws_bytes) return cudaErrorInvalidValue;
  __nv_bfloat16* Z = reinterpret_cast<__nv_bfloat16*>(w + p.off_Z);
  float* norms = reinterpret_cast<float*>(w + p.off_norm);
  double* csum = reinterpret_cast<double*>(w + p.off_csum);
  __nv_bfloat16* Wb = reinterpret_cast<__nv_bfloat16*>(w + p.off_W);
  unsigned long long* racc = reinterpret_cast<unsigned long long*>(w + p.off_racc);
  double* partials = reinterpret_cast<double*>(w + p.off_stats);
  PrepTcArgs pa{src.X, src.Y, src.dtype, src.ldx, src.ldy, g.m, g.n, p.mp, p.np, g.d, p.dp, p.dp, nullptr, nullptr, 0,
                kf.tanh_features, 0, Z, norms, nullptr, kf, src.blk_x, src.blk_y};
  pa.f16 = c.f16;
  prep_set_peers(pa, src);
  const double wscale = w_scale_for(c, kf);
  if ((e = launch_prep_tc(pa, p.Mp, 1, s)) != cudaSuccess) return e;
  peer_after_pull(s);   // (peer calls: the caller's "rows pulled" event, see smmd_peer_set_pull_event)
  ++*launches;
  if ((e = cudaMemsetAsync(racc, 0, (size_t)p.Mp * 8, s)) != cudaSuccess) return e;
  const bool dot = kf.family == FAM_RQ && kf.add_dot > 0.f;
  if (dot) {
    if ((e = launch_colsum_tc(Z, p.dp, p.dp, g.m, p.mp, g.n, csum, c.f16, s)) != cudaSuccess) return e;
    ++*launches;
  }
  CUtensorMap tzi, tzj, tz, twd, twt, tws;
  if (!smmd_host::make_tmap_bf16_2d(&tzi, Z, p.Mp, p.dp, p.dp, BM)) return cudaErrorUnknown;
  if (!smmd_host::make_tmap_bf16_2d(&tzj, Z, p.Mp, p.dp, p.dp, BNW)) return cudaErrorUnknown;
  if (!smmd_host::make_tmap_bf16_2d(&tz, Z, p.Mp, p.dp, p.dp, BNF)) return cudaErrorUnknown;
  if (!smmd_host::make_tmap_bf16_2d(&twd, Wb, p.Mp, p.Mp, p.Mp, BM)) return cudaErrorUnknown;
  if (!smmd_host::make_tmap_bf16_2d(&twt, Wb, p.Mp, p.Mp, p.Mp, 64)) return cudaErrorUnknown;
  if (!smmd_host::make_tmap_bf16_2d(&tws, Wb, p.Mp, p.Mp, p.Mp, 32)) return cudaErrorUnknown;   // pass-1 stores
  // fixed-point scale of the row sums of W: |W| <= cmax * kd bound, |r_i| <= that * Mp < 2^61 after scaling
  const double cmax = 4.0 * wscale * std::max(std::max(fabs(c.a_xx), fabs(c.a_yy)), fabs(c.a_xy));
  const double wb = cmax * kd_abs_bound(kf);
  int ex = 0;
  frexp(wb * (double)p.Mp, &ex);           // wb * Mp < 2^ex
  const int shift = std::max(-100, std::min(100, 61 - ex));
  SymWgenArgs ga;
  ga.kf = kf;
  ga.geo = p.geo;
  ga.m = g.m;
  ga.n = g.n;
  ga.mp = p.mp;
  ga.np = p.np;
  ga.Mp = p.Mp;
  ga.c_xx = (float)(4.0 * c.a_xx * wscale);
  ga.c_yy = (float)(4.0 * c.a_yy * wscale);
  ga.c_xy = (float)(4.0 * c.a_xy * wscale);
  ga.norms = norms;
  ga.nkp = (int)(p.dp / 64);
  ga.rscale = (float)ldexp(1.0, shift);
  ga.rclamp = (float)(wb * (double)p.Mp);
  ga.racc = racc;
  ga.partials = partials;
  prof_begin(s);
  const int only = tuning().sym_only;   // developer timing knob: 1 = pass 1 only, 2 = pass 2 only (results meaningless)
  if (only != 2) {
    if ((e = launch_sym_wgen(variant, c.f16 != 0, tzi, tzj, tws, ga, p.grid1, s)) != cudaSuccess) return e;
    ++*launches;
  }
  prof_mark(s);
  SymWzArgs za;
  za.nmb = p.nmb;
  za.FB = p.FB;
  za.dp = (int)p.dp;
  za.KT = p.KT;
  za.S = p.S;
  za.ksteps = p.ksteps;
  za.Opart = reinterpret_cast<float*>(w + p.off_O);
  za.tonly = 0;
  za.ell = 1;
  za.Mp = (int)p.Mp;
  za.f16 = c.f16;
  if ((e = cudaFuncSetAttribute(tc_sym_wz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSzSmem)) != cudaSuccess)
    return e;
  if (only != 1) {
    tc_sym_wz_kernel<<<p.grid2, kThreads, kSzSmem, s>>>(twd, twt, tz, za);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*launches;
  }
  prof_end(s);
  SymFinArgs fr;
  fr.kf = kf;
  fr.m = g.m;
  fr.n = g.n;
  fr.mp = p.mp;
  fr.np = p.np;
  fr.d = g.d;
  fr.dp = (int)p.dp;
  fr.FB = p.FB;
  fr.S = p.S;
  fr.a_xx = c.a_xx;
  fr.a_yy = c.a_yy;
  fr.a_xy = c.a_xy;
  fr.rinv = ldexp(1.0, -shift);
  fr.src = src;
  fr.norms = norms;
  fr.csum = dot ? csum : nullptr;
  fr.Opart = za.Opart;
  fr.racc = racc;
  fr.dX = dX;
  fr.dY = dY;
  fr.partials = partials + (int64_t)p.grid1 * 6;
  fr.symf = 0;
  fr.ell = 1;
  fr.fgeo = SymfGeo{};
  fr.Odir = nullptr;
  fr.gscale = (float)(1.0 / wscale);
  sym_finalize_rows_kernel<<<(unsigned)p.fin_blocks, 256, 0, s>>>(fr);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  ++*launches;
  e = launch_finalize_partials(kf, g, partials, (int64_t)p.grid1 + p.fin_blocks, scalars, s);
  if (e != cudaSuccess) return e;
  ++*launches;
  return cudaSuccess;
}

}  // namespace tc
}  // namespace smmd

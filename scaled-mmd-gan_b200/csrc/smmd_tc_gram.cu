// smmd_tc_gram.cu -- Gram + reduction kernels: tc_stream_kernel (value-only MMD^2) and tc_macro_kernel (KID, 3-sample
// sums).  Overview of the tensor-core path: smmd_tc.cu.
#include "smmd_tc_common.cuh"

namespace smmd {
namespace tc {
namespace {

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
      mbar_wait_sleep(&acc_full[grp], (uint32_t)((tc >> 1) & 1), 64);
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
      mbar_wait_sleep(acc_full, fph, 200);
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

size_t tc_value_only_workspace_bytes(int64_t m, int64_t n, int64_t d, int precision) {
  return stream_plan(m, n, d, 1, precision == SMMD_PREC_BF16X3).off_end + 4096;
}

cudaError_t tc_run_value_only(const KernelFn& kf, TcVariant variant, const Geometry& g, const Coefs& c, const SrcLayout& src, int precision, double* scalars,
                              float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches, const char** path) {
  const void* X = src.X;
  const void* Y = src.Y;
  const int dtype = src.dtype;
  const int64_t ldx = src.ldx, ldy = src.ldy;
  char* w = static_cast<char*>(ws);
  cudaError_t e;
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
  if ((e = launch_prep_tc(pa, p.Mp, 1, s)) != cudaSuccess) return e;
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

}  // namespace tc

using namespace tc;

// Shapes the macro-tile kernel covers: the tile index of a subset is decoded with R <= 64 row blocks of 256 (stacked
// rows 2 * round_up(msub, 256) <= 16384, i.e. msub <= 8192) and the stacked operand matrix is addressed with 32-bit TMA
// row coordinates.  Anything else runs on the exact path (AUTO) or is refused before a launch (explicit bf16 / bf16x3).
bool tc_kid_supported(int64_t d, int64_t msub, int64_t nsub) {
  if (!(d >= 1 && d <= 65536) || msub < 1 || nsub < 1) return false;
  const int64_t Mp = 2 * round_up(msub, 256);
  return Mp / 256 <= 64 && nsub * Mp < ((int64_t)1 << 31);
}

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
  if ((e = launch_prep_tc(pa, p.Mp, (unsigned)nsub, s)) != cudaSuccess) return e;
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

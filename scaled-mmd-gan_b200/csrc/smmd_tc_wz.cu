// smmd_tc_wz.cu -- wide features (d > 256): W row panels (tc_wgen_kernel, pass 1) + O = W Z (tc_wz_kernel, pass 2) +
// wz_finalize_rows_kernel.  Overview of the tensor-core path: smmd_tc.cu.
#include "smmd_tc_common.cuh"

namespace smmd {
namespace tc {
namespace {

constexpr int kFinRowsPerWarp = 4;
constexpr int kFinRowsPerCta = 8 * kFinRowsPerWarp;

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
constexpr int kWgThreads = kRoleThreads;                   // 16 epilogue warps + producer + issuer + 2 idle (setmaxnreg works per warpgroup)

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

  // (nested on purpose: ptxas allocates registers per setmaxnreg region only when each region is a branch of its own)
  if (warp >= kWgEpiWarps) {
  reg_dec<kCtlRegs>();
  if (warp == kWgEpiWarps) {
    // ===================== TMA producer =====================
    uint32_t st = 0, ph = 0;
    for (int w = 0; w < a.nwin; ++w) {
    int rbl = rbl_first, t = t_first;
    for (int pos = wp0; pos < wp1; ++pos) {
      const int ct = w * a.Wc + t;
      const int32_t arow = rb_of(a.rbi0 + rbl_ld(rbl)) * BM, brow = ct < a.CT ? col0_of(ct) : 0;
      for (int p = 0; p < (ct < a.CT ? a.nkp : 0); ++p) {
        mbar_wait_sleep(&empty[st], ph ^ 1, 128);
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
    constexpr uint32_t idesc = make_idesc(PAIR ? 2 * BM : BM, BNW, operand_fmt<Math>(), false, false);
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
      else mbar_wait_sleep(&acc_empty[ab], aph ^ 1, 64);
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
  }
  } else {
    reg_inc<kEpiRegs>();
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
        mbar_wait_sleep(&acc_full[grp], (tc >> 1) & 1, 64);
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
  int f16;        // operands are IEEE half (fp16 tier) instead of bf16
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
    const uint32_t idesc = make_idesc(BM, (uint32_t)nf, a.f16 ? kFmtF16 : kFmtBF16, false, true);   // B = Z tile, MN-major
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
      mbar_wait_sleep(acc_full, 0, 1000);
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
  float gscale;              // 1 / (power-of-two scale W was carried with), see w_scale_for()
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
          gv[e] = (rs * zz[e] - oo[e]) * a.gscale;
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
cudaError_t launch_wgen(TcVariant v, bool f16, const CUtensorMap& tm, const CUtensorMap& tb, const WgenArgs& a, int grid,
                        int pair, cudaStream_t s) {
  if (f16) {
    switch (v) {
      case TV_RBF1: return launch_wgen_t<F16Of<MathRbf1>>(tm, tb, a, grid, pair, s);
      case TV_RBF_LADDER5: return launch_wgen_t<F16Of<MathRbfLadder<5>>>(tm, tb, a, grid, pair, s);
      case TV_RBF_GENERIC: return launch_wgen_t<F16Of<MathGeneric<FAM_RBF>>>(tm, tb, a, grid, pair, s);
      case TV_RQ3_DEFAULT: return launch_wgen_t<F16Of<MathRq3Default>>(tm, tb, a, grid, pair, s);
      case TV_RQ_GENERIC: return launch_wgen_t<F16Of<MathGeneric<FAM_RQ>>>(tm, tb, a, grid, pair, s);
      case TV_DISTANCE: return launch_wgen_t<F16Of<MathDistance>>(tm, tb, a, grid, pair, s);
      default: return cudaErrorInvalidValue;
    }
  }
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

}  // namespace

size_t tc_wz_workspace_bytes(const Geometry& g) { return wz_plan(g.m, g.n, g.d, g.x0, g.x1, g.y0, g.y1).off_end; }

cudaError_t tc_run_wz(const KernelFn& kf, TcVariant variant, const Geometry& g, const Coefs& c, const SrcLayout& src, double* scalars,
                      float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches, const char** path) {
  const void* X = src.X;
  const void* Y = src.Y;
  const int dtype = src.dtype;
  const int64_t ldx = src.ldx, ldy = src.ldy;
  char* w = static_cast<char*>(ws);
  cudaError_t e;
  *path = "tc_bf16_wz";
  const WzPlan p = wz_plan(g.m, g.n, g.d, g.x0, g.x1, g.y0, g.y1);
  if (p.off_end > ws_bytes) return cudaErrorInvalidValue;
  __nv_bfloat16* Z = reinterpret_cast<__nv_bfloat16*>(w + p.off_Z);
  float* norms = reinterpret_cast<float*>(w + p.off_norm);
  double* csum = reinterpret_cast<double*>(w + p.off_csum);
  __nv_bfloat16* Wb = reinterpret_cast<__nv_bfloat16*>(w + p.off_W);
  PrepTcArgs pa{X, Y, dtype, ldx, ldy, g.m, g.n, p.mp, p.np, g.d, p.dp, p.dp, nullptr, nullptr, 0,
                kf.tanh_features, 0, Z, norms, nullptr, kf, src.blk_x, src.blk_y};
  pa.f16 = c.f16;
  prep_set_peers(pa, src);
  const double wscale = w_scale_for(c, kf);
  if ((e = launch_prep_tc(pa, p.Mp, 1, s)) != cudaSuccess) return e;
  peer_after_pull(s);   // (peer calls: the caller's "rows pulled" event, see smmd_peer_set_pull_event)
  ++*launches;
  const bool dot = kf.family == FAM_RQ && kf.add_dot > 0.f;
  if (dot) {
    if ((e = launch_colsum_tc(Z, p.dp, p.dp, g.m, p.mp, g.n, csum, c.f16, s)) != cudaSuccess) return e;
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
    ga.c_xx = (float)(4.0 * c.a_xx * wscale);
    ga.c_yy = (float)(4.0 * c.a_yy * wscale);
    ga.c_xy = (float)(4.0 * c.a_xy * wscale);
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
    if ((e = launch_wgen(variant, c.f16 != 0, t1, t1b, ga, q.grid1, q.pair, s)) != cudaSuccess) return e;
    ++*launches;
    WzArgs za;
    za.nmb = q.nmb;
    za.FB = p.FB;
    za.dp = (int)p.dp;
    za.KT = p.KT;
    za.S = q.S;
    za.ksteps = q.ksteps;
    za.Opart = reinterpret_cast<float*>(w + p.off_O);
    za.f16 = c.f16;
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
    fr.gscale = (float)(1.0 / wscale);
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

}  // namespace tc
}  // namespace smmd

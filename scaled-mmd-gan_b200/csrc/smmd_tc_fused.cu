// smmd_tc_fused.cu -- fused MMD^2 forward + backward on tcgen05 (d <= 256): tc_fused_kernel + tc_finalize_rows_kernel.
// Overview of the tensor-core path: smmd_tc.cu.
#include "smmd_tc_common.cuh"

namespace smmd {
namespace tc {
namespace {

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

// KSPLIT = column slices per tile: each of the two epilogue groups has 4*KSPLIT warps (TMEM lane quarter x
// column slice).  Warp roles: warps [0, 8*KSPLIT) = epilogue, then the TMA producer, and LAST the UMMA issuer:
// the warp scheduler favours the highest warp id on a sub-partition, and the single issuing thread is on the
// critical path of the whole CTA (measured: as warp 1 it needed ~3100 cycles per tile, most of it waiting for
// issue slots behind the epilogue warps and polling mbarriers at ~100-150 cycles per poll).
//
// mbarriers (all phases tracked with running counters, no div/mod):
//   zj_full[nst]   TMA -> UMMA issuer                      (Zj tile landed)
//   zj_empty[nst]  UMMA #2 commit -> TMA producer
//   w_free[2][2]   UMMA #2 commit per 32-column half -> the epilogue group that owns the W buffer
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
  uint64_t* w_free = w_full + 2;        // [2][2]  UMMA #2 has read 32-column half h of group g's W buffer
  uint64_t* zi_full = w_free + 4;
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
    for (int i = 0; i < 4; ++i) mbar_init(&w_free[i], 1);
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
        mbar_wait_sleep(zi_empty, (unit & 1) ^ 1, 128);
        if (elect_one()) {
          mbar_arrive_expect_tx(zi_full, ZI_BYTES);
          for (int p = 0; p < NPANEL; ++p) tma_load_2d(sZi + p * (BM * 128), &tmap_zi, zi_full, p * 64, rb * BM);
        }
        __syncwarp();
        for (int t = t0; t < t0 + TU; ++t) {
          mbar_wait_sleep(&zj_empty[st], ph ^ 1, 128);
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
        // static order, UMMA #1 three tiles ahead of UMMA #2; 2 waits per tile
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
              for (int kk = 0; kk < BNF / 16; ++kk) {
                umma_ts2(tmem + TM_O, wad + kk * 8, blo + kk * (2048 >> 4), hi, idesc2, (b2 > 0 || kk > 0) ? 1u : 0u);
                if (kk & 1) umma_commit(&w_free[wb2 * 2 + (kk >> 1)]);   // this half of the W buffer may be overwritten
              }
              umma_commit(&zj_empty[st2]);     // frees the Zj stage (producer)
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
    uint32_t gk = 0;                        // tiles this group has processed (W buffer phase)
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
          uint64_t* wfree_g = w_free + grp * 2;
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
              // the group's previous tile shares this W buffer: wait, half by half, until its UMMA #2 has read it
              // (one commit per 32-column half: waiting for the whole UMMA #2 stalled every tile by ~800 cycles)
              if (gk && ((half * CH_PER + 2 * it) & 1) == 0) mbar_wait(&wfree_g[(half * CH_PER + 2 * it) >> 1], (gk - 1) & 1);
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
              if (gk && ((half * CH_PER + ch) & 1) == 0) mbar_wait(&wfree_g[(half * CH_PER + ch) >> 1], (gk - 1) & 1);
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
          ++gk;
        }
        // advance the ring state by one tile of the CTA's stream
        par ^= 1;
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

// ================================================================================================
// tile-pair variant of the fused kernel
// ================================================================================================
// UMMA #1 at N = 64 is bound by the shared-memory read of its A operand (Z_i: 4 KB per instruction for 32 tensor
// cycles, ~65 cycles each, 1040 per tile) and made the issuer the longest role.  Here one UMMA #1 covers TWO
// column tiles (N = 128, same A read, tensor-bound at 64 cycles): 1040 cycles per PAIR.  What makes room for the
// 128-column S buffers in tensor memory is writing W in place: a warp overwrites the first half of the 32 S
// columns it has just read with its 16 packed bf16x2 W columns (chunk 0 -> columns [0,8), chunk 1 -> [8,16) of its
// own slice, both already consumed), so there are no separate W buffers:
//   TMEM  O[dp <= 256] | pair buffer 0 [tile A: 64 | tile B: 64] | pair buffer 1 [128]
// and an S/W buffer lives from UMMA #1 of its pair until UMMA #2 of both tiles has completed (pb_free).
// The Z_j ring is panel-major ([panel][stage][64 rows x 128 B], 4 stages) so that the two tiles of a pair are 128
// contiguous rows per k-panel for UMMA #1, while UMMA #2 reads one tile MN-major with LBO = the panel stride.
// Group g of the epilogue takes tile g of every pair; work chunks are cut at even tile positions.
constexpr uint32_t TMP_S = 256;
constexpr int kPairStages = 4;
inline int fused_pair_smem(int npanel) { return 1024 + npanel * kZiRowBytes + kPairStages * npanel * kZjRowBytes + 512; }

template <class Math>
__global__ void __launch_bounds__(576, 1)   // 18 warps: 5 on two sub-partitions -> 16384 / (5 * 32) = 102 -> 96 registers is the cap (the setmaxnreg re-split of
                                            // the symmetric kernels measured 0.5% slower here: this epilogue fits 96 registers)
tc_fused_pair_kernel(const __grid_constant__ CUtensorMap tmap_zi, const __grid_constant__ CUtensorMap tmap_zj,
                     const __grid_constant__ FusedArgs a) {
  constexpr int KSPLIT = 2, NPART = 4, CH_PER = 2, EPI_WARPS = 16, NST = kPairStages;
  const int NPANEL = a.npanel, DP = a.dp;
  const int ZI_BYTES = NPANEL * kZiRowBytes;
  constexpr int PANEL_STRIDE = NST * kZjRowBytes;     // bytes between k-panels of the ring
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sZi = smem;
  uint8_t* sZj = smem + ZI_BYTES;                     // [NPANEL][NST][64 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sZj + NPANEL * PANEL_STRIDE);
  uint64_t* zj_full = bars;             // [NST]
  uint64_t* zj_empty = bars + NST;      // [NST]
  uint64_t* s_full = bars + 2 * NST;    // [2]  UMMA #1 of a pair -> both epilogue groups
  // w_full is indexed [pair buffer][group]: a fast group may finish pair P+1 before the issuer has consumed the slow
  // group's arrival for pair P (nothing orders them: W is written in place), and a barrier that completes two phases
  // between waits deadlocks a parity wait.  The same (buffer, group) barrier is next used by pair P+2, which cannot
  // start before UMMA #2 of pair P -- i.e. after the issuer has consumed it.
  uint64_t* w_full = s_full + 2;        // [2][2]  epilogue group g -> issuer (W of its tile is in place)
  uint64_t* pb_free = w_full + 4;       // [2]  UMMA #2 of both tiles of a pair done -> issuer may refill the buffer
  uint64_t* zi_full = pb_free + 2;
  uint64_t* zi_empty = zi_full + 1;
  uint64_t* o_full = zi_empty + 1;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);
  float* sParams = reinterpret_cast<float*>(tmem_slot + 4);

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
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int64_t pos0 = (int64_t)blockIdx.x * a.chunk;   // even (fused_plan)
  const int64_t pos1 = pos0 + a.chunk < a.total_tiles ? pos0 + a.chunk : a.total_tiles;
  auto rb_of = [&](int64_t rbi) -> int { return rbi < a.nrb_x ? a.rb_x0 + (int)rbi : a.rb_y0 + (int)(rbi - a.nrb_x); };

  if (warp == EPI_WARPS) {
    // ===================== TMA producer =====================
    uint32_t unit = 0, st = 0, ph = 0;
    int rbi = (int)(pos0 / a.T);
    int t0 = (int)(pos0 - (int64_t)rbi * a.T);
    for (int64_t left = pos1 - pos0; left > 0; ++rbi, t0 = 0, ++unit) {
      const int TU = (int)std::min<int64_t>(a.T - t0, left);
      const int rb = rb_of(rbi);
      mbar_wait_sleep(zi_empty, (unit & 1) ^ 1, 128);
      if (elect_one()) {
        mbar_arrive_expect_tx(zi_full, ZI_BYTES);
        for (int p = 0; p < NPANEL; ++p) tma_load_2d(sZi + p * (BM * 128), &tmap_zi, zi_full, p * 64, rb * BM);
      }
      __syncwarp();
      for (int t = t0; t < t0 + TU; ++t) {
        mbar_wait_sleep(&zj_empty[st], ph ^ 1, 128);
        if (elect_one()) {
          mbar_arrive_expect_tx(&zj_full[st], NPANEL * kZjRowBytes);
          uint8_t* dst = sZj + st * kZjRowBytes;
          for (int p = 0; p < NPANEL; ++p) tma_load_2d(dst + p * PANEL_STRIDE, &tmap_zj, &zj_full[st], p * 64, t * BNF);
        }
        __syncwarp();
        if (++st == (uint32_t)NST) {
          st = 0;
          ph ^= 1;
        }
      }
      left -= TU;
    }
  } else if (warp == EPI_WARPS + 1) {
    // ===================== UMMA issuer =====================
    constexpr uint32_t idesc1 = make_idesc(BM, 2 * BNF, operand_fmt<Math>(), false, false);
    const uint32_t idesc2 = make_idesc(BM, (uint32_t)DP, operand_fmt<Math>(), false, true);
    const uint32_t hi = desc_hi_sw128(1024);
    const uint32_t zi_lo = desc_lo(smem_u32(sZi), 16);
    const uint32_t zj_lo1 = desc_lo(smem_u32(sZj), 16);                 // K-major B of UMMA #1 (two stages = 128 rows)
    const uint32_t zj_lo2 = desc_lo(smem_u32(sZj), PANEL_STRIDE);       // MN-major B of UMMA #2: LBO = panel stride
    constexpr uint32_t stage_step = (uint32_t)kZjRowBytes >> 4, panel_step = (uint32_t)PANEL_STRIDE >> 4;
    uint32_t unit = 0;
    uint32_t st1 = 0, ph1 = 0, gp1 = 0;        // UMMA #1 stream: first stage of the pair / phase, global pair counter
    uint32_t st2 = 0, gp2 = 0;                 // UMMA #2 stream: stage, global pair counter
    int rbi = (int)(pos0 / a.T);
    int t0 = (int)(pos0 - (int64_t)rbi * a.T);
    for (int64_t left = pos1 - pos0; left > 0; ++rbi, t0 = 0, ++unit) {
      const int TU = (int)std::min<int64_t>(a.T - t0, left);
      const int NP = TU >> 1;
      mbar_wait(zi_full, unit & 1);
      // static order: UMMA #1 runs two pairs ahead of the UMMA #2 of a pair
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
            const uint32_t wad = tmem + TMP_S + pb * 128 + g * 64;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < BNF / 16; ++kk)   // W slice kk sits at columns (kk/2)*32 + (kk%2)*8 of the tile
                umma_ts2(tmem + TM_O, wad + (kk >> 1) * 32 + (kk & 1) * 8, blo + kk * (2048 >> 4), hi, idesc2,
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
        if (j < NP) {   // ---- UMMA #1 for pair j: S[128 x 128] = Zi * [Zj(2j); Zj(2j+1)]^T
          const uint32_t pb = gp1 & 1;
          mbar_wait(&zj_full[st1], ph1);
          mbar_wait(&zj_full[st1 + 1], ph1);
          if (gp1 >= 2) mbar_wait(&pb_free[pb], ((gp1 >> 1) - 1) & 1);
          tc_fence_after();
          const uint32_t blo = zj_lo1 + st1 * stage_step;
          const uint32_t sad = tmem + TMP_S + pb * 128;
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
      left -= TU;
    }
  } else {
    // ===================== epilogue groups: group g takes tile g of every pair =====================
    const int grp = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int part = grp * KSPLIT + half;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const Math math(a.kf, sParams);
    const float kscale = math.k_scale(), kdscale = math.kd_scale();
    uint32_t unit = 0, gp = 0;   // global pair counter
    int slot = 0;
    int rbi = (int)(pos0 / a.T);
    int t0 = (int)(pos0 - (int64_t)rbi * a.T);
    const int mp = (int)a.mp, mvalid = (int)a.m, yvalid = (int)(a.mp + a.n);
    for (int64_t left = pos1 - pos0; left > 0; ++rbi, t0 = 0, ++unit, ++slot) {
      const int TU = (int)std::min<int64_t>(a.T - t0, left);
      const int rb = rb_of(rbi);
      const int gi = rb * BM + r;
      const bool rowX = gi < mp;
      const float ni = a.norms[gi];
      float2 rsum = make_float2(0.f, 0.f);
      double dsame = 0.0, dcross = 0.0;
      for (int lp = 0; lp < (TU >> 1); ++lp, ++gp) {
        const uint32_t pb = gp & 1;
        const int c0 = (t0 + 2 * lp + grp) * BNF;
        const bool colX = c0 < mp;
        const bool same = (colX == rowX);
        const float2 cw = bc2((same ? (rowX ? a.c_xx : a.c_yy) : a.c_xy) * kdscale);
        const int lim = colX ? mvalid : yvalid;
        const bool special = (c0 + BNF > lim) || ((c0 >> 7) == rb);
        const float* nj = a.norms + c0 + half * (CH_PER * 16);
        // this warp's 32 S columns; its 16 packed W columns go over the first half of them
        const uint32_t s_addr = tmem + TMP_S + pb * 128 + grp * 64 + half * 32 + lane_base;
        float2 tsum = make_float2(0.f, 0.f);
        mbar_wait(&s_full[pb], (gp >> 1) & 1);
        tc_fence_after();
        if (!special) {
          uint32_t va[16], vb[16], wpk[8];
          tmem_ld_x16(s_addr, va);
          tmem_ld_wait();
          tmem_ld_x16(s_addr + 16, vb);
          fused_chunk16<Math, false>(math, va, nj, ni, cw, 0, 0, 0, tsum, rsum, wpk);
          tmem_st_x8(s_addr, wpk);            // columns [0, 8): read with chunk 0
          tmem_ld_wait();
          fused_chunk16<Math, false>(math, vb, nj + 16, ni, cw, 0, 0, 0, tsum, rsum, wpk);
          tmem_st_x8(s_addr + 8, wpk);        // columns [8, 16): read with chunk 0
        } else {
#pragma unroll 1
          for (int ch = 0; ch < CH_PER; ++ch) {
            uint32_t v[16], wpk[8];
            tmem_ld_x16(s_addr + ch * 16, v);
            tmem_ld_wait();
            fused_chunk16<Math, true>(math, v, nj + ch * 16, ni, cw, c0 + half * (CH_PER * 16) + ch * 16, lim, gi, tsum,
                                      rsum, wpk);
            tmem_st_x8(s_addr + ch * 8, wpk);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&w_full[pb * 2 + grp]);
        if (same) dsame += (double)((tsum.x + tsum.y) * kscale);
        else dcross += (double)((tsum.x + tsum.y) * kscale);
      }
      // ---- unit end: drain O (this thread's slice of the feature columns) ----
      mbar_wait(o_full, unit & 1);
      tc_fence_after();
      {
        const int64_t sl = (int64_t)blockIdx.x * a.slots + slot;
        const int seg = DP / NPART;
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
  float gscale;      // 1 / (power-of-two scale W was carried with): 1 for bf16 operands, see w_scale_for()
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
            gv[e] = (rs * zz[e] - oo[e]) * a.gscale;         // W already carries the factor 4 a_ij
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
          float gv = (rs * z - oacc[t]) * a.gscale;
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

struct FusedPlan {
  int64_t mp, np, Mp, dp;
  int nrb_x, rb_x0, nrb_y, rb_y0, T, grid, slots;
  int64_t total, chunk;
  size_t off_Z, off_norm, off_csum, off_O, off_r, off_s, off_stats, off_end;
};

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
  p.chunk += p.chunk & 1;   // chunks start at even tile positions (the tile-pair kernel works on pairs; T is even)
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

template <class Math>
cudaError_t launch_fused_pair_t(const CUtensorMap& tzi, const CUtensorMap& tzj, const FusedArgs& a, int grid, cudaStream_t s) {
  const int smem = fused_pair_smem(a.npanel);
  if (smem > kMaxSmem) return cudaErrorInvalidConfiguration;
  auto kern = tc_fused_pair_kernel<Math>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, 576, smem, s>>>(tzi, tzj, a);
  return cudaGetLastError();
}

cudaError_t launch_fused_pair(TcVariant v, bool f16, const CUtensorMap& tzi, const CUtensorMap& tzj, const FusedArgs& a, int grid,
                              cudaStream_t s) {
  if (f16) {
    switch (v) {
      case TV_RBF1: return launch_fused_pair_t<F16Of<MathRbf1>>(tzi, tzj, a, grid, s);
      case TV_RBF_LADDER5: return launch_fused_pair_t<F16Of<MathRbfLadder<5>>>(tzi, tzj, a, grid, s);
      case TV_RBF_GENERIC: return launch_fused_pair_t<F16Of<MathGeneric<FAM_RBF>>>(tzi, tzj, a, grid, s);
      case TV_RQ3_DEFAULT: return launch_fused_pair_t<F16Of<MathRq3Default>>(tzi, tzj, a, grid, s);
      case TV_RQ_GENERIC: return launch_fused_pair_t<F16Of<MathGeneric<FAM_RQ>>>(tzi, tzj, a, grid, s);
      case TV_DISTANCE: return launch_fused_pair_t<F16Of<MathDistance>>(tzi, tzj, a, grid, s);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (v) {
    case TV_RBF1: return launch_fused_pair_t<MathRbf1>(tzi, tzj, a, grid, s);
    case TV_RBF_LADDER5: return launch_fused_pair_t<MathRbfLadder<5>>(tzi, tzj, a, grid, s);
    case TV_RBF_GENERIC: return launch_fused_pair_t<MathGeneric<FAM_RBF>>(tzi, tzj, a, grid, s);
    case TV_RQ3_DEFAULT: return launch_fused_pair_t<MathRq3Default>(tzi, tzj, a, grid, s);
    case TV_RQ_GENERIC: return launch_fused_pair_t<MathGeneric<FAM_RQ>>(tzi, tzj, a, grid, s);
    case TV_DISTANCE: return launch_fused_pair_t<MathDistance>(tzi, tzj, a, grid, s);
    case TV_NULL: return launch_fused_pair_t<MathNull>(tzi, tzj, a, grid, s);
    default: return cudaErrorInvalidValue;
  }
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

}  // namespace

#ifdef SMMD_PIPE_TIMING
void pipe_timing_dump_impl(bool reset) {
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

size_t tc_fused_workspace_bytes(const Geometry& g) { return fused_plan(g.m, g.n, g.d, g.x0, g.x1, g.y0, g.y1).off_end; }

cudaError_t tc_run_fused(const KernelFn& kf, TcVariant variant, const Geometry& g, const Coefs& c, const SrcLayout& src, double* scalars,
                         float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches, const char** path) {
  const void* X = src.X;
  const void* Y = src.Y;
  const int dtype = src.dtype;
  const int64_t ldx = src.ldx, ldy = src.ldy;
  char* w = static_cast<char*>(ws);
  cudaError_t e;
  const FusedPlan p = fused_plan(g.m, g.n, g.d, g.x0, g.x1, g.y0, g.y1);
  if (p.off_end > ws_bytes) return cudaErrorInvalidValue;
  *path = "tc_bf16_fused";
  __nv_bfloat16* Z = reinterpret_cast<__nv_bfloat16*>(w + p.off_Z);
  float* norms = reinterpret_cast<float*>(w + p.off_norm);
  double* csum = reinterpret_cast<double*>(w + p.off_csum);
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
  CUtensorMap tzi, tzj;
  if (!smmd_host::make_tmap_bf16_2d(&tzi, Z, p.Mp, p.dp, p.dp, BM)) return cudaErrorUnknown;
  if (!smmd_host::make_tmap_bf16_2d(&tzj, Z, p.Mp, p.dp, p.dp, BNF)) return cudaErrorUnknown;
  FusedArgs fa;
  fa.kf = kf;
  fa.m = g.m;
  fa.n = g.n;
  fa.mp = p.mp;
  fa.np = p.np;
  fa.c_xx = (float)(4.0 * c.a_xx * wscale);
  fa.c_yy = (float)(4.0 * c.a_yy * wscale);
  fa.c_xy = (float)(4.0 * c.a_xy * wscale);
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
  // (the fp16 operand tier exists for the tile-pair kernel only)
  e = (tuning().fused_pair || c.f16) ? launch_fused_pair(variant, c.f16 != 0, tzi, tzj, fa, p.grid, s)
                                     : launch_fused(variant, tzi, tzj, fa, p.grid, s);
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
  fr.gscale = (float)(1.0 / wscale);
  const unsigned fin_blocks = (unsigned)((fr.ox + fr.oy + kFinRowsPerCta - 1) / kFinRowsPerCta);
  tc_finalize_rows_kernel<<<fin_blocks, 256, 0, s>>>(fr);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  ++*launches;
  e = launch_finalize_partials(kf, g, fr.partials, fin_blocks, scalars, s);
  if (e != cudaSuccess) return e;
  ++*launches;
  return cudaSuccess;
}

}  // namespace tc

#ifdef SMMD_PIPE_TIMING
void pipe_timing_dump(bool reset) { tc::pipe_timing_dump_impl(reset); }
#endif

}  // namespace smmd

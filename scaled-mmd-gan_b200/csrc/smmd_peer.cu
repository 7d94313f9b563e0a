// smmd_peer.cu -- the exchange steps of the sharded loss as kernels over peer-mapped memory (NVLink / NVSwitch):
//   peer_publish_kernel + peer_signal_kernel   local rows -> own exchange slot (bf16 / fp32), data flags raised in every peer
//   (prep_tc_kernel, smmd_tc.cu)               pulls each peer's rows once its flag is up: the all_gather, fused into the prep
//   peer_combine_kernel                        partial sums -> every peer's slot, flags, wait, add in rank order, MMD^2
//   peer_small_mmd2_kernel<DMAX>               latency-bound shapes: publish + pull + loss + gradients + sum exchange +
//                                              combine in ONE launch (global batch <= 1024 rows, d <= 64, exact fp32 math)
// Layout and protocol: smmd_peer.cuh.  New capability (the reference's towers never exchange features,
// gan/core/model.py:186-218); the arithmetic is that of small_mmd2_kernel / finalize (smmd_simt.cu).
#include <algorithm>
#include <cuda_bf16.h>
#include "smmd_kfun.cuh"
#include "smmd_internal.h"

namespace smmd {

namespace {

constexpr int kPeerRowsPerCta = 8;

struct PeerBases {
  void* base[kPeerMax];
};

// ---- big problems: publish the local rows ------------------------------------------------------------------
// rows [0, blk_x) = X_local, [blk_x, blk_x + blk_y) = Y_local, pitch d; bf16 (the tensor-core operand format: half
// the NVLink bytes, and exactly the values every rank will multiply) or fp32 (fp16 operand tier).
__global__ void __launch_bounds__(256) peer_publish_kernel(const float* X, const float* Y, int64_t ld, int64_t blk_x,
                                                           int64_t blk_y, int64_t d, int to_bf16, void* own,
                                                           size_t slot_bytes, unsigned long long step_arg) {
  const int64_t rows = blk_x + blk_y;
  void* slot = peer_data_slot(own, peer_step_of(step_arg, own), slot_bytes);
  const bool vec = (d % 8 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
  if (vec) {
    const int64_t per_row = d / 8, total = rows * per_row;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = e / per_row, c = (e - r * per_row) * 8;
      const float* src = (r < blk_x ? X + r * ld : Y + (r - blk_x) * ld) + c;
      const float4 lo = *reinterpret_cast<const float4*>(src), hi = *reinterpret_cast<const float4*>(src + 4);
      if (to_bf16) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(lo.x, lo.y), h1 = __floats2bfloat162_rn(lo.z, lo.w);
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(hi.x, hi.y), h3 = __floats2bfloat162_rn(hi.z, hi.w);
        *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(slot) + r * d + c) =
            make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                       *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
      } else {
        float* dst = static_cast<float*>(slot) + r * d + c;
        *reinterpret_cast<float4*>(dst) = lo;
        *reinterpret_cast<float4*>(dst + 4) = hi;
      }
    }
  } else {
    const int64_t total = rows * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = e / d, c = e - r * d;
      const float v = r < blk_x ? X[r * ld + c] : Y[(r - blk_x) * ld + c];
      if (to_bf16) static_cast<__nv_bfloat16*>(slot)[e] = __float2bfloat16_rn(v);
      else static_cast<float*>(slot)[e] = v;
    }
  }
}

// The publish kernel has completed (stream order): its rows sit in this GPU's memory, which is where a peer's NVLink read
// looks.  Raise data_flag[self] = step in every peer's buffer.
__global__ void peer_signal_kernel(PeerBases pb, int world, int self, size_t flag_off, unsigned long long step_arg) {
  const unsigned long long step = peer_step_of(step_arg, pb.base[self]);
  __threadfence_system();
  const int t = threadIdx.x;
  if (t < world && t != self) st_release_sys(peer_flag(pb.base[t], flag_off, step, self), step);
}

// ---- partial sums: exchange + combine (the all_reduce of 7 doubles and smmd_mmd2_combine in one kernel) -------
// One warp.  Lane t < world writes this rank's 16 scalars into rank t's slot [step & 1][self] and raises sums_flag[self]
// there; then lane t waits for rank t's flag in the own buffer; lane 0 adds the slots in rank order (the same order on
// every rank: identical bits everywhere) and forms MMD^2.
__device__ __forceinline__ void peer_exchange_and_combine(const KernelFn& kf, double m, double n, int biased,
                                                          const double (&mine)[SMMD_NUM_SCALARS], double* scalars,
                                                          const PeerBases& pb, int world, int self,
                                                          unsigned long long step, int lane) {
  const size_t slot = ((size_t)(step & 1) * kPeerMax) * SMMD_NUM_SCALARS * sizeof(double);
  if (lane < world) {
    char* dst_base = static_cast<char*>(pb.base[lane]);
    double* dst = reinterpret_cast<double*>(dst_base + kPeerOffSums + slot) + (size_t)self * SMMD_NUM_SCALARS;
#pragma unroll
    for (int i = 0; i < SMMD_NUM_SCALARS; ++i) st_relaxed_sys_f64(dst + i, mine[i]);
    __threadfence_system();
    st_release_sys(peer_flag(pb.base[lane], kPeerOffSumsFlag, step, self), step);
  }
  char* own = static_cast<char*>(pb.base[self]);
  if (lane < world) peer_wait_flag(peer_flag(pb.base[self], kPeerOffSumsFlag, step, lane), step);
  __syncwarp();
  if (lane == 0) {
    const double* in = reinterpret_cast<const double*>(own + kPeerOffSums + slot);
    double t[SMMD_NUM_SCALARS];
    for (int i = 0; i < SMMD_NUM_SCALARS; ++i) t[i] = 0.0;
    for (int r = 0; r < world; ++r)
      for (int i = SMMD_S_SUM_XX; i <= SMMD_S_NONFINITE; ++i) t[i] += ld_relaxed_sys_f64(in + (size_t)r * SMMD_NUM_SCALARS + i);
    t[SMMD_S_MMD2] = mmd2_from_sums(kf, m, n, biased, t[SMMD_S_SUM_XX], t[SMMD_S_SUM_YY], t[SMMD_S_SUM_XY],
                                    t[SMMD_S_SUM_YX], t[SMMD_S_DIAG_X], t[SMMD_S_DIAG_Y]);
    t[SMMD_S_NONFINITE] = t[SMMD_S_NONFINITE] != 0.0 ? 1.0 : 0.0;
    for (int i = 0; i < SMMD_NUM_SCALARS; ++i) scalars[i] = t[i];
    // this rank has completed `step` (last action of the call: the next call with step = 0 reads counter + 1)
    *reinterpret_cast<volatile unsigned long long*>(own + kPeerOffStep) = step;
  }
}

__global__ void __launch_bounds__(32) peer_combine_kernel(KernelFn kf, int64_t m, int64_t n, int biased, double* scalars,
                                                          PeerBases pb, int world, int self, unsigned long long step_arg) {
  const unsigned long long step = peer_step_of(step_arg, pb.base[self]);
  double mine[SMMD_NUM_SCALARS];
#pragma unroll
  for (int i = 0; i < SMMD_NUM_SCALARS; ++i) mine[i] = scalars[i];
  __syncwarp();
  peer_exchange_and_combine(kf, (double)m, (double)n, biased, mine, scalars, pb, world, self, step, (int)threadIdx.x);
}

// ---- latency-bound shapes: everything in one launch ----------------------------------------------------------
struct PeerSmallArgs {
  KernelFn kf;
  const float* X;            // local rows
  const float* Y;
  int64_t ld;
  int m, n, d;               // GLOBAL sizes
  int blk_x, blk_y;          // rows per rank
  int pitch;                 // shared-memory row pitch in floats (odd: conflict-free column walks)
  float a_xx, a_yy, a_xy;
  int diag_in_sum, biased;
  float* dX;                 // local gradients (nullable)
  float* dY;
  double* partials;          // [gridDim.x][6]
  unsigned int* counters;    // [2], zeroed in front of the launch
  double* scalars;
  PeerBases pb;
  int world, self;
  unsigned long long step;   // 0: read on the device (own counter + 1)
  size_t slot_bytes;
};

template <int DMAX>
__global__ void __launch_bounds__(256) peer_small_mmd2_kernel(PeerSmallArgs a) {
  extern __shared__ float sm[];
  const int M = a.m + a.n, rows_local = a.blk_x + a.blk_y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* Zs = sm;                                           // [M][pitch]: the global batch, X rows then Y rows
  float* accS = sm + (size_t)M * a.pitch + (warp * 32 + lane) * (DMAX + 1);
  __shared__ double red[kPeerRowsPerCta][6];
  __shared__ int is_last;
  char* own = static_cast<char*>(a.pb.base[a.self]);
  const unsigned long long step = peer_step_of(a.step, own);
  const size_t slot_off = kPeerOffData + (size_t)(step & 1) * a.slot_bytes;

  // 1. publish: this CTA's share of the local rows -> own slot (fp32, pitch d); the last CTA to finish raises the flags
  if (a.world > 1) {
    float* slot = reinterpret_cast<float*>(own + slot_off);
    const int total = rows_local * a.d;
    for (int e = blockIdx.x * 256 + tid; e < total; e += gridDim.x * 256) {
      const int r = e / a.d, c = e - r * a.d;
      slot[e] = r < a.blk_x ? a.X[(int64_t)r * a.ld + c] : a.Y[(int64_t)(r - a.blk_x) * a.ld + c];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(a.counters, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last) {
      __threadfence_system();
      if (tid < a.world && tid != a.self)
        st_release_sys(peer_flag(a.pb.base[tid], kPeerOffDataFlag, step, a.self), step);
    }
    // 2. every peer's rows are published
    if (tid < a.world && tid != a.self)
      peer_wait_flag(peer_flag(own, kPeerOffDataFlag, step, tid), step);
    __syncthreads();
  }
  // 3. pull the global batch into shared memory (independent loads: one NVLink round trip, not one per column)
  for (int e = tid; e < M * a.d; e += 256) {
    const int j = e / a.d, c = e - j * a.d;
    const bool inX = j < a.m;
    const int idx = inX ? j : j - a.m;
    const int blk = inX ? a.blk_x : a.blk_y;
    const int r = idx / blk, loc = idx - r * blk;
    float v;
    if (r == a.self) v = inX ? a.X[(int64_t)loc * a.ld + c] : a.Y[(int64_t)loc * a.ld + c];
    else v = __ldcg(reinterpret_cast<const float*>(static_cast<const char*>(a.pb.base[r]) + slot_off) +
                    (size_t)((inX ? 0 : a.blk_x) + loc) * a.d + c);
    Zs[j * a.pitch + c] = v;
  }
  __syncthreads();
  // 4. this rank's rows against all columns (the arithmetic of small_mmd2_kernel)
  const int il = blockIdx.x * kPeerRowsPerCta + warp;   // local row: X_local rows then Y_local rows
  const bool row_valid = il < rows_local;
  const bool rowX = il < a.blk_x;
  const int ig = rowX ? a.self * a.blk_x + il : a.m + a.self * a.blk_y + (il - a.blk_x);   // stacked global index
  double q[6] = {0, 0, 0, 0, 0, 0};
  if (row_valid) {
    const float* zi_p = Zs + ig * a.pitch;
    float zi[DMAX], acc[DMAX];
    float ni = 0.f;
#pragma unroll
    for (int c = 0; c < DMAX; ++c) {
      zi[c] = c < a.d ? zi_p[c] : 0.f;
      ni = fmaf(zi[c], zi[c], ni);
      acc[c] = 0.f;
    }
    double s_same = 0.0, s_cross = 0.0;
    for (int j = lane; j < M; j += 32) {
      const bool colX = j < a.m;
      const float* zj_p = Zs + j * a.pitch;
      float zj[DMAX];
      float S = 0.f, Dd = 0.f, nj = 0.f;
#pragma unroll
      for (int c = 0; c < DMAX; ++c) {
        zj[c] = c < a.d ? zj_p[c] : 0.f;
        const float e = zi[c] - zj[c];
        S = fmaf(zi[c], zj[c], S);
        Dd = fmaf(e, e, Dd);
        nj = fmaf(zj[c], zj[c], nj);
      }
      const PairVal pv = eval_exact(a.kf, S, Dd, ni, nj);
      const bool same = (colX == rowX);
      const float aco = same ? (rowX ? a.a_xx : a.a_yy) : a.a_xy;
      float wd = 0.f, wg = 0.f;
      if (j == ig) {
        wg = a.diag_in_sum ? 2.f * aco * pv.kg : 0.f;
      } else {
        wd = 4.f * aco * pv.kd;
        wg = 2.f * aco * pv.kg;
        if (same) s_same += (double)pv.k;
        else s_cross += (double)pv.k;
      }
      if (a.dX) {
#pragma unroll
        for (int c = 0; c < DMAX; ++c) acc[c] = fmaf(wd, zi[c] - zj[c], fmaf(wg, zj[c], acc[c]));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_same += __shfl_xor_sync(0xffffffffu, s_same, o);
      s_cross += __shfl_xor_sync(0xffffffffu, s_cross, o);
    }
    const double dg = a.kf.family == FAM_RQ ? (double)a.kf.const_diag + (double)a.kf.add_dot * (double)ni
                                            : (double)diag_value(a.kf, ni);
    q[rowX ? 0 : 1] = s_same;
    q[rowX ? 2 : 3] = s_cross;
    q[rowX ? 4 : 5] = dg;
    if (a.dX) {
#pragma unroll
      for (int c = 0; c < DMAX; ++c) accS[c] = acc[c];
      __syncwarp();
      const float* wbase = sm + (size_t)M * a.pitch + warp * 32 * (DMAX + 1);
      float* out = rowX ? a.dX + (int64_t)il * a.d : a.dY + (int64_t)(il - a.blk_x) * a.d;
      for (int c = lane; c < a.d; c += 32) {
        float t = 0.f;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) t += wbase[l * (DMAX + 1) + c];
        out[c] = t;
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) red[warp][i] = q[i];
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) {
      double t = 0.0;
      for (int w = 0; w < kPeerRowsPerCta; ++w) t += red[w][i];
      a.partials[(int64_t)blockIdx.x * 6 + i] = t;
    }
    __threadfence();
    is_last = atomicAdd(a.counters + 1, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // 5. this rank's sums -> every peer, wait for theirs, combine (first warp of the last CTA)
  if (tid < 6) {
    double t = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(a.partials + (int64_t)b * 6 + tid);
    red[0][tid] = t;
  }
  __syncthreads();
  if (warp == 0) {
    double mine[SMMD_NUM_SCALARS];
#pragma unroll
    for (int i = 0; i < SMMD_NUM_SCALARS; ++i) mine[i] = 0.0;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      mine[SMMD_S_SUM_XX + i] = red[0][i];
      bad = bad || !isfinite(red[0][i]);
    }
    mine[SMMD_S_NONFINITE] = bad ? 1.0 : 0.0;
    peer_exchange_and_combine(a.kf, (double)a.m, (double)a.n, a.biased, mine, a.scalars, a.pb, a.world, a.self, step, lane);
  }
}

PeerBases bases_of(const smmd_peer_table& pt) {
  PeerBases pb;
  for (int i = 0; i < kPeerMax; ++i) pb.base[i] = i < pt.world ? pt.base[i] : nullptr;
  return pb;
}

}  // namespace

cudaError_t launch_peer_publish(const float* X, const float* Y, int64_t ld, int64_t blk_x, int64_t blk_y, int64_t d,
                                int to_bf16, const smmd_peer_table& pt, uint64_t step, cudaStream_t s) {
  const int64_t work = (blk_x + blk_y) * d / 8 + 1;
  const unsigned grid = (unsigned)std::min<int64_t>((work + 255) / 256, 148 * 8);
  peer_publish_kernel<<<grid, 256, 0, s>>>(X, Y, ld, blk_x, blk_y, d, to_bf16, pt.base[pt.rank],
                                          peer_slot_bytes(blk_x + blk_y, d), (unsigned long long)step);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (pt.world > 1) {
    peer_signal_kernel<<<1, 32, 0, s>>>(bases_of(pt), pt.world, pt.rank, kPeerOffDataFlag, (unsigned long long)step);
    e = cudaGetLastError();
  }
  return e;
}

PeerSrc make_peer_src(const smmd_peer_table& pt, int64_t rows_local, int64_t d, uint64_t step) {
  PeerSrc ps;
  ps.on = 1;
  ps.world = pt.world;
  ps.self = pt.rank;
  for (int i = 0; i < kPeerMax; ++i) ps.base[i] = i < pt.world ? pt.base[i] : nullptr;
  ps.slot_bytes = peer_slot_bytes(rows_local, d);
  ps.step = (unsigned long long)step;
  return ps;
}

cudaError_t launch_peer_combine(const KernelFn& kf, const Geometry& g, double* scalars, const smmd_peer_table& pt,
                                uint64_t step, cudaStream_t s) {
  peer_combine_kernel<<<1, 32, 0, s>>>(kf, g.m, g.n, g.biased, scalars, bases_of(pt), pt.world, pt.rank,
                                       (unsigned long long)step);
  return cudaGetLastError();
}

// one-launch path: the whole global batch must fit shared memory next to the per-lane gradient rows
static int peer_small_dmax(int64_t d) { return d <= 4 ? 4 : d <= 16 ? 16 : d <= 32 ? 32 : 64; }
static size_t peer_small_smem(int64_t M, int64_t d) {
  const int64_t pitch = d | 1;
  return (size_t)(M * pitch + (int64_t)kPeerRowsPerCta * 32 * (peer_small_dmax(d) + 1)) * sizeof(float);
}
bool peer_small_eligible(const KernelFn& kf, const Geometry& g) {
  return !kf.tanh_features && kf.family != FAM_POLY && g.d <= 64 && g.m + g.n <= 1024 &&
         peer_small_smem(g.m + g.n, g.d) <= 200 * 1024;
}

cudaError_t launch_peer_small_mmd2(const KernelFn& kf, const Geometry& g, const Coefs& c, const float* X, const float* Y,
                                   int64_t ld, float* dX, float* dY, double* partials, unsigned int* counters,
                                   double* scalars, const smmd_peer_table& pt, uint64_t step, cudaStream_t s) {
  PeerSmallArgs a;
  a.kf = kf;
  a.X = X;
  a.Y = Y;
  a.ld = ld;
  a.m = (int)g.m;
  a.n = (int)g.n;
  a.d = (int)g.d;
  a.blk_x = (int)(g.m / pt.world);
  a.blk_y = (int)(g.n / pt.world);
  a.pitch = (int)(g.d | 1);
  a.a_xx = (float)c.a_xx;
  a.a_yy = (float)c.a_yy;
  a.a_xy = (float)c.a_xy;
  a.diag_in_sum = c.diag_in_sum;
  a.biased = g.biased;
  a.dX = dX;
  a.dY = dY;
  a.partials = partials;
  a.counters = counters;
  a.scalars = scalars;
  a.pb = bases_of(pt);
  a.world = pt.world;
  a.self = pt.rank;
  a.step = (unsigned long long)step;
  a.slot_bytes = peer_slot_bytes(a.blk_x + a.blk_y, g.d);
  cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), s);
  if (e != cudaSuccess) return e;
  const int rows_local = a.blk_x + a.blk_y;
  const unsigned grid = (unsigned)((rows_local + kPeerRowsPerCta - 1) / kPeerRowsPerCta);
  const size_t smem = peer_small_smem(g.m + g.n, g.d);
  auto go = [&](auto kern) -> cudaError_t {
    if (smem > 48 * 1024) {
      cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e2 != cudaSuccess) return e2;
    }
    kern<<<grid, 256, smem, s>>>(a);
    return cudaGetLastError();
  };
  switch (peer_small_dmax(g.d)) {
    case 4: return go(peer_small_mmd2_kernel<4>);
    case 16: return go(peer_small_mmd2_kernel<16>);
    case 32: return go(peer_small_mmd2_kernel<32>);
    default: return go(peer_small_mmd2_kernel<64>);
  }
}

}  // namespace smmd

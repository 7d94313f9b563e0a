// umma2_probe.cu -- hardware probe for cta_group::2 UMMA (one MMA spans a CTA pair):
//   D[256 x 256] = A[256 x 64] * B[256 x 64]^T, bf16 in, fp32 out.  CTA r of the pair holds rows [128 r, +128) of A
//   and of B in its own shared memory (K-major SW128 panels); the leader issues 4 UMMAs M=256 N=256 K=16; D rows
//   [128 r, +128) land in CTA r's tensor memory.  Exit code 0 = matches the CPU.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include "../sm100_ptx.cuh"
#include "../tmap_host.h"

using namespace sm100;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2_kernel(const __grid_constant__ CUtensorMap tmap, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 128 rows x 128 B
  uint8_t* sB = smem + 128 * 128;     // 128 rows x 128 B (this CTA's half of the N = 256 B rows)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * 128 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(&bars[0], 1);   // operands landed (leader's barrier collects both CTAs' bytes)
    mbar_init(&bars[1], 1);   // accumulator ready (commit multicast to both CTAs)
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_pair<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    const uint32_t leader_bar = map_to_cta(smem_u32(&bars[0]), 0);
    if (rank == 0) mbar_arrive_expect_tx(&bars[0], 4 * 128 * 128);
    tma_load_2d_pair(sA, &tmap, leader_bar, 0, (int)rank * 128);
    tma_load_2d_pair(sB, &tmap, leader_bar, 0, 256 + (int)rank * 128);
    if (rank == 0) {
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc(256, 256, kFmtBF16, false, false);
      const uint32_t hi = desc_hi_sw128(1024);
      const uint32_t a_lo = desc_lo(smem_u32(sA), 16), b_lo = desc_lo(smem_u32(sB), 16);
      for (int k = 0; k < 4; ++k) umma_ss2_pair(tmem, a_lo + k * 2, b_lo + k * 2, hi, idesc, k ? 1u : 0u);
      umma_commit_pair(&bars[1], 3);
    }
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  // epilogue: thread = TMEM lane = row (rank * 128 + tid)
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < 256; c += 16) {
    uint32_t v[16];
    tmem_ld_x16(tmem + c + lane_base, v);
    tmem_ld_wait();
    for (int e = 0; e < 16; ++e) out[(size_t)(rank * 128 + tid) * 256 + c + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) tmem_dealloc_pair<256>(tmem);
}

int main() {
  const int R = 512, K = 64;
  std::vector<__nv_bfloat16> h(R * K);
  std::vector<float> hf(R * K);
  srand(7);
  for (int i = 0; i < R * K; ++i) {
    hf[i] = (float)((rand() % 7) - 3);
    h[i] = __float2bfloat16(hf[i]);
  }
  __nv_bfloat16* d;
  float* out;
  cudaMalloc(&d, R * K * 2);
  cudaMalloc(&out, 256 * 256 * 4);
  cudaMemset(out, 0xff, 256 * 256 * 4);
  cudaMemcpy(d, h.data(), R * K * 2, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  if (!smmd_host::make_tmap_bf16_2d(&tm, d, R, K, K, 128)) { printf("tmap failed\n"); return 2; }
  const int smem = 1024 + 2 * 128 * 128 + 256;
  cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe2_kernel<<<2, 128, smem>>>(tm, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 3; }
  std::vector<float> ho(256 * 256);
  cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  int bad = 0;
  for (int i = 0; i < 256; ++i)
    for (int j = 0; j < 256; ++j) {
      float ref = 0;
      for (int k = 0; k < K; ++k) ref += hf[i * K + k] * hf[(256 + j) * K + k];
      double err = fabs(ref - ho[i * 256 + j]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3 && bad < 5) { printf("mismatch D[%d][%d] = %g, ref %g\n", i, j, ho[i * 256 + j], ref); ++bad; }
    }
  printf("umma2_probe: max |err| = %g -> %s\n", maxerr, maxerr < 1e-3 ? "OK" : "FAIL");
  return maxerr < 1e-3 ? 0 : 1;
}

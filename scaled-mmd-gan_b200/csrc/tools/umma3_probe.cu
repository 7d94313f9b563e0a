// umma3_probe.cu -- hardware probe for the MN-major *A* operand of tcgen05.mma read from shared memory
// (the symmetric two-pass path reads a stored W tile "transposed": W[j rows][i cols] is the A operand of
// O_i += W_ji^T Z_j, i.e. A is [M = i][K = j] with M contiguous in memory).
//   O[128 x 256] = sum_k Wt[k][m] * Z[k][f],  Wt: [64 K-rows][128 M-cols] bf16 (two 64x64 SW128 boxes),
//                                             Z : [64 K-rows][256 features] bf16 (four 64x64 SW128 boxes, MN-major B)
// Usage: umma3_probe <lbo_bytes> <sbo_bytes> <kstep_bytes>     (exit code 0 = matches the CPU)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../sm100_ptx.cuh"
#include "../tmap_host.h"

using namespace sm100;

constexpr uint32_t BOX_BYTES = 64 * 128;   // 64 rows x 128 B

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_z, float* outO,
             uint32_t lbo, uint32_t sbo, uint32_t kstep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                    // 2 boxes: M panels [0,64), [64,128)
  uint8_t* sZ = smem + 2 * BOX_BYTES;    // 4 boxes: feature panels
  uint64_t* bars = reinterpret_cast<uint64_t*>(sZ + 4 * BOX_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 6 * BOX_BYTES);
    for (int p = 0; p < 2; ++p) tma_load_2d(sW + p * BOX_BYTES, &tmap_w, &bars[0], p * 64, 0);
    for (int p = 0; p < 4; ++p) tma_load_2d(sZ + p * BOX_BYTES, &tmap_z, &bars[0], p * 64, 0);
  }
  mbar_wait(&bars[0], 0);
  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc(128, 256, kFmtBF16, true, true);   // A and B both MN-major
    for (int kk = 0; kk < 4; ++kk) {   // K = 64 rows, 16 per MMA
      uint64_t da = make_smem_desc_sw128(smem_u32(sW) + kk * kstep, lbo, sbo);
      uint64_t db = make_smem_desc_sw128(smem_u32(sZ) + kk * 2048, BOX_BYTES, 1024);
      umma_ss(tmem, da, db, idesc, kk > 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < 8; ++c) {
    uint32_t v[32];
    tmem_ld_x32(tmem + lane_base + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) outO[tid * 256 + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
  uint32_t lbo = argc > 1 ? atoi(argv[1]) : 8192, sbo = argc > 2 ? atoi(argv[2]) : 1024;
  uint32_t kstep = argc > 3 ? atoi(argv[3]) : 2048;
  const int K = 64, M = 128, F = 256;
  std::vector<__nv_bfloat16> hW(K * M), hZ(K * F);
  std::vector<float> fW(K * M), fZ(K * F);
  srand(11);
  for (int i = 0; i < K * M; ++i) { float v = ((rand() % 2001) - 1000) / 1000.0f; hW[i] = __float2bfloat16(v); fW[i] = __bfloat162float(hW[i]); }
  for (int i = 0; i < K * F; ++i) { float v = ((rand() % 2001) - 1000) / 1000.0f; hZ[i] = __float2bfloat16(v); fZ[i] = __bfloat162float(hZ[i]); }
  __nv_bfloat16 *dW, *dZ; float* dO;
  CK(cudaMalloc(&dW, K * M * 2)); CK(cudaMalloc(&dZ, K * F * 2)); CK(cudaMalloc(&dO, M * F * 4));
  CK(cudaMemcpy(dW, hW.data(), K * M * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dZ, hZ.data(), K * F * 2, cudaMemcpyHostToDevice));
  CUtensorMap tw, tz;
  if (!smmd_host::make_tmap_bf16_2d(&tw, dW, K, M, M, 64)) { printf("tensor map encode failed\n"); return 2; }
  if (!smmd_host::make_tmap_bf16_2d(&tz, dZ, K, F, F, 64)) { printf("tensor map encode failed\n"); return 2; }
  size_t smem = 6 * BOX_BYTES + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(tw, tz, dO, lbo, sbo, kstep);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> O(M * F);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double err = 0, mag = 0;
  for (int m = 0; m < M; ++m)
    for (int f = 0; f < F; ++f) {
      double o = 0;
      for (int k = 0; k < K; ++k) o += (double)fW[k * M + m] * fZ[k * F + f];
      err = fmax(err, fabs(o - O[m * F + f]));
      mag = fmax(mag, fabs(o));
    }
  bool ok = err < 1e-3 * fmax(1.0, mag);
  printf("probe3 A-MN-major lbo=%u sbo=%u kstep=%u : max|dO|=%.3e (|O|max %.3f) %s\n", lbo, sbo, kstep, err, mag,
         ok ? "OK" : "FAIL");
  return ok ? 0 : 1;
}

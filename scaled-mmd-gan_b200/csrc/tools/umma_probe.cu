// umma_probe.cu -- hardware probe for the UMMA encodings the production kernels rely on.
//   stage 1: S[128x128]  = Zi[128xK] * Zj[128xK]^T   (SS, both operands K-major SW128, K = 256)
//   stage 2: O[128x256] += W[128x128] * Zj[128x256]   (A = W from TMEM (mode ts) or from smem (mode ss),
//                                                      B = the SAME Zj panels read MN-major)
// Usage: umma_probe <mode: ts|ss> <lbo_bytes> <sbo_bytes>      (exit code 0 = both stages match the CPU)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../sm100_ptx.cuh"
#include "../tmap_host.h"

using namespace sm100;

constexpr int KD = 256;        // feature dim
constexpr int NP = KD / 64;    // 64-wide panels
constexpr uint32_t PANEL_BYTES = 128 * 128;  // 128 rows x 128 B

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap, float* outS, float* outO, uint32_t lbo, uint32_t sbo,
             int mode_ts, float wscale) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sZi = smem;                         // NP panels
  uint8_t* sZj = smem + NP * PANEL_BYTES;      // NP panels
  uint8_t* sW = smem + 2 * NP * PANEL_BYTES;   // 2 panels (128 x 128 bf16, K-major SW128) for mode ss
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 2 * PANEL_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem;          // cols [0,128)   S accumulator (fp32)
  const uint32_t tW = tmem + 128;    // cols [128,192) W as packed bf16x2
  const uint32_t tO = tmem + 256;    // cols [256,512) O accumulator

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 2 * NP * PANEL_BYTES);
    for (int p = 0; p < NP; ++p) {
      tma_load_2d(sZi + p * PANEL_BYTES, &tmap, &bars[0], p * 64, 0);
      tma_load_2d(sZj + p * PANEL_BYTES, &tmap, &bars[0], p * 64, 128);
    }
  }
  mbar_wait(&bars[0], 0);

  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc1 = make_idesc(128, 128, kFmtBF16, false, false);
    uint32_t acc = 0;
    for (int p = 0; p < NP; ++p)
      for (int k = 0; k < 4; ++k) {
        uint64_t da = make_smem_desc_sw128(smem_u32(sZi + p * PANEL_BYTES) + k * 32, 16, 1024);
        uint64_t db = make_smem_desc_sw128(smem_u32(sZj + p * PANEL_BYTES) + k * 32, 16, 1024);
        umma_ss(tS, da, db, idesc1, acc);
        acc = 1;
      }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();

  // every thread owns one row of S
  const int row = tid;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld_x32(tS + lane_base + c * 32, v);
    tmem_ld_wait();
    uint32_t w[16];
    for (int i = 0; i < 32; ++i) outS[row * 128 + c * 32 + i] = __uint_as_float(v[i]);
    for (int i = 0; i < 16; ++i)
      w[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * wscale, __uint_as_float(v[2 * i + 1]) * wscale);
    if (mode_ts) {
      tmem_st_x16(tW + lane_base + c * 16, w);
    } else {
      // K-major SW128 panel layout: element (r, k) -> panel k/64, row r, 16-B chunk ((k%64)/8) ^ (r%8)
      for (int i = 0; i < 16; ++i) {
        int k = c * 32 + 2 * i;
        int panel = k >> 6, chunk = (k & 63) >> 3, within = (k & 7) * 2;
        uint32_t off = panel * PANEL_BYTES + row * 128 + ((chunk ^ (row & 7)) << 4) + within;
        *reinterpret_cast<uint32_t*>(sW + off) = w[i];
      }
    }
  }
  if (mode_ts) tmem_st_wait();
  else fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc2 = make_idesc(128, 256, kFmtBF16, false, true);
    for (int kk = 0; kk < 8; ++kk) {   // K = 128 rows of Zj, 16 per MMA
      uint64_t db = make_smem_desc_sw128(smem_u32(sZj) + kk * 2048, lbo, sbo);
      if (mode_ts) {
        umma_ts(tO, tW + kk * 8, db, idesc2, kk > 0);
      } else {
        uint64_t da = make_smem_desc_sw128(smem_u32(sW + (kk >> 2) * PANEL_BYTES) + (kk & 3) * 32, 16, 1024);
        umma_ss(tO, da, db, idesc2, kk > 0);
      }
    }
    umma_commit(&bars[2]);
  }
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  for (int c = 0; c < 8; ++c) {
    uint32_t v[32];
    tmem_ld_x32(tO + lane_base + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) outO[row * 256 + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "ts";
  uint32_t lbo = argc > 2 ? atoi(argv[2]) : 16384, sbo = argc > 3 ? atoi(argv[3]) : 1024;
  int mode_ts = (mode[0] == 't');
  const int R = 256;
  std::vector<__nv_bfloat16> hZ(R * KD);
  std::vector<float> fZ(R * KD);
  srand(7);
  for (int i = 0; i < R * KD; ++i) {
    float v = ((rand() % 2001) - 1000) / 1000.0f;
    hZ[i] = __float2bfloat16(v);
    fZ[i] = __bfloat162float(hZ[i]);
  }
  __nv_bfloat16* dZ; float *dS, *dO;
  CK(cudaMalloc(&dZ, R * KD * 2)); CK(cudaMalloc(&dS, 128 * 128 * 4)); CK(cudaMalloc(&dO, 128 * 256 * 4));
  CK(cudaMemcpy(dZ, hZ.data(), R * KD * 2, cudaMemcpyHostToDevice));
  CUtensorMap tm;
  if (!smmd_host::make_tmap_bf16_2d(&tm, dZ, R, KD, KD, 128)) { printf("tensor map encode failed\n"); return 2; }
  size_t smem = 2 * NP * PANEL_BYTES + 2 * PANEL_BYTES + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float wscale = 1.0f / 64.0f;
  probe_kernel<<<1, 128, smem>>>(tm, dS, dO, lbo, sbo, mode_ts, wscale);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> S(128 * 128), O(128 * 256);
  CK(cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  // CPU reference
  double errS = 0, errO = 0, magO = 0;
  std::vector<float> W(128 * 128);
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      double s = 0;
      for (int k = 0; k < KD; ++k) s += (double)fZ[i * KD + k] * fZ[(128 + j) * KD + k];
      errS = fmax(errS, fabs(s - S[i * 128 + j]));
      W[i * 128 + j] = __bfloat162float(__float2bfloat16(S[i * 128 + j] * wscale));
    }
  for (int i = 0; i < 128; ++i)
    for (int c = 0; c < 256; ++c) {
      double o = 0;
      for (int j = 0; j < 128; ++j) o += (double)W[i * 128 + j] * fZ[(128 + j) * KD + c];
      errO = fmax(errO, fabs(o - O[i * 256 + c]));
      magO = fmax(magO, fabs(o));
    }
  bool ok1 = errS < 1e-3, ok2 = errO < 1e-3 * fmax(1.0, magO);
  printf("probe mode=%s lbo=%u sbo=%u : stage1 max|dS|=%.3e %s ; stage2 max|dO|=%.3e (|O|max %.3f) %s\n", mode, lbo,
         sbo, errS, ok1 ? "OK" : "FAIL", errO, magO, ok2 ? "OK" : "FAIL");
  return (ok1 && ok2) ? 0 : 1;
}

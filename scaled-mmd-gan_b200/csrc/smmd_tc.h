// smmd_tc.h -- interface of the tcgen05 tensor-core path (smmd_tc.cu dispatches to smmd_tc_fused / _wz / _gram.cu).
#pragma once
#include "smmd_internal.h"

namespace smmd {

// fused fwd(+bwd) MMD^2 on tensor cores
bool tc_mmd2_supported(int64_t d, int want_grad);
bool tc_mmd2_covers(const KernelFn& kf, const Geometry& g, int want_grad);
size_t tc_mmd2_workspace_bytes(const Geometry& g, int want_grad, int precision);
cudaError_t tc_mmd2_run(const KernelFn& kf, const Geometry& g, const Coefs& c, const SrcLayout& src, int precision,
                        double* scalars, float* dX, float* dY, void* ws, size_t ws_bytes, cudaStream_t s, int* launches,
                        const char** path);

// path-selection options (smmd_tc_common.cuh: Tuning); false = unknown name
bool tc_set_option(const char* name, long long value);
bool tc_small_kernel_disabled();

// KID: batched Gram sums over subsets; returns per-row statistics [nsub][2m][RS_COUNT] inside the workspace
bool tc_kid_supported(int64_t d, int64_t msub, int64_t nsub);
size_t tc_kid_workspace_bytes(int64_t msub, int64_t d, int64_t nsub, int precision);
cudaError_t tc_kid_run(const KernelFn& kf, const void* G, const void* R, int dtype, int64_t ldg, int64_t ldr, int64_t d,
                       const int32_t* idx_g, const int32_t* idx_r, int64_t first, int64_t nsub, int64_t msub,
                       int precision, int want_second_order, void* ws, size_t ws_bytes, double** stats_out,
                       cudaStream_t s, int* launches, const char** path);

}  // namespace smmd

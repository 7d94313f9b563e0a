// placeholder until the tcgen05 kernels land
#include "smmd_tc.h"
namespace smmd {
bool tc_mmd2_supported(int64_t, int) { return false; }
size_t tc_mmd2_workspace_bytes(int64_t, int64_t, int64_t, int, int) { return 0; }
cudaError_t tc_mmd2_run(const KernelFn&, const Geometry&, const Coefs&, const void*, const void*, int, int64_t, int64_t,
                        int, double*, float*, float*, void*, size_t, cudaStream_t, int*, const char**) {
  return cudaErrorNotSupported;
}
bool tc_kid_supported(int64_t) { return false; }
size_t tc_kid_workspace_bytes(int64_t, int64_t, int64_t, int) { return 0; }
cudaError_t tc_kid_run(const KernelFn&, const void*, const void*, int, int64_t, int64_t, int64_t, const int32_t*,
                       const int32_t*, int64_t, int64_t, int64_t, int, int, void*, size_t, double**, cudaStream_t, int*,
                       const char**) {
  return cudaErrorNotSupported;
}
}  // namespace smmd

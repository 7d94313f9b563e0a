// smmd_torch.cpp -- PyTorch C++ extension wrapper (smmd._C) around the C ABI of libsmmd.so for the latency-bound calls.
//
// The shapes every shipped YAML of the reference runs (batch 64, dof_dim 1..16: gan/core/model.py:313-319 with
// configs/*.yml) take ~10 us of GPU time per loss forward + backward; through the ctypes wrapper the host needed
// 24-29 us per call (three torch.empty from Python, ctypes argument marshalling, stream / device queries).  This
// wrapper does the same work from C++: output allocation, current stream, per-stream workspace and the one call
// into the library.  It binds include/smmd.h directly (no kernels of its own); smmd/mmd.py uses it when present and
// the ctypes path otherwise -- both end in the same CUDA kernels, there is no CPU path.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <ATen/cuda/CUDAGraphsUtils.cuh>

#include <map>
#include <mutex>
#include <tuple>

#include "../../../include/smmd.h"

namespace {

std::mutex g_mu;
std::map<std::pair<int, void*>, at::Tensor> g_ws;   // (device, stream) -> workspace, grown on demand

at::Tensor workspace(size_t nbytes, const at::Device& dev, cudaStream_t stream) {
  auto opts = at::TensorOptions().dtype(at::kByte).device(dev);
  if (at::cuda::currentStreamCaptureStatus() != at::cuda::CaptureStatus::None)
    return at::empty({(int64_t)nbytes}, opts);       // graph capture: the buffer belongs to the graph's pool
  std::lock_guard<std::mutex> lock(g_mu);
  auto key = std::make_pair((int)dev.index(), (void*)stream);
  auto it = g_ws.find(key);
  if (it == g_ws.end() || (size_t)it->second.numel() < nbytes) {
    if (it != g_ws.end()) g_ws.erase(it);
    it = g_ws.emplace(key, at::empty({(int64_t)nbytes}, opts)).first;
  }
  return it->second;
}

// problem_addr: address of a filled `smmd_problem` (the ctypes structure smmd/mmd.py caches per call signature).
// Returns (status, scalars[16] f64, dX, dY) for the owned rows; dX / dY are undefined tensors when want_grad is false.
// A non-zero smmd_status is RETURNED (the Python layer raises its SmmdError from it), not thrown.
std::tuple<int64_t, at::Tensor, at::Tensor, at::Tensor> mmd2_fwd_bwd(int64_t problem_addr, const at::Tensor& X, const at::Tensor& Y,
                                                            int64_t own_m, int64_t own_n, bool want_grad, int64_t ws_bytes) {
  TORCH_CHECK(X.is_cuda() && Y.is_cuda(), "smmd: features must be CUDA tensors; there is no CPU fallback");
  const smmd_problem* p = reinterpret_cast<const smmd_problem*>(problem_addr);
  const c10::cuda::CUDAGuard guard(X.device());
  const cudaStream_t stream = c10::cuda::getCurrentCUDAStream(X.device().index()).stream();
  at::Tensor scalars = at::empty({SMMD_NUM_SCALARS}, X.options().dtype(at::kDouble));
  at::Tensor dX, dY;
  if (want_grad) {
    dX = at::empty({own_m, p->d}, X.options().dtype(at::kFloat));
    dY = at::empty({own_n, p->d}, X.options().dtype(at::kFloat));
  }
  at::Tensor ws = workspace((size_t)ws_bytes, X.device(), stream);
  const int st = smmd_mmd2_fwd_bwd(p, X.data_ptr(), Y.data_ptr(), scalars.data_ptr<double>(),
                                   want_grad ? dX.data_ptr<float>() : nullptr, want_grad ? dY.data_ptr<float>() : nullptr,
                                   ws.data_ptr(), (size_t)ws_bytes, (void*)stream);
  return std::make_tuple((int64_t)st, scalars, dX, dY);
}

void release_workspaces() {
  std::lock_guard<std::mutex> lock(g_mu);
  g_ws.clear();
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "smmd._C: C++ wrapper of libsmmd.so's smmd_mmd2_fwd_bwd (outputs, stream and workspace handled in C++)";
  m.def("mmd2_fwd_bwd", &mmd2_fwd_bwd, "fused MMD^2 forward + backward (include/smmd.h: smmd_mmd2_fwd_bwd)");
  m.def("release_workspaces", &release_workspaces);
}

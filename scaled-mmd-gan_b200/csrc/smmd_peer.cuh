// smmd_peer.cuh -- exchange over peer-mapped memory (NVLink / NVSwitch) for the sharded loss: buffer layout, the
// system-scope flag protocol, and the description of the peers' published rows that the operand preparation reads.
//
// Every rank owns one exchange buffer (include/smmd.h, smmd_peer_buffer_bytes), mapped into all peers:
//   [0, 256)        data_flag[2][16] u64: data_flag[step & 1][r] = last step of that parity whose rows rank r has published
//                   (written BY rank r)
//   [256, 512)      sums_flag[2][16] u64: same for rank r's partial sums
//   [512, 520)      step counter u64 (local use only): the last step this rank has completed; a call made with step = 0
//                   takes "counter + 1" on the device, which makes the call sequence CUDA-graph capturable
//   [1024, 5120)    sums[2][16][16] f64: slot (step & 1), source rank r, the 16 scalars of smmd_scalar
//   [8192, ...)     two data slots (step & 1) of rows_local x d elements (sized for fp32): the rank's own rows
// Protocol per step (all stores that cross GPUs are followed by a system-scope fence and a release store of the flag;
// readers poll the flag in their OWN memory with acquire loads and then pull the data over NVLink):
//   publish rows -> raise data_flag[self] in every peer -> each peer's preparation pulls the rows once the flag is up
//   write sums into every peer's slot -> raise sums_flag[self] there -> every rank adds the slots in rank order.
// Slot reuse needs no extra barrier: step k + 2 of a rank follows its step k in stream order, step k completes only after
// every peer has raised its step-k sums flag, and a peer does that after it has finished reading step k's rows.  Flags are
// kept per slot parity, so two consecutive steps of one rank may be in flight on two streams (step k + 1's flags say
// nothing about step k); calls of the same parity must be stream-ordered.
#pragma once
#include <cstddef>
#include <cstdint>
#include "../../include/smmd.h"

namespace smmd {

constexpr int kPeerMax = SMMD_MAX_PEERS;
constexpr size_t kPeerOffDataFlag = 0;
constexpr size_t kPeerOffSumsFlag = 256;
constexpr size_t kPeerOffStep = 512;
constexpr size_t kPeerOffSums = 1024;
constexpr size_t kPeerOffData = 8192;

__host__ __device__ inline size_t peer_flag_index(unsigned long long step, int r) { return (size_t)(step & 1) * kPeerMax + (size_t)r; }
__host__ __device__ inline size_t peer_slot_bytes(int64_t rows_local, int64_t d) {
  return ((size_t)rows_local * (size_t)d * 4 + 255) / 256 * 256;
}
__host__ __device__ inline size_t peer_buffer_bytes(int64_t rows_local, int64_t d) {
  return kPeerOffData + 2 * peer_slot_bytes(rows_local, d);
}

// The rows every rank has published for this step, as the operand preparation sees them: block r of the global X (Y)
// rows lives in rank r's data slot (step & 1) at rows [0, blk_x) ([blk_x, blk_x + blk_y)), pitch = d elements.
// Passed by value to kernels.  step = 0: the step number is read on the device (own counter + 1).
struct PeerSrc {
  int on;                               // 0 = not a peer call
  int world, self;
  void* base[kPeerMax];                 // every rank's exchange buffer as mapped here
  size_t slot_bytes;                    // peer_slot_bytes(rows_local, d)
  unsigned long long step;
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// step of this call: given by the host, or (step == 0) one more than the last step this rank completed
__device__ __forceinline__ unsigned long long peer_step_of(unsigned long long step, void* own_base) {
  if (step) return step;
  return *reinterpret_cast<volatile unsigned long long*>(static_cast<char*>(own_base) + kPeerOffStep) + 1ull;
}
__device__ __forceinline__ char* peer_data_slot(void* base, unsigned long long step, size_t slot_bytes) {
  return static_cast<char*>(base) + kPeerOffData + (size_t)(step & 1) * slot_bytes;
}
__device__ __forceinline__ unsigned long long* peer_flag(void* base, size_t flag_off, unsigned long long step, int r) {
  return reinterpret_cast<unsigned long long*>(static_cast<char*>(base) + flag_off) + peer_flag_index(step, r);
}
// Bounded wait for flag >= step (a peer that never arrives is an error, not a hang): ~4 s of polling, then trap.
__device__ __forceinline__ void peer_wait_flag(const unsigned long long* flag, unsigned long long step) {
  if (ld_acquire_sys(flag) >= step) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
      __nanosleep(64);
      if (ld_acquire_sys(flag) >= step) return;
    }
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 4000000000ull) __trap();
  }
}
#endif

}  // namespace smmd

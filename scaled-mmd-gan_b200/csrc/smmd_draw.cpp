// smmd_draw.cpp -- host-side helper (no GPU work): the subset draw of polynomial_mmd_averages
// (gan/compute_scores.py:219-222: per subset np.random.choice(len(codes_g), m, replace=False), then the same for codes_r),
// reproduced bit for bit from numpy's legacy global RNG state (MT19937).
//
// numpy: choice(n, m, replace=False) = permutation(n)[:m]; permutation = Fisher-Yates shuffle of arange(n) from the top,
// one random_interval(i) per position: masked 32-bit outputs, rejected while above i.  One call = 200 such shuffles of 50000
// indices at the scorer's size: ~1.4e7 outputs of ONE sequential stream, 107 ms in numpy on the bench box's host.
//
// Here the stream is walked twice:
//   scan   (one thread, sequential by nature): for every shuffle, remember the generator state it starts from and skip
//          over it -- counting only how many outputs it consumes, which needs the accept / reject decisions but not the
//          array.  Sixteen outputs are classified at a time: a masked value <= i - 16 is accepted whatever the fifteen before
//          it did, one > i is rejected whatever they did, and only a value in (i - 16, i] (16 in 2^k of them) sends the block
//          to the one-at-a-time path.
//   shuffle (all threads, one shuffle at a time, each from its remembered starting state): the real shuffle, branch-free
//          (a rejected draw swaps position i with itself and does not advance), first m entries out.
// The state left behind is the scan's: exactly where numpy's loop would have left the global stream.
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../../include/smmd.h"

namespace {

// Hot loops get an AVX2 clone next to the baseline one (resolved once at load time by the dynamic linker).
#define SMMD_CLONES __attribute__((target_clones("avx2", "default")))

SMMD_CLONES void mt_temper(const uint32_t* __restrict__ key, uint32_t* __restrict__ out) {
  for (int i = 0; i < 624; ++i) {
    uint32_t y = key[i];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    out[i] = y;
  }
}
SMMD_CLONES void mt_refill(uint32_t* key, uint32_t* __restrict__ out) {
  constexpr int N = 624, M = 397;
  constexpr uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
  int i = 0;
  for (; i < N - M; ++i) {
    const uint32_t y = (key[i] & UP) | (key[i + 1] & LO);
    key[i] = key[i + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
  }
  for (; i < N - 1; ++i) {
    const uint32_t y = (key[i] & UP) | (key[i + 1] & LO);
    key[i] = key[i + (M - N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
  }
  const uint32_t y = (key[N - 1] & UP) | (key[0] & LO);
  key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
  for (int k = 0; k < 624; ++k) {
    uint32_t t = key[k];
    t ^= t >> 11;
    t ^= (t << 7) & 0x9d2c5680u;
    t ^= (t << 15) & 0xefc60000u;
    t ^= t >> 18;
    out[k] = t;
  }
}

struct alignas(64) Mt19937 {
  uint32_t key[624];
  uint32_t out[624];   // tempered outputs of the current block
  int pos;             // next output (624 = block used up)
  void refill() {
    mt_refill(key, out);
    pos = 0;
  }
  void temper() { mt_temper(key, out); }
  inline uint32_t next() {
    if (pos == 624) refill();
    return out[pos++];
  }
};

inline uint32_t mask_of(uint32_t v) {   // smallest all-ones mask >= v
  v |= v >> 1;
  v |= v >> 2;
  v |= v >> 4;
  v |= v >> 8;
  v |= v >> 16;
  return v;
}

// The generator state a shuffle starts from (the tempered block is rebuilt from the key on restore).
struct Snap {
  uint32_t key[624];
  int pos;
};
inline void snap_of(const Mt19937& g, Snap& s) {
  memcpy(s.key, g.key, sizeof(s.key));
  s.pos = g.pos;
}
inline void restore(Mt19937& g, const Snap& s) {
  memcpy(g.key, s.key, sizeof(g.key));
  g.pos = s.pos;
  if (g.pos < 624) g.temper();
}

// scan: consume exactly the outputs a shuffle of arange(n) consumes.
// classify<B>: B outputs at once.  With lo = i - B: a masked value <= lo is accepted whatever the draws before it in the
// block did (i has dropped by fewer than B), one > i is rejected whatever they did; a value in (lo, i] depends on them --
// then the block is walked one draw at a time.  Values and i are < 2^31 (n < 2^31): signed compares, which vectorise.
template <int B>
static inline bool classify(const uint32_t* o, uint32_t i, uint32_t mask, uint32_t& sure_out) {
  const int lo = (int)(i - B), hi = (int)i, mk = (int)mask;
  int sure = 0, amb = 0;
  for (int t = 0; t < B; ++t) {
    const int v = (int)o[t] & mk;
    sure += v <= lo;
    amb += (v > lo) & (v <= hi);
  }
  sure_out = (uint32_t)sure;
  return amb == 0;
}
template <int B>
static inline bool skip_block(Mt19937& g, uint32_t& i, uint32_t mask) {
  if (!(g.pos + B <= 624 && i > (uint32_t)B && i - B > (mask >> 1))) return false;   // (the mask cannot change within B draws)
  uint32_t sure;
  if (classify<B>(g.out + g.pos, i, mask, sure)) {
    i -= sure;
    g.pos += B;
  } else {
    for (int t = 0; t < B; ++t) i -= (g.out[g.pos++] & mask) <= i ? 1u : 0u;
  }
  return true;
}
SMMD_CLONES void skip_shuffle(Mt19937& g, int64_t n) {
  uint32_t i = (uint32_t)(n - 1);
  uint32_t mask = mask_of(i);
  while (i >= 1) {
    if (g.pos == 624) g.refill();
    if (skip_block<64>(g, i, mask)) continue;
    if (skip_block<16>(g, i, mask)) continue;
    const uint32_t v = g.out[g.pos++] & mask;
    i -= v <= i ? 1u : 0u;
    mask >>= (i <= (mask >> 1)) ? 1 : 0;
  }
}

// permutation(n)[:m]: numpy's _shuffle_raw order, no data-dependent branch
void shuffle_take(Mt19937& g, int32_t* __restrict__ scratch, int64_t n, int32_t m, int32_t* __restrict__ out) {
  for (int64_t k = 0; k < n; ++k) scratch[k] = (int32_t)k;
  uint32_t i = (uint32_t)(n - 1);
  uint32_t mask = mask_of(i);
  while (i >= 1) {
    const uint32_t v = g.next() & mask;
    const uint32_t acc = v <= i ? 1u : 0u;
    const uint32_t j = i ^ ((v ^ i) & (0u - acc));   // acc ? v : i
    const int32_t a = scratch[i], b = scratch[j];
    scratch[i] = b;
    scratch[j] = a;
    i -= acc;
    mask >>= (i <= (mask >> 1)) ? 1 : 0;
  }
  for (int32_t k = 0; k < m; ++k) out[k] = scratch[k];
}

}  // namespace

extern "C" SMMD_API int smmd_draw_subsets_mt19937(uint32_t* key, int32_t* pos, int64_t len_g, int64_t len_r, int32_t n_subsets,
                                         int32_t subset_size, int32_t* idx_g, int32_t* idx_r) {
  if (!key || !pos || !idx_g || !idx_r) return SMMD_EINVAL;
  if (*pos < 0 || *pos > 624 || n_subsets < 0 || subset_size < 0) return SMMD_EINVAL;
  if (len_g < 1 || len_r < 1 || len_g >= ((int64_t)1 << 31) || len_r >= ((int64_t)1 << 31)) return SMMD_ESHAPE;
  if (subset_size > len_g || subset_size > len_r) return SMMD_ESHAPE;   // numpy: "Cannot take a larger sample than population"
  const int64_t nmax = len_g > len_r ? len_g : len_r;
  const int jobs = 2 * n_subsets;   // job 2s = subset s of g, job 2s + 1 = subset s of r (numpy's draw order)
  auto job_n = [&](int j) { return (j & 1) ? len_r : len_g; };
  auto job_out = [&](int j) { return ((j & 1) ? idx_r : idx_g) + (size_t)(j >> 1) * subset_size; };
  alignas(64) Mt19937 g;
  memcpy(g.key, key, sizeof(g.key));
  g.pos = *pos;
  if (g.pos < 624) g.temper();

  // helper threads for the shuffle phase: SMMD_DRAW_THREADS (total threads incl. this one; 1 = sequential), default
  // min(16, hardware threads)
  unsigned hw = std::thread::hardware_concurrency();
  int workers = (int)(hw > 1 ? hw - 1 : 0);
  if (workers > 15) workers = 15;
  if (const char* e = getenv("SMMD_DRAW_THREADS")) {
    const int t = atoi(e);
    if (t >= 1 && t <= 64) workers = t - 1;
  }
  if ((int64_t)jobs * nmax < 400000 || jobs < 4) workers = 0;   // small draws: one thread, no scan
  if (workers == 0) {
    std::vector<int32_t> scratch((size_t)nmax);
    for (int j = 0; j < jobs; ++j) shuffle_take(g, scratch.data(), job_n(j), subset_size, job_out(j));
  } else {
    // scan first (alone: a spinning or busy sibling thread slows this sequential pass more than pipelining would gain),
    // then every thread -- this one included -- takes shuffles until none are left
    std::vector<Snap> snaps((size_t)jobs);
    for (int j = 0; j < jobs; ++j) {
      snap_of(g, snaps[(size_t)j]);
      skip_shuffle(g, job_n(j));
    }
    std::atomic<int> claim{0};
    auto work = [&]() {
      std::vector<int32_t> scratch((size_t)nmax);
      Mt19937 lg;
      for (;;) {
        const int j = claim.fetch_add(1, std::memory_order_relaxed);
        if (j >= jobs) return;
        restore(lg, snaps[(size_t)j]);
        shuffle_take(lg, scratch.data(), job_n(j), subset_size, job_out(j));
      }
    };
    std::vector<std::thread> pool;
    pool.reserve((size_t)workers);
    for (int w = 0; w < workers; ++w) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
  }
  memcpy(key, g.key, sizeof(g.key));
  *pos = g.pos;
  return SMMD_OK;
}

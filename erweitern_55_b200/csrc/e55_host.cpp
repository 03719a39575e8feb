// Host-side input compaction for the host-buffer predict (e55_predict): a small thread pool turns the reference's
// float32 planes (network.py:255-274; 26 600 B per position) into the bit-packed form the tower kernel expands itself
// (one occupancy word + one value per plane, 2 128 B per position) while earlier chunks are already on the GPU, so the
// 27 MB host->device copy that bounds the float path at batch 1024 shrinks 12.5x.  Exact by construction: a plane is
// packed only if all its non-zero entries hold one and the same value (occupancy, repetition, hand-count, colour and
// ply planes all do); anything else is reported and the caller sends that chunk as plain float planes.
// Host-only translation unit (g++, AVX2).
#include "e55_host.h"

#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace e55host {

namespace {

constexpr int kSq = 25;

// One 5x5 plane -> (occupancy bits, value).  False if two different non-zero values occur.
inline bool pack_plane(const float* q, uint32_t& bits, float& value) {
  const __m256 zero = _mm256_setzero_ps();
  const __m256 v0 = _mm256_loadu_ps(q), v1 = _mm256_loadu_ps(q + 8), v2 = _mm256_loadu_ps(q + 16), v3 = _mm256_loadu_ps(q + 17);
  const uint32_t m0 = static_cast<uint32_t>(_mm256_movemask_ps(_mm256_cmp_ps(v0, zero, _CMP_NEQ_UQ)));
  const uint32_t m1 = static_cast<uint32_t>(_mm256_movemask_ps(_mm256_cmp_ps(v1, zero, _CMP_NEQ_UQ)));
  const uint32_t m2 = static_cast<uint32_t>(_mm256_movemask_ps(_mm256_cmp_ps(v2, zero, _CMP_NEQ_UQ)));
  const uint32_t m3 = static_cast<uint32_t>(_mm256_movemask_ps(_mm256_cmp_ps(v3, zero, _CMP_NEQ_UQ)));
  bits = m0 | (m1 << 8) | (m2 << 16) | (((m3 >> 7) & 1u) << 24);
  if (bits == 0) { value = 1.f; return true; }
  value = q[__builtin_ctz(bits)];
  const __m256 s = _mm256_set1_ps(value);
  auto uniform = [&](__m256 v) {   // every lane is 0 or the value
    return _mm256_movemask_ps(_mm256_or_ps(_mm256_cmp_ps(v, s, _CMP_EQ_OQ), _mm256_cmp_ps(v, zero, _CMP_EQ_OQ)));
  };
  return uniform(v0) == 0xFF && uniform(v1) == 0xFF && uniform(v2) == 0xFF && (uniform(v3) & 0x80);
}

bool pack_positions(const float* planes, int first, int count, int in_planes, uint32_t* bits, float* scale) {
  bool exact = true;
  for (int p = first; p < first + count; ++p) {
    const float* src = planes + static_cast<size_t>(p) * in_planes * kSq;
    uint32_t* b = bits + static_cast<size_t>(p) * in_planes;
    float* s = scale + static_cast<size_t>(p) * in_planes;
    for (int c = 0; c < in_planes; ++c) exact &= pack_plane(src + c * kSq, b[c], s[c]);   // (loads stay inside the plane)
  }
  return exact;
}

}  // namespace

struct Packer {
  static constexpr int kBlock = 8;   // positions per work item
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable cv;
  std::atomic<uint64_t> generation{0};
  bool stop = false;
  // the current job
  const float* planes = nullptr;
  uint32_t* bits = nullptr;
  float* scale = nullptr;
  int n = 0, in_planes = 0, n_blocks = 0;
  std::atomic<int> next{0};
  std::atomic<int> idle{0};                       // workers that have left the current job
  std::vector<std::atomic<uint8_t>> block_state;  // 0 pending, 1 packed exactly, 2 not representable

  explicit Packer(int n_threads) : block_state(0) {
    for (int i = 0; i < n_threads; ++i) threads.emplace_back([this] { work(); });
  }
  ~Packer() {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
      generation.fetch_add(1, std::memory_order_release);
    }
    cv.notify_all();
    for (auto& t : threads) t.join();
  }

  void run_blocks() {
    for (;;) {
      const int b = next.fetch_add(1, std::memory_order_relaxed);
      if (b >= n_blocks) return;
      const int first = b * kBlock, count = std::min(kBlock, n - first);
      const bool exact = pack_positions(planes, first, count, in_planes, bits, scale);
      block_state[b].store(exact ? 1 : 2, std::memory_order_release);
    }
  }

  void work() {
    uint64_t seen = 0;
    for (;;) {
      // spin a little for back-to-back calls, then sleep
      uint64_t g = generation.load(std::memory_order_acquire);
      for (int spin = 0; g == seen && spin < 2000; ++spin) { _mm_pause(); g = generation.load(std::memory_order_acquire); }
      if (g == seen) {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return generation.load(std::memory_order_acquire) != seen; });
        g = generation.load(std::memory_order_acquire);
      }
      seen = g;
      if (stop) return;
      run_blocks();
      idle.fetch_add(1, std::memory_order_release);
    }
  }

  void start(const float* planes_, int n_, int in_planes_, uint32_t* bits_, float* scale_) {
    planes = planes_; n = n_; in_planes = in_planes_; bits = bits_; scale = scale_;
    n_blocks = (n + kBlock - 1) / kBlock;
    if (static_cast<int>(block_state.size()) < n_blocks) block_state = std::vector<std::atomic<uint8_t>>(n_blocks);
    for (int b = 0; b < n_blocks; ++b) block_state[b].store(0, std::memory_order_relaxed);
    next.store(0, std::memory_order_relaxed);
    idle.store(0, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(m);
      generation.fetch_add(1, std::memory_order_release);
    }
    cv.notify_all();
  }

  // Wait until positions [first, first + count) are packed; true if all of them were representable.
  bool wait_range(int first, int count) {
    bool exact = true;
    for (int b = first / kBlock; b <= (first + count - 1) / kBlock; ++b) {
      uint8_t st;
      while ((st = block_state[b].load(std::memory_order_acquire)) == 0) {
        const int mine = next.fetch_add(1, std::memory_order_relaxed);   // lend a hand instead of just waiting
        if (mine < n_blocks) {
          const int f = mine * kBlock;
          block_state[mine].store(pack_positions(planes, f, std::min(kBlock, n - f), in_planes, bits, scale) ? 1 : 2, std::memory_order_release);
        } else {
          _mm_pause();
        }
      }
      exact &= st == 1;
    }
    return exact;
  }

  // All workers have left the job (its buffers may be reused or freed).
  void finish() {
    run_blocks();   // nothing left normally; covers a pool whose threads have not woken up yet
    while (idle.load(std::memory_order_acquire) < static_cast<int>(threads.size())) _mm_pause();
  }
};

Packer* packer_create(int threads) { return threads > 0 ? new (std::nothrow) Packer(threads) : nullptr; }
void packer_destroy(Packer* p) { delete p; }
int packer_threads(const Packer* p) { return p ? static_cast<int>(p->threads.size()) : 0; }
void packer_start(Packer* p, const float* planes, int n, int in_planes, uint32_t* bits, float* scale) {
  p->start(planes, n, in_planes, bits, scale);
}
bool packer_wait_range(Packer* p, int first, int count) { return p->wait_range(first, count); }
void packer_finish(Packer* p) { p->finish(); }
bool pack_planes_inline(const float* planes, int n, int in_planes, uint32_t* bits, float* scale) {
  return pack_positions(planes, 0, n, in_planes, bits, scale);
}

}  // namespace e55host

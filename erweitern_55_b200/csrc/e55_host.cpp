// Host-side input compaction for the host-buffer predict (e55_predict): a small thread pool turns the reference's
// float32 planes (network.py:255-274; 26 600 B per position) into the bit-packed form the tower kernel expands itself
// (one occupancy word + one value per plane, 2 128 B per position) while earlier chunks are already on the GPU, so the
// 27 MB host->device copy that bounds the float path at batch 1024 shrinks 12.5x.  Exact by construction: a plane is
// packed only if all its non-zero entries hold one and the same value (occupancy, repetition, hand-count, colour and
// ply planes all do); anything else is reported and the caller sends that chunk as plain float planes.
// The same pool copies results out of the library's pinned staging into pageable caller buffers (parallel_copy).
// Host-only translation unit (g++): AVX2 baseline, AVX-512 kernels chosen at run time where the CPU has them.
#include "e55_host.h"

#include <immintrin.h>
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
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

// (bits / scale of position p start at word p * stride: stride = in_planes for two separate arrays, 2 * in_planes for
// the interleaved staging the host-buffer predict copies to the device in one piece)
bool pack_positions_avx2(const float* planes, int first, int count, int in_planes, uint32_t* bits, float* scale, int stride) {
  bool exact = true;
  for (int p = first; p < first + count; ++p) {
    const float* src = planes + static_cast<size_t>(p) * in_planes * kSq;
    uint32_t* b = bits + static_cast<size_t>(p) * stride;
    float* s = scale + static_cast<size_t>(p) * stride;
    for (int c = 0; c < in_planes; ++c) exact &= pack_plane(src + c * kSq, b[c], s[c]);   // (loads stay inside the plane)
  }
  return exact;
}

// AVX-512: a plane is one full and one 9-lane masked load, the occupancy word is the two compare masks, and the
// uniformity test is one more compare per half -- branch-free, so the loop runs at the rate the planes stream in.
__attribute__((target("avx512f,avx512vl,avx512bw,avx512dq")))
bool pack_positions_avx512(const float* planes, int first, int count, int in_planes, uint32_t* bits, float* scale, int stride) {
  const __m512 zero = _mm512_setzero_ps();
  uint32_t bad = 0;
  for (int p = first; p < first + count; ++p) {
    const float* src = planes + static_cast<size_t>(p) * in_planes * kSq;
    uint32_t* b = bits + static_cast<size_t>(p) * stride;
    float* s = scale + static_cast<size_t>(p) * stride;
    for (int c = 0; c < in_planes; ++c) {
      const float* q = src + c * kSq;
      _mm_prefetch(reinterpret_cast<const char*>(q + 512), _MM_HINT_T0);
      const __m512 lo = _mm512_loadu_ps(q);
      const __m512 hi = _mm512_maskz_loadu_ps(0x1FF, q + 16);
      const uint32_t k0 = _mm512_cmp_ps_mask(lo, zero, _CMP_NEQ_UQ), k1 = _mm512_cmp_ps_mask(hi, zero, _CMP_NEQ_UQ);
      const uint32_t w = k0 | (k1 << 16);
      float v = q[w ? __builtin_ctz(w) : 0];
      v = w ? v : 1.f;
      const __m512 sv = _mm512_set1_ps(v);
      const uint32_t e0 = _mm512_cmp_ps_mask(lo, sv, _CMP_EQ_OQ), e1 = _mm512_cmp_ps_mask(hi, sv, _CMP_EQ_OQ);
      bad |= (k0 & ~e0) | (k1 & ~e1);   // a non-zero entry that is not the value (NaN included)
      b[c] = w;
      s[c] = v;
    }
  }
  return bad == 0;
}

using PackFn = bool (*)(const float*, int, int, int, uint32_t*, float*, int);

PackFn choose_pack() {
  const char* env = getenv("E55_HOST_ISA");
  if (env && !strcmp(env, "avx2")) return pack_positions_avx2;
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512bw") &&
      __builtin_cpu_supports("avx512dq"))
    return pack_positions_avx512;
  return pack_positions_avx2;
}

const PackFn g_pack = choose_pack();

inline bool pack_positions(const float* planes, int first, int count, int in_planes, uint32_t* bits, float* scale, int stride) {
  return g_pack(planes, first, count, in_planes, bits, scale, stride);
}

}  // namespace

struct Packer {
  static constexpr int kBlock = 8;   // positions per work item
  static constexpr size_t kCopyBlock = 128 << 10;   // bytes per copy work item
  static constexpr int kSpinUs = 400;               // how long an idle worker polls before it sleeps
  std::vector<std::thread> threads;
  std::atomic<uint32_t> generation{0};   // bumped per job; sleeping workers wait on it (a futex: no mutex hand-over on wake-up)
  std::atomic<bool> stop{false};
  // the current job: pack (planes != null) or copy (copy_src != null)
  const float* planes = nullptr;
  uint32_t* bits = nullptr;
  float* scale = nullptr;
  int n = 0, in_planes = 0, stride = 0, n_blocks = 0;
  const uint8_t* copy_src = nullptr;
  uint8_t* copy_dst = nullptr;         // null with copy_src set: read only (bandwidth probe)
  size_t copy_bytes = 0;
  std::atomic<uint64_t> read_sink{0};
  std::atomic<int> next{0};
  std::atomic<int> idle{0};                       // workers that have left the current job
  std::vector<std::atomic<uint8_t>> block_state;  // 0 pending, 1 packed exactly, 2 not representable

  explicit Packer(int n_threads) : block_state(0) {
    // The workers inherit the process's affinity mask (several ranks on one box each own a block of cores: bench.py / the
    // launcher set it), so they stay on this rank's cores.  Binding each worker to one core of the mask is optional
    // (E55_PIN_THREADS=1): the calling thread is not bound, and when the scheduler leaves it on a core whose bound worker
    // is polling, every call waits for time slices (measured: 3.4 ms instead of 0.34 ms per batch-1024 call, at random).
    std::vector<int> cpus;
    cpu_set_t set;
    const char* pin = getenv("E55_PIN_THREADS");
    if (pin && !strcmp(pin, "1") && sched_getaffinity(0, sizeof(set), &set) == 0)
      for (int c = 0; c < CPU_SETSIZE; ++c)
        if (CPU_ISSET(c, &set)) cpus.push_back(c);
    for (int i = 0; i < n_threads; ++i) {
      threads.emplace_back([this] { work(); });
      if (static_cast<int>(cpus.size()) > n_threads) {
        cpu_set_t one;
        CPU_ZERO(&one);
        CPU_SET(cpus[1 + i], &one);
        pthread_setaffinity_np(threads.back().native_handle(), sizeof(one), &one);
      }
    }
  }
  ~Packer() {
    stop.store(true, std::memory_order_release);
    generation.fetch_add(1, std::memory_order_release);
    generation.notify_all();
    for (auto& t : threads) t.join();
  }

  void run_blocks() {
    if (copy_src) {
      for (;;) {
        const int b = next.fetch_add(1, std::memory_order_relaxed);
        if (b >= n_blocks) return;
        const size_t o = static_cast<size_t>(b) * kCopyBlock, len = std::min(kCopyBlock, copy_bytes - o);
        if (copy_dst) {
          memcpy(copy_dst + o, copy_src + o, len);
        } else {                       // stream the block through the core: 64-bit loads, four independent sums
          const uint64_t* q = reinterpret_cast<const uint64_t*>(copy_src + o);
          uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
          size_t i = 0;
          for (; i + 4 <= len / 8; i += 4) { a0 += q[i]; a1 += q[i + 1]; a2 += q[i + 2]; a3 += q[i + 3]; }
          read_sink.fetch_add(a0 + a1 + a2 + a3, std::memory_order_relaxed);
        }
      }
    }
    for (;;) {
      const int b = next.fetch_add(1, std::memory_order_relaxed);
      if (b >= n_blocks) return;
      const int first = b * kBlock, count = std::min(kBlock, n - first);
      const bool exact = pack_positions(planes, first, count, in_planes, bits, scale, stride);
      block_state[b].store(exact ? 1 : 2, std::memory_order_release);
    }
  }

  void work() {
    uint32_t seen = 0;
    for (;;) {
      // spin for a while (back-to-back calls of a search loop find the pool awake), then sleep on the counter
      uint32_t g = generation.load(std::memory_order_acquire);
      const auto t0 = std::chrono::steady_clock::now();
      for (int spin = 0; g == seen; ++spin) {
        _mm_pause();
        g = generation.load(std::memory_order_acquire);
        if ((spin & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(kSpinUs)) break;
      }
      while (g == seen) {
        generation.wait(seen, std::memory_order_acquire);
        g = generation.load(std::memory_order_acquire);
      }
      seen = g;
      if (stop.load(std::memory_order_acquire)) return;
      run_blocks();
      idle.fetch_add(1, std::memory_order_release);
    }
  }

  void publish() {
    next.store(0, std::memory_order_relaxed);
    idle.store(0, std::memory_order_relaxed);
    generation.fetch_add(1, std::memory_order_release);
    generation.notify_all();
  }

  void start(const float* planes_, int n_, int in_planes_, uint32_t* bits_, float* scale_, int stride_) {
    copy_src = nullptr;
    planes = planes_; n = n_; in_planes = in_planes_; bits = bits_; scale = scale_; stride = stride_ ? stride_ : in_planes_;
    n_blocks = (n + kBlock - 1) / kBlock;
    if (static_cast<int>(block_state.size()) < n_blocks) block_state = std::vector<std::atomic<uint8_t>>(n_blocks);
    for (int b = 0; b < n_blocks; ++b) block_state[b].store(0, std::memory_order_relaxed);
    publish();
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
          block_state[mine].store(pack_positions(planes, f, std::min(kBlock, n - f), in_planes, bits, scale, stride) ? 1 : 2, std::memory_order_release);
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

  // memcpy shared by the pool and the calling thread; returns when every byte is in place.
  void parallel_copy(void* dst, const void* src, size_t bytes) {
    if (bytes <= 2 * kCopyBlock) { memcpy(dst, src, bytes); return; }
    planes = nullptr;
    copy_src = static_cast<const uint8_t*>(src); copy_dst = static_cast<uint8_t*>(dst); copy_bytes = bytes;
    n_blocks = static_cast<int>((bytes + kCopyBlock - 1) / kCopyBlock);
    publish();
    finish();
    copy_src = nullptr;
  }
};

// Reads `bytes` at `src` once with the pool's threads and the caller (nothing is written): GB/s of the pass.
double Packer_read_pass(Packer* p, const void* src, size_t bytes) {
  const auto t0 = std::chrono::steady_clock::now();
  p->planes = nullptr;
  p->copy_src = static_cast<const uint8_t*>(src); p->copy_dst = nullptr; p->copy_bytes = bytes;
  p->n_blocks = static_cast<int>((bytes + Packer::kCopyBlock - 1) / Packer::kCopyBlock);
  p->publish();
  p->finish();
  p->copy_src = nullptr;
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return static_cast<double>(bytes) / s / 1e9;
}

Packer* packer_create(int threads) { return threads > 0 ? new (std::nothrow) Packer(threads) : nullptr; }
void packer_destroy(Packer* p) { delete p; }
int packer_threads(const Packer* p) { return p ? static_cast<int>(p->threads.size()) : 0; }
void packer_start(Packer* p, const float* planes, int n, int in_planes, uint32_t* bits, float* scale, int stride) {
  p->start(planes, n, in_planes, bits, scale, stride);
}
bool packer_wait_range(Packer* p, int first, int count) { return p->wait_range(first, count); }
void packer_finish(Packer* p) { p->finish(); }
void packer_parallel_copy(Packer* p, void* dst, const void* src, size_t bytes) {
  if (p) p->parallel_copy(dst, src, bytes); else memcpy(dst, src, bytes);
}
bool pack_planes_inline(const float* planes, int n, int in_planes, uint32_t* bits, float* scale, int stride) {
  return pack_positions(planes, 0, n, in_planes, bits, scale, stride ? stride : in_planes);
}
double packer_read_bandwidth(Packer* p, const void* src, size_t bytes) { return p ? Packer_read_pass(p, src, bytes) : 0.0; }
const char* pack_isa() { return g_pack == pack_positions_avx512 ? "avx512" : "avx2"; }

}  // namespace e55host

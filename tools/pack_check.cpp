// Host-side plane packer (csrc/e55_host.cpp) against a scalar restatement, with the thread pool, over rotating buffers;
// also the harness of the sanitizer builds (tests/test_sanitizers.py) and a throughput probe:
//   g++ -O3 -std=c++20 -march=x86-64-v3 -pthread -Ierweitern_55_b200/csrc tools/pack_check.cpp erweitern_55_b200/csrc/e55_host.cpp -o /tmp/pack_check
//   /tmp/pack_check <threads> <calls>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "e55_host.h"

int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 4, calls = argc > 2 ? atoi(argv[2]) : 30;
  const int n = 1000, ip = 266, kSq = 25, nbuf = 3;   // ragged: not a multiple of the pool's block of 8... nor of the chunk
  uint64_t rng = 12345;
  auto rnd = [&] { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  std::vector<std::vector<float>> bufs(nbuf, std::vector<float>(static_cast<size_t>(n) * ip * kSq, 0.f));
  std::vector<std::vector<uint8_t>> inexact(nbuf, std::vector<uint8_t>(n, 0));
  for (int b = 0; b < nbuf; ++b)
    for (int p = 0; p < n; ++p)
      for (int c = 0; c < ip; ++c) {
        float* q = bufs[b].data() + (static_cast<size_t>(p) * ip + c) * kSq;
        const uint64_t kind = rnd() % 16;
        const float v = kind < 12 ? 1.f : static_cast<float>(rnd() % 9 + 1) / 2.f;
        if (kind == 15) for (int s = 0; s < kSq; ++s) q[s] = v;                       // constant plane
        else for (int s = 0; s < kSq; ++s) q[s] = rnd() % 5 == 0 ? v : (rnd() % 64 == 0 ? -0.f : 0.f);
        if (rnd() % 4000 == 0) { q[rnd() % 12] = 0.25f; q[12 + rnd() % 13] = 0.75f; inexact[b][p] = 1; }   // two values: not packable
      }
  std::vector<uint32_t> bits(static_cast<size_t>(n) * ip);
  std::vector<float> scale(static_cast<size_t>(n) * ip);
  e55host::Packer* pool = e55host::packer_create(threads);
  long bad = 0;
  double us = 0;
  for (int call = 0; call < calls; ++call) {
    const int b = call % nbuf;
    const auto t0 = std::chrono::steady_clock::now();
    e55host::packer_start(pool, bufs[b].data(), n, ip, bits.data(), scale.data());
    std::vector<uint8_t> chunk_exact;
    for (int off = 0; off < n; off += 256) chunk_exact.push_back(e55host::packer_wait_range(pool, off, std::min(256, n - off)));
    e55host::packer_finish(pool);
    us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    for (int p = 0; p < n; ++p) {
      bool chunk_has_inexact = false;
      for (int k = p / 256 * 256; k < std::min(n, p / 256 * 256 + 256); ++k) chunk_has_inexact |= inexact[b][k] != 0;
      if (chunk_exact[p / 256] == chunk_has_inexact) { ++bad; continue; }
      if (inexact[b][p]) continue;
      for (int c = 0; c < ip; ++c) {
        const float* q = bufs[b].data() + (static_cast<size_t>(p) * ip + c) * kSq;
        uint32_t want = 0;
        float value = 1.f;
        for (int s = kSq - 1; s >= 0; --s) if (q[s] != 0.f) { want |= 1u << s; value = q[s]; }
        if (bits[static_cast<size_t>(p) * ip + c] != want || scale[static_cast<size_t>(p) * ip + c] != value) ++bad;
      }
    }
  }
  // the pool's parallel copy (results leaving the pinned staging): ragged size, compared byte for byte
  {
    const size_t bytes = bufs[0].size() * sizeof(float) - 13;
    std::vector<uint8_t> out(bytes + 64, 0xAB);
    for (int rep = 0; rep < 3; ++rep) e55host::packer_parallel_copy(pool, out.data(), bufs[rep % nbuf].data(), bytes);
    if (memcmp(out.data(), bufs[2 % nbuf].data(), bytes) != 0) ++bad;
    for (size_t i = bytes; i < bytes + 64; ++i) if (out[i] != 0xAB) ++bad;
  }
  e55host::packer_destroy(pool);
  printf("pack_check (%s): %d threads, %d calls of %d positions, %.0f us per call, %ld mismatches: %s\n", e55host::pack_isa(), threads, calls, n, us / calls, bad,
         bad ? "FAILED" : "ok");
  return bad ? 1 : 0;
}

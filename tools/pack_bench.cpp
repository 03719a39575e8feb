// Throughput of the host-side plane packer (csrc/e55_host.cpp) from DRAM, beside the pool's pure read pass over the same
// buffers: how close the packer is to the rate at which a host's cores can read the planes at all (bench.py's
// host_read_gbs).  No GPU needed.
//   g++ -O3 -std=c++20 -march=x86-64-v3 -pthread -Ierweitern_55_b200/csrc tools/pack_bench.cpp erweitern_55_b200/csrc/e55_host.cpp -o /tmp/pack_bench
//   /tmp/pack_bench <pool threads> [batch 1024] [rotation buffers 12] [passes 5]
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "e55_host.h"

int main(int argc, char** argv) {
  const int threads = std::max(1, argc > 1 ? atoi(argv[1]) : 4), n = argc > 2 ? atoi(argv[2]) : 1024, nbuf = argc > 3 ? atoi(argv[3]) : 12;
  const int passes = argc > 4 ? atoi(argv[4]) : 5;
  const int ip = 266, kSq = 25;
  const size_t per = static_cast<size_t>(n) * ip * kSq;
  std::vector<std::vector<float>> bufs(nbuf, std::vector<float>(per, 0.f));
  uint64_t rng = 99;
  auto rnd = [&] { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  for (auto& b : bufs)
    for (size_t i = 0; i < per; i += 7) b[i] = (rnd() & 3) == 0 ? 1.f : 0.f;   // sparse occupancy planes
  std::vector<uint32_t> stage(static_cast<size_t>(n) * ip * 2);
  e55host::Packer* pool = e55host::packer_create(threads);
  auto secs = [](auto t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
  std::vector<double> pack_gbs, read_gbs;
  for (int pass = 0; pass < passes; ++pass) {
    auto t0 = std::chrono::steady_clock::now();
    for (int b = 0; b < nbuf; ++b) {
      e55host::packer_start(pool, bufs[b].data(), n, ip, stage.data(), reinterpret_cast<float*>(stage.data()) + ip, 2 * ip);
      for (int off = 0; off < n; off += 128) e55host::packer_wait_range(pool, off, std::min(128, n - off));
      e55host::packer_finish(pool);
    }
    pack_gbs.push_back(static_cast<double>(nbuf) * per * 4 / secs(t0) / 1e9);
    t0 = std::chrono::steady_clock::now();
    for (int b = 0; b < nbuf; ++b) e55host::packer_read_bandwidth(pool, bufs[b].data(), per * 4);
    read_gbs.push_back(static_cast<double>(nbuf) * per * 4 / secs(t0) / 1e9);
  }
  std::sort(pack_gbs.begin(), pack_gbs.end());
  std::sort(read_gbs.begin(), read_gbs.end());
  printf("%s packer, %d pool threads + caller, %d x %.1f MB: pack %.1f GB/s (%.2f M positions/s), read-only %.1f GB/s (medians of %d)\n",
         e55host::pack_isa(), threads, nbuf, per * 4 / 1e6, pack_gbs[passes / 2], pack_gbs[passes / 2] / (ip * kSq * 4) * 1e3,
         read_gbs[passes / 2], passes);
  e55host::packer_destroy(pool);
  return 0;
}

// Host-side helpers of libe55.so that are plain C++ (compiled by g++, see e55_host.cpp).
#pragma once
#include <cstddef>
#include <cstdint>

namespace e55host {

struct Packer;   // thread pool that bit-packs float planes (exactly, or reports that it cannot)

Packer* packer_create(int threads);
void packer_destroy(Packer* p);
int packer_threads(const Packer* p);
// Pack planes [n][in_planes][25] into bits/scale in the background, front to back.  Position i's words start at
// bits + i * stride and scale + i * stride (stride 0 = in_planes: two plain [n][in_planes] arrays).
void packer_start(Packer* p, const float* planes, int n, int in_planes, uint32_t* bits, float* scale, int stride = 0);
// Block until positions [first, first + count) are done; false if one of them is not exactly representable
// (some plane holds two different non-zero values).
bool packer_wait_range(Packer* p, int first, int count);
// Block until every worker has left the job started last.
void packer_finish(Packer* p);
// memcpy shared by the pool's threads and the caller (no pack job may be running); p may be null (plain memcpy).
void packer_parallel_copy(Packer* p, void* dst, const void* src, size_t bytes);
// One read-only pass over [src, src + bytes) by the pool's threads and the caller: GB/s (the host-memory roofline of
// everything that has to read float planes: 26 600 bytes per position).
double packer_read_bandwidth(Packer* p, const void* src, size_t bytes);
// "avx512" or "avx2": the packing kernel chosen for this CPU (E55_HOST_ISA=avx2 forces the baseline).
const char* pack_isa();
// Same packing on the calling thread.
bool pack_planes_inline(const float* planes, int n, int in_planes, uint32_t* bits, float* scale, int stride = 0);

}  // namespace e55host

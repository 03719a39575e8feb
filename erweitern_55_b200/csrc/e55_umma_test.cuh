// Hardware self-test and micro-benchmark of the tcgen05 operand addressing the tower kernel relies on:
// no-swizzle K-major operands with SBO = 128 B so that rows are linear, and an A descriptor whose
// start address is displaced by an arbitrary number of rows (16-byte granularity) = a 3x3 tap shift.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "e55_ptx.cuh"

namespace e55 {
namespace ummatest {

constexpr int kM = 128, kN = 128, kK = 64;
constexpr int kRowsA = kM + 64;  // extra rows so the descriptor can be displaced by up to 64 rows

// A_full [kRowsA][kK] row-major bf16, B [kN][kK] row-major bf16, D [kM][kN] fp32:
// D[m][n] = sum_k A_full[m + row_shift][k] * B[n][k]
__global__ void __launch_bounds__(128)
selftest_kernel(const __nv_bfloat16* __restrict__ a_full, const __nv_bfloat16* __restrict__ b,
                float* __restrict__ d, int row_shift) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                               // [kK/8 planes][kRowsA rows][16 B]
  uint8_t* sB = smem + (kK / 8) * kRowsA * 16;      // [kK/8 planes][kN rows][16 B]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  for (int i = threadIdx.x; i < kRowsA * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    *reinterpret_cast<__nv_bfloat16*>(sA + ((k >> 3) * kRowsA + r) * 16 + (k & 7) * 2) = a_full[i];
  }
  for (int i = threadIdx.x; i < kN * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    *reinterpret_cast<__nv_bfloat16*>(sB + ((k >> 3) * kN + r) * 16 + (k & 7) * 2) = b[i];
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc<128>(&tmem_base_s);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16_f32(kM, kN);
    const uint32_t a0 = ptx::smem_u32(sA) + row_shift * 16;
    const uint32_t b0 = ptx::smem_u32(sB);
    for (int ks = 0; ks < kK / 16; ++ks) {
      const uint64_t da = ptx::umma_desc_kmajor_noswz(a0 + ks * 2 * kRowsA * 16, kRowsA * 16, 128);
      const uint64_t db = ptx::umma_desc_kmajor_noswz(b0 + ks * 2 * kN * 16, kN * 16, 128);
      ptx::umma_bf16_ss(tmem_base, da, db, idesc, ks > 0);
    }
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after_sync();

  const int warp = threadIdx.x >> 5;
  const int row = threadIdx.x;  // TMEM lane = output row
#pragma unroll 1
  for (int c0 = 0; c0 < kN; c0 += 32) {
    uint32_t r[32];
    ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) d[row * kN + c0 + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<128>(tmem_base);
}

// Tensor-pipe rate from shared memory with the tower kernel's access pattern: per round, 9 displaced A
// descriptors x 8 K-steps x (128 x NDIM x 16) UMMAs, rotating over the TMEM accumulators.
template <int NDIM>
__global__ void __launch_bounds__(128)
bench_kernel(int iters, float* __restrict__ sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int kPlanes = 16;               // 128 input channels
  constexpr int kRows = 128 * 2 + 64;       // two M tiles + guard
  uint8_t* sA = smem;                       // [16][kRows][16]
  uint8_t* sB = smem + kPlanes * kRows * 16;  // 4 K-steps of weights: [8 planes][NDIM][16]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  for (int i = threadIdx.x; i < (kPlanes * kRows * 16 + 8 * NDIM * 16) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // bf16 ~0.0078
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tmem_base_s);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16_f32(128, NDIM);
    const uint32_t a0 = ptx::smem_u32(sA) + 8 * 16;
    const uint32_t b0 = ptx::smem_u32(sB);
    constexpr int kAcc = 512 / NDIM;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const int shift = (tap / 3 - 1) * 6 + (tap % 3 - 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t db = ptx::umma_desc_kmajor_noswz(b0 + (ks & 3) * 2 * NDIM * 16, NDIM * 16, 128);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint64_t da = ptx::umma_desc_kmajor_noswz(
                a0 + (t * 128 + shift) * 16 + ks * 2 * kRows * 16, kRows * 16, 128);
            ptx::umma_bf16_ss(tmem_base + ((it * 2 + t) % kAcc) * NDIM, da, db, idesc, (tap | ks) != 0);
          }
        }
      }
    }
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after_sync();
  uint32_t r[32];
  ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16), r);
  ptx::tmem_ld_wait();
  if (r[0] == 0x12345678u) sink[blockIdx.x] = __uint_as_float(r[1]);
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem_base);
}

}  // namespace ummatest
}  // namespace e55

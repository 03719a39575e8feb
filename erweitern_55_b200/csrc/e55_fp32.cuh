// fp32 CUDA-core path: the reference's arithmetic type (float32), used for tight parity against the
// oracle and as the on-device cross-check of the tensor-core path.  Activations NCHW fp32 in HBM,
// one launch per layer, BN already folded into (w, bias) by the host.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace e55 {
namespace fp32 {

constexpr int kSq = 25;       // 5x5 board (network.py:85)
constexpr int kPad = 7;       // board padded by one zero ring ('same' padding of a 3x3 conv)
constexpr int kPatch = 52;    // 49 padded squares rounded up to a float4 multiple
constexpr int kChunk = 8;     // input channels staged per step

// out[n][co][sq] = act( bias[co] + sum_{tap,ci} w[tap][ci][co] * in[n][ci][sq + tap] (+ res[n][co][sq]) )
// Restates Conv2D(3x3,'same') + folded BN (+Add) (+ReLU): network.py:92-95,136-144.
// blockDim.x = cout (one output channel per thread), PB positions per block.
template <int PB>
__global__ void __launch_bounds__(256)
conv3x3_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ res, float* __restrict__ out, int n, int cin, int cout, int relu) {
  __shared__ __align__(16) float s_in[PB][kChunk][kPatch];
  const int co = threadIdx.x;
  const int n0 = blockIdx.x * PB;

  for (int i = threadIdx.x; i < PB * kChunk * kPatch; i += blockDim.x) (&s_in[0][0][0])[i] = 0.f;

  float acc[PB][kSq];
#pragma unroll
  for (int p = 0; p < PB; ++p)
#pragma unroll
    for (int s = 0; s < kSq; ++s) acc[p][s] = 0.f;

  for (int c0 = 0; c0 < cin; c0 += kChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < PB * kChunk * kSq; i += blockDim.x) {
      const int p = i / (kChunk * kSq);
      const int c = (i / kSq) % kChunk;
      const int s = i % kSq;
      float v = 0.f;
      if (n0 + p < n && c0 + c < cin) v = in[(static_cast<size_t>(n0 + p) * cin + c0 + c) * kSq + s];
      s_in[p][c][(s / 5 + 1) * kPad + (s % 5) + 1] = v;
    }
    __syncthreads();
    const int cmax = min(kChunk, cin - c0);
    for (int c = 0; c < cmax; ++c) {
      float wt[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) wt[t] = w[(static_cast<size_t>(t) * cin + c0 + c) * cout + co];
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        float patch[kPatch];
        const float4* src = reinterpret_cast<const float4*>(&s_in[p][c][0]);
#pragma unroll
        for (int q = 0; q < kPatch / 4; ++q) {
          const float4 v = src[q];
          patch[4 * q + 0] = v.x; patch[4 * q + 1] = v.y; patch[4 * q + 2] = v.z; patch[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int y = 0; y < 5; ++y)
#pragma unroll
          for (int x = 0; x < 5; ++x) {
            float a = acc[p][y * 5 + x];
#pragma unroll
            for (int t = 0; t < 9; ++t) a = fmaf(wt[t], patch[(y + t / 3) * kPad + x + t % 3], a);
            acc[p][y * 5 + x] = a;
          }
      }
    }
  }

  const float b = bias[co];
#pragma unroll
  for (int p = 0; p < PB; ++p) {
    if (n0 + p >= n) break;
    const size_t base = (static_cast<size_t>(n0 + p) * cout + co) * kSq;
#pragma unroll
    for (int s = 0; s < kSq; ++s) {
      float v = acc[p][s] + b;
      if (res != nullptr) v += res[base + s];
      if (relu) v = fmaxf(v, 0.f);
      out[base + s] = v;
    }
  }
}

// Policy tail: Conv2D(1x1, cin -> 69) + bias, Flatten of NCHW (index c*25 + sq): network.py:106-108.
// One block per position; w is [cin][cout].
__global__ void __launch_bounds__(256)
policy_out_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ policy, int cin, int cout) {
  extern __shared__ float s_act[];  // [cin][25]
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < cin * kSq; i += blockDim.x) s_act[i] = in[static_cast<size_t>(n) * cin * kSq + i];
  __syncthreads();
  for (int o = threadIdx.x; o < cout * kSq; o += blockDim.x) {
    const int co = o / kSq, s = o % kSq;
    float a = bias[co];
    for (int c = 0; c < cin; ++c) a = fmaf(w[c * cout + co], s_act[c * kSq + s], a);
    policy[static_cast<size_t>(n) * cout * kSq + o] = a;
  }
}

// Value head: Conv2D(1x1 -> 1)+BN+ReLU, Flatten, Dense(25->128)+ReLU, Dense(128->1)+tanh
// (network.py:111-120).  One block of 128 threads per position.
__global__ void __launch_bounds__(128)
value_head_kernel(const float* __restrict__ in, const float* __restrict__ wv, float bv,
                  const float* __restrict__ k1, const float* __restrict__ b1,
                  const float* __restrict__ k2, float b2, float* __restrict__ value, int cin) {
  extern __shared__ float s_act[];  // [cin][25]
  __shared__ float s_v[kSq];
  __shared__ float s_part[4];
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < cin * kSq; i += blockDim.x) s_act[i] = in[static_cast<size_t>(n) * cin * kSq + i];
  __syncthreads();
  if (threadIdx.x < kSq) {
    float a = bv;
    for (int c = 0; c < cin; ++c) a = fmaf(wv[c], s_act[c * kSq + threadIdx.x], a);
    s_v[threadIdx.x] = fmaxf(a, 0.f);
  }
  __syncthreads();
  float h = b1[threadIdx.x];
#pragma unroll
  for (int s = 0; s < kSq; ++s) h = fmaf(s_v[s], k1[s * 128 + threadIdx.x], h);
  float part = fmaxf(h, 0.f) * k2[threadIdx.x];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) value[n] = tanhf(s_part[0] + s_part[1] + s_part[2] + s_part[3] + b2);
}

// Packed planes -> fp32 NCHW planes: x[n][c][sq] = bit sq of bits[n][c] ? scale[n][c] : 0.
__global__ void __launch_bounds__(256)
unpack_planes_kernel(const uint32_t* __restrict__ bits, const float* __restrict__ scale, float* __restrict__ planes,
                     size_t total /* n * in_planes * 25 */) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t nc = i / kSq;
    const int sq = static_cast<int>(i % kSq);
    planes[i] = ((bits[nc] >> sq) & 1u) ? scale[nc] : 0.f;
  }
}

// probs = softmax over legal entries of one policy row; illegal entries 0 (all-illegal row -> zeros).
// Warp-shuffle reductions for the row max and the row sum.  One block of 256 threads per position.
__global__ void __launch_bounds__(256)
masked_softmax_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ mask,
                      float* __restrict__ probs, int width) {
  __shared__ float s_red[8];
  __shared__ float s_bcast;
  const size_t base = static_cast<size_t>(blockIdx.x) * width;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  float m = -INFINITY;
  for (int i = threadIdx.x; i < width; i += blockDim.x)
    if (mask[base + i]) m = fmaxf(m, logits[base + i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = s_red[0];
    for (int i = 1; i < 8; ++i) mm = fmaxf(mm, s_red[i]);
    s_bcast = mm;
  }
  __syncthreads();
  const float rowmax = s_bcast;
  __syncthreads();

  float sum = 0.f;
  for (int i = threadIdx.x; i < width; i += blockDim.x)
    if (mask[base + i]) sum += __expf(logits[base + i] - rowmax);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_red[wid] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float ss = 0.f;
    for (int i = 0; i < 8; ++i) ss += s_red[i];
    s_bcast = ss;
  }
  __syncthreads();
  const float inv = s_bcast > 0.f ? 1.f / s_bcast : 0.f;
  for (int i = threadIdx.x; i < width; i += blockDim.x)
    probs[base + i] = mask[base + i] ? __expf(logits[base + i] - rowmax) * inv : 0.f;
}

}  // namespace fp32
// Legal-move policy tail in compact form: for position p the logits at index[offset[p] .. offset[p+1]) (the policy
// indices of its legal moves, at most a few dozen of the 1725) are soft-maxed and written to priors at the same
// offsets -- what minishogilib.MCTS.evaluate does with a policy row (mcts.py:60,105-106), without moving the other
// ~1700 logits to the host.  One warp per position, warp-shuffle max/sum reductions.
__global__ void __launch_bounds__(256)
legal_priors_kernel(const float* __restrict__ logits, const int32_t* __restrict__ offset, const uint16_t* __restrict__ index,
                    float* __restrict__ priors, int n, int width) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n) return;
  const int beg = offset[p], end = offset[p + 1];
  const float* row = logits + static_cast<size_t>(p) * width;
  float m = -INFINITY;
  for (int j = beg + lane; j < end; j += 32) m = fmaxf(m, row[index[j]]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int j = beg + lane; j < end; j += 32) sum += __expf(row[index[j]] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  for (int j = beg + lane; j < end; j += 32) priors[j] = __expf(row[index[j]] - m) * inv;
}

}  // namespace e55

// Tensor-core path: the whole network (stem, residual tower, policy conv, policy 1x1, value head) for
// a batch of positions in ONE persistent, warp-specialised sm_100a kernel running on CTA pairs.
//
// Restates network.py:87-146 with BatchNorm folded into bf16 weights.  Every 3x3 convolution is an
// implicit GEMM on tcgen05:  D[row, cout] += A[row + shift(tap), cin] * W[tap][cin, cout]
//
//   * rows  = squares of a "wide board": 4 positions side by side, each 5 squares + 1 zero column wide,
//             5 board rows -> row = y*24 + p*6 + x  (120 rows, one 128-row UMMA tile).  A 3x3 tap is then
//             the constant row displacement (dy*24 + dx); the zero columns and 25 zero guard rows
//             above/below supply the 'same' padding.  M efficiency = 100/128.
//   * A     = activations, bf16, resident in shared memory for the whole network, stored as 16 channel
//             planes [plane = c/8][row][c%8]: the no-swizzle K-major UMMA layout with SBO = 128 B, in
//             which rows are linear, so a tap is just start-address + shift*16 B (e55_selftest_umma).
//   * B     = weights, bf16, pre-packed on the host into the same layout per chunk (board row of taps x
//             pair of 32-channel K blocks) and streamed L2 -> smem by bulk TMA through a 5-stage mbarrier ring.
//   * D     = fp32 accumulators in TMEM (128 lanes = rows, 128 columns = output channels), two per group
//             (layer parity) so the next layer's MMAs start while the epilogue still drains the last one.
//
// CTA pair (cta_group::2).  A 128x128x16 SS-mode UMMA issued by one CTA reads 4 KB of A and 4 KB of B per
// 64 cycles = the full 128 B/clk of shared-memory bandwidth, so weight-ring writes and epilogue traffic
// would steal tensor throughput.  Two CTAs therefore pair up: each holds (and streams) only HALF of every
// weight chunk (64 of the 128 output channels) and the leader CTA issues M=256 UMMAs that cover one
// 128-row tile of each CTA: 96 B/clk per SM, half the L2->SM weight traffic, twice the ring depth.
//
// Each CTA evaluates 8 positions per round as two independent 4-position groups A and B (so a pair does
// 16 positions per round).  Both groups consume the SAME weight stream (a ring stage is released after
// both used it); the epilogue of one group (TMEM -> +bias (+residual) -> ReLU -> bf16 -> smem, in place)
// overlaps the other group's MMAs, and it is progressive: each finished 32-channel K block of the new
// activations is signalled separately and the next layer's chunks are ordered K-block-major.
// The residual operand never leaves registers: the thread that owns a row keeps its block input as packed
// bf16x2 registers.  The fp32 -> bf16 conversion of the network input is fused (epilogue warps).
//
// Warp roles per CTA (352 threads): warps 0-7 epilogue + input conversion, shared by both groups (TMEM
// lane quadrant = warp%4, 64-channel half = warp/4); warp 8 weight producer (bulk TMA, own half);
// warps 9 / 10: in the leader CTA the MMA issuers of group A / B, in the peer CTA warp 9 relays
// "my half of the stage has landed" to the leader.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "e55_ptx.cuh"

namespace e55 {
namespace tc {

// ---- geometry -----------------------------------------------------------------------------------
constexpr int kPosPerGroup = 4;
constexpr int kPosPerRound = 8;        // per CTA; a pair does 16
constexpr int kW = 24;                 // wide-board width
constexpr int kRowsValid = 120;        // 5 * kW
constexpr int kTile = 128;             // rows per CTA tile (UMMA M = 256 across the pair)
constexpr int kGuard = 25;             // kW + 1
constexpr int kValueHidden = 128;      // Dense(25 -> 128) of the value head (network.py:116-117)

// Per-tower-width geometry.  F = 128 (reference net): two 4-position groups per CTA.  F = 256 (scaled tower,
// BASELINE.json configs[4]): one group per CTA (the activations of a second one would not fit), N = 256.
template <int F>
struct Geo {
  static_assert(F == 128 || F == 256, "tower width must be 128 or 256");
  static constexpr int kGroups = F == 128 ? 2 : 1;
  static constexpr int kPlanes = F / 8;                                   // 8-channel activation planes
  static constexpr int kRowsTotal = kGuard + kGroups * (kTile + kGuard);  // 331 / 178
  static constexpr int kPlaneBytes = kRowsTotal * 16;
  static constexpr int kActBytes = kPlanes * kPlaneBytes;                 // 84,736 / 91,136
  static constexpr int kKBlocks = F / 32;                                 // 32-channel activation K blocks
  static constexpr int kHalfCh = F / 2;                                   // channels per epilogue thread
  static constexpr int kCb = kHalfCh / 32;                                // 32-column TMEM loads per thread
  static constexpr bool kBiasInSmem = F == 128;                           // 43 x 256 biases do not fit
};
constexpr int kMaxStages = 5;
constexpr int kStageBytes = 24576;     // this CTA's half of a chunk: 3 taps x 2 K blocks x 2 K-steps x 64 x 16 x 2 B
constexpr int kPolicyCh = 69;
constexpr int kPolicyN = 96;           // policy 1x1 output channels padded (each CTA holds N/2 = 48 columns)
constexpr int kSq = 25;
constexpr int kThreads = 352;
constexpr int kTmemCols = 512;         // groups x 2 accumulators (layer parity) x F columns
constexpr uint32_t kDescHi = 8u | (1u << 14);  // SBO = 128 B, descriptor version 1, no swizzle

__host__ __device__ constexpr int group_origin_row(int g) { return kGuard + g * (kTile + kGuard); }

enum : uint32_t { kFirstOfSlice = 1, kLastOfSlice = 2, kFirstOfLayer = 4, kLastOfLayer = 8, kSingleTap = 16 };
enum : uint32_t { kRelu = 1, kHasRes = 2, kUpdRes = 4, kValueTap = 8, kFinal = 16 };

// One ring stage worth of weights: the taps of one board row (dx = -1,0,1; or the single tap of a 1x1
// layer) for a pair of 32-channel K blocks of one input slice.  The epilogue finishes K blocks 0 and 2
// first (first 32 channels of each 64-channel half), then 1 and 3, so a layer's chunks come in that
// order: {0,2} x 3 board rows, then {1,3} x 3 board rows; each chunk only needs "its" K blocks.
// In the weight image the two 64-column halves of a chunk follow each other (CTA rank 0, then rank 1).
struct alignas(16) Chunk {
  uint32_t gofs;     // byte offset into the weight image
  uint16_t half16;   // bytes of one CTA's half, in 16 B units
  uint16_t lend;     // chunk index one past the end of this chunk's layer
  int16_t a_off;     // A-descriptor displacement in 16 B units: first plane * rows + dy * board width
  uint8_t nk0;       // K=16 steps per tap on the first K block (planes a_off..)
  uint8_t nk1;       // K=16 steps per tap on the second K block, half the planes further on (0 = none)
  uint8_t need;      // bit kb: activation K block kb of the previous layer must be written before this chunk
  uint8_t ndim8;     // UMMA N / 8 of the layer
  uint8_t flags;     // bits 0-3: first/last of slice/layer; bit 4: single tap (1x1 layer) instead of dx = -1,0,1
  uint8_t layer;
};

struct LayerInfo {
  int chunk_begin, chunk_end;
  int ndim;          // UMMA N (host-side bookkeeping; the kernel reads Chunk::ndim8)
  uint32_t flags;
};

// Which input slices the epilogue warps have written into the four stem-input buffers b = group * 2 + plane half, as much
// as the barrier protocol needs of it: before slice number k >= 1 of a buffer is written, the MMAs that read slice k - 1
// must have released it, which is phase k - 1 of the buffer's `bar_in_empty` barrier, i.e. a wait on parity (k - 1) & 1.
// One register (bit b: slices written so far is odd; bit 4 + b: there was one) instead of four counters: nothing that can
// overflow in a launch of any length, and no array for the compiler to put into local memory (section 4.1 of DESIGN.md).
// Host-callable so that `e55_selftest_slice_book` can hold it to plain counters without a GPU.
struct SliceBook {
  uint32_t state = 0u;
  __host__ __device__ __forceinline__ bool none_yet(int b) const { return ((state >> (4 + b)) & 1u) == 0u; }
  __host__ __device__ __forceinline__ uint32_t release_parity(int b) const { return ((state >> b) & 1u) ^ 1u; }
  __host__ __device__ __forceinline__ void wrote(int b) { state = (state ^ (1u << b)) | (16u << b); }
};

struct Params {
  const uint8_t* wimg;
  const Chunk* chunks;
  const LayerInfo* layers;
  const float* bias;       // [nlayers][F]
  const float* value_w;    // [F]
  const float* fc1_k;      // [25][128]
  const float* fc1_b;      // [128]
  const float* fc2_k;      // [128]
  float value_b, fc2_b;
  const float* planes;       // [n][in_planes][25] fp32 NCHW, as Network.get_inputs builds them (or nullptr)
  const uint32_t* bits;      // packed input instead of `planes`: [n][in_planes] occupancy bit per square ...
  const float* scale;        // ... and [n][in_planes] value of the set squares of that plane
  int packed_stride;         // words from one position's bits (scale) to the next one's: in_planes, or 2 * in_planes when interleaved
  int in_planes;
  int in_planes8;            // ceil(in_planes / 16) * 2: 8-channel planes incl. zero padding
  int nslices;
  int nstages;               // weight ring depth
  int lag;                   // chunks group A is kept ahead of group B (0 = free running)
  int nchunks_total;         // chunks of the full network (size of the shared-memory chunk table)
  int nlayers;               // layers of the full network (tower convs + policy conv + policy 1x1)
  int nlayers_run;           // == nlayers, or fewer for a debug dump
  int n;                     // positions
  int nrounds;               // ceil(n / pos_per_round)
  int pos_per_round;         // 8 (groups A and B), or 4 (group A only: latency mode for small batches)
  float* policy;             // [n][1725] raw logits (device memory or mapped page-locked host memory), or nullptr
  float* value;              // [n]
  // fused legal-move tail (mcts.py:60,105-106: softmax over the legal moves' logits), all optional:
  const int32_t* move_off;   // CSR offsets [n+1] (move_stride == 0), or counts [n] (move_stride > 0), or nullptr
  const uint16_t* move_idx;  // policy indices of the legal moves: [move_off[i], move_off[i+1]) or [i*stride, +count[i])
  float* priors;             // softmax over exactly those logits, same indexing as move_idx
  int move_stride;
  const uint8_t* mask;       // dense form: [n][1725] legal mask -> probs [n][1725] (illegal entries 0), or nullptr
  float* probs;
  int tail;                  // 1: the logits are staged in shared memory and leave through the policy tail (priors / probs wanted,
                             // or `policy` is to be written as whole rows: mapped host memory); 0: the epilogue warps write `policy` directly
  uint4* dbg;                // [rounds*2][16][128] packed activations after layer nlayers_run-1
  unsigned long long* prof;  // optional cycle counters / timeline of CTA 0 (kProf instantiation only)
};

// debug: packed activations [group][16][128 rows][8] bf16 -> fp32 NCHW [n][128][25]
__global__ void unpack_debug_kernel(const uint4* __restrict__ dbg, float* __restrict__ out, int n, int filters) {
  const int pos = blockIdx.x;
  const int group = pos / kPosPerGroup, p = pos % kPosPerGroup;
  const int kPlanes = filters / 8;
  for (int idx = threadIdx.x; idx < kPlanes * kSq; idx += blockDim.x) {
    const int pl = idx / kSq, sq = idx % kSq;
    const int r = (sq / 5) * kW + p * 6 + sq % 5;
    const uint4 v = dbg[(static_cast<size_t>(group) * kPlanes + pl) * kTile + r];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t bits = (e & 1) ? (w[e >> 1] & 0xFFFF0000u) : (w[e >> 1] << 16);
      out[(static_cast<size_t>(pos) * filters + pl * 8 + e) * kSq + sq] = __uint_as_float(bits);
    }
  }
}

__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// D[tmem] (+)= A * B^T across the CTA pair (M = 256), descriptors given as (lo, hi) words.
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// arrive on the leader CTA's copy of `bar` (plain local arrive when this CTA is the leader)
__device__ __forceinline__ void arrive_on_leader(uint64_t* bar, uint32_t rank) {
  if (rank == 0) ptx::mbar_arrive(bar);
  else ptx::mbar_arrive_cluster(bar, 0);
}

// Epilogue of one half (F/2 channels) of a 128 x F accumulator tile for the thread that owns `row`:
// TMEM -> +bias (+residual) -> ReLU -> bf16 -> shared memory (in place: the next layer's A operand).
// taddr / bias / act_row / s_wv / dbg_row / bar_part already point at this thread's half.
template <int F, bool kRes, bool kUpd, bool kVal>
__device__ __forceinline__ float epilogue_half(uint32_t taddr, const float* __restrict__ bias, uint8_t* act_row,
                                               uint32_t (&res)[F / 4], const float* __restrict__ s_wv,
                                               bool row_valid, uint4* dbg_row, uint64_t* bar_part, int lane,
                                               uint32_t rank) {
  using G = Geo<F>;
  float vacc = 0.f;
  uint32_t v[2][32];
  ptx::tmem_ld_32x32b_x32(taddr, v[0]);
#pragma unroll
  for (int cb = 0; cb < G::kCb; ++cb) {
    ptx::tmem_ld_wait();
    if (cb + 1 < G::kCb) ptx::tmem_ld_32x32b_x32(taddr + (cb + 1) * 32, v[(cb + 1) & 1]);  // overlaps the math below
    const uint32_t(&vv)[32] = v[cb & 1];
    uint32_t o[16];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + cb * 32 + j4 * 4);
      float a0 = __uint_as_float(vv[4 * j4 + 0]) + bb.x;
      float a1 = __uint_as_float(vv[4 * j4 + 1]) + bb.y;
      float a2 = __uint_as_float(vv[4 * j4 + 2]) + bb.z;
      float a3 = __uint_as_float(vv[4 * j4 + 3]) + bb.w;
      if (kRes) {
        const uint32_t r0 = res[cb * 16 + 2 * j4], r1 = res[cb * 16 + 2 * j4 + 1];
        a0 += bf16lo(r0); a1 += bf16hi(r0); a2 += bf16lo(r1); a3 += bf16hi(r1);
      }
      const uint32_t p0 = ptx::pack_relu_bf16x2(a0, a1), p1 = ptx::pack_relu_bf16x2(a2, a3);
      o[2 * j4] = p0; o[2 * j4 + 1] = p1;
      if (kUpd) { res[cb * 16 + 2 * j4] = p0; res[cb * 16 + 2 * j4 + 1] = p1; }
      if (kVal) {
        const float4 wv = *reinterpret_cast<const float4*>(s_wv + cb * 32 + j4 * 4);
        vacc += wv.x * bf16lo(p0) + wv.y * bf16hi(p0) + wv.z * bf16lo(p1) + wv.w * bf16hi(p1);
      }
    }
    if (row_valid) {
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const uint4 pk = make_uint4(o[4 * j4], o[4 * j4 + 1], o[4 * j4 + 2], o[4 * j4 + 3]);
        *reinterpret_cast<uint4*>(act_row + (cb * 4 + j4) * G::kPlaneBytes) = pk;
        if (dbg_row) dbg_row[(cb * 4 + j4) * kTile] = pk;
      }
    }
    // this 32-channel K block of the new activations is complete: tell the leader's MMA issuer
    ptx::tc_fence_before_sync();
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) arrive_on_leader(&bar_part[cb], rank);
  }
  return vacc;
}

// Policy tail of one 4-position group, run by the weight-producer warp once all eight epilogue warps have staged the
// group's logits in its (by then idle) activation planes: the logits leave as one coalesced 27.6 KB piece (device memory
// or mapped page-locked host memory) and/or the legal moves of each position are soft-maxed with warp shuffles -- what
// minishogilib.MCTS.evaluate does with a policy row (mcts.py:60,105-106).  Then the planes go back to the next round's
// input: the staging has overwritten the zero columns / rows that give the convolutions their padding, so those are
// cleared first.
// Staging layout of a group's logits in its activation planes: five policy channels per plane strip (128 rows x 16 B),
// each a row of [4 positions][25 squares] floats -- for an epilogue thread (one position, one square) the channel is
// the only variable, and its offset a compile-time constant of the unrolled loop.
template <int F>
__host__ __device__ constexpr int stage_offset(int c) { return (c / 5) * Geo<F>::kPlaneBytes + (c % 5) * (kPosPerGroup * kSq * 4); }

template <int F>
__device__ __forceinline__ void policy_tail(const Params& prm, uint8_t* s_act, uint64_t* bar_in_empty, int g, int pos0, int lane) {
  constexpr int kPlaneBytes = Geo<F>::kPlaneBytes;
  constexpr int kRow = kPolicyCh * kSq;
  uint8_t* stage = s_act + group_origin_row(g) * 16;
  auto staged = [&](int L) -> float {   // L = p * 1725 + c * 25 + sq
    const int p = L / kRow, j = L - p * kRow, c = j / kSq, sq = j - c * kSq;
    return *reinterpret_cast<const float*>(stage + stage_offset<F>(c) + (p * kSq + sq) * 4);
  };
  const int npos = min(kPosPerGroup, prm.n - pos0);   // <= 0: the whole group lies past the batch
  if (npos == kPosPerGroup && prm.policy != nullptr && !(prm.tail & 2)) {
    // rows of consecutive positions are consecutive in memory, and a full group's four rows (27 600 bytes from a
    // position index that is a multiple of four) start and end on 16-byte boundaries: 1 725 float4 stores, 512 bytes per
    // warp instruction -- when the destination is mapped host memory the size of the PCIe write transactions follows.
    // The lane's first element advances by 128 floats per step: (position, channel, square) are carried, not divided out.
    float4* dst = reinterpret_cast<float4*>(prm.policy + static_cast<size_t>(pos0) * kRow);
    int e0 = lane * 4;                                   // first of the lane's four consecutive elements
    int c = e0 / kSq, sq = e0 - c * kSq, p = 0;          // (e0 < 128 < 1725: position 0)
    for (int v = lane; v < kPosPerGroup * kRow / 4; v += 32) {
      float x[4];
      int cc = c, ss = sq, pp = p;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        x[k] = *reinterpret_cast<const float*>(stage + stage_offset<F>(cc) + (pp * kSq + ss) * 4);
        if (++ss == kSq) { ss = 0; if (++cc == kPolicyCh) { cc = 0; ++pp; } }
      }
      dst[v] = make_float4(x[0], x[1], x[2], x[3]);
      sq += 128 - 5 * kSq; c += 5;                       // + 128 elements = 5 channels and 3 squares
      if (sq >= kSq) { sq -= kSq; c += 1; }
      if (c >= kPolicyCh) { c -= kPolicyCh; p += 1; }
    }
  } else if (npos > 0 && prm.policy != nullptr) {
    // a ragged last group: the same piece, element by element
    float* dst = prm.policy + static_cast<size_t>(pos0) * kRow;
    int c = lane / kSq, sq = lane - c * kSq, p = 0;
    for (int i = lane; i < npos * kRow; i += 32) {
      dst[i] = *reinterpret_cast<const float*>(stage + stage_offset<F>(c) + (p * kSq + sq) * 4);
      sq += 32 - kSq; c += 1;
      if (sq >= kSq) { sq -= kSq; c += 1; }
      if (c >= kPolicyCh) { c -= kPolicyCh; p += 1; }
    }
  }
  if (npos > 0 && prm.priors != nullptr) {
    for (int p = 0; p < npos; ++p) {
      const int pos = pos0 + p, base = p * kRow;
      int beg, end;
      if (prm.move_stride > 0) { beg = pos * prm.move_stride; end = beg + max(prm.move_off[pos], 0); }
      else { beg = prm.move_off[pos]; end = prm.move_off[pos + 1]; }
      float m = -INFINITY;
      for (int j = beg + lane; j < end; j += 32) m = fmaxf(m, staged(base + prm.move_idx[j]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.f;
      for (int j = beg + lane; j < end; j += 32) sum += __expf(staged(base + prm.move_idx[j]) - m);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      for (int j = beg + lane; j < end; j += 32) prm.priors[j] = __expf(staged(base + prm.move_idx[j]) - m) * inv;
    }
  }
  if (npos > 0 && prm.probs != nullptr) {
    for (int p = 0; p < npos; ++p) {
      const int base = p * kRow;
      const uint8_t* mk = prm.mask + static_cast<size_t>(pos0 + p) * kRow;
      float* out = prm.probs + static_cast<size_t>(pos0 + p) * kRow;
      float m = -INFINITY;
      for (int i = lane; i < kRow; i += 32) if (mk[i]) m = fmaxf(m, staged(base + i));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.f;
      for (int i = lane; i < kRow; i += 32) if (mk[i]) sum += __expf(staged(base + i) - m);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      for (int i = lane; i < kRow; i += 32) out[i] = mk[i] ? __expf(staged(base + i) - m) * inv : 0.f;
    }
  }
  __syncwarp();
  constexpr int kStripsUsed = (kPolicyCh + 4) / 5;                       // 14 of the planes hold staged logits
  constexpr int kPadRows = kRowsValid / 6 + (kTile - kRowsValid);        // 20 sixth-column rows + 8 rows past the board
  for (int i = lane; i < kStripsUsed * kPadRows; i += 32) {
    const int strip = i / kPadRows, k = i % kPadRows;
    const int prow = k < kRowsValid / 6 ? k * 6 + 5 : kRowsValid + (k - kRowsValid / 6);
    *reinterpret_cast<uint4*>(stage + strip * kPlaneBytes + prow * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  ptx::fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    ptx::mbar_arrive(&bar_in_empty[g * 2 + 0]);
    ptx::mbar_arrive(&bar_in_empty[g * 2 + 1]);
  }
}

// Packed network input is small (2 128 bytes per position): the positions of a CTA's round fit the little L1 that is left
// beside 220 KB of shared memory, so they are pulled into it before the conversion loops ask for them word by word.
__device__ __forceinline__ void prefetch_packed_l1(const Params& prm, int pos0, int npos, int tid, int nthreads) {
  const int last = min(pos0 + npos, prm.n);
  if (pos0 >= last) return;
  const bool interleaved = prm.scale == reinterpret_cast<const float*>(prm.bits) + prm.in_planes;   // [bits | scale] per position
  const char* b0 = reinterpret_cast<const char*>(prm.bits + static_cast<size_t>(pos0) * prm.packed_stride);
  const size_t bytes = static_cast<size_t>(last - pos0) * prm.packed_stride * 4 - (interleaved ? 0 : (prm.packed_stride - prm.in_planes) * 4);
  for (size_t o = static_cast<size_t>(tid) * 128; o < bytes; o += static_cast<size_t>(nthreads) * 128)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(b0 + o));
  if (!interleaved) {
    const char* s0 = reinterpret_cast<const char*>(prm.scale + static_cast<size_t>(pos0) * prm.packed_stride);
    for (size_t o = static_cast<size_t>(tid) * 128; o < bytes; o += static_cast<size_t>(nthreads) * 128)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(s0 + o));
  }
}

struct SmemLayout {
  static constexpr int ring = 0;
  __host__ __device__ static constexpr int act(int nstages) { return ring + nstages * kStageBytes; }
};

template <int F, bool kProf>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tower_kernel(const Params prm) {
  using G = Geo<F>;
  constexpr int kPlanes = G::kPlanes;
  constexpr int kPlaneBytes = G::kPlaneBytes;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_ring = smem + SmemLayout::ring;
  uint8_t* s_act = smem + SmemLayout::act(prm.nstages);
  float* s_bias = reinterpret_cast<float*>(s_act + G::kActBytes);   // [nlayers][F] when it fits (F = 128)
  float* s_wv = s_bias + (G::kBiasInSmem ? prm.nlayers * F : 0);
  float* s_v = s_wv + F;                         // [2 groups][2 halves][4 positions][25]
  Chunk* s_chunks = reinterpret_cast<Chunk*>(s_v + 4 * kPosPerGroup * kSq);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_chunks + prm.nchunks_total);
  const uint32_t nstages = prm.nstages;
  // identical layout in both CTAs (multicast commits and remote arrives address barriers by offset)
  uint64_t* bar_full = bars;                          // [stages] stage landed (leader: both CTAs' halves)
  uint64_t* bar_empty = bars + kMaxStages;            // [stages] both groups' MMAs on the stage are done
  uint64_t* bar_tmem_full = bars + 2 * kMaxStages;    // [2] accumulator of a group complete
  uint64_t* bar_act_part = bar_tmem_full + 2;         // [8] leader only: activation K block (group*4 + kb) written by both CTAs
  uint64_t* bar_in_full = bar_act_part + 8;           // [2 groups][2 plane halves] leader only: stem input slice written
  uint64_t* bar_in_empty = bar_in_full + 4;           // [2][2] that half of the planes may be overwritten
  uint64_t* bar_v = bar_in_empty + 4;                 // [2] value-conv partials of a group written by all 8 warps
  uint64_t* bar_p = bar_v + 2;                        // [2] logits of a group staged in shared memory by all 8 warps
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_p + 2);
  volatile uint32_t* s_apos = s_tmem + 1;             // chunks issued so far by group A's issuer

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int nl = prm.nlayers_run;
  const uint32_t nchunks = prm.layers[nl - 1].chunk_end;
  const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;
  const int npair_rounds = (prm.nrounds + 1) >> 1;    // a pair-round = one round in each CTA

  long long t_entry = 0;
  if (kProf) t_entry = clock64();
  // ---- one-time setup -----------------------------------------------------------------------
  for (int i = threadIdx.x; i < G::kActBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(s_act)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (G::kBiasInSmem)
    for (int i = threadIdx.x; i < prm.nlayers * F; i += kThreads) s_bias[i] = prm.bias[i];
  for (int i = threadIdx.x; i < F; i += kThreads) s_wv[i] = prm.value_w[i];
  const float* bias_base = G::kBiasInSmem ? s_bias : prm.bias;
  for (int i = threadIdx.x; i < prm.nchunks_total; i += kThreads) s_chunks[i] = prm.chunks[i];
  if (threadIdx.x == 0) {
    *s_apos = 0u;
    for (uint32_t s = 0; s < nstages; ++s) {
      ptx::mbar_init(&bar_full[s], rank == 0 ? 2 : 1);   // leader: own TMA + the peer's relay
      ptx::mbar_init(&bar_empty[s], 2);
    }
    for (int kb = 0; kb < 8; ++kb) ptx::mbar_init(&bar_act_part[kb], 8);               // 4 warps x 2 CTAs
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(&bar_tmem_full[g], 1);
      for (int hh = 0; hh < 2; ++hh) {
        ptx::mbar_init(&bar_in_full[g * 2 + hh], 16);                                // 8 warps x 2 CTAs
        ptx::mbar_init(&bar_in_empty[g * 2 + hh], 1);
      }
      ptx::mbar_init(&bar_v[g], 8);
      ptx::mbar_init(&bar_p[g], 8);
    }
    ptx::fence_mbar_init();
  }
  if (prm.planes != nullptr) {  // pull this CTA's first round of network input into L2 while the setup completes
    const int r0 = pair0 * 2 + static_cast<int>(rank);
    const size_t first = static_cast<size_t>(r0) * prm.pos_per_round * prm.in_planes * kSq;
    const size_t last = min(static_cast<size_t>(r0 + 1) * prm.pos_per_round, static_cast<size_t>(prm.n)) * prm.in_planes * kSq;
    for (size_t i = first + threadIdx.x * 32; i < last; i += kThreads * 32)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.planes + i));
  }
  if (prm.bits != nullptr) prefetch_packed_l1(prm, (pair0 * 2 + static_cast<int>(rank)) * prm.pos_per_round, prm.pos_per_round, threadIdx.x, kThreads);
  if (warp == 9) ptx::tmem_alloc_2sm<kTmemCols>(s_tmem);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();   // the peer's barriers are initialised before anyone arrives on them remotely
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  if (kProf && blockIdx.x == 0 && threadIdx.x == 0) { prm.prof[12] = t_entry; prm.prof[13] = clock64(); }

  if (warp == 8) {
    // ===== weight producer: this CTA's half (64 output channels) of every chunk; then, with the round's last chunk on
    // its way, the policy tails of the round's groups (this warp holds no epilogue state, so the tail costs the epilogue
    // warps neither registers nor time) ==================
    uint32_t s = 0, ph = 1, round_seq = 0;
    long long t_wait = 0;
    for (int pr = pair0; pr < npair_rounds; pr += pair_stride, ++round_seq) {
      if (lane == 0) {
        for (uint32_t c = 0; c < nchunks; ++c) {
          const Chunk ch = s_chunks[c];
          long long t0 = 0;
          if (kProf) t0 = clock64();
          ptx::mbar_wait(&bar_empty[s], ph);
          if (kProf) t_wait += clock64() - t0;
          const uint32_t bytes = static_cast<uint32_t>(ch.half16) * 16u;
          ptx::mbar_arrive_expect_tx(&bar_full[s], bytes);
          ptx::tma_bulk_g2s(s_ring + s * kStageBytes, prm.wimg + ch.gofs + rank * bytes, bytes, &bar_full[s]);
          if (++s == nstages) { s = 0; ph ^= 1; }
        }
      }
      __syncwarp();
      if (prm.tail && nl == prm.nlayers) {
        const int r = pr * 2 + static_cast<int>(rank);
        const bool has_b = G::kGroups == 2 && prm.pos_per_round == kPosPerRound && pr * 2 * kPosPerRound + kPosPerGroup < prm.n;
        for (int g = 0; g < (has_b ? 2 : 1); ++g) {
          ptx::mbar_wait(&bar_p[g], round_seq & 1);
          policy_tail<F>(prm, s_act, bar_in_empty, g, r * prm.pos_per_round + g * kPosPerGroup, lane);
        }
      }
    }
    if (kProf && blockIdx.x == 0 && lane == 0) prm.prof[8] = t_wait;
  } else if (warp == 9 && rank == 1) {
    // ===== peer CTA: relay "my half of stage s has landed" to the leader's issuers ==================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int pr = pair0; pr < npair_rounds; pr += pair_stride) {
        for (uint32_t c = 0; c < nchunks; ++c) {
          ptx::mbar_wait(&bar_full[s], ph);
          ptx::mbar_arrive_cluster(&bar_full[s], 0);
          if (++s == nstages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if ((warp == 9 || (warp == 10 && G::kGroups == 2)) && rank == 0) {
    // ===== MMA issuers (leader CTA): warp 9 drives group A, warp 10 group B, for BOTH CTAs' tiles
    // (cta_group::2, M = 256).  Warp-uniform loop, one elected lane issues.  Both walk the same chunk
    // sequence; a ring stage is released (in both CTAs) after both groups used it. ==================
    const int g = warp == 9 ? 0 : 1;
    const uint32_t a_lo_base = ((ptx::smem_u32(s_act) + group_origin_row(g) * 16) >> 4) |
                               (static_cast<uint32_t>(kPlaneBytes >> 4) << 16);
    const uint32_t b_lo_base = ptx::smem_u32(s_ring) >> 4;
    uint32_t s = 0, ph = 0, round_seq = 0, stem_idx = 0, stem_full0 = 0u, stem_full1 = 0u;   // (scalars: no local-memory arrays)
    long long t_start = 0, t_issue = 0, t_full = 0, t_act = 0;
    if (kProf) t_start = clock64();
    for (int pr = pair0; pr < npair_rounds; pr += pair_stride, ++round_seq) {
      const bool has_b = G::kGroups == 2 && prm.pos_per_round == kPosPerRound && pr * 2 * kPosPerRound + kPosPerGroup < prm.n;
      if (g == 1 && !has_b) break;  // only the last pair-round can lack group B
      const uint32_t gbase = round_seq * nchunks;
      for (uint32_t c = 0; c < nchunks; ++c) {
        const Chunk ch = s_chunks[c];
        long long t0 = 0, t1 = 0, t2 = 0;
        if (kProf) t0 = clock64();
        if (g == 1 && prm.lag > 0) {
          // Keep group A `lag` chunks ahead so that one group's epilogue overlaps the other group's
          // MMAs; once A has finished B's current layer B runs freely to the end of that layer.
          const uint32_t want = min(gbase + c + 1 + prm.lag, gbase + static_cast<uint32_t>(ch.lend));
          while (*s_apos < want) {
          }
        }
        ptx::mbar_wait_cluster(&bar_full[s], ph);
        if (kProf) t1 = clock64();
        const uint32_t epoch = round_seq * nl + ch.layer;
        if ((ch.flags & kFirstOfSlice) && ch.layer == 0) {
          // stem input slices alternate between the two halves of the activation planes
          if (ch.flags & kFirstOfLayer) stem_idx = 0;
          const uint32_t hh = stem_idx & 1;
          ptx::mbar_wait_cluster(&bar_in_full[g * 2 + hh], (hh ? stem_full1 : stem_full0) & 1);
          if (hh) ++stem_full1; else ++stem_full0;
        }
        if (ch.need && epoch > 0) {
          // activation K blocks written by the previous layer's (progressive) epilogue in both CTAs
#pragma unroll
          for (int kb = 0; kb < G::kKBlocks; ++kb)
            if (ch.need & (1u << kb)) ptx::mbar_wait_cluster(&bar_act_part[g * 4 + kb], (epoch - 1) & 1);
        }
        if (kProf) t2 = clock64();
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {
          const uint32_t ndim = ch.ndim8 * 8u;
          const uint32_t idesc = ptx::umma_idesc_bf16_f32(2 * kTile, ndim);
          const uint32_t d = tmem_base + (g * 2 + (ch.layer & 1)) * F;
          const int ntaps = (ch.flags & kSingleTap) ? 1 : 3;
          const uint32_t a_chunk = a_lo_base + static_cast<int>(ch.a_off) - (ntaps == 3 ? 1 : 0);
          const uint32_t b_step = ndim;                          // 2 planes x (ndim/2 rows) x 16 B, in 16 B units
          uint32_t b_lo = (b_lo_base + s * (kStageBytes >> 4)) | ((ndim >> 1) << 16);   // LBO = (ndim/2) x 16 B
          uint32_t acc = (ch.flags & kFirstOfLayer) ? 0u : 1u;
          constexpr uint32_t kStepK = 2 * (kPlaneBytes >> 4);    // one K=16 step = 2 planes
          constexpr uint32_t kBlock2 = (kPlanes / 2) * (kPlaneBytes >> 4);   // second K block: half the planes on
          if (ntaps == 3 && ch.nk0 == 2 && ch.nk1 == 2) {
#pragma unroll
            for (int t = 0; t < 3; ++t)
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                umma_pair(d, a_chunk + t + (k4 >> 1) * kBlock2 + (k4 & 1) * kStepK, b_lo, idesc, acc);
                acc = 1u;
                b_lo += b_step;
              }
          } else if (ntaps == 3 && ch.nk0 == 2 && ch.nk1 == 0) {
#pragma unroll
            for (int t = 0; t < 3; ++t)
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                umma_pair(d, a_chunk + t + ks * kStepK, b_lo, idesc, acc);
                acc = 1u;
                b_lo += b_step;
              }
          } else {
            for (int t = 0; t < ntaps; ++t) {
              for (int ks = 0; ks < ch.nk0; ++ks) {
                umma_pair(d, a_chunk + t + ks * kStepK, b_lo, idesc, acc);
                acc = 1u;
                b_lo += b_step;
              }
              for (int ks = 0; ks < ch.nk1; ++ks) {
                umma_pair(d, a_chunk + t + kBlock2 + ks * kStepK, b_lo, idesc, acc);
                acc = 1u;
                b_lo += b_step;
              }
            }
          }
          ptx::umma_commit_2sm(&bar_empty[s], 3);
          if (!has_b) ptx::umma_commit_2sm(&bar_empty[s], 3);
          if ((ch.flags & kLastOfSlice) && ch.layer == 0 && stem_idx + 2 < static_cast<uint32_t>(prm.nslices))
            ptx::umma_commit_2sm(&bar_in_empty[g * 2 + (stem_idx & 1)], 3);   // a later slice reuses this half
          if (ch.flags & kLastOfLayer) {
            ptx::umma_commit_2sm(&bar_tmem_full[g], 3);
            if (ch.layer == nl - 1 && !(prm.tail && nl == prm.nlayers)) {   // round end: both halves are free for the next
              ptx::umma_commit_2sm(&bar_in_empty[g * 2 + 0], 3);              // round's input (when the policy tail runs, it
              ptx::umma_commit_2sm(&bar_in_empty[g * 2 + 1], 3);              // stages the logits in these planes and releases them)
            }
          }
        }
        if ((ch.flags & kLastOfSlice) && ch.layer == 0) ++stem_idx;
        __syncwarp();
        if (g == 0 && lane == 0) *s_apos = gbase + c + 1;
        if (++s == nstages) { s = 0; ph ^= 1; }
        if (kProf) {
          const long long t3 = clock64();
          t_full += t1 - t0; t_act += t2 - t1; t_issue += t3 - t2;
          if (blockIdx.x == 0 && lane == 0 && round_seq == 0 && c < 1024) {
            unsigned long long* tr = prm.prof + 16 + g * 4096 + c * 4;
            tr[0] = t0; tr[1] = t1; tr[2] = t2; tr[3] = t3;
          }
        }
      }
    }
    if (kProf && blockIdx.x == 0 && lane == 0) {
      unsigned long long* o = prm.prof + (g ? 4 : 0);
      o[0] = clock64() - t_start; o[1] = t_issue; o[2] = t_full; o[3] = t_act;
      if (g == 0) prm.prof[9] = t_start;
    }
  } else if (warp < 8) {
    // ===== epilogue warps 0-7, shared by both groups: warp w owns TMEM lane quadrant w%4 (32 tile rows)
    // and the 64-channel half w/4 of whichever group's accumulator is complete; they also convert the
    // fp32 network input into the bf16 wide-board planes (stem input slices). =======================
    const int q = warp & 3;
    const int h = warp >> 2;
    const int row = q * 32 + lane;                       // TMEM lane == tile row
    const int by = row / kW, bp = (row % kW) / 6, bx = row % 6;
    const bool row_valid = row < kRowsValid && bx < 5;
    const int sq = by * 5 + bx;
    uint32_t res_a[F / 4], res_b[G::kGroups == 2 ? F / 4 : 1];   // block inputs of this row/half, packed bf16x2
#pragma unroll
    for (int i = 0; i < F / 4; ++i) res_a[i] = 0u;
#pragma unroll
    for (int i = 0; i < (G::kGroups == 2 ? F / 4 : 1); ++i) res_b[i] = 0u;
    uint32_t round_seq = 0;
    // input slices written so far per (group, plane half): one register of parity/seen bits (a launch of 131072 positions
    // on 124 SMs converts ~400 slices per buffer).  Not an array: a dynamically indexed array lives in local memory, and
    // during the input phase, with the L1 full of streaming loads, every access to it was an ~600-cycle miss
    // (tools/tc_trace.py).
    SliceBook slices;
    long long t_wait = 0, t_work = 0;

    for (int pr = pair0; pr < npair_rounds; pr += pair_stride, ++round_seq) {
      const int r = pr * 2 + static_cast<int>(rank);     // this CTA's round (may lie past the batch: all absent)
      const bool has_b = G::kGroups == 2 && prm.pos_per_round == kPosPerRound && pr * 2 * kPosPerRound + kPosPerGroup < prm.n;
      int next_layer0 = 0, next_layer1 = has_b ? 0 : nl;
      int next_slice0 = 0, next_slice1 = has_b ? 0 : prm.nslices;
      // the value-head tail of group g is run by warp 4g once all 8 warps have written their partials
      const bool do_value = nl == prm.nlayers;
      bool fc_pending0 = do_value && warp == 0, fc_pending1 = do_value && warp == 4 && has_b;
      if (prm.planes != nullptr && pr + pair_stride < npair_rounds) {  // pull the next round's network input into L2
        const int rn = (pr + pair_stride) * 2 + static_cast<int>(rank);
        const size_t first = static_cast<size_t>(rn) * prm.pos_per_round * prm.in_planes * kSq;
        const size_t last = min(static_cast<size_t>(rn + 1) * prm.pos_per_round, static_cast<size_t>(prm.n)) * prm.in_planes * kSq;
        for (size_t i = first + threadIdx.x * 32; i < last; i += 256 * 32)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.planes + i));
      }
      if (prm.bits != nullptr && round_seq > 0) prefetch_packed_l1(prm, r * prm.pos_per_round, prm.pos_per_round, threadIdx.x, 256);
      // ---- phase 1: the stem's input.  All slices of both groups are converted before any epilogue of the round is
      // served (the stem epilogues need every slice anyway).  No residual of the previous round is live here, and
      // saying so (the registers are cleared) gives the conversion the registers to keep all its loads in flight:
      // with the residuals alive the compiler had serialised the four items of a pass into four dependent round trips
      // to L2, and the stem was bound by them instead of by its MMAs.
#pragma unroll
      for (int i = 0; i < F / 4; ++i) res_a[i] = 0u;
#pragma unroll
      for (int i = 0; i < (G::kGroups == 2 ? F / 4 : 1); ++i) res_b[i] = 0u;
      // One conversion item = 8 consecutive channels of one square of one position = one 16-byte row entry.  Which items
      // a thread converts does not depend on the slice or the group, so the split of the item index into (position,
      // plane, square), the source offset and the shared-memory offset are worked out once per round, not per event
      // (the events had been bound by their instruction stream: ~60 instructions per item, each event a walk through
      // cold instruction cache).
      constexpr int kSlicePlanes = kPlanes / 2;
      constexpr int kItems = kPosPerGroup * kSlicePlanes * kSq;
      constexpr int kPasses = (kItems + 1023) / 1024;
      int it_meta[kPasses][4], it_src[kPasses][4], it_dst[kPasses][4];
#pragma unroll
      for (int pass = 0; pass < kPasses; ++pass)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int it = static_cast<int>(threadIdx.x) + (pass * 4 + u) * 256;
          const int q = it / kSq, s_ = it - q * kSq, c8 = q % kSlicePlanes, p = q / kSlicePlanes;
          it_meta[pass][u] = s_ | (c8 << 8) | (p << 16) | ((it < kItems ? 1 : 0) << 24);
          it_src[pass][u] = prm.bits == nullptr ? (p * prm.in_planes + c8 * 8) * kSq + s_ : p * prm.packed_stride + c8 * 8;
          it_dst[pass][u] = c8 * kPlaneBytes + ((s_ / 5) * kW + p * 6 + s_ % 5) * 16;
        }
      while (next_slice0 < prm.nslices || next_slice1 < prm.nslices) {
        // slice k goes to plane half k&1 once the MMAs that read that half's previous slice are done (the group that is
        // further behind with its input goes first, so neither starves the other)
        int g = -1;
        long long tl0 = 0;
        if (kProf) tl0 = clock64();
        while (g < 0) {
          const int g0 = next_slice1 < next_slice0 ? 1 : 0;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int gg = g0 ^ k;
            const int ns = gg ? next_slice1 : next_slice0;
            if (g < 0 && ns < prm.nslices) {
              const int b = gg * 2 + (ns & 1);
              // the release of the buffer's previous slice is phase (count - 1) of its barrier
              if (slices.none_yet(b) || ptx::mbar_test_wait(&bar_in_empty[b], slices.release_parity(b))) g = gg;
            }
          }
        }
        // fp32 NCHW planes -> bf16 wide-board planes of this group (network.py:255-274 built the input):
        // one item = 8 consecutive channels of one square of one position = one 16-byte row entry
        const int sl = g ? next_slice1 : next_slice0;
        long long tc0 = 0;
        if (kProf) tc0 = clock64();
        const int hh = sl & 1;
        const int pl0 = sl * kSlicePlanes;                       // first input plane (8 channels each) of the slice
        const int npl = min(kSlicePlanes, prm.in_planes8 - pl0);
        const int pos0 = r * prm.pos_per_round + g * kPosPerGroup;
        uint8_t* dst_base = s_act + hh * kSlicePlanes * kPlaneBytes + group_origin_row(g) * 16;
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
          float f[4][8];
          // phase 1: issue every global load of up to 4 items before touching any result
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int meta = it_meta[pass][u];
            const int s_ = meta & 31, c8 = (meta >> 8) & 31, p = (meta >> 16) & 7;
            const bool live = (meta >> 24) != 0 && c8 < npl && pos0 + p < prm.n;
            const int nch = prm.in_planes - (pl0 + c8) * 8;      // channels of this 8-channel plane that exist (>= 8: all)
            if (prm.bits == nullptr) {
              const float* src = prm.planes + (static_cast<size_t>(pos0) * prm.in_planes + pl0 * 8) * kSq + it_src[pass][u];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[u][e] = (live && e < nch) ? __ldg(src + e * kSq) : 0.f;
            } else {
              // packed planes: x[c][sq] = bit sq of bits[c] ? scale[c] : 0
              const size_t o = static_cast<size_t>(pos0) * prm.packed_stride + pl0 * 8 + it_src[pass][u];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const bool ok = live && e < nch;
                const uint32_t wbits = ok ? __ldg(prm.bits + o + e) : 0u;
                const float sc = ok ? __ldg(prm.scale + o + e) : 0.f;
                f[u][e] = ((wbits >> s_) & 1u) ? sc : 0.f;
              }
            }
          }
          // phase 2: pack and store
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int meta = it_meta[pass][u];
            if ((meta >> 24) != 0 && ((meta >> 8) & 31) < npl) {
              const uint4 pk = make_uint4(ptx::pack_bf16x2(f[u][0], f[u][1]), ptx::pack_bf16x2(f[u][2], f[u][3]),
                                          ptx::pack_bf16x2(f[u][4], f[u][5]), ptx::pack_bf16x2(f[u][6], f[u][7]));
              *reinterpret_cast<uint4*>(dst_base + it_dst[pass][u]) = pk;
            }
          }
        }
        long long tfence = 0;
        if (kProf) tfence = clock64();
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive_on_leader(&bar_in_full[g * 2 + hh], rank);
        if (kProf && blockIdx.x == 0 && threadIdx.x == 0 && round_seq == 0) {
          unsigned long long* tr = prm.prof + 16 + 8192 + 512 + (g * 8 + sl) * 4;
          tr[0] = tc0; tr[1] = clock64(); tr[2] = tl0; tr[3] = tfence;
        }
        if (g) next_slice1 = sl + 1; else next_slice0 = sl + 1;
        slices.wrote(g * 2 + hh);
      }
      // ---- phase 2: epilogues, value tails ----
      while (next_layer0 < nl || next_layer1 < nl || fc_pending0 || fc_pending1) {
        long long te0 = 0, te1 = 0;
        if (kProf) te0 = clock64();
        int g;
        bool value_fc = false;
        for (;;) {  // serve whichever event is ready first (group A preferred: it runs ahead)
          if (next_layer0 < nl && ptx::mbar_test_wait(&bar_tmem_full[0], (round_seq * nl + next_layer0) & 1)) { g = 0; break; }
          if (next_layer1 < nl && ptx::mbar_test_wait(&bar_tmem_full[1], (round_seq * nl + next_layer1) & 1)) { g = 1; break; }
          if (fc_pending0 && ptx::mbar_test_wait(&bar_v[0], round_seq & 1)) { g = 0; value_fc = true; break; }
          if (fc_pending1 && ptx::mbar_test_wait(&bar_v[1], round_seq & 1)) { g = 1; value_fc = true; break; }
        }
        if (value_fc) {
          // value head tail of group g on one warp: ReLU(conv1x1) -> Dense(25->128)+ReLU -> Dense(128->1)+tanh
          // (network.py:111-120).  Lane l owns hidden units l, l+32, l+64, l+96.
          if (g) fc_pending1 = false; else fc_pending0 = false;
          float hsum[4][kPosPerGroup];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float b1 = prm.fc1_b[lane + 32 * i];
#pragma unroll
            for (int p = 0; p < kPosPerGroup; ++p) hsum[i][p] = b1;
          }
          for (int s_ = 0; s_ < kSq; ++s_) {
            float v1[kPosPerGroup];
#pragma unroll
            for (int p = 0; p < kPosPerGroup; ++p)
              v1[p] = fmaxf(s_v[((g * 2 + 0) * kPosPerGroup + p) * kSq + s_] +
                            s_v[((g * 2 + 1) * kPosPerGroup + p) * kSq + s_] + prm.value_b, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float kk = __ldg(prm.fc1_k + s_ * kValueHidden + lane + 32 * i);
#pragma unroll
              for (int p = 0; p < kPosPerGroup; ++p) hsum[i][p] = fmaf(v1[p], kk, hsum[i][p]);
            }
          }
          float part[kPosPerGroup] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float k2 = __ldg(prm.fc2_k + lane + 32 * i);
#pragma unroll
            for (int p = 0; p < kPosPerGroup; ++p) part[p] = fmaf(fmaxf(hsum[i][p], 0.f), k2, part[p]);
          }
#pragma unroll
          for (int p = 0; p < kPosPerGroup; ++p)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part[p] += __shfl_xor_sync(0xffffffffu, part[p], o);
          const int pos = r * prm.pos_per_round + g * kPosPerGroup + lane;
          if (lane < kPosPerGroup && pos < prm.n) {
            const float t = lane == 0 ? part[0] : lane == 1 ? part[1] : lane == 2 ? part[2] : part[3];
            prm.value[pos] = tanhf(t + prm.fc2_b);
          }
          continue;
        }
        ptx::tc_fence_after_sync();
        if (kProf) te1 = clock64();
        const int L = g ? next_layer1 : next_layer0;
        if (g) next_layer1 = L + 1; else next_layer0 = L + 1;
        const uint32_t lf = prm.layers[L].flags;   // (L1-resident; a copy in shared memory measured 0.7 % slower)
        const float* bias = bias_base + L * F + h * G::kHalfCh;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (g * 2 + (L & 1)) * F + h * G::kHalfCh;
        uint64_t* bar_part = &bar_act_part[g * 4 + h * G::kCb];
        const int pos_base = r * prm.pos_per_round + g * kPosPerGroup;

        if (!(lf & kFinal)) {
          uint8_t* act_row = s_act + (h * (G::kHalfCh / 8)) * kPlaneBytes + (group_origin_row(g) + row) * 16;
          uint4* dbg_row = (prm.dbg != nullptr && L == nl - 1 && pos_base < prm.n)
                               ? prm.dbg + (static_cast<size_t>(pos_base / kPosPerGroup) * kPlanes + h * (G::kHalfCh / 8)) * kTile + row : nullptr;
          const float* wv = s_wv + h * G::kHalfCh;
          float vacc = 0.f;
#define E55_EPI(RES)                                                                                          \
          if (lf & kValueTap) {                                                                               \
            if (lf & kHasRes) vacc = epilogue_half<F, true, true, true>(taddr, bias, act_row, RES, wv, row_valid, dbg_row, bar_part, lane, rank);   \
            else vacc = epilogue_half<F, false, true, true>(taddr, bias, act_row, RES, wv, row_valid, dbg_row, bar_part, lane, rank);  \
          } else if (lf & kHasRes) {                                                                          \
            epilogue_half<F, true, true, false>(taddr, bias, act_row, RES, wv, row_valid, dbg_row, bar_part, lane, rank);              \
          } else if (lf & kUpdRes) {                                                                          \
            epilogue_half<F, false, true, false>(taddr, bias, act_row, RES, wv, row_valid, dbg_row, bar_part, lane, rank);             \
          } else {                                                                                            \
            epilogue_half<F, false, false, false>(taddr, bias, act_row, RES, wv, row_valid, dbg_row, bar_part, lane, rank);            \
          }
          if (G::kGroups == 1 || g == 0) { E55_EPI(res_a) } else if constexpr (G::kGroups == 2) { E55_EPI(res_b) }
#undef E55_EPI
          // value 1x1 conv: this thread's half of the channel dot product (summed in the value tail)
          if (lf & kValueTap) {
            if (row_valid) s_v[((g * 2 + h) * kPosPerGroup + bp) * kSq + sq] = vacc;
            __syncwarp();
            if (lane == 0 && do_value) ptx::mbar_arrive(&bar_v[g]);
          }
        } else {
          // policy 1x1 output: fp32 logits, index c*25 + sq (Flatten of NCHW)
          if (!prm.tail) {
            // ... straight to HBM from the accumulator registers
            const int pos = pos_base + bp;
            const bool out_valid = row_valid && pos < prm.n && prm.policy != nullptr;
            float* dst = prm.policy + static_cast<size_t>(pos) * (kPolicyCh * kSq) + sq;
#pragma unroll
            for (int cb = 0; cb < G::kCb; ++cb) {
              if (h * G::kHalfCh + cb * 32 >= kPolicyN) break;  // columns beyond the padded policy width do not exist
              uint32_t v[32];
              ptx::tmem_ld_32x32b_x32(taddr + cb * 32, v);
              ptx::tmem_ld_wait();
              if (out_valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int c = h * G::kHalfCh + cb * 32 + j;
                  if (c < kPolicyCh) dst[c * kSq] = __uint_as_float(v[j]) + bias[cb * 32 + j];
                }
              }
            }
          } else {
            // ... staged in the group's activation planes (nothing reads those any more: this layer's MMAs, their last
            // readers, are complete) for the policy tail
            uint8_t* stage_row = s_act + group_origin_row(g) * 16 + (bp * kSq + sq) * 4;
#pragma unroll
            for (int cb = 0; cb < G::kCb; ++cb) {
              if (h * G::kHalfCh + cb * 32 >= kPolicyN) break;
              uint32_t v[32];
              ptx::tmem_ld_32x32b_x32(taddr + cb * 32, v);
              ptx::tmem_ld_wait();
              if (row_valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  // (h is 0 or 1: both offsets are compile-time constants, selected per thread)
                  const int c0 = cb * 32 + j, c1 = G::kHalfCh + cb * 32 + j;
                  if (h == 0 ? c0 < kPolicyCh : c1 < kPolicyCh)
                    *reinterpret_cast<float*>(stage_row + (h == 0 ? stage_offset<F>(c0) : stage_offset<F>(c1 < kPolicyCh ? c1 : 0))) =
                        __uint_as_float(v[j]) + bias[cb * 32 + j];
                }
              }
            }
            __syncwarp();
            if (lane == 0 && do_value) ptx::mbar_arrive(&bar_p[g]);
          }
          // no activations written, but keep the per-layer barrier phases in step
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0)
            for (int cb = 0; cb < G::kCb; ++cb) arrive_on_leader(&bar_part[cb], rank);
        }
        if (kProf) {
          const long long te2 = clock64();
          t_wait += te1 - te0; t_work += te2 - te1;
          if (blockIdx.x == 0 && threadIdx.x == 0 && round_seq == 0) {
            unsigned long long* tr = prm.prof + 16 + 8192 + (g * nl + L) * 4;
            tr[0] = te1; tr[1] = te2; tr[2] = g; tr[3] = L;
          }
        }
      }

    }
    if (kProf && blockIdx.x == 0 && threadIdx.x == 0) {
      prm.prof[10] = t_wait;
      prm.prof[11] = t_work;
    }
  }

  // neither CTA may exit (or free TMEM) while the other can still signal its barriers / read its smem
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (kProf && blockIdx.x == 0 && threadIdx.x == 0) prm.prof[14] = clock64();
  ptx::cluster_sync();
  if (warp == 9) ptx::tmem_dealloc_2sm<kTmemCols>(tmem_base);
  if (kProf && blockIdx.x == 0 && threadIdx.x == 0) prm.prof[15] = clock64();
}

}  // namespace tc
}  // namespace e55
#include "e55_tc_split.cuh"
namespace e55 {
namespace tc {

// =================================================================================================
// host side
// =================================================================================================
struct HeadWeights {
  const float* policy_w; const float* policy_b;   // host, folded [cin][69], [69]
  const float* d_value_w; float value_b;          // device [filters]
  const float* d_fc1_k; const float* d_fc1_b; const float* d_fc2_k; float fc2_b;  // device
};

struct State {
  int blocks = 0, filters = 0, in_planes = 0, sms = 0;
  int nlayers = 0, nslices = 0, in_planes8 = 0;
  int nstages = 0, nchunks = 0;
  int lag = 0;                      // free-running groups measured fastest (tools/tc_prof.py, E55_LAG)
  bool throughput_rounds = false;   // set by callers that run several chunk kernels concurrently: never use latency mode
  bool no_split = false;            // set by callers that pipeline many small launches: throughput per SM matters more than latency
  int split_max = 0;                // batches up to this size run on the channel-split kernel (e55_tc_split.cuh): an experiment that
                                    // did NOT pay (98 us against 75 us at batch 16, see its header), off unless E55_SPLIT_MAX says so
  size_t smem_bytes = 0;
  uint8_t* d_wimg = nullptr;
  Chunk* d_chunks = nullptr;
  LayerInfo* d_layers = nullptr;
  float* d_bias = nullptr;
  std::vector<LayerInfo> layers;
  HeadWeights heads{};
  uint4* d_dbg = nullptr;
  unsigned long long* d_prof = nullptr;  // set by e55_debug_profile
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // when set, recorded around the tower kernel launch
  int capacity = 0;
  bool attr_set = false, split_attr_set = false;
};

inline void init(State& s, int blocks, int filters, int in_planes, int sms) {
  s.blocks = blocks; s.filters = filters; s.in_planes = in_planes; s.sms = sms;
  s.nlayers = 1 + 2 * blocks + 2;
  if (const char* env = getenv("E55_SPLIT_MAX")) s.split_max = atoi(env);   // 0 switches the latency kernel off
  s.in_planes8 = ((in_planes + 15) / 16) * 2;
  const int slice_planes = filters / 16;   // stem input slices fill half of the activation planes, alternating
  s.nslices = slice_planes > 0 ? (s.in_planes8 + slice_planes - 1) / slice_planes : 0;
}

inline bool supported(const State& s) { return (s.filters == 128 || s.filters == 256) && s.nlayers <= 255; }

struct HostGeo { int groups, planes, rows_total, act_bytes, kblocks; bool bias_in_smem; };
inline HostGeo host_geo(int filters) {
  if (filters == 128)
    return HostGeo{Geo<128>::kGroups, Geo<128>::kPlanes, Geo<128>::kRowsTotal, Geo<128>::kActBytes, Geo<128>::kKBlocks, Geo<128>::kBiasInSmem};
  return HostGeo{Geo<256>::kGroups, Geo<256>::kPlanes, Geo<256>::kRowsTotal, Geo<256>::kActBytes, Geo<256>::kKBlocks, Geo<256>::kBiasInSmem};
}

inline void free_weights(State& s) {
  cudaFree(s.d_wimg); cudaFree(s.d_chunks); cudaFree(s.d_layers); cudaFree(s.d_bias);
  s.d_wimg = nullptr; s.d_chunks = nullptr; s.d_layers = nullptr; s.d_bias = nullptr;
  s.layers.clear();
}

inline void free_buffers(State& s) {
  cudaFree(s.d_dbg);
  s.d_dbg = nullptr; s.capacity = 0;
}

inline cudaError_t alloc_buffers(State& s, int cap) {
  if (!supported(s)) return cudaSuccess;
  free_buffers(s);
  const size_t groups = (static_cast<size_t>(cap) + kPosPerGroup - 1) / kPosPerGroup + 4;
  cudaError_t e = cudaMalloc(&s.d_dbg, groups * (s.filters / 8) * kTile * sizeof(uint4));
  if (e != cudaSuccess) return e;
  s.capacity = cap;
  return cudaSuccess;
}

inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return static_cast<uint16_t>(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

// conv_w[i]: folded fp32 [9][cin][128] for the stem, tower convs and the policy conv (in that order);
// heads.policy_w: folded [128][69].
inline cudaError_t set_weights(State& s, const std::vector<std::vector<float>>& conv_w,
                               const std::vector<std::vector<float>>& conv_b, const HeadWeights& heads) {
  if (!supported(s)) return cudaSuccess;
  free_weights(s);
  s.heads = heads;
  const int F = s.filters;
  const HostGeo geo = host_geo(F);
  const int kPlanes = geo.planes;
  std::vector<Chunk> chunks;
  std::vector<uint16_t> img;
  std::vector<float> bias(static_cast<size_t>(s.nlayers) * F, 0.f);
  s.layers.assign(s.nlayers, LayerInfo{});

  // chunk = (layer, slice, K-block set, board row dy): ntaps taps x (nk0 + nk1) K-steps, MMA order
  // (tap, K block, K-step).  image: [half: cout 0..ndim/2-1 | ndim/2..ndim-1][tap][K block][plane j][n][8]
  struct KB { int ch0; int nk; };  // first input channel and K-steps of a K block
  auto emit = [&](int layer, int tap0, int ntaps, int dy, int first_plane, KB k0, KB k1, int ndim, int cin, int cout,
                  const float* w, uint32_t flags, uint32_t need) {
    Chunk ch{};
    ch.gofs = static_cast<uint32_t>(img.size() * 2);
    ch.half16 = static_cast<uint16_t>(ntaps * (k0.nk + k1.nk) * 2 * (ndim / 2));
    ch.a_off = static_cast<int16_t>(first_plane * geo.rows_total + dy * kW);
    ch.nk0 = static_cast<uint8_t>(k0.nk);
    ch.nk1 = static_cast<uint8_t>(k1.nk);
    ch.need = static_cast<uint8_t>(need);
    ch.ndim8 = static_cast<uint8_t>(ndim / 8);
    ch.flags = static_cast<uint8_t>(flags | (ntaps == 1 ? kSingleTap : 0u));
    ch.layer = static_cast<uint8_t>(layer);
    for (int half = 0; half < 2; ++half)
      for (int t = 0; t < ntaps; ++t)
        for (int blk = 0; blk < 2; ++blk) {
          const KB& k = blk ? k1 : k0;
          for (int j = 0; j < k.nk * 2; ++j)
            for (int nn = 0; nn < ndim / 2; ++nn)
              for (int e = 0; e < 8; ++e) {
                const int ci = k.ch0 + j * 8 + e;
                const int n = half * (ndim / 2) + nn;
                const float v = (ci < cin && n < cout) ? w[(static_cast<size_t>(tap0 + t) * cin + ci) * cout + n] : 0.f;
                img.push_back(f32_to_bf16_rne(v));
              }
        }
    chunks.push_back(ch);
  };

  for (int L = 0; L < s.nlayers; ++L) {
    LayerInfo& li = s.layers[L];
    li.chunk_begin = static_cast<int>(chunks.size());
    const bool is_final = L == s.nlayers - 1;
    const bool is_stem = L == 0;
    const bool is_policy_conv = L == s.nlayers - 2;
    const int cin = is_stem ? s.in_planes : F;
    const int cout = is_final ? kPolicyCh : F;
    const int ndim = is_final ? kPolicyN : F;
    const float* w = is_final ? heads.policy_w : conv_w[L].data();
    const float* b = is_final ? heads.policy_b : conv_b[L].data();
    for (int o = 0; o < cout; ++o) bias[static_cast<size_t>(L) * F + o] = b[o];
    li.ndim = ndim;
    li.flags = 0;
    if (is_final) li.flags = kFinal;
    else if (is_stem) li.flags = kRelu | kUpdRes;
    else if (is_policy_conv) li.flags = kRelu;
    else if ((L - 1) % 2 == 0) li.flags = kRelu;                // first conv of a block
    else li.flags = kRelu | kHasRes | kUpdRes;                  // second conv: relu(y + x)
    if (L == s.nlayers - 3) li.flags |= kValueTap;              // trunk output feeds the value head
    const int planes_total = is_stem ? s.in_planes8 : kPlanes;  // 8-channel planes of the input
    const int slice_planes = is_stem ? kPlanes / 2 : kPlanes;   // stem: double-buffered half-plane slices
    const int nslices = (planes_total + slice_planes - 1) / slice_planes;
    for (int sl = 0; sl < nslices; ++sl) {
      const int pl0 = sl * slice_planes;                          // first input plane of the slice
      const int npl = std::min(slice_planes, planes_total - pl0);
      const int base = is_stem ? (sl & 1) * slice_planes : 0;     // activation plane the slice starts at
      const bool first_sl = sl == 0, last_sl = sl == nslices - 1;
      if (is_final) {
        // 1x1 conv: one chunk, every K-step over all planes, needs every activation K block
        emit(L, 0, 1, 0, 0, KB{0, npl / 2}, KB{0, 0}, ndim, cin, cout, w,
             kFirstOfSlice | kLastOfSlice | kFirstOfLayer | kLastOfLayer, (1u << geo.kblocks) - 1u);
        continue;
      }
      // K-block sets of this slice, in the order in which the previous layer's epilogue finishes them
      // (its two thread halves work through their K blocks in parallel)
      struct Set { int plane; KB k0, k1; uint32_t need; };
      std::vector<Set> sets;
      const int hk = geo.kblocks / 2;  // K blocks per epilogue half
      if (is_stem && F == 128) {
        sets.push_back(Set{base, KB{pl0 * 8, npl / 2}, KB{0, 0}, 0u});   // <= 4 K-steps over <= 8 consecutive planes
      } else if (is_stem) {
        for (int kb = 0; kb * 4 < npl; ++kb)
          sets.push_back(Set{base + kb * 4, KB{(pl0 + kb * 4) * 8, std::min(2, (npl - kb * 4) / 2)}, KB{0, 0}, 0u});
      } else if (npl == kPlanes && F == 128) {
        for (int pi = 0; pi < hk; ++pi)   // two K blocks per chunk: {pi, pi + hk}
          sets.push_back(Set{pi * 4, KB{(pl0 + pi * 4) * 8, 2}, KB{(pl0 + (pi + hk) * 4) * 8, 2}, (1u << pi) | (1u << (pi + hk))});
      } else if (npl == kPlanes) {
        for (int pi = 0; pi < hk; ++pi)   // one K block per chunk (N = 256 fills the stage): pi, pi + hk, ...
          for (int hh = 0; hh < 2; ++hh) {
            const int kb = pi + hh * hk;
            sets.push_back(Set{kb * 4, KB{(pl0 + kb * 4) * 8, 2}, KB{0, 0}, 1u << kb});
          }
      } else {
        for (int kb = 0; kb * 4 < npl; ++kb)
          sets.push_back(Set{kb * 4, KB{(pl0 + kb * 4) * 8, std::min(2, (npl - kb * 4) / 2)}, KB{0, 0}, 1u << kb});
      }
      for (size_t si = 0; si < sets.size(); ++si)
        for (int dy = 0; dy < 3; ++dy) {
          uint32_t fl = 0;
          const bool first = si == 0 && dy == 0, last = si + 1 == sets.size() && dy == 2;
          if (first) fl |= kFirstOfSlice;
          if (last) fl |= kLastOfSlice;
          if (first && first_sl) fl |= kFirstOfLayer;
          if (last && last_sl) fl |= kLastOfLayer;
          uint32_t need = 0;
          if (!is_stem && dy == 0) need = sets[si].need;          // K blocks written by the previous epilogue
          if (is_stem && first && first_sl) need = (1u << geo.kblocks) - 1u;   // previous round fully drained
          emit(L, dy * 3, 3, dy - 1, sets[si].plane, sets[si].k0, sets[si].k1, ndim, cin, cout, w, fl, need);
        }
    }
    li.chunk_end = static_cast<int>(chunks.size());
    for (int c = li.chunk_begin; c < li.chunk_end; ++c) chunks[c].lend = static_cast<uint16_t>(li.chunk_end);
  }

  cudaError_t e;
  if ((e = cudaMalloc(&s.d_wimg, img.size() * 2)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(s.d_wimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&s.d_chunks, chunks.size() * sizeof(Chunk))) != cudaSuccess) return e;
  if ((e = cudaMemcpy(s.d_chunks, chunks.data(), chunks.size() * sizeof(Chunk), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&s.d_layers, s.layers.size() * sizeof(LayerInfo))) != cudaSuccess) return e;
  if ((e = cudaMemcpy(s.d_layers, s.layers.data(), s.layers.size() * sizeof(LayerInfo), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&s.d_bias, bias.size() * 4)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(s.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return e;

  s.nchunks = static_cast<int>(chunks.size());
  const size_t fixed = geo.act_bytes + (geo.bias_in_smem ? static_cast<size_t>(s.nlayers) * F * 4 : 0) + F * 4 +
                       4 * kPosPerGroup * kSq * 4 + chunks.size() * sizeof(Chunk) + (2 * kMaxStages + 24) * 8 + 16;
  s.nstages = static_cast<int>(std::min<size_t>(kMaxStages, (232448 - std::min<size_t>(fixed, 232448)) / kStageBytes));
  s.smem_bytes = fixed + static_cast<size_t>(s.nstages) * kStageBytes;
  s.attr_set = false;
  s.split_attr_set = false;
  return cudaSuccess;
}

// nlayers_run < nlayers: stop after that many layers and leave the packed activations in d_dbg.
struct PackedInput { const uint32_t* bits = nullptr; const float* scale = nullptr; int stride = 0; };   // stride 0 = in_planes
// What the fused policy tail produces besides (or instead of) the raw logits; see Params.
struct PolicyTail {
  const int32_t* move_off = nullptr; const uint16_t* move_idx = nullptr; float* priors = nullptr; int move_stride = 0;
  const uint8_t* mask = nullptr; float* probs = nullptr;
  bool rows = false;   // write `policy` through the tail as whole rows (for mapped host memory) even if nothing else is wanted
};

inline cudaError_t run(State& s, const float* d_planes, int n, float* d_policy, float* d_value, int nlayers_run,
                       cudaStream_t stream, int* launched, std::string* why, PackedInput packed = PackedInput(),
                       PolicyTail tail = PolicyTail()) {
  if (!supported(s)) { *why = "the tcgen05 path supports 128 or 256 filters"; return cudaSuccess; }
  if (s.nstages < 3) { *why = "network too deep for the shared-memory layer tables"; return cudaSuccess; }
  if (!s.attr_set) {
    const int sm = static_cast<int>(s.smem_bytes);
    cudaError_t e = s.filters == 128
        ? cudaFuncSetAttribute(tower_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)
        : cudaFuncSetAttribute(tower_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    if (e == cudaSuccess)
      e = s.filters == 128 ? cudaFuncSetAttribute(tower_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)
                           : cudaFuncSetAttribute(tower_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    if (e != cudaSuccess) { *why = "cudaFuncSetAttribute"; return e; }
    s.attr_set = true;
  }
  // small batches (the reference's own: 1 and 16 positions): the channel-split latency kernel, one group per CTA pair
  if (s.filters == 128 && !s.throughput_rounds && !s.no_split && n <= s.split_max && (s.d_prof == nullptr || getenv("E55_SPLIT_PROF"))) {
    if (!s.split_attr_set) {
      cudaError_t e = cudaFuncSetAttribute(tower_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(s.smem_bytes));
      if (e != cudaSuccess) { *why = "cudaFuncSetAttribute (split kernel)"; return e; }
      s.split_attr_set = true;
    }
    Params p{};
    p.wimg = s.d_wimg; p.chunks = s.d_chunks; p.layers = s.d_layers; p.bias = s.d_bias;
    p.value_w = s.heads.d_value_w; p.fc1_k = s.heads.d_fc1_k; p.fc1_b = s.heads.d_fc1_b; p.fc2_k = s.heads.d_fc2_k;
    p.value_b = s.heads.value_b; p.fc2_b = s.heads.fc2_b;
    p.planes = d_planes; p.bits = packed.bits; p.scale = packed.scale; p.packed_stride = packed.stride ? packed.stride : s.in_planes;
    p.in_planes = s.in_planes; p.in_planes8 = s.in_planes8; p.nslices = s.nslices;
    p.nstages = s.nstages; p.nchunks_total = s.nchunks; p.nlayers = s.nlayers; p.nlayers_run = nlayers_run; p.n = n;
    p.nrounds = (n + kPosPerGroup - 1) / kPosPerGroup; p.pos_per_round = kPosPerGroup;
    p.policy = d_policy; p.value = d_value;
    p.move_off = tail.move_off; p.move_idx = tail.move_idx; p.priors = tail.priors; p.move_stride = tail.move_stride;
    p.mask = tail.mask; p.probs = tail.probs;
    p.tail = (tail.priors != nullptr || tail.probs != nullptr || tail.rows) ? 1 : 0;
    p.dbg = nlayers_run < s.nlayers ? s.d_dbg : nullptr;
    p.lag = s.lag;
    p.prof = s.d_prof;
    const int clusters = std::min(p.nrounds, std::max(1, s.sms / 2));
    if (s.ev0) cudaEventRecord(s.ev0, stream);
    tower_split_kernel<<<2 * clusters, kSplitThreads, s.smem_bytes, stream>>>(p);
    if (s.ev1) cudaEventRecord(s.ev1, stream);
    ++*launched;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) *why = "tower_split_kernel launch";
    return e;
  }
  // latency mode: while the batch fits the machine with one 4-position group per CTA, leave group B out
  const int ppr = (s.filters != 128 || (!s.throughput_rounds && n <= kPosPerGroup * 2 * std::max(1, s.sms / 2))) ? kPosPerGroup : kPosPerRound;
  const int rounds = (n + ppr - 1) / ppr;
  Params p{};
  p.wimg = s.d_wimg; p.chunks = s.d_chunks; p.layers = s.d_layers; p.bias = s.d_bias;
  p.value_w = s.heads.d_value_w; p.fc1_k = s.heads.d_fc1_k; p.fc1_b = s.heads.d_fc1_b; p.fc2_k = s.heads.d_fc2_k;
  p.value_b = s.heads.value_b; p.fc2_b = s.heads.fc2_b;
  p.planes = d_planes; p.bits = packed.bits; p.scale = packed.scale; p.packed_stride = packed.stride ? packed.stride : s.in_planes; p.in_planes = s.in_planes; p.in_planes8 = s.in_planes8; p.nslices = s.nslices;
  p.nstages = s.nstages; p.nchunks_total = s.nchunks; p.lag = std::min(s.lag, s.nstages - 2);
  p.nlayers = s.nlayers; p.nlayers_run = nlayers_run; p.n = n; p.nrounds = rounds; p.pos_per_round = ppr;
  p.policy = d_policy; p.value = d_value;
  p.move_off = tail.move_off; p.move_idx = tail.move_idx; p.priors = tail.priors; p.move_stride = tail.move_stride;
  p.mask = tail.mask; p.probs = tail.probs;
  p.tail = (tail.priors != nullptr || tail.probs != nullptr || tail.rows) ? 1 : 0;
  static const bool scalar_rows = getenv("E55_TAIL_VEC") && atoi(getenv("E55_TAIL_VEC")) == 0;   // A/B: element-wise row stores
  if (p.tail && scalar_rows) p.tail |= 2;
  p.dbg = nlayers_run < s.nlayers ? s.d_dbg : nullptr;
  p.prof = s.d_prof;
  const int pairs = std::min((rounds + 1) / 2, std::max(1, s.sms / 2));
  const int grid = 2 * pairs;
  if (s.ev0) cudaEventRecord(s.ev0, stream);
  if (s.filters == 128) {
    if (p.prof) tower_kernel<128, true><<<grid, kThreads, s.smem_bytes, stream>>>(p);
    else tower_kernel<128, false><<<grid, kThreads, s.smem_bytes, stream>>>(p);
  } else {
    if (p.prof) tower_kernel<256, true><<<grid, kThreads, s.smem_bytes, stream>>>(p);
    else tower_kernel<256, false><<<grid, kThreads, s.smem_bytes, stream>>>(p);
  }
  if (s.ev1) cudaEventRecord(s.ev1, stream);
  ++*launched;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) *why = "tower_kernel launch";
  return e;
}

inline cudaError_t forward(State& s, const float* d_planes, int n, float* d_policy, float* d_value,
                           cudaStream_t stream, int* launched, std::string* why, PolicyTail tail = PolicyTail()) {
  return run(s, d_planes, n, d_policy, d_value, s.nlayers, stream, launched, why, PackedInput(), tail);
}

inline cudaError_t forward_packed(State& s, const uint32_t* d_bits, const float* d_scale, int n, float* d_policy,
                                  float* d_value, cudaStream_t stream, int* launched, std::string* why, int stride = 0,
                                  PolicyTail tail = PolicyTail()) {
  return run(s, nullptr, n, d_policy, d_value, s.nlayers, stream, launched, why, PackedInput{d_bits, d_scale, stride}, tail);
}

// Activations after `nconv` 3x3 convolutions (1 = stem ... 1+2*blocks = trunk) as fp32 NCHW.
inline cudaError_t debug_layers(State& s, const float* d_planes, int n, int nconv, float* d_out_nchw,
                                cudaStream_t stream, int* launched, std::string* why) {
  cudaError_t e = run(s, d_planes, n, nullptr, nullptr, nconv, stream, launched, why);
  if (e != cudaSuccess || !why->empty()) return e;
  unpack_debug_kernel<<<n, 128, 0, stream>>>(s.d_dbg, d_out_nchw, n, s.filters);
  ++*launched;
  e = cudaGetLastError();
  if (e != cudaSuccess) *why = "unpack_debug_kernel launch";
  return e;
}

}  // namespace tc
}  // namespace e55

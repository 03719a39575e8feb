// Channel-split variant of the tower kernel (e55_tc.cuh) for the batches the reference's search really sends: the root
// (1 position) and leaf batches of 16 (mcts.py:57-58,101-102; usi.py:119-120).  AN EXPERIMENT THAT DID NOT PAY, kept
// (off by default: E55_SPLIT_MAX=<batch> switches it on; tests/test_gpu_parity.py::test_channel_split_kernel_is_bit_identical
// keeps it honest) because the measurement explains where the latency floor of this design is.
//
// In the throughput kernel a CTA pair evaluates two 4-position tiles with cta_group::2 UMMAs of M = 256, N = 128: every
// layer costs 72 UMMAs however few positions there are, plus an epilogue nothing overlaps when a CTA has one group only
// -- 7.0 k cycles per layer (5.2 k MMAs + 1.8 k exposed epilogue), 75 us for the network.  Here the two CTAs of a cluster
// work on the SAME 4-position group and split the OUTPUT CHANNELS of every layer: each holds the group's complete
// activations, streams its own half of every weight chunk (the very halves of the throughput kernel's weight image) and
// issues cta_group::1 UMMAs of M = 128, N = 64, and its epilogue (32 channels per thread instead of 64) hands each finished
// 32-channel K block to both issuers: to its own through the usual generic-proxy stores + fence, to the peer's as bulk
// shared-to-shared copies through the async proxy that complete on the peer's K-block mbarrier.  A CTA's epilogue stores
// only after BOTH CTAs' MMAs of the layer are complete (the planes are overwritten in place and the peer's tensor core may
// still be reading them: each epilogue warp tells the peer when it has seen its own accumulator complete).
//
// Measured (tools/split_trace.py, batch 16): 8.8 k cycles per layer, 98 us for the network -- SLOWER.  The 72 N = 64 UMMAs
// of a layer take 6.1 k cycles, 85 each, against 72 for the pair's M = 256 x N = 128 instruction: in SS mode an UMMA costs
// its shared-memory operand reads (4 KB of A for the 128 rows whatever N is, + 2 KB of B: ~6 KB at ~80 B/clk in both
// kernels), not its FLOPs, so halving N halves nothing, and the exchange adds 2.6 k cycles per layer.  The latency floor
// of the implicit-GEMM-from-shared-memory design is therefore 72 UMMAs x ~70 cycles = 5 k cycles per layer and tile, ~50 us
// for 17 layers plus stem and set-up; the 75 us of the throughput kernel's latency mode are within 1.5x of it, and the way
// below leads through fewer A re-reads per layer (wider N: the 20x256 tower runs at 64 B/clk per SM), not through more SMs.
// One more lesson, paid for in hours: UMMA shared-memory descriptors take the offset inside the CTA's own shared memory
// (14 bits); in the second CTA of a cluster the shared-window address has upper bits set and must be masked.
//
// Arithmetic is the throughput kernel's, operation for operation (same K order, same epilogue expression; the value
// head's channel sum is continued from the lower 32 channels' partial exactly as one thread of the throughput kernel
// does it), so results are bit-identical.  Reference net only (128 filters).
// (Included by e55_tc.cuh after the throughput kernel: shares its parameter block, chunk tables, weight image and tail.)
#pragma once

namespace e55 {
namespace tc {

constexpr int kSplitThreads = 320;     // warps 0-7 epilogue + input conversion, 8 weight producer (+ policy tail), 9 MMA issuer
constexpr int kSplitTmemCols = 256;    // two accumulators (layer parity) of 128 columns, 64 (48) of them used

__device__ __forceinline__ void umma_single(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSplitThreads, 1)
tower_split_kernel(const Params prm) {
  constexpr int F = 128;
  using G = Geo<F>;
  constexpr int kPlanes = G::kPlanes;
  constexpr int kPlaneBytes = G::kPlaneBytes;
  extern __shared__ __align__(128) uint8_t smem[];
  // same carve as the throughput kernel (same dynamic shared memory size): ring | planes | biases | value taps | chunks | barriers
  uint8_t* s_ring = smem + SmemLayout::ring;
  uint8_t* s_act = smem + SmemLayout::act(prm.nstages);
  float* s_bias = reinterpret_cast<float*>(s_act + G::kActBytes);
  float* s_wv = s_bias + prm.nlayers * F;
  float* s_v = s_wv + F;                                // [2 channel halves = CTA ranks][4 positions][25] (rank 0's copy is read)
  float* s_part = s_v + 2 * kPosPerGroup * kSq;         // [128 rows] value partial of the lower 32 channels of this CTA's half
  Chunk* s_chunks = reinterpret_cast<Chunk*>(s_v + 4 * kPosPerGroup * kSq);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_chunks + prm.nchunks_total);
  const uint32_t nstages = prm.nstages;
  uint64_t* bar_full = bars;                            // [stages] this CTA's half of the stage has landed
  uint64_t* bar_empty = bars + kMaxStages;              // [stages] this CTA's MMAs on the stage are done
  uint64_t* bar_tmem_full = bars + 2 * kMaxStages;      // [1] this CTA's accumulator of the layer is complete
  uint64_t* bar_peer = bar_tmem_full + 1;               // [1] ... and so is the peer's (its 8 epilogue warps say so)
  uint64_t* bar_act_part = bar_peer + 1;                // [4] activation K block kb written into THIS CTA's planes (by CTA kb / 2)
  uint64_t* bar_in_full = bar_act_part + 4;             // [2 plane halves] stem input slice written
  uint64_t* bar_in_empty = bar_in_full + 2;             // [2] that half of the planes may be overwritten
  uint64_t* bar_v = bar_in_empty + 2;                   // [1] rank 0: both CTAs' value partials written
  uint64_t* bar_p = bar_v + 1;                          // [1] rank 0: both CTAs' logits staged
  uint64_t* bar_s0 = bar_p + 1;                         // [1] lower-half value partials of this CTA written
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_s0 + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t peer = rank ^ 1u;
  const int nl = prm.nlayers_run;                       // == prm.nlayers, or fewer for a debug dump of the activations
  const bool full = nl == prm.nlayers;
  const uint32_t nchunks = prm.layers[nl - 1].chunk_end;
  const int cl0 = blockIdx.x >> 1, cl_stride = gridDim.x >> 1;
  const int nrounds = prm.nrounds;                      // one 4-position group per cluster and round

  // ---- one-time setup -----------------------------------------------------------------------
  for (int i = threadIdx.x; i < G::kActBytes / 16; i += kSplitThreads)
    reinterpret_cast<uint4*>(s_act)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < prm.nlayers * F; i += kSplitThreads) s_bias[i] = prm.bias[i];
  for (int i = threadIdx.x; i < F; i += kSplitThreads) s_wv[i] = prm.value_w[i];
  for (int i = threadIdx.x; i < prm.nchunks_total; i += kSplitThreads) s_chunks[i] = prm.chunks[i];
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < nstages; ++s) {
      ptx::mbar_init(&bar_full[s], 1);
      ptx::mbar_init(&bar_empty[s], 1);
    }
    ptx::mbar_init(bar_tmem_full, 1);
    ptx::mbar_init(bar_peer, 8);                        // the peer's 8 epilogue warps
    for (int kb = 0; kb < 4; ++kb) ptx::mbar_init(&bar_act_part[kb], 4);   // the 4 warps (lane quadrants) that own the block
    for (int hh = 0; hh < 2; ++hh) {
      ptx::mbar_init(&bar_in_full[hh], 8);
      ptx::mbar_init(&bar_in_empty[hh], 1);
    }
    ptx::mbar_init(bar_v, 8);                           // 4 upper-half warps x 2 CTAs
    ptx::mbar_init(bar_p, 16);                          // 8 warps x 2 CTAs
    ptx::mbar_init(bar_s0, 4);
    ptx::fence_mbar_init();
  }
  if (warp == 9) ptx::tmem_alloc<kSplitTmemCols>(s_tmem);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();   // the peer's barriers and zeroed planes exist before anyone touches them remotely
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 8) {
    // ===== weight producer: this CTA's half (64 output channels) of every chunk; on rank 0 then the policy tail ==========
    uint32_t s = 0, ph = 1, round_seq = 0;
    for (int rd = cl0; rd < nrounds; rd += cl_stride, ++round_seq) {
      if (lane == 0) {
        for (uint32_t c = 0; c < nchunks; ++c) {
          const Chunk ch = s_chunks[c];
          ptx::mbar_wait(&bar_empty[s], ph);
          const uint32_t bytes = static_cast<uint32_t>(ch.half16) * 16u;
          ptx::mbar_arrive_expect_tx(&bar_full[s], bytes);
          ptx::tma_bulk_g2s(s_ring + s * kStageBytes, prm.wimg + ch.gofs + rank * bytes, bytes, &bar_full[s]);
          if (++s == nstages) { s = 0; ph ^= 1; }
        }
      }
      __syncwarp();
      if (prm.tail && full && rank == 0) {
        ptx::mbar_wait_cluster(bar_p, round_seq & 1);
        policy_tail<F>(prm, s_act, bar_in_empty, 0, rd * kPosPerGroup, lane);   // (releases rank 0's planes) ...
        if (lane == 0) {                                                         // ... and the peer's
          ptx::mbar_arrive_cluster(&bar_in_empty[0], peer);
          ptx::mbar_arrive_cluster(&bar_in_empty[1], peer);
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer: cta_group::1, M = 128 (the group's tile), N = this CTA's half of the layer's output channels ======
    // (the descriptors' 14-bit address field takes the offset inside this CTA's shared memory: in the second CTA of a
    // cluster the shared-window address carries the CTA's position in its upper bits, which must not reach the LBO field)
    const uint32_t a_lo_base = (((ptx::smem_u32(s_act) + group_origin_row(0) * 16) >> 4) & 0x3FFFu) | (static_cast<uint32_t>(kPlaneBytes >> 4) << 16);
    const uint32_t b_lo_base = (ptx::smem_u32(s_ring) >> 4) & 0x3FFFu;
    uint32_t s = 0, ph = 0, round_seq = 0, stem_idx = 0, stem_full[2] = {0u, 0u};
    for (int rd = cl0; rd < nrounds; rd += cl_stride, ++round_seq) {
      for (uint32_t c = 0; c < nchunks; ++c) {
        const Chunk ch = s_chunks[c];
        const bool tri = prm.prof != nullptr && blockIdx.x == 0 && lane == 0 && round_seq == 0 && (ch.flags & kFirstOfLayer);
        if (tri) prm.prof[16 + ch.layer * 8 + 5] = clock64();
        ptx::mbar_wait(&bar_full[s], ph);
        const uint32_t epoch = round_seq * nl + ch.layer;
        if ((ch.flags & kFirstOfSlice) && ch.layer == 0) {
          if (ch.flags & kFirstOfLayer) stem_idx = 0;
          const uint32_t hh = stem_idx & 1;
          ptx::mbar_wait(&bar_in_full[hh], stem_full[hh] & 1);
          ++stem_full[hh];
        }
        if (ch.need && epoch > 0) {
#pragma unroll
          for (int kb = 0; kb < G::kKBlocks; ++kb)
            if (ch.need & (1u << kb)) ptx::mbar_wait_cluster(&bar_act_part[kb], (epoch - 1) & 1);
        }
        ptx::tc_fence_after_sync();
        if (tri) prm.prof[16 + ch.layer * 8 + 6] = clock64();
        if (ptx::elect_one()) {
          const uint32_t nhalf = ch.ndim8 * 4u;                               // N of this CTA: half the layer's
          const uint32_t idesc = ptx::umma_idesc_bf16_f32(kTile, nhalf);
          const uint32_t d = tmem_base + (ch.layer & 1) * F;
          const int ntaps = (ch.flags & kSingleTap) ? 1 : 3;
          const uint32_t a_chunk = a_lo_base + static_cast<int>(ch.a_off) - (ntaps == 3 ? 1 : 0);
          const uint32_t b_step = ch.ndim8 * 8u;                              // 2 planes x (N/2 rows) x 16 B, in 16 B units
          uint32_t b_lo = (b_lo_base + s * (kStageBytes >> 4)) | (nhalf << 16);   // LBO = (N/2) x 16 B
          uint32_t acc = (ch.flags & kFirstOfLayer) ? 0u : 1u;
          constexpr uint32_t kStepK = 2 * (kPlaneBytes >> 4);
          constexpr uint32_t kBlock2 = (kPlanes / 2) * (kPlaneBytes >> 4);
          if (ntaps == 3 && ch.nk0 == 2 && ch.nk1 == 2) {
#pragma unroll
            for (int t = 0; t < 3; ++t)
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                umma_single(d, a_chunk + t + (k4 >> 1) * kBlock2 + (k4 & 1) * kStepK, b_lo, idesc, acc);
                acc = 1u;
                b_lo += b_step;
              }
          } else {
            for (int t = 0; t < ntaps; ++t) {
              for (int ks = 0; ks < ch.nk0; ++ks) {
                umma_single(d, a_chunk + t + ks * kStepK, b_lo, idesc, acc);
                acc = 1u;
                b_lo += b_step;
              }
              for (int ks = 0; ks < ch.nk1; ++ks) {
                umma_single(d, a_chunk + t + kBlock2 + ks * kStepK, b_lo, idesc, acc);
                acc = 1u;
                b_lo += b_step;
              }
            }
          }
          ptx::umma_commit(&bar_empty[s]);
          if ((ch.flags & kLastOfSlice) && ch.layer == 0 && stem_idx + 2 < static_cast<uint32_t>(prm.nslices))
            ptx::umma_commit(&bar_in_empty[stem_idx & 1]);
          if (ch.flags & kLastOfLayer) {
            ptx::umma_commit(bar_tmem_full);
            if (ch.layer == nl - 1 && !(prm.tail && full)) {   // round end: both halves are free for the next round's input
              ptx::umma_commit(&bar_in_empty[0]);
              ptx::umma_commit(&bar_in_empty[1]);
            }
          }
        }
        if ((ch.flags & kLastOfSlice) && ch.layer == 0) ++stem_idx;
        __syncwarp();
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps 0-7: TMEM lane quadrant q (32 tile rows), 32-channel half h of THIS CTA's 64 output channels ====
    const int q = warp & 3;
    const int h = warp >> 2;
    const int row = q * 32 + lane;
    const int by = row / kW, bp = (row % kW) / 6, bx = row % 6;
    const bool row_valid = row < kRowsValid && bx < 5;
    const int sq = by * 5 + bx;
    const int kb = 2 * static_cast<int>(rank) + h;          // the 32-channel K block this thread produces: channels 32 kb ..
    uint32_t res[16];                                        // block input of this row, these 32 channels, packed bf16x2
#pragma unroll
    for (int i = 0; i < 16; ++i) res[i] = 0u;
    uint32_t round_seq = 0, conv_cnt[2] = {0u, 0u};
    uint8_t* act_row = s_act + (kb * 4) * kPlaneBytes + (group_origin_row(0) + row) * 16;
    const uint32_t act_row_peer = ptx::mapa_u32(act_row, peer);

    for (int rd = cl0; rd < nrounds; rd += cl_stride, ++round_seq) {
      const int pos0 = rd * kPosPerGroup;
      // (nothing of the previous round is live in the residual registers: clearing them says so to the compiler, which
      // then has the registers to keep every load of a conversion pass in flight)
#pragma unroll
      for (int i = 0; i < 16; ++i) res[i] = 0u;
      // ---- the stem's input: every slice into this CTA's planes (both CTAs convert the whole input) ----
      for (int sl = 0; sl < prm.nslices; ++sl) {
        constexpr int kSlicePlanes = kPlanes / 2;
        const int hh = sl & 1;
        if (conv_cnt[hh] != 0) ptx::mbar_wait(&bar_in_empty[hh], (conv_cnt[hh] - 1) & 1);
        const int pl0 = sl * kSlicePlanes;
        const int npl = min(kSlicePlanes, prm.in_planes8 - pl0);
        constexpr int nitems = kPosPerGroup * kSlicePlanes * kSq;
        for (int it0 = threadIdx.x; it0 < nitems; it0 += 256 * 4) {
          float f[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int it = it0 + u * 256;
            const int qq = it / kSq, s_ = it - qq * kSq, c8 = qq % kSlicePlanes, p = qq / kSlicePlanes;
            const bool live = it < nitems && c8 < npl && pos0 + p < prm.n;
            const int ch0 = (pl0 + c8) * 8;
            if (prm.bits == nullptr) {
              const float* src = prm.planes + (static_cast<size_t>(pos0 + p) * prm.in_planes + ch0) * kSq + s_;
#pragma unroll
              for (int e = 0; e < 8; ++e) f[u][e] = (live && ch0 + e < prm.in_planes) ? __ldg(src + e * kSq) : 0.f;
            } else {
              const size_t o = static_cast<size_t>(pos0 + p) * prm.packed_stride + ch0;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const bool ok = live && ch0 + e < prm.in_planes;
                const uint32_t wbits = ok ? __ldg(prm.bits + o + e) : 0u;
                const float sc = ok ? __ldg(prm.scale + o + e) : 0.f;
                f[u][e] = ((wbits >> s_) & 1u) ? sc : 0.f;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int it = it0 + u * 256;
            const int qq = it / kSq, s_ = it - qq * kSq, c8 = qq % kSlicePlanes, p = qq / kSlicePlanes;
            if (it < nitems && c8 < npl) {
              const uint4 pk = make_uint4(ptx::pack_bf16x2(f[u][0], f[u][1]), ptx::pack_bf16x2(f[u][2], f[u][3]),
                                          ptx::pack_bf16x2(f[u][4], f[u][5]), ptx::pack_bf16x2(f[u][6], f[u][7]));
              const int rr = (s_ / 5) * kW + p * 6 + s_ % 5;
              *reinterpret_cast<uint4*>(s_act + (hh * kSlicePlanes + c8) * kPlaneBytes + (group_origin_row(0) + rr) * 16) = pk;
            }
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar_in_full[hh]);
        ++conv_cnt[hh];
      }
      // ---- the layers, one after the other ----
      for (int L = 0; L < nl; ++L) {
        const uint32_t par = (round_seq * nl + L) & 1;
        const bool tr = prm.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && round_seq == 0;
        if (tr) prm.prof[16 + L * 8 + 0] = clock64();
        ptx::mbar_wait(bar_tmem_full, par);
        ptx::tc_fence_after_sync();
        if (tr) prm.prof[16 + L * 8 + 1] = clock64();
        if (lane == 0) ptx::mbar_arrive_cluster(bar_peer, peer);    // "my CTA's MMAs of this layer are complete"
        const uint32_t lf = prm.layers[L].flags;
        const float* bias = s_bias + L * F;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (L & 1) * F + h * 32;
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr, v);
        ptx::tmem_ld_wait();
        if (!(lf & kFinal)) {
          const float* bb = bias + kb * 32;
          uint32_t o[16];
          float vacc = 0.f;
          const bool val = (lf & kValueTap) != 0 && full;
          if (val && h == 1) {                                      // continue the channel sum the lower half started
            ptx::mbar_wait(bar_s0, round_seq & 1);
            vacc = s_part[row];
          }
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float a0 = __uint_as_float(v[4 * j4 + 0]) + bb[4 * j4 + 0];
            float a1 = __uint_as_float(v[4 * j4 + 1]) + bb[4 * j4 + 1];
            float a2 = __uint_as_float(v[4 * j4 + 2]) + bb[4 * j4 + 2];
            float a3 = __uint_as_float(v[4 * j4 + 3]) + bb[4 * j4 + 3];
            if (lf & kHasRes) {
              const uint32_t r0 = res[2 * j4], r1 = res[2 * j4 + 1];
              a0 += bf16lo(r0); a1 += bf16hi(r0); a2 += bf16lo(r1); a3 += bf16hi(r1);
            }
            const uint32_t p0 = ptx::pack_relu_bf16x2(a0, a1), p1 = ptx::pack_relu_bf16x2(a2, a3);
            o[2 * j4] = p0; o[2 * j4 + 1] = p1;
            if (lf & kUpdRes) { res[2 * j4] = p0; res[2 * j4 + 1] = p1; }
            if (val) {
              const float4 wv = *reinterpret_cast<const float4*>(s_wv + kb * 32 + j4 * 4);
              vacc += wv.x * bf16lo(p0) + wv.y * bf16hi(p0) + wv.z * bf16lo(p1) + wv.w * bf16hi(p1);
            }
          }
          // the planes are overwritten in place: only once the PEER's tensor core has finished reading them too
          if (tr) prm.prof[16 + L * 8 + 2] = clock64();
          ptx::mbar_wait_cluster(bar_peer, par);
          if (tr) prm.prof[16 + L * 8 + 3] = clock64();
          if (row_valid) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              *reinterpret_cast<uint4*>(act_row + j4 * kPlaneBytes) = make_uint4(o[4 * j4], o[4 * j4 + 1], o[4 * j4 + 2], o[4 * j4 + 3]);
          }
          ptx::tc_fence_before_sync();
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // this warp's 32 rows of the block's four planes: to my own issuer at once, to the peer's planes as bulk copies
            // through the async proxy, which complete on the peer's K-block barrier (so its tensor core sees them)
            ptx::mbar_arrive(&bar_act_part[kb]);
            ptx::mbar_arrive_expect_tx_cluster(&bar_act_part[kb], peer, 4 * 32 * 16);
            const uint32_t bar_peer_kb = ptx::mapa_u32(&bar_act_part[kb], peer);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              ptx::bulk_s2s_cluster(act_row + j4 * kPlaneBytes, act_row_peer + j4 * kPlaneBytes, 32 * 16, bar_peer_kb);
            if (tr) prm.prof[16 + L * 8 + 4] = clock64();
          }
          if (val) {
            if (h == 0) {
              s_part[row] = vacc;
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(bar_s0);
            } else {
              // this CTA's 64-channel partial (exactly one epilogue thread's sum in the throughput kernel) -> rank 0
              if (row_valid) {
                float* dst = &s_v[(static_cast<int>(rank) * kPosPerGroup + bp) * kSq + sq];
                if (rank == 0) *dst = vacc; else ptx::st_cluster_f32(ptx::mapa_u32(dst, 0), vacc);
              }
              __syncwarp();
              if (lane == 0) { if (rank == 0) ptx::mbar_arrive(bar_v); else ptx::mbar_arrive_cluster_release(bar_v, 0); }
            }
          }
        } else {
          // policy 1x1: this CTA holds channels [48 rank, 48 rank + 48); this thread columns h*32 .. of them
          const int pos = pos0 + bp;
          if (!prm.tail) {
            const bool out_valid = row_valid && pos < prm.n && prm.policy != nullptr;
            float* dst = prm.policy + static_cast<size_t>(pos) * (kPolicyCh * kSq) + sq;
            if (out_valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int cl = h * 32 + j, c = (kPolicyN / 2) * static_cast<int>(rank) + cl;
                if (cl < kPolicyN / 2 && c < kPolicyCh) dst[c * kSq] = __uint_as_float(v[j]) + bias[c];
              }
            }
          } else {
            // staged in rank 0's planes for the policy tail: not before rank 0's tensor core has finished with them
            ptx::mbar_wait_cluster(bar_peer, par);
            uint8_t* stage_row = s_act + group_origin_row(0) * 16 + (bp * kSq + sq) * 4;
            const uint32_t stage_row0 = ptx::mapa_u32(stage_row, 0);
            if (row_valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int cl = h * 32 + j, c = (kPolicyN / 2) * static_cast<int>(rank) + cl;
                if (cl < kPolicyN / 2 && c < kPolicyCh) {
                  const float x = __uint_as_float(v[j]) + bias[c];
                  if (rank == 0) *reinterpret_cast<float*>(stage_row + stage_offset<F>(c)) = x;
                  else ptx::st_cluster_f32(stage_row0 + stage_offset<F>(c), x);
                }
              }
            }
            __syncwarp();
            if (lane == 0) { if (rank == 0) ptx::mbar_arrive(bar_p); else ptx::mbar_arrive_cluster_release(bar_p, 0); }
          }
          // no activations written, but keep the K-block barrier phases in step
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(&bar_act_part[kb]);
            ptx::mbar_arrive_cluster(&bar_act_part[kb], peer);
          }
        }
      }
      if (!full && prm.dbg != nullptr && rank == static_cast<uint32_t>(prm.lag & 1)) {   // (e55_debug_set_lag(1): the peer's copy)
        // debug dump: the planes after layer nl - 1, complete once all four K blocks (two of them the peer's) have landed
        const uint32_t par = (round_seq * nl + nl - 1) & 1;
        for (int k = 0; k < 4; ++k) ptx::mbar_wait_cluster(&bar_act_part[k], par);
        if (pos0 < prm.n)
          for (int i = threadIdx.x; i < kPlanes * kTile; i += 256) {
            const int pl = i / kTile, r_ = i % kTile;
            prm.dbg[(static_cast<size_t>(pos0 / kPosPerGroup) * kPlanes + pl) * kTile + r_] =
                *reinterpret_cast<const uint4*>(s_act + pl * kPlaneBytes + (group_origin_row(0) + r_) * 16);
          }
      }
      // ---- value head tail (rank 0, warp 0): ReLU(conv1x1) -> Dense(25->128)+ReLU -> Dense(128->1)+tanh (network.py:111-120) ----
      if (full && rank == 0 && warp == 0) {
        ptx::mbar_wait_cluster(bar_v, round_seq & 1);
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
            v1[p] = fmaxf(s_v[(0 * kPosPerGroup + p) * kSq + s_] + s_v[(1 * kPosPerGroup + p) * kSq + s_] + prm.value_b, 0.f);
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
        const int pos = pos0 + lane;
        if (lane < kPosPerGroup && pos < prm.n) {
          const float t = lane == 0 ? part[0] : lane == 1 ? part[1] : lane == 2 ? part[2] : part[3];
          prm.value[pos] = tanhf(t + prm.fc2_b);
        }
      }
    }
  }

  // neither CTA may exit (or free TMEM) while the other can still signal its barriers / write its shared memory
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 9) ptx::tmem_dealloc<kSplitTmemCols>(tmem_base);
}

}  // namespace tc
}  // namespace e55

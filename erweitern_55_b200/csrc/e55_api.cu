// libe55.so -- C ABI of the B200-native policy/value evaluator (see include/e55.h).
// Host side: context, Keras weight-list parsing, BatchNorm folding, weight images, launch plumbing.
#include "../../include/e55.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "e55_host.h"
#include "e55_fp32.cuh"
#include "e55_tc.cuh"
#include "e55_umma_test.cuh"
#include "e55_gpu_selfplay.cuh"

namespace {

constexpr int kSq = 25;
constexpr int kPolicyCh = 69;
constexpr int kValueHidden = 128;
constexpr double kBnEps = 1e-3;  // Keras BatchNormalization default epsilon

thread_local std::string g_create_error;

struct ConvF32 {
  float* w = nullptr;  // [taps][cin][cout], BN folded
  float* b = nullptr;  // [cout]
  int cin = 0, cout = 0, taps = 0;
};

}  // namespace

struct e55_service;

struct e55_ctx {
  e55_service* service = nullptr;   // cross-game leaf aggregation (e55_service.inl)
  int device = 0, blocks = 0, filters = 0, in_planes = 0;
  int capacity = 0;
  int precision = E55_PRECISION_BF16_TC;
  bool has_weights = false;
  int64_t launches = 0;
  cudaStream_t stream = nullptr;
  // host-buffer predict pipeline: chunks of the batch go H2D -> kernels -> D2H on rotating side streams
  static constexpr int kPipeStreams = 8;
  cudaStream_t pipe_stream[kPipeStreams] = {};
  cudaEvent_t pipe_done[kPipeStreams] = {};
  cudaEvent_t pipe_start = nullptr;
  mutable std::mutex mu;
  mutable std::string err;

  // fp32 weight images
  std::vector<ConvF32> convs;  // stem, 2 per block, policy conv (all 3x3)
  ConvF32 policy_out;          // 1x1 -> 69
  float* value_w = nullptr;    // [filters] folded value conv
  float value_b = 0.f;
  float* fc1_k = nullptr;      // [25][128]
  float* fc1_b = nullptr;      // [128]
  float* fc2_k = nullptr;      // [128]
  float fc2_b = 0.f;

  // tensor-core path state
  e55::tc::State tc;

  // optional CUDA-event timing of the dominant kernel (e55_set_kernel_timing)
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double timed_ms = 0.0;
  int64_t timed_launches = 0;
  bool ev_pending = false;

  // device work buffers (sized for `capacity` positions)
  float* d_in = nullptr;
  float* d_act[3] = {nullptr, nullptr, nullptr};
  float* d_policy = nullptr;
  float* d_value = nullptr;
  float* d_probs = nullptr;
  uint8_t* d_mask = nullptr;
  uint32_t* d_bits = nullptr;   // packed input staging: [capacity][in_planes]
  float* d_scale = nullptr;
  // host-side input compaction for e55_predict (e55_host.cpp): thread pool + pinned staging of the packed form
  int host_threads = -1;           // -1: chosen from the core count on first use; 0: off
  bool host_auto = true;           // the caller left the choice to the library: measure both pipelines, use the faster
  int64_t pipe_calls = 0;
  static constexpr int kSplitLevels = 8;
  static constexpr int kSplitEighths[kSplitLevels] = {0, 1, 2, 3, 4, 5, 6, 8};   // eighths of a large batch that travel as float planes
  double split_us[kSplitLevels] = {};        // running mean of microseconds per position at each split
  int split_samples[kSplitLevels] = {};
  int split_level = 0;                       // the split of the latest large call
  e55host::Packer* packer = nullptr;
  uint32_t* h_pack = nullptr;      // pinned [h_cap][bits in_planes | scale in_planes]: one copy per chunk takes both
  uint32_t* d_pack = nullptr;      // its device twin [d_pack_cap][2 * in_planes]
  int d_pack_cap = 0;
  int h_cap = 0;
  float* h_out_policy = nullptr;   // pinned staging of logits for pageable caller buffers: [h_out_cap][1725]
  int h_out_cap = 0;
  float* h_out_value = nullptr;    // pinned, device-mapped staging of values: the kernel writes them here directly
  void* d_out_value_mapped = nullptr;
  int h_out_value_cap = 0;
  std::vector<cudaEvent_t> chunk_ev;   // one per chunk of the host-buffer pipeline (grown on demand)
  // e55_predict_packed_moves: one pinned request block and one answer block per call, and their device twins
  uint8_t *h_req = nullptr, *h_ans = nullptr, *d_req = nullptr, *d_ans = nullptr;
  size_t req_cap = 0, ans_cap = 0;
  // compact legal-move tail (e55_predict_packed_moves): CSR move lists and their priors
  static constexpr int kMovesPerPosition = 128;   // capacity per position on average; a single position may use more
  int32_t* d_move_off = nullptr;   // [capacity + 1]
  uint16_t* d_move_idx = nullptr;  // [capacity * kMovesPerPosition]
  float* d_priors = nullptr;       // [capacity * kMovesPerPosition]
  // CUDA graphs of the small-batch host call (one per batch size seen twice): copy in + tower kernel, replayed
  struct SmallGraph { cudaGraphExec_t exec = nullptr; int calls = 0; };
  std::map<int, SmallGraph> small_graphs;
  // playout-cap oscillation of the GPU-resident self-play (e55_gpu_selfplay_set_cap; selfplay.py:45-61)
  int gpu_cap_enable = 0, gpu_cap_full = 800, gpu_cap_fast = 128;
  float gpu_cap_frac = 0.25f;
};

namespace {

int fail(const e55_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return code;
}

#define E55_CUDA(ctx, expr)                                                                    \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(ctx, E55_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));     \
  } while (0)

void free_conv(ConvF32& c) {
  cudaFree(c.w); cudaFree(c.b);
  c = ConvF32();
}

void free_weights(e55_ctx* c) {
  for (auto& cv : c->convs) free_conv(cv);
  c->convs.clear();
  free_conv(c->policy_out);
  cudaFree(c->value_w); cudaFree(c->fc1_k); cudaFree(c->fc1_b); cudaFree(c->fc2_k);
  c->value_w = c->fc1_k = c->fc1_b = c->fc2_k = nullptr;
  e55::tc::free_weights(c->tc);
  c->has_weights = false;
}

void free_buffers(e55_ctx* c) {
  cudaFree(c->d_in);
  for (auto& p : c->d_act) { cudaFree(p); p = nullptr; }
  cudaFree(c->d_policy); cudaFree(c->d_value); cudaFree(c->d_probs); cudaFree(c->d_mask);
  cudaFree(c->d_bits); cudaFree(c->d_scale);
  cudaFree(c->d_move_off); cudaFree(c->d_move_idx); cudaFree(c->d_priors);
  c->d_move_off = nullptr; c->d_move_idx = nullptr; c->d_priors = nullptr;
  c->d_in = c->d_policy = c->d_value = c->d_probs = c->d_scale = nullptr;
  c->d_mask = nullptr; c->d_bits = nullptr;
  e55::tc::free_buffers(c->tc);
  c->capacity = 0;
}

// The graphs hold device pointers, kernel parameters and the weight image: anything that moves those drops them.
void drop_small_graphs(e55_ctx* c) {
  for (auto& kv : c->small_graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  c->small_graphs.clear();
}

int ensure_capacity(e55_ctx* c, int n) {
  if (n <= c->capacity) return E55_OK;
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  drop_small_graphs(c);
  free_buffers(c);
  int cap = 64;
  while (cap < n) cap *= 2;
  const size_t act = static_cast<size_t>(cap) * c->filters * kSq * sizeof(float);
  E55_CUDA(c, cudaMalloc(&c->d_in, static_cast<size_t>(cap) * c->in_planes * kSq * sizeof(float)));
  for (auto& p : c->d_act) E55_CUDA(c, cudaMalloc(&p, act));
  E55_CUDA(c, cudaMalloc(&c->d_policy, static_cast<size_t>(cap) * kPolicyCh * kSq * sizeof(float)));
  E55_CUDA(c, cudaMalloc(&c->d_value, static_cast<size_t>(cap) * sizeof(float)));
  E55_CUDA(c, cudaMalloc(&c->d_probs, static_cast<size_t>(cap) * kPolicyCh * kSq * sizeof(float)));
  E55_CUDA(c, cudaMalloc(&c->d_mask, static_cast<size_t>(cap) * kPolicyCh * kSq));
  E55_CUDA(c, cudaMalloc(&c->d_bits, static_cast<size_t>(cap) * c->in_planes * sizeof(uint32_t)));
  E55_CUDA(c, cudaMalloc(&c->d_scale, static_cast<size_t>(cap) * c->in_planes * sizeof(float)));
  E55_CUDA(c, cudaMalloc(&c->d_move_off, (static_cast<size_t>(cap) + 1) * sizeof(int32_t)));
  E55_CUDA(c, cudaMalloc(&c->d_move_idx, static_cast<size_t>(cap) * e55_ctx::kMovesPerPosition * sizeof(uint16_t)));
  E55_CUDA(c, cudaMalloc(&c->d_priors, static_cast<size_t>(cap) * e55_ctx::kMovesPerPosition * sizeof(float)));
  cudaError_t e = e55::tc::alloc_buffers(c->tc, cap);
  if (e != cudaSuccess) return fail(c, E55_ERR_CUDA, std::string("tc buffers: ") + cudaGetErrorString(e));
  c->capacity = cap;
  return E55_OK;
}

// ---- Keras weight list -> named layers ----------------------------------------------------------
struct Tensor {
  const float* data;
  std::vector<int64_t> shape;
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};
struct BnRef { const float *gamma = nullptr, *beta = nullptr, *mean = nullptr, *var = nullptr; int c = 0; };
struct LayerRef {
  enum Kind { CONV, DENSE } kind;
  Tensor kernel, bias;
  BnRef bn;  // c == 0 when no BatchNorm follows
};

// Groups the flat list into conv/dense layers, each with the BatchNorm quadruple that follows it.
bool group_layers(const std::vector<Tensor>& ts, std::vector<LayerRef>& out, std::string& why) {
  size_t i = 0;
  while (i < ts.size()) {
    const Tensor& k = ts[i];
    if (k.shape.size() == 4 || k.shape.size() == 2) {
      if (i + 1 >= ts.size() || ts[i + 1].shape.size() != 1 || ts[i + 1].shape[0] != k.shape.back()) {
        why = "kernel tensor " + std::to_string(i) + " is not followed by its bias";
        return false;
      }
      LayerRef l;
      l.kind = k.shape.size() == 4 ? LayerRef::CONV : LayerRef::DENSE;
      l.kernel = k; l.bias = ts[i + 1];
      out.push_back(l);
      i += 2;
    } else if (k.shape.size() == 1) {
      if (i + 3 >= ts.size()) { why = "truncated BatchNorm group at tensor " + std::to_string(i); return false; }
      const int64_t c = k.shape[0];
      for (int j = 1; j < 4; ++j)
        if (ts[i + j].shape.size() != 1 || ts[i + j].shape[0] != c) {
          why = "BatchNorm group at tensor " + std::to_string(i) + " has inconsistent shapes";
          return false;
        }
      // attach to the most recent conv with matching channel count that has no BN yet
      bool attached = false;
      for (size_t j = out.size(); j-- > 0;) {
        if (out[j].kind == LayerRef::CONV && out[j].bn.c == 0 && out[j].kernel.shape[3] == c) {
          out[j].bn = BnRef{ts[i].data, ts[i + 1].data, ts[i + 2].data, ts[i + 3].data, static_cast<int>(c)};
          attached = true;
          break;
        }
      }
      if (!attached) { why = "BatchNorm group at tensor " + std::to_string(i) + " has no convolution"; return false; }
      i += 4;
    } else {
      why = "tensor " + std::to_string(i) + " has unsupported rank";
      return false;
    }
  }
  return true;
}

// Fold inference BN into conv (double math): w'[t][ci][co] = w*s[co], b' = (b-mean)*s + beta,
// s = gamma / sqrt(var + eps).   Keras HWIO flattening == [tap][ci][co].
void fold(const LayerRef& l, std::vector<float>& w, std::vector<float>& b) {
  const int64_t cout = l.kernel.shape.back();
  const int64_t rows = l.kernel.numel() / cout;
  w.resize(rows * cout);
  b.resize(cout);
  std::vector<double> s(cout, 1.0);
  if (l.bn.c)
    for (int64_t o = 0; o < cout; ++o) s[o] = static_cast<double>(l.bn.gamma[o]) / std::sqrt(static_cast<double>(l.bn.var[o]) + kBnEps);
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t o = 0; o < cout; ++o) w[r * cout + o] = static_cast<float>(static_cast<double>(l.kernel.data[r * cout + o]) * s[o]);
  for (int64_t o = 0; o < cout; ++o) {
    double v = l.bias.data[o];
    if (l.bn.c) v = (v - l.bn.mean[o]) * s[o] + l.bn.beta[o];
    b[o] = static_cast<float>(v);
  }
}

cudaError_t upload(float** dst, const std::vector<float>& src) {
  cudaError_t e = cudaMalloc(dst, src.size() * sizeof(float));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dst, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice);
}

// ---- fp32 forward -------------------------------------------------------------------------------
int launch_conv3x3(e55_ctx* c, const ConvF32& cv, const float* in, const float* res, float* out, int n, int relu) {
  if (n <= 148) {
    e55::fp32::conv3x3_kernel<1><<<n, cv.cout, 0, c->stream>>>(in, cv.w, cv.b, res, out, n, cv.cin, cv.cout, relu);
  } else {
    e55::fp32::conv3x3_kernel<4><<<(n + 3) / 4, cv.cout, 0, c->stream>>>(in, cv.w, cv.b, res, out, n, cv.cin, cv.cout, relu);
  }
  ++c->launches;
  E55_CUDA(c, cudaGetLastError());
  return E55_OK;
}

// trunk_out (optional) receives a pointer to the trunk output buffer.
int forward_fp32(e55_ctx* c, const float* d_planes, int n, float* d_policy, float* d_value, const float** trunk_out) {
  float* x = c->d_act[0];
  float* y = c->d_act[1];
  float* z = c->d_act[2];
  int rc = launch_conv3x3(c, c->convs[0], d_planes, nullptr, x, n, 1);  // network.py:92-95
  if (rc) return rc;
  for (int b = 0; b < c->blocks; ++b) {                                 // network.py:98-99,124-146
    if ((rc = launch_conv3x3(c, c->convs[1 + 2 * b], x, nullptr, y, n, 1))) return rc;
    if ((rc = launch_conv3x3(c, c->convs[2 + 2 * b], y, x, z, n, 1))) return rc;
    float* t = x; x = z; z = t;
  }
  if (trunk_out) *trunk_out = x;
  if ((rc = launch_conv3x3(c, c->convs[1 + 2 * c->blocks], x, nullptr, y, n, 1))) return rc;  // network.py:102-105
  const size_t act_smem = static_cast<size_t>(c->filters) * kSq * sizeof(float);
  e55::fp32::policy_out_kernel<<<n, 256, act_smem, c->stream>>>(y, c->policy_out.w, c->policy_out.b, d_policy,
                                                                c->filters, kPolicyCh);
  ++c->launches;
  E55_CUDA(c, cudaGetLastError());
  e55::fp32::value_head_kernel<<<n, 128, act_smem, c->stream>>>(x, c->value_w, c->value_b, c->fc1_k, c->fc1_b,
                                                                c->fc2_k, c->fc2_b, d_value, c->filters);
  ++c->launches;
  E55_CUDA(c, cudaGetLastError());
  return E55_OK;
}

// Folds the previous launch's event pair into the running total (events are reused, so this blocks
// until that launch has finished; timing mode is for measurement runs only).
int collect_timing(e55_ctx* c) {
  if (!c->ev_pending) return E55_OK;
  float ms = 0.f;
  E55_CUDA(c, cudaEventSynchronize(c->ev1));
  E55_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->timed_ms += ms;
  ++c->timed_launches;
  c->ev_pending = false;
  return E55_OK;
}

int forward_tc_on(e55_ctx* c, cudaStream_t stream, const float* d_planes, int n, float* d_policy, float* d_value,
                  e55::tc::PolicyTail tail = e55::tc::PolicyTail()) {
  std::string why;
  int launched = 0;
  c->tc.ev0 = c->tc.ev1 = nullptr;
  cudaError_t e = e55::tc::forward(c->tc, d_planes, n, d_policy, d_value, stream, &launched, &why, tail);
  c->launches += launched;
  if (e != cudaSuccess) return fail(c, E55_ERR_CUDA, "tensor-core forward: " + why + ": " + cudaGetErrorString(e));
  if (!why.empty()) return fail(c, E55_ERR_UNSUPPORTED, why);
  return E55_OK;
}

int forward(e55_ctx* c, const float* d_planes, int n, float* d_policy, float* d_value) {
  if (c->precision == E55_PRECISION_FP32) return forward_fp32(c, d_planes, n, d_policy, d_value, nullptr);
  std::string why;
  int launched = 0;
  if (c->timing) {
    int rc = collect_timing(c);
    if (rc) return rc;
    c->tc.ev0 = c->ev0; c->tc.ev1 = c->ev1;
    c->ev_pending = true;
  } else {
    c->tc.ev0 = c->tc.ev1 = nullptr;
  }
  cudaError_t e = e55::tc::forward(c->tc, d_planes, n, d_policy, d_value, c->stream, &launched, &why);
  c->launches += launched;
  if (e != cudaSuccess) return fail(c, E55_ERR_CUDA, "tensor-core forward: " + why + ": " + cudaGetErrorString(e));
  if (!why.empty()) return fail(c, E55_ERR_UNSUPPORTED, why);
  return E55_OK;
}

// Forward from packed device input on `stream`: the tcgen05 kernel expands the bits itself; the fp32 path
// expands them into the context's plane buffer first (main stream only).
int forward_packed_on(e55_ctx* c, cudaStream_t stream, const uint32_t* d_bits, const float* d_scale, int n, int plane_off,
                      float* d_policy, float* d_value, int stride = 0, e55::tc::PolicyTail tail = e55::tc::PolicyTail()) {
  if (c->precision == E55_PRECISION_FP32) {
    if (stride || tail.priors || tail.probs) return fail(c, E55_ERR_UNSUPPORTED, "interleaved packed input and the fused policy tail belong to the tensor-core path");
    const size_t total = static_cast<size_t>(n) * c->in_planes * kSq;
    float* planes = c->d_in + static_cast<size_t>(plane_off) * c->in_planes * kSq;
    e55::fp32::unpack_planes_kernel<<<static_cast<unsigned>(std::min<size_t>((total + 255) / 256, 4096)), 256, 0, stream>>>(
        d_bits, d_scale, planes, total);
    ++c->launches;
    E55_CUDA(c, cudaGetLastError());
    return forward_fp32(c, planes, n, d_policy, d_value, nullptr);
  }
  std::string why;
  int launched = 0;
  c->tc.ev0 = c->tc.ev1 = nullptr;
  cudaError_t e = e55::tc::forward_packed(c->tc, d_bits, d_scale, n, d_policy, d_value, stream, &launched, &why, stride, tail);
  c->launches += launched;
  if (e != cudaSuccess) return fail(c, E55_ERR_CUDA, "tensor-core forward: " + why + ": " + cudaGetErrorString(e));
  if (!why.empty()) return fail(c, E55_ERR_UNSUPPORTED, why);
  return E55_OK;
}

// Chunk size of the pipelined host-buffer predict: full chunks while much is left, then halves (multiples of 64).
int pipe_chunk(int left, int chunk) {
  if (left > chunk + chunk / 2) return std::min(chunk, left);
  const int half = std::max(64, ((left / 2 + 63) / 64) * 64);
  return half >= left ? left : half;
}

void sync_pipe_streams(e55_ctx* c) {
  for (int i = 0; i < e55_ctx::kPipeStreams; ++i)
    if (c->pipe_stream[i]) cudaStreamSynchronize(c->pipe_stream[i]);
  cudaStreamSynchronize(c->stream);
}

constexpr int kSmallBatch = 64;   // below this a host-buffer predict is one copy in, one kernel, one copy out

// Pinned staging of the packed form of `n` positions, and its device twin.
int ensure_host_staging(e55_ctx* c, int n) {
  if (n <= c->h_cap && n <= c->d_pack_cap) return E55_OK;
  sync_pipe_streams(c);
  drop_small_graphs(c);
  cudaFreeHost(c->h_pack); cudaFree(c->d_pack);
  c->h_pack = nullptr; c->d_pack = nullptr; c->h_cap = 0; c->d_pack_cap = 0;
  int cap = 256;
  while (cap < n) cap *= 2;
  const size_t bytes = static_cast<size_t>(cap) * 2 * c->in_planes * sizeof(uint32_t);
  if (cudaMallocHost(&c->h_pack, bytes) != cudaSuccess || cudaMalloc(&c->d_pack, bytes) != cudaSuccess) {
    const std::string msg = std::string("packed staging: ") + cudaGetErrorString(cudaGetLastError());
    cudaFreeHost(c->h_pack); cudaFree(c->d_pack);
    c->h_pack = nullptr; c->d_pack = nullptr;
    return fail(c, E55_ERR_CUDA, msg);
  }
  c->h_cap = cap; c->d_pack_cap = cap;
  return E55_OK;
}

// The packing pool and pinned staging for `n` positions (null: compaction is off or unavailable).
e55host::Packer* host_packer(e55_ctx* c, int n) {
  if (c->host_threads < 0) {
    cpu_set_t set;
    int hw = static_cast<int>(std::thread::hardware_concurrency());
    if (sched_getaffinity(0, sizeof(set), &set) == 0) hw = CPU_COUNT(&set);   // the cores this process may actually use
    int t = hw - 1;                                                           // the calling thread packs too while it waits
    if (const char* env = getenv("E55_HOST_THREADS")) t = atoi(env);
    c->host_threads = std::max(0, std::min(t, 31));
  }
  if (c->host_threads == 0) return nullptr;
  if (c->packer && e55host::packer_threads(c->packer) != c->host_threads) { e55host::packer_destroy(c->packer); c->packer = nullptr; }
  if (!c->packer) c->packer = e55host::packer_create(c->host_threads);
  if (c->packer && ensure_host_staging(c, n) != E55_OK) return nullptr;
  return c->packer;
}

int check_predict_args(e55_ctx* c, const void* a, const void* b, const void* d, int n) {
  if (!c) return E55_ERR_INVALID;
  if (!a || !b || !d) return fail(c, E55_ERR_INVALID, "null buffer");
  if (n <= 0) return fail(c, E55_ERR_INVALID, "batch size must be >= 1");
  if (!c->has_weights) return fail(c, E55_ERR_NO_WEIGHTS, "predict called before e55_set_weights");
  return E55_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int e55_version(void) { return 1; }

const char* e55_last_error(const e55_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int e55_create(e55_ctx** out, int device, int blocks, int filters, int in_planes, int max_batch) {
  if (!out) return fail(nullptr, E55_ERR_INVALID, "out is null");
  *out = nullptr;
  if (blocks < 0 || blocks > 64 || in_planes < 1 || in_planes > 4096 || max_batch < 1)
    return fail(nullptr, E55_ERR_INVALID, "bad architecture arguments");
  if (filters != 128 && filters != 256)
    return fail(nullptr, E55_ERR_UNSUPPORTED, "filters must be 128 or 256");
  int ndev = 0;
  // the host-buffer pipeline keeps up to eight chunk streams busy: with the default of 8 hardware queues, streams share
  // queues and a late chunk waits behind an early chunk's read-back (seen as ~100 us per batch-1024 call); only
  // effective when this is the process's first CUDA use, hence also set by the Python package on import
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, E55_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, E55_ERR_INVALID, "device index out of range");
  cudaDeviceProp prop;
  E55_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, E55_ERR_NO_DEVICE, std::string("device is not sm_100 (Blackwell B200): ") + prop.name);
  E55_CUDA(nullptr, cudaSetDevice(device));
  e55_ctx* c = new (std::nothrow) e55_ctx();
  if (!c) return fail(nullptr, E55_ERR_INVALID, "out of host memory");
  c->device = device; c->blocks = blocks; c->filters = filters; c->in_planes = in_planes;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return fail(nullptr, E55_ERR_CUDA, cudaGetErrorString(e)); }
  for (int i = 0; i < e55_ctx::kPipeStreams && e == cudaSuccess; ++i) {
    e = cudaStreamCreateWithFlags(&c->pipe_stream[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->pipe_done[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->pipe_start, cudaEventDisableTiming);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); e55_destroy(c); return E55_ERR_CUDA; }
  e55::tc::init(c->tc, blocks, filters, in_planes, prop.multiProcessorCount);
  int rc = ensure_capacity(c, max_batch);
  if (rc != E55_OK) { g_create_error = c->err; e55_destroy(c); return rc; }
  *out = c;
  return E55_OK;
}

void e55_destroy(e55_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  e55_service_stop(c);
  if (c->stream) cudaStreamSynchronize(c->stream);
  e55host::packer_destroy(c->packer);
  drop_small_graphs(c);
  cudaFreeHost(c->h_pack); cudaFree(c->d_pack);
  cudaFreeHost(c->h_out_policy); cudaFreeHost(c->h_out_value);
  for (cudaEvent_t ev : c->chunk_ev) cudaEventDestroy(ev);
  cudaFreeHost(c->h_req); cudaFreeHost(c->h_ans); cudaFree(c->d_req); cudaFree(c->d_ans);
  free_weights(c);
  free_buffers(c);
  if (c->ev0) { cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); }
  if (c->tc.d_prof) cudaFree(c->tc.d_prof);
  for (int i = 0; i < e55_ctx::kPipeStreams; ++i) {
    if (c->pipe_stream[i]) { cudaStreamSynchronize(c->pipe_stream[i]); cudaStreamDestroy(c->pipe_stream[i]); }
    if (c->pipe_done[i]) cudaEventDestroy(c->pipe_done[i]);
  }
  if (c->pipe_start) cudaEventDestroy(c->pipe_start);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int e55_set_weights(e55_ctx* c, int n_tensors, const float* const* data, const int64_t* const* shapes,
                    const int* ndims) {
  if (!c) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!data || !shapes || !ndims || n_tensors <= 0) return fail(c, E55_ERR_INVALID, "null weight list");
  std::vector<Tensor> ts(n_tensors);
  for (int i = 0; i < n_tensors; ++i) {
    if (!data[i] || !shapes[i] || ndims[i] < 1 || ndims[i] > 4)
      return fail(c, E55_ERR_INVALID, "weight tensor " + std::to_string(i) + " is malformed");
    ts[i].data = data[i];
    ts[i].shape.assign(shapes[i], shapes[i] + ndims[i]);
  }
  std::vector<LayerRef> layers;
  std::string why;
  if (!group_layers(ts, layers, why)) return fail(c, E55_ERR_INVALID, why);

  // trunk: the first 1 + 2*blocks convolutions, in order (network.py:92-99)
  const int n_trunk = 1 + 2 * c->blocks;
  if (static_cast<int>(layers.size()) != n_trunk + 5)
    return fail(c, E55_ERR_INVALID, "expected " + std::to_string(n_trunk + 5) + " conv/dense layers, got " +
                                        std::to_string(layers.size()));
  const int F = c->filters;
  auto is_conv = [](const LayerRef& l, int kh, int64_t cin, int64_t cout, bool bn) {
    return l.kind == LayerRef::CONV && l.kernel.shape[0] == kh && l.kernel.shape[1] == kh &&
           l.kernel.shape[2] == cin && l.kernel.shape[3] == cout && (l.bn.c != 0) == bn;
  };
  for (int i = 0; i < n_trunk; ++i)
    if (!is_conv(layers[i], 3, i == 0 ? c->in_planes : F, F, true))
      return fail(c, E55_ERR_INVALID, "trunk convolution " + std::to_string(i) + " has the wrong shape or no BatchNorm");
  // heads: identified by shape, any interleaving (Keras orders them by graph depth)
  const LayerRef *value_conv = nullptr, *policy_conv = nullptr, *policy_out = nullptr, *fc1 = nullptr, *fc2 = nullptr;
  for (size_t i = n_trunk; i < layers.size(); ++i) {
    const LayerRef& l = layers[i];
    if (is_conv(l, 1, F, 1, true) && !value_conv) value_conv = &l;
    else if (is_conv(l, 3, F, F, true) && !policy_conv) policy_conv = &l;
    else if (is_conv(l, 1, F, kPolicyCh, false) && !policy_out) policy_out = &l;
    else if (l.kind == LayerRef::DENSE && l.kernel.shape[0] == kSq && l.kernel.shape[1] == kValueHidden && !fc1) fc1 = &l;
    else if (l.kind == LayerRef::DENSE && l.kernel.shape[0] == kValueHidden && l.kernel.shape[1] == 1 && !fc2) fc2 = &l;
    else return fail(c, E55_ERR_INVALID, "unrecognised head layer at position " + std::to_string(i));
  }
  if (!value_conv || !policy_conv || !policy_out || !fc1 || !fc2)
    return fail(c, E55_ERR_INVALID, "a head layer is missing from the weight list");

  E55_CUDA(c, cudaSetDevice(c->device));
  // nothing may still read the old weight images: the context stream, the host pipeline's chunk streams and (the
  // service's batches run on streams of their own) the whole device
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < e55_ctx::kPipeStreams; ++i) E55_CUDA(c, cudaStreamSynchronize(c->pipe_stream[i]));
  E55_CUDA(c, cudaDeviceSynchronize());
  drop_small_graphs(c);
  free_weights(c);

  std::vector<float> w, b;
  std::vector<std::vector<float>> conv_w, conv_b;  // kept for the tensor-core packer
  auto add_conv3 = [&](const LayerRef& l) -> cudaError_t {
    fold(l, w, b);
    ConvF32 cv;
    cv.cin = static_cast<int>(l.kernel.shape[2]); cv.cout = static_cast<int>(l.kernel.shape[3]); cv.taps = 9;
    cudaError_t e = upload(&cv.w, w);
    if (e == cudaSuccess) e = upload(&cv.b, b);
    c->convs.push_back(cv);
    conv_w.push_back(w); conv_b.push_back(b);
    return e;
  };
  for (int i = 0; i < n_trunk; ++i) E55_CUDA(c, add_conv3(layers[i]));
  E55_CUDA(c, add_conv3(*policy_conv));

  std::vector<float> pw, pb;
  fold(*policy_out, pw, pb);
  c->policy_out.cin = F; c->policy_out.cout = kPolicyCh; c->policy_out.taps = 1;
  E55_CUDA(c, upload(&c->policy_out.w, pw));
  E55_CUDA(c, upload(&c->policy_out.b, pb));

  std::vector<float> vw, vb;
  fold(*value_conv, vw, vb);
  E55_CUDA(c, upload(&c->value_w, vw));
  c->value_b = vb[0];
  std::vector<float> k1(fc1->kernel.data, fc1->kernel.data + kSq * kValueHidden);
  std::vector<float> b1(fc1->bias.data, fc1->bias.data + kValueHidden);
  std::vector<float> k2(fc2->kernel.data, fc2->kernel.data + kValueHidden);
  E55_CUDA(c, upload(&c->fc1_k, k1));
  E55_CUDA(c, upload(&c->fc1_b, b1));
  E55_CUDA(c, upload(&c->fc2_k, k2));
  c->fc2_b = fc2->bias.data[0];

  e55::tc::HeadWeights hw{pw.data(), pb.data(), c->value_w, c->value_b, c->fc1_k, c->fc1_b, c->fc2_k, c->fc2_b};
  cudaError_t e = e55::tc::set_weights(c->tc, conv_w, conv_b, hw);
  if (e != cudaSuccess) return fail(c, E55_ERR_CUDA, std::string("tensor-core weight image: ") + cudaGetErrorString(e));
  c->has_weights = true;
  return E55_OK;
}

int e55_set_precision(e55_ctx* c, int precision) {
  if (!c) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  if (precision != E55_PRECISION_FP32 && precision != E55_PRECISION_BF16_TC)
    return fail(c, E55_ERR_INVALID, "unknown precision");
  if (precision == E55_PRECISION_BF16_TC && !e55::tc::supported(c->tc))
    return fail(c, E55_ERR_UNSUPPORTED, "the tcgen05 path supports 128 or 256 filters");
  drop_small_graphs(c);
  c->precision = precision;
  return E55_OK;
}

int e55_get_precision(const e55_ctx* c, int* precision) {
  if (!c || !precision) return E55_ERR_INVALID;
  *precision = c->precision;
  return E55_OK;
}

int e55_predict_device(e55_ctx* c, const float* d_planes, int n, float* d_policy, float* d_value) {
  int rc = check_predict_args(c, d_planes, d_policy, d_value, n);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  if ((rc = ensure_capacity(c, n))) return rc;
  return forward(c, d_planes, n, d_policy, d_value);
}

namespace {

constexpr int kPipeChunk = 256;

// True if `p` is page-locked host memory the copy engines can reach directly (cudaHostAlloc / cudaHostRegister).
bool host_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// Device-side address of page-locked host memory (null: not mapped, or direct writes are switched off).  The tower
// kernel's policy tail writes whole 6.9 KB rows, so logits can go straight over PCIe into the caller's array: no
// read-back copy, and nothing waits behind a copy engine.  E55_DIRECT_HOST=0 restores the copies (for comparison).
float* mapped_device_pointer(void* host) {
  static const bool direct = !(getenv("E55_DIRECT_HOST") && !strcmp(getenv("E55_DIRECT_HOST"), "0"));
  void* d = nullptr;
  if (!direct || cudaHostGetDevicePointer(&d, host, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return static_cast<float*>(d);
}

// Pinned staging of results.  Values (4 bytes per position) always go through it: the kernel writes them straight into
// this page-locked block (it is device-mapped: no read-back copy at all) and the call copies them to the caller's array
// at the end.  Logits go through it only for callers whose policy buffer is pageable (malloc, plain numpy).
int ensure_out_staging(e55_ctx* c, int n, bool need_policy) {
  int cap = 256;
  while (cap < n) cap *= 2;
  if (n > c->h_out_value_cap) {
    sync_pipe_streams(c);
    drop_small_graphs(c);
    cudaFreeHost(c->h_out_value);
    c->h_out_value = nullptr; c->h_out_value_cap = 0;
    E55_CUDA(c, cudaHostAlloc(&c->h_out_value, static_cast<size_t>(cap) * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
    E55_CUDA(c, cudaHostGetDevicePointer(&c->d_out_value_mapped, c->h_out_value, 0));
    c->h_out_value_cap = cap;
  }
  if (need_policy && n > c->h_out_cap) {
    sync_pipe_streams(c);
    drop_small_graphs(c);
    cudaFreeHost(c->h_out_policy);
    c->h_out_policy = nullptr; c->h_out_cap = 0;
    E55_CUDA(c, cudaMallocHost(&c->h_out_policy, static_cast<size_t>(cap) * kPolicyCh * kSq * sizeof(float)));
    c->h_out_cap = cap;
  }
  return E55_OK;
}

// Host-buffer predict as a pipeline of chunks on rotating side streams: H2D -> tower kernel -> D2H, so that PCIe
// transfers in both directions overlap each other and the kernels (which run side by side on disjoint SMs).
//   * The first n - n_float positions are bit-packed on the host chunk by chunk (exactly; e55_host.cpp) while earlier
//     chunks are already on the GPU, so 2 128 instead of 26 600 bytes per position cross PCIe.  A chunk that is not
//     exactly representable travels as float planes.
//   * The last n_float positions are not packed at all but sent as float planes in between the packed chunks -- the
//     copy engine moves them while the cores pack the rest, so both the PCIe link and the host cores are busy for the
//     whole call (pinned inputs only: a pageable buffer is read by the cores either way).  n_float == n is the plain
//     float pipeline.
//   * Results go straight into the caller's buffers when those are pinned; otherwise into the library's pinned staging,
//     from where the pool copies each chunk out as soon as its event has fired, while later chunks are still on the GPU.
// Results are identical whichever way a position travels.
int predict_host_pipeline(e55_ctx* c, e55host::Packer* packer, const float* planes, int n, int n_float, float* policy, float* value,
                          bool pol_pinned) {
  const int n_pack = n - n_float;
  const size_t ip = static_cast<size_t>(c->in_planes), in_pp = ip * kSq, pol_pp = static_cast<size_t>(kPolicyCh) * kSq;
  int rc = E55_OK;
  if ((rc = ensure_out_staging(c, n, !pol_pinned))) return rc;
  float* out_policy = pol_pinned ? policy : c->h_out_policy;
  float* direct_policy = mapped_device_pointer(out_policy);      // non-null: the kernels write the logits into it themselves
  float* d_value = static_cast<float*>(c->d_out_value_mapped);   // written by the kernels over PCIe, 4 bytes per position
  e55::tc::PolicyTail tail;
  tail.rows = direct_policy != nullptr;
  const size_t ip2 = 2 * ip;
  static const bool trace = getenv("E55_HOST_TRACE") != nullptr;   // prints where a call's time goes (host clock, us)
  const auto t_begin = std::chrono::steady_clock::now();
  auto now_us = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count(); };
  std::vector<double> t_sent;
  std::vector<cudaEvent_t> tev;   // trace mode: [chunk][before H2D, after H2D, after kernel, after D2H] (+ one at the start)
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t ev = nullptr;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, st);
    tev.push_back(ev);
  };
  if (n_pack > 0) e55host::packer_start(packer, planes, n_pack, c->in_planes, c->h_pack, reinterpret_cast<float*>(c->h_pack) + ip, static_cast<int>(ip2));
  mark(c->stream);
  cudaError_t e = cudaEventRecord(c->pipe_start, c->stream);
  struct Sent { int off, m; };
  std::vector<Sent> sent;
  int used = 0;
  auto send = [&](int off, int m, bool packed, bool full_chunk) {
    const int chunk = static_cast<int>(sent.size());
    cudaStream_t st = c->pipe_stream[chunk % e55_ctx::kPipeStreams];
    if (chunk < e55_ctx::kPipeStreams) { e = cudaStreamWaitEvent(st, c->pipe_start, 0); used = chunk + 1; }
    while (e == cudaSuccess && static_cast<int>(c->chunk_ev.size()) <= chunk) {
      cudaEvent_t ev = nullptr;
      e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e == cudaSuccess) c->chunk_ev.push_back(ev);
    }
    c->tc.throughput_rounds = full_chunk;   // full chunks run side by side on few SMs each; a short tail spreads out
    mark(st);
    if (packed) {
      if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_pack + off * ip2, c->h_pack + off * ip2, m * ip2 * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
      mark(st);
      if (e == cudaSuccess) rc = forward_packed_on(c, st, c->d_pack + off * ip2, reinterpret_cast<const float*>(c->d_pack + off * ip2) + ip, m, off,
                                                   (direct_policy ? direct_policy : c->d_policy) + off * pol_pp, d_value + off, static_cast<int>(ip2), tail);
    } else {
      if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_in + off * in_pp, planes + off * in_pp, m * in_pp * sizeof(float), cudaMemcpyHostToDevice, st);
      mark(st);
      if (e == cudaSuccess) rc = forward_tc_on(c, st, c->d_in + off * in_pp, m, (direct_policy ? direct_policy : c->d_policy) + off * pol_pp, d_value + off, tail);
    }
    mark(st);
    c->tc.throughput_rounds = false;
    if (e == cudaSuccess && rc == E55_OK && !direct_policy)
      e = cudaMemcpyAsync(out_policy + off * pol_pp, c->d_policy + off * pol_pp, m * pol_pp * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == E55_OK) e = cudaEventRecord(c->chunk_ev[chunk], st);
    mark(st);
    sent.push_back(Sent{off, m});
    if (trace) t_sent.push_back(now_us());
  };
  // Packed chunks.  The read-back of the logits (6.9 KB per position) is the longest stage of a large call, so results
  // have to start flowing early and keep flowing: two chunks of 64 positions first, then chunks of 128, each a kernel in
  // latency mode (one 4-position group per CTA: 16 / 32 CTAs, ~76 us), up to four of which run side by side.
  static const int first_chunk = getenv("E55_FIRST_CHUNK") ? std::max(64, atoi(getenv("E55_FIRST_CHUNK"))) : 64;
  static const int body_chunk = getenv("E55_BODY_CHUNK") ? std::max(64, atoi(getenv("E55_BODY_CHUNK"))) : 128;
  constexpr int kFloatChunk = 128;   // 3.4 MB: ~65 us on the link, about what packing the next chunk takes
  int off = 0, foff = n_pack;
  while ((off < n_pack || foff < n) && e == cudaSuccess && rc == E55_OK) {
    // the float share keeps step with the packed share (and goes first: nothing has to be waited for)
    while (foff < n && e == cudaSuccess && rc == E55_OK &&
           (off >= n_pack || static_cast<int64_t>(foff - n_pack) * n_pack <= static_cast<int64_t>(off) * n_float)) {
      // all floats: full chunks while much is left, then halving chunks, so that what remains after the last byte
      // has arrived -- one kernel and one read-back -- is short
      const int m = n_pack == 0 ? pipe_chunk(n - foff, kPipeChunk) : std::min(kFloatChunk, n - foff);
      send(foff, m, false, n_pack == 0 && m >= kPipeChunk);
      foff += m;
    }
    if (off < n_pack && e == cudaSuccess && rc == E55_OK) {
      const int left = n_pack - off;
      int m = off < 2 * first_chunk ? first_chunk : body_chunk;
      if (left < m + 32) m = left;   // no crumbs at the end
      send(off, m, e55host::packer_wait_range(packer, off, m), false);
      off += m;
    }
  }
  if (n_pack > 0) e55host::packer_finish(packer);   // the pool reads the caller's planes: it must be done before we return, whatever happened
  const double t_packed = trace ? now_us() : 0.0;
  if (rc != E55_OK || e != cudaSuccess) {
    sync_pipe_streams(c);   // nothing of this call may still be in flight when the caller gets its buffers back
    if (rc != E55_OK) return rc;
    E55_CUDA(c, e);
  }
  if (!pol_pinned) {
    for (size_t k = 0; k < sent.size() && e == cudaSuccess; ++k) {
      e = cudaEventSynchronize(c->chunk_ev[k]);
      if (e != cudaSuccess) break;
      e55host::packer_parallel_copy(packer, policy + sent[k].off * pol_pp, out_policy + sent[k].off * pol_pp, sent[k].m * pol_pp * sizeof(float));
    }
  }
  for (int i = 0; i < used && e == cudaSuccess; ++i) {
    e = cudaEventRecord(c->pipe_done[i], c->pipe_stream[i]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->pipe_done[i], 0);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) sync_pipe_streams(c);
  E55_CUDA(c, e);
  memcpy(value, c->h_out_value, static_cast<size_t>(n) * sizeof(float));
  if (trace) {
    fprintf(stderr, "[e55 host pipeline] n %d (%d as floats): chunks sent at", n, n_float);
    for (size_t k = 0; k < sent.size(); ++k) fprintf(stderr, " %.0f(%d)", t_sent[k], sent[k].m);
    fprintf(stderr, " us, pool done %.0f us, results home %.0f us\n", t_packed, now_us());
    if (tev.size() == 1 + 4 * sent.size()) {
      fprintf(stderr, "[e55 host pipeline]   device clock, us since the call's first event: chunk = H2D start..end, kernel end, D2H end:");
      for (size_t k = 0; k < sent.size(); ++k) {
        float t[4];
        for (int j = 0; j < 4; ++j) cudaEventElapsedTime(&t[j], tev[0], tev[1 + 4 * k + j]);
        fprintf(stderr, " [%d: %.0f..%.0f, %.0f, %.0f]", sent[k].m, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3);
      }
      fprintf(stderr, "\n");
    }
    for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
  }
  return E55_OK;
}

}  // namespace

int e55_predict(e55_ctx* c, const float* planes, int n, float* policy, float* value) {
  int rc = check_predict_args(c, planes, policy, value, n);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  if ((rc = ensure_capacity(c, n))) return rc;
  const bool tc_path = c->precision == E55_PRECISION_BF16_TC && !c->timing;
  if (tc_path && n >= kSmallBatch) {
    // Larger batch on the tensor-core path: chunks travel as a pipeline.  The cores pack one part of the batch, the rest
    // travels as floats meanwhile (pinned inputs).  Which split is fastest depends on the host (cores and memory bandwidth
    // to spare against the PCIe link), so unless the caller fixed the thread count the first calls try them all and later
    // ones take the fastest, looking at its neighbours again now and then.
    const bool in_pinned = host_pinned(planes), pol_pinned = host_pinned(policy);
    e55host::Packer* packer = host_packer(c, n);
    int level = packer ? 0 : e55_ctx::kSplitLevels - 1;   // index into kSplitEighths: how much of the batch travels as floats
    const bool adaptive = packer && c->host_auto && in_pinned && n >= 2 * kPipeChunk;
    if (adaptive) {
      const int64_t k = c->pipe_calls++;
      int best = 0;
      for (int i = 1; i < e55_ctx::kSplitLevels; ++i)
        if (c->split_us[i] < c->split_us[best]) best = i;
      if (k < 2 * e55_ctx::kSplitLevels) level = static_cast<int>(k % e55_ctx::kSplitLevels);   // every split twice, ...
      else if (k % 32 == 0) level = std::max(0, std::min(e55_ctx::kSplitLevels - 1, best + ((k / 32) % 2 ? 1 : -1)));   // ... a neighbour now and then, ...
      else level = best;                                                                                            // ... else the fastest
    }
    if (const char* env = (packer && in_pinned) ? getenv("E55_HOST_SPLIT") : nullptr)   // fixed share (eighths of the batch as floats), for measurements
      for (int i = 0; i < e55_ctx::kSplitLevels; ++i)
        if (e55_ctx::kSplitEighths[i] == atoi(env)) level = i;
    const int eighths = e55_ctx::kSplitEighths[level];
    const int n_float = eighths >= 8 ? n : (n / 8 * eighths) / 64 * 64;
    const auto t0 = std::chrono::steady_clock::now();
    rc = predict_host_pipeline(c, packer, planes, n, n_float, policy, value, pol_pinned);
    if (adaptive) {
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / n;
      double& avg = c->split_us[level];
      avg = c->split_samples[level]++ < 2 ? us : 0.75 * avg + 0.25 * us;   // (the first call of a kind also pays for its set-up)
      c->split_level = level;
    }
    return rc;
  }
  // Small batch (the reference's own: root 1, leaves 16; mcts.py:57-58,101-102): one copy in, one kernel, one copy out.
  // A pageable input is bit-packed by the calling thread into pinned staging on the way (2 KB instead of 26 KB per
  // position for the copy engine, and no driver-side staging copy); anything not exactly representable goes as floats.
  // The logits come back through pinned staging, the values are written into mapped host memory by the kernel.
  if (tc_path) {
    static const bool trace_small = getenv("E55_HOST_TRACE") != nullptr;
    const auto ts0 = std::chrono::steady_clock::now();
    auto us_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t).count(); };
    double t_checks = 0, t_pack = 0, t_issue = 0, t_wait = 0;
    const bool pol_pinned = host_pinned(policy);
    if ((rc = ensure_out_staging(c, n, !pol_pinned))) return rc;
    float* d_value = static_cast<float*>(c->d_out_value_mapped);
    float* out_policy = pol_pinned ? policy : c->h_out_policy;
    float* direct_policy = mapped_device_pointer(out_policy);   // the kernel writes the logits into host memory itself
    e55::tc::PolicyTail tail;
    tail.rows = direct_policy != nullptr;
    float* k_policy = direct_policy ? direct_policy : c->d_policy;
    const size_t ip = static_cast<size_t>(c->in_planes), ip2 = 2 * ip;
    const bool try_pack = c->host_threads != 0 && !host_pinned(planes) && ensure_host_staging(c, n) == E55_OK;
    if (trace_small) t_checks = us_since(ts0);
    if (try_pack &&
        e55host::pack_planes_inline(planes, n, c->in_planes, c->h_pack, reinterpret_cast<float*>(c->h_pack) + ip, static_cast<int>(ip2))) {
      if (trace_small) t_pack = us_since(ts0);
      // The search sends the same few batch sizes over and over (1 and 16), so copy + kernel can be replayed as one CUDA
      // graph from the second call of a size on (every pointer in it is library-owned staging: the graph stays valid).
      // Measured (tools/small_batch_probe.py, B200): batch 16 124.5 us per call with the graph against 118.7 us with the
      // two stream operations, batch 1 95-98 against 96 -- the call is two launches, there is nothing for a graph to
      // fold, and its launch costs more than they do.  Hence off unless E55_GRAPH=1.
      static const bool use_graphs = getenv("E55_GRAPH") && atoi(getenv("E55_GRAPH")) != 0;
      e55_ctx::SmallGraph* sg = (use_graphs && !pol_pinned && direct_policy) ? &c->small_graphs[n] : nullptr;
      if (sg && sg->exec) {
        E55_CUDA(c, cudaGraphLaunch(sg->exec, c->stream));
        ++c->launches;
      } else {
        const bool capture = sg && sg->calls >= 1;
        cudaGraph_t graph = nullptr;
        if (capture) E55_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t ce = cudaMemcpyAsync(c->d_pack, c->h_pack, n * ip2 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
        if (ce == cudaSuccess)
          rc = forward_packed_on(c, c->stream, c->d_pack, reinterpret_cast<const float*>(c->d_pack) + ip, n, 0, k_policy, d_value, static_cast<int>(ip2), tail);
        if (capture) {
          const cudaError_t ee = cudaStreamEndCapture(c->stream, &graph);
          if (ce == cudaSuccess) ce = ee;
          if (ce == cudaSuccess && rc == E55_OK) ce = cudaGraphInstantiate(&sg->exec, graph, 0);
          if (graph) cudaGraphDestroy(graph);
          if (ce == cudaSuccess && rc == E55_OK) ce = cudaGraphLaunch(sg->exec, c->stream);
        }
        if (rc) return rc;
        E55_CUDA(c, ce);
        if (sg) ++sg->calls;
      }
    } else {
      E55_CUDA(c, cudaMemcpyAsync(c->d_in, planes, static_cast<size_t>(n) * ip * kSq * sizeof(float), cudaMemcpyHostToDevice, c->stream));
      if ((rc = forward_tc_on(c, c->stream, c->d_in, n, k_policy, d_value, tail))) {
        cudaStreamSynchronize(c->stream);   // the copy may still be reading the caller's (page-locked) planes
        return rc;
      }
    }
    const size_t pol_bytes = static_cast<size_t>(n) * kPolicyCh * kSq * sizeof(float);
    if (!direct_policy) E55_CUDA(c, cudaMemcpyAsync(out_policy, c->d_policy, pol_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (trace_small) t_issue = us_since(ts0);
    E55_CUDA(c, cudaStreamSynchronize(c->stream));
    if (trace_small) t_wait = us_since(ts0);
    if (!pol_pinned) memcpy(policy, c->h_out_policy, pol_bytes);
    memcpy(value, c->h_out_value, static_cast<size_t>(n) * sizeof(float));
    if (trace_small)
      fprintf(stderr, "[e55 small call] n %d: checks %.1f, packed %.1f, issued %.1f, GPU done %.1f, results copied %.1f us\n", n, t_checks, t_pack,
              t_issue, t_wait, us_since(ts0));
    return E55_OK;
  }
  const size_t in_bytes = static_cast<size_t>(n) * c->in_planes * kSq * sizeof(float);
  E55_CUDA(c, cudaMemcpyAsync(c->d_in, planes, in_bytes, cudaMemcpyHostToDevice, c->stream));
  if ((rc = forward(c, c->d_in, n, c->d_policy, c->d_value))) {
    cudaStreamSynchronize(c->stream);
    return rc;
  }
  E55_CUDA(c, cudaMemcpyAsync(policy, c->d_policy, static_cast<size_t>(n) * kPolicyCh * kSq * sizeof(float),
                              cudaMemcpyDeviceToHost, c->stream));
  E55_CUDA(c, cudaMemcpyAsync(value, c->d_value, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost,
                              c->stream));
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  return E55_OK;
}

int e55_predict_packed_device(e55_ctx* c, const uint32_t* d_bits, const float* d_scale, int n, float* d_policy,
                              float* d_value) {
  int rc = check_predict_args(c, d_bits, d_policy, d_value, n);
  if (rc) return rc;
  if (!d_scale) return fail(c, E55_ERR_INVALID, "null scale buffer");
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  if ((rc = ensure_capacity(c, n))) return rc;
  return forward_packed_on(c, c->stream, d_bits, d_scale, n, 0, d_policy, d_value);
}

int e55_predict_packed(e55_ctx* c, const uint32_t* bits, const float* scale, int n, float* policy, float* value) {
  int rc = check_predict_args(c, bits, policy, value, n);
  if (rc) return rc;
  if (!scale) return fail(c, E55_ERR_INVALID, "null scale buffer");
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  if ((rc = ensure_capacity(c, n))) return rc;
  const size_t in_pp = static_cast<size_t>(c->in_planes), pol_pp = static_cast<size_t>(kPolicyCh) * kSq;
  constexpr int kChunk = kPipeChunk;
  const bool pipelined = c->precision == E55_PRECISION_BF16_TC && n >= 2 * kChunk;
  cudaError_t e = pipelined ? cudaEventRecord(c->pipe_start, c->stream) : cudaSuccess;
  int used = 0;
  for (int off = 0, i = 0, m = 0; off < n && e == cudaSuccess && rc == E55_OK; off += m, ++i) {
    m = pipelined ? pipe_chunk(n - off, kChunk) : n;   // one-group-per-CTA kernels of <= 256 positions, two at a time
    cudaStream_t st = pipelined ? c->pipe_stream[i % e55_ctx::kPipeStreams] : c->stream;
    if (pipelined && i < e55_ctx::kPipeStreams) { e = cudaStreamWaitEvent(st, c->pipe_start, 0); used = i + 1; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_bits + off * in_pp, bits + off * in_pp, m * in_pp * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_scale + off * in_pp, scale + off * in_pp, m * in_pp * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = forward_packed_on(c, st, c->d_bits + off * in_pp, c->d_scale + off * in_pp, m, off, c->d_policy + off * pol_pp, c->d_value + off);
    if (e == cudaSuccess && rc == E55_OK) e = cudaMemcpyAsync(policy + off * pol_pp, c->d_policy + off * pol_pp, m * pol_pp * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == E55_OK) e = cudaMemcpyAsync(value + off, c->d_value + off, m * sizeof(float), cudaMemcpyDeviceToHost, st);
  }
  for (int i = 0; i < used && e == cudaSuccess && rc == E55_OK; ++i) {
    e = cudaEventRecord(c->pipe_done[i], c->pipe_stream[i]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->pipe_done[i], 0);
  }
  if (e == cudaSuccess && rc == E55_OK) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess || rc != E55_OK) sync_pipe_streams(c);   // nothing of this call may still be in flight when the caller gets its buffers back
  if (rc != E55_OK) return rc;
  E55_CUDA(c, e);
  return E55_OK;
}

int e55_predict_packed_moves(e55_ctx* c, const uint32_t* bits, const float* scale, int n, const int32_t* move_offset,
                             const uint16_t* move_index, float* priors, float* value) {
  int rc = check_predict_args(c, bits, priors, value, n);
  if (rc) return rc;
  if (!scale || !move_offset || !move_index) return fail(c, E55_ERR_INVALID, "null buffer");
  const int total = move_offset[n];
  if (move_offset[0] != 0 || total < 0) return fail(c, E55_ERR_INVALID, "move_offset must start at 0 and be non-decreasing");
  for (int i = 0; i < n; ++i)
    if (move_offset[i + 1] < move_offset[i]) return fail(c, E55_ERR_INVALID, "move_offset must start at 0 and be non-decreasing");
  for (int j = 0; j < total; ++j)
    if (move_index[j] >= kPolicyCh * kSq) return fail(c, E55_ERR_INVALID, "move index outside the 1725 policy entries");
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  if ((rc = ensure_capacity(c, n))) return rc;
  // This is the call a single search makes for every leaf batch (latency matters): the whole request travels as ONE
  // copy from a pinned staging block, the whole answer as ONE copy back.
  auto up16 = [](size_t x) { return (x + 15) & ~size_t(15); };
  const size_t w = static_cast<size_t>(n) * c->in_planes;
  const size_t o_scale = up16(w * 4), o_off = o_scale + up16(w * 4), o_idx = o_off + up16((static_cast<size_t>(n) + 1) * 4);
  const size_t in_bytes = o_idx + up16(static_cast<size_t>(total) * 2);
  const size_t o_value = up16(static_cast<size_t>(total) * 4), out_bytes = o_value + up16(static_cast<size_t>(n) * 4);
  if (in_bytes > c->req_cap || out_bytes > c->ans_cap) {
    E55_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFreeHost(c->h_req); cudaFreeHost(c->h_ans); cudaFree(c->d_req); cudaFree(c->d_ans);
    c->h_req = c->h_ans = c->d_req = c->d_ans = nullptr;
    c->req_cap = c->ans_cap = 0;
    const size_t rq = std::max<size_t>(2 * in_bytes, 1 << 16), an = std::max<size_t>(2 * out_bytes, 1 << 14);
    E55_CUDA(c, cudaMallocHost(&c->h_req, rq));
    E55_CUDA(c, cudaMallocHost(&c->h_ans, an));
    E55_CUDA(c, cudaMalloc(&c->d_req, rq));
    E55_CUDA(c, cudaMalloc(&c->d_ans, an));
    c->req_cap = rq; c->ans_cap = an;
  }
  memcpy(c->h_req, bits, w * 4);
  memcpy(c->h_req + o_scale, scale, w * 4);
  memcpy(c->h_req + o_off, move_offset, (static_cast<size_t>(n) + 1) * 4);
  memcpy(c->h_req + o_idx, move_index, static_cast<size_t>(total) * 2);
  E55_CUDA(c, cudaMemcpyAsync(c->d_req, c->h_req, in_bytes, cudaMemcpyHostToDevice, c->stream));
  float* d_priors = reinterpret_cast<float*>(c->d_ans);
  float* d_value = reinterpret_cast<float*>(c->d_ans + o_value);
  const int32_t* d_off = reinterpret_cast<const int32_t*>(c->d_req + o_off);
  const uint16_t* d_idx = reinterpret_cast<const uint16_t*>(c->d_req + o_idx);
  if (c->precision == E55_PRECISION_BF16_TC) {
    // the tower kernel soft-maxes the legal moves itself (policy tail): the 1725 logits never leave shared memory
    e55::tc::PolicyTail tail;
    tail.move_off = d_off; tail.move_idx = d_idx; tail.priors = d_priors;
    if ((rc = forward_packed_on(c, c->stream, reinterpret_cast<const uint32_t*>(c->d_req), reinterpret_cast<const float*>(c->d_req + o_scale), n, 0,
                                nullptr, d_value, 0, tail))) return rc;
  } else {
    if ((rc = forward_packed_on(c, c->stream, reinterpret_cast<const uint32_t*>(c->d_req), reinterpret_cast<const float*>(c->d_req + o_scale), n, 0,
                                c->d_policy, d_value))) return rc;
    e55::legal_priors_kernel<<<(n + 7) / 8, 256, 0, c->stream>>>(c->d_policy, d_off, d_idx, d_priors, n, kPolicyCh * kSq);
    ++c->launches;
    E55_CUDA(c, cudaGetLastError());
  }
  E55_CUDA(c, cudaMemcpyAsync(c->h_ans, c->d_ans, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(priors, c->h_ans, static_cast<size_t>(total) * 4);
  memcpy(value, c->h_ans + o_value, static_cast<size_t>(n) * 4);
  return E55_OK;
}

int e55_predict_masked_device(e55_ctx* c, const float* d_planes, const uint8_t* d_mask, int n, float* d_probs,
                              float* d_value) {
  int rc = check_predict_args(c, d_planes, d_probs, d_value, n);
  if (rc) return rc;
  if (!d_mask) return fail(c, E55_ERR_INVALID, "null legal mask");
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  if ((rc = ensure_capacity(c, n))) return rc;
  if (c->precision == E55_PRECISION_BF16_TC) {   // masked softmax in the tower kernel's policy tail
    e55::tc::PolicyTail tail;
    tail.mask = d_mask; tail.probs = d_probs;
    return forward_tc_on(c, c->stream, d_planes, n, nullptr, d_value, tail);
  }
  if ((rc = forward(c, d_planes, n, c->d_policy, d_value))) return rc;
  e55::fp32::masked_softmax_kernel<<<n, 256, 0, c->stream>>>(c->d_policy, d_mask, d_probs, kPolicyCh * kSq);
  ++c->launches;
  E55_CUDA(c, cudaGetLastError());
  return E55_OK;
}

int e55_predict_masked(e55_ctx* c, const float* planes, const uint8_t* mask, int n, float* probs, float* value) {
  int rc = check_predict_args(c, planes, probs, value, n);
  if (rc) return rc;
  if (!mask) return fail(c, E55_ERR_INVALID, "null legal mask");
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  if ((rc = ensure_capacity(c, n))) return rc;
  const size_t pol = static_cast<size_t>(n) * kPolicyCh * kSq;
  E55_CUDA(c, cudaMemcpyAsync(c->d_in, planes, static_cast<size_t>(n) * c->in_planes * kSq * sizeof(float),
                              cudaMemcpyHostToDevice, c->stream));
  E55_CUDA(c, cudaMemcpyAsync(c->d_mask, mask, pol, cudaMemcpyHostToDevice, c->stream));
  if (c->precision == E55_PRECISION_BF16_TC) {   // masked softmax in the tower kernel's policy tail
    e55::tc::PolicyTail tail;
    tail.mask = c->d_mask; tail.probs = c->d_probs;
    if ((rc = forward_tc_on(c, c->stream, c->d_in, n, nullptr, c->d_value, tail))) return rc;
  } else {
    if ((rc = forward(c, c->d_in, n, c->d_policy, c->d_value))) return rc;
    e55::fp32::masked_softmax_kernel<<<n, 256, 0, c->stream>>>(c->d_policy, c->d_mask, c->d_probs, kPolicyCh * kSq);
    ++c->launches;
    E55_CUDA(c, cudaGetLastError());
  }
  E55_CUDA(c, cudaMemcpyAsync(probs, c->d_probs, pol * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  E55_CUDA(c, cudaMemcpyAsync(value, c->d_value, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost,
                              c->stream));
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  return E55_OK;
}

int e55_host_alloc(size_t bytes, void** out) {
  if (!out || bytes == 0) return fail(nullptr, E55_ERR_INVALID, "bad host allocation request");
  *out = nullptr;
  E55_CUDA(nullptr, cudaHostAlloc(out, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
  return E55_OK;
}

int e55_host_free(void* p) {
  if (p) E55_CUDA(nullptr, cudaFreeHost(p));
  return E55_OK;
}

int e55_sync(e55_ctx* c) {
  if (!c) return E55_ERR_INVALID;
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  return E55_OK;
}

int e55_stream(e55_ctx* c, void** stream_out) {
  if (!c || !stream_out) return E55_ERR_INVALID;
  *stream_out = static_cast<void*>(c->stream);
  return E55_OK;
}

int e55_set_host_threads(e55_ctx* c, int threads) {
  if (!c || threads < -1 || threads > 256) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  c->host_threads = threads;
  c->host_auto = threads < 0;
  c->pipe_calls = 0; c->split_level = 0;
  for (int i = 0; i < e55_ctx::kSplitLevels; ++i) { c->split_us[i] = 0.0; c->split_samples[i] = 0; }
  if (threads == 0) { e55host::packer_destroy(c->packer); c->packer = nullptr; }
  return E55_OK;
}

int e55_host_read_bandwidth(e55_ctx* c, const void* buf, size_t bytes, int passes, double* gbs_out) {
  if (!c || !buf || bytes < (1u << 20) || passes < 1 || !gbs_out) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  e55host::Packer* packer = host_packer(c, 1);
  if (!packer) return fail(c, E55_ERR_UNSUPPORTED, "the host thread pool is switched off (e55_set_host_threads)");
  double best = 0.0;
  for (int i = 0; i < passes; ++i) best = std::max(best, e55host::packer_read_bandwidth(packer, buf, bytes));
  *gbs_out = best;
  return E55_OK;
}

int e55_get_host_pipeline(e55_ctx* c, int* threads_out, double* compact_us_out, double* float_us_out) {
  if (!c || !threads_out || !compact_us_out || !float_us_out) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  *threads_out = c->packer ? e55host::packer_threads(c->packer) : 0;
  int best = 0;                                       // the fastest split that packs at all
  for (int i = 1; i < e55_ctx::kSplitLevels - 1; ++i)
    if (c->split_samples[i] && (!c->split_samples[best] || c->split_us[i] < c->split_us[best])) best = i;
  *compact_us_out = c->split_us[best];
  *float_us_out = c->split_us[e55_ctx::kSplitLevels - 1];
  return E55_OK;
}

int e55_get_host_split(e55_ctx* c, int* float_eighths_out, double* us_per_position_out) {
  if (!c || !float_eighths_out) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  *float_eighths_out = e55_ctx::kSplitEighths[c->split_level];
  if (us_per_position_out)
    for (int i = 0; i < e55_ctx::kSplitLevels; ++i) us_per_position_out[i] = c->split_us[i];
  return E55_OK;
}

int e55_get_arch(const e55_ctx* c, int* blocks_out, int* filters_out, int* in_planes_out) {
  if (!c || !blocks_out || !filters_out || !in_planes_out) return E55_ERR_INVALID;
  *blocks_out = c->blocks; *filters_out = c->filters; *in_planes_out = c->in_planes;
  return E55_OK;
}

int e55_launch_count(const e55_ctx* c, int64_t* count_out) {
  if (!c || !count_out) return E55_ERR_INVALID;
  *count_out = c->launches;
  return E55_OK;
}

int e55_debug_activations(e55_ctx* c, const float* planes, int n, int n_convs, float* act_out) {
  int rc = check_predict_args(c, planes, act_out, act_out, n);
  if (rc) return rc;
  if (n_convs < 1 || n_convs > 1 + 2 * c->blocks) return fail(c, E55_ERR_INVALID, "n_convs out of range");
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  if ((rc = ensure_capacity(c, n))) return rc;
  E55_CUDA(c, cudaMemcpyAsync(c->d_in, planes, static_cast<size_t>(n) * c->in_planes * kSq * sizeof(float),
                              cudaMemcpyHostToDevice, c->stream));
  const float* result = nullptr;
  if (c->precision == E55_PRECISION_FP32) {
    float* x = c->d_act[0];
    float* y = c->d_act[1];
    float* z = c->d_act[2];
    if ((rc = launch_conv3x3(c, c->convs[0], c->d_in, nullptr, x, n, 1))) return rc;
    result = x;
    for (int i = 1; i < n_convs; ++i) {
      if (i % 2 == 1) {
        if ((rc = launch_conv3x3(c, c->convs[i], x, nullptr, y, n, 1))) return rc;
        result = y;
      } else {
        if ((rc = launch_conv3x3(c, c->convs[i], y, x, z, n, 1))) return rc;
        float* t = x; x = z; z = t;
        result = x;
      }
    }
  } else {
    std::string why;
    int launched = 0;
    cudaError_t e = e55::tc::debug_layers(c->tc, c->d_in, n, n_convs, c->d_act[0], c->stream, &launched, &why);
    c->launches += launched;
    if (e != cudaSuccess) return fail(c, E55_ERR_CUDA, "debug activations: " + why + ": " + cudaGetErrorString(e));
    if (!why.empty()) return fail(c, E55_ERR_UNSUPPORTED, why);
    result = c->d_act[0];
  }
  E55_CUDA(c, cudaMemcpyAsync(act_out, result, static_cast<size_t>(n) * c->filters * kSq * sizeof(float),
                              cudaMemcpyDeviceToHost, c->stream));
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  return E55_OK;
}

int e55_debug_legal_priors(e55_ctx* c, const float* logits, int n, const int32_t* move_offset, const uint16_t* move_index, float* priors) {
  if (!c || !logits || !move_offset || !move_index || !priors || n <= 0) return fail(c, E55_ERR_INVALID, "bad arguments");
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure_capacity(c, n);
  if (rc) return rc;
  const int total = move_offset[n];
  if (total < 0 || static_cast<size_t>(total) > static_cast<size_t>(c->capacity) * e55_ctx::kMovesPerPosition)
    return fail(c, E55_ERR_INVALID, "too many moves");
  E55_CUDA(c, cudaMemcpyAsync(c->d_policy, logits, static_cast<size_t>(n) * kPolicyCh * kSq * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  E55_CUDA(c, cudaMemcpyAsync(c->d_move_off, move_offset, (static_cast<size_t>(n) + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  E55_CUDA(c, cudaMemcpyAsync(c->d_move_idx, move_index, static_cast<size_t>(total) * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
  e55::legal_priors_kernel<<<(n + 7) / 8, 256, 0, c->stream>>>(c->d_policy, c->d_move_off, c->d_move_idx, c->d_priors, n, kPolicyCh * kSq);
  ++c->launches;
  E55_CUDA(c, cudaGetLastError());
  E55_CUDA(c, cudaMemcpyAsync(priors, c->d_priors, static_cast<size_t>(total) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  return E55_OK;
}

int e55_set_kernel_timing(e55_ctx* c, int enable) {
  if (!c) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  if (enable && !c->ev0) {
    E55_CUDA(c, cudaEventCreate(&c->ev0));
    E55_CUDA(c, cudaEventCreate(&c->ev1));
  }
  c->timing = enable != 0;
  c->timed_ms = 0.0; c->timed_launches = 0; c->ev_pending = false;
  return E55_OK;
}

int e55_get_kernel_timing(e55_ctx* c, double* total_ms_out, int64_t* launches_out) {
  if (!c || !total_ms_out || !launches_out) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  int rc = collect_timing(c);
  if (rc) return rc;
  *total_ms_out = c->timed_ms;
  *launches_out = c->timed_launches;
  return E55_OK;
}

int e55_debug_set_lag(e55_ctx* c, int lag) {
  if (!c || lag < 0 || lag > 16) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  c->tc.lag = lag;
  return E55_OK;
}

static constexpr size_t kProfWords = 16 + 8192 + 1024;

int e55_debug_profile(e55_ctx* c, int enable, uint64_t* counters_out) {
  if (!c) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  if (counters_out && c->tc.d_prof)
    E55_CUDA(c, cudaMemcpy(counters_out, c->tc.d_prof, kProfWords * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (enable && !c->tc.d_prof) E55_CUDA(c, cudaMalloc(&c->tc.d_prof, kProfWords * sizeof(uint64_t)));
  if (!enable && c->tc.d_prof) { cudaFree(c->tc.d_prof); c->tc.d_prof = nullptr; }
  if (c->tc.d_prof) E55_CUDA(c, cudaMemsetAsync(c->tc.d_prof, 0, kProfWords * sizeof(uint64_t), c->stream));   // ordered with the kernels
  return E55_OK;
}

int e55_selftest_slice_book(int64_t slices, uint64_t seed, int64_t* mismatches_out) {
  if (!mismatches_out || slices < 0) return fail(nullptr, E55_ERR_INVALID, "bad selftest arguments");
  // the kernel's bookkeeping against plain 64-bit counters, buffers visited in a pseudo-random order
  e55::tc::SliceBook book;
  uint64_t count[4] = {0, 0, 0, 0};
  int64_t bad = 0;
  uint64_t x = seed * 2862933555777941757ull + 3037000493ull;
  for (int64_t i = 0; i < slices; ++i) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    const int b = static_cast<int>(x >> 62);
    for (int k = 0; k < 4; ++k) {
      if (book.none_yet(k) != (count[k] == 0)) ++bad;
      if (count[k] > 0 && book.release_parity(k) != static_cast<uint32_t>((count[k] - 1) & 1)) ++bad;
    }
    book.wrote(b);
    ++count[b];
  }
  *mismatches_out = bad;
  return E55_OK;
}

int e55_selftest_umma(int device, int row_shift, float* max_abs_err_out) {
  using namespace e55::ummatest;
  if (!max_abs_err_out || row_shift < 0 || row_shift > 64) return fail(nullptr, E55_ERR_INVALID, "bad selftest arguments");
  E55_CUDA(nullptr, cudaSetDevice(device));
  std::vector<__nv_bfloat16> ha(kRowsA * kK), hb(kN * kK);
  std::vector<float> fa(ha.size()), fb(hb.size());
  uint32_t s = 12345u + row_shift;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return static_cast<float>((s >> 9) & 0xFF) / 128.f - 1.f; };
  for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2bfloat16(rnd()); fa[i] = __bfloat162float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { hb[i] = __float2bfloat16(rnd()); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da = nullptr, *db = nullptr;
  float* dd = nullptr;
  E55_CUDA(nullptr, cudaMalloc(&da, ha.size() * 2));
  E55_CUDA(nullptr, cudaMalloc(&db, hb.size() * 2));
  E55_CUDA(nullptr, cudaMalloc(&dd, kM * kN * 4));
  E55_CUDA(nullptr, cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  E55_CUDA(nullptr, cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  E55_CUDA(nullptr, cudaMemset(dd, 0xFF, kM * kN * 4));
  const size_t smem = (kK / 8) * (kRowsA + kN) * 16;
  E55_CUDA(nullptr, cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  selftest_kernel<<<1, 128, smem>>>(da, db, dd, row_shift);
  E55_CUDA(nullptr, cudaGetLastError());
  E55_CUDA(nullptr, cudaDeviceSynchronize());
  std::vector<float> hd(kM * kN);
  E55_CUDA(nullptr, cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dd);
  float worst = 0.f;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double ref = 0.0;
      for (int k = 0; k < kK; ++k) ref += static_cast<double>(fa[(m + row_shift) * kK + k]) * fb[n * kK + k];
      const float err = std::fabs(hd[m * kN + n] - static_cast<float>(ref));
      if (!(err <= worst)) worst = std::isnan(err) ? INFINITY : err;
    }
  *max_abs_err_out = worst;
  return E55_OK;
}

int e55_bench_umma(int device, int n_dim, int iters, float* tflops_out) {
  using namespace e55::ummatest;
  if (!tflops_out || iters < 1 || (n_dim != 128 && n_dim != 256)) return fail(nullptr, E55_ERR_INVALID, "bad bench arguments");
  E55_CUDA(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  E55_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  const int grid = prop.multiProcessorCount;
  float* sink = nullptr;
  E55_CUDA(nullptr, cudaMalloc(&sink, grid * sizeof(float)));
  const size_t smem = 16 * (128 * 2 + 64) * 16 + 8 * static_cast<size_t>(n_dim) * 16;
  cudaEvent_t e0, e1;
  E55_CUDA(nullptr, cudaEventCreate(&e0));
  E55_CUDA(nullptr, cudaEventCreate(&e1));
  float ms = 0.f;
  for (int rep = 0; rep < 2; ++rep) {
    E55_CUDA(nullptr, cudaEventRecord(e0));
    if (n_dim == 128) {
      E55_CUDA(nullptr, cudaFuncSetAttribute(bench_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      bench_kernel<128><<<grid, 128, smem>>>(iters, sink);
    } else {
      E55_CUDA(nullptr, cudaFuncSetAttribute(bench_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      bench_kernel<256><<<grid, 128, smem>>>(iters, sink);
    }
    E55_CUDA(nullptr, cudaGetLastError());
    E55_CUDA(nullptr, cudaEventRecord(e1));
    E55_CUDA(nullptr, cudaEventSynchronize(e1));
    E55_CUDA(nullptr, cudaEventElapsedTime(&ms, e0, e1));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
  const double flops = 2.0 * 128.0 * n_dim * 16.0 * 144.0 * iters * grid;
  *tflops_out = static_cast<float>(flops / (ms * 1e-3) / 1e12);
  return E55_OK;
}

}  // extern "C"

#include "e55_service.inl"
#include "e55_gpu_selfplay.inl"

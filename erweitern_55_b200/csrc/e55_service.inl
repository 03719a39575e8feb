// Cross-game leaf aggregation: many search threads hand in the small leaf batches
// mcts.py produces (root: 1, leaves: Config.batch_size = 16; mcts.py:57-58,91-106) and the service coalesces them
// into one large forward per GPU.  Included by e55_api.cu (needs e55_ctx internals).
//
// Host side: a ring of pinned staging batches.  A request reserves a slot range of the open batch under a short
// lock, copies its own input in and later its own output out (so host copies run on the callers' threads), and
// holds a ticket in between; blocking calls wait on the batch's generation counter (futex-style), asynchronous
// callers (e55_service_submit_packed_moves / e55_service_fetch) keep several tickets in flight from one thread.
// Device side (tensor-core path): three device slots with their own stream and buffers, so the H2D copy of batch k+1
// overlaps the tower kernel of batch k and the next kernel starts on the SMs the previous one leaves; a dispatcher thread closes and launches batches, a completer thread waits
// for their events and wakes the readers.  The fp32 path runs one batch at a time on the context stream.

namespace {

constexpr int kStaging = 5;       // host staging batches in the ring
constexpr int kDeviceSlots = 3;   // batches in flight on the GPU

enum BatchState { kFree = 0, kOpen, kClosed, kDone };

struct ServiceBatch {
  // pinned host staging
  float* planes = nullptr;      // [cap][in_planes][25]   (float-plane requests)
  uint32_t* bits = nullptr;     // [cap][in_planes]       (packed requests)
  float* scale = nullptr;
  float* policy = nullptr;      // [cap][1725]
  float* value = nullptr;       // [cap]
  int32_t* move_off = nullptr;  // [cap + 1]              (packed + legal-move requests: CSR move lists)
  uint16_t* move_idx = nullptr; // [cap * kMovesPerPosition]
  float* priors = nullptr;      // [cap * kMovesPerPosition]
  int kind = 0;                 // 0 = empty, 1 = float planes, 2 = packed, 3 = packed + legal moves -> priors
  int reserved = 0;             // positions handed out
  int moves_reserved = 0;       // move slots handed out (kind 3)
  int state = kFree;            // guarded by e55_service::m
  std::atomic<int> filled{0};   // positions whose input copy has finished
  std::atomic<int> readers{0};              // tickets that still have to fetch their output
  std::atomic<uint64_t> generation{0};      // bumped when the batch's results are ready
  int status = 0;               // result code of the forward
  cudaEvent_t done = nullptr;
  cudaEvent_t k0 = nullptr, k1 = nullptr;   // E55_SERVICE_TRACE: around the kernels of the batch
  std::chrono::steady_clock::time_point first_arrival;
};

struct DeviceSlot {
  cudaStream_t stream = nullptr;
  float* planes = nullptr;
  uint32_t* bits = nullptr;
  float* scale = nullptr;
  float* policy = nullptr;
  float* value = nullptr;
  int32_t* move_off = nullptr;
  uint16_t* move_idx = nullptr;
  float* priors = nullptr;
};

}  // namespace

struct e55_service {
  std::thread dispatcher, completer;
  std::mutex m;
  std::condition_variable cv_submit;    // dispatcher waits for requests
  std::condition_variable cv_space;     // submitters wait for room in an open batch
  std::condition_variable cv_free;      // dispatcher waits for a staging batch to be released
  std::condition_variable cv_slot;      // dispatcher waits for a device slot
  std::condition_variable cv_inflight;  // completer waits for launched batches
  ServiceBatch batch[kStaging];
  DeviceSlot slot[kDeviceSlots];
  int open = -1;                        // index of the batch currently accepting requests (-1: none free)
  int inflight[kDeviceSlots] = {};         // FIFO of launched batches
  int n_inflight = 0, next_slot = 0;
  int cap = 0, max_wait_us = 0;
  int idle_min = 0;                     // an idle GPU takes a batch of at least this many positions at once
  bool stop = false, dispatcher_done = false;
  bool force_close = false;             // a request does not fit the open batch: dispatch it now
  int64_t n_batches = 0, n_positions = 0;
  // E55_SERVICE_TRACE=1: where a batch spends its time (microseconds, summed over batches; printed at stop)
  bool trace = false;
  double t_fill = 0, t_slot = 0, t_copywait = 0, t_launch = 0, t_gpu = 0, t_idle = 0, t_kernels = 0, t_interval = 0;
  int64_t n_kernels = 0;
  cudaEvent_t prev_k1 = nullptr;
  std::chrono::steady_clock::time_point launched_at[kStaging], last_complete;
  int64_t closed_full = 0, closed_idle = 0, closed_timeout = 0, closed_forced = 0;
};

namespace {

// Enqueue one closed batch.  Tensor-core path: asynchronous on the slot's stream (returns E55_PENDING: completion is
// signalled by b.done); fp32 path: synchronous on the context stream with the context's own buffers.
int service_launch(e55_ctx* c, e55_service* s, ServiceBatch& b, int slot_index) {
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  const int n = b.reserved;
  const size_t pol = static_cast<size_t>(n) * kPolicyCh * kSq;
  const size_t w = static_cast<size_t>(n) * c->in_planes;
  if (b.kind == 3) b.move_off[n] = b.moves_reserved;
  const bool async = c->precision == E55_PRECISION_BF16_TC;
  int rc = E55_OK;
  DeviceSlot d = s->slot[slot_index];
  if (!async) {   // the fp32 kernels work in the context's activation buffers: one batch at a time
    if ((rc = ensure_capacity(c, n))) return rc;
    d.stream = c->stream; d.planes = c->d_in; d.bits = c->d_bits; d.scale = c->d_scale; d.policy = c->d_policy;
    d.value = c->d_value; d.move_off = c->d_move_off; d.move_idx = c->d_move_idx; d.priors = c->d_priors;
  }
  if (b.kind == 1) {
    E55_CUDA(c, cudaMemcpyAsync(d.planes, b.planes, w * kSq * sizeof(float), cudaMemcpyHostToDevice, d.stream));
    rc = async ? forward_tc_on(c, d.stream, d.planes, n, d.policy, d.value) : forward(c, d.planes, n, d.policy, d.value);
  } else {
    E55_CUDA(c, cudaMemcpyAsync(d.bits, b.bits, w * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream));
    E55_CUDA(c, cudaMemcpyAsync(d.scale, b.scale, w * sizeof(float), cudaMemcpyHostToDevice, d.stream));
    if (b.kind == 3) {
      E55_CUDA(c, cudaMemcpyAsync(d.move_off, b.move_off, (static_cast<size_t>(n) + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, d.stream));
      if (b.moves_reserved)
        E55_CUDA(c, cudaMemcpyAsync(d.move_idx, b.move_idx, static_cast<size_t>(b.moves_reserved) * sizeof(uint16_t), cudaMemcpyHostToDevice, d.stream));
    }
    if (s->trace) cudaEventRecord(b.k0, d.stream);
    if (b.kind == 3 && async) {   // the tower kernel's policy tail soft-maxes the legal moves: no logits in HBM, one launch
      e55::tc::PolicyTail tail;
      tail.move_off = d.move_off; tail.move_idx = d.move_idx; tail.priors = d.priors;
      rc = forward_packed_on(c, d.stream, d.bits, d.scale, n, 0, nullptr, d.value, 0, tail);
    } else {
      rc = forward_packed_on(c, d.stream, d.bits, d.scale, n, 0, d.policy, d.value);
    }
  }
  if (rc) return rc;
  if (b.kind == 3) {   // only the legal moves' priors travel back
    if (!async) {
      e55::legal_priors_kernel<<<(n + 7) / 8, 256, 0, d.stream>>>(d.policy, d.move_off, d.move_idx, d.priors, n, kPolicyCh * kSq);
      ++c->launches;
      E55_CUDA(c, cudaGetLastError());
    }
    if (s->trace) cudaEventRecord(b.k1, d.stream);
    if (b.moves_reserved)
      E55_CUDA(c, cudaMemcpyAsync(b.priors, d.priors, static_cast<size_t>(b.moves_reserved) * sizeof(float), cudaMemcpyDeviceToHost, d.stream));
  } else {
    E55_CUDA(c, cudaMemcpyAsync(b.policy, d.policy, pol * sizeof(float), cudaMemcpyDeviceToHost, d.stream));
  }
  E55_CUDA(c, cudaMemcpyAsync(b.value, d.value, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost, d.stream));
  if (!async) {
    E55_CUDA(c, cudaStreamSynchronize(d.stream));
    return E55_OK;
  }
  E55_CUDA(c, cudaEventRecord(b.done, d.stream));
  return E55_PENDING;
}

// Results of `b` are in its staging buffers (or the forward failed): wake the readers.  Called with s->m held.
void service_finish(e55_service* s, ServiceBatch& b, int rc) {
  b.status = rc;
  b.state = kDone;
  ++s->n_batches;
  s->n_positions += b.reserved;
  b.generation.fetch_add(1, std::memory_order_release);
  b.generation.notify_all();
}

// With s->m held: make a released staging batch the open one, if there is none.
void service_open_next(e55_service* s) {
  if (s->open >= 0) return;
  for (int i = 0; i < kStaging; ++i) {
    ServiceBatch& b = s->batch[i];
    if (b.state != kFree) continue;
    b.kind = 0; b.reserved = 0; b.moves_reserved = 0; b.filled.store(0, std::memory_order_relaxed);
    b.state = kOpen;
    s->open = i;
    s->cv_space.notify_all();
    return;
  }
}

void service_dispatch_loop(e55_ctx* c) {
  e55_service* s = c->service;
  std::unique_lock<std::mutex> lk(s->m);
  for (;;) {
    while (!s->stop && s->open < 0) {
      service_open_next(s);
      if (s->open < 0) s->cv_free.wait(lk);
    }
    if (s->open < 0) break;   // stopping
    const int idx = s->open;
    ServiceBatch& b = s->batch[idx];
    s->cv_submit.wait(lk, [&] { return s->stop || b.reserved > 0; });
    if (b.reserved == 0) break;   // stopping with nothing pending
    // Give other games a moment to join: the batch goes out when it is full, when a request did not fit, when
    // the GPU has nothing to do and the batch is worth a launch, or max_wait_us after its first request.
    const auto deadline = b.first_arrival + std::chrono::microseconds(s->max_wait_us);
    while (!s->stop && !s->force_close && b.reserved < s->cap && !(s->n_inflight == 0 && b.reserved >= s->idle_min) &&
           std::chrono::steady_clock::now() < deadline)
      s->cv_submit.wait_until(lk, deadline);
    const auto t_closed = std::chrono::steady_clock::now();
    if (s->trace) {
      if (b.reserved >= s->cap) ++s->closed_full; else if (s->force_close) ++s->closed_forced;
      else if (t_closed >= deadline) ++s->closed_timeout; else ++s->closed_idle;
      s->t_fill += std::chrono::duration<double, std::micro>(t_closed - b.first_arrival).count();
    }
    s->force_close = false;
    b.state = kClosed;
    s->open = -1;
    service_open_next(s);   // submitters continue in the next staging batch while this one is launched
    s->cv_slot.wait(lk, [&] { return s->n_inflight < (c->precision == E55_PRECISION_BF16_TC ? kDeviceSlots : 1); });
    const int slot_index = s->next_slot;
    s->next_slot = (s->next_slot + 1) % kDeviceSlots;
    if (s->trace && s->n_inflight == 0) s->t_idle += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - s->last_complete).count();
    lk.unlock();
    const auto t_slot = std::chrono::steady_clock::now();
    while (b.filled.load(std::memory_order_acquire) < b.reserved) std::this_thread::yield();
    const auto t_copied = std::chrono::steady_clock::now();
    const int rc = service_launch(c, s, b, slot_index);
    lk.lock();
    if (s->trace) {
      const auto t_launched = std::chrono::steady_clock::now();
      s->t_slot += std::chrono::duration<double, std::micro>(t_slot - t_closed).count();
      s->t_copywait += std::chrono::duration<double, std::micro>(t_copied - t_slot).count();
      s->t_launch += std::chrono::duration<double, std::micro>(t_launched - t_copied).count();
      s->launched_at[idx] = t_launched;
    }
    if (rc == E55_PENDING) {
      s->inflight[s->n_inflight++] = idx;
      s->cv_inflight.notify_one();
    } else {
      service_finish(s, b, rc);
    }
  }
  s->dispatcher_done = true;
  s->cv_inflight.notify_all();
}

void service_complete_loop(e55_ctx* c) {
  e55_service* s = c->service;
  cudaSetDevice(c->device);
  std::unique_lock<std::mutex> lk(s->m);
  for (;;) {
    s->cv_inflight.wait(lk, [&] { return s->n_inflight > 0 || s->dispatcher_done; });
    if (s->n_inflight == 0) return;
    ServiceBatch& b = s->batch[s->inflight[0]];
    lk.unlock();
    const cudaError_t e = cudaEventSynchronize(b.done);
    int rc = E55_OK;
    if (e != cudaSuccess) {
      std::lock_guard<std::mutex> g(c->mu);
      rc = fail(c, E55_ERR_CUDA, std::string("evaluation service forward: ") + cudaGetErrorString(e));
    }
    lk.lock();
    if (s->trace) {
      float ms = 0.f;
      if (b.kind == 3 && cudaEventElapsedTime(&ms, b.k0, b.k1) == cudaSuccess) { s->t_kernels += ms * 1e3; ++s->n_kernels; }
      if (s->prev_k1 && b.kind == 3 && cudaEventElapsedTime(&ms, s->prev_k1, b.k1) == cudaSuccess) s->t_interval += ms * 1e3;
      s->prev_k1 = b.kind == 3 ? b.k1 : nullptr;
      s->last_complete = std::chrono::steady_clock::now();
      s->t_gpu += std::chrono::duration<double, std::micro>(s->last_complete - s->launched_at[s->inflight[0]]).count();
    }
    for (int i = 1; i < s->n_inflight; ++i) s->inflight[i - 1] = s->inflight[i];
    --s->n_inflight;
    service_finish(s, b, rc);
    s->cv_slot.notify_one();
    if (s->n_inflight == 0) s->cv_submit.notify_one();   // the GPU is idle: the dispatcher may send what it has
  }
}

struct Ticket {   // mirrors e55_ticket
  int32_t batch, off, moff, n, total, kind;
  uint64_t gen;
};

// Reserve room for a request and copy its input into the staging batch.  kind 1 (planes), 2 (bits + scale) or
// 3 (bits + scale + legal-move lists).  Returns E55_OK, E55_PENDING when may_block is false and no open batch has
// room right now, or an error.
int service_submit(e55_ctx* c, int kind, const float* planes, const uint32_t* bits, const float* scale, int n,
                   const int32_t* move_offset, const uint16_t* move_index, bool may_block, Ticket* t) {
  e55_service* s = c->service;
  auto fail_locked = [&](int code, const char* msg) {   // many threads call in: serialise the error string
    std::lock_guard<std::mutex> g(c->mu);
    return fail(c, code, msg);
  };
  if (!s) return fail_locked(E55_ERR_INVALID, "evaluation service is not running (e55_service_start)");
  if (n <= 0 || !t || (kind == 1 ? !planes : (!bits || !scale)) || (kind == 3 && (!move_offset || !move_index)))
    return fail_locked(E55_ERR_INVALID, "bad service request");
  if (n > s->cap) return fail_locked(E55_ERR_INVALID, "request larger than the service batch");
  const int total = kind == 3 ? move_offset[n] : 0;
  if (kind == 3 && (move_offset[0] != 0 || total < 0 || total > s->cap * e55_ctx::kMovesPerPosition))
    return fail_locked(E55_ERR_INVALID, "bad move lists in service request");
  if (kind == 3) {
    // every list must lie inside this request's own range: the batch is shared with other callers' requests
    for (int i = 0; i < n; ++i)
      if (move_offset[i] > move_offset[i + 1]) return fail_locked(E55_ERR_INVALID, "move_offset must be non-decreasing");
    for (int j = 0; j < total; ++j)
      if (move_index[j] >= kPolicyCh * kSq) return fail_locked(E55_ERR_INVALID, "move index outside the 1725 policy entries");
  }
  if (!c->has_weights) return fail_locked(E55_ERR_NO_WEIGHTS, "predict called before e55_set_weights");
  std::unique_lock<std::mutex> lk(s->m);
  ServiceBatch* b = nullptr;
  for (;;) {
    if (s->stop) { lk.unlock(); return fail_locked(E55_ERR_INVALID, "evaluation service stopped"); }
    if (s->open >= 0) {
      b = &s->batch[s->open];
      if ((b->kind == 0 || b->kind == kind) && b->reserved + n <= s->cap &&
          b->moves_reserved + total <= s->cap * e55_ctx::kMovesPerPosition) break;
      if (b->reserved > 0) { s->force_close = true; s->cv_submit.notify_one(); }   // cannot take this request: dispatch it now
    }
    if (!may_block) return E55_PENDING;
    s->cv_space.wait(lk);
  }
  const int off = b->reserved;
  if (off == 0) { b->kind = kind; b->first_arrival = std::chrono::steady_clock::now(); }
  b->reserved += n;
  const int moff = b->moves_reserved;
  b->moves_reserved += total;
  b->readers.fetch_add(1, std::memory_order_relaxed);
  *t = Ticket{s->open, off, moff, n, total, kind, b->generation.load(std::memory_order_relaxed)};
  if (off == 0 || b->reserved == s->cap || (s->n_inflight == 0 && b->reserved >= s->idle_min)) s->cv_submit.notify_one();
  lk.unlock();

  const size_t ip = static_cast<size_t>(c->in_planes);
  if (kind == 1) {
    memcpy(b->planes + off * ip * kSq, planes, static_cast<size_t>(n) * ip * kSq * sizeof(float));
  } else {
    memcpy(b->bits + off * ip, bits, static_cast<size_t>(n) * ip * sizeof(uint32_t));
    memcpy(b->scale + off * ip, scale, static_cast<size_t>(n) * ip * sizeof(float));
    if (kind == 3) {
      for (int i = 0; i < n; ++i) b->move_off[off + i] = moff + move_offset[i];
      for (int j = 0; j < total; ++j)     // an index outside the policy row would read foreign memory: clamp it
        b->move_idx[moff + j] = move_index[j] < kPolicyCh * kSq ? move_index[j] : static_cast<uint16_t>(kPolicyCh * kSq - 1);
    }
  }
  b->filled.fetch_add(n, std::memory_order_release);
  return E55_OK;
}

// Copy the results of a ticket out and release it.  `out` receives the policy rows (kinds 1, 2) or the priors of
// the listed moves (kind 3).  Returns E55_PENDING when the batch has not finished and wait is false.
int service_fetch(e55_ctx* c, const Ticket& t, float* out, float* value, bool wait) {
  e55_service* s = c->service;
  if (!s || t.batch < 0 || t.batch >= kStaging || !out || !value) return E55_ERR_INVALID;
  ServiceBatch* b = &s->batch[t.batch];
  while (b->generation.load(std::memory_order_acquire) == t.gen) {
    if (!wait) return E55_PENDING;
    b->generation.wait(t.gen);
  }
  const int rc = b->status;
  if (rc == E55_OK) {
    if (t.kind == 3) memcpy(out, b->priors + t.moff, static_cast<size_t>(t.total) * sizeof(float));
    else memcpy(out, b->policy + static_cast<size_t>(t.off) * kPolicyCh * kSq, static_cast<size_t>(t.n) * kPolicyCh * kSq * sizeof(float));
    memcpy(value, b->value + t.off, static_cast<size_t>(t.n) * sizeof(float));
  }
  if (b->readers.fetch_sub(1, std::memory_order_acq_rel) == 1) {   // last reader: hand the staging batch back
    std::lock_guard<std::mutex> lk(s->m);
    b->state = kFree;
    s->cv_free.notify_one();
  }
  return rc;
}

int service_predict(e55_ctx* c, int kind, const float* planes, const uint32_t* bits, const float* scale, int n,
                    float* out, float* value, const int32_t* move_offset = nullptr, const uint16_t* move_index = nullptr) {
  if (!out || !value) {
    std::lock_guard<std::mutex> g(c->mu);
    return fail(c, E55_ERR_INVALID, "bad service request");
  }
  Ticket t;
  const int rc = service_submit(c, kind, planes, bits, scale, n, move_offset, move_index, true, &t);
  if (rc != E55_OK) return rc;
  return service_fetch(c, t, out, value, true);
}

void service_free(e55_service* s) {
  for (auto& b : s->batch) {
    cudaFreeHost(b.planes); cudaFreeHost(b.bits); cudaFreeHost(b.scale); cudaFreeHost(b.policy); cudaFreeHost(b.value);
    cudaFreeHost(b.move_off); cudaFreeHost(b.move_idx); cudaFreeHost(b.priors);
    if (b.done) cudaEventDestroy(b.done);
    if (b.k0) cudaEventDestroy(b.k0);
    if (b.k1) cudaEventDestroy(b.k1);
  }
  for (auto& d : s->slot) {
    cudaFree(d.planes); cudaFree(d.bits); cudaFree(d.scale); cudaFree(d.policy); cudaFree(d.value);
    cudaFree(d.move_off); cudaFree(d.move_idx); cudaFree(d.priors);
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  delete s;
}

}  // namespace

extern "C" {

int e55_service_start(e55_ctx* c, int max_batch, int max_wait_us) {
  if (!c || max_batch < 1 || max_batch > 65536 || max_wait_us < 0) return E55_ERR_INVALID;
  if (c->service) return fail(c, E55_ERR_INVALID, "evaluation service already running");
  E55_CUDA(c, cudaSetDevice(c->device));
  {
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = ensure_capacity(c, max_batch);
    if (rc) return rc;
  }
  e55_service* s = new (std::nothrow) e55_service();
  if (!s) return fail(c, E55_ERR_INVALID, "out of host memory");
  s->cap = max_batch; s->max_wait_us = max_wait_us;
  s->idle_min = std::max(1, max_batch / 4);
  if (const char* env = getenv("E55_SERVICE_IDLE_MIN")) s->idle_min = std::max(1, atoi(env));
  s->trace = getenv("E55_SERVICE_TRACE") != nullptr;
  s->last_complete = std::chrono::steady_clock::now();
  const size_t ip = static_cast<size_t>(c->in_planes), cap = static_cast<size_t>(max_batch);
  const size_t moves = cap * e55_ctx::kMovesPerPosition, pol = cap * kPolicyCh * kSq;
  cudaError_t e = cudaSuccess;
  for (auto& b : s->batch) {
    if (e == cudaSuccess) e = cudaMallocHost(&b.planes, cap * ip * kSq * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&b.bits, cap * ip * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(&b.scale, cap * ip * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&b.policy, pol * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&b.value, cap * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&b.move_off, (cap + 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMallocHost(&b.move_idx, moves * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMallocHost(&b.priors, moves * sizeof(float));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&b.k0);
    if (e == cudaSuccess) e = cudaEventCreate(&b.k1);
  }
  for (auto& d : s->slot) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&d.planes, cap * ip * kSq * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d.bits, cap * ip * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d.scale, cap * ip * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d.policy, pol * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d.value, cap * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d.move_off, (cap + 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d.move_idx, moves * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMalloc(&d.priors, moves * sizeof(float));
  }
  if (e != cudaSuccess) { service_free(s); return fail(c, E55_ERR_CUDA, std::string("evaluation service buffers: ") + cudaGetErrorString(e)); }
  c->service = s;
  s->dispatcher = std::thread(service_dispatch_loop, c);
  s->completer = std::thread(service_complete_loop, c);
  return E55_OK;
}

int e55_service_stop(e55_ctx* c) {
  if (!c) return E55_ERR_INVALID;
  e55_service* s = c->service;
  if (!s) return E55_OK;
  {
    std::lock_guard<std::mutex> lk(s->m);
    s->stop = true;
  }
  s->cv_submit.notify_all();
  s->cv_space.notify_all();
  s->cv_free.notify_all();
  s->cv_slot.notify_all();
  s->dispatcher.join();
  s->completer.join();
  if (s->trace && s->n_batches > 0) {
    const double nb = static_cast<double>(s->n_batches);
    fprintf(stderr, "[e55 service] %lld batches, mean %.1f positions; per batch (us): fill %.1f, wait for slot %.1f, wait for copies %.1f, "
            "launch %.1f, launch->complete %.1f, GPU idle before launch %.1f, kernels %.1f, between kernel ends %.1f; closed: %lld full, %lld idle-GPU, %lld timeout, %lld forced\n",
            static_cast<long long>(s->n_batches), static_cast<double>(s->n_positions) / nb, s->t_fill / nb, s->t_slot / nb, s->t_copywait / nb,
            s->t_launch / nb, s->t_gpu / nb, s->t_idle / nb, s->n_kernels ? s->t_kernels / s->n_kernels : 0.0, s->n_kernels ? s->t_interval / s->n_kernels : 0.0,
            static_cast<long long>(s->closed_full), static_cast<long long>(s->closed_idle),
            static_cast<long long>(s->closed_timeout), static_cast<long long>(s->closed_forced));
  }
  c->service = nullptr;
  service_free(s);
  return E55_OK;
}

int e55_service_predict(e55_ctx* c, const float* planes, int n, float* policy, float* value) {
  if (!c) return E55_ERR_INVALID;
  return service_predict(c, 1, planes, nullptr, nullptr, n, policy, value);
}

int e55_service_predict_packed(e55_ctx* c, const uint32_t* bits, const float* scale, int n, float* policy, float* value) {
  if (!c) return E55_ERR_INVALID;
  return service_predict(c, 2, nullptr, bits, scale, n, policy, value);
}

int e55_service_predict_packed_moves(e55_ctx* c, const uint32_t* bits, const float* scale, int n, const int32_t* move_offset,
                                     const uint16_t* move_index, float* priors, float* value) {
  if (!c) return E55_ERR_INVALID;
  return service_predict(c, 3, nullptr, bits, scale, n, priors, value, move_offset, move_index);
}

static_assert(sizeof(Ticket) == sizeof(e55_ticket), "e55_ticket layout");

int e55_service_submit_packed_moves(e55_ctx* c, const uint32_t* bits, const float* scale, int n, const int32_t* move_offset,
                                    const uint16_t* move_index, int may_block, e55_ticket* ticket_out) {
  if (!c || !ticket_out) return E55_ERR_INVALID;
  return service_submit(c, 3, nullptr, bits, scale, n, move_offset, move_index, may_block != 0, reinterpret_cast<Ticket*>(ticket_out));
}

int e55_service_fetch(e55_ctx* c, const e55_ticket* ticket, float* out, float* value, int wait) {
  if (!c || !ticket) return E55_ERR_INVALID;
  Ticket t;
  memcpy(&t, ticket, sizeof(t));
  return service_fetch(c, t, out, value, wait != 0);
}

int e55_service_stats(e55_ctx* c, int64_t* batches_out, int64_t* positions_out) {
  if (!c || !batches_out || !positions_out) return E55_ERR_INVALID;
  e55_service* s = c->service;
  if (!s) return fail(c, E55_ERR_INVALID, "evaluation service is not running");
  std::lock_guard<std::mutex> lk(s->m);
  *batches_out = s->n_batches;
  *positions_out = s->n_positions;
  return E55_OK;
}

// Synthetic self-play load (tree operations excluded): what the evaluation service itself sustains.  `workers`
// threads each keep `leaf_batch`-position requests of the kind the search sends (bit-packed planes + ~30 legal moves
// per position, mcts.py:91-106) in flight on 48 asynchronous tickets per thread (enough to keep the GPU fed)
// until every worker has had simulations / leaf_batch requests answered for each of its `moves` moves.
int e55_bench_selfplay(e55_ctx* c, int workers, int moves, int simulations, int leaf_batch, double* moves_per_s_out,
                       double* evals_per_s_out, double* mean_batch_out) {
  if (!c || workers < 1 || moves < 1 || simulations < 1 || leaf_batch < 1 || leaf_batch > 256 || !moves_per_s_out || !evals_per_s_out || !mean_batch_out)
    return E55_ERR_INVALID;
  if (!c->service) return fail(c, E55_ERR_INVALID, "evaluation service is not running");
  const int ip = c->in_planes;
  constexpr int kMaxTickets = 64, kMovesPerPosition = 30;
  const int kTickets = getenv("E55_BENCH_TICKETS") ? std::min(kMaxTickets, std::max(1, atoi(getenv("E55_BENCH_TICKETS")))) : 48;
  std::atomic<int> failures{0};
  int64_t b0 = 0, p0 = 0, b1 = 0, p1 = 0;
  e55_service_stats(c, &b0, &p0);
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int w = 0; w < workers; ++w)
    th.emplace_back([&, w] {
      std::vector<uint32_t> bits(static_cast<size_t>(leaf_batch) * ip);
      std::vector<float> scale(bits.size(), 1.f), priors(static_cast<size_t>(leaf_batch) * kMovesPerPosition), value(leaf_batch);
      std::vector<int32_t> off(leaf_batch + 1);
      std::vector<uint16_t> idx(priors.size());
      uint32_t x = 0x9E3779B9u * (w + 1);
      for (auto& v : bits) { x = x * 1664525u + 1013904223u; v = (x >> 7) & (x >> 13) & (x >> 19) & 0x1FFFFFFu; }  // ~12 % density
      for (int i = 0; i <= leaf_batch; ++i) off[i] = i * kMovesPerPosition;
      for (auto& v : idx) { x = x * 1664525u + 1013904223u; v = static_cast<uint16_t>((x >> 8) % (kPolicyCh * kSq)); }
      const long requests = static_cast<long>(moves) * ((simulations + leaf_batch - 1) / leaf_batch);
      Ticket tickets[kMaxTickets];
      bool live[kMaxTickets] = {};
      long sent = 0, answered = 0;
      while (answered < requests && !failures.load()) {
        bool progressed = false;
        int oldest = -1;
        for (int k = 0; k < kTickets; ++k) {
          if (live[k]) {
            const int rc = service_fetch(c, tickets[k], priors.data(), value.data(), false);
            if (rc == E55_PENDING) { if (oldest < 0 || tickets[k].gen < tickets[oldest].gen) oldest = k; continue; }
            if (rc != E55_OK) { ++failures; break; }
            live[k] = false; ++answered; progressed = true;
          }
          if (!live[k] && sent < requests) {
            const int rc = service_submit(c, 3, nullptr, bits.data(), scale.data(), leaf_batch, off.data(), idx.data(), false, &tickets[k]);
            if (rc == E55_OK) { live[k] = true; ++sent; progressed = true; }
            else if (rc != E55_PENDING) { ++failures; break; }
          }
        }
        if (!progressed && oldest >= 0) {   // nothing moved: sleep on an outstanding answer
          if (service_fetch(c, tickets[oldest], priors.data(), value.data(), true) != E55_OK) { ++failures; break; }
          live[oldest] = false; ++answered;
        } else if (!progressed) {
          std::this_thread::yield();
        }
      }
      for (int k = 0; k < kTickets; ++k)     // every ticket has to be fetched
        if (live[k]) service_fetch(c, tickets[k], priors.data(), value.data(), true);
    });
  for (auto& t : th) t.join();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (failures.load()) return E55_ERR_CUDA;
  e55_service_stats(c, &b1, &p1);
  *moves_per_s_out = static_cast<double>(workers) * moves / secs;
  *evals_per_s_out = static_cast<double>(p1 - p0) / secs;
  *mean_batch_out = b1 > b0 ? static_cast<double>(p1 - p0) / static_cast<double>(b1 - b0) : 0.0;
  return E55_OK;
}

}  // extern "C"

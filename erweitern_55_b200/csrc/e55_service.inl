// Cross-game leaf aggregation: many search threads call a blocking predict with the
// small leaf batches mcts.py produces (root: 1, leaves: Config.batch_size = 16; mcts.py:57-58,91-106) and the
// service coalesces them into one large forward per GPU.  Included by e55_api.cu (needs e55_ctx internals).
//
// Two pinned staging batches alternate: while one is in flight on the GPU the other one fills.  Submitters
// reserve a slot range under the lock and copy their own input / fetch their own output outside it, so host
// copies run in parallel on the callers' threads; the dispatcher thread only closes batches and drives CUDA.

namespace {

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
  std::atomic<int> filled{0};   // positions whose input copy has finished
  std::atomic<int> readers{0};              // requests that still have to fetch their output
  std::atomic<uint64_t> generation{0};      // bumped when the batch's results are ready (futex-style wait/notify)
  int status = 0;               // result code of the forward
  bool in_flight = false;
  std::chrono::steady_clock::time_point first_arrival;
};

}  // namespace

struct e55_service {
  std::thread worker;
  std::mutex m;
  std::condition_variable cv_submit;   // dispatcher waits for work
  std::condition_variable cv_space;    // submitters wait for room in the open batch
  ServiceBatch batch[2];
  int open = 0;                        // index of the batch currently accepting requests
  int cap = 0, max_wait_us = 0;
  bool stop = false;
  bool force_close = false;            // a request does not fit the open batch: dispatch it now
  int64_t n_batches = 0, n_positions = 0;
};

namespace {

int service_forward(e55_ctx* c, ServiceBatch& b) {
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  const int n = b.reserved;
  int rc = ensure_capacity(c, n);
  if (rc) return rc;
  const size_t pol = static_cast<size_t>(n) * kPolicyCh * kSq;
  if (b.kind == 1) {
    E55_CUDA(c, cudaMemcpyAsync(c->d_in, b.planes, static_cast<size_t>(n) * c->in_planes * kSq * sizeof(float),
                                cudaMemcpyHostToDevice, c->stream));
    if ((rc = forward(c, c->d_in, n, c->d_policy, c->d_value))) return rc;
  } else {
    const size_t w = static_cast<size_t>(n) * c->in_planes;
    E55_CUDA(c, cudaMemcpyAsync(c->d_bits, b.bits, w * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    E55_CUDA(c, cudaMemcpyAsync(c->d_scale, b.scale, w * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (b.kind == 3) {
      b.move_off[n] = b.moves_reserved;
      E55_CUDA(c, cudaMemcpyAsync(c->d_move_off, b.move_off, (static_cast<size_t>(n) + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
      if (b.moves_reserved)
        E55_CUDA(c, cudaMemcpyAsync(c->d_move_idx, b.move_idx, static_cast<size_t>(b.moves_reserved) * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    }
    if ((rc = forward_packed_on(c, c->stream, c->d_bits, c->d_scale, n, 0, c->d_policy, c->d_value))) return rc;
  }
  if (b.kind == 3) {   // only the legal moves' priors travel back
    e55::legal_priors_kernel<<<(n + 7) / 8, 256, 0, c->stream>>>(c->d_policy, c->d_move_off, c->d_move_idx, c->d_priors, n, kPolicyCh * kSq);
    ++c->launches;
    E55_CUDA(c, cudaGetLastError());
    if (b.moves_reserved)
      E55_CUDA(c, cudaMemcpyAsync(b.priors, c->d_priors, static_cast<size_t>(b.moves_reserved) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  } else {
    E55_CUDA(c, cudaMemcpyAsync(b.policy, c->d_policy, pol * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  }
  E55_CUDA(c, cudaMemcpyAsync(b.value, c->d_value, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  return E55_OK;
}

void service_loop(e55_ctx* c) {
  e55_service* s = c->service;
  std::unique_lock<std::mutex> lk(s->m);
  for (;;) {
    ServiceBatch& b = s->batch[s->open];
    s->cv_submit.wait(lk, [&] { return s->stop || b.reserved > 0; });
    if (s->stop && b.reserved == 0) return;
    // give other games a moment to join, unless the batch is already full
    const auto deadline = b.first_arrival + std::chrono::microseconds(s->max_wait_us);
    while (!s->stop && !s->force_close && b.reserved < s->cap && std::chrono::steady_clock::now() < deadline)
      s->cv_submit.wait_until(lk, deadline);
    s->force_close = false;
    // close it: the other staging batch becomes the open one as soon as its readers are gone
    ServiceBatch& next = s->batch[s->open ^ 1];
    lk.unlock();
    for (int r; (r = next.readers.load(std::memory_order_acquire)) != 0;) next.readers.wait(r);
    lk.lock();
    next.kind = 0; next.reserved = 0; next.moves_reserved = 0; next.filled.store(0);
    b.in_flight = true;
    s->open ^= 1;
    s->cv_space.notify_all();
    lk.unlock();
    while (b.filled.load(std::memory_order_acquire) < b.reserved) std::this_thread::yield();
    const int rc = service_forward(c, b);
    b.status = rc;
    b.generation.fetch_add(1, std::memory_order_release);   // waiters wake without touching the mutex
    b.generation.notify_all();
    lk.lock();
    b.in_flight = false;
    ++s->n_batches;
    s->n_positions += b.reserved;
  }
}

// Blocking aggregated predict: kind 1 (planes), 2 (bits + scale) or 3 (bits + scale + legal-move lists; `policy`
// then receives the priors of those moves, move_offset[n] floats).
int service_predict(e55_ctx* c, int kind, const float* planes, const uint32_t* bits, const float* scale, int n,
                    float* policy, float* value, const int32_t* move_offset = nullptr, const uint16_t* move_index = nullptr) {
  e55_service* s = c->service;
  auto fail_locked = [&](int code, const char* msg) {   // many threads call in: serialise the error string
    std::lock_guard<std::mutex> g(c->mu);
    return fail(c, code, msg);
  };
  if (!s) return fail_locked(E55_ERR_INVALID, "evaluation service is not running (e55_service_start)");
  if (n <= 0 || !policy || !value || (kind == 1 ? !planes : (!bits || !scale)) || (kind == 3 && (!move_offset || !move_index)))
    return fail_locked(E55_ERR_INVALID, "bad service request");
  if (n > s->cap) return fail_locked(E55_ERR_INVALID, "request larger than the service batch");
  const int total = kind == 3 ? move_offset[n] : 0;
  if (kind == 3 && (move_offset[0] != 0 || total < 0 || total > s->cap * e55_ctx::kMovesPerPosition))
    return fail_locked(E55_ERR_INVALID, "bad move lists in service request");
  if (!c->has_weights) return fail_locked(E55_ERR_NO_WEIGHTS, "predict called before e55_set_weights");
  std::unique_lock<std::mutex> lk(s->m);
  ServiceBatch* b = nullptr;
  for (;;) {
    if (s->stop) { lk.unlock(); return fail_locked(E55_ERR_INVALID, "evaluation service stopped"); }
    b = &s->batch[s->open];
    if (!b->in_flight && (b->kind == 0 || b->kind == kind) && b->reserved + n <= s->cap &&
        b->moves_reserved + total <= s->cap * e55_ctx::kMovesPerPosition) break;
    if (b->reserved > 0) s->force_close = true;   // the open batch cannot take this request: dispatch it now
    s->cv_submit.notify_one();
    s->cv_space.wait(lk);
  }
  const int off = b->reserved;
  if (off == 0) { b->kind = kind; b->first_arrival = std::chrono::steady_clock::now(); }
  b->reserved += n;
  const int moff = b->moves_reserved;
  b->moves_reserved += total;
  b->readers.fetch_add(1, std::memory_order_relaxed);
  const uint64_t gen = b->generation.load(std::memory_order_relaxed);
  s->cv_submit.notify_one();
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

  while (b->generation.load(std::memory_order_acquire) == gen) b->generation.wait(gen);
  const int rc = b->status;
  if (rc == E55_OK) {
    if (kind == 3) memcpy(policy, b->priors + moff, static_cast<size_t>(total) * sizeof(float));
    else memcpy(policy, b->policy + static_cast<size_t>(off) * kPolicyCh * kSq, static_cast<size_t>(n) * kPolicyCh * kSq * sizeof(float));
    memcpy(value, b->value + off, static_cast<size_t>(n) * sizeof(float));
  }
  if (b->readers.fetch_sub(1, std::memory_order_release) == 1) b->readers.notify_all();
  return rc;
}

void service_free(e55_service* s) {
  for (auto& b : s->batch) {
    cudaFreeHost(b.planes); cudaFreeHost(b.bits); cudaFreeHost(b.scale); cudaFreeHost(b.policy); cudaFreeHost(b.value);
    cudaFreeHost(b.move_off); cudaFreeHost(b.move_idx); cudaFreeHost(b.priors);
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
  const size_t ip = static_cast<size_t>(c->in_planes);
  for (auto& b : s->batch) {
    cudaError_t e = cudaMallocHost(&b.planes, max_batch * ip * kSq * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&b.bits, max_batch * ip * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(&b.scale, max_batch * ip * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&b.policy, static_cast<size_t>(max_batch) * kPolicyCh * kSq * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&b.value, max_batch * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&b.move_off, (static_cast<size_t>(max_batch) + 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMallocHost(&b.move_idx, static_cast<size_t>(max_batch) * e55_ctx::kMovesPerPosition * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMallocHost(&b.priors, static_cast<size_t>(max_batch) * e55_ctx::kMovesPerPosition * sizeof(float));
    if (e != cudaSuccess) { service_free(s); return fail(c, E55_ERR_CUDA, std::string("pinned staging: ") + cudaGetErrorString(e)); }
  }
  c->service = s;
  s->worker = std::thread(service_loop, c);
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
  s->worker.join();
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

int e55_service_stats(e55_ctx* c, int64_t* batches_out, int64_t* positions_out) {
  if (!c || !batches_out || !positions_out) return E55_ERR_INVALID;
  e55_service* s = c->service;
  if (!s) return fail(c, E55_ERR_INVALID, "evaluation service is not running");
  std::lock_guard<std::mutex> lk(s->m);
  *batches_out = s->n_batches;
  *positions_out = s->n_positions;
  return E55_OK;
}

// Synthetic self-play load (tree operations excluded): `workers` threads each play `moves` moves the way
// mcts.MCTS.run drives the network -- one root evaluation of batch 1, then simulations / leaf_batch
// evaluations of leaf_batch positions (mcts.py:57-58,70,91-106) -- on random packed positions, all through
// the evaluation service.
int e55_bench_selfplay(e55_ctx* c, int workers, int moves, int simulations, int leaf_batch, double* moves_per_s_out,
                       double* evals_per_s_out, double* mean_batch_out) {
  if (!c || workers < 1 || moves < 1 || simulations < 1 || leaf_batch < 1 || !moves_per_s_out || !evals_per_s_out || !mean_batch_out)
    return E55_ERR_INVALID;
  if (!c->service) return fail(c, E55_ERR_INVALID, "evaluation service is not running");
  const int ip = c->in_planes;
  std::atomic<int> failures{0};
  int64_t b0 = 0, p0 = 0, b1 = 0, p1 = 0;
  e55_service_stats(c, &b0, &p0);
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int w = 0; w < workers; ++w)
    th.emplace_back([&, w] {
      std::vector<uint32_t> bits(static_cast<size_t>(leaf_batch) * ip);
      std::vector<float> scale(bits.size(), 1.f), policy(static_cast<size_t>(leaf_batch) * kPolicyCh * kSq), value(leaf_batch);
      uint32_t x = 0x9E3779B9u * (w + 1);
      for (auto& v : bits) { x = x * 1664525u + 1013904223u; v = (x >> 7) & (x >> 13) & (x >> 19) & 0x1FFFFFFu; }  // ~12 % density
      const int iters = (simulations + leaf_batch - 1) / leaf_batch;
      for (int mv = 0; mv < moves && !failures.load(); ++mv) {
        if (e55_service_predict_packed(c, bits.data(), scale.data(), 1, policy.data(), value.data()) != E55_OK) { ++failures; break; }
        for (int it = 0; it < iters; ++it)
          if (e55_service_predict_packed(c, bits.data(), scale.data(), leaf_batch, policy.data(), value.data()) != E55_OK) { ++failures; break; }
      }
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

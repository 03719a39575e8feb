// Host driver of the GPU-resident self-play (e55_gpu_selfplay.cuh).  Included by e55_api.cu.

namespace {

std::once_flag g_engine_once;
cudaError_t g_engine_status = cudaSuccess;

// One device copy of the engine's attack tables, and room for the engine's recursion (pawn-drop mate test, perft).
cudaError_t device_engine_ready() {
  std::call_once(g_engine_once, [] {
    mshogi::Tables* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(mshogi::Tables));
    if (e == cudaSuccess) e = cudaMemcpy(d, &mshogi::tables(), sizeof(mshogi::Tables), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(mshogi::g_device_tables, &d, sizeof(d));
    if (e == cudaSuccess) e = cudaDeviceSetLimit(cudaLimitStackSize, 24 * 1024);
    g_engine_status = e;
  });
  return g_engine_status;
}

template <class T>
cudaError_t dev_alloc(T*& ptr, size_t count, std::vector<void*>& owned) {
  cudaError_t e = cudaMalloc(&ptr, count * sizeof(T));
  if (e == cudaSuccess) { owned.push_back(ptr); e = cudaMemset(ptr, 0, count * sizeof(T)); }
  return e;
}

}  // namespace

extern "C" {

int e55_selftest_engine(int device, int depth, int64_t* nodes_out) {
  if (depth < 1 || depth > 6 || !nodes_out) return E55_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess || device_engine_ready() != cudaSuccess) return E55_ERR_CUDA;
  long long* d_out = nullptr;
  if (cudaMalloc(&d_out, sizeof(long long)) != cudaSuccess) return E55_ERR_CUDA;
  e55::gs::perft_kernel<<<1, 32>>>(depth, d_out);
  long long h = 0;
  const cudaError_t e = cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d_out);
  if (e != cudaSuccess) return E55_ERR_CUDA;
  *nodes_out = h;
  return E55_OK;
}

int e55_gpu_selfplay_run(e55_ctx* c, int games, long total_moves, int simulations, int use_dirichlet, int sampling_moves,
                         int max_moves, uint64_t seed, e55_selfplay_stats* out, uint16_t* records, int record_cap,
                         int* n_records_out) {
  namespace gs = e55::gs;
  if (!c || !out || games < 1 || games > 65536 || total_moves < 1 || simulations < gs::kLeaf || simulations > 1024 ||
      simulations % gs::kLeaf != 0 || max_moves < 1 || record_cap < 0 || (record_cap > 0 && (!records || !n_records_out)))
    return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->has_weights) return fail(c, E55_ERR_NO_WEIGHTS, "self-play before e55_set_weights");
  if (c->precision != E55_PRECISION_BF16_TC || c->in_planes != mshogi::kInputPlanes || c->filters != 128)
    return fail(c, E55_ERR_UNSUPPORTED, "GPU-resident self-play runs the 266-plane, 128-filter network on the tensor-core path");
  E55_CUDA(c, cudaSetDevice(c->device));
  E55_CUDA(c, device_engine_ready());
  std::vector<void*> owned;
  gs::Params p{};
  const size_t G = static_cast<size_t>(games), slots = G * gs::kLeaf;
  float* d_policy = nullptr;
  cudaError_t e = cudaSuccess;
  auto chain = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  chain(dev_alloc(p.games, G, owned));
  chain(dev_alloc(p.nodes, G * gs::kMaxNodes, owned));
  chain(dev_alloc(p.edges, G * gs::kMaxEdges, owned));
  chain(dev_alloc(p.child, G * gs::kMaxEdges, owned));
  chain(dev_alloc(p.bits, slots * mshogi::kInputPlanes, owned));
  chain(dev_alloc(p.scale, slots * mshogi::kInputPlanes, owned));
  chain(dev_alloc(p.move_idx, slots * gs::kMoveStride, owned));
  chain(dev_alloc(p.move_cnt, slots, owned));
  chain(dev_alloc(p.priors, slots * gs::kMoveStride, owned));
  chain(dev_alloc(p.value, slots, owned));
  chain(dev_alloc(p.shared, 1, owned));
  chain(dev_alloc(d_policy, slots * kPolicyCh * kSq, owned));
  if (record_cap > 0) chain(dev_alloc(p.records, static_cast<size_t>(record_cap) * (gs::kRecordMoves + 2), owned));
  gs::Shared* h_shared = nullptr;
  chain(cudaMallocHost(&h_shared, sizeof(gs::Shared)));
  auto cleanup = [&] {
    for (void* q : owned) cudaFree(q);
    if (h_shared) cudaFreeHost(h_shared);
  };
  if (e != cudaSuccess) { cleanup(); return fail(c, E55_ERR_CUDA, std::string("GPU self-play buffers: ") + cudaGetErrorString(e)); }
  p.n_games = games; p.simulations = simulations; p.use_dirichlet = use_dirichlet; p.sampling_moves = sampling_moves;
  p.max_moves = std::min(max_moves, gs::kMaxPly); p.record_cap = record_cap; p.move_budget = total_moves;
  p.c_puct = 1.25f; p.dirichlet_alpha = 0.15f; p.dirichlet_eps = 0.25f;
  cudaStream_t st = c->stream;
  const int blocks = (games + 63) / 64, n = static_cast<int>(slots);
  const auto t0 = std::chrono::steady_clock::now();
  gs::init_kernel<<<blocks, 64, 0, st>>>(p, seed);
  int rc = E55_OK;
  memset(h_shared, 0, sizeof(*h_shared));
  for (long round = 0; rc == E55_OK; ++round) {
    gs::select_kernel<<<blocks, 64, 0, st>>>(p);
    ++c->launches;
    rc = forward_packed_on(c, st, p.bits, p.scale, n, 0, d_policy, p.value);
    if (rc) break;
    gs::legal_priors_strided_kernel<<<(n + 7) / 8, 256, 0, st>>>(d_policy, p.move_cnt, p.move_idx, p.priors, n, kPolicyCh * kSq, gs::kMoveStride);
    gs::expand_kernel<<<blocks, 64, 0, st>>>(p);
    c->launches += 2;
    if (round % 4 == 3) {   // look at the counters now and then
      e = cudaMemcpyAsync(h_shared, p.shared, sizeof(gs::Shared), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { rc = fail(c, E55_ERR_CUDA, std::string("GPU self-play: ") + cudaGetErrorString(e)); break; }
      if (static_cast<long>(h_shared->moves) >= total_moves) break;
    }
  }
  if (rc == E55_OK) {
    e = cudaMemcpyAsync(h_shared, p.shared, sizeof(gs::Shared), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && record_cap > 0) {
      const int nrec = std::min(h_shared->n_records, record_cap);
      *n_records_out = nrec;
      if (nrec > 0) e = cudaMemcpy(records, p.records, static_cast<size_t>(nrec) * (gs::kRecordMoves + 2) * sizeof(uint16_t), cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) rc = fail(c, E55_ERR_CUDA, std::string("GPU self-play: ") + cudaGetErrorString(e));
  }
  out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  out->moves = static_cast<long>(h_shared->moves);
  out->evals = static_cast<long>(h_shared->evals);
  out->games = static_cast<long>(h_shared->games);
  out->mean_game_plies = h_shared->games ? static_cast<double>(h_shared->plies_in_finished_games) / static_cast<double>(h_shared->games) : 0.0;
  out->mean_gpu_batch = static_cast<double>(n);
  if (rc == E55_OK && h_shared->pool_overflows)
    fprintf(stderr, "[e55 gpu self-play] %llu tree-pool overflows (searches cut short)\n", h_shared->pool_overflows);
  cleanup();
  return rc;
}

}  // extern "C"

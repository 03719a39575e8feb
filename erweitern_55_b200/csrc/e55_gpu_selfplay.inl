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

// Zeroed device memory; the zeroing is ordered on `stream` like everything else the run does (a cudaMemset on the
// default stream is not ordered against kernels on a non-blocking stream and could wipe what the first kernel wrote).
template <class T>
cudaError_t dev_alloc(T*& ptr, size_t count, std::vector<void*>& owned, cudaStream_t stream) {
  cudaError_t e = cudaMalloc(&ptr, count * sizeof(T));
  if (e == cudaSuccess) { owned.push_back(ptr); e = cudaMemsetAsync(ptr, 0, count * sizeof(T), stream); }
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

int e55_selftest_encode(int device, int n, uint64_t seed, int64_t* mismatches_out, int64_t* repetitions_out, int64_t* path_plies_out) {
  namespace gs = e55::gs;
  if (n < 1 || n > 65536 || !mismatches_out || !repetitions_out || !path_plies_out) return E55_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess || device_engine_ready() != cudaSuccess) return E55_ERR_CUDA;
  const size_t N = static_cast<size_t>(n);
  gs::Game* d_games = nullptr;
  uint32_t* d_bits = nullptr;
  float* d_scale = nullptr;
  uint16_t* d_moves = nullptr;
  int32_t* d_info = nullptr;
  cudaError_t e = cudaMalloc(&d_games, N * sizeof(gs::Game));
  if (e == cudaSuccess) e = cudaMalloc(&d_bits, N * mshogi::kInputPlanes * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_scale, N * mshogi::kInputPlanes * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_moves, N * gs::kSelftestMoves * sizeof(uint16_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_info, N * 4 * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemset(d_games, 0, N * sizeof(gs::Game));
  std::vector<uint32_t> bits(N * mshogi::kInputPlanes);
  std::vector<float> scale(N * mshogi::kInputPlanes);
  std::vector<uint16_t> moves(N * gs::kSelftestMoves);
  std::vector<int32_t> info(N * 4);
  if (e == cudaSuccess) {
    gs::encode_selftest_kernel<<<(n + 63) / 64, 64>>>(d_games, n, seed, d_bits, d_scale, d_moves, d_info);
    e = cudaDeviceSynchronize();
  }
  if (e == cudaSuccess) e = cudaMemcpy(bits.data(), d_bits, bits.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(scale.data(), d_scale, scale.size() * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(moves.data(), d_moves, moves.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(info.data(), d_info, info.size() * sizeof(int32_t), cudaMemcpyDeviceToHost);
  cudaFree(d_games); cudaFree(d_bits); cudaFree(d_scale); cudaFree(d_moves); cudaFree(d_info);
  if (e != cudaSuccess) { cudaGetLastError(); return E55_ERR_CUDA; }
  int64_t bad = 0, reps = 0, path = 0;
  uint32_t hb[mshogi::kInputPlanes];
  float hs[mshogi::kInputPlanes];
  for (size_t i = 0; i < N; ++i) {
    mshogi::Position pos;
    const int n_moves = info[i * 4];
    bool legal = n_moves >= 0 && n_moves <= gs::kSelftestMoves;
    for (int k = 0; legal && k < n_moves; ++k) {
      const mshogi::Move mv = gs::unpack_move(moves[i * gs::kSelftestMoves + k]);
      mshogi::MoveList l;
      pos.generate_moves(l);
      legal = std::find(l.begin(), l.end(), mv) != l.end();
      if (legal) pos.do_move(mv);
    }
    if (!legal) { ++bad; continue; }
    mshogi::encode_packed(pos, hb, hs);
    bool rep, my_check, op_check;
    pos.is_repetition(rep, my_check, op_check);
    mshogi::MoveList l;
    pos.generate_moves(l);
    const int flags = (rep ? 1 : 0) | (my_check ? 2 : 0) | (op_check ? 4 : 0) | (pos.repetition_count() << 8);
    const bool same = memcmp(hb, bits.data() + i * mshogi::kInputPlanes, sizeof(hb)) == 0 &&
                      memcmp(hs, scale.data() + i * mshogi::kInputPlanes, sizeof(hs)) == 0 && flags == info[i * 4 + 2] &&
                      l.size() == info[i * 4 + 3];
    if (!same) ++bad;
    if (pos.repetition_count() >= 2) ++reps;
    path += info[i * 4 + 1];
  }
  *mismatches_out = bad;
  *repetitions_out = reps;
  *path_plies_out = path;
  return E55_OK;
}

int e55_gpu_selfplay_set_cap(e55_ctx* c, int enable, int full_simulations, int fast_simulations, float frac) {
  namespace gs = e55::gs;
  if (!c) return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  if (enable && (full_simulations < gs::kLeaf || full_simulations > 1024 || fast_simulations < 1 || fast_simulations > 1024 || !(frac >= 0.f && frac <= 1.f)))
    return fail(c, E55_ERR_INVALID, "playout-cap oscillation: simulations in [16, 1024] / [1, 1024], frac in [0, 1]");
  c->gpu_cap_enable = enable != 0; c->gpu_cap_full = full_simulations; c->gpu_cap_fast = fast_simulations; c->gpu_cap_frac = frac;
  return E55_OK;
}

// ---- debug stepper: ONE game's search, a round at a time, with the caller playing the network ---------------------------
// (tests hold select_kernel / expand_kernel to oracle/mcts_oracle.py selection by selection with it)
struct e55_gpu_search {
  e55_ctx* c = nullptr;
  e55::gs::Params p{};
  std::vector<void*> owned;
  std::vector<e55::gs::GNode> nodes;
  std::vector<e55::gs::GEdge> edges;
  e55::gs::Game game;
};

int e55_gpu_search_open(e55_ctx* c, const uint16_t* prefix, int n_prefix, int simulations, int reuse_tree, e55_gpu_search** out) {
  namespace gs = e55::gs;
  if (!c || !out || n_prefix < 0 || (n_prefix > 0 && !prefix) || simulations < gs::kLeaf || simulations > 1024 || simulations % gs::kLeaf != 0)
    return E55_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  E55_CUDA(c, device_engine_ready());
  e55_gpu_search* h = new e55_gpu_search();
  h->c = c;
  gs::Params& p = h->p;
  cudaError_t e = cudaSuccess;
  auto chain = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  uint16_t* d_prefix = nullptr;
  unsigned long long* d_ans = nullptr;
  chain(dev_alloc(p.games, 1, h->owned, c->stream));
  chain(dev_alloc(p.nodes, 2 * gs::kMaxNodes, h->owned, c->stream));
  chain(dev_alloc(p.edges, 2 * gs::kMaxEdges, h->owned, c->stream));
  chain(dev_alloc(p.child, 2 * gs::kMaxEdges, h->owned, c->stream));
  chain(dev_alloc(p.old_of, gs::kMaxNodes, h->owned, c->stream));
  chain(dev_alloc(p.bits, gs::kLeaf * mshogi::kInputPlanes, h->owned, c->stream));
  chain(dev_alloc(p.scale, gs::kLeaf * mshogi::kInputPlanes, h->owned, c->stream));
  chain(dev_alloc(p.move_idx, gs::kLeaf * gs::kMoveStride, h->owned, c->stream));
  chain(dev_alloc(p.move_cnt, gs::kLeaf, h->owned, c->stream));
  chain(dev_alloc(p.priors, gs::kLeaf * gs::kMoveStride, h->owned, c->stream));
  chain(dev_alloc(p.value, gs::kLeaf, h->owned, c->stream));
  chain(dev_alloc(p.shared, 1, h->owned, c->stream));
  chain(dev_alloc(p.rec_buf, gs::kRecordWords, h->owned, c->stream));
  chain(dev_alloc(p.records, gs::kRecordWords, h->owned, c->stream));
  chain(dev_alloc(p.req_board, 1, h->owned, c->stream));
  chain(dev_alloc(p.req_seq, 1, h->owned, c->stream));
  chain(dev_alloc(d_ans, 1, h->owned, c->stream));
  chain(dev_alloc(d_prefix, static_cast<size_t>(std::max(n_prefix, 1)), h->owned, c->stream));
  p.ans = d_ans;
  p.n_games = 1; p.game0 = 0; p.simulations = simulations; p.use_dirichlet = 0; p.reuse_tree = reuse_tree; p.mate_depth = 0;
  p.speculate = 0; p.sampling_moves = 0; p.max_moves = gs::kMaxPly; p.record_cap = 1; p.move_budget = 1 << 30;
  p.c_puct = 1.25f; p.dirichlet_alpha = 0.15f; p.dirichlet_eps = 0.25f;
  p.cap_enable = c->gpu_cap_enable; p.cap_full = c->gpu_cap_full; p.cap_fast = c->gpu_cap_fast; p.cap_frac = c->gpu_cap_frac;
  if (p.cap_enable) p.dirichlet_eps = 0.f;   // (the stepper's searches are deterministic: noise of weight 0 leaves the priors alone)
  if (e == cudaSuccess && n_prefix > 0) e = cudaMemcpyAsync(d_prefix, prefix, n_prefix * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    gs::init_kernel<<<1, 64, 0, c->stream>>>(p, 1);
    gs::prefix_kernel<<<1, 64, 0, c->stream>>>(p, d_prefix, n_prefix);
    e = cudaStreamSynchronize(c->stream);
  }
  if (e != cudaSuccess) {
    for (void* q : h->owned) cudaFree(q);
    delete h;
    return fail(c, E55_ERR_CUDA, std::string("GPU search stepper: ") + cudaGetErrorString(e));
  }
  h->nodes.resize(gs::kMaxNodes);
  h->edges.resize(gs::kMaxEdges);
  *out = h;
  return E55_OK;
}

void e55_gpu_search_close(e55_gpu_search* h) {
  if (!h) return;
  cudaSetDevice(h->c->device);
  cudaStreamSynchronize(h->c->stream);
  for (void* q : h->owned) cudaFree(q);
  delete h;
}

// One select_kernel round.  info[8] = {plies played, sims_done, n_sel, n_nodes, n_edges, games_finished, rec_words, 0};
// per selection s < n_sel: sel_leaf[s] (node, -1 = abandoned), sel_slot[s] (network slot or -1), sel_path_len[s] and
// sel_path[s*96 ..] (the moves from the root, packed); per slot k < 16: slot_cnt[k] legal moves, slot_moves[k*256 ..]
// (packed, in edge order) and slot_idx[k*256 ..] (their policy indices).
int e55_gpu_search_select(e55_gpu_search* h, int32_t* info, int32_t* sel_leaf, int32_t* sel_slot, int32_t* sel_path_len, uint16_t* sel_path,
                          int32_t* slot_cnt, uint16_t* slot_moves, uint16_t* slot_idx) {
  namespace gs = e55::gs;
  if (!h || !info || !sel_leaf || !sel_slot || !sel_path_len || !sel_path || !slot_cnt || !slot_moves || !slot_idx) return E55_ERR_INVALID;
  e55_ctx* c = h->c;
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  gs::select_kernel<<<1, 64, 0, c->stream>>>(h->p);
  ++c->launches;
  E55_CUDA(c, cudaGetLastError());
  E55_CUDA(c, cudaMemcpyAsync(&h->game, h->p.games, sizeof(gs::Game), cudaMemcpyDeviceToHost, c->stream));
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  const gs::Game& g = h->game;
  E55_CUDA(c, cudaMemcpy(h->nodes.data(), h->p.nodes + static_cast<size_t>(g.pool) * gs::kMaxNodes, gs::kMaxNodes * sizeof(gs::GNode), cudaMemcpyDeviceToHost));
  E55_CUDA(c, cudaMemcpy(h->edges.data(), h->p.edges + static_cast<size_t>(g.pool) * gs::kMaxEdges, gs::kMaxEdges * sizeof(gs::GEdge), cudaMemcpyDeviceToHost));
  E55_CUDA(c, cudaMemcpy(slot_cnt, h->p.move_cnt, gs::kLeaf * sizeof(int32_t), cudaMemcpyDeviceToHost));
  E55_CUDA(c, cudaMemcpy(slot_idx, h->p.move_idx, gs::kLeaf * gs::kMoveStride * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  info[0] = g.root.ply; info[1] = g.sims_done; info[2] = g.n_sel; info[3] = g.n_nodes; info[4] = g.n_edges; info[5] = g.games_finished;
  info[6] = g.rec_words; info[7] = 0;
  memset(slot_moves, 0, gs::kLeaf * gs::kMoveStride * sizeof(uint16_t));
  for (int s = 0; s < g.n_sel; ++s) {
    sel_leaf[s] = g.leaf[s];
    sel_slot[s] = g.leaf_slot[s];
    int len = 0;
    uint16_t rev[gs::kMaxPath + 1];
    for (int node = g.leaf[s]; node >= 0 && h->nodes[node].parent >= 0 && len <= gs::kMaxPath; node = h->nodes[node].parent)
      rev[len++] = h->edges[h->nodes[node].parent_edge].mv;
    sel_path_len[s] = len;
    for (int i = 0; i < len; ++i) sel_path[s * gs::kMaxPath + i] = rev[len - 1 - i];
    if (g.leaf_slot[s] >= 0 && g.leaf[s] >= 0) {
      const gs::GNode& leaf = h->nodes[g.leaf[s]];
      for (int i = 0; i < leaf.n_edges && i < gs::kMoveStride; ++i) slot_moves[g.leaf_slot[s] * gs::kMoveStride + i] = h->edges[leaf.first_edge + i].mv;
    }
  }
  return E55_OK;
}

// The network's answers for the round's slots (priors float32 [16][256] in edge order, value float32 [16] in [-1, 1]), then
// one expand_kernel round.  info[8] as above, after the round; record receives the game's record so far (info[6] words:
// per ply move, sum_N, Q * 65535, k and k (move, N) pairs), or, when the round has ended a game (info[5] counts them),
// the finished game: its four header words first, info[6] = words behind them.
int e55_gpu_search_expand(e55_gpu_search* h, const float* priors, const float* value, int32_t* info, uint16_t* record) {
  namespace gs = e55::gs;
  if (!h || !priors || !value || !info || !record) return E55_ERR_INVALID;
  e55_ctx* c = h->c;
  std::lock_guard<std::mutex> lk(c->mu);
  E55_CUDA(c, cudaSetDevice(c->device));
  const int finished_before = h->game.games_finished;
  E55_CUDA(c, cudaMemcpyAsync(h->p.priors, priors, gs::kLeaf * gs::kMoveStride * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  E55_CUDA(c, cudaMemcpyAsync(h->p.value, value, gs::kLeaf * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  gs::expand_kernel<<<1, 64, 0, c->stream>>>(h->p);
  ++c->launches;
  E55_CUDA(c, cudaGetLastError());
  E55_CUDA(c, cudaMemcpyAsync(&h->game, h->p.games, sizeof(gs::Game), cudaMemcpyDeviceToHost, c->stream));
  E55_CUDA(c, cudaStreamSynchronize(c->stream));
  const gs::Game& g = h->game;
  info[0] = g.root.ply; info[1] = g.sims_done; info[2] = g.n_sel; info[3] = g.n_nodes; info[4] = g.n_edges; info[5] = g.games_finished;
  info[6] = g.rec_words; info[7] = 0;
  if (g.games_finished != finished_before) {
    E55_CUDA(c, cudaMemcpy(record, h->p.records, gs::kRecordWords * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    info[6] = record[2];
    info[7] = 1;
    gs::Shared zero{};
    E55_CUDA(c, cudaMemcpy(h->p.shared, &zero, sizeof(zero), cudaMemcpyHostToDevice));   // (the next finished game is record 0 again)
  } else {
    E55_CUDA(c, cudaMemcpy(record, h->p.rec_buf, gs::kRecordWords * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  }
  return E55_OK;
}

int e55_gpu_selfplay_run(e55_ctx* c, int games, long total_moves, int simulations, int mate_depth, int reuse_tree, int use_dirichlet,
                         int sampling_moves, int max_moves, uint64_t seed, e55_selfplay_stats* out, uint16_t* records, int record_cap,
                         int* n_records_out) {
  namespace gs = e55::gs;
  if (!c || !out || games < 1 || games > 65536 || total_moves < 1 || simulations < gs::kLeaf || simulations > 1024 ||
      simulations % gs::kLeaf != 0 || max_moves < 1 || mate_depth < 0 || mate_depth > 9 || record_cap < 0 || (record_cap > 0 && (!records || !n_records_out)))
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
  cudaError_t e = cudaSuccess;
  auto chain = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  chain(dev_alloc(p.games, G, owned, c->stream));
  chain(dev_alloc(p.nodes, G * 2 * gs::kMaxNodes, owned, c->stream));
  chain(dev_alloc(p.edges, G * 2 * gs::kMaxEdges, owned, c->stream));
  chain(dev_alloc(p.child, G * 2 * gs::kMaxEdges, owned, c->stream));
  chain(dev_alloc(p.old_of, G * gs::kMaxNodes, owned, c->stream));
  chain(dev_alloc(p.bits, slots * mshogi::kInputPlanes, owned, c->stream));
  chain(dev_alloc(p.scale, slots * mshogi::kInputPlanes, owned, c->stream));
  chain(dev_alloc(p.move_idx, slots * gs::kMoveStride, owned, c->stream));
  chain(dev_alloc(p.move_cnt, slots, owned, c->stream));
  chain(dev_alloc(p.priors, slots * gs::kMoveStride, owned, c->stream));
  chain(dev_alloc(p.value, slots, owned, c->stream));
  chain(dev_alloc(p.shared, 1, owned, c->stream));
  if (record_cap > 0) {
    chain(dev_alloc(p.records, static_cast<size_t>(record_cap) * gs::kRecordWords, owned, c->stream));
    chain(dev_alloc(p.rec_buf, G * gs::kRecordWords, owned, c->stream));
  }
  // mate-search mailbox (see Params): requests come down every round, answers go up a round or two later
  unsigned long long* d_ans = nullptr;
  chain(dev_alloc(p.req_board, G, owned, c->stream));
  chain(dev_alloc(p.req_seq, G, owned, c->stream));
  chain(dev_alloc(d_ans, G, owned, c->stream));
  p.ans = d_ans;
  gs::Shared* h_shared = nullptr;            // [2 cohorts][2]: one per cohort and round in flight
  mshogi::Board* h_board = nullptr;          // [2][G]
  uint32_t* h_seq = nullptr;                 // [2][G]
  unsigned long long* h_ans = nullptr;       // [2][G]
  chain(cudaMallocHost(&h_shared, 4 * sizeof(gs::Shared)));
  chain(cudaMallocHost(&h_board, 2 * G * sizeof(mshogi::Board)));
  chain(cudaMallocHost(&h_seq, 2 * G * sizeof(uint32_t)));
  chain(cudaMallocHost(&h_ans, 2 * G * sizeof(unsigned long long)));
  cudaEvent_t ev[2][2] = {};
  cudaEvent_t ready = nullptr;
  for (auto& pair : ev)
    for (auto& x : pair) chain(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
  chain(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  const int sms_before = c->tc.sms;
  auto cleanup = [&] {
    c->tc.sms = sms_before;
    for (void* q : owned) cudaFree(q);
    cudaFreeHost(h_shared); cudaFreeHost(h_board); cudaFreeHost(h_seq); cudaFreeHost(h_ans);
    for (auto& pair : ev)
      for (auto& x : pair) if (x) cudaEventDestroy(x);
    if (ready) cudaEventDestroy(ready);
  };
  if (e != cudaSuccess) { cleanup(); return fail(c, E55_ERR_CUDA, std::string("GPU self-play buffers: ") + cudaGetErrorString(e)); }
  p.n_games = games; p.game0 = 0; p.simulations = simulations; p.use_dirichlet = use_dirichlet; p.reuse_tree = reuse_tree; p.mate_depth = mate_depth;
  p.speculate = !(getenv("E55_GPU_SPECULATE") && atoi(getenv("E55_GPU_SPECULATE")) == 0);   // search on while the host looks for a mate
  p.sampling_moves = sampling_moves; p.max_moves = std::min(max_moves, gs::kMaxPly); p.record_cap = record_cap; p.move_budget = total_moves;
  p.c_puct = 1.25f; p.dirichlet_alpha = 0.15f; p.dirichlet_eps = 0.25f;
  p.cap_enable = c->gpu_cap_enable; p.cap_full = c->gpu_cap_full; p.cap_fast = c->gpu_cap_fast; p.cap_frac = c->gpu_cap_frac;
  // Two cohorts of games on two streams: while the network evaluates one cohort's leaves the search kernels of the other
  // run beside it.  The tower kernel is a persistent kernel with one CTA per SM that leaves no room for anything else
  // on its SMs, so for the run it gets 24 SMs fewer and the (latency-bound, small) search kernels get those
  // (measured at 16384 games: 8.65 k moves/s with one cohort; 8.95 k / 9.18 k / 9.54 k / 9.42 k with 12 / 16 / 24 / 32 spare SMs).
  // Search blocks do not share an SM with a tower CTA even where registers and shared memory would allow it (tried with
  // the tower's shared-memory carve-out on the search kernels: the two cohorts just run in lock step), hence spare SMs.
  const int want_cohorts = getenv("E55_GPU_COHORTS") ? atoi(getenv("E55_GPU_COHORTS")) : 0;   // 0: by size
  const int n_cohorts = want_cohorts == 1 || games < 128 ? 1 : (want_cohorts == 2 || games >= 2048 ? 2 : 1);
  const int spare_sms = getenv("E55_GPU_SPARE_SMS") ? atoi(getenv("E55_GPU_SPARE_SMS")) : 24;
  if (n_cohorts == 2) c->tc.sms = std::max(2, sms_before - std::max(0, spare_sms));
  struct Cohort {
    gs::Params p;
    cudaStream_t st;
    int g0, games, blocks, n;
  } co[2];
  for (int k = 0; k < n_cohorts; ++k) {
    Cohort& q = co[k];
    q.g0 = k == 0 ? 0 : games / 2;
    q.games = n_cohorts == 1 ? games : (k == 0 ? games / 2 : games - games / 2);
    q.blocks = (q.games + 63) / 64;
    q.n = q.games * gs::kLeaf;
    q.st = k == 0 ? c->stream : c->pipe_stream[0];
    const size_t g0 = static_cast<size_t>(q.g0), s0 = g0 * gs::kLeaf;
    q.p = p;
    q.p.n_games = q.games; q.p.game0 = q.g0;
    q.p.games += g0; q.p.nodes += g0 * 2 * gs::kMaxNodes; q.p.edges += g0 * 2 * gs::kMaxEdges; q.p.child += g0 * 2 * gs::kMaxEdges;
    q.p.old_of += g0 * gs::kMaxNodes; q.p.bits += s0 * mshogi::kInputPlanes; q.p.scale += s0 * mshogi::kInputPlanes;
    q.p.move_idx += s0 * gs::kMoveStride; q.p.move_cnt += s0; q.p.priors += s0 * gs::kMoveStride; q.p.value += s0;
    if (q.p.rec_buf) q.p.rec_buf += g0 * gs::kRecordWords;
    q.p.req_board += g0; q.p.req_seq += g0; q.p.ans += g0;
  }
  const int n = static_cast<int>(slots);
  std::vector<unsigned long long> answers(G, 0ull);   // the latest answer of every game (request number 0 = none)
  std::vector<uint32_t> answered(G, 0u);
  cpu_set_t cpus;
  int host_cores = static_cast<int>(std::thread::hardware_concurrency());
  if (sched_getaffinity(0, sizeof(cpus), &cpus) == 0) host_cores = CPU_COUNT(&cpus);
  const int mate_threads = std::max(1, std::min(8, host_cores - 1));
  long mate_searches = 0;
  std::atomic<long> mate_exhausted{0};
  static const long mate_budget = getenv("E55_MATE_BUDGET") ? atol(getenv("E55_MATE_BUDGET")) : 1000000;
  // The mate searches of the requests of games [g0, g0 + count) in snapshot `buf`: selfplay.py:35-36 on the host.
  auto answer_requests = [&](int buf, int g0, int count) {
    const mshogi::Board* boards = h_board + static_cast<size_t>(buf) * G;
    const uint32_t* seq = h_seq + static_cast<size_t>(buf) * G;
    std::vector<int> todo;
    for (int g = g0; g < g0 + count; ++g)
      if (seq[g] != 0 && seq[g] != answered[g]) todo.push_back(g);
    if (todo.empty()) return;
    mate_searches += static_cast<long>(todo.size());
    auto work = [&](size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i) {
        const int g = todo[i];
        mshogi::Board b = boards[g];
        mshogi::Move mv;
        bool exhausted = false;
        const bool found = b.solve_checkmate_dfs(mate_depth, mv, mate_budget, &exhausted);   // bounded: the loop below waits for these
        if (exhausted) mate_exhausted.fetch_add(1, std::memory_order_relaxed);
        answers[g] = (static_cast<unsigned long long>(seq[g]) << 16) | (found ? gs::pack_move(mv) : 0xFFFFu);
        answered[g] = seq[g];
      }
    };
    const int nt = static_cast<int>(std::min<size_t>(mate_threads, (todo.size() + 31) / 32));
    if (nt <= 1) { work(0, todo.size()); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, todo.size() * t / nt, todo.size() * (t + 1) / nt);
    for (auto& t : th) t.join();
  };
  const auto t0 = std::chrono::steady_clock::now();
  int rc = E55_OK;
  e = cudaEventRecord(ready, c->stream);            // the buffers are zeroed on the context stream: the second stream waits for that
  for (int k = 0; k < n_cohorts && e == cudaSuccess; ++k) {
    if (k > 0) e = cudaStreamWaitEvent(co[k].st, ready, 0);
    gs::init_kernel<<<co[k].blocks, 64, 0, co[k].st>>>(co[k].p, seed);
  }
  if (e != cudaSuccess) rc = fail(c, E55_ERR_CUDA, std::string("GPU self-play: ") + cudaGetErrorString(e));
  memset(h_shared, 0, 4 * sizeof(gs::Shared));
  gs::Shared last{};
  bool stalled = false;
  unsigned long long progress_moves = 0;
  auto progress_at = std::chrono::steady_clock::now();
  const bool trace = getenv("E55_SERVICE_TRACE") != nullptr;
  cudaEvent_t te[5] = {};
  double t_sel = 0, t_net = 0, t_pri = 0, t_exp = 0;
  long t_rounds = 0;
  if (trace) for (auto& x : te) cudaEventCreate(&x);
  // E55_GPU_TIMELINE=1: when each kernel of rounds 40..43 started and ended on either stream (no synchronisation in between)
  constexpr long kTl0 = 40, kTlRounds = 4;
  const bool timeline = getenv("E55_GPU_TIMELINE") != nullptr;
  cudaEvent_t tl[kTlRounds][2][5] = {};
  if (timeline) for (auto& a : tl) for (auto& b : a) for (auto& x : b) cudaEventCreate(&x);
  for (long round = 0; rc == E55_OK; ++round) {
    const int buf = static_cast<int>(round & 1);
    for (int k = 0; k < n_cohorts && rc == E55_OK; ++k) {
      Cohort& q = co[k];
      cudaStream_t st = q.st;
      const bool timed = trace && !timeline && k == 0 && round % 16 == 8;
      cudaEvent_t* tle = timeline && round >= kTl0 && round < kTl0 + kTlRounds ? tl[round - kTl0][k] : nullptr;
      if (timed) cudaEventRecord(te[0], st);
      if (tle) cudaEventRecord(tle[0], st);
      gs::select_kernel<<<q.blocks, 64, 0, st>>>(q.p);
      ++c->launches;
      if (timed) cudaEventRecord(te[1], st);
      if (tle) cudaEventRecord(tle[1], st);
      {   // the tower kernel's policy tail soft-maxes each slot's legal moves: the logits never reach HBM
        e55::tc::PolicyTail tail;
        tail.move_off = q.p.move_cnt; tail.move_idx = q.p.move_idx; tail.priors = q.p.priors; tail.move_stride = gs::kMoveStride;
        rc = forward_packed_on(c, st, q.p.bits, q.p.scale, q.n, 0, nullptr, q.p.value, 0, tail);
      }
      if (rc) break;
      if (timed) cudaEventRecord(te[2], st);
      if (tle) cudaEventRecord(tle[2], st);
      if (timed) cudaEventRecord(te[3], st);
      if (tle) cudaEventRecord(tle[3], st);
      gs::expand_kernel<<<q.blocks, 64, 0, st>>>(q.p);
      ++c->launches;
      if (tle) cudaEventRecord(tle[4], st);
      if (timed) {
        cudaEventRecord(te[4], st);
        cudaEventSynchronize(te[4]);
        float ms;
        cudaEventElapsedTime(&ms, te[0], te[1]); t_sel += ms;
        cudaEventElapsedTime(&ms, te[1], te[2]); t_net += ms;
        cudaEventElapsedTime(&ms, te[2], te[3]); t_pri += ms;
        cudaEventElapsedTime(&ms, te[3], te[4]); t_exp += ms;
        ++t_rounds;
      }
      // this round's counters and mate-search requests come down behind it ...
      e = cudaMemcpyAsync(h_shared + 2 * k + buf, p.shared, sizeof(gs::Shared), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess && mate_depth > 0)
        e = cudaMemcpyAsync(h_board + static_cast<size_t>(buf) * G + q.g0, q.p.req_board, q.games * sizeof(mshogi::Board), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess && mate_depth > 0)
        e = cudaMemcpyAsync(h_seq + static_cast<size_t>(buf) * G + q.g0, q.p.req_seq, q.games * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaEventRecord(ev[k][buf], st);
      if (e != cudaSuccess) rc = fail(c, E55_ERR_CUDA, std::string("GPU self-play: ") + cudaGetErrorString(e));
    }
    // ... while the host, the GPU busy with this round, looks at the previous one: stop? which mates?
    for (int k = 0; k < n_cohorts && rc == E55_OK && round > 0; ++k) {
      Cohort& q = co[k];
      // (polled, with a deadline: a round takes tens of milliseconds; minutes mean something is wrong, and the caller
      // should hear about it rather than wait for ever)
      const auto wait_start = std::chrono::steady_clock::now();
      while ((e = cudaEventQuery(ev[k][buf ^ 1])) == cudaErrorNotReady) {
        if (std::chrono::steady_clock::now() - wait_start > std::chrono::seconds(60)) { stalled = true; break; }
        std::this_thread::sleep_for(std::chrono::microseconds(50));
      }
      if (stalled) { cudaGetLastError(); rc = fail(c, E55_ERR_CUDA, "GPU self-play: a round did not finish within 60 s"); break; }
      if (e == cudaSuccess) {
        if (h_shared[2 * k + (buf ^ 1)].moves >= last.moves) last = h_shared[2 * k + (buf ^ 1)];
        if (last.moves != progress_moves) { progress_moves = last.moves; progress_at = std::chrono::steady_clock::now(); }
        else if (std::chrono::steady_clock::now() - progress_at > std::chrono::seconds(30)) {
          fprintf(stderr, "[e55 gpu self-play] stalled after %ld rounds: %llu moves (%llu mate), %llu selections, %llu evaluations, %llu games, "
                  "%ld mate searches answered, %llu overflows\n", round, last.moves, last.mate_moves, last.selections, last.evals, last.games,
                  mate_searches, last.pool_overflows);
          rc = fail(c, E55_ERR_CUDA, "GPU self-play: no move was played for 30 s");
          break;
        }
        if (mate_depth > 0) {
          answer_requests(buf ^ 1, q.g0, q.games);
          memcpy(h_ans + static_cast<size_t>(buf) * G + q.g0, answers.data() + q.g0, q.games * sizeof(unsigned long long));
          e = cudaMemcpyAsync(d_ans + q.g0, h_ans + static_cast<size_t>(buf) * G + q.g0, q.games * sizeof(unsigned long long), cudaMemcpyHostToDevice, q.st);
        }
      }
      if (e != cudaSuccess) rc = fail(c, E55_ERR_CUDA, std::string("GPU self-play: ") + cudaGetErrorString(e));
    }
    if (static_cast<long>(last.moves) >= total_moves) break;
  }
  if (rc == E55_OK) {
    for (int k = 0; k < n_cohorts && e == cudaSuccess; ++k) e = cudaStreamSynchronize(co[k].st);
    if (e == cudaSuccess) e = cudaMemcpy(h_shared, p.shared, sizeof(gs::Shared), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && record_cap > 0) {
      const int nrec = std::min(h_shared->n_records, record_cap);
      *n_records_out = nrec;
      if (nrec > 0) e = cudaMemcpy(records, p.records, static_cast<size_t>(nrec) * gs::kRecordWords * sizeof(uint16_t), cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) rc = fail(c, E55_ERR_CUDA, std::string("GPU self-play: ") + cudaGetErrorString(e));
  } else if (!stalled) {
    for (int k = 0; k < n_cohorts; ++k) cudaStreamSynchronize(co[k].st);
  }
  if (stalled) { c->tc.sms = sms_before; return rc; }   // the device may be hung: freeing the buffers would wait for it, so they are left behind
  out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  out->moves = static_cast<long>(h_shared->moves);
  out->evals = static_cast<long>(h_shared->evals);
  out->games = static_cast<long>(h_shared->games);
  out->mean_game_plies = h_shared->games ? static_cast<double>(h_shared->plies_in_finished_games) / static_cast<double>(h_shared->games) : 0.0;
  out->mean_gpu_batch = static_cast<double>(n) / n_cohorts;   // slots per network launch
  out->tree_overflows = static_cast<long>(h_shared->pool_overflows);
  out->searches_suspended = static_cast<long>(h_shared->searches_suspended);
  out->mate_searches = mate_searches;
  out->mate_exhausted = mate_exhausted.load();
  out->records_truncated = static_cast<long>(h_shared->records_truncated);
  if (timeline && rc == E55_OK) {
    for (long r = 0; r < kTlRounds; ++r)
      for (int k = 0; k < n_cohorts; ++k) {
        float t[5] = {};
        bool ok = true;
        for (int i = 0; i < 5; ++i) ok &= cudaEventElapsedTime(&t[i], tl[0][0][0], tl[r][k][i]) == cudaSuccess;
        if (ok) fprintf(stderr, "[e55 gpu self-play] round %ld cohort %d: select %.2f-%.2f, network -%.2f, priors -%.2f, expand -%.2f ms\n", kTl0 + r, k,
                        t[0], t[1], t[2], t[3], t[4]);
      }
    cudaGetLastError();
  }
  if (timeline) for (auto& a : tl) for (auto& b : a) for (auto& x : b) cudaEventDestroy(x);
  if (trace && t_rounds)
    fprintf(stderr, "[e55 gpu self-play] per round (ms): select %.2f, network %.2f, legal priors %.2f, expand %.2f\n", t_sel / t_rounds,
            t_net / t_rounds, t_pri / t_rounds, t_exp / t_rounds);
  if (trace) for (auto& x : te) cudaEventDestroy(x);
  if (trace)
    fprintf(stderr, "[e55 gpu self-play] %llu moves (%llu by the mate search; %ld mate searches on %d host threads), %llu selections, %llu evaluations, "
            "%llu nodes kept by tree reuse, %llu overflows\n",
            h_shared->moves, h_shared->mate_moves, mate_searches, mate_threads, h_shared->selections, h_shared->evals, h_shared->kept_nodes,
            h_shared->pool_overflows);
  if (rc == E55_OK && h_shared->pool_overflows)
    fprintf(stderr, "[e55 gpu self-play] %llu descents abandoned on a full tree pool (the 90 %% suspension should make this impossible)\n", h_shared->pool_overflows);
  cleanup();
  return rc;
}

}  // extern "C"

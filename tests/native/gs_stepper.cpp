// TEST INFRASTRUCTURE (never part of the product).  The debug stepper of the GPU-resident search (e55_gpu_search_open /
// _select / _expand / _close, csrc/e55_gpu_selfplay.inl) over the HOST build of the same kernels (tests/native/shim):
// same arguments, same outputs, no device.  tests/search_walk.py holds either one to oracle/mcts_oracle.py selection by
// selection; this one runs in the CPU suite, on more and longer searches than the GPU test can afford.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "e55_gpu_selfplay.cuh"

using namespace e55::gs;

struct gs_search {
  Params p{};
  std::vector<void*> owned;
  int finished_before = 0;
};

namespace {
template <class T>
void zalloc(T*& ptr, size_t n, std::vector<void*>& owned) {
  ptr = static_cast<T*>(calloc(n ? n : 1, sizeof(T)));
  owned.push_back(ptr);
}
void one_thread() { blockDim.x = 1; threadIdx.x = 0; blockIdx.x = 0; }
void fill_info(const Game& g, int32_t* info) {
  info[0] = g.root.ply; info[1] = g.sims_done; info[2] = g.n_sel; info[3] = g.n_nodes; info[4] = g.n_edges; info[5] = g.games_finished;
  info[6] = g.rec_words; info[7] = 0;
}
}  // namespace

extern "C" {

int gs_search_open(const uint16_t* prefix, int n_prefix, int simulations, int reuse_tree, int cap_enable, int cap_full, int cap_fast,
                   float cap_frac, gs_search** out) {
  if (!out || n_prefix < 0 || simulations < kLeaf || simulations % kLeaf) return -1;
  gs_search* h = new gs_search();
  Params& p = h->p;
  unsigned long long* ans = nullptr;
  zalloc(p.games, 1, h->owned);
  zalloc(p.nodes, 2 * static_cast<size_t>(kMaxNodes), h->owned);
  zalloc(p.edges, 2 * static_cast<size_t>(kMaxEdges), h->owned);
  zalloc(p.child, 2 * static_cast<size_t>(kMaxEdges), h->owned);
  zalloc(p.old_of, kMaxNodes, h->owned);
  zalloc(p.bits, kLeaf * mshogi::kInputPlanes, h->owned);
  zalloc(p.scale, kLeaf * mshogi::kInputPlanes, h->owned);
  zalloc(p.move_idx, kLeaf * kMoveStride, h->owned);
  zalloc(p.move_cnt, kLeaf, h->owned);
  zalloc(p.priors, kLeaf * kMoveStride, h->owned);
  zalloc(p.value, kLeaf, h->owned);
  zalloc(p.shared, 1, h->owned);
  zalloc(p.rec_buf, kRecordWords, h->owned);
  zalloc(p.records, kRecordWords, h->owned);
  zalloc(p.req_board, 1, h->owned);
  zalloc(p.req_seq, 1, h->owned);
  zalloc(ans, 1, h->owned);
  p.ans = ans;
  // as e55_gpu_search_open sets them
  p.n_games = 1; p.game0 = 0; p.simulations = simulations; p.use_dirichlet = 0; p.reuse_tree = reuse_tree; p.mate_depth = 0;
  p.speculate = 0; p.sampling_moves = 0; p.max_moves = kMaxPly; p.record_cap = 1; p.move_budget = 1 << 30;
  p.c_puct = 1.25f; p.dirichlet_alpha = 0.15f; p.dirichlet_eps = 0.25f;
  p.cap_enable = cap_enable; p.cap_full = cap_full; p.cap_fast = cap_fast; p.cap_frac = cap_frac;
  if (p.cap_enable) p.dirichlet_eps = 0.f;
  one_thread();
  init_kernel(p, 1);
  prefix_kernel(p, prefix, n_prefix);
  *out = h;
  return 0;
}

void gs_search_close(gs_search* h) {
  if (!h) return;
  for (void* q : h->owned) free(q);
  delete h;
}

int gs_search_select(gs_search* h, int32_t* info, int32_t* sel_leaf, int32_t* sel_slot, int32_t* sel_path_len, uint16_t* sel_path,
                     int32_t* slot_cnt, uint16_t* slot_moves, uint16_t* slot_idx) {
  if (!h) return -1;
  one_thread();
  select_kernel(h->p);
  const Game& g = h->p.games[0];
  const GNode* nodes = h->p.nodes + static_cast<size_t>(g.pool) * kMaxNodes;
  const GEdge* edges = h->p.edges + static_cast<size_t>(g.pool) * kMaxEdges;
  memcpy(slot_cnt, h->p.move_cnt, kLeaf * sizeof(int32_t));
  memcpy(slot_idx, h->p.move_idx, kLeaf * kMoveStride * sizeof(uint16_t));
  fill_info(g, info);
  memset(slot_moves, 0, kLeaf * kMoveStride * sizeof(uint16_t));
  for (int s = 0; s < g.n_sel; ++s) {
    sel_leaf[s] = g.leaf[s];
    sel_slot[s] = g.leaf_slot[s];
    int len = 0;
    uint16_t rev[kMaxPath + 1];
    for (int node = g.leaf[s]; node >= 0 && nodes[node].parent >= 0 && len <= kMaxPath; node = nodes[node].parent)
      rev[len++] = edges[nodes[node].parent_edge].mv;
    sel_path_len[s] = len;
    for (int i = 0; i < len; ++i) sel_path[s * kMaxPath + i] = rev[len - 1 - i];
    if (g.leaf_slot[s] >= 0 && g.leaf[s] >= 0) {
      const GNode& leaf = nodes[g.leaf[s]];
      for (int i = 0; i < leaf.n_edges && i < kMoveStride; ++i) slot_moves[g.leaf_slot[s] * kMoveStride + i] = edges[leaf.first_edge + i].mv;
    }
  }
  return 0;
}

int gs_search_expand(gs_search* h, const float* priors, const float* value, int32_t* info, uint16_t* record) {
  if (!h) return -1;
  const int finished_before = h->p.games[0].games_finished;
  memcpy(h->p.priors, priors, kLeaf * kMoveStride * sizeof(float));
  memcpy(h->p.value, value, kLeaf * sizeof(float));
  one_thread();
  expand_kernel(h->p);
  const Game& g = h->p.games[0];
  fill_info(g, info);
  if (g.games_finished != finished_before) {
    memcpy(record, h->p.records, kRecordWords * sizeof(uint16_t));
    info[6] = record[2];
    info[7] = 1;
    *h->p.shared = Shared{};
  } else {
    memcpy(record, h->p.rec_buf, kRecordWords * sizeof(uint16_t));
  }
  return 0;
}

}  // extern "C"

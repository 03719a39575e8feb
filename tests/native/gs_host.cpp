// TEST INFRASTRUCTURE (never part of the product; the product path is the CUDA build of the same header).
//
// The GPU-resident self-play kernels of erweitern_55_b200/csrc/e55_gpu_selfplay.cuh -- init_kernel, select_kernel,
// expand_kernel, the very source the GPU runs -- compiled for the host through tests/native/shim/cuda_runtime.h and
// driven one "thread" (= one game) at a time, the way e55_gpu_selfplay.inl drives them on the device: select ->
// network -> expand per round, mate-search answers a round late.  The network is a pseudo-network (hash of the encoded
// planes -> logits, soft-maxed over the legal moves as the tower kernel's policy tail does; hash -> value), sharp or
// flat on request, so that searches go deep and wide, trees fill, games end every way the rules know.
//
// What it is for (tests/test_gpu_search_on_host.py; compute-sanitizer is closed on the GPU pool):
//   * the kernels run under AddressSanitizer / UndefinedBehaviorSanitizer (build with -fsanitize=address,undefined):
//     every index into the tree pools, path arrays, move lists and record buffers is checked;
//   * after every round the trees are audited (see audit()): no virtual loss left behind, node counts equal to their
//     parent edges' counts, child links consistent, pools within bounds;
//   * finished games are written out in the kernels' record layout for the product's decoder and the rules oracle, and
//     replayed here on the host engine's Position: every move legal, every game ended for the reason the rules give.
//
// usage: gs_host games total_moves simulations mate_depth reuse_tree use_dirichlet max_moves seed sharp cap_enable
//                cap_full cap_fast cap_frac record_cap records_out
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "e55_gpu_selfplay.cuh"

using namespace e55::gs;

namespace {

template <class T>
T* zalloc(size_t n) {
  T* p = static_cast<T*>(calloc(n ? n : 1, sizeof(T)));
  if (!p) { fprintf(stderr, "out of memory\n"); exit(2); }
  return p;
}

template <class K>
void launch(K kernel, int n) {
  blockDim.x = 1;
  threadIdx.x = 0;
  for (int gi = 0; gi < n; ++gi) { blockIdx.x = static_cast<unsigned>(gi); kernel(); }
}

uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// the pseudo-network on one slot: legal-move softmax of hashed logits (what the policy tail leaves in p.priors), value
void pseudo_network(const Params& p, size_t slot, float sharp) {
  const int k = p.move_cnt[slot];
  if (k <= 0) return;
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (int c = 0; c < mshogi::kInputPlanes; ++c) {
    uint32_t s;
    memcpy(&s, &p.scale[slot * mshogi::kInputPlanes + c], 4);
    h = mix(h ^ p.bits[slot * mshogi::kInputPlanes + c] ^ (static_cast<uint64_t>(s) << 32) ^ static_cast<uint64_t>(c) * 0x100000001b3ull);
  }
  const uint16_t* idx = p.move_idx + slot * kMoveStride;
  float* pr = p.priors + slot * kMoveStride;
  float m = -1e30f, sum = 0.f;
  for (int j = 0; j < k; ++j) {
    pr[j] = sharp * (static_cast<float>(mix(h ^ (static_cast<uint64_t>(idx[j]) * 0x9E3779B1ull)) >> 40) / 16777216.f - 0.5f);
    m = pr[j] > m ? pr[j] : m;
  }
  for (int j = 0; j < k; ++j) { pr[j] = std::exp(pr[j] - m); sum += pr[j]; }
  for (int j = 0; j < k; ++j) pr[j] /= sum;
  p.value[slot] = static_cast<float>(mix(h ^ 0x5851f42d4c957f2dull) >> 40) / 8388608.f - 1.f;   // [-1, 1)
}

struct Audit {
  long rounds = 0, bad = 0;
  int max_nodes = 0, max_edges = 0;
  char first[256] = {0};
  void fail(int gi, const char* what, long a, long b) {
    if (!bad) snprintf(first, sizeof(first), "round %ld game %d: %s (%ld, %ld)", rounds, gi, what, a, b);
    ++bad;
  }
};

// The state every game's tree must be in between two rounds.
void audit(const Params& p, Audit& a) {
  ++a.rounds;
  for (int gi = 0; gi < p.n_games; ++gi) {
    const Game& g = p.games[gi];
    if (g.pool != 0 && g.pool != 1) { a.fail(gi, "pool", g.pool, 0); continue; }
    const GNode* nodes = p.nodes + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxNodes;
    const GEdge* edges = p.edges + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxEdges;
    const int32_t* child = p.child + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxEdges;
    if (g.n_nodes < 1 || g.n_nodes > kMaxNodes) { a.fail(gi, "n_nodes", g.n_nodes, kMaxNodes); continue; }
    if (g.n_edges < 0 || g.n_edges > kMaxEdges) { a.fail(gi, "n_edges", g.n_edges, kMaxEdges); continue; }
    a.max_nodes = g.n_nodes > a.max_nodes ? g.n_nodes : a.max_nodes;
    a.max_edges = g.n_edges > a.max_edges ? g.n_edges : a.max_edges;
    if (g.root.ply < 0 || g.root.ply > kMaxPly) a.fail(gi, "ply", g.root.ply, 0);
    if (g.sims_done > g.sim_target + kMaxSel) a.fail(gi, "sims_done", g.sims_done, g.sim_target);
    if (g.rec_words < 0 || g.rec_words > kRecordWords - 4) a.fail(gi, "rec_words", g.rec_words, 0);
    if (nodes[0].parent != -1) a.fail(gi, "root parent", nodes[0].parent, 0);
    long edge_sum = 0;
    for (int i = 0; i < g.n_nodes; ++i) {
      const GNode& nd = nodes[i];
      if (nd.claimed) a.fail(gi, "claimed left set", i, 0);
      if (nd.expanded || nd.n_edges) {
        if (nd.first_edge < 0 || nd.n_edges < 0 || nd.first_edge + nd.n_edges > g.n_edges) { a.fail(gi, "edge range", nd.first_edge, nd.n_edges); continue; }
        edge_sum += nd.n_edges;
        float psum = 0.f;
        for (int e = 0; e < nd.n_edges; ++e) {
          const GEdge& ed = edges[nd.first_edge + e];
          if (ed.vloss != 0) a.fail(gi, "virtual loss left behind", i, ed.vloss);
          if (ed.n < 0 || ed.w < -1e-3f || ed.w > ed.n + 1e-3f) a.fail(gi, "edge statistics", ed.n, static_cast<long>(ed.w));
          if (!(ed.prior >= 0.f && ed.prior <= 1.0001f)) a.fail(gi, "prior", i, e);
          psum += ed.prior;
          const int c = child[nd.first_edge + e];
          if (c < -1 || c >= g.n_nodes) { a.fail(gi, "child index", c, g.n_nodes); continue; }
          if (c >= 0) {
            if (nodes[c].parent != i || nodes[c].parent_edge != nd.first_edge + e) a.fail(gi, "child link", c, i);
            if (nodes[c].n != ed.n) a.fail(gi, "node count != parent edge count", nodes[c].n, ed.n);
          } else if (ed.n != 0) a.fail(gi, "visited edge without a node", i, e);
        }
        if (nd.expanded && nd.n_edges > 0 && (psum < 0.99f || psum > 1.01f)) a.fail(gi, "priors do not sum to one", i, static_cast<long>(psum * 1000));
      }
      if (i > 0 && (nd.parent < 0 || nd.parent >= g.n_nodes)) a.fail(gi, "parent index", nd.parent, i);
    }
    if (edge_sum != g.n_edges) a.fail(gi, "edge pool not the sum of the nodes' edges", edge_sum, g.n_edges);
  }
}

// A finished game on the host engine: every move legal, the end the rules give.  Returns 0 or the first problem.
const char* replay(const uint16_t* rec, int max_moves, bool pruned_targets, long& plies_out, int& reason) {
  const int plies = rec[0], winner = rec[1], words = rec[2];
  mshogi::Position pos;
  int w = 4, ply = 0;
  while (w < 4 + words) {
    const uint16_t mv = rec[w] & 0x7FFF;
    const int sum_n = rec[w + 1], k = rec[w + 3];
    std::vector<mshogi::Move> legal = pos.generate_moves();
    const mshogi::Move m = unpack_move(mv);
    const mshogi::Move* found = nullptr;
    for (const auto& l : legal)
      if (l == m) found = &l;
    if (!found) return "illegal move";
    int total = 0, chosen = 0;
    for (int i = 0; i < k; ++i) {
      const mshogi::Move vm = unpack_move(rec[w + 4 + 2 * i]);
      bool ok = false;
      for (const auto& l : legal) ok = ok || l == vm;
      if (!ok) return "visit list holds an illegal move";
      if (rec[w + 5 + 2 * i] == 0) return "visit count 0 recorded";
      total += rec[w + 5 + 2 * i];
      if ((rec[w + 4 + 2 * i] & 0x7FFF) == mv) chosen = rec[w + 5 + 2 * i];
    }
    if (k > 0 && total != sum_n) return "sum_N is not the sum of the visit counts";
    if (k > 0 && !pruned_targets && chosen == 0) return "the move played has no visits in a target ply";
    pos.do_move(*found);
    w += 4 + 2 * k;
    ++ply;
  }
  if (w != 4 + words) return "record words do not end on a ply";
  plies_out = ply;
  if (rec[3]) { reason = 3; return nullptr; }            // truncated record: nothing more to hold it to
  if (ply != plies) return "ply count";
  bool rep, my_check, op_check;
  pos.is_repetition(rep, my_check, op_check);
  if (rep) {
    reason = 0;
    const int want = my_check ? 1 - pos.side : (op_check ? pos.side : 1);
    return winner == want ? nullptr : "winner of a repetition";
  }
  if (pos.generate_moves().empty()) { reason = 1; return winner == 1 - pos.side ? nullptr : "winner of a game without moves"; }
  reason = 2;
  if (ply < max_moves && ply < kMaxPly) return "game ended early";
  return winner == 2 ? nullptr : "winner of a game that reached the length limit";
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 16) { fprintf(stderr, "usage: see the header of gs_host.cpp\n"); return 2; }
  int a = 1;
  const int G = atoi(argv[a++]);
  const long total_moves = atol(argv[a++]);
  const int simulations = atoi(argv[a++]), mate_depth = atoi(argv[a++]), reuse_tree = atoi(argv[a++]), use_dirichlet = atoi(argv[a++]);
  const int max_moves = atoi(argv[a++]);
  const uint64_t seed = strtoull(argv[a++], nullptr, 10);
  const float sharp = static_cast<float>(atof(argv[a++]));
  const int cap_enable = atoi(argv[a++]), cap_full = atoi(argv[a++]), cap_fast = atoi(argv[a++]);
  const float cap_frac = static_cast<float>(atof(argv[a++]));
  const int record_cap = atoi(argv[a++]);
  const char* records_out = argv[a++];
  if (G < 1 || simulations % kLeaf) { fprintf(stderr, "bad arguments\n"); return 2; }

  Params p{};
  const size_t slots = static_cast<size_t>(G) * kLeaf;
  p.games = zalloc<Game>(G);
  p.nodes = zalloc<GNode>(static_cast<size_t>(G) * 2 * kMaxNodes);
  p.edges = zalloc<GEdge>(static_cast<size_t>(G) * 2 * kMaxEdges);
  p.child = zalloc<int32_t>(static_cast<size_t>(G) * 2 * kMaxEdges);
  p.old_of = zalloc<int32_t>(static_cast<size_t>(G) * kMaxNodes);
  p.bits = zalloc<uint32_t>(slots * mshogi::kInputPlanes);
  p.scale = zalloc<float>(slots * mshogi::kInputPlanes);
  p.move_idx = zalloc<uint16_t>(slots * kMoveStride);
  p.move_cnt = zalloc<int32_t>(slots);
  p.priors = zalloc<float>(slots * kMoveStride);
  p.value = zalloc<float>(slots);
  p.shared = zalloc<Shared>(1);
  p.records = zalloc<uint16_t>(static_cast<size_t>(record_cap) * kRecordWords);
  p.rec_buf = zalloc<uint16_t>(static_cast<size_t>(G) * kRecordWords);
  p.req_board = zalloc<mshogi::Board>(G);
  p.req_seq = zalloc<uint32_t>(G);
  unsigned long long* ans = zalloc<unsigned long long>(G);
  p.ans = ans;
  p.n_games = G; p.game0 = 0; p.speculate = 1; p.simulations = simulations; p.use_dirichlet = use_dirichlet; p.reuse_tree = reuse_tree;
  p.sampling_moves = 10; p.max_moves = max_moves < kMaxPly ? max_moves : kMaxPly; p.record_cap = record_cap; p.mate_depth = mate_depth;
  p.move_budget = total_moves;
  p.c_puct = 1.25f; p.dirichlet_alpha = 0.15f; p.dirichlet_eps = 0.25f;
  p.cap_enable = cap_enable; p.cap_full = cap_full; p.cap_fast = cap_fast; p.cap_frac = cap_frac;

  launch([&] { init_kernel(p, seed); }, G);
  std::vector<uint32_t> answered(G, 0u);
  std::vector<unsigned long long> pending(G, 0ull);      // answers worked out last round: they reach the device this round
  Audit au;
  long mate_searches = 0, mate_found = 0, rounds = 0;
  long evals_by_slots = 0;
  while (static_cast<long>(p.shared->moves) < total_moves) {
    launch([&] { select_kernel(p); }, G);
    for (size_t s = 0; s < slots; ++s) {
      if (p.move_cnt[s] < 0 || p.move_cnt[s] > kMoveStride) { fprintf(stderr, "move count %d in slot %zu\n", p.move_cnt[s], s); return 1; }
      if (p.move_cnt[s] > 0) ++evals_by_slots;
      pseudo_network(p, s, sharp);
    }
    launch([&] { expand_kernel(p); }, G);
    audit(p, au);
    // the host's part (e55_gpu_selfplay.inl): last round's answers go up, this round's requests are worked out
    for (int g = 0; g < G; ++g) ans[g] = pending[g];
    if (mate_depth > 0)
      for (int g = 0; g < G; ++g)
        if (p.req_seq[g] != 0 && p.req_seq[g] != answered[g]) {
          mshogi::Board b = p.req_board[g];
          mshogi::Move mv;
          bool exhausted = false;
          const bool found = b.solve_checkmate_dfs(mate_depth, mv, 200000, &exhausted);
          ++mate_searches;
          mate_found += found;
          pending[g] = (static_cast<unsigned long long>(p.req_seq[g]) << 16) | (found ? pack_move(mv) : 0xFFFFu);
          answered[g] = p.req_seq[g];
        }
    if (++rounds > 4000000) { fprintf(stderr, "no end\n"); return 1; }
  }
  // the finished games
  const int nrec = p.shared->n_records < record_cap ? p.shared->n_records : record_cap;
  long bad_games = 0, plies = 0, reasons[4] = {0, 0, 0, 0};
  char first_bad[128] = {0};
  for (int r = 0; r < nrec; ++r) {
    long pl = 0;
    int reason = -1;
    const char* err = replay(p.records + static_cast<size_t>(r) * kRecordWords, p.max_moves, cap_enable != 0, pl, reason);
    if (err) { if (!bad_games) snprintf(first_bad, sizeof(first_bad), "record %d: %s", r, err); ++bad_games; }
    else { plies += pl; ++reasons[reason]; }
  }
  if (records_out[0] != '-') {
    FILE* f = fopen(records_out, "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", records_out); return 2; }
    fwrite(p.records, sizeof(uint16_t) * kRecordWords, static_cast<size_t>(nrec), f);
    fclose(f);
  }
  const Shared& s = *p.shared;
  printf("{\"rounds\": %ld, \"moves\": %llu, \"games\": %llu, \"evals\": %llu, \"evals_by_slots\": %ld, \"selections\": %llu, \"mate_moves\": %llu, "
         "\"mate_searches\": %ld, \"mate_found\": %ld, \"pool_overflows\": %llu, \"searches_suspended\": %llu, \"records_truncated\": %llu, "
         "\"kept_nodes\": %llu, \"n_records\": %d, \"records_checked\": %d, \"bad_games\": %ld, \"first_bad_game\": \"%s\", "
         "\"ended_by_repetition\": %ld, \"ended_without_moves\": %ld, \"ended_by_length\": %ld, \"truncated_records_seen\": %ld, "
         "\"replayed_plies\": %ld, \"audit_rounds\": %ld, \"audit_failures\": %ld, \"first_audit_failure\": \"%s\", \"max_nodes\": %d, "
         "\"max_edges\": %d, \"node_pool\": %d, \"edge_pool\": %d, \"suspend_nodes\": %d}\n",
         rounds, s.moves, s.games, s.evals, evals_by_slots, s.selections, s.mate_moves, mate_searches, mate_found, s.pool_overflows,
         s.searches_suspended, s.records_truncated, s.kept_nodes, s.n_records, nrec, bad_games, first_bad, reasons[0], reasons[1], reasons[2],
         reasons[3], plies, au.rounds, au.bad, au.first, au.max_nodes, au.max_edges, kMaxNodes, kMaxEdges, kNodesSuspend);
  void* owned[] = {p.games, p.nodes, p.edges, p.child, p.old_of, p.bits, p.scale, p.move_idx, p.move_cnt, p.priors, p.value, p.shared,
                   p.records, p.rec_buf, p.req_board, p.req_seq, ans};
  for (void* q : owned) free(q);
  return au.bad || bad_games ? 1 : 0;
}

// Batched PUCT search tree for minishogi: the native counterpart of minishogilib.MCTS as the reference drives it
// (mcts.py:27-205, selfplay.py:41-89).  minishogilib's implementation is not visible in
// the reference tree, so this is the AlphaZero/KataGo search restated from those call sites:
//   set_root(position, reuse_tree)            mcts.py:50      keep the subtree of the moves played since the last root
//   expanded(node)                            mcts.py:56
//   evaluate(node, position, policy, value)   mcts.py:60,105  terminal test, legal-move masking + softmax -> priors
//   add_noise(root)                           mcts.py:64      Dirichlet noise on the root priors
//   select_leaf(root, position, forced)       mcts.py:97      PUCT descent with virtual loss; moves applied to `position`
//   backpropagate(leaf)                       mcts.py:110     walk the parent links back to the root
//   get_usage / get_playouts / get_nodes / info / best_move / softmax_sample / softmax_sample_among_top_moves /
//   dump(node, target_pruning, remove_zeros)  mcts.py:72,88,114-116,152-187; gamerecord.py:5 "(sum_N, Q, [(m, N)])"
// Values are in [0,1] for the side to move at a node (mcts.py:59,103); an edge is scored for the side that plays it.
// Forced playouts and policy target pruning follow Wu, "Accelerating Self-Play Learning in Go" (k = 2), the source
// of the reference's playout-cap options (selfplay.py:45-61).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "minishogi.hpp"

namespace mshogi {

// 16 bytes per edge: the PUCT scan over a node's children walks these, and a search's working set is mostly edges.
// (The child node index lives in a parallel array: only the chosen edge needs it.)
struct Edge {
  float prior = 0.f;
  float w = 0.f;        // sum of values from the perspective of the side to move at the parent
  int32_t n = 0;
  uint16_t mv = 0;      // the move, packed: from (31 = drop) | to << 5 | piece << 10 | promotes << 14
  uint16_t vloss = 0;
  static uint16_t pack(const Move& m) {
    return static_cast<uint16_t>((m.is_drop() ? 31 : m.from) | (m.to << 5) | (m.piece << 10) | ((m.promotes ? 1 : 0) << 14));
  }
  Move move() const {   // `captured` is not stored: whoever plays the move reads it off the board
    Move m;
    const int from = mv & 31;
    m.from = static_cast<int8_t>(from == 31 ? -1 : from);
    m.to = static_cast<int8_t>((mv >> 5) & 31);
    m.piece = static_cast<int8_t>((mv >> 10) & 15);
    m.promotes = static_cast<int8_t>((mv >> 14) & 1);
    return m;
  }
};
static_assert(sizeof(Edge) == 16, "edge layout");

struct Node {
  int32_t parent = -1;          // node index of the parent (-1: root)
  int32_t parent_edge = -1;     // absolute index (into the edge pool) of the edge leading here
  int32_t first_edge = 0;       // edges live in one pool per tree: [first_edge, first_edge + n_edges)
  int32_t n_edges = 0;
  int32_t n = 0;                // completed visits
  float value = 0.f;            // network / terminal value for the side to move (set by evaluate)
  bool expanded = false;
  bool terminal = false;
  bool claimed = false;         // selected, not yet evaluated, in the current leaf batch (native driver only)
  bool in_check = false;        // known once expanded: the side to move here is in check (= the move into it checks)
};

class Mcts {
 public:
  float c_puct = 1.25f;
  float dirichlet_alpha = 0.15f;   // AlphaZero's value for shogi
  float dirichlet_eps = 0.25f;
  float forced_k = 2.f;
  size_t capacity_bytes = size_t(1) << 28;
  std::vector<Node> nodes;
  std::vector<Edge> edges;      // pool; cleared, not freed, between searches
  std::vector<int32_t> child;   // per edge: node index, -1 = not created yet
  std::mt19937_64 rng{0x55ULL};

  Mcts() { clear(); }
  explicit Mcts(double memory_gb) { set_memory(memory_gb); clear(); }
  void set_memory(double gb) { capacity_bytes = gb > 0 ? static_cast<size_t>(gb * 1073741824.0) : (size_t(1) << 28); }

  void clear() {
    nodes.clear(); edges.clear(); child.clear(); nodes.emplace_back();
    if (nodes.capacity() < 4096) { nodes.reserve(4096); edges.reserve(65536); child.reserve(65536); }   // one 800-simulation search without regrowth
    root_ply_ = -1; root_hash_ = game_hash_ = 0;
  }
  int32_t root() const { return 0; }
  bool valid(int32_t node) const { return node >= 0 && node < static_cast<int32_t>(nodes.size()); }
  Move edge_move(int32_t node, int i) const { return edges[nodes[node].first_edge + i].move(); }
  Edge* edges_of(const Node& nd) { return edges.data() + nd.first_edge; }
  const Edge* edges_of(const Node& nd) const { return edges.data() + nd.first_edge; }
  double usage() const {
    return static_cast<double>(nodes.size() * sizeof(Node) + edges.size() * (sizeof(Edge) + sizeof(int32_t))) / static_cast<double>(capacity_bytes);
  }
  int node_count() const { return static_cast<int>(nodes.size()); }

  // The root for `pos`.  With reuse_tree, when `pos` continues the game the current root belongs to, the subtree
  // reached by the moves played since is kept (compacted to the front of the pools); otherwise the tree is emptied.
  int32_t set_root(const Position& pos, bool reuse_tree) {
    if (reuse_tree && root_ply_ >= 0 && pos.ply >= root_ply_ && pos.history_size() == pos.ply + 1 &&
        pos.entry(0).hash == game_hash_ && pos.entry(root_ply_).hash == root_hash_) {
      int32_t node = 0;
      for (int i = root_ply_ + 1; i <= pos.ply && node >= 0; ++i) {
        const Node& nd = nodes[node];
        int32_t next = -1;
        if (nd.expanded && !nd.terminal)
          for (int e = 0; e < nd.n_edges; ++e)
            if (edges[nd.first_edge + e].mv == Edge::pack(pos.entry(i).move)) { next = child[nd.first_edge + e]; break; }
        node = next;
      }
      if (node >= 0) {
        if (node != 0) compact(node);
        root_ply_ = pos.ply; root_hash_ = pos.hash();
        return 0;
      }
    }
    clear();
    root_ply_ = pos.ply; root_hash_ = pos.hash(); game_hash_ = pos.entry(0).hash;
    return 0;
  }

  // Terminal test in the reference's order (selfplay.py:17-31): perpetual check, repetition, no moves.
  static bool terminal_value(Position& pos, MoveList& moves, float& v) {
    bool rep, my_check, op_check;
    pos.is_repetition(rep, my_check, op_check);
    if (my_check) { v = 0.f; return true; }                        // the perpetual checker (side to move) loses
    if (op_check) { v = 1.f; return true; }
    if (rep) { v = pos.side == WHITE ? 1.f : 0.f; return true; }   // sennichite: the second player wins
    pos.generate_moves(moves);
    if (moves.empty()) { v = 0.f; return true; }
    return false;
  }

  // Walk down from `from` with PUCT, applying moves to `pos` and adding virtual loss along the way.
  int32_t select_leaf(int32_t from, Position& pos, bool forced_playouts = false) {
    int32_t node = from;
    for (;;) {
      const Node nd = nodes[node];        // by value: nodes may grow below
      if (!nd.expanded || nd.terminal) break;
      Edge* ed = edges.data() + nd.first_edge;
      int total = 0;
      for (int i = 0; i < nd.n_edges; ++i) total += ed[i].n + ed[i].vloss;
      const float sq = std::sqrt(static_cast<float>(total) + 1e-8f);
      int best = -1;
      if (forced_playouts && node == from)    // a root child that has been visited gets at least sqrt(k P N) playouts
        for (int i = 0; i < nd.n_edges && best < 0; ++i)
          if (ed[i].n > 0 && static_cast<float>(ed[i].n + ed[i].vloss) < std::sqrt(forced_k * ed[i].prior * static_cast<float>(total)))
            best = i;
      if (best < 0) {
        best = 0;
        float best_score = -1e30f;
        for (int i = 0; i < nd.n_edges; ++i) {
          const Edge& e = ed[i];
          const int n = e.n + e.vloss;
          const float q = n > 0 ? e.w / static_cast<float>(n) : 0.f;   // virtual losses count as visits worth 0
          const float score = q + c_puct * e.prior * sq / (1.f + static_cast<float>(n));
          if (score > best_score) { best_score = score; best = i; }
        }
      }
      Edge& e = ed[best];
      if (e.vloss < 0xFFFF) ++e.vloss;
      int32_t& ch = child[nd.first_edge + best];
      const Move mv = e.move();   // (do_move reads the captured piece off the board)
      // whether the move gives check is known from the child's expansion; only moves into new leaves have to look
      pos.do_move(mv, ch >= 0 && nodes[ch].expanded ? static_cast<int>(nodes[ch].in_check) : -1);
      if (ch < 0) {
        ch = static_cast<int32_t>(nodes.size());
        nodes.emplace_back();
        nodes.back().parent = node;
        nodes.back().parent_edge = nd.first_edge + best;
      }
      node = ch;
    }
    return node;
  }

  // Mark an unevaluated leaf as taken by the current batch; false when another selection already holds it (or it
  // needs no evaluation).
  bool claim(int32_t node) {
    Node& nd = nodes[node];
    if (nd.expanded || nd.terminal || nd.claimed) return false;
    nd.claimed = true;
    return true;
  }

  void set_terminal(int32_t node, float v) {
    Node& nd = nodes[node];
    nd.claimed = false;
    nd.terminal = true;
    nd.value = v;
  }

  // Expand a leaf from the network's 1725 logits: priors = softmax over the logits of the legal moves.
  void expand_from_logits(int32_t node, int side, const MoveList& moves, const float* policy, float value01) {
    Edge* ed = begin_expand(node, moves, value01);
    if (!ed) return;
    float mx = -1e30f;
    for (int i = 0; i < moves.size(); ++i) {
      ed[i].prior = policy[policy_index(moves.m[i], side)];
      mx = std::max(mx, ed[i].prior);
    }
    float sum = 0.f;
    for (int i = 0; i < moves.size(); ++i) { ed[i].prior = std::exp(ed[i].prior - mx); sum += ed[i].prior; }
    const float inv = 1.f / sum;
    for (int i = 0; i < moves.size(); ++i) ed[i].prior *= inv;
  }
  // Expand a leaf from priors the GPU already normalised over `moves` (e55_predict_packed_moves).
  void expand_from_priors(int32_t node, const MoveList& moves, const float* priors, float value01) {
    Edge* ed = begin_expand(node, moves, value01);
    if (!ed) return;
    for (int i = 0; i < moves.size(); ++i) ed[i].prior = priors[i];
  }
  // minishogilib's evaluate(): terminal test + expansion in one call.
  void evaluate(int32_t node, Position& pos, const float* policy, float value01) {
    const Node& nd = nodes[node];
    if (nd.expanded || nd.terminal) return;
    MoveList moves;
    float v;
    if (terminal_value(pos, moves, v)) { set_terminal(node, v); return; }
    expand_from_logits(node, pos.side, moves, policy, value01);
  }

  // The value stored at the leaf is for the side to move there; each edge is scored for the side that chose it.
  void backpropagate(int32_t leaf) {
    int32_t node = leaf;
    float v = nodes[node].value;
    ++nodes[node].n;
    while (nodes[node].parent >= 0) {
      v = 1.f - v;
      Edge& e = edges[nodes[node].parent_edge];
      if (e.vloss > 0) --e.vloss;
      ++e.n;
      e.w += v;
      node = nodes[node].parent;
      ++nodes[node].n;
    }
  }

  void add_noise(int32_t node) {
    Node& nd = nodes[node];
    if (!nd.expanded || nd.n_edges == 0) return;
    std::gamma_distribution<float> gamma(dirichlet_alpha, 1.f);
    std::vector<float> g(nd.n_edges);
    float sum = 0.f;
    for (float& x : g) { x = gamma(rng); sum += x; }
    if (sum <= 0.f) return;
    Edge* ed = edges_of(nd);
    for (int i = 0; i < nd.n_edges; ++i) ed[i].prior = (1.f - dirichlet_eps) * ed[i].prior + dirichlet_eps * g[i] / sum;
  }

  int playouts(int32_t node, bool child_sum) const {
    const Node& nd = nodes[node];
    if (!child_sum) return nd.n;
    int total = 0;
    for (int i = 0; i < nd.n_edges; ++i) total += edges[nd.first_edge + i].n;
    return total;
  }

  int best_edge(int32_t node = 0) const {
    const Node& nd = nodes[node];
    if (nd.n_edges == 0) return -1;
    const Edge* ed = edges_of(nd);
    int best = 0;
    for (int i = 1; i < nd.n_edges; ++i)
      if (ed[i].n > ed[best].n) best = i;
    return best;
  }
  // Sample an edge with probability proportional to n^(1/temperature) (temperature 1: the visit distribution,
  // selfplay.py:65-66).
  int sample_edge(int32_t node, float temperature) {
    const Node& nd = nodes[node];
    if (nd.n_edges == 0) return -1;
    const Edge* ed = edges_of(nd);
    std::vector<double> wgt(nd.n_edges);
    for (int i = 0; i < nd.n_edges; ++i) wgt[i] = ed[i].n;
    return sample_weighted(wgt, temperature, best_edge(node));
  }
  // Sample among the visited children whose win rate is within `away` of the most visited child's (usi.py:126-128).
  int sample_edge_among_top(int32_t node, float away, float temperature) {
    const Node& nd = nodes[node];
    const int best = best_edge(node);
    if (best < 0) return -1;
    const Edge* ed = edges_of(nd);
    const float qbest = ed[best].n > 0 ? ed[best].w / static_cast<float>(ed[best].n) : 0.f;
    std::vector<double> wgt(nd.n_edges, 0.0);
    for (int i = 0; i < nd.n_edges; ++i)
      if (ed[i].n > 0 && ed[i].w / static_cast<float>(ed[i].n) >= qbest - away) wgt[i] = ed[i].n;
    return sample_weighted(wgt, temperature, best);
  }

  // Principal variation (most visited edges) and the win rate of its first move for the side to move at `node`.
  void info(int32_t node, std::vector<Move>& pv, float& q) const {
    pv.clear();
    q = 0.5f;
    bool first = true;
    while (valid(node)) {
      const Node& nd = nodes[node];
      if (!nd.expanded || nd.terminal || nd.n_edges == 0) break;
      const int b = best_edge(node);
      const Edge& e = edges[nd.first_edge + b];
      if (e.n == 0) break;
      if (first) { q = e.w / static_cast<float>(e.n); first = false; }
      pv.push_back(e.move());
      node = child[nd.first_edge + b];
    }
  }

  // Visit counts of the children: (sum_N, Q, [(move, N)]).  With target_pruning, forced playouts are subtracted
  // from every child but the most visited one as long as its PUCT score stays below the best child's, and children
  // left with a single playout are pruned.
  void dump(int32_t node, bool target_pruning, bool remove_zeros, int& sum_n, float& q,
            std::vector<std::pair<Move, int>>& out) const {
    out.clear();
    sum_n = 0; q = 0.f;
    const Node& nd = nodes[node];
    if (!nd.expanded || nd.n_edges == 0) return;
    const Edge* ed = edges_of(nd);
    std::vector<int> cnt(nd.n_edges);
    int total = 0;
    float wsum = 0.f;
    for (int i = 0; i < nd.n_edges; ++i) { cnt[i] = ed[i].n; total += ed[i].n; wsum += ed[i].w; }
    q = total > 0 ? wsum / static_cast<float>(total) : nd.value;
    if (target_pruning && total > 0) {
      const int best = best_edge(node);
      const float sq = std::sqrt(static_cast<float>(total));
      auto puct = [&](int i, int n) {
        const float qi = ed[i].n > 0 ? ed[i].w / static_cast<float>(ed[i].n) : 0.f;   // utility estimate held constant
        return qi + c_puct * ed[i].prior * sq / (1.f + static_cast<float>(n));
      };
      const float best_score = puct(best, cnt[best]);
      for (int i = 0; i < nd.n_edges; ++i) {
        if (i == best || cnt[i] == 0) continue;
        int forced = static_cast<int>(std::sqrt(forced_k * ed[i].prior * static_cast<float>(total)));
        while (forced > 0 && cnt[i] > 0 && puct(i, cnt[i] - 1) < best_score) { --cnt[i]; --forced; }
        if (cnt[i] <= 1) cnt[i] = 0;
      }
    }
    for (int i = 0; i < nd.n_edges; ++i) {
      sum_n += cnt[i];
      if (cnt[i] > 0 || !remove_zeros) out.emplace_back(ed[i].move(), cnt[i]);
    }
  }

  std::string describe(int32_t node) const {   // MCTS.print / MCTS.debug: one line per child, most visited first
    const Node& nd = nodes[node];
    std::vector<int> order(nd.n_edges);
    for (int i = 0; i < nd.n_edges; ++i) order[i] = i;
    const Edge* ed = edges_of(nd);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return ed[a].n > ed[b].n; });
    std::string s = "playouts " + std::to_string(playouts(node, true)) + " value " + std::to_string(nd.value) + "\n";
    for (int i : order) {
      char line[128];
      snprintf(line, sizeof(line), "  %-6s N %6d  Q %.4f  P %.4f\n", ed[i].move().sfen().c_str(), ed[i].n,
               ed[i].n > 0 ? ed[i].w / static_cast<float>(ed[i].n) : 0.f, ed[i].prior);
      s += line;
    }
    return s;
  }

  std::string visualize(int32_t node, int node_num) const {   // the `node_num` most visited nodes below `node` as dot
    std::string s = "digraph game_tree {\n";
    std::vector<int32_t> frontier{node}, chosen;
    auto visits = [&](int32_t nidx) { return nodes[nidx].n; };
    while (!frontier.empty() && static_cast<int>(chosen.size()) < node_num) {
      size_t bi = 0;
      for (size_t i = 1; i < frontier.size(); ++i)
        if (visits(frontier[i]) > visits(frontier[bi])) bi = i;
      const int32_t cur = frontier[bi];
      frontier.erase(frontier.begin() + static_cast<long>(bi));
      chosen.push_back(cur);
      const Node& nd = nodes[cur];
      if (nd.parent >= 0 && cur != node) {
        const Edge& e = edges[nd.parent_edge];
        s += "  n" + std::to_string(nd.parent) + " -> n" + std::to_string(cur) + " [label=\"" + e.move().sfen() + "\"];\n";
      }
      s += "  n" + std::to_string(cur) + " [label=\"N " + std::to_string(nd.n) + "\\nV " + std::to_string(nd.value).substr(0, 5) + "\"];\n";
      if (nd.expanded && !nd.terminal)
        for (int i = 0; i < nd.n_edges; ++i)
          if (child[nd.first_edge + i] >= 0 && edges[nd.first_edge + i].n > 0) frontier.push_back(child[nd.first_edge + i]);
    }
    return s + "}\n";
  }

 private:
  int root_ply_ = -1;
  uint64_t root_hash_ = 0, game_hash_ = 0;

  Edge* begin_expand(int32_t node, const MoveList& moves, float value01) {
    Node& nd = nodes[node];
    nd.claimed = false;
    if (nd.expanded || nd.terminal) return nullptr;
    nd.value = value01;
    nd.first_edge = static_cast<int32_t>(edges.size());
    nd.n_edges = moves.size();
    nd.in_check = moves.in_check;
    nd.expanded = true;
    edges.resize(edges.size() + moves.size());
    child.resize(edges.size(), -1);
    Edge* ed = edges.data() + nd.first_edge;
    for (int i = 0; i < moves.size(); ++i) { ed[i] = Edge(); ed[i].mv = Edge::pack(moves.m[i]); }
    return ed;
  }

  int sample_weighted(std::vector<double>& wgt, float temperature, int fallback) {
    double total = 0.0;
    const double inv_t = temperature > 0.f ? 1.0 / temperature : 1.0;
    for (double& x : wgt) { x = x > 0.0 ? std::pow(x, inv_t) : 0.0; total += x; }
    if (!(total > 0.0)) return fallback;
    double pick = std::uniform_real_distribution<double>(0.0, total)(rng);
    for (size_t i = 0; i < wgt.size(); ++i) {
      pick -= wgt[i];
      if (pick < 0.0 && wgt[i] > 0.0) return static_cast<int>(i);
    }
    return fallback;
  }

  // Move the subtree under `new_root` to the front of fresh pools (breadth first) and drop everything else.
  void compact(int32_t new_root) {
    std::vector<Node> nn;
    std::vector<Edge> ne;
    std::vector<int32_t> nc;
    nn.reserve(std::max<size_t>(4096, nodes.size() / 2));      // room for the next search: no regrowth copies
    ne.reserve(std::max<size_t>(65536, edges.size() / 2));
    nc.reserve(ne.capacity());
    std::vector<int32_t> old_of;           // new index -> old index
    old_of.push_back(new_root);
    nn.push_back(nodes[new_root]);
    nn[0].parent = -1; nn[0].parent_edge = -1;
    for (size_t i = 0; i < nn.size(); ++i) {
      const Node src = nodes[old_of[i]];
      nn[i].claimed = false;
      if (!src.expanded || src.n_edges == 0) { nn[i].first_edge = 0; nn[i].n_edges = 0; continue; }
      const int32_t first = static_cast<int32_t>(ne.size());
      nn[i].first_edge = first;
      for (int e = 0; e < src.n_edges; ++e) {
        Edge ed = edges[src.first_edge + e];
        ed.vloss = 0;
        const int32_t old_child = child[src.first_edge + e];
        int32_t new_child = -1;
        if (old_child >= 0) {
          new_child = static_cast<int32_t>(nn.size());
          old_of.push_back(old_child);
          nn.push_back(nodes[old_child]);
          nn.back().parent = static_cast<int32_t>(i);
          nn.back().parent_edge = first + e;
        }
        ne.push_back(ed);
        nc.push_back(new_child);
      }
    }
    nodes.swap(nn);
    edges.swap(ne);
    child.swap(nc);
  }
};

}  // namespace mshogi

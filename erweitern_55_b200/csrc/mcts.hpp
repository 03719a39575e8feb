// Batched PUCT search for minishogi: the native counterpart of minishogilib.MCTS as mcts.py drives it
// (set_root / select_leaf with virtual loss / evaluate / backpropagate, mcts.py:50-110).
// minishogilib's implementation is not visible in the reference tree, so this is the standard AlphaZero
// search restated from those call sites: priors = softmax of the policy logits over the legal moves
// ("legal-move masking + softmax happen inside evaluate", mcts.py:60,105-106), values in [0,1] for the side
// to move (mcts.py:59,103), leaf_batch leaves selected per network call with virtual loss.
#pragma once
#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

#include "minishogi.hpp"

namespace mshogi {

struct Edge {
  Move move;
  float prior = 0.f;
  float w = 0.f;        // sum of values from the perspective of the side to move at the parent
  int32_t n = 0;
  int32_t vloss = 0;
  int32_t child = -1;   // node index, -1 = not created yet
};

struct Node {
  int32_t first_edge = 0;       // edges live in one pool per tree: [first_edge, first_edge + n_edges)
  int32_t n_edges = 0;
  bool expanded = false;
  bool terminal = false;
  float terminal_value = 0.f;   // for the side to move at this node
  int32_t pending = -1;         // >= 0: already selected in the current batch (index of that selection)
  float value = 0.f;            // network / terminal value for the side to move (set by evaluate)
};

struct Selection {
  std::vector<std::pair<int32_t, int32_t>> path;   // (node, edge index) from the root down
  int32_t leaf = -1;
  int32_t duplicate_of = -1;                       // selection index evaluating the same leaf in this batch
};

class Mcts {
 public:
  float c_puct = 1.25f;
  std::vector<Node> nodes;
  std::vector<Edge> edges;      // pool; cleared, not freed, between searches

  void clear() { nodes.clear(); edges.clear(); nodes.emplace_back(); }
  Edge* edges_of(const Node& nd) { return edges.data() + nd.first_edge; }
  const Edge* edges_of(const Node& nd) const { return edges.data() + nd.first_edge; }
  int32_t root() const { return 0; }

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

  // Walk down from the root with PUCT, applying moves to `pos` and adding virtual loss along the way.
  void select_leaf(Position& pos, Selection& sel, int32_t selection_index) {
    sel.path.clear();
    sel.duplicate_of = -1;
    int32_t node = 0;
    for (;;) {
      const Node nd = nodes[node];        // by value: nodes may grow below
      if (!nd.expanded || nd.terminal) break;
      Edge* ed = edges.data() + nd.first_edge;
      int total = 0;
      for (int i = 0; i < nd.n_edges; ++i) total += ed[i].n + ed[i].vloss;
      const float sq = std::sqrt(static_cast<float>(total) + 1e-8f);
      int best = 0;
      float best_score = -1e30f;
      for (int i = 0; i < nd.n_edges; ++i) {
        const Edge& e = ed[i];
        const int n = e.n + e.vloss;
        const float q = n > 0 ? e.w / static_cast<float>(n) : 0.f;   // virtual losses count as visits worth 0
        const float score = q + c_puct * e.prior * sq / (1.f + static_cast<float>(n));
        if (score > best_score) { best_score = score; best = i; }
      }
      Edge& e = ed[best];
      ++e.vloss;
      sel.path.emplace_back(node, best);
      pos.do_move(e.move);
      if (e.child < 0) {
        e.child = static_cast<int32_t>(nodes.size());
        nodes.emplace_back();
      }
      node = e.child;
    }
    sel.leaf = node;
    Node& leaf = nodes[node];
    if (!leaf.expanded && !leaf.terminal) {
      if (leaf.pending >= 0) sel.duplicate_of = leaf.pending;
      else leaf.pending = selection_index;
    }
  }

  // Expand a leaf: priors = softmax over the logits of the legal moves.  `policy` is the network's 1725 logits.
  void evaluate(int32_t node, Position& pos, const MoveList& moves, const float* policy, float value01) {
    Node& nd = nodes[node];
    nd.pending = -1;
    if (nd.expanded || nd.terminal) return;
    nd.value = value01;
    nd.first_edge = static_cast<int32_t>(edges.size());
    nd.n_edges = moves.size();
    edges.resize(edges.size() + moves.size());
    Edge* ed = edges.data() + nd.first_edge;
    float mx = -1e30f;
    for (int i = 0; i < moves.size(); ++i) {
      ed[i] = Edge();
      ed[i].move = moves.m[i];
      ed[i].prior = policy[policy_index(moves.m[i], pos.side)];
      mx = std::max(mx, ed[i].prior);
    }
    float sum = 0.f;
    for (int i = 0; i < moves.size(); ++i) { ed[i].prior = std::exp(ed[i].prior - mx); sum += ed[i].prior; }
    for (int i = 0; i < moves.size(); ++i) ed[i].prior /= sum;
    nd.expanded = true;
  }
  void set_terminal(int32_t node, float v) {
    Node& nd = nodes[node];
    nd.pending = -1;
    nd.terminal = true;
    nd.terminal_value = nd.value = v;
  }

  // `value01` is for the side to move at the leaf; the edge into a node is scored for the side that chose it.
  void backpropagate(const Selection& sel, float value01) {
    float v = value01;
    for (int i = static_cast<int>(sel.path.size()) - 1; i >= 0; --i) {
      v = 1.f - v;
      Edge& e = edges[nodes[sel.path[i].first].first_edge + sel.path[i].second];
      --e.vloss;
      ++e.n;
      e.w += v;
    }
  }

  int best_edge() const {
    const Edge* ed = edges_of(nodes[0]);
    int best = 0;
    for (int i = 1; i < nodes[0].n_edges; ++i)
      if (ed[i].n > ed[best].n) best = i;
    return best;
  }
  template <class Rng>
  int sample_edge(Rng& rng) const {   // proportional to visit counts (temperature 1, selfplay.py:65-66)
    const Edge* ed = edges_of(nodes[0]);
    long total = 0;
    for (int i = 0; i < nodes[0].n_edges; ++i) total += ed[i].n;
    if (total <= 0) return best_edge();
    long pick = static_cast<long>(std::uniform_int_distribution<long>(0, total - 1)(rng));
    for (int i = 0; i < nodes[0].n_edges; ++i) {
      pick -= ed[i].n;
      if (pick < 0) return i;
    }
    return best_edge();
  }
};

}  // namespace mshogi

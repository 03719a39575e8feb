// C ABI of the native minishogi engine, the search tree and the self-play driver.
// Host-only translation unit (g++): it reaches the network through the public C ABI of include/e55.h only.
#include "../../include/e55.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <new>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "mcts.hpp"

struct e55_pos {
  mshogi::Position p;
};

struct e55_mcts {
  mshogi::Mcts tree;
  // scratch of the native search steps (e55_mcts_search_root / e55_mcts_search_batches)
  std::vector<uint32_t> bits;
  std::vector<float> scale, priors, value;
  std::vector<int32_t> move_off;
  std::vector<uint16_t> move_idx;
  std::vector<int32_t> leaf;
  std::vector<mshogi::Position> leaf_pos;
  std::vector<mshogi::MoveList> leaf_moves;
  std::vector<int> eval_slot;
  explicit e55_mcts(double gb) : tree(gb) {}
};

namespace {

using mshogi::kInputPlanes;
constexpr int kPolicySize = 1725;

int copy_out(const std::string& s, char* buf, int buflen) {
  if (!buf || buflen <= static_cast<int>(s.size())) return E55_ERR_INVALID;
  memcpy(buf, s.c_str(), s.size() + 1);
  return E55_OK;
}

bool find_legal(mshogi::Position& p, const char* move_sfen, mshogi::Move& out) {
  mshogi::Move m;
  if (!move_sfen || !p.parse_move(move_sfen, m)) return false;
  mshogi::MoveList legal;
  p.generate_moves(legal);
  for (const mshogi::Move& l : legal)
    if (l == m) { out = l; return true; }
  return false;
}

bool is_default_net(e55_ctx* ctx) {
  int blocks = 0, filters = 0, in_planes = 0;
  return ctx && e55_get_arch(ctx, &blocks, &filters, &in_planes) == E55_OK && in_planes == kInputPlanes;
}

// ---- one search = what mcts.MCTS.run does for one move (mcts.py:32-138) ---------------------------------------
struct SearchScratch {
  std::vector<uint32_t>* bits;
  std::vector<float>*scale, *priors, *value;
  std::vector<int32_t>* move_off;
  std::vector<uint16_t>* move_idx;
  std::vector<int32_t>* leaf;
  std::vector<mshogi::Position>* leaf_pos;
  std::vector<mshogi::MoveList>* leaf_moves;
  std::vector<int>* eval_slot;
  void size_for(int leaf_batch) {
    if (static_cast<int>(leaf->size()) >= leaf_batch) return;
    bits->resize(static_cast<size_t>(leaf_batch) * kInputPlanes);
    scale->resize(bits->size());
    priors->resize(static_cast<size_t>(leaf_batch) * 256);     // MoveList holds at most 256 moves
    move_idx->resize(priors->size());
    move_off->resize(static_cast<size_t>(leaf_batch) + 1);
    value->resize(leaf_batch);
    leaf->resize(leaf_batch);
    leaf_pos->resize(leaf_batch);
    leaf_moves->resize(leaf_batch);
    eval_slot->resize(leaf_batch);
  }
};

struct Evaluator {
  e55_ctx* ctx;
  bool use_service;
  long evals = 0;
  // priors[j] = softmax over the logits at move_idx[move_off[i] .. move_off[i+1]) of position i, computed on the GPU
  int operator()(const uint32_t* bits, const float* scale, int n, const int32_t* move_off, const uint16_t* move_idx,
                 float* priors, float* value) {
    evals += n;
    return use_service ? e55_service_predict_packed_moves(ctx, bits, scale, n, move_off, move_idx, priors, value)
                       : e55_predict_packed_moves(ctx, bits, scale, n, move_off, move_idx, priors, value);
  }
};

// Step 1 + 2 of MCTS.run: set the root, evaluate it when it is not expanded yet, add noise (mcts.py:49-64).
// Returns E55_OK, or 1 when the root is a terminal position (nothing to search).
int search_root(mshogi::Mcts& tree, SearchScratch& s, Evaluator& nn, const mshogi::Position& root_pos, bool reuse_tree,
                bool use_dirichlet) {
  using namespace mshogi;
  s.size_for(1);
  const int32_t root = tree.set_root(root_pos, reuse_tree);
  const Node& nd = tree.nodes[root];
  if (!nd.expanded && !nd.terminal) {
    Position scratch = root_pos;
    MoveList moves;
    float tv;
    if (Mcts::terminal_value(scratch, moves, tv)) {
      tree.set_terminal(root, tv);
    } else {
      encode_packed(root_pos, s.bits->data(), s.scale->data());
      (*s.move_off)[0] = 0; (*s.move_off)[1] = moves.size();
      for (int i = 0; i < moves.size(); ++i) (*s.move_idx)[i] = static_cast<uint16_t>(policy_index(moves.m[i], root_pos.side));
      const int rc = nn(s.bits->data(), s.scale->data(), 1, s.move_off->data(), s.move_idx->data(), s.priors->data(), s.value->data());
      if (rc != E55_OK) return rc;
      tree.expand_from_priors(root, moves, s.priors->data(), ((*s.value)[0] + 1.f) * 0.5f);
    }
  }
  if (tree.nodes[root].terminal) return 1;
  if (use_dirichlet) tree.add_noise(root);
  return E55_OK;
}

// `n_batches` iterations of MCTS.run's main loop (mcts.py:91-110): select leaf_batch leaves with virtual loss,
// evaluate the new ones in ONE network call, expand, back up.
int search_batches(mshogi::Mcts& tree, SearchScratch& s, Evaluator& nn, const mshogi::Position& root_pos, int n_batches,
                   int leaf_batch, bool forced_playouts) {
  using namespace mshogi;
  s.size_for(leaf_batch);
  std::vector<int32_t>& leaf = *s.leaf;
  std::vector<Position>& leaf_pos = *s.leaf_pos;
  std::vector<MoveList>& leaf_moves = *s.leaf_moves;
  std::vector<int>& eval_slot = *s.eval_slot;
  int32_t* move_off = s.move_off->data();
  uint16_t* move_idx = s.move_idx->data();
  for (int it = 0; it < n_batches; ++it) {
    int n_eval = 0;
    move_off[0] = 0;
    for (int b = 0; b < leaf_batch; ++b) {
      leaf_pos[b] = root_pos;
      leaf[b] = tree.select_leaf(0, leaf_pos[b], forced_playouts);
      eval_slot[b] = -1;
      if (!tree.claim(leaf[b])) continue;                       // expanded, terminal, or a duplicate in this batch
      float v;
      if (Mcts::terminal_value(leaf_pos[b], leaf_moves[b], v)) { tree.set_terminal(leaf[b], v); continue; }
      eval_slot[b] = n_eval;
      encode_packed(leaf_pos[b], s.bits->data() + static_cast<size_t>(n_eval) * kInputPlanes,
                    s.scale->data() + static_cast<size_t>(n_eval) * kInputPlanes);
      const MoveList& lm = leaf_moves[b];                       // the policy entries of the legal moves go along
      uint16_t* idx = move_idx + move_off[n_eval];
      for (int i = 0; i < lm.size(); ++i) idx[i] = static_cast<uint16_t>(policy_index(lm.m[i], leaf_pos[b].side));
      move_off[n_eval + 1] = move_off[n_eval] + lm.size();
      ++n_eval;
    }
    if (n_eval > 0) {
      const int rc = nn(s.bits->data(), s.scale->data(), n_eval, move_off, move_idx, s.priors->data(), s.value->data());
      if (rc != E55_OK) return rc;
    }
    for (int b = 0; b < leaf_batch; ++b)
      if (eval_slot[b] >= 0)
        tree.expand_from_priors(leaf[b], leaf_moves[b], s.priors->data() + move_off[eval_slot[b]],
                                ((*s.value)[eval_slot[b]] + 1.f) * 0.5f);
    for (int b = 0; b < leaf_batch; ++b) tree.backpropagate(leaf[b]);
  }
  return E55_OK;
}

// ---- self-play ---------------------------------------------------------------------------------------------------
// One game the way selfplay.run plays it (selfplay.py:10-113) under client.py's search settings (client.py:27-35):
// repetition / no-move termination, optional mate search, otherwise `simulations` simulations in batches of
// `leaf_batch` through the evaluation service; the first `sampling_moves` plies sample from the visit counts.
struct SelfplayShared {
  e55_ctx* ctx;
  e55_selfplay_config cfg;
  std::atomic<long> moves_left{0}, moves_done{0}, evals_done{0}, games_done{0}, plies_in_finished_games{0};
  std::atomic<int> failures{0};
  std::mutex sink_mu;
};

struct SelfplayWorker {
  SelfplayShared* sh = nullptr;
  uint64_t seed = 0;
  e55_mcts m{0.2};

  void run() {
    using namespace mshogi;
    const e55_selfplay_config& cfg = sh->cfg;
    SearchScratch s{&m.bits, &m.scale, &m.priors, &m.value, &m.move_off, &m.move_idx, &m.leaf, &m.leaf_pos, &m.leaf_moves, &m.eval_slot};
    Evaluator nn{sh->ctx, true};
    m.tree.rng.seed(seed);
    const int n_batches = (cfg.simulations + cfg.leaf_batch - 1) / cfg.leaf_batch;
    std::string record;
    std::vector<std::pair<Move, int>> visits;
    while (sh->moves_left.load() > 0 && sh->failures.load() == 0) {
      Position pos;
      pos.set_start_position();
      m.tree.clear();
      int winner = 2;
      record.clear();
      for (int mv = 0; mv < cfg.max_moves; ++mv) {
        bool rep, my_check, op_check;                                            // selfplay.py:17-26
        pos.is_repetition(rep, my_check, op_check);
        if (my_check) { winner = 1 - pos.side; break; }
        if (op_check) { winner = pos.side; break; }
        if (rep) { winner = 1; break; }
        MoveList legal;
        pos.generate_moves(legal);
        if (legal.empty()) { winner = 1 - pos.side; break; }                    // selfplay.py:28-31
        if (sh->moves_left.fetch_sub(1) <= 0) { sh->evals_done.fetch_add(nn.evals); return; }   // move budget exhausted mid-game
        Move next;
        bool mate = cfg.mate_depth > 0 && pos.solve_checkmate_dfs(cfg.mate_depth, next);   // selfplay.py:35-42
        if (mate) {
          m.tree.clear();
          if (cfg.sink) record += next.sfen() + " 1 1.0 1 " + next.sfen() + " 1\n";     // selfplay.py:82-84
        } else {
          int rc = search_root(m.tree, s, nn, pos, cfg.reuse_tree != 0, cfg.use_dirichlet != 0);
          if (rc == E55_OK) rc = search_batches(m.tree, s, nn, pos, n_batches, cfg.leaf_batch, false);
          if (rc != E55_OK) { sh->failures.fetch_add(1); sh->evals_done.fetch_add(nn.evals); return; }
          const int chosen = pos.ply < cfg.sampling_moves ? m.tree.sample_edge(0, 1.f) : m.tree.best_edge(0);
          next = m.tree.edges_of(m.tree.nodes[0])[chosen].move;
          if (cfg.sink) {                                                       // search.dump(root), selfplay.py:89
            int sum_n;
            float q;
            m.tree.dump(0, false, true, sum_n, q, visits);
            record += next.sfen() + " " + std::to_string(sum_n) + " " + std::to_string(q) + " " + std::to_string(visits.size());
            for (const auto& v : visits) record += " " + v.first.sfen() + " " + std::to_string(v.second);
            record += "\n";
          }
        }
        pos.do_move(next);
        sh->moves_done.fetch_add(1);
      }
      sh->evals_done.fetch_add(nn.evals);
      nn.evals = 0;
      sh->games_done.fetch_add(1);
      sh->plies_in_finished_games.fetch_add(pos.ply);
      if (cfg.sink) {
        record = "winner " + std::to_string(winner) + " ply " + std::to_string(pos.ply) + "\n" + record;
        std::lock_guard<std::mutex> lk(sh->sink_mu);
        cfg.sink(cfg.sink_user, record.c_str());
      }
    }
    sh->evals_done.fetch_add(nn.evals);
  }
};

}  // namespace

extern "C" {

// ---- Position ------------------------------------------------------------------------------------------------
e55_pos* e55_pos_create(void) { return new (std::nothrow) e55_pos(); }
void e55_pos_destroy(e55_pos* p) { delete p; }
e55_pos* e55_pos_copy(const e55_pos* p) { return p ? new (std::nothrow) e55_pos(*p) : nullptr; }

int e55_pos_set_sfen(e55_pos* p, const char* sfen) {
  if (!p || !sfen) return E55_ERR_INVALID;
  return p->p.set_sfen(sfen) ? E55_OK : E55_ERR_INVALID;
}
int e55_pos_sfen(const e55_pos* p, char* buf, int buflen) { return p ? copy_out(p->p.sfen(), buf, buflen) : E55_ERR_INVALID; }
int e55_pos_side_to_move(const e55_pos* p) { return p ? p->p.side : E55_ERR_INVALID; }
int e55_pos_ply(const e55_pos* p) { return p ? p->p.ply : E55_ERR_INVALID; }

int e55_pos_generate_moves(e55_pos* p, char* buf, int buflen) {
  if (!p) return E55_ERR_INVALID;
  std::string s;
  mshogi::MoveList moves;
  p->p.generate_moves(moves);
  for (const mshogi::Move& m : moves) { if (!s.empty()) s += ' '; s += m.sfen(); }
  if (copy_out(s, buf, buflen) != E55_OK) return E55_ERR_INVALID;
  return moves.size();
}

int e55_pos_do_move(e55_pos* p, const char* move_sfen) {
  mshogi::Move m;
  if (!p || !find_legal(p->p, move_sfen, m)) return E55_ERR_INVALID;
  p->p.do_move(m);
  return E55_OK;
}

int e55_pos_is_repetition(const e55_pos* p, int out3[3]) {
  if (!p || !out3) return E55_ERR_INVALID;
  bool a, b, c;
  p->p.is_repetition(a, b, c);
  out3[0] = a; out3[1] = b; out3[2] = c;
  return E55_OK;
}

int e55_pos_solve_checkmate_dfs(e55_pos* p, int depth, char* move_buf, int buflen) {
  if (!p) return E55_ERR_INVALID;
  mshogi::Move m;
  const bool found = p->p.solve_checkmate_dfs(depth, m);
  if (found && copy_out(m.sfen(), move_buf, buflen) != E55_OK) return E55_ERR_INVALID;
  return found ? 1 : 0;
}

int e55_pos_encode_packed(const e55_pos* p, uint32_t* bits, float* scale) {
  if (!p || !bits || !scale) return E55_ERR_INVALID;
  mshogi::encode_packed(p->p, bits, scale);
  return E55_OK;
}

int e55_pos_encode_packed_batch(const e55_pos* const* ps, int n, uint32_t* bits, float* scale) {
  if (!ps || n < 0 || !bits || !scale) return E55_ERR_INVALID;
  for (int i = 0; i < n; ++i) {
    if (!ps[i]) return E55_ERR_INVALID;
    mshogi::encode_packed(ps[i]->p, bits + static_cast<size_t>(i) * kInputPlanes, scale + static_cast<size_t>(i) * kInputPlanes);
  }
  return E55_OK;
}

int e55_pos_policy_index(e55_pos* p, const char* move_sfen) {
  mshogi::Move m;
  if (!p || !find_legal(p->p, move_sfen, m)) return E55_ERR_INVALID;
  return mshogi::policy_index(m, p->p.side);
}

int e55_pos_print(const e55_pos* p, char* buf, int buflen) {
  if (!p) return E55_ERR_INVALID;
  return copy_out(p->p.pretty(), buf, buflen);
}

// ---- MCTS ----------------------------------------------------------------------------------------------------
e55_mcts* e55_mcts_create(double memory_gb) { return new (std::nothrow) e55_mcts(memory_gb); }
void e55_mcts_destroy(e55_mcts* m) { delete m; }

int e55_mcts_clear(e55_mcts* m) {
  if (!m) return E55_ERR_INVALID;
  m->tree.clear();
  return E55_OK;
}

int e55_mcts_configure(e55_mcts* m, float c_puct, float dirichlet_alpha, float dirichlet_eps, uint64_t seed) {
  if (!m || !(c_puct > 0.f) || !(dirichlet_alpha > 0.f) || dirichlet_eps < 0.f || dirichlet_eps > 1.f) return E55_ERR_INVALID;
  m->tree.c_puct = c_puct;
  m->tree.dirichlet_alpha = dirichlet_alpha;
  m->tree.dirichlet_eps = dirichlet_eps;
  m->tree.rng.seed(seed);
  return E55_OK;
}

int e55_mcts_set_root(e55_mcts* m, const e55_pos* pos, int reuse_tree) {
  if (!m || !pos) return E55_ERR_INVALID;
  return m->tree.set_root(pos->p, reuse_tree != 0);
}

int e55_mcts_expanded(const e55_mcts* m, int node) {
  if (!m || !m->tree.valid(node)) return E55_ERR_INVALID;
  const mshogi::Node& nd = m->tree.nodes[node];
  return (nd.expanded || nd.terminal) ? 1 : 0;
}

int e55_mcts_evaluate(e55_mcts* m, int node, e55_pos* pos, const float* policy, float value01) {
  if (!m || !pos || !policy || !m->tree.valid(node)) return E55_ERR_INVALID;
  m->tree.evaluate(node, pos->p, policy, value01);
  return E55_OK;
}

int e55_mcts_add_noise(e55_mcts* m, int node) {
  if (!m || !m->tree.valid(node)) return E55_ERR_INVALID;
  m->tree.add_noise(node);
  return E55_OK;
}

int e55_mcts_select_leaf(e55_mcts* m, int node, e55_pos* pos, int forced_playouts) {
  if (!m || !pos || !m->tree.valid(node)) return E55_ERR_INVALID;
  return m->tree.select_leaf(node, pos->p, forced_playouts != 0);
}

int e55_mcts_backpropagate(e55_mcts* m, int leaf) {
  if (!m || !m->tree.valid(leaf)) return E55_ERR_INVALID;
  m->tree.backpropagate(leaf);
  return E55_OK;
}

double e55_mcts_get_usage(const e55_mcts* m) { return m ? m->tree.usage() : 0.0; }
int e55_mcts_get_nodes(const e55_mcts* m) { return m ? m->tree.node_count() : E55_ERR_INVALID; }

int e55_mcts_get_playouts(const e55_mcts* m, int node, int child_sum) {
  if (!m || !m->tree.valid(node)) return E55_ERR_INVALID;
  return m->tree.playouts(node, child_sum != 0);
}

static int edge_move_out(e55_mcts* m, int node, int edge, char* buf, int buflen) {
  if (edge < 0) return 0;
  const mshogi::Node& nd = m->tree.nodes[node];
  return copy_out(m->tree.edges_of(nd)[edge].move.sfen(), buf, buflen) == E55_OK ? 1 : E55_ERR_INVALID;
}

int e55_mcts_best_move(e55_mcts* m, int node, char* buf, int buflen) {
  if (!m || !m->tree.valid(node)) return E55_ERR_INVALID;
  return edge_move_out(m, node, m->tree.best_edge(node), buf, buflen);
}

int e55_mcts_softmax_sample(e55_mcts* m, int node, float temperature, char* buf, int buflen) {
  if (!m || !m->tree.valid(node)) return E55_ERR_INVALID;
  return edge_move_out(m, node, m->tree.sample_edge(node, temperature), buf, buflen);
}

int e55_mcts_softmax_sample_among_top_moves(e55_mcts* m, int node, float away, float temperature, char* buf, int buflen) {
  if (!m || !m->tree.valid(node)) return E55_ERR_INVALID;
  return edge_move_out(m, node, m->tree.sample_edge_among_top(node, away, temperature), buf, buflen);
}

int e55_mcts_info(const e55_mcts* m, int node, char* pv_buf, int buflen, float* q_out) {
  if (!m || !q_out || !m->tree.valid(node)) return E55_ERR_INVALID;
  std::vector<mshogi::Move> pv;
  m->tree.info(node, pv, *q_out);
  std::string s;
  for (const mshogi::Move& mv : pv) { if (!s.empty()) s += ' '; s += mv.sfen(); }
  if (copy_out(s, pv_buf, buflen) != E55_OK) return E55_ERR_INVALID;
  return static_cast<int>(pv.size());
}

int e55_mcts_dump(const e55_mcts* m, int node, int target_pruning, int remove_zeros, int* sum_n_out, float* q_out,
                  char* moves_buf, int buflen, int* counts, int max_counts) {
  if (!m || !sum_n_out || !q_out || !counts || !m->tree.valid(node)) return E55_ERR_INVALID;
  std::vector<std::pair<mshogi::Move, int>> out;
  m->tree.dump(node, target_pruning != 0, remove_zeros != 0, *sum_n_out, *q_out, out);
  if (static_cast<int>(out.size()) > max_counts) return E55_ERR_INVALID;
  std::string s;
  for (size_t i = 0; i < out.size(); ++i) {
    if (i) s += ' ';
    s += out[i].first.sfen();
    counts[i] = out[i].second;
  }
  if (copy_out(s, moves_buf, buflen) != E55_OK) return E55_ERR_INVALID;
  return static_cast<int>(out.size());
}

int e55_mcts_describe(const e55_mcts* m, int node, char* buf, int buflen) {
  if (!m || !m->tree.valid(node)) return E55_ERR_INVALID;
  return copy_out(m->tree.describe(node), buf, buflen);
}

int e55_mcts_visualize(const e55_mcts* m, int node, int node_num, char* buf, int buflen) {
  if (!m || !m->tree.valid(node) || node_num < 1) return E55_ERR_INVALID;
  return copy_out(m->tree.visualize(node, node_num), buf, buflen);
}

int e55_mcts_search_root(e55_mcts* m, e55_ctx* ctx, const e55_pos* pos, int reuse_tree, int use_dirichlet, int use_service) {
  if (!m || !ctx || !pos || !is_default_net(ctx)) return E55_ERR_INVALID;
  SearchScratch s{&m->bits, &m->scale, &m->priors, &m->value, &m->move_off, &m->move_idx, &m->leaf, &m->leaf_pos, &m->leaf_moves, &m->eval_slot};
  Evaluator nn{ctx, use_service != 0};
  return search_root(m->tree, s, nn, pos->p, reuse_tree != 0, use_dirichlet != 0);
}

int e55_mcts_search_batches(e55_mcts* m, e55_ctx* ctx, const e55_pos* pos, int n_batches, int leaf_batch,
                            int forced_playouts, int use_service) {
  if (!m || !ctx || !pos || n_batches < 1 || leaf_batch < 1 || leaf_batch > 4096 || !is_default_net(ctx)) return E55_ERR_INVALID;
  const mshogi::Node& root = m->tree.nodes[0];
  if (!root.expanded || root.terminal) return E55_ERR_INVALID;   // e55_mcts_search_root first
  SearchScratch s{&m->bits, &m->scale, &m->priors, &m->value, &m->move_off, &m->move_idx, &m->leaf, &m->leaf_pos, &m->leaf_moves, &m->eval_slot};
  Evaluator nn{ctx, use_service != 0};
  const int rc = search_batches(m->tree, s, nn, pos->p, n_batches, leaf_batch, forced_playouts != 0);
  return rc != E55_OK ? rc : n_batches * leaf_batch;
}

// ---- self-play -----------------------------------------------------------------------------------------------
void e55_selfplay_default_config(e55_selfplay_config* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->threads = 1; cfg->total_moves = 1;
  cfg->simulations = 800; cfg->leaf_batch = 16;          // client.py:28-29
  cfg->mate_depth = 7;                                   // selfplay.py:36
  cfg->reuse_tree = 1; cfg->use_dirichlet = 1;           // client.py:31-32
  cfg->sampling_moves = 10; cfg->max_moves = 512;        // selfplay.py:10
  cfg->seed = 1;
}

int e55_selfplay_run_ex(e55_ctx* c, const e55_selfplay_config* cfg, e55_selfplay_stats* out) {
  if (!c || !cfg || !out || cfg->threads < 1 || cfg->total_moves < 1 || cfg->simulations < 1 || cfg->leaf_batch < 1 ||
      cfg->leaf_batch > 256 || cfg->mate_depth < 0 || cfg->max_moves < 1)
    return E55_ERR_INVALID;
  int64_t b0 = 0, p0 = 0, b1 = 0, p1 = 0;
  if (e55_service_stats(c, &b0, &p0) != E55_OK) return E55_ERR_INVALID;   // the evaluation service must be running
  if (!is_default_net(c)) return E55_ERR_INVALID;                         // self-play needs the 266-plane network
  SelfplayShared sh;
  sh.ctx = c; sh.cfg = *cfg;
  sh.moves_left.store(cfg->total_moves);
  std::vector<SelfplayWorker> workers(cfg->threads);
  std::vector<std::thread> th;
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < cfg->threads; ++i) {
    workers[i].sh = &sh;
    workers[i].seed = cfg->seed * 0x9E3779B97F4A7C15ull + static_cast<uint64_t>(i);
    th.emplace_back([&workers, i] { workers[i].run(); });
  }
  for (auto& t : th) t.join();
  out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  e55_service_stats(c, &b1, &p1);
  out->moves = sh.moves_done.load(); out->evals = sh.evals_done.load(); out->games = sh.games_done.load();
  out->mean_game_plies = out->games ? static_cast<double>(sh.plies_in_finished_games.load()) / static_cast<double>(out->games) : 0.0;
  out->mean_gpu_batch = b1 > b0 ? static_cast<double>(p1 - p0) / static_cast<double>(b1 - b0) : 0.0;
  return sh.failures.load() ? E55_ERR_CUDA : E55_OK;
}

int e55_selfplay_run(e55_ctx* c, int threads, long total_moves, int simulations, int leaf_batch, int mate_depth,
                     uint64_t seed, e55_selfplay_stats* out) {
  e55_selfplay_config cfg;
  e55_selfplay_default_config(&cfg);
  cfg.threads = threads; cfg.total_moves = total_moves; cfg.simulations = simulations; cfg.leaf_batch = leaf_batch;
  cfg.mate_depth = mate_depth; cfg.seed = seed;
  return e55_selfplay_run_ex(c, &cfg, out);
}

}  // extern "C"

// C ABI of the native minishogi engine + self-play driver.  Included by e55_api.cu.
#include <random>

#include "mcts.hpp"

struct e55_pos {
  mshogi::Position p;
};

namespace {

int copy_out(const std::string& s, char* buf, int buflen) {
  if (!buf || buflen <= static_cast<int>(s.size())) return E55_ERR_INVALID;
  memcpy(buf, s.c_str(), s.size() + 1);
  return E55_OK;
}

bool find_legal(mshogi::Position& p, const char* move_sfen, mshogi::Move& out) {
  mshogi::Move m;
  if (!move_sfen || !p.parse_move(move_sfen, m)) return false;
  for (const mshogi::Move& l : p.generate_moves())
    if (l == m) { out = l; return true; }
  return false;
}

// One self-play game the way selfplay.run plays it (selfplay.py:10-113): repetition / no-move termination,
// optional mate search, otherwise `simulations` MCTS simulations in batches of `leaf_batch` through the
// evaluation service; the first 10 plies sample from the visit counts, later plies take the most visited move.
struct SelfplayWorker {
  e55_ctx* ctx;
  int simulations, leaf_batch, mate_depth, max_moves;
  std::atomic<long>* moves_left;
  std::atomic<long>* moves_done;
  std::atomic<long>* evals_done;
  std::atomic<long>* games_done;
  std::atomic<long>* plies_in_finished_games;
  std::atomic<int>* failures;
  uint64_t seed;

  std::vector<uint32_t> bits;
  std::vector<float> scale, policy, value;
  std::vector<mshogi::Selection> sels;
  std::vector<mshogi::Position> leaf_pos;
  std::vector<mshogi::MoveList> leaf_moves;
  std::vector<int> eval_slot;
  mshogi::Mcts tree;

  bool nn(int n) {
    if (e55_service_predict_packed(ctx, bits.data(), scale.data(), n, policy.data(), value.data()) != E55_OK) {
      failures->fetch_add(1);
      return false;
    }
    evals_done->fetch_add(n);
    return true;
  }

  bool search(mshogi::Position& root_pos, int& chosen, std::mt19937_64& rng) {
    using namespace mshogi;
    tree.clear();
    MoveList root_moves;
    float tv;
    Position scratch = root_pos;
    if (Mcts::terminal_value(scratch, root_moves, tv)) return false;
    encode_packed(root_pos, bits.data(), scale.data());                       // root evaluation, batch 1 (mcts.py:55-60)
    if (!nn(1)) return false;
    tree.evaluate(0, root_pos, root_moves, policy.data(), (value[0] + 1.f) * 0.5f);
    int done = 0;
    while (done < simulations) {
      int n_eval = 0;
      for (int b = 0; b < leaf_batch; ++b) {                                  // select (virtual loss), mcts.py:91-98
        leaf_pos[b] = root_pos;
        tree.select_leaf(leaf_pos[b], sels[b], b);
        eval_slot[b] = -1;
        Node& leaf = tree.nodes[sels[b].leaf];
        if (leaf.expanded || leaf.terminal || sels[b].duplicate_of >= 0) continue;
        float v;
        if (Mcts::terminal_value(leaf_pos[b], leaf_moves[b], v)) { tree.set_terminal(sels[b].leaf, v); continue; }
        eval_slot[b] = n_eval;
        encode_packed(leaf_pos[b], bits.data() + static_cast<size_t>(n_eval) * kInputPlanes,
                      scale.data() + static_cast<size_t>(n_eval) * kInputPlanes);
        ++n_eval;
      }
      if (n_eval > 0 && !nn(n_eval)) return false;                            // one network call per batch, mcts.py:101-102
      for (int b = 0; b < leaf_batch; ++b)                                    // expand, mcts.py:104-106
        if (eval_slot[b] >= 0)
          tree.evaluate(sels[b].leaf, leaf_pos[b], leaf_moves[b], policy.data() + static_cast<size_t>(eval_slot[b]) * 1725,
                        (value[eval_slot[b]] + 1.f) * 0.5f);
      for (int b = 0; b < leaf_batch; ++b) {                                  // back up, mcts.py:109-110
        const Node& leaf = tree.nodes[sels[b].leaf];
        tree.backpropagate(sels[b], leaf.terminal ? leaf.terminal_value : leaf.value);
      }
      done += leaf_batch;
    }
    chosen = root_pos.ply < 10 ? tree.sample_edge(rng) : tree.best_edge();
    return true;
  }

  void run() {
    using namespace mshogi;
    bits.resize(static_cast<size_t>(leaf_batch) * kInputPlanes);
    scale.resize(bits.size());
    policy.resize(static_cast<size_t>(leaf_batch) * 1725);
    value.resize(leaf_batch);
    sels.resize(leaf_batch);
    leaf_pos.resize(leaf_batch);
    leaf_moves.resize(leaf_batch);
    eval_slot.resize(leaf_batch);
    std::mt19937_64 rng(seed);
    while (moves_left->load() > 0 && failures->load() == 0) {
      Position pos;
      pos.set_start_position();
      for (int mv = 0; mv < max_moves; ++mv) {
        if (moves_left->fetch_sub(1) <= 0) return;
        Move next;
        bool have = false;
        if (mate_depth > 0 && pos.solve_checkmate_dfs(mate_depth, next)) have = true;     // selfplay.py:35-42
        if (!have) {
          int chosen = 0;
          if (!search(pos, chosen, rng)) break;                                           // terminal position
          next = tree.edges_of(tree.nodes[0])[chosen].move;
        }
        pos.do_move(next);
        moves_done->fetch_add(1);
      }
      games_done->fetch_add(1);
      plies_in_finished_games->fetch_add(pos.ply);
    }
  }
};

}  // namespace

extern "C" {

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
  const std::vector<mshogi::Move> moves = p->p.generate_moves();
  for (const mshogi::Move& m : moves) { if (!s.empty()) s += ' '; s += m.sfen(); }
  if (copy_out(s, buf, buflen) != E55_OK) return E55_ERR_INVALID;
  return static_cast<int>(moves.size());
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

int e55_pos_policy_index(e55_pos* p, const char* move_sfen) {
  mshogi::Move m;
  if (!p || !find_legal(p->p, move_sfen, m)) return E55_ERR_INVALID;
  return mshogi::policy_index(m, p->p.side);
}

int e55_selfplay_run(e55_ctx* c, int threads, long total_moves, int simulations, int leaf_batch, int mate_depth,
                     uint64_t seed, e55_selfplay_stats* out) {
  if (!c || !out || threads < 1 || total_moves < 1 || simulations < 1 || leaf_batch < 1 || leaf_batch > 256 || mate_depth < 0)
    return E55_ERR_INVALID;
  if (!c->service) return fail(c, E55_ERR_INVALID, "evaluation service is not running (e55_service_start)");
  if (c->in_planes != mshogi::kInputPlanes) return fail(c, E55_ERR_INVALID, "self-play needs the 266-plane network");
  std::atomic<long> left{total_moves}, moves{0}, evals{0}, games{0}, plies{0};
  std::atomic<int> failures{0};
  int64_t b0 = 0, p0 = 0, b1 = 0, p1 = 0;
  e55_service_stats(c, &b0, &p0);
  std::vector<SelfplayWorker> workers(threads);
  std::vector<std::thread> th;
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < threads; ++i) {
    SelfplayWorker& w = workers[i];
    w.ctx = c; w.simulations = simulations; w.leaf_batch = leaf_batch; w.mate_depth = mate_depth; w.max_moves = 512;
    w.moves_left = &left; w.moves_done = &moves; w.evals_done = &evals; w.games_done = &games;
    w.plies_in_finished_games = &plies; w.failures = &failures;
    w.seed = seed * 0x9E3779B97F4A7C15ull + i;
    th.emplace_back([&w] { w.run(); });
  }
  for (auto& t : th) t.join();
  out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  e55_service_stats(c, &b1, &p1);
  out->moves = moves.load(); out->evals = evals.load(); out->games = games.load();
  out->mean_game_plies = games.load() ? static_cast<double>(plies.load()) / games.load() : 0.0;
  out->mean_gpu_batch = b1 > b0 ? static_cast<double>(p1 - p0) / static_cast<double>(b1 - b0) : 0.0;
  return failures.load() ? E55_ERR_CUDA : E55_OK;
}

}  // extern "C"

// C ABI of the native minishogi engine, the search tree and the self-play driver.
// Host-only translation unit (g++): it reaches the network through the public C ABI of include/e55.h only.
#include "../../include/e55.h"

#include <sched.h>
#include <sys/resource.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
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
// Each step is split into a "prepare" half that leaves a network request in the scratch buffers and a "finish" half
// that consumes the answer, so that a caller can either run them back to back around a blocking predict
// (e55_mcts_search_*) or keep many games in flight from one thread (the self-play worker below).
struct SearchScratch {
  std::vector<uint32_t>* bits;
  std::vector<float>*scale, *priors, *value;
  std::vector<int32_t>* move_off;
  std::vector<uint16_t>* move_idx;
  std::vector<int32_t>* leaf;
  std::vector<mshogi::Position>* leaf_pos;
  std::vector<mshogi::MoveList>* leaf_moves;
  std::vector<int>* eval_slot;
  explicit SearchScratch(e55_mcts& m)
      : bits(&m.bits), scale(&m.scale), priors(&m.priors), value(&m.value), move_off(&m.move_off), move_idx(&m.move_idx),
        leaf(&m.leaf), leaf_pos(&m.leaf_pos), leaf_moves(&m.leaf_moves), eval_slot(&m.eval_slot) {}
  void size_for(int leaf_batch) {
    if (static_cast<int>(leaf->size()) >= leaf_batch) return;
    bits->resize(static_cast<size_t>(leaf_batch) * kInputPlanes);
    scale->resize(bits->size());
    priors->resize(static_cast<size_t>(leaf_batch) * 256);     // MoveList holds at most 256 moves
    move_idx->resize(priors->size());
    move_off->resize(static_cast<size_t>(leaf_batch) + 1);
    value->resize(leaf_batch);
    leaf->resize(leaf_batch);
    leaf_pos->clear();                                           // (forked scratch positions must not be copied around)
    leaf_pos->resize(leaf_batch);
    leaf_moves->resize(leaf_batch);
    eval_slot->resize(leaf_batch);
  }
};

// priors[j] = softmax over the logits at move_idx[move_off[i] .. move_off[i+1]) of position i, computed on the GPU
int evaluate_request(e55_ctx* ctx, bool use_service, SearchScratch& s, int n) {
  return use_service
             ? e55_service_predict_packed_moves(ctx, s.bits->data(), s.scale->data(), n, s.move_off->data(), s.move_idx->data(),
                                                s.priors->data(), s.value->data())
             : e55_predict_packed_moves(ctx, s.bits->data(), s.scale->data(), n, s.move_off->data(), s.move_idx->data(),
                                        s.priors->data(), s.value->data());
}

enum RootStatus { kRootNeedsEval = 0, kRootTerminal = 1, kRootReady = 2 };

// Step 1 of MCTS.run: set the root (mcts.py:49-50).  kRootNeedsEval: a one-position request is in the scratch.
RootStatus prepare_root(mshogi::Mcts& tree, SearchScratch& s, const mshogi::Position& root_pos, bool reuse_tree) {
  using namespace mshogi;
  s.size_for(1);
  const int32_t root = tree.set_root(root_pos, reuse_tree);
  const Node& nd = tree.nodes[root];
  if (nd.terminal) return kRootTerminal;
  if (nd.expanded) return kRootReady;
  Position scratch = root_pos;
  MoveList& moves = (*s.leaf_moves)[0];
  float tv;
  if (Mcts::terminal_value(scratch, moves, tv)) { tree.set_terminal(root, tv); return kRootTerminal; }
  encode_packed(root_pos, s.bits->data(), s.scale->data());
  (*s.move_off)[0] = 0; (*s.move_off)[1] = moves.size();
  for (int i = 0; i < moves.size(); ++i) (*s.move_idx)[i] = static_cast<uint16_t>(policy_index(moves.m[i], root_pos.side));
  return kRootNeedsEval;
}
// Step 2 (mcts.py:55-60): expand the root from the answer.
void finish_root(mshogi::Mcts& tree, SearchScratch& s) {
  tree.expand_from_priors(0, (*s.leaf_moves)[0], s.priors->data(), ((*s.value)[0] + 1.f) * 0.5f);
}

// First half of one iteration of MCTS.run's main loop (mcts.py:91-102): select leaf_batch leaves with virtual loss
// and encode the new ones.  Returns the number of positions in the request (0: nothing to evaluate).
int prepare_batch(mshogi::Mcts& tree, SearchScratch& s, const mshogi::Position& root_pos, int leaf_batch, bool forced_playouts) {
  using namespace mshogi;
  s.size_for(leaf_batch);
  std::vector<int32_t>& leaf = *s.leaf;
  std::vector<Position>& leaf_pos = *s.leaf_pos;
  std::vector<MoveList>& leaf_moves = *s.leaf_moves;
  std::vector<int>& eval_slot = *s.eval_slot;
  int32_t* move_off = s.move_off->data();
  uint16_t* move_idx = s.move_idx->data();
  int n_eval = 0;
  move_off[0] = 0;
  for (int b = 0; b < leaf_batch; ++b) {
    leaf_pos[b].fork_from(root_pos);                          // shares the root's history, owns only the path below it
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
  return n_eval;
}
// Second half (mcts.py:103-110): expand the evaluated leaves, back every selection up.
void finish_batch(mshogi::Mcts& tree, SearchScratch& s, int leaf_batch) {
  const int32_t* move_off = s.move_off->data();
  for (int b = 0; b < leaf_batch; ++b) {
    const int slot = (*s.eval_slot)[b];
    if (slot >= 0)
      tree.expand_from_priors((*s.leaf)[b], (*s.leaf_moves)[b], s.priors->data() + move_off[slot], ((*s.value)[slot] + 1.f) * 0.5f);
  }
  for (int b = 0; b < leaf_batch; ++b) tree.backpropagate((*s.leaf)[b]);
}

// ---- self-play ---------------------------------------------------------------------------------------------------
// Games are played the way selfplay.run plays them (selfplay.py:10-113) under client.py's search settings
// (client.py:27-35): repetition / no-move termination, optional mate search, otherwise `simulations` simulations in
// batches of `leaf_batch`; the first `sampling_moves` plies sample from the visit counts.  The games are small state
// machines over the asynchronous evaluation service, kept in one pool that all host threads work through: a thread
// takes whichever game can advance (its answer has arrived, or it has CPU work to do), so a core never sleeps on a
// network call while any game has work, and only sleeps -- on the oldest outstanding request -- when none has.
struct SelfplayShared {
  e55_ctx* ctx;
  e55_selfplay_config cfg;
  std::vector<struct SelfplayGame> games;      // the pool (threads x games_per_thread)
  std::atomic<uint64_t> submit_seq{0};
  std::atomic<long> moves_left{0}, moves_done{0}, evals_done{0}, games_done{0}, plies_in_finished_games{0};
  std::atomic<int> failures{0};
  std::mutex sink_mu;
};

struct SelfplayGame {
  enum State { kNewMove, kSubmitRoot, kWaitRoot, kSelect, kSubmitBatch, kWaitBatch, kStopped };
  std::atomic<State> state{kNewMove};   // written only by the thread holding `busy`; others read it to skip the game
  mshogi::Position pos;
  e55_mcts m{0.2};
  e55_ticket ticket{};
  uint64_t seq = 0;                  // order of the outstanding request among all games (batches complete in order)
  std::atomic_flag busy = ATOMIC_FLAG_INIT;   // a thread is working on this game
  int batches_left = 0, n_eval = 0, plies = 0, winner = 2;
  bool full_search = true;           // playout-cap oscillation: this move's search is a full one (a learning target)
  int sim_target = 0;                // simulations of this move's search
  std::string record;
  std::vector<std::pair<mshogi::Move, int>> visits;
  SelfplayGame() = default;
  SelfplayGame(SelfplayGame&&) noexcept {}   // only to let std::vector size the pool; games never move afterwards
};

struct SelfplayWorker {
  SelfplayShared* sh = nullptr;
  int index = 0;
  long evals = 0;

  void start_game(SelfplayGame& g) {
    g.pos.set_start_position();
    g.m.tree.clear();
    g.record.clear();
    g.winner = 2; g.plies = 0;
    g.state = SelfplayGame::kNewMove;
  }
  void end_game(SelfplayGame& g) {
    sh->games_done.fetch_add(1);
    sh->plies_in_finished_games.fetch_add(g.pos.ply);
    if (sh->cfg.sink) {
      g.record = "winner " + std::to_string(g.winner) + " ply " + std::to_string(g.pos.ply) + "\n" + g.record;
      std::lock_guard<std::mutex> lk(sh->sink_mu);
      sh->cfg.sink(sh->cfg.sink_user, g.record.c_str());
    }
    if (sh->moves_left.load() > 0 && sh->failures.load() == 0) start_game(g);
    else g.state = SelfplayGame::kStopped;
  }
  bool noise(const SelfplayGame& g) const { return sh->cfg.cap_enable ? g.full_search : sh->cfg.use_dirichlet != 0; }
  void play(SelfplayGame& g, const mshogi::Move& mv) {
    g.pos.do_move(mv);
    ++g.plies;
    sh->moves_done.fetch_add(1);
    g.state = SelfplayGame::kNewMove;
  }
  int submit(SelfplayGame& g, SearchScratch& s, int n, bool may_block) {
    const int rc = e55_service_submit_packed_moves(sh->ctx, s.bits->data(), s.scale->data(), n, s.move_off->data(),
                                                   s.move_idx->data(), may_block ? 1 : 0, &g.ticket);
    if (rc == E55_OK) { evals += n; g.seq = sh->submit_seq.fetch_add(1, std::memory_order_relaxed); }
    else if (rc != E55_PENDING) { sh->failures.fetch_add(1); g.state = SelfplayGame::kStopped; }
    return rc;
  }

  // Advance one game as far as it goes without waiting.  Returns true if anything happened.
  bool step(SelfplayGame& g, bool block) {
    using namespace mshogi;
    const e55_selfplay_config& cfg = sh->cfg;
    SearchScratch s(g.m);
    bool progressed = false;
    for (;;) {
      switch (g.state.load(std::memory_order_relaxed)) {
        case SelfplayGame::kStopped:
          return progressed;
        case SelfplayGame::kNewMove: {
          progressed = true;
          bool rep, my_check, op_check;                                          // selfplay.py:17-26
          g.pos.is_repetition(rep, my_check, op_check);
          if (my_check) { g.winner = 1 - g.pos.side; end_game(g); break; }
          if (op_check) { g.winner = g.pos.side; end_game(g); break; }
          if (rep) { g.winner = 1; end_game(g); break; }
          if (!g.pos.has_legal_move()) { g.winner = 1 - g.pos.side; end_game(g); break; }   // selfplay.py:28-31
          if (g.plies >= cfg.max_moves) { end_game(g); break; }
          if (sh->failures.load() || sh->moves_left.fetch_sub(1) <= 0) { g.state = SelfplayGame::kStopped; break; }   // budget spent mid-game
          Move mate_move;
          if (cfg.mate_depth > 0 && g.pos.solve_checkmate_dfs(cfg.mate_depth, mate_move)) {  // selfplay.py:35-42
            g.m.tree.clear();
            if (cfg.sink) g.record += mate_move.sfen() + " 1 1.0 1 " + mate_move.sfen() + " 1\n";   // selfplay.py:82-84
            play(g, mate_move);
            break;
          }
          // selfplay.py:45-61: under playout-cap oscillation a move is searched fully (N simulations, forced playouts,
          // noise, fresh tree, pruned targets) with probability frac, else fast (n, counting the reused tree's visits)
          g.full_search = true;
          g.sim_target = cfg.simulations;
          bool reuse = cfg.reuse_tree != 0;
          if (cfg.cap_enable) {
            g.full_search = std::uniform_real_distribution<float>(0.f, 1.f)(g.m.tree.rng) < cfg.cap_frac;
            g.sim_target = g.full_search ? cfg.cap_full : cfg.cap_fast;
            reuse = !g.full_search;
          }
          g.batches_left = (g.sim_target + cfg.leaf_batch - 1) / cfg.leaf_batch;
          const RootStatus rs = prepare_root(g.m.tree, s, g.pos, reuse);
          if (rs == kRootNeedsEval) { g.state = SelfplayGame::kSubmitRoot; break; }
          if (noise(g)) g.m.tree.add_noise(0);
          g.state = SelfplayGame::kSelect;
          break;
        }
        case SelfplayGame::kSubmitRoot: {
          const int rc = submit(g, s, 1, block);
          if (rc == E55_PENDING) return progressed;
          if (rc == E55_OK) { g.state = SelfplayGame::kWaitRoot; progressed = true; }
          break;
        }
        case SelfplayGame::kWaitRoot: {
          const int rc = e55_service_fetch(sh->ctx, &g.ticket, s.priors->data(), s.value->data(), block ? 1 : 0);
          if (rc == E55_PENDING) return progressed;
          if (rc != E55_OK) { sh->failures.fetch_add(1); g.state = SelfplayGame::kStopped; break; }
          progressed = true;
          finish_root(g.m.tree, s);
          if (noise(g)) g.m.tree.add_noise(0);
          g.state = SelfplayGame::kSelect;
          break;
        }
        case SelfplayGame::kSelect: {
          progressed = true;
          // "immediate" (mcts.py:84-88, set for fast searches): the visits the reused tree already has count
          if (cfg.cap_enable && !g.full_search && g.m.tree.playouts(0, true) >= g.sim_target) g.batches_left = 0;
          if (g.batches_left == 0) {                                             // selfplay.py:63-68
            const int chosen = g.pos.ply < cfg.sampling_moves ? g.m.tree.sample_edge(0, 1.f) : g.m.tree.best_edge(0);
            const Move next = g.m.tree.edge_move(0, chosen);
            if (cfg.sink) {                                                      // search.dump(root), selfplay.py:89
              int sum_n;
              float q;
              g.m.tree.dump(0, cfg.cap_enable && g.full_search, true, sum_n, q, g.visits);   // target pruning for full searches
              if (cfg.cap_enable && !g.full_search) g.record += "~";                          // selfplay.py:90-91: not a learning target
              g.record += next.sfen() + " " + std::to_string(sum_n) + " " + std::to_string(q) + " " + std::to_string(g.visits.size());
              for (const auto& v : g.visits) g.record += " " + v.first.sfen() + " " + std::to_string(v.second);
              g.record += "\n";
            }
            play(g, next);
            break;
          }
          --g.batches_left;
          g.n_eval = prepare_batch(g.m.tree, s, g.pos, cfg.leaf_batch, cfg.cap_enable && g.full_search);   // forced playouts
          if (g.n_eval == 0) { finish_batch(g.m.tree, s, cfg.leaf_batch); break; }
          g.state = SelfplayGame::kSubmitBatch;
          break;
        }
        case SelfplayGame::kSubmitBatch: {
          const int rc = submit(g, s, g.n_eval, block);
          if (rc == E55_PENDING) return progressed;
          if (rc == E55_OK) { g.state = SelfplayGame::kWaitBatch; progressed = true; }
          break;
        }
        case SelfplayGame::kWaitBatch: {
          const int rc = e55_service_fetch(sh->ctx, &g.ticket, s.priors->data(), s.value->data(), block ? 1 : 0);
          if (rc == E55_PENDING) return progressed;
          if (rc != E55_OK) { sh->failures.fetch_add(1); g.state = SelfplayGame::kStopped; break; }
          progressed = true;
          finish_batch(g.m.tree, s, cfg.leaf_batch);
          g.state = SelfplayGame::kSelect;
          break;
        }
      }
      block = false;   // at most one blocking call per step
      if (g.state == SelfplayGame::kWaitRoot || g.state == SelfplayGame::kWaitBatch) return progressed;   // let the others run
    }
  }

  void run() {
    std::vector<SelfplayGame>& games = sh->games;
    const size_t n_games = games.size();
    // Each thread has a home range of the pool (its games stay warm in its cache); only when none of those can
    // advance does it look through everybody else's for one that can.
    const size_t per = n_games / static_cast<size_t>(sh->cfg.threads);
    const size_t start = per * static_cast<size_t>(index);
    const size_t home_n = index + 1 == sh->cfg.threads ? n_games - start : per;
    for (;;) {
      bool any_alive = false, progressed = false;
      for (size_t i = 0; i < n_games; ++i) {
        if (i == home_n && progressed) break;                           // home games moved: stay home
        SelfplayGame& g = games[(start + i) % n_games];
        if (g.state == SelfplayGame::kStopped) continue;
        any_alive = true;
        if (g.busy.test_and_set(std::memory_order_acquire)) continue;   // another thread has it
        const bool moved = step(g, false);
        g.busy.clear(std::memory_order_release);
        progressed |= moved;
        if (moved && i >= home_n) break;                                // helped out once: back to the home range
      }
      if (!any_alive) break;
      if (progressed) continue;
      // Nothing could move: sleep on the oldest outstanding answer that no other thread is sleeping on (the service
      // always completes it), or -- holding no ticket at all -- on room for a request that did not fit.
      SelfplayGame* oldest = nullptr;
      for (size_t i = 0; i < n_games; ++i) {
        SelfplayGame& g = games[(start + i) % n_games];
        if (g.state != SelfplayGame::kWaitRoot && g.state != SelfplayGame::kWaitBatch && g.state != SelfplayGame::kSubmitRoot &&
            g.state != SelfplayGame::kSubmitBatch) continue;
        if (g.busy.test_and_set(std::memory_order_acquire)) continue;
        const bool waiting = g.state == SelfplayGame::kWaitRoot || g.state == SelfplayGame::kWaitBatch;
        const bool better = !oldest || (waiting && (oldest->state == SelfplayGame::kSubmitRoot || oldest->state == SelfplayGame::kSubmitBatch ||
                                                    g.seq < oldest->seq));
        if (better && (waiting || !oldest)) {
          if (oldest) oldest->busy.clear(std::memory_order_release);
          oldest = &g;
        } else {
          g.busy.clear(std::memory_order_release);
        }
      }
      if (oldest) {
        step(*oldest, true);
        oldest->busy.clear(std::memory_order_release);
      } else {
        std::this_thread::sleep_for(std::chrono::microseconds(30));   // every candidate is in another thread's hands
      }
    }
    sh->evals_done.fetch_add(evals);   // (a game only stops between moves, so no ticket is left unfetched)
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
  return copy_out(m->tree.edge_move(node, edge).sfen(), buf, buflen) == E55_OK ? 1 : E55_ERR_INVALID;
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
  SearchScratch s(*m);
  const RootStatus rs = prepare_root(m->tree, s, pos->p, reuse_tree != 0);
  if (rs == kRootTerminal) return 1;
  if (rs == kRootNeedsEval) {
    const int rc = evaluate_request(ctx, use_service != 0, s, 1);
    if (rc != E55_OK) return rc;
    finish_root(m->tree, s);
  }
  if (use_dirichlet) m->tree.add_noise(0);
  return E55_OK;
}

int e55_mcts_search_batches(e55_mcts* m, e55_ctx* ctx, const e55_pos* pos, int n_batches, int leaf_batch,
                            int forced_playouts, int use_service) {
  if (!m || !ctx || !pos || n_batches < 1 || leaf_batch < 1 || leaf_batch > 4096 || !is_default_net(ctx)) return E55_ERR_INVALID;
  const mshogi::Node& root = m->tree.nodes[0];
  if (!root.expanded || root.terminal) return E55_ERR_INVALID;   // e55_mcts_search_root first
  SearchScratch s(*m);
  for (int it = 0; it < n_batches; ++it) {
    const int n_eval = prepare_batch(m->tree, s, pos->p, leaf_batch, forced_playouts != 0);
    if (n_eval > 0) {
      const int rc = evaluate_request(ctx, use_service != 0, s, n_eval);
      if (rc != E55_OK) return rc;
    }
    finish_batch(m->tree, s, leaf_batch);
  }
  return n_batches * leaf_batch;
}

// ---- self-play -----------------------------------------------------------------------------------------------
void e55_selfplay_default_config(e55_selfplay_config* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->total_moves = 1;
  cfg->simulations = 800; cfg->leaf_batch = 16;          // client.py:28-29
  cfg->mate_depth = 7;                                   // selfplay.py:36
  cfg->reuse_tree = 1; cfg->use_dirichlet = 1;           // client.py:31-32
  cfg->cap_enable = 0; cfg->cap_full = 800; cfg->cap_fast = 128; cfg->cap_frac = 0.25f;   // selfplay.py:10
  cfg->sampling_moves = 10; cfg->max_moves = 512;        // selfplay.py:10
  cfg->threads = 0; cfg->games_per_thread = 0;           // chosen from the core count
  cfg->seed = 1;
}

int e55_selfplay_run_ex(e55_ctx* c, const e55_selfplay_config* cfg_in, e55_selfplay_stats* out) {
  if (!c || !cfg_in || !out) return E55_ERR_INVALID;
  e55_selfplay_config resolved = *cfg_in;
  if (resolved.threads <= 0) {   // leave room for the service's dispatcher and completer threads
    cpu_set_t set;
    int hw = static_cast<int>(std::thread::hardware_concurrency());
    if (sched_getaffinity(0, sizeof(set), &set) == 0) hw = CPU_COUNT(&set);   // the cores this process may actually use
    resolved.threads = hw > 12 ? hw - 2 : (hw > 1 ? hw - 1 : 1);   // the service's two threads need about one core (a spinning completer beats a sleeping one)
  }
  if (resolved.games_per_thread <= 0) resolved.games_per_thread = 12;   // enough to hide the GPU round trip, few enough to stay in cache
  const e55_selfplay_config* cfg = &resolved;
  if (cfg->total_moves < 1 || cfg->simulations < 1 || cfg->leaf_batch < 1 ||
      cfg->leaf_batch > 256 || cfg->mate_depth < 0 || cfg->max_moves < 1 ||
      (cfg->cap_enable && (cfg->cap_full < 1 || cfg->cap_fast < 1 || !(cfg->cap_frac >= 0.f && cfg->cap_frac <= 1.f))))
    return E55_ERR_INVALID;
  int64_t b0 = 0, p0 = 0, b1 = 0, p1 = 0;
  if (e55_service_stats(c, &b0, &p0) != E55_OK) return E55_ERR_INVALID;   // the evaluation service must be running
  if (!is_default_net(c)) return E55_ERR_INVALID;                         // self-play needs the 266-plane network
  SelfplayShared sh;
  sh.ctx = c; sh.cfg = *cfg;
  sh.moves_left.store(cfg->total_moves);
  sh.games = std::vector<SelfplayGame>(static_cast<size_t>(cfg->threads) * static_cast<size_t>(cfg->games_per_thread));
  for (size_t i = 0; i < sh.games.size(); ++i) {
    sh.games[i].m.tree.rng.seed(cfg->seed * 0x9E3779B97F4A7C15ull + i);
    sh.games[i].pos.set_start_position();
  }
  std::vector<SelfplayWorker> workers(cfg->threads);
  std::vector<std::thread> th;
  const auto t0 = std::chrono::steady_clock::now();
  rusage ru0{};
  getrusage(RUSAGE_SELF, &ru0);
  for (int i = 0; i < cfg->threads; ++i) {
    workers[i].sh = &sh;
    workers[i].index = i;
    th.emplace_back([&workers, i] { workers[i].run(); });
  }
  for (auto& t : th) t.join();
  *out = e55_selfplay_stats{};
  out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (getenv("E55_SERVICE_TRACE")) {
    rusage ru1{};
    getrusage(RUSAGE_SELF, &ru1);
    auto secs = [](const timeval& a, const timeval& b) { return static_cast<double>(b.tv_sec - a.tv_sec) + 1e-6 * static_cast<double>(b.tv_usec - a.tv_usec); };
    fprintf(stderr, "[e55 selfplay] %.2f s wall, process CPU: %.2f s user + %.2f s system = %.1f cores busy; %ld voluntary + %ld involuntary context switches\n",
            out->seconds, secs(ru0.ru_utime, ru1.ru_utime), secs(ru0.ru_stime, ru1.ru_stime),
            (secs(ru0.ru_utime, ru1.ru_utime) + secs(ru0.ru_stime, ru1.ru_stime)) / out->seconds, ru1.ru_nvcsw - ru0.ru_nvcsw, ru1.ru_nivcsw - ru0.ru_nivcsw);
  }
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

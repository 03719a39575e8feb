// GPU-resident self-play: the whole loop of selfplay.run / mcts.MCTS.run -- tree descent
// with virtual loss, move generation, repetition test, input encoding, expansion, back-up, move choice -- runs on the
// GPU next to the network, one thread per game, thousands of games at once.  The host only launches kernels: no
// positions, planes, priors or trees cross PCIe, and the number of host cores stops mattering (the host-side driver
// of e55_game.cpp needs ~1.5 ms of a core per move; an 8-GPU box with four cores per GPU starves it).
//
// One round = what one iteration of MCTS.run's main loop does for every game (mcts.py:91-110):
//   select_kernel   per game: selections with virtual loss until kLeaf network slots are filled; each new leaf gets its
//                   terminal test, its legal moves (stored as the node's edges straight away) and its bit-packed
//                   planes in a batch slot
//   tower_kernel    all slots of a cohort of games in one launch (e55_tc.cuh); its policy tail soft-maxes each slot's
//                   legal moves (fixed-stride lists) into the priors: no logits in HBM, no second kernel
//   expand_kernel   per game: priors and values into the new nodes, back-up of every selection; after the last
//                   simulation of a move: choose it (sampled from the visit counts for the first plies), play it,
//                   test for the end of the game, start the next search (or the next game)
// Rules, encoding and policy indexing are the very functions the host engine uses (minishogi.hpp, compiled for the
// device); the search follows csrc/mcts.hpp, tree reuse included (keep_subtree).  Differences to the host driver: the
// root evaluation takes a round of its own, and the mate search before every move is answered by the host
// (request_mate) while the game already searches.  The host side (e55_gpu_selfplay.inl) runs the games as two cohorts
// on two streams, so that one cohort's kernels here overlap the other's tower launch.  Included by e55_api.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "minishogi.hpp"

namespace e55 {
namespace gs {

using namespace mshogi;

constexpr int kLeaf = 16;              // network slots per game per round (Config.batch_size, mcts.py:13)
constexpr int kMaxSel = 24;            // selections per game per round: until the 16 slots are filled, at most this many
constexpr int kMaxPly = 512;           // selfplay.run's max_moves
// (pool and record sizes can be overridden at compile time: the host build of these kernels for the sanitizer runs,
// tests/native/gs_host.cpp, shrinks them so that the suspension, overflow and truncation paths are walked all the time)
#ifndef E55_GS_MAX_NODES
#define E55_GS_MAX_NODES 4096
#endif
#ifndef E55_GS_MAX_EDGES
#define E55_GS_MAX_EDGES (112 * 1024)
#endif
#ifndef E55_GS_RECORD_WORDS
#define E55_GS_RECORD_WORDS 16384
#endif
constexpr int kMaxNodes = E55_GS_MAX_NODES;   // per game and pool: the subtree kept from the last search + one new node per simulation
constexpr int kMaxEdges = E55_GS_MAX_EDGES;   // per game and pool (late positions with pieces in hand have 100+ legal moves)
constexpr int kNodesSuspend = kMaxNodes * 9 / 10 - kMaxSel;    // mcts.py:72: "if the memory is used over 90%, suspend the search"
constexpr int kEdgesSuspend = kMaxEdges * 9 / 10 - kMaxSel * 256;
constexpr int kMoveStride = 256;       // move-list slots per batch position (MoveList capacity: nothing is ever cut)
constexpr int kMaxPath = 96;           // deepest descent below the root
constexpr int kRecordWords = E55_GS_RECORD_WORDS;   // uint16 words of one game record (see finish_game)

struct GNode {
  int32_t parent, parent_edge, first_edge;
  int32_t n;
  float value;
  int16_t n_edges;
  uint8_t expanded, terminal, claimed, in_check;
};

struct alignas(16) GEdge {            // as mshogi::Edge (csrc/mcts.hpp)
  float prior, w;
  int32_t n;
  uint16_t mv, vloss;
};

__host__ __device__ inline uint16_t pack_move(const Move& m) {
  return static_cast<uint16_t>((m.is_drop() ? 31 : m.from) | (m.to << 5) | (m.piece << 10) | ((m.promotes ? 1 : 0) << 14));
}
__host__ __device__ inline Move unpack_move(uint16_t mv) {
  Move m;
  const int from = mv & 31;
  m.from = static_cast<int8_t>(from == 31 ? -1 : from);
  m.to = static_cast<int8_t>((mv >> 5) & 31);
  m.piece = static_cast<int8_t>((mv >> 10) & 15);
  m.promotes = static_cast<int8_t>((mv >> 14) & 1);
  return m;
}

struct Game {
  Board root;                          // the position being searched (root.ply = plies played)
  uint64_t hashes[kMaxPly + 2];        // key of the position after each ply of the game
  uint8_t gave_check[kMaxPly + 2];
  uint8_t rep_count[kMaxPly + 2];
  Snapshot snaps[kHistory];            // the last 8 positions, slot = ply % 8
  uint64_t seen[2][8];                 // 512-bit Bloom filter of the keys of the game, per side to move: "never seen" needs no scan
  uint16_t moves[kMaxPly];             // the game so far
  uint64_t rng;
  int32_t n_nodes, n_edges;
  int32_t pool;                        // which of the game's two tree pools holds the current tree
  int32_t sims_done;                   // simulations of the current search (-1: the root is not evaluated yet)
  int32_t sim_target;                  // simulations this move's search gets (Params::simulations, or N / n under playout-cap oscillation)
  int32_t full_search;                 // playout-cap oscillation: this move's search is a full one (forced playouts, noise, fresh tree, pruned targets)
  int32_t leaf[kMaxSel];               // this round's selections
  int8_t leaf_slot[kMaxSel];           // the network slot a selection is evaluated in (-1: none)
  int32_t n_sel;
  int32_t games_finished;
  int32_t rec_words;                   // words of the game's record written so far
  int32_t rec_truncated;               // the record buffer was too small: a ply or its visit list is missing
  int32_t need_mate;                   // the mate search of the move about to be searched is still to come (from the host)
  uint32_t req_seq;                    // number of the game's latest mate-search request
};

struct Shared {                        // counters of the whole run
  unsigned long long moves, evals, games, plies_in_finished_games, pool_overflows, selections, kept_nodes, mate_moves;
  unsigned long long searches_suspended;   // searches ended early because a tree pool was over 90 % full (mcts.py:72)
  unsigned long long records_truncated;    // finished games whose record did not fit the record buffer
  int n_records;
};

struct Params {
  Game* games;
  GNode* nodes;        // [n_games][2 pools][kMaxNodes]: a kept subtree is compacted from one pool into the other
  GEdge* edges;        // [n_games][2][kMaxEdges]
  int32_t* child;      // [n_games][2][kMaxEdges]
  int32_t* old_of;     // [n_games][kMaxNodes]: scratch of the compaction
  uint32_t* bits;      // [n_games * kLeaf][266]
  float* scale;
  uint16_t* move_idx;  // [n_games * kLeaf][kMoveStride]
  int32_t* move_cnt;   // [n_games * kLeaf]   (0: slot not used this round)
  float* priors;       // [n_games * kLeaf][kMoveStride]
  float* value;        // [n_games * kLeaf]
  Shared* shared;
  uint16_t* records;   // [record_cap][kRecordWords]: finished games, in the layout of finish_game
  uint16_t* rec_buf;   // [n_games][kRecordWords]: the record of the game in progress (null: no records wanted)
  // mate search (selfplay.py:35-42): divergent recursion with a long tail, so the host does it -- a game posts its root
  // here, the driver answers a round or two later (e55_gpu_selfplay.inl); the game searches on meanwhile (speculate)
  Board* req_board;    // [n_games]
  uint32_t* req_seq;   // [n_games]: 0 = no request yet, else the number of the game's latest request
  const unsigned long long* ans;   // [n_games]: (request number << 16) | packed mate move, 0xFFFF = no mate
  int n_games, game0, speculate, simulations, use_dirichlet, reuse_tree, sampling_moves, max_moves, record_cap, mate_depth;   // game0: index of the cohort's first game in the run
  long long move_budget;
  float c_puct, dirichlet_alpha, dirichlet_eps;
  // playout-cap oscillation (selfplay.py:45-61,90-91; off as in client.py unless enabled): a move is searched fully (cap_full
  // simulations, forced playouts, Dirichlet noise, fresh tree, pruned visit counts recorded as the learning target) with
  // probability cap_frac, else fast (cap_fast simulations counting the visits the reused tree already has, no noise; no target)
  int cap_enable, cap_full, cap_fast;
  float cap_frac;
};

// ---- small device helpers ---------------------------------------------------------------------------------------
__device__ inline uint64_t next_u64(uint64_t& s) {   // xorshift64*
  s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
  return s * 0x2545F4914F6CDD1Dull;
}
__device__ inline float next_unit(uint64_t& s) { return (static_cast<float>(next_u64(s) >> 40) + 0.5f) * (1.f / 16777216.f); }
__device__ inline float next_normal(uint64_t& s) {
  const float u1 = next_unit(s), u2 = next_unit(s);
  return sqrtf(-2.f * logf(u1)) * cosf(6.28318530718f * u2);
}
__device__ inline float next_gamma(uint64_t& s, float alpha) {   // Marsaglia-Tsang; alpha < 1 through gamma(alpha + 1)
  const float boost = alpha < 1.f ? powf(next_unit(s), 1.f / alpha) : 1.f;
  const float a = alpha < 1.f ? alpha + 1.f : alpha;
  const float d = a - 1.f / 3.f, c = 1.f / sqrtf(9.f * d);
  for (int it = 0; it < 64; ++it) {
    const float x = next_normal(s), v0 = 1.f + c * x;
    if (v0 <= 0.f) continue;
    const float v = v0 * v0 * v0, u = next_unit(s);
    if (logf(u) < 0.5f * x * x + d - d * v + d * logf(v)) return d * v * boost;
  }
  return d * boost;
}

__device__ inline void take_snapshot(const Board& b, Snapshot& sn, int rep) {
  for (int c = 0; c < 2; ++c)
    for (int t = 1; t < PIECE_TYPES; ++t) sn.pieces[c][t - 1] = b.bb[c][t];
  for (int c = 0; c < 2; ++c)
    for (int k = 0; k < kHandTypes; ++k) sn.hand[c][k] = b.hand[c][k];
  sn.rep_count = static_cast<uint8_t>(rep > 255 ? 255 : rep);
}

// The history a descent sees: the game's plies [0, ply0] in global memory, then the path below the root in local
// arrays.  Index = ply.
struct PathView {
  const Game* g;
  int ply0, depth;                   // depth = plies below the root
  uint64_t hash[kMaxPath + 1];       // [1..depth]
  uint8_t check[kMaxPath + 1];
  uint8_t rep[kMaxPath + 1];
  Snapshot snap[kHistory];           // snapshots of the path positions, slot = ply % 8 (only the last 8 matter)
  __device__ uint64_t key(int ply) const { return ply <= ply0 ? g->hashes[ply] : hash[ply - ply0]; }
  __device__ bool gave_check(int ply) const { return ply <= ply0 ? g->gave_check[ply] != 0 : check[ply - ply0] != 0; }
  __device__ int rep_count(int ply) const { return ply <= ply0 ? g->rep_count[ply] : rep[ply - ply0]; }
  __device__ const Snapshot& snapshot(int ply) const { return ply <= ply0 ? g->snaps[ply % kHistory] : snap[ply % kHistory]; }
  __device__ int last() const { return ply0 + depth; }
};

__device__ inline bool maybe_seen(const Game& g, int parity, uint64_t key) { return (g.seen[parity][(key >> 6) & 7] >> (key & 63)) & 1ull; }
__device__ inline void mark_seen(Game& g, int parity, uint64_t key) { g.seen[parity][(key >> 6) & 7] |= 1ull << (key & 63); }

// Position::push_history for a descent: key, check flag, repetition count, snapshot of the position after `m`.
__device__ inline void path_push(PathView& pv, const Board& after, uint64_t key, bool check) {
  const int d = ++pv.depth, ply = pv.ply0 + d;
  pv.hash[d] = key;
  pv.check[d] = check ? 1 : 0;
  int c = 1;
  // earlier occurrences with the same side to move: first along the path (local), then -- only if the game's filter
  // does not rule it out -- through the game's history
  int i = ply - 2;
  for (; i > pv.ply0; i -= 2)
    if (pv.hash[i - pv.ply0] == key) { c += pv.rep[i - pv.ply0]; break; }
  if (c == 1 && i <= pv.ply0 && maybe_seen(*pv.g, ply & 1, key))
    for (; i >= 0; i -= 2)
      if (pv.g->hashes[i] == key) { c += pv.g->rep_count[i]; break; }
  pv.rep[d] = static_cast<uint8_t>(c > 255 ? 255 : c);
  take_snapshot(after, pv.snap[ply % kHistory], c);
}

// Position::is_repetition on a PathView (minishogi.hpp): fourth occurrence = repetition; a side all of whose moves
// since the first occurrence gave check is the perpetual checker.
__device__ inline void path_repetition(const PathView& pv, bool& rep, bool& my_check, bool& op_check) {
  rep = my_check = op_check = false;
  const int last = pv.last();
  if (pv.rep_count(last) < 4) return;
  const uint64_t h = pv.key(last);
  int first = last;
  for (int i = last; i >= 0; i -= 2)
    if (pv.key(i) == h) first = i;
  rep = true;
  bool mine = true, theirs = true, any_mine = false, any_theirs = false;
  for (int i = first + 1; i <= last; ++i) {
    const bool mover_is_me = ((last - i) & 1) == 1;
    if (mover_is_me) { any_mine = true; mine = mine && pv.gave_check(i); }
    else { any_theirs = true; theirs = theirs && pv.gave_check(i); }
  }
  my_check = any_mine && mine;
  op_check = any_theirs && theirs;
}

// encode_packed (minishogi.hpp) on a PathView.
__device__ inline void path_encode(const PathView& pv, const Board& pos, uint32_t* bits, float* scale) {
  for (int c = 0; c < kInputPlanes; ++c) { bits[c] = 0u; scale[c] = 1.f; }
  const int me = pos.side, last = pv.last();
  for (int t = 0; t < kHistory; ++t) {
    const int idx = last - t;
    if (idx < 0) break;
    const Snapshot& sn = pv.snapshot(idx);
    uint32_t* b = bits + t * kPlanesPerStep;
    float* sc = scale + t * kPlanesPerStep;
    if (me == BLACK) {
      for (int k = 0; k < 10; ++k) { b[k] = sn.pieces[BLACK][k]; b[10 + k] = sn.pieces[WHITE][k]; }
    } else {
      for (int k = 0; k < 10; ++k) { b[k] = rotate180(sn.pieces[WHITE][k]); b[10 + k] = rotate180(sn.pieces[BLACK][k]); }
    }
    const int rc = sn.rep_count >= 3 ? 3 : sn.rep_count;
    if (rc >= 1) b[20 + rc - 1] = kAllSquares;
    for (int k = 0; k < kHandTypes; ++k) {
      if (sn.hand[me][k]) { b[23 + k] = kAllSquares; sc[23 + k] = sn.hand[me][k] / 2.f; }
      if (sn.hand[me ^ 1][k]) { b[28 + k] = kAllSquares; sc[28 + k] = sn.hand[me ^ 1][k] / 2.f; }
    }
  }
  if (me == WHITE) bits[264] = kAllSquares;
  if (pos.ply > 0) { bits[265] = kAllSquares; scale[265] = static_cast<float>(pos.ply > 512 ? 512 : pos.ply) / 512.f; }
}

// ---- game bookkeeping ---------------------------------------------------------------------------------------------
// What kind of search the move now on the board gets (selfplay.py:45-61).
__device__ inline void begin_move(Game& g, const Params& p) {
  g.full_search = 1;
  g.sim_target = p.simulations;
  if (p.cap_enable) {
    g.full_search = next_unit(g.rng) < p.cap_frac ? 1 : 0;
    g.sim_target = g.full_search ? p.cap_full : p.cap_fast;
  }
}
__device__ inline bool noise_on(const Game& g, const Params& p) { return p.cap_enable ? g.full_search != 0 : p.use_dirichlet != 0; }

__device__ inline void reset_tree(Game& g, GNode* nodes) {
  g.n_nodes = 1; g.n_edges = 0; g.sims_done = -1;
  GNode& r = nodes[0];
  r.parent = -1; r.parent_edge = -1; r.first_edge = 0; r.n = 0; r.value = 0.f; r.n_edges = 0;
  r.expanded = r.terminal = r.claimed = r.in_check = 0;
}

__device__ inline void start_game(Game& g, GNode* nodes) {
  Board& b = g.root;
  b.clear_board();
  // rbsgk/4p/5/P4/KGSBR b - 1
  const PieceType top[5] = {ROOK, BISHOP, SILVER, GOLD, KING}, bottom[5] = {KING, GOLD, SILVER, BISHOP, ROOK};
  for (int c = 0; c < 5; ++c) { b.put(sq_of(0, c), top[c], WHITE); b.put(sq_of(4, c), bottom[c], BLACK); }
  b.put(sq_of(1, 4), PAWN, WHITE);
  b.put(sq_of(3, 0), PAWN, BLACK);
  g.hashes[0] = b.compute_hash();
  for (int k = 0; k < 8; ++k) g.seen[0][k] = g.seen[1][k] = 0ull;
  mark_seen(g, 0, g.hashes[0]);
  g.gave_check[0] = 0;
  g.rep_count[0] = 1;
  take_snapshot(b, g.snaps[0], 1);
  reset_tree(g, nodes);
}

// Position::do_move on the game itself (after a move was chosen).
__device__ inline void game_play(Game& g, const Move& mv) {
  Board& b = g.root;
  Move m = mv;
  if (!m.is_drop()) m.captured = static_cast<int8_t>(Board::type_of(b.board[m.to]));
  const uint64_t key = b.hash_after(g.hashes[b.ply], m);
  g.moves[b.ply] = pack_move(m);
  b.apply(m);
  ++b.ply;
  const int ply = b.ply;
  g.hashes[ply] = key;
  g.gave_check[ply] = b.in_check(b.side) ? 1 : 0;
  int c = 1;
  for (int i = ply - 2; i >= 0; i -= 2)
    if (g.hashes[i] == key) { c += g.rep_count[i]; break; }
  g.rep_count[ply] = static_cast<uint8_t>(c > 255 ? 255 : c);
  mark_seen(g, ply & 1, key);
  take_snapshot(b, g.snaps[ply % kHistory], c);
}

// Mcts::add_noise on the root (mcts.py:63-64).
__device__ inline void root_noise(Game& g, const GNode& root, GEdge* edges, float* scratch, float alpha, float eps) {
  float sum = 0.f;
  for (int i = 0; i < root.n_edges; ++i) { scratch[i] = next_gamma(g.rng, alpha); sum += scratch[i]; }
  if (sum > 0.f)
    for (int i = 0; i < root.n_edges; ++i)
      edges[root.first_edge + i].prior = (1.f - eps) * edges[root.first_edge + i].prior + eps * scratch[i] / sum;
}

// Mcts::compact: the subtree under `new_root` moves, breadth first, to the front of the game's other pool.
__device__ inline void keep_subtree(Game& g, const Params& p, int gi, int new_root) {
  const size_t from = static_cast<size_t>(gi) * 2 + g.pool, to = static_cast<size_t>(gi) * 2 + (g.pool ^ 1);
  const GNode* on = p.nodes + from * kMaxNodes; const GEdge* oe = p.edges + from * kMaxEdges; const int32_t* oc = p.child + from * kMaxEdges;
  GNode* nn = p.nodes + to * kMaxNodes; GEdge* ne = p.edges + to * kMaxEdges; int32_t* nc = p.child + to * kMaxEdges;
  int32_t* old_of = p.old_of + static_cast<size_t>(gi) * kMaxNodes;
  int n_nodes = 1, n_edges = 0;
  old_of[0] = new_root;
  nn[0] = on[new_root];
  nn[0].parent = -1; nn[0].parent_edge = -1;
  for (int i = 0; i < n_nodes; ++i) {
    const GNode src = on[old_of[i]];
    nn[i].claimed = 0;
    if (!src.expanded || src.n_edges == 0) { nn[i].first_edge = 0; nn[i].n_edges = 0; continue; }
    const int first = n_edges;
    nn[i].first_edge = first;
    for (int e = 0; e < src.n_edges; ++e) {
      GEdge ed = oe[src.first_edge + e];
      ed.vloss = 0;
      const int old_child = oc[src.first_edge + e];
      int new_child = -1;
      if (old_child >= 0) {
        new_child = n_nodes++;
        old_of[new_child] = old_child;
        nn[new_child] = on[old_child];
        nn[new_child].parent = i;
        nn[new_child].parent_edge = first + e;
      }
      ne[first + e] = ed;
      nc[first + e] = new_child;
    }
    n_edges += src.n_edges;
  }
  g.pool ^= 1;
  g.n_nodes = n_nodes;
  g.n_edges = n_edges;
}

// One ply of the game record, as selfplay.run keeps it (selfplay.py:80-92; gamerecord.py:5): the move, then
// search.dump(root) = (sum_N, Q, [(move, N), ...]) -- four words (move, sum_N, Q * 65535, k) and k (move, N) pairs.
// A mate-search move is (1, 1.0, [(move, 1)]).  A record that does not fit its buffer (16 384 words: ~350 plies at 20
// visited moves each) first loses visit lists (k = 0), then plies; either way the finished record carries a flag
// (header word 3), its reader reports it as incomplete and keeps the plies without visits out of the learning targets.
__device__ inline void record_ply(Game& g, const Params& p, int gi, uint16_t mv, const GNode* root, const GEdge* edges) {
  if (!p.rec_buf) return;
  constexpr int kCap = kRecordWords - 4;            // finish_game puts a four-word header in front
  uint16_t* rec = p.rec_buf + static_cast<size_t>(gi) * kRecordWords;
  int w = g.rec_words;
  const bool prune = root && p.cap_enable && g.full_search;     // Mcts::dump with target_pruning (Wu 2019; mcts.py:187)
  const bool target = !root || !p.cap_enable || g.full_search;  // selfplay.py:90-91
  int k = 0, sum_n = 0, total = 0, best = 0;
  float wsum = 0.f, sq = 0.f, best_score = 0.f;
  if (root) {
    for (int i = 0; i < root->n_edges; ++i) {
      const GEdge& e = edges[root->first_edge + i];
      total += e.n; wsum += e.w;
      if (e.n > edges[root->first_edge + best].n) best = i;
    }
    sq = sqrtf(static_cast<float>(total));
  }
  // the count a child is recorded with: forced playouts are taken back from every child but the most visited one while
  // its PUCT score, with its utility held, stays below the best child's; children left with one playout are dropped
  auto puct = [&](const GEdge& e, int n) {
    const float qi = e.n > 0 ? __fdiv_rn(e.w, static_cast<float>(e.n)) : 0.f;
    return __fadd_rn(qi, __fdiv_rn(__fmul_rn(__fmul_rn(p.c_puct, e.prior), sq), __fadd_rn(1.f, static_cast<float>(n))));
  };
  if (prune && total > 0) best_score = puct(edges[root->first_edge + best], edges[root->first_edge + best].n);
  auto count_of = [&](int i) {
    const GEdge& e = edges[root->first_edge + i];
    int c = e.n;
    if (prune && total > 0 && i != best && c > 0) {
      int forced = static_cast<int>(sqrtf(__fmul_rn(__fmul_rn(2.f, e.prior), static_cast<float>(total))));
      while (forced > 0 && c > 0 && puct(e, c - 1) < best_score) { --c; --forced; }
      if (c <= 1) c = 0;
    }
    return c;
  };
  if (root)
    for (int i = 0; i < root->n_edges; ++i) { const int c = count_of(i); sum_n += c; k += c > 0; }
  else
    k = 1;                                          // a mate-search move: (1, 1.0, [(move, 1)])
  if (w + 4 > kCap) { g.rec_truncated = 1; return; }   // not even the move fits: the record ends here (flagged)
  const bool fits = w + 4 + 2 * k <= kCap;
  if (!fits) g.rec_truncated = 1;                   // the move is kept, its visit list is not (k = 0)
  rec[w++] = static_cast<uint16_t>(mv | (target ? 0u : 0x8000u));   // bit 15: not a learning target
  rec[w++] = static_cast<uint16_t>(root ? (sum_n > 65535 ? 65535 : sum_n) : 1);
  rec[w++] = static_cast<uint16_t>(root ? (total > 0 ? wsum / static_cast<float>(total) : root->value) * 65535.f + 0.5f : 65535.f);
  rec[w++] = static_cast<uint16_t>(fits ? k : 0);
  if (fits) {
    if (!root) { rec[w++] = mv; rec[w++] = 1; }
    else
      for (int i = 0; i < root->n_edges; ++i) {
        const int c = count_of(i);
        if (c > 0) { rec[w++] = edges[root->first_edge + i].mv; rec[w++] = static_cast<uint16_t>(c > 65535 ? 65535 : c); }
      }
  }
  g.rec_words = w;
}

// A move has been chosen: play it and look at the game (selfplay.py:17-31,78).  True if the game is over.
__device__ inline bool play_and_judge(Game& g, const Params& p, const Move& mv, int& winner) {
  atomicAdd(&p.shared->moves, 1ull);
  game_play(g, mv);
  const int ply = g.root.ply;
  if (g.rep_count[ply] >= 4) {
    PathView pv;
    pv.g = &g; pv.ply0 = ply; pv.depth = 0;
    bool rep, my_check, op_check;
    path_repetition(pv, rep, my_check, op_check);
    winner = my_check ? 1 - g.root.side : (op_check ? g.root.side : 1);
    return true;
  }
  if (!g.root.has_legal_move()) { winner = 1 - g.root.side; return true; }
  return ply >= p.max_moves || ply >= kMaxPly;
}

__device__ inline void finish_game(Game& g, const Params& p, int gi, GNode* nodes, int winner) {
  atomicAdd(&p.shared->games, 1ull);
  atomicAdd(&p.shared->plies_in_finished_games, static_cast<unsigned long long>(g.root.ply));
  if (p.records && p.rec_buf) {        // the finished game: [plies, winner, words, 0] + its plies (record_ply)
    const int r = atomicAdd(&p.shared->n_records, 1);
    if (r < p.record_cap) {
      uint16_t* out = p.records + static_cast<size_t>(r) * kRecordWords;
      const uint16_t* rec = p.rec_buf + static_cast<size_t>(gi) * kRecordWords;
      out[0] = static_cast<uint16_t>(g.root.ply);
      out[1] = static_cast<uint16_t>(winner);
      const int words = g.rec_words < kRecordWords - 4 ? g.rec_words : kRecordWords - 4;
      out[2] = static_cast<uint16_t>(words);
      out[3] = static_cast<uint16_t>(g.rec_truncated ? 1 : 0);
      for (int i = 0; i < words; ++i) out[4 + i] = rec[i];
    }
    if (g.rec_truncated) atomicAdd(&p.shared->records_truncated, 1ull);
  }
  g.rec_words = 0;
  g.rec_truncated = 0;
  ++g.games_finished;
  start_game(g, nodes);
  begin_move(g, p);
}

// Ask the host for the mate search of the position now on the board.
__device__ inline void request_mate(Game& g, const Params& p, int gi) {
  if (p.mate_depth <= 0) { g.need_mate = 0; return; }
  g.need_mate = 1;
  ++g.req_seq;
  p.req_board[gi] = g.root;
  __threadfence();
  p.req_seq[gi] = g.req_seq;
}

// ---- kernels ------------------------------------------------------------------------------------------------------
__global__ void init_kernel(Params p, uint64_t seed) {
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= p.n_games) return;
  Game& g = p.games[gi];
  g.rng = (seed + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(p.game0 + gi + 1)) | 1ull;
  for (int i = 0; i < 8; ++i) next_u64(g.rng);
  g.games_finished = 0;
  g.pool = 0;
  g.req_seq = 0;
  g.rec_words = 0;
  g.rec_truncated = 0;
  start_game(g, p.nodes + static_cast<size_t>(gi) * 2 * kMaxNodes);
  begin_move(g, p);
  request_mate(g, p, gi);
}

// Debug stepper (e55_gpu_search_*): the moves of `prefix` are played on every game before its first search.
__global__ void prefix_kernel(Params p, const uint16_t* prefix, int n_prefix) {
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= p.n_games) return;
  Game& g = p.games[gi];
  for (int i = 0; i < n_prefix; ++i) game_play(g, unpack_move(prefix[i]));
  reset_tree(g, p.nodes + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxNodes);
  begin_move(g, p);
  request_mate(g, p, gi);
}

// Terminal test of a leaf in the reference's order (selfplay.py:17-31; Mcts::terminal_value).
__device__ inline bool leaf_terminal(Board& pos, const PathView& pv, MoveList& moves, float& v) {
  bool rep, my_check, op_check;
  path_repetition(pv, rep, my_check, op_check);
  if (my_check) { v = 0.f; return true; }
  if (op_check) { v = 1.f; return true; }
  if (rep) { v = pos.side == WHITE ? 1.f : 0.f; return true; }
  pos.generate_moves(moves);
  if (moves.empty()) { v = 0.f; return true; }
  return false;
}

__global__ void __launch_bounds__(64) select_kernel(Params p) {
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= p.n_games) return;
  Game& g = p.games[gi];
  GNode* nodes = p.nodes + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxNodes;
  GEdge* edges = p.edges + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxEdges;
  int32_t* child = p.child + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxEdges;
  for (int k = 0; k < kLeaf; ++k) p.move_cnt[static_cast<size_t>(gi) * kLeaf + k] = 0;
  g.n_sel = 0;
  // selfplay.py:35-42: a mate within mate_depth plies is played at once, without a search (and the tree is dropped).
  // The answer to the game's request comes from the host a round or two later.  Until then the game searches on
  // regardless (p.speculate; a found mate drops that work, which is rare) -- only the end of a search waits for it.
  const unsigned long long answer = g.need_mate ? p.ans[gi] : 0ull;
  if (g.need_mate && static_cast<uint32_t>(answer >> 16) == g.req_seq) {
    const uint16_t mate = static_cast<uint16_t>(answer & 0xFFFFull);
    g.need_mate = 0;
    if (mate != 0xFFFF) {
      atomicAdd(&p.shared->mate_moves, 1ull);
      int winner = 2;
      record_ply(g, p, gi, mate, nullptr, nullptr);
      if (play_and_judge(g, p, unpack_move(mate), winner)) finish_game(g, p, gi, nodes, winner);
      else { reset_tree(g, nodes); begin_move(g, p); }
      request_mate(g, p, gi);          // the next move starts with a mate search too
    }
  }
  if (g.need_mate && (!p.speculate || g.sims_done >= g.sim_target || (g.sims_done >= 0 && nodes[0].terminal))) return;
  if (g.sims_done >= g.sim_target) return;   // the answer was all that was missing: expand_kernel ends the search
  // mcts.py:71-73: "If the memory is used over 90%, suspend the search" -- the tree pools of a game are its memory.  The
  // search ends with the simulations done so far (expand_kernel chooses the move from them); nothing ever overflows.
  if (g.sims_done >= 0 && (g.n_nodes > kNodesSuspend || g.n_edges > kEdgesSuspend)) {
    atomicAdd(&p.shared->searches_suspended, 1ull);
    g.sims_done = g.sim_target;
    g.n_sel = 0;
    return;
  }
  PathView pv;
  MoveList moves;
  // A round selects until the game's 16 network slots are filled (selections that end in an expanded, terminal or
  // already claimed node need none), at most kMaxSel times; an unevaluated root takes a round of its own.
  const int max_sel = nodes[0].expanded ? kMaxSel : 1;
  const bool forced = p.cap_enable && g.full_search;
  int used = 0, s = 0;
  for (; s < max_sel && used < kLeaf; ++s) {
    const size_t slot = static_cast<size_t>(gi) * kLeaf + used;
    g.leaf_slot[s] = -1;
    Board pos = g.root;
    pv.g = &g; pv.ply0 = pos.ply; pv.depth = 0;
    int node = 0;
    // PUCT descent with virtual loss (Mcts::select_leaf)
    while (nodes[node].expanded && !nodes[node].terminal && pv.depth < kMaxPath) {
      const GNode nd = nodes[node];
      GEdge* ed = edges + nd.first_edge;
      // (an edge is one aligned 16-byte load: {prior, w, n, mv | vloss << 16})
      const uint4* raw = reinterpret_cast<const uint4*>(ed);
      int total = 0;
      for (int i = 0; i < nd.n_edges; ++i) { const uint4 r = raw[i]; total += static_cast<int>(r.z) + static_cast<int>(r.w >> 16); }
      // Q + c_puct P sqrt(sum N) / (1 + N), virtual losses counting as visits worth nothing: the very expression, in the
      // very order of operations, of Mcts::select_leaf (csrc/mcts.hpp) and oracle/mcts_oracle.py -- the three trees take
      // the same path bit for bit (tests/test_gpu_parity.py::test_gpu_search_walks_like_the_oracle)
      const float sq = sqrtf(static_cast<float>(total) + 1e-8f);
      int best = -1;
      float best_score = -1e30f;
      if (forced && node == 0)   // forced playouts (Wu 2019; mcts.py:97): a visited root child gets at least sqrt(2 P N) playouts
        for (int i = 0; i < nd.n_edges && best < 0; ++i) {
          const uint4 r = raw[i];
          const int nvi = static_cast<int>(r.z) + static_cast<int>(r.w >> 16);
          if (static_cast<int>(r.z) > 0 &&
              static_cast<float>(nvi) < sqrtf(__fmul_rn(__fmul_rn(2.f, __uint_as_float(r.x)), static_cast<float>(total))))
            best = i;
        }
      const int n_scan = best < 0 ? nd.n_edges : 0;   // (no forced child: the PUCT scan)
      for (int i = 0; i < n_scan; ++i) {
        const uint4 r = raw[i];
        const int nvi = static_cast<int>(r.z) + static_cast<int>(r.w >> 16);
        const float nv = static_cast<float>(nvi);
        const float q = nvi > 0 ? __fdiv_rn(__uint_as_float(r.y), nv) : 0.f;
        const float score = __fadd_rn(q, __fdiv_rn(__fmul_rn(__fmul_rn(p.c_puct, __uint_as_float(r.x)), sq), __fadd_rn(1.f, nv)));
        if (score > best_score) { best_score = score; best = i; }
      }
      if (best < 0) best = 0;
      GEdge& e = ed[best];
      if (e.vloss < 0xFFFF) ++e.vloss;
      Move m = unpack_move(e.mv);
      m.captured = static_cast<int8_t>(Board::type_of(pos.board[m.to]));
      const uint64_t key = pos.hash_after(pv.key(pv.last()), m);
      pos.apply(m);
      ++pos.ply;
      int& ch = child[nd.first_edge + best];
      const bool check = ch >= 0 && nodes[ch].expanded ? nodes[ch].in_check != 0 : pos.in_check(pos.side);
      path_push(pv, pos, key, check);
      if (ch < 0) {
        if (g.n_nodes >= kMaxNodes) {
          // (cannot happen while searches are suspended at 90 %; kept as a safety net that leaves the tree consistent:
          // the virtual loss of the abandoned descent is taken back, and expand_kernel does not count the selection)
          atomicAdd(&p.shared->pool_overflows, 1ull);
          --e.vloss;
          for (int up = node; nodes[up].parent >= 0; up = nodes[up].parent)
            if (edges[nodes[up].parent_edge].vloss > 0) --edges[nodes[up].parent_edge].vloss;
          node = -1;
          break;
        }
        ch = g.n_nodes++;
        GNode& nn = nodes[ch];
        nn.parent = node; nn.parent_edge = nd.first_edge + best; nn.first_edge = 0; nn.n = 0; nn.value = 0.5f; nn.n_edges = 0;
        nn.expanded = nn.terminal = nn.claimed = nn.in_check = 0;
      }
      node = ch;
    }
    g.leaf[s] = node;
    if (node < 0) continue;
    GNode& leaf = nodes[node];
    if (leaf.expanded || leaf.terminal || leaf.claimed) continue;   // nothing to evaluate (or a duplicate of this round)
    leaf.claimed = 1;
    float tv;
    if (leaf_terminal(pos, pv, moves, tv)) { leaf.terminal = 1; leaf.claimed = 0; leaf.value = tv; continue; }
    if (g.n_edges + moves.size() > kMaxEdges) {      // out of room: treat as a draw-ish dead end, never evaluated
      atomicAdd(&p.shared->pool_overflows, 1ull);
      leaf.terminal = 1; leaf.claimed = 0; leaf.value = 0.5f;
      continue;
    }
    // the legal moves become the node's edges right away; their priors arrive with the network's answer
    leaf.first_edge = g.n_edges;
    leaf.n_edges = static_cast<int16_t>(moves.size());
    leaf.in_check = moves.in_check ? 1 : 0;
    g.n_edges += moves.size();
    uint16_t* idx = p.move_idx + slot * kMoveStride;
    for (int i = 0; i < moves.size(); ++i) {
      GEdge& e = edges[leaf.first_edge + i];
      e.prior = 0.f; e.w = 0.f; e.n = 0; e.mv = pack_move(moves.m[i]); e.vloss = 0;
      child[leaf.first_edge + i] = -1;
      idx[i] = static_cast<uint16_t>(policy_index(moves.m[i], pos.side));
    }
    p.move_cnt[slot] = moves.size();
    path_encode(pv, pos, p.bits + slot * kInputPlanes, p.scale + slot * kInputPlanes);
    g.leaf_slot[s] = static_cast<int8_t>(used++);
  }
  g.n_sel = s;
}

// legal_priors_kernel (e55_fp32.cuh) for fixed-stride move lists: position p owns index/priors [p*stride, +count[p]).
// Not launched any more (round 2: the tower kernel's policy tail does this from shared memory); kept as the plain statement
// of what the strided tail computes.
__global__ void __launch_bounds__(256)
legal_priors_strided_kernel(const float* __restrict__ logits, const int32_t* __restrict__ count, const uint16_t* __restrict__ index,
                            float* __restrict__ priors, int n, int width, int stride) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n) return;
  const int k = count[p];
  if (k <= 0) return;
  const float* row = logits + static_cast<size_t>(p) * width;
  const uint16_t* idx = index + static_cast<size_t>(p) * stride;
  float* out = priors + static_cast<size_t>(p) * stride;
  float m = -INFINITY;
  for (int j = lane; j < k; j += 32) m = fmaxf(m, row[idx[j]]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int j = lane; j < k; j += 32) sum += __expf(row[idx[j]] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  for (int j = lane; j < k; j += 32) out[j] = __expf(row[idx[j]] - m) * inv;
}

__global__ void __launch_bounds__(64) expand_kernel(Params p) {
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= p.n_games) return;
  Game& g = p.games[gi];
  GNode* nodes = p.nodes + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxNodes;
  GEdge* edges = p.edges + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxEdges;
  int32_t* child = p.child + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxEdges;
  if (g.n_sel == 0 && (g.need_mate || g.sims_done < g.sim_target)) return;   // the game sat this round out
  const bool root_round = g.sims_done < 0;
  int evaluated = 0;
  // priors and values into the leaves evaluated this round (Mcts::expand_from_priors)
  for (int s = 0; s < g.n_sel; ++s) {
    if (g.leaf_slot[s] < 0) continue;
    const size_t slot = static_cast<size_t>(gi) * kLeaf + g.leaf_slot[s];
    const int k = p.move_cnt[slot];
    ++evaluated;
    GNode& leaf = nodes[g.leaf[s]];
    const float* pr = p.priors + slot * kMoveStride;
    for (int i = 0; i < k; ++i) edges[leaf.first_edge + i].prior = pr[i];
    leaf.value = (p.value[slot] + 1.f) * 0.5f;      // mcts.py:59,103
    leaf.expanded = 1;
    leaf.claimed = 0;
    if (g.leaf[s] == 0 && noise_on(g, p))           // the freshly expanded root (the slot serves as scratch)
      root_noise(g, leaf, edges, p.priors + slot * kMoveStride, p.dirichlet_alpha, p.dirichlet_eps);
  }
  // back every selection up (Mcts::backpropagate)
  int done = 0;
  for (int s = 0; s < g.n_sel; ++s) {
    int node = g.leaf[s];
    if (node < 0) continue;              // a descent abandoned for lack of room is not a simulation
    ++done;
    float v = nodes[node].value;
    ++nodes[node].n;
    while (nodes[node].parent >= 0) {
      v = 1.f - v;
      GEdge& e = edges[nodes[node].parent_edge];
      if (e.vloss > 0) --e.vloss;
      ++e.n;
      e.w += v;
      node = nodes[node].parent;
      ++nodes[node].n;
    }
  }
  if (evaluated) atomicAdd(&p.shared->evals, static_cast<unsigned long long>(evaluated));
  atomicAdd(&p.shared->selections, static_cast<unsigned long long>(g.n_sel));
  if (root_round) { g.sims_done = 0; if (!nodes[0].terminal) return; }   // the root evaluation is not a simulation (mcts.py:55-60)
  else g.sims_done += done;
  if (p.cap_enable && !g.full_search && g.sims_done >= 0 && g.sims_done < g.sim_target) {
    // "immediate" (mcts.py:84-88, set for fast searches): the visits the reused tree already has count
    int visits = 0;
    for (int i = 0; i < nodes[0].n_edges; ++i) visits += edges[nodes[0].first_edge + i].n;
    if (visits >= g.sim_target) g.sims_done = g.sim_target;
  }
  if (g.sims_done < g.sim_target && !nodes[0].terminal) return;
  if (g.need_mate) return;             // a mate found by the host goes first (select_kernel)

  // ---- the search of this move is over: choose, play, look at the game (selfplay.py:63-78,17-31) -----------------
  const GNode& root = nodes[0];
  bool over = root.terminal || root.n_edges == 0;
  int winner = 2, kept_root = -1;
  if (over) {
    // terminal roots can only be positions without moves here (repetition ends a game before it is searched)
    winner = 1 - g.root.side;
  } else {
    const GEdge* ed = edges + root.first_edge;
    int chosen = 0;
    if (g.root.ply < p.sampling_moves) {
      long total = 0;
      for (int i = 0; i < root.n_edges; ++i) total += ed[i].n;
      long pick = total > 0 ? static_cast<long>(next_u64(g.rng) % static_cast<uint64_t>(total)) : 0;
      for (int i = 0; i < root.n_edges; ++i) { pick -= ed[i].n; if (pick < 0) { chosen = i; break; } }
    } else {
      for (int i = 1; i < root.n_edges; ++i)
        if (ed[i].n > ed[chosen].n) chosen = i;
    }
    kept_root = child[root.first_edge + chosen];
    record_ply(g, p, gi, ed[chosen].mv, &root, edges);
    over = play_and_judge(g, p, unpack_move(ed[chosen].mv), winner);
  }
  if (over) {
    finish_game(g, p, gi, nodes, winner);
  } else {
    begin_move(g, p);                    // the kind of search the next move gets decides whether its tree starts afresh
    const bool reuse = p.cap_enable ? !g.full_search : p.reuse_tree != 0;
    if (reuse && kept_root >= 0 && nodes[kept_root].expanded && !nodes[kept_root].terminal) {
      // Mcts::set_root with reuse_tree (mcts.py:50; client.py:32): the chosen move's subtree becomes the tree; its root is
      // expanded already, so the next search starts with simulations at once -- after fresh noise on the root priors
      keep_subtree(g, p, gi, kept_root);
      atomicAdd(&p.shared->kept_nodes, static_cast<unsigned long long>(g.n_nodes));
      g.sims_done = 0;
      if (p.cap_enable && !g.full_search) {   // "immediate": a kept tree with enough visits needs no search at all
        const GNode* kn = p.nodes + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxNodes;
        const GEdge* ke = p.edges + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxEdges;
        int visits = 0;
        for (int i = 0; i < kn[0].n_edges; ++i) visits += ke[kn[0].first_edge + i].n;
        if (visits >= g.sim_target) g.sims_done = g.sim_target;
      }
      if (noise_on(g, p)) {
        GNode* kn = p.nodes + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxNodes;
        root_noise(g, kn[0], p.edges + (static_cast<size_t>(gi) * 2 + g.pool) * kMaxEdges, p.priors + static_cast<size_t>(gi) * kLeaf * kMoveStride,
                   p.dirichlet_alpha, p.dirichlet_eps);
      }
    } else {
      reset_tree(g, nodes);
    }
  }
  request_mate(g, p, gi);              // selfplay.py:35-36: every move starts with the mate search
}

// Validation of the device-side history: every thread plays a random game with the game bookkeeping (game_play), walks
// a random path below it with the descent's bookkeeping (path_push), encodes the position it arrives at
// (path_encode) and judges it (path_repetition).  The host replays the moves on a Position and must get the same planes,
// repetition verdict and number of legal moves (e55_selftest_encode).  Moves that take the previous one back are
// preferred now and then, so that positions do repeat.
constexpr int kSelftestMoves = 192;   // moves per thread: at most 128 of the game + 48 of the path
__global__ void encode_selftest_kernel(Game* games, int n, uint64_t seed, uint32_t* bits, float* scale, uint16_t* moves_out, int32_t* info) {
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= n) return;
  Game& g = games[gi];
  uint64_t rng = (seed + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(gi + 1)) | 1ull;
  for (int i = 0; i < 8; ++i) next_u64(rng);
  GNode scratch_root;
  g.pool = 0;
  start_game(g, &scratch_root);
  MoveList moves;
  uint16_t* out = moves_out + static_cast<size_t>(gi) * kSelftestMoves;
  int n_out = 0;
  auto pick = [&](const MoveList& l, int n_played) {
    if (n_played >= 2 && (next_u64(rng) & 3) != 0) {        // take my previous move back, if that is legal
      const Move prev = unpack_move(out[n_played - 2]);
      if (!prev.is_drop() && !prev.promotes)
        for (int i = 0; i < l.size(); ++i)
          if (l.m[i].from == prev.to && l.m[i].to == prev.from && !l.m[i].promotes) return i;
    }
    return static_cast<int>(next_u64(rng) % static_cast<uint64_t>(l.size()));
  };
  const int game_plies = static_cast<int>(next_u64(rng) % 129), path_plies = static_cast<int>(next_u64(rng) % 49);
  for (int i = 0; i < game_plies; ++i) {
    g.root.generate_moves(moves);
    if (moves.empty()) break;
    const Move m = moves.m[pick(moves, n_out)];
    out[n_out++] = pack_move(m);
    game_play(g, m);
  }
  Board pos = g.root;
  PathView pv;
  pv.g = &g; pv.ply0 = pos.ply; pv.depth = 0;
  for (int i = 0; i < path_plies; ++i) {
    pos.generate_moves(moves);
    if (moves.empty()) break;
    Move m = moves.m[pick(moves, n_out)];
    out[n_out++] = pack_move(m);
    m.captured = static_cast<int8_t>(Board::type_of(pos.board[m.to]));      // as select_kernel plays a move
    const uint64_t key = pos.hash_after(pv.key(pv.last()), m);
    pos.apply(m);
    ++pos.ply;
    path_push(pv, pos, key, pos.in_check(pos.side));
  }
  path_encode(pv, pos, bits + static_cast<size_t>(gi) * kInputPlanes, scale + static_cast<size_t>(gi) * kInputPlanes);
  bool rep, my_check, op_check;
  path_repetition(pv, rep, my_check, op_check);
  pos.generate_moves(moves);
  info[gi * 4 + 0] = n_out;
  info[gi * 4 + 1] = pv.depth;
  info[gi * 4 + 2] = (rep ? 1 : 0) | (my_check ? 2 : 0) | (op_check ? 4 : 0) | (pv.rep_count(pv.last()) << 8);
  info[gi * 4 + 3] = moves.size();
}

// Device perft from the start position (validation of the device build of the engine).
__device__ long long perft_rec(Board& b, int depth) {
  MoveList l;
  b.generate_moves(l);
  if (depth <= 1) return l.size();
  long long total = 0;
  for (int i = 0; i < l.size(); ++i) {
    const Move m = l.m[i];
    b.apply(m); ++b.ply;
    total += perft_rec(b, depth - 1);
    --b.ply; b.revert(m);
  }
  return total;
}
__global__ void perft_kernel(int depth, long long* out) {
  // the root's moves are spread over the threads of the block
  __shared__ long long acc;
  if (threadIdx.x == 0) acc = 0;
  __syncthreads();
  Board b;
  b.clear_board();
  const PieceType top[5] = {ROOK, BISHOP, SILVER, GOLD, KING}, bottom[5] = {KING, GOLD, SILVER, BISHOP, ROOK};
  for (int c = 0; c < 5; ++c) { b.put(sq_of(0, c), top[c], WHITE); b.put(sq_of(4, c), bottom[c], BLACK); }
  b.put(sq_of(1, 4), PAWN, WHITE);
  b.put(sq_of(3, 0), PAWN, BLACK);
  MoveList l;
  b.generate_moves(l);
  long long mine = 0;
  if (depth <= 1) { if (threadIdx.x == 0) mine = l.size(); }
  else
    for (int i = threadIdx.x; i < l.size(); i += blockDim.x) {
      const Move m = l.m[i];
      b.apply(m); ++b.ply;
      mine += perft_rec(b, depth - 1);
      --b.ply; b.revert(m);
    }
  atomicAdd(reinterpret_cast<unsigned long long*>(&acc), static_cast<unsigned long long>(mine));
  __syncthreads();
  if (threadIdx.x == 0) *out = acc;
}

}  // namespace gs
}  // namespace e55

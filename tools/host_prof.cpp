// Host-side profile of the self-play loop without a GPU: links e55_game.cpp against stub network entry points that
// return cheap pseudo-random logits, so tree / move generation / mate search / encoding costs can be measured here.
//   g++ -O3 -std=c++20 -march=x86-64-v3 -pthread -I. tools/host_prof.cpp erweitern_55_b200/csrc/e55_game.cpp -o /tmp/host_prof
// The same build with -O1 -g -fsanitize=address,undefined or -fsanitize=thread is how the engine, the tree and the
// pooled self-play driver were checked for memory errors and data races (clean on 20 000 moves, 6 threads).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../include/e55.h"

static std::atomic<long> g_calls{0}, g_positions{0};

extern "C" {
static float g_table[8192];
static void fill(const uint32_t* bits, int n, float* policy, float* value) {
  if (g_table[1] == 0.f) {
    uint32_t x = 99u;
    for (float& v : g_table) { x = x * 1664525u + 1013904223u; v = static_cast<float>(x >> 8) * (4.f / 16777216.f) - 2.f; }
  }
  for (int i = 0; i < n; ++i) {   // about what the service's copy-out costs: one 6.9 KB memcpy per position
    const uint32_t x = (bits[i * 266] * 2654435761u + bits[i * 266 + 5] * 40503u + bits[i * 266 + 38] * 977u) >> 7;
    memcpy(policy + static_cast<size_t>(i) * 1725, g_table + x % (8192 - 1725), 1725 * sizeof(float));
    value[i] = g_table[x % 8192] * 0.5f;
  }
}
// what the service costs the caller: copy-in of the packed input and the move lists, copy-out of the priors
int e55_service_predict_packed_moves(e55_ctx*, const uint32_t* bits, const float* scale, int n, const int32_t* off,
                                     const uint16_t* idx, float* priors, float* value) {
  g_calls++; g_positions += n;
  static thread_local uint32_t stage_b[266 * 256];
  static thread_local float stage_s[266 * 256];
  static thread_local uint16_t stage_i[256 * 256];
  memcpy(stage_b, bits, sizeof(uint32_t) * 266 * n); memcpy(stage_s, scale, sizeof(float) * 266 * n);
  memcpy(stage_i, idx, sizeof(uint16_t) * off[n]);
  if (g_table[1] == 0.f) { float dummy[1725]; float v; fill(bits, 1, dummy, &v); }
  for (int i = 0; i < n; ++i) {
    const uint32_t x = (bits[i * 266] * 2654435761u + bits[i * 266 + 5] * 40503u + bits[i * 266 + 38] * 977u) >> 7;
    const int k = off[i + 1] - off[i];
    float sum = 0.f;
    for (int j = 0; j < k; ++j) { priors[off[i] + j] = 2.1f + g_table[(x + stage_i[off[i] + j]) % 8192]; sum += priors[off[i] + j]; }
    for (int j = 0; j < k; ++j) priors[off[i] + j] /= sum;
    value[i] = g_table[x % 8192] * 0.5f;
  }
  return 0;
}
int e55_predict_packed_moves(e55_ctx* c, const uint32_t* bits, const float* s, int n, const int32_t* off, const uint16_t* idx,
                             float* priors, float* value) {
  return e55_service_predict_packed_moves(c, bits, s, n, off, idx, priors, value);
}
// asynchronous form: answered at once, handed over at fetch
struct Pending { std::vector<float> priors, value; };
static std::mutex g_mu;
static std::vector<Pending*> g_pending;
int e55_service_submit_packed_moves(e55_ctx* c, const uint32_t* bits, const float* s, int n, const int32_t* off, const uint16_t* idx,
                                    int, e55_ticket* t) {
  Pending* p = new Pending{std::vector<float>(off[n]), std::vector<float>(n)};
  e55_service_predict_packed_moves(c, bits, s, n, off, idx, p->priors.data(), p->value.data());
  std::lock_guard<std::mutex> lk(g_mu);
  int slot = -1;
  for (size_t i = 0; i < g_pending.size(); ++i) if (!g_pending[i]) { slot = static_cast<int>(i); break; }
  if (slot < 0) { slot = static_cast<int>(g_pending.size()); g_pending.push_back(nullptr); }
  g_pending[slot] = p;
  t->batch = slot; t->n = n; t->total = off[n];
  return 0;
}
int e55_service_fetch(e55_ctx*, const e55_ticket* t, float* priors, float* value, int) {
  Pending* p;
  { std::lock_guard<std::mutex> lk(g_mu); p = g_pending[t->batch]; g_pending[t->batch] = nullptr; }
  memcpy(priors, p->priors.data(), sizeof(float) * p->priors.size());
  memcpy(value, p->value.data(), sizeof(float) * p->value.size());
  delete p;
  return 0;
}
int e55_get_arch(const e55_ctx*, int* b, int* f, int* ip) { *b = 7; *f = 128; *ip = 266; return 0; }
int e55_service_stats(e55_ctx*, int64_t* b, int64_t* p) { *b = g_calls; *p = g_positions; return 0; }
}

// `host_prof fuzz <games> [simulations]`: the position / tree API the Python classes bind (e55_pos_*, e55_mcts_*), driven
// the way mcts.py drives minishogilib -- immediate mode, random policies, tree reuse -- with every query in between.
// Meant for the sanitizer builds; returns non-zero on any inconsistency it sees itself.
static int fuzz(int games, int sims) {
  uint64_t rng = 0x55aa1234beefull;
  auto rnd = [&] { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  std::vector<float> policy(1725);
  std::vector<uint32_t> bits(266);
  std::vector<float> scale(266);
  char buf[16384], mv[16], sf[4096];
  int counts[700];
  long plies = 0, mates = 0, reps = 0;
  e55_mcts* tree = e55_mcts_create(0.05);
  e55_mcts_configure(tree, 1.25f, 0.15f, 0.25f, 7);
  for (int g = 0; g < games; ++g) {
    e55_pos* pos = e55_pos_create();
    e55_mcts_clear(tree);
    for (int ply = 0; ply < 200; ++ply) {
      const int n_moves = e55_pos_generate_moves(pos, buf, sizeof(buf));
      int rep[3];
      e55_pos_is_repetition(pos, rep);
      if (n_moves <= 0 || rep[0]) { reps += rep[0]; break; }
      if (e55_pos_sfen(pos, sf, sizeof(sf)) != 0) return 10;
      e55_pos* again = e55_pos_create();                       // SFEN -> the same position
      if (e55_pos_set_sfen(again, sf) != 0) return 11;
      char sf2[4096];
      e55_pos_sfen(again, sf2, sizeof(sf2));
      // (board, side and hands; the move number restarts: set_sfen counts plies from the position it is given)
      if (strncmp(sf, sf2, strrchr(sf, ' ') - sf) != 0) {
        fprintf(stderr, "sfen round trip: '%s' -> '%s'\n", sf, sf2);
        return 12;
      }
      e55_pos_destroy(again);
      e55_pos_encode_packed(pos, bits.data(), scale.data());
      e55_pos_print(pos, buf, sizeof(buf));
      if (e55_pos_solve_checkmate_dfs(pos, 3, mv, sizeof(mv)) == 1) {
        ++mates;
        if (e55_pos_do_move(pos, mv) != 0) return 13;
        continue;
      }
      const int root = e55_mcts_set_root(tree, pos, (g & 1));
      if (root < 0) return 14;
      for (float& v : policy) v = static_cast<float>(rnd() % 2048) / 512.f - 2.f;
      if (!e55_mcts_expanded(tree, root)) e55_mcts_evaluate(tree, root, pos, policy.data(), 0.5f);
      e55_mcts_add_noise(tree, root);
      for (int s = 0; s < sims; ++s) {
        e55_pos* leaf_pos = e55_pos_copy(pos);
        const int leaf = e55_mcts_select_leaf(tree, root, leaf_pos, s % 3 == 0);
        if (leaf < 0) return 15;
        if (!e55_mcts_expanded(tree, leaf)) {
          for (int k = 0; k < 64; ++k) policy[rnd() % 1725] = static_cast<float>(rnd() % 2048) / 512.f - 2.f;
          e55_mcts_evaluate(tree, leaf, leaf_pos, policy.data(), static_cast<float>(rnd() % 1000) / 999.f);
        }
        e55_mcts_backpropagate(tree, leaf);
        e55_pos_destroy(leaf_pos);
      }
      float q;
      int sum_n;
      e55_mcts_info(tree, root, buf, sizeof(buf), &q);
      e55_mcts_describe(tree, root, buf, sizeof(buf));
      e55_mcts_visualize(tree, root, 20, buf, sizeof(buf));
      const int entries = e55_mcts_dump(tree, root, ply & 1, 1, &sum_n, &q, buf, sizeof(buf), counts, 700);
      if (entries <= 0 || e55_mcts_get_playouts(tree, root, 1) < sims || e55_mcts_get_nodes(tree) <= 0 || e55_mcts_get_usage(tree) < 0) return 16;
      int ok;
      switch (rnd() % 3) {
        case 0: ok = e55_mcts_best_move(tree, root, mv, sizeof(mv)); break;
        case 1: ok = e55_mcts_softmax_sample(tree, root, 1.0f, mv, sizeof(mv)); break;
        default: ok = e55_mcts_softmax_sample_among_top_moves(tree, root, 0.1f, 1.0f, mv, sizeof(mv)); break;
      }
      if (ok != 1 || e55_pos_policy_index(pos, mv) < 0 || e55_pos_do_move(pos, mv) != 0) return 17;
      if (e55_pos_do_move(pos, "9z9z") == 0) return 18;        // nonsense must be refused
      ++plies;
    }
    e55_pos_destroy(pos);
  }
  e55_mcts_destroy(tree);
  printf("fuzz: %d games, %ld searched plies, %ld mate moves, %ld repetition ends: ok\n", games, plies, mates, reps);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && strcmp(argv[1], "fuzz") == 0) {
    const int r = fuzz(argc > 2 ? atoi(argv[2]) : 20, argc > 3 ? atoi(argv[3]) : 64);
    if (r) fprintf(stderr, "fuzz: check %d failed\n", r);
    return r;
  }
  const int threads = argc > 1 ? atoi(argv[1]) : 1;
  const long moves = argc > 2 ? atol(argv[2]) : 200;
  const int mate = argc > 3 ? atoi(argv[3]) : 7;
  const int sims = argc > 4 ? atoi(argv[4]) : 800;
  { uint32_t b[266] = {}; float pol[1725], v; fill(b, 1, pol, &v); }   // the stub's logit table, before any worker thread looks at it
  e55_selfplay_config cfg;
  e55_selfplay_default_config(&cfg);
  cfg.threads = threads; cfg.total_moves = moves; cfg.mate_depth = mate; cfg.simulations = sims;
  cfg.games_per_thread = argc > 5 ? atoi(argv[5]) : 1;
  e55_selfplay_stats st;
  const int rc = e55_selfplay_run_ex(reinterpret_cast<e55_ctx*>(0x1), &cfg, &st);
  printf("rc %d: %ld moves, %ld evals, %ld games (%.1f plies) in %.3f s -> %.1f moves/s, %.2f us/eval, %.2f ms/move/thread\n", rc,
         st.moves, st.evals, st.games, st.mean_game_plies, st.seconds, st.moves / st.seconds,
         1e6 * st.seconds * threads / (st.evals > 0 ? st.evals : 1), 1e3 * st.seconds * threads / st.moves);
  return rc;
}
